/* wiggletools_b200 -- the reference's own operator API for the merge+reduce hot
 * path, re-exported by libwiggletools_b200.so with the reference's struct
 * layouts, so that the UNMODIFIED front end (src/commandParser.c, the writers,
 * the readers) links against the GPU engine instead of multiplexer.o,
 * multiSet.o, reducers.o and setComparisons.o.
 *
 * Every prototype below replaces the reference definition cited next to it
 * (paths relative to /root/reference/src).  Types mirror
 * wiggleIterator.h:21-35, multiplexer.h:21-36 and multiSet.h:20-30 field for
 * field (x86-64 offsets are static_asserted in csrc/dropin.cpp) because callers
 * read the fields directly.
 *
 * Behavioural contract kept from the reference:
 *  - constructors pop once, so the first record is readable on return
 *    (wiggleIterator.c:20-34);
 *  - user/arity errors print a message and exit(1) (reducers.c:111-114,
 *    setComparisons.c:125-128,374-377);
 *  - one interned, stable char* per chromosome is handed out.
 * Deliberate differences: a GPU-backed Multiplexer is batch:
 * its rows (values[]/inplay[]) are not materialised, so popMultiplexer() on it
 * (mwrite, paste) exits with a message; FTestReduction exits (the reference's
 * is defective, setComparisons.c:183,199).
 */
#ifndef WIGGLETOOLS_B200_DROPIN_H
#define WIGGLETOOLS_B200_DROPIN_H

#ifdef __cplusplus
extern "C" {
#endif

typedef char wt_bool; /* the reference's `bool` macro, wiggletools.h:18-22 */

typedef struct wiggleIterator_st WiggleIterator;
typedef struct multiplexer_st Multiplexer;
typedef struct multiset_st Multiset;

struct wiggleIterator_st {          /* wiggleIterator.h:21-35 */
	char *chrom;
	int start;
	int finish;
	double value;
	void *valuePtr;
	wt_bool done;
	int strand;
	void *data;
	void (*pop)(WiggleIterator *);
	void (*seek)(WiggleIterator *, const char *, int, int);
	wt_bool overlaps;
	double default_value;
	WiggleIterator *append;
};

struct multiplexer_st {             /* multiplexer.h:21-36 */
	char *chrom;
	int start;
	int finish;
	double *values;
	double *default_values;
	int count, inplay_count;
	wt_bool *inplay;
	WiggleIterator **iters;
	wt_bool done;
	wt_bool strict;
	void (*pop)(Multiplexer *);
	void (*seek)(Multiplexer *, const char *, int, int);
	void *starts, *finishes;        /* FibHeap* in the reference; unused here */
	void *data;
};

struct multiset_st {                /* multiSet.h:20-30 */
	char *chrom;
	int start;
	int finish;
	double **values;
	int count, inplay_count;
	wt_bool *inplay;
	Multiplexer **multis;
	wt_bool done;
	void *starts, *finishes;
};

/* -- multiplexer.c --------------------------------------------------------- */
Multiplexer *newMultiplexer(WiggleIterator **iters, int count, wt_bool strict);            /* :157-169 */
Multiplexer *newCoreMultiplexer(void *data, int count, void (*pop)(Multiplexer *),
                                void (*seek)(Multiplexer *, const char *, int, int));      /* :143-155 */
void popMultiplexer(Multiplexer *multi);                                                   /* :22-25 */
void runMultiplexer(Multiplexer *multi);                                                   /* :27-30 */
void seekMultiplexer(Multiplexer *multi, const char *chrom, int start, int finish);        /* :32-35 */

/* -- multiSet.c ------------------------------------------------------------ */
Multiset *newMultiset(Multiplexer **multis, int count);                                    /* :115-128 */
void popMultiset(Multiset *multi);                                                         /* :80-101 */
void seekMultiset(Multiset *multi, const char *chrom, int start, int finish);              /* :103-113 */

/* -- reducers.c ------------------------------------------------------------ */
WiggleIterator *SelectReduction(Multiplexer *, int);        /* :40-72 (generic, any multiplexer) */
WiggleIterator *FillInReduction(Multiplexer *, wt_bool);    /* :78-119 */
WiggleIterator *MaxReduction(Multiplexer *);                /* :125-186 */
WiggleIterator *MinReduction(Multiplexer *);                /* :192-253 */
WiggleIterator *SumReduction(Multiplexer *);                /* :259-307 */
WiggleIterator *ProductReduction(Multiplexer *);            /* :313-361 */
WiggleIterator *MeanReduction(Multiplexer *);               /* :367-422 */
WiggleIterator *VarianceReduction(Multiplexer *);           /* :428-505 */
WiggleIterator *StdDevReduction(Multiplexer *);             /* :511-589 */
WiggleIterator *EntropyReduction(Multiplexer *);            /* :640-666 */
WiggleIterator *CVReduction(Multiplexer *);                 /* :672-750 */
WiggleIterator *MedianReduction(Multiplexer *);             /* :780-834 */

/* -- setComparisons.c ------------------------------------------------------ */
WiggleIterator *TTestReduction(Multiset *);                 /* :35-131 */
WiggleIterator *FTestReduction(Multiset *);                 /* :152-235: exits, see above */
WiggleIterator *MWUReduction(Multiset *);                   /* :269-390 */

/* -- statistics.c (the two whole-genome scalars on the path) --------------- */
WiggleIterator *AUCIntegrator(WiggleIterator *);            /* :103-127 */
WiggleIterator *PearsonIntegrator(Multiplexer *);           /* :414-465 */
WiggleIterator *NDPearsonIntegrator(Multiset *);            /* :490-556 (rank = 2 quirk of :551 kept) */

/* -- bigWiggleReader.c: BigWig source on the library's batch decoder ---------- */
WiggleIterator *BigWiggleReader(char *filename, wt_bool holdFire);                         /* :147-151 */

/* -- wiggleIterator.c / unaryOps.c: the few core entry points the path calls;
 * defined here too so the library is self-contained (a host program that
 * already defines them -- the reference CLI -- keeps its own). ---------------- */
WiggleIterator *newWiggleIterator(void *data, void (*pop)(WiggleIterator *),
                                  void (*seek)(WiggleIterator *, const char *, int, int),
                                  double default_value, wt_bool overlapping);              /* wiggleIterator.c:20-34 */
void pop(WiggleIterator *);                                                                /* wiggleIterator.c:57-60 */
void seek(WiggleIterator *, const char *, int, int);                                       /* wiggleIterator.c:67-70 */
void runWiggleIterator(WiggleIterator *);                                                  /* wiggleIterator.c:62-65 */
void destroyWiggleIterator(WiggleIterator *);                                              /* wiggleIterator.c:52-55 */
WiggleIterator *UnionWiggleIterator(WiggleIterator *);                                     /* unaryOps.c:60-91 */
WiggleIterator *NonOverlappingWiggleIterator(WiggleIterator *);                            /* unaryOps.c:98-103 */

/* -- wigWriter.c: the default text sink on the device -------------------------------
 * What toFile() / toStdout() (src/wigWriter.c:278-290) do for an iterator returned by
 * one of the reductions above: compress (write only), 10000-record writer blocks,
 * fixedStep / bedGraph lines with "%lf" values -- computed per chromosome on the
 * device right after the reduction, so that segments never cross PCIe and the host
 * only fwrites text.  wb200_sink_accepts() says whether `wi` is such an iterator in
 * its freshly constructed state; anything else stays with the reference's writer. */
#include <stdio.h>
/* -- unaryOps.c: the element-wise operators (scale, offset, ln / log, exp, pow, abs, default;
 * src/unaryOps.c:210-229,650-960) as one constructor.  Popped on the host the iterator
 * behaves as the reference's; as input of one of the reductions above its maps run on
 * the device (wb200_set_input_maps).  op: WB200_MAP_* of wiggletools_b200.h, 0 for
 * `default`; arg: the CLI-level argument (log base <= 0 means ln). */
WiggleIterator *wb200_mapped_iterator(WiggleIterator *i, int op, double arg);

int wb200_sink_accepts(WiggleIterator *wi);
void wb200_sink_run(WiggleIterator *wi, FILE *out, wt_bool bedGraph);

/* -- apply.c: ApplyMultiplexer (src/apply.c:340-353) for a dataset that is a fresh reduction of this
 * library: per run of regions on one chromosome the reduction is sought to the run's span, its
 * segments stay in HBM and the integrators (WB200_STAT_* of wiggletools_b200.h, one per output
 * column, standing for the statistics[] constructors of src/statistics.c:62-326) run there as
 * segmented reductions.  Same row protocol as the reference's (chrom/start/finish = the region,
 * values[i] = statistic i, first row on display on return).  NULL when the dataset is anything
 * else: the caller then keeps the reference's ApplyMultiplexer. */
Multiplexer *wb200_apply_multiplexer(WiggleIterator *regions, const int *stat_codes, int count, WiggleIterator *dataset, wt_bool strict);

#ifdef __cplusplus
}
#endif
#endif
