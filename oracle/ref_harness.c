/* TEST INFRASTRUCTURE ONLY (oracle/).  Nothing under wiggletools_b200/ may
 * include, link or call this file.
 *
 * Harness that drives the UNMODIFIED reference hot path (compiled from
 * /root/reference/src by oracle/Makefile into oracle/_ref/libwtref.so) on
 * in-memory tracks, so that the C restatement (oracle/wt_oracle.c) and the
 * CUDA engine can be checked against the reference's own code.
 *
 * It contains three things, none of them copied from the reference:
 *  1. link stubs for the reader constructors whose libraries (libBigWig,
 *     htslib) are absent: BigWiggleReader, BigBedReader, BamReader, BcfReader,
 *     libBigWigInit  (declared at /root/reference/src/wiggletools.h:36-42,143);
 *  2. an in-memory WiggleIterator source (arrays -> pop), built with the
 *     reference's own newWiggleIterator (src/wiggleIterator.c:20-34);
 *  3. C entry points (wtref_*) that build newMultiplexer / newMultiset and the
 *     reducers exactly as src/commandParser.c:500-569,635-651,683-694 does and
 *     drain them with the consumer loop of src/wiggleIterator.c:62-65.
 */
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#ifdef WB200_DROPIN
/* Same harness, compiled against the GPU library's drop-in API instead of the
 * reference objects (oracle/_ref/libwtdropin.so): the parity tests then drive
 * newMultiplexer / MeanReduction / ... of libwiggletools_b200.so exactly the
 * way they drive the reference. */
#include "wiggletools_dropin.h"
#define bool char
#define true 1
#define false 0
#else
#include "wiggleIterator.h"   /* reference header: struct wiggleIterator_st */
#include "multiplexer.h"
#include "multiSet.h"
#endif

/* ---- 1. stubs for absent third-party readers --------------------------- */
#ifndef WB200_DROPIN
static void wtref_absent(const char *what) {
	fprintf(stderr, "oracle/_ref: %s needs a library that is not in this image\n", what);
	exit(1);
}
#ifndef WTREF_WITH_BIGWIG   /* oracle/_ref/libwtrefbw.so links the reference's own bigWiggleReader.c over oracle/bigwig_shim.c */
WiggleIterator *BigWiggleReader(char *f, bool h) { (void)f; (void)h; wtref_absent("BigWiggleReader"); return NULL; }
#endif
WiggleIterator *BigBedReader(char *f, bool h) { (void)f; (void)h; wtref_absent("BigBedReader"); return NULL; }
WiggleIterator *BamReader(char *f, bool h, bool r) { (void)f; (void)h; (void)r; wtref_absent("BamReader"); return NULL; }
WiggleIterator *BcfReader(char *f, bool h) { (void)f; (void)h; wtref_absent("BcfReader"); return NULL; }
#ifndef WTREF_WITH_BIGWIG
void libBigWigInit(int n) { (void)n; }
#endif
#endif

/* ---- 2. in-memory source ------------------------------------------------ */
typedef struct {
	const int32_t *chrom;   /* index into names[] per interval */
	const int32_t *start;   /* 1-based, inclusive */
	const int32_t *end;     /* exclusive */
	const double *value;
	int64_t n;
	double default_value;
} wt_track;

typedef struct {
	int32_t *chrom, *start, *end;
	double *value;
	int64_t cap;
} wt_out;

typedef struct {
	const wt_track *t;
	char **names;
	int64_t i;
} MemSource;

static void MemSourcePop(WiggleIterator *wi) {
	MemSource *m = (MemSource *) wi->data;
	if (m->i >= m->t->n) {
		wi->done = true;
		return;
	}
	wi->chrom = m->names[m->t->chrom[m->i]];
	wi->start = m->t->start[m->i];
	wi->finish = m->t->end[m->i];
	wi->value = m->t->value[m->i];
	m->i++;
}

static void MemSourceSeek(WiggleIterator *wi, const char *chrom, int start, int finish) {
	/* Region restriction in the style of the reference's binary readers
	 * (src/bigWiggleReader.c:125-145): skip records ending at or before start,
	 * clip the first start. */
	MemSource *m = (MemSource *) wi->data;
	(void) finish;
	m->i = 0;
	while (m->i < m->t->n &&
	       (strcmp(m->names[m->t->chrom[m->i]], chrom) < 0 ||
	        (strcmp(m->names[m->t->chrom[m->i]], chrom) == 0 && m->t->end[m->i] <= start)))
		m->i++;
	wi->done = false;
	MemSourcePop(wi);
	if (!wi->done && strcmp(wi->chrom, chrom) == 0 && wi->start < start)
		wi->start = start;
}

static WiggleIterator *MemSourceIterator(const wt_track *t, char **names) {
	MemSource *m = (MemSource *) calloc(1, sizeof(MemSource));
	m->t = t;
	m->names = names;
	return newWiggleIterator(m, &MemSourcePop, &MemSourceSeek, t->default_value, false);
}

/* ---- 3. entry points ---------------------------------------------------- */
enum {
	WT_SUM = 0, WT_PRODUCT = 1, WT_MEAN = 2, WT_VAR = 3, WT_STDDEV = 4,
	WT_ENTROPY = 5, WT_CV = 6, WT_MIN = 7, WT_MAX = 8, WT_MEDIAN = 9,
	WT_FILLIN = 10, WT_TRIMFILL = 11,
	WT_TTEST = 16, WT_MWU = 17
};

static int chrom_index(char **names, int nnames, const char *c) {
	for (int i = 0; i < nnames; i++)
		if (names[i] == c) return i;
	for (int i = 0; i < nnames; i++)
		if (!strcmp(names[i], c)) return i;
	return -1;
}

static int64_t drain(WiggleIterator *wi, char **names, int nnames, wt_out *out) {
	int64_t n = 0;
	while (!wi->done) {
		if (out && n < out->cap) {
			out->chrom[n] = chrom_index(names, nnames, wi->chrom);
			out->start[n] = wi->start;
			out->end[n] = wi->finish;
			out->value[n] = wi->value;
		}
		n++;
		pop(wi);
	}
	return n;
}

static Multiplexer *build_multiplexer(int n, const wt_track *tracks, char **names, int strict) {
	WiggleIterator **iters = (WiggleIterator **) calloc(n, sizeof(WiggleIterator *));
	for (int i = 0; i < n; i++)
		iters[i] = MemSourceIterator(&tracks[i], names);
	Multiplexer *multi = newMultiplexer(iters, n, strict ? true : false);
	free(iters); /* newMultiplexer copies the array (src/multiplexer.c:160) */
	return multi;
}

static WiggleIterator *build_reduction(int op, Multiplexer *multi) {
	switch (op) {
	case WT_SUM: return SumReduction(multi);
	case WT_PRODUCT: return ProductReduction(multi);
	case WT_MEAN: return MeanReduction(multi);
	case WT_VAR: return VarianceReduction(multi);
	case WT_STDDEV: return StdDevReduction(multi);
	case WT_ENTROPY: return EntropyReduction(multi);
	case WT_CV: return CVReduction(multi);
	case WT_MIN: return MinReduction(multi);
	case WT_MAX: return MaxReduction(multi);
	case WT_MEDIAN: return MedianReduction(multi);
	case WT_FILLIN: return FillInReduction(multi, false);
	case WT_TRIMFILL: return FillInReduction(multi, true);
	}
	fprintf(stderr, "wtref: unknown op %d\n", op);
	exit(1);
}

/* `<op> [strict] t0 t1 ...` -> segments.  Returns the number of records the
 * reference produced (records beyond out->cap are counted, not stored).
 * *default_out receives the reducer's derived default_value. */
int64_t wtref_reduce(int op, int ntracks, const wt_track *tracks, int strict,
                     int nnames, char **names, wt_out *out, double *default_out) {
	Multiplexer *multi = build_multiplexer(ntracks, tracks, names, strict);
	WiggleIterator *wi = build_reduction(op, multi);
	if (default_out) *default_out = wi->default_value;
	return drain(wi, names, nnames, out);
}

/* `ttest A... : B...` / `wilcoxon A... : B...` */
int64_t wtref_setcompare(int op, int n1, const wt_track *t1, int n2, const wt_track *t2,
                         int nnames, char **names, wt_out *out) {
	Multiplexer **multis = (Multiplexer **) calloc(2, sizeof(Multiplexer *));
	multis[0] = build_multiplexer(n1, t1, names, 0);
	multis[1] = build_multiplexer(n2, t2, names, 0);
	Multiset *ms = newMultiset(multis, 2);
	WiggleIterator *wi = (op == WT_TTEST) ? TTestReduction(ms) : MWUReduction(ms);
	return drain(wi, names, nnames, out);
}

/* `AUC <op> t0 t1 ...` : whole-genome scalar of a reduction. */
double wtref_auc(int op, int ntracks, const wt_track *tracks, int strict, int nnames, char **names) {
	(void) nnames;
	Multiplexer *multi = build_multiplexer(ntracks, tracks, names, strict);
	WiggleIterator *wi = AUCIntegrator(build_reduction(op, multi));
	runWiggleIterator(wi);
	return *(double *) wi->data; /* result is the first field (src/statistics.c:26-31) */
}

/* `pearson X Y` */
double wtref_pearson(const wt_track *x, const wt_track *y, int nnames, char **names) {
	(void) nnames;
	wt_track two[2] = { *x, *y };
	Multiplexer *multi = build_multiplexer(2, two, names, 0);
	WiggleIterator *wi = PearsonIntegrator(multi);
	runWiggleIterator(wi);
	return *(double *) wi->data;
}

/* `ndpearson A1 A2 .. : B1 B2 ..` */
double wtref_ndpearson(int n, const wt_track *a, const wt_track *b, int nnames, char **names) {
	(void) nnames;
	Multiplexer **multis = (Multiplexer **) calloc(2, sizeof(Multiplexer *));
	multis[0] = build_multiplexer(n, a, names, 0);
	multis[1] = build_multiplexer(n, b, names, 0);
	WiggleIterator *wi = NDPearsonIntegrator(newMultiset(multis, 2));
	runWiggleIterator(wi);
	return *(double *) wi->data;
}

/* Raw multiplexer rows (what mWigWriter sees): values[count] + inplay[count]
 * per segment; row_values/row_inplay hold cap*count entries. */
int64_t wtref_rows(int ntracks, const wt_track *tracks, int strict, int nnames, char **names,
                   wt_out *out, double *row_values, char *row_inplay) {
	Multiplexer *multi = build_multiplexer(ntracks, tracks, names, strict);
	int64_t n = 0;
	while (!multi->done) {
		if (n < out->cap) {
			out->chrom[n] = chrom_index(names, nnames, multi->chrom);
			out->start[n] = multi->start;
			out->end[n] = multi->finish;
			out->value[n] = 0;
			for (int i = 0; i < ntracks; i++) {
				row_values[n * ntracks + i] = multi->values[i];
				row_inplay[n * ntracks + i] = multi->inplay[i];
			}
		}
		n++;
		popMultiplexer(multi);
	}
	return n;
}

#ifndef WB200_DROPIN
/* Writer-side compress (src/unaryOps.c:235-263) applied to a reduction. */
int64_t wtref_reduce_compressed(int op, int ntracks, const wt_track *tracks, int strict,
                                int nnames, char **names, wt_out *out) {
	Multiplexer *multi = build_multiplexer(ntracks, tracks, names, strict);
	WiggleIterator *wi = CompressionWiggleIterator(build_reduction(op, multi));
	return drain(wi, names, nnames, out);
}

/* The reference's own sink on an in-memory segment list: TeeWiggleIterator +
 * runWiggleIterator, i.e. what toFile() does (src/wigWriter.c:261-284). */
WiggleIterator *TeeWiggleIterator(WiggleIterator *, FILE *, bool, bool);
int64_t wtref_write(int64_t n, const int32_t *chrom, const int32_t *start, const int32_t *end,
                    const double *value, int nnames, char **names, int bedgraph, const char *path) {
	(void) nnames;
	FILE *f = fopen(path, "w");
	if (!f) return -1;
	wt_track t = { chrom, start, end, value, n, 0.0 };
	WiggleIterator *src = MemSourceIterator(&t, names);
	runWiggleIterator(TeeWiggleIterator(src, f, bedgraph ? true : false, false));
	fclose(f);
	return n;
}

/* The reference's own unary operators chained over an in-memory track (codes as in
 * wto_map of wt_oracle.c), built the way src/commandParser.c:115-211 builds them. */
WiggleIterator *ScaleWiggleIterator(WiggleIterator *, double);
WiggleIterator *ShiftWiggleIterator(WiggleIterator *, double);
WiggleIterator *NaturalLogWiggleIterator(WiggleIterator *);
WiggleIterator *LogWiggleIterator(WiggleIterator *, double);
WiggleIterator *NaturalExpWiggleIterator(WiggleIterator *);
WiggleIterator *PowerWiggleIterator(WiggleIterator *, double);
WiggleIterator *AbsWiggleIterator(WiggleIterator *);
WiggleIterator *DefaultValueWiggleIterator(WiggleIterator *, double);
int64_t wtref_map(const wt_track *t, int nmaps, const int *codes, const double *args, int nnames, char **names,
                  wt_out *out, double *default_out) {
	WiggleIterator *wi = MemSourceIterator(t, names);
	for (int c = 0; c < nmaps; c++) {
		switch (codes[c]) {
		case 1: wi = ScaleWiggleIterator(wi, args[c]); break;
		case 2: wi = ShiftWiggleIterator(wi, args[c]); break;
		case 3: wi = args[c] > 0 ? LogWiggleIterator(wi, args[c]) : NaturalLogWiggleIterator(wi); break;
		case 4: wi = NaturalExpWiggleIterator(wi); break;
		case 5: wi = PowerWiggleIterator(wi, args[c]); break;
		case 6: wi = AbsWiggleIterator(wi); break;
		case 7: wi = DefaultValueWiggleIterator(wi, args[c]); break;
		}
	}
	if (default_out) *default_out = wi->default_value;
	return drain(wi, names, nnames, out);
}

/* The reference's own `apply`: ApplyMultiplexer(regions, statistics, count, dataset, strict)
 * (src/apply.c:330-344) over an in-memory dataset and in-memory regions; one row of `count`
 * statistics per region (codes as in wto_apply of wt_oracle.c). */
WiggleIterator *MeanIntegrator(WiggleIterator *);
WiggleIterator *SpanIntegrator(WiggleIterator *);
WiggleIterator *MaxIntegrator(WiggleIterator *);
WiggleIterator *MinIntegrator(WiggleIterator *);
WiggleIterator *VarianceIntegrator(WiggleIterator *);
WiggleIterator *StandardDeviationIntegrator(WiggleIterator *);
WiggleIterator *CoefficientOfVariationIntegrator(WiggleIterator *);
Multiplexer *ApplyMultiplexer(WiggleIterator *, WiggleIterator *(**)(WiggleIterator *), int, WiggleIterator *, bool);
int64_t wtref_apply(const wt_track *dataset, const wt_track *regions, int nstats, const int *codes, int strict,
                    int nnames, char **names, double *out) {
	(void) nnames;
	WiggleIterator *(**stats)(WiggleIterator *) = calloc(nstats, sizeof(*stats));
	for (int k = 0; k < nstats; k++) {
		switch (codes[k]) {
		case 1: stats[k] = &AUCIntegrator; break;
		case 2: stats[k] = &MeanIntegrator; break;
		case 3: stats[k] = &VarianceIntegrator; break;
		case 4: stats[k] = &StandardDeviationIntegrator; break;
		case 5: stats[k] = &CoefficientOfVariationIntegrator; break;
		case 6: stats[k] = &MaxIntegrator; break;
		case 7: stats[k] = &MinIntegrator; break;
		case 8: stats[k] = &SpanIntegrator; break;
		}
	}
	Multiplexer *apply = ApplyMultiplexer(MemSourceIterator(regions, names), stats, nstats, MemSourceIterator(dataset, names),
	                                      strict ? true : false);
	int64_t r = 0;
	while (!apply->done) {
		for (int k = 0; k < nstats; k++) out[r * nstats + k] = apply->values[k];
		r++;
		popMultiplexer(apply);
	}
	return r;
}

#endif /* !WB200_DROPIN */

#ifdef WTREF_WITH_BIGWIG
/* A BigWig file through the reference's own reader (src/bigWiggleReader.c: 10000-bp stretch
 * queries, boxing, producer thread) -- records as the reference's consumers see them. */
WiggleIterator *BigWiggleReader(char *, bool);
int64_t wtref_read_bigwig(char *path, int nnames, char **names, wt_out *out) {
	WiggleIterator *wi = BigWiggleReader(path, false);
	return drain(wi, names, nnames, out);
}
#endif

#ifndef WB200_DROPIN
double wtref_tdist_Q(double t, double nu);
double gsl_cdf_tdist_Q(double, double);
double wtref_tdist_Q(double t, double nu) { return gsl_cdf_tdist_Q(t, nu); }

/* The reference CLI (src/wiggletools.c:22-34 minus libBigWigInit): lets tests
 * run `wiggletools <program>` in-process on the reference's text fixtures. */
int wtref_cli(int argc, char **argv) {
	rollYourOwn(argc, argv);
	return 0;
}
#endif /* !WB200_DROPIN */
