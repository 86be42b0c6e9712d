/* wiggletools_b200 -- C ABI of the B200 interval-stream reduction engine.
 *
 * This is the drop-in boundary for ONE hot path of Ensembl/WiggleTools:
 *   N sorted (chrom,start,end,value) tracks -> union-breakpoint segments
 *   (Multiplexer / Multiset) -> per-segment N-wide statistic -> optional
 *   whole-genome scalars (AUC, Pearson).
 *
 * Two layers are exported by libwiggletools_b200.so:
 *
 *  (1) the batch engine (wb200_*): plain pointers and sizes, one call per
 *      chromosome shard.  This is what a binding (ctypes / cgo / JNI) uses.
 *  (2) the reference's own operator API for the path (newMultiplexer,
 *      popMultiplexer, newMultiset, SumReduction ... MWUReduction,
 *      AUCIntegrator, PearsonIntegrator) with the reference's struct layouts,
 *      declared in include/wiggletools_dropin.h, so that the unmodified
 *      src/commandParser.c and writers link against it.
 *
 * Conventions (all from the reference): coordinates are 1-based half-open
 * [start,end) int32 (src/wigReader.c:123-124); values are double; tracks are
 * sorted, non-overlapping within a track (newMultiplexer runs inputs through
 * NonOverlappingWiggleIterator, src/multiplexer.c:163).
 *
 * Errors: every function returns 0 on success or a negative WB200_E* code and
 * records a message readable with wb200_last_error().  There is no CPU
 * fallback: without a CUDA device the create call fails.
 */
#ifndef WIGGLETOOLS_B200_H
#define WIGGLETOOLS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Statistic computed per union segment.  Each replaces the reference
 * constructor + pop function cited. */
enum wb200_op {
	WB200_SUM     = 0,  /* SumReduction      src/reducers.c:259-307 */
	WB200_PRODUCT = 1,  /* ProductReduction  src/reducers.c:313-361 */
	WB200_MEAN    = 2,  /* MeanReduction     src/reducers.c:367-422 */
	WB200_VAR     = 3,  /* VarianceReduction src/reducers.c:428-505 */
	WB200_STDDEV  = 4,  /* StdDevReduction   src/reducers.c:511-589 */
	WB200_ENTROPY = 5,  /* EntropyReduction  src/reducers.c:640-666 (runs the stddev pop, :665) */
	WB200_CV      = 6,  /* CVReduction       src/reducers.c:672-750 */
	WB200_MIN     = 7,  /* MinReduction      src/reducers.c:192-253 */
	WB200_MAX     = 8,  /* MaxReduction      src/reducers.c:125-186 */
	WB200_MEDIAN  = 9,  /* MedianReduction   src/reducers.c:780-834 */
	WB200_FILLIN  = 10, /* FillInReduction(multi, false)  src/reducers.c:78-119 (exactly 2 tracks) */
	WB200_TRIMFILL = 11,/* FillInReduction(multi, true)  (only where track 0 is in play) */
	WB200_TTEST   = 16, /* TTestReduction    src/setComparisons.c:35-131 (over a Multiset of 2) */
	WB200_MWU     = 17  /* MWUReduction      src/setComparisons.c:269-390 */
};

enum wb200_error {
	WB200_OK = 0,
	WB200_ENODEV = -1,     /* no CUDA device / driver */
	WB200_ECUDA = -2,      /* CUDA runtime error (message has the detail) */
	WB200_EINVAL = -3,     /* bad argument (op, counts, unsorted input ...) */
	WB200_ENOMEM = -4,     /* device or pinned-host allocation failed */
	WB200_ECAPACITY = -5,  /* caller-provided output capacity too small */
	WB200_EUNSUPPORTED = -6
};

#define WB200_MAX_TRACKS 1023  /* per set; cover counters are 10-bit fields.  WB200_MWU: at most 1024 tracks in the two
                                * sets together (per-warp sort rows in shared memory), else WB200_EUNSUPPORTED */

typedef struct wb200_ctx wb200_ctx;

/* One track restricted to one chromosome (host or device pointers, depending
 * on the call).  Mirrors the fields a WiggleIterator streams
 * (src/wiggleIterator.h:21-35: start, finish, value, default_value). */
typedef struct wb200_track {
	const int32_t *start;
	const int32_t *end;
	const double *value;       /* fp64 values, or NULL when value32 is given */
	int64_t n;
	double default_value;
	const float *value32;      /* optional narrow transport: float32 values (what BigWig files hold), widened to fp64
	                            * on the device -- 12 instead of 16 bytes per interval over PCIe; used when value is NULL
	                            * (host entry points only) */
} wb200_track;

/* Segments of one chromosome.  For *_host calls the arrays are owned by the
 * context (pinned host memory) and stay valid until the next call on it. */
typedef struct wb200_segments {
	int32_t *start;
	int32_t *end;          /* NULL in breakpoint form (see wb200_set_breakpoint_output) */
	double *value;
	int64_t n;
	double default_value;  /* the reducer's derived default (constructor semantics) */
	double auc;            /* sum over segments of (end-start)*value, NaN skipped (AUCPop, src/statistics.c:103-120) */
	/* breakpoint form: end[k] = start[k+1] for every segment except the n_ends listed ones (those followed by a gap,
	 * and the last), whose ends are end_value[r] at segment index end_index[r] (ascending) */
	int64_t n_ends;
	const int32_t *end_index;
	const int32_t *end_value;
} wb200_segments;

/* Timing / accounting of the last call (for bench.py). */
typedef struct wb200_stats {
	int64_t intervals;       /* I: input intervals consumed */
	int64_t segments;        /* S: segments emitted */
	int64_t windows;
	int64_t slow_segments;   /* segments evaluated per segment (row / sweep paths, fix-ups of the exact paths) */
	int32_t kernels;         /* kernel launches issued by the call */
	int32_t fast_path;       /* 0 row or sweep path, 1 exact sum path, 2 exact moments path */
	float ms_count;          /* device time of the cursor kernel (historic name) */
	float ms_emit;           /* device time of the emit kernel(s) */
	float ms_total;          /* device time of the whole call (events on the call's stream) */
	int64_t h2d_bytes, d2h_bytes;
	int32_t limbs;           /* exact sum path: 32-bit accumulator limbs per position (1..3), else 0 */
	int32_t sweep_kernel;    /* 1: the call ran on the sum / mean sweep kernel (csrc/fastsum.cu), 0: emit_kernel */
} wb200_stats;

/* -- lifetime ------------------------------------------------------------ */
/* device: CUDA ordinal.  Fails (NULL, message via wb200_last_error(NULL)) when
 * no device is usable. */
wb200_ctx *wb200_create(int device);
void wb200_destroy(wb200_ctx *ctx);
const char *wb200_last_error(const wb200_ctx *ctx);
const wb200_stats *wb200_last_stats(const wb200_ctx *ctx);
/* Use an externally owned stream (e.g. torch's current stream) for all work. */
int wb200_set_stream(wb200_ctx *ctx, void *cuda_stream);
/* Tuning knobs: window size in bp (power of two, 0 = automatic). */
int wb200_set_window(wb200_ctx *ctx, int window_bp);
/* Narrow read-back for wb200_reduce_host_batch: segments come back in BREAKPOINT FORM -- start[] and value[] in full,
 * end[] only where it is not the next segment's start (a gap follows, or the segment is the last): 12 instead of 16
 * bytes per segment over PCIe on gap-free data, the layout of a bedGraph-style step function.  The reference's
 * consumers read one record at a time (pop), so a binding derives `finish` from the next start as it goes; no host
 * pass over the arrays is needed.  A shard with explicit ends on more than a quarter of its segments is sent in
 * the full form (end != NULL).  0 (default): always the full form. */
int wb200_set_breakpoint_output(wb200_ctx *ctx, int on);

/* Pin the calling host thread to the CPUs local to the context's GPU (its PCIe root's NUMA
 * node), so that staging buffers allocated afterwards live in that node's memory: matters
 * on multi-socket boxes with one process per GPU.  Returns the number of CPUs in the mask,
 * 0 if the topology is not exposed (nothing changed). */
int wb200_bind_host_thread(wb200_ctx *ctx);

/* Pinned host memory for staging (cudaHostAlloc). */
void *wb200_host_alloc(int64_t bytes);
void wb200_host_free(void *p);

/* -- the hot path, host buffers (end-to-end entry) ----------------------- */
/* Replaces: newMultiplexer(iters,count,strict) + <op>Reduction(multi) + the
 * consumer loop, for one chromosome (src/multiplexer.c:157-169,
 * src/reducers.c, src/wiggleIterator.c:62-65).  For WB200_TTEST / WB200_MWU
 * the first nset1 tracks are set A and the rest set B (newMultiset of two
 * multiplexers, src/multiSet.c:115-128); otherwise pass nset1 = ntracks.
 * H2D staging, kernels and D2H of the result all happen inside the call. */
int wb200_reduce_host(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict,
                      const wb200_track *tracks, wb200_segments *out);

/* Raw multiplexer rows of the chromosome the last wb200_reduce_host() call on the context reduced
 * (inputs, cursor table and segments are still resident in HBM): for its segments
 * [first, first+count), values[(k-first)*ntracks + i] is track i's value where it is in play, else its
 * default_value, and inplay[...] is 1 / 0.  Replaces what popCoreMultiplexer2 leaves in
 * Multiplexer.values[] / inplay[] (src/multiplexer.c:76-96), read by TeeMultiplexerPop
 * (src/mWigWriter.c:179-225: mwrite, mwrite_bg, paste) and SelectReductionPop (src/reducers.c:40-65).
 * values / inplay are host arrays of count*ntracks entries. */
int wb200_rows_host(wb200_ctx *ctx, int64_t first, int64_t count, double *values, unsigned char *inplay);

/* Same, for several chromosome shards in one call: tracks[c*ntracks + i] is track
 * i on shard c, out[c] receives shard c's segments (all valid until the next
 * call on the context).  H2D staging of shard c+1, the kernels of shard c and the
 * D2H of shard c-1 overlap on three streams with double-buffered device memory;
 * results are bit-identical to nshards calls of wb200_reduce_host. */
int wb200_reduce_host_batch(wb200_ctx *ctx, int op, int nshards, int ntracks, int nset1, int strict,
                            const wb200_track *tracks, wb200_segments *out);

/* -- the hot path, device-resident buffers ------------------------------- */
/* d_start/d_end/d_value: all tracks of the chromosome concatenated; h_off[i]
 * .. h_off[i+1] is track i (host array of ntracks+1 offsets).  Positions lie in
 * [pos_lo, pos_hi].  Output arrays (device) must hold out_cap segments; an
 * upper bound is min(2*I+1, pos_hi-pos_lo+1).  *n_out receives S. */
int wb200_reduce_device(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict,
                        const int32_t *d_start, const int32_t *d_end, const double *d_value,
                        const int64_t *h_off, const double *h_defaults,
                        int32_t pos_lo, int32_t pos_hi,
                        int32_t *d_out_start, int32_t *d_out_end, double *d_out_value,
                        int64_t out_cap, int64_t *n_out, double *default_out, double *auc_out);

/* Several HBM-resident chromosome shards in one call (all with the same tracks count and
 * defaults).  The value samples of all shards are read back with one synchronisation, then the
 * cursor kernel and the sweep of every shard are queued back to back and all results are read at
 * the end: two host round trips per call instead of two per chromosome.  Results are identical
 * to nshards calls of wb200_reduce_device. */
typedef struct wb200_device_shard {
	const int32_t *d_start;      /* in: device, all tracks of the shard concatenated */
	const int32_t *d_end;
	const double *d_value;
	const int64_t *h_off;        /* in: host, ntracks+1 offsets */
	int32_t pos_lo, pos_hi;
	int32_t *d_out_start;        /* in: device output arrays of out_cap segments */
	int32_t *d_out_end;
	double *d_out_value;
	int64_t out_cap;
	int64_t n_out;               /* out: S */
	double auc;                  /* out */
	int64_t slow_segments;       /* out */
	float ms_cursor, ms_sweep;   /* out: device time of the two kernels of this shard (CUDA events) */
	int32_t limbs;               /* out: accumulator limbs of the exact sum path, else 0 */
	int32_t sweep_kernel;        /* out: 1 = fastsum_kernel ran */
} wb200_device_shard;
int wb200_reduce_device_batch(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict, const double *h_defaults,
                              int nshards, wb200_device_shard *shards, double *default_out);

/* -- several GPUs of one box, one process ------------------------ */
/* Chromosome shards are packed onto the devices (longest processing time first, by interval
 * count); one host thread per device runs that device's context through wb200_reduce_host_batch
 * (or wb200_pearson_host).  No per-segment data crosses GPUs.  The whole-genome scalars do, over
 * NCCL (libnccl.so.2, loaded at run time): the AUC partials with ncclAllReduce(ncclDouble,
 * ncclSum), the six Pearson moments of every device with ncclAllGather, folded in device order
 * with wb200_pearson_merge -- what the reference does with one process per genome region
 * (python/wiggletools/parallelWiggleTools.py:66-68) and a sum of the printed AUCs. */
typedef struct wb200_multi wb200_multi;
wb200_multi *wb200_multi_create(int ndev, const int *devices);   /* ndev <= 0: every visible device; devices NULL: 0..ndev-1 */
void wb200_multi_destroy(wb200_multi *m);
const char *wb200_multi_last_error(const wb200_multi *m);
int wb200_multi_ndev(const wb200_multi *m);
int wb200_multi_owner(const wb200_multi *m, int shard);           /* device index that ran the shard in the last call */
/* tracks[c*ntracks + i], out[c] as in wb200_reduce_host_batch (out[c]'s arrays belong to the context of the
 * device that owns shard c, valid until the next call); *auc_total = the NCCL-all-reduced integral. */
int wb200_multi_reduce_host(wb200_multi *m, int op, int nshards, int ntracks, int nset1, int strict,
                            const wb200_track *tracks, wb200_segments *out, double *auc_total);
/* x[c], y[c]: the two tracks on shard c, in chromosome order; moments = the six running moments of the genome. */
int wb200_multi_pearson_host(wb200_multi *m, int nshards, const wb200_track *x, const wb200_track *y, double moments[6]);

/* -- whole-genome scalars ------------------------------------------------ */
/* Pearson correlation of two tracks over one chromosome shard: returns the six
 * running moments so that shards (chromosomes, GPUs) can be merged exactly as
 * the reference's sequential update would (PearsonPop, src/statistics.c:414-458).
 * moments = {count, sum_X, sum_Y, T_XX, T_XY, T_YY}. */
int wb200_pearson_host(wb200_ctx *ctx, const wb200_track *x, const wb200_track *y, double moments[6]);
/* Same for tracks px and py of an ntracks-wide multiplexer: segments where only other
 * tracks are in play still count, with the two tracks' default values (what
 * NDPearsonPop sees through a Multiset, src/statistics.c:490-542). */
int wb200_pearson_pair_host(wb200_ctx *ctx, int ntracks, const wb200_track *tracks, int px, int py, double moments[6]);
/* Fold shard b into a (the sequential update applied to pre-reduced shards). */
void wb200_pearson_merge(double a[6], const double b[6]);
double wb200_pearson_result(const double moments[6]);

/* -- host BigWig batch decoder ------------------------- */
/* Stands where the reference goes through libBigWig one record at a time
 * (src/bigWiggleReader.c:36-123): a whole chromosome is decoded at once (R-tree
 * walk, parallel zlib inflate) into pinned SoA arrays.  Chromosomes are indexed
 * in strcmp order of their names (the reference's visiting order, :97).
 * Coordinates come out 1-based half-open; `stretch` (default 10000) reproduces
 * the reference's chopping of intervals at 1+10000k (quirk Q9), 0 disables it. */
typedef struct wb200_bigwig wb200_bigwig;
wb200_bigwig *wb200_bw_open(const char *path);               /* NULL on error, see wb200_bw_last_error(NULL) */
void wb200_bw_close(wb200_bigwig *bw);
const char *wb200_bw_last_error(const wb200_bigwig *bw);
int wb200_bw_nchrom(const wb200_bigwig *bw);
const char *wb200_bw_chrom_name(const wb200_bigwig *bw, int i);
int64_t wb200_bw_chrom_length(const wb200_bigwig *bw, int i);
void wb200_bw_set_stretch(wb200_bigwig *bw, int stretch_bp);
/* out->start/end/value are owned by the handle and valid until its next decode. */
int wb200_bw_decode_chrom(wb200_bigwig *bw, int chrom_index, wb200_track *out);

/* -- element-wise maps on the inputs --------------------- */
/* `mean map ln a b c`, `sum scale 2 x : ...`: the reference wraps each source in a
 * unary iterator that rewrites one record at a time (src/unaryOps.c:650-960,
 * built by src/commandParser.c:115-211).  Here a chain of up to 8 maps per track
 * is applied to the staged values in HBM before the sweep:
 *   SCALE  s   NaN -> NaN, else s*v                         ScaleWiggleIteratorPop  :650-665
 *   OFFSET s   s+v                                          ShiftWiggleIteratorPop  :722-734
 *   LOG    l   records with v <= 0 are dropped, NaN stays,
 *              else log(v)/l  (l = 1 for `ln`, log(base) for `log base`)   LogWiggleIteratorPop :760-779
 *   EXP    l   exp(v*l)       (l = 1 for `exp`)             ExpWiggleIteratorPop    :823-835
 *   POW    s   (s<0 and v<=0) or NaN -> NaN, else pow(v,s)  PowerWiggleIteratorPop  :873-889
 *   ABS        NaN -> NaN, else |v|                         AbsWiggleIteratorPop    :934-949
 * wb200_set_input_maps() arms the chains for the NEXT wb200_reduce_host /
 * wb200_reduce_format_host call on the context (one chromosome; call it again for
 * the next one).  wb200_track.default_value must already be the mapped default:
 * wb200_map_default() applies one map's constructor rule (some round through
 * float, as the reference does). */
enum wb200_map_op {
	WB200_MAP_SCALE = 1, WB200_MAP_OFFSET = 2, WB200_MAP_LOG = 3, WB200_MAP_EXP = 4, WB200_MAP_POW = 5, WB200_MAP_ABS = 6
};
typedef struct wb200_map { int32_t op; int32_t reserved; double arg; } wb200_map;
int wb200_set_input_maps(wb200_ctx *ctx, int ntracks, const wb200_map *const *maps, const int *nmaps);
double wb200_map_default(int op, double arg, double default_value);

/* -- regional integrators on the device ------------------ */
/* `apply` / `apply_paste`: for every region of a sorted BED-like list, the integrators of
 * src/statistics.c:62-326 over the dataset restricted to the region -- what ApplyMultiplexerPop /
 * computeApplyValues (src/apply.c:259-321) compute through a per-base buffer
 * (BufferedWiggleIterator, apply.c:58-160): bases the dataset does not cover count as one run of
 * its default value (rounded through float, as the buffer keeps it) unless `strict`; NaN records
 * are skipped by every integrator.  One call per chromosome: the dataset is a sorted, disjoint
 * segment list (typically a reduction's output still in HBM), regions are 1-based half-open and
 * may overlap.  out[r*nstats + k] = statistic codes[k] of region r. */
enum wb200_stat {
	WB200_STAT_AUC = 1,     /* AUCIntegrator      src/statistics.c:103-127 */
	WB200_STAT_MEAN = 2,    /* MeanIntegrator     src/statistics.c:62-97 */
	WB200_STAT_VAR = 3,     /* VarianceIntegrator src/statistics.c:209-268 */
	WB200_STAT_STDDEV = 4,  /* StandardDeviationIntegrator :270-296 */
	WB200_STAT_CV = 5,      /* CoefficientOfVariationIntegrator :298-326 */
	WB200_STAT_MAX = 6,     /* MaxIntegrator      :157-181 */
	WB200_STAT_MIN = 7,     /* MinIntegrator      :183-207 */
	WB200_STAT_SPAN = 8     /* SpanIntegrator     :129-155 */
};
int wb200_apply_device(wb200_ctx *ctx, const int32_t *d_start, const int32_t *d_end, const double *d_value, int64_t nseg,
                       double default_value, int64_t nreg, const int32_t *rstart, const int32_t *rend,
                       int nstats, const int *codes, int strict, double *out);
/* Same with the dataset's segments in host memory (staged inside the call). */
int wb200_apply_host(wb200_ctx *ctx, const int32_t *start, const int32_t *end, const double *value, int64_t nseg,
                     double default_value, int64_t nreg, const int32_t *rstart, const int32_t *rend,
                     int nstats, const int *codes, int strict, double *out);
/* Same with the segments of the last wb200_reduce_host() call on the context as the dataset (they are
 * still resident in HBM): `apply_paste out meanI regions.bed mean a.bw b.bw c.bw` per chromosome. */
int wb200_apply_last(wb200_ctx *ctx, double default_value, int64_t nreg, const int32_t *rstart, const int32_t *rend,
                     int nstats, const int *codes, int strict, double *out);

/* -- text sink on the device --------------------------- */
/* Stands where the reference's default sink works on one host thread: the
 * writer-side CompressionWiggleIterator (src/unaryOps.c:235-263; `write` only),
 * the 10000-record block hand-over of TeeWiggleIteratorPop (src/wigWriter.c:159-200)
 * and printBlock (src/wigWriter.c:55-119): "fixedStep chrom=.. start=.. step=1"
 * runs for records shorter than 2 bp (left again above 5 bp), bedGraph lines
 * otherwise, values as "%lf".  The bytes returned are exactly what
 * toFile()/toStdout() (src/wigWriter.c:278-290) print for the same records.
 *
 * One call per chromosome, in the order the records reach the sink; `state`
 * (zero-initialised before the first call) carries what the reference's writer
 * carries from one record to the next: the number of records so far (its
 * per-block reset) and the point-by-point flag.  *text points to pinned host
 * memory owned by the context, valid until the next call on it.
 * WB200_EUNSUPPORTED for |value| >= 2^63 (the device formatter prints up to 19
 * integer digits). */
enum wb200_sink_mode {
	WB200_SINK_WIG = 0,       /* `write`:    toFile(wi, f, false, ..)  src/wigWriter.c:278 */
	WB200_SINK_BEDGRAPH = 1   /* `write_bg`: toFile(wi, f, true, ..)   no compression, bedGraph lines only */
};
typedef struct wb200_sink_state {
	int64_t records;          /* records handed to the sink so far */
	int32_t point_by_point;   /* printBlock's pointByPoint after the last record */
	int32_t reserved;
} wb200_sink_state;
int wb200_format_device(wb200_ctx *ctx, int mode, const char *chrom,
                        const int32_t *d_start, const int32_t *d_end, const double *d_value, int64_t n,
                        wb200_sink_state *state, const char **text, int64_t *nbytes);
/* Reduction and sink in one call (what `write out.wig mean a b c` runs per
 * chromosome): wb200_reduce_host without the read-back of the segments, then
 * wb200_format_device on them.  n_segments / auc may be NULL. */
int wb200_reduce_format_host(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict, const wb200_track *tracks,
                             int mode, const char *chrom, wb200_sink_state *state,
                             const char **text, int64_t *nbytes, int64_t *n_segments, double *auc);
/* Same with the segments in host memory (staged to HBM inside the call). */
int wb200_format_host(wb200_ctx *ctx, int mode, const char *chrom,
                      const int32_t *start, const int32_t *end, const double *value, int64_t n,
                      wb200_sink_state *state, const char **text, int64_t *nbytes);

/* Student-t upper tail used by WB200_TTEST (stands where the reference calls
 * gsl_cdf_tdist_Q, src/setComparisons.c:117); exported for tests. */
double wb200_tdist_Q(double t, double nu);

/* Diagnostic: the same function as the DEVICE evaluates it (one thread per pair): out[i] = Q(t[i], nu[i]). */
int wb200_debug_tdist_device(wb200_ctx *ctx, int n, const double *t, const double *nu, double *out);
/* Development: dynamic shared memory (bytes) one CTA of emit_kernel requests for `op` over n_tracks with
 * window_bp-wide windows (csrc/engine.cu: make_layout + walk_layout).  Host arithmetic only. */
long long wb200_debug_window_smem_bytes(int op, int n_tracks, int window_bp);

/* Diagnostic: the device formatter's "%lf" compiled for the host (same source function as the
 * kernel's): writes the text of v into out32 (>= 32 bytes), returns its length, -1 for |v| >= 2^63. */
int wb200_debug_format_value(double v, char *out32);

/* Diagnostic: per-phase cycle counters of the window kernel (thread 0 of every CTA, summed over
 * windows) since the last call; zeros unless the library was built with -DWB_DIAG_PHASES. */
void wb200_debug_phases(wb200_ctx *ctx, uint64_t out[16]);

#ifdef __cplusplus
}
#endif
#endif
