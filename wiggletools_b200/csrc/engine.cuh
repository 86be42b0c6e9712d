// wiggletools_b200 -- device-side job description shared by the kernels
// (engine.cu) and the host C ABI (capi.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace wb {

#ifndef WB_THREADS
#define WB_THREADS 256
#endif
constexpr int THREADS = WB_THREADS;    // threads per CTA in the window kernels
constexpr int TLOG = THREADS == 512 ? 9 : 8;
static_assert(THREADS == 256 || THREADS == 512, "window kernels are written for 256 or 512 threads");
constexpr int WARPS = THREADS / 32;
constexpr int MAX_TRACKS = 1023;       // per set: 10-bit cover fields

// Cover word layout (one int32 per breakpoint slot): the prefix sum over slots
// gives, for the segment that starts at the slot,
//   bits  0.. 9  tracks of set A in play   (all tracks for a plain multiplexer)
//   bits 10..19  tracks of set B in play
//   bits 20..31  tracks in play whose value is not exactly representable in
//                the fixed-point accumulator ("poison": NaN, Inf, out of range)
constexpr int COV_B_SHIFT = 10;
constexpr int COV_P_SHIFT = 20;
constexpr int COV_MASK = 1023;

enum Op : int {
	OP_SUM = 0, OP_PRODUCT = 1, OP_MEAN = 2, OP_VAR = 3, OP_STDDEV = 4, OP_ENTROPY = 5,
	OP_CV = 6, OP_MIN = 7, OP_MAX = 8, OP_MEDIAN = 9, OP_FILLIN = 10, OP_TRIMFILL = 11, OP_TTEST = 16, OP_MWU = 17,
	OP_PEARSON = 32   // internal: 2-track moments (PearsonPop, src/statistics.c:414-458)
};

struct Job {
	// inputs: all tracks of one chromosome, concatenated
	const int32_t *start;
	const int32_t *end;
	const double *value;
	const int64_t *off;     // [N+1] device
	const double *dflt;     // [N] device: default_value per track
	const int64_t *dflt_fx; // [N] device: fixed-point defaults (fast path)
	int N;                  // tracks
	int64_t I;              // total intervals (length of start/end/value)
	int vec;                // inputs are 16-byte aligned: 128-bit interval loads
	int n1;                 // tracks in set A (== N for a plain multiplexer)
	int strict;
	int op;
	// windowing
	int64_t p0;             // position of window 0, slot 0
	int wlog;               // log2(window size in bp)
	int nw;                 // number of windows
	// scratch
	int32_t *cursor;        // [(nw+1) * cstride] first interval of track i with end >= window start (row nw: end of data)
	int cstride;            // row stride of the cursor table: N rounded up to 4 (rows are 16-byte aligned: bulk-copy tiles)
	unsigned long long *state; // [nw] look-back words: 2 flag bits | segments (aggregate or inclusive prefix)
	unsigned int *ticket;   // window ticket counter
	long long *total;       // S, written by the last window
	int lb_warps;           // warps used by the first look-back step (1..WARPS)
	// outputs
	int32_t *out_start;
	int32_t *out_end;
	double *out_value;
	int64_t out_cap;
	double *auc_partial;    // [nw * WARPS] per-warp sums of (end-start)*value (NaN skipped), or null
	unsigned long long *slow_count; // segments that took the row path
	double *moments;        // [nw * 6] per-window Pearson moments (OP_PEARSON), or null
	int px, py;             // OP_PEARSON: the two tracks correlated (the others only shape the cover)
	long long *slow_list;   // fast path: output indices of segments to recompute (fix-up kernel)
	long long slow_cap;
	double *auc_fix;        // AUC contribution of the fixed-up segments
	// fixed-point fast path (sum / mean)
	int fast;               // 1: scan path enabled for this op
	int fx_q;               // value = m * 2^fx_q
	int fx_mbits;           // |m| < 2^fx_mbits
	int fx_limbs;           // sum / mean: 32-bit limbs per position (1..3)
	int64_t fx_base;        // sum of fixed-point defaults (all tracks absent)
	int sum1;               // moments path: the plain sum lives in ONE 32-bit limb (N * 2^fx_mbits < 2^31)
	unsigned long long *diag; // [16] diagnostic builds: per-phase cycle counters
	int *flags;             // input validation (cursor kernel): bit 0 unsorted / overlapping / empty interval, bit 1 outside [p0, pos_hi]
	int pos_hi;             // last position the windows cover
	double mwu_mu, mwu_sigma; // MWU normal approximation moments (host-computed, integer division quirk)
};

// Pearson moments m = {n, sum_X, sum_Y, T_XX, T_XY, T_YY}; fold shard b into a.
// Pairwise form of the reference's sequential update (src/statistics.c:443-456).
__host__ __device__ inline void merge6(double *a, const double *b) {
	if (b[0] == 0) return;
	if (a[0] == 0) { for (int k = 0; k < 6; k++) a[k] = b[k]; return; }
	const double n = a[0] + b[0];
	const double dx = a[1] / a[0] - b[1] / b[0], dy = a[2] / a[0] - b[2] / b[0];
	const double w = a[0] * b[0] / n;
	a[3] += b[3] + w * dx * dx;
	a[4] += b[4] + w * dx * dy;
	a[5] += b[5] + w * dy * dy;
	a[0] = n; a[1] += b[1]; a[2] += b[2];
}

// register-friendly form of the same merge (no arrays)
struct Mom { double n, sx, sy, txx, txy, tyy; };
__host__ __device__ inline void merge_mom(Mom &a, const Mom &b) {
	if (b.n == 0) return;
	if (a.n == 0) { a = b; return; }
	const double n = a.n + b.n;
	const double dx = a.sx / a.n - b.sx / b.n, dy = a.sy / a.n - b.sy / b.n;
	const double w = a.n * b.n / n;
	a.txx += b.txx + w * dx * dx;
	a.txy += b.txy + w * dx * dy;
	a.tyy += b.tyy + w * dy * dy;
	a.n = n; a.sx += b.sx; a.sy += b.sy;
}

// launch helpers implemented in engine.cu
size_t window_smem_bytes(int wlog, int N, bool emit, int op, int limb_arrays);
cudaError_t launch_cursor(const Job &job, int64_t max_track_len, cudaStream_t s);
// auc_per_window: how many AUC partials per window the chosen kernel writes to job.auc_partial
cudaError_t launch_emit(const Job &job, cudaStream_t s, int *auc_per_window);
// sum / mean sweep kernel (fastsum.cu)
bool fastsum_supported(const Job &job);
cudaError_t launch_fastsum(const Job &job, int lw, double inv_scale, double lim, double scale, cudaStream_t s);
cudaError_t launch_fixup(const Job &job, long long count, cudaStream_t s);
cudaError_t launch_tdist(const double *t, const double *nu, int n, double *out, cudaStream_t s);
cudaError_t launch_rows(const Job &job, long long first, long long count, double *values, unsigned char *inplay, cudaStream_t s);
cudaError_t launch_sample_maxabs(const double *value, int64_t n, int64_t stride, double *out2, cudaStream_t s);
cudaError_t launch_reduce_auc(const double *partial, int count, double *out, cudaStream_t s);
cudaError_t launch_reduce_moments(const double *moments, int nw, double *out6, cudaStream_t s);
// regional integrators over HBM-resident segments (apply.cu)
cudaError_t launch_apply(const int32_t *start, const int32_t *end, const double *value, long long nseg, double dflt, int strict,
                         const int32_t *rstart, const int32_t *rend, long long nreg, int nstats, const int *codes, double *out,
                         bool var_family, cudaStream_t s);
// element-wise maps on staged inputs (maps.cu)
constexpr int MAX_MAP_CHAIN = 8;
enum MapCode : int { MAP_SCALE = 1, MAP_OFFSET = 2, MAP_LOG = 3, MAP_EXP = 4, MAP_POW = 5, MAP_ABS = 6 };
struct MapOp { int op; int pad; double arg; };
cudaError_t launch_map(double *value, const int64_t *d_off, const MapOp *d_chains, const int *d_nchain, int N, int64_t max_len,
                       unsigned char *keep, int *any_dropped, cudaStream_t s);
cudaError_t launch_map_gather(const int32_t *hp, const long long *d_R, long long R, const int32_t *s_in, const int32_t *e_in,
                              const double *v_in, int32_t *s_out, int32_t *e_out, double *v_out,
                              const int64_t *d_off, int N, int64_t *d_new_off, cudaStream_t s);
// narrow transport (maps.cu): float32 -> fp64 widening of staged values; segments whose end is not the next start
cudaError_t launch_widen(const float *in, double *out, long long n, cudaStream_t s);
cudaError_t launch_explicit_ends(const int32_t *start, const int32_t *end, long long S, unsigned char *H, cudaStream_t s);
cudaError_t launch_gather_ends(const int32_t *hp, long long R, const int32_t *end, int32_t *out, cudaStream_t s);
// positions of the set bytes of H[0..n): hp[0..R), hp[R] = n (count / scan / scatter of writer.cu)
cudaError_t sink_compact(const unsigned char *H, int64_t n, int *tile_count, long long *tile_base, long long *R, int32_t *hp, cudaStream_t s);
// text sink (writer.cu): record heads of the compress rule, then the two formatting passes
cudaError_t sink_heads(const int32_t *start, const int32_t *end, const double *value, int64_t n, unsigned char *D, unsigned char *H,
                       int *tile_count, long long *tile_base, long long *R, int32_t *hp, cudaStream_t s);
int sink_tiles(int64_t n);
int sink_debug_format_value(double v, char *out);
int sink_block_records();
cudaError_t sink_format(bool write, const int32_t *start, const int32_t *end, const double *value, const int32_t *hp,
                        long long R, long long rec_base, int wig, int pbp_in, const char *chrom,
                        int *block_bytes, const long long *block_off, char *text, int *flags, int nblocks, cudaStream_t s);

} // namespace wb
