// wiggletools_b200 -- sm_100a kernels of the interval-stream reduction engine.
//
// What the reference does one segment at a time on one CPU thread
// (popCoreMultiplexer2, /root/reference/src/multiplexer.c:98-121, followed by
// an O(N) reducer loop such as MeanReductionPop, src/reducers.c:367-402), this
// file does per genomic WINDOW of 256*2^VLOG base pairs, one CTA per window:
//
//   1. events : every interval adds +delta at its start position and -delta at
//               its end position of a position-indexed shared-memory array: a
//               packed cover counter (how many tracks of set A/B are in play)
//               and the value as a fixed-point integer split in 32-bit limbs
//               (native shared-memory atomics; integer adds commute, so the
//               result is order-independent and exact -- no floating-point
//               atomics anywhere).  Replaces the two Fibonacci heaps of
//               multiplexer.c:37-96;
//   2. scan   : prefix sum over positions = state of the sweep after each
//               breakpoint = cover set + exact sum of the segment that starts
//               there (the reference recomputes this with an O(N) loop per
//               segment; here it is O(1) per segment);
//   3. compact: positions where an interval starts or ends are the breakpoints;
//               they are packed into a dense list;
//   4. look-back: the window's output offset comes from a decoupled look-back
//               over the preceding windows (single pass, no counting kernel);
//   5. emit   : segments with a non-empty cover (or a full one under `strict`,
//               multiplexer.c:123-128) are written coalesced.
//
// Kernel variants (emit_kernel<MODE,...>):
//   MODE_FAST    sum / mean from one exact accumulator;
//   MODE_MOMENTS var / stddev / CV from exact sum and sum of squares;
//   MODE_WALK    min / max / median from an incrementally maintained sorted array (up to WB_SEL_MIN_N = 24 tracks);
//   MODE_SELECT  min / max / median from the window's sorted universe (more tracks: bucketed selection, bitset sweep);
//   MODE_ROW     everything else: row_eval(), a per-segment loop over the N
//                tracks in index order that mirrors the reference's arithmetic
//                operation for operation (also the fix-up path for segments whose
//                inputs do not fit the fixed-point formats: NaN, Inf, huge range).
// cursor_kernel builds, per window and track, the first interval that can touch
// the window, so that every window starts independently.
#include "engine.cuh"
#include "device_util.cuh"
#include <cstdio>
#include <cstdlib>
#include <cuda_pipeline.h>
#include <cfloat>
#include <cstdlib>
#include <climits>
#include <cmath>

namespace wb {

#ifndef WB_PREF
#define WB_PREF 0               // events phase: 0 plain loads, 1 register double buffering, 2 L1 prefetch, 3 L2 prefetch
#endif
#ifndef WB_LATETICKET
#define WB_LATETICKET 1         // next window drawn when the emit phase starts (0: two windows ahead); see fastsum.cu
#endif
#ifndef WB_FAST_MINB
#define WB_FAST_MINB 3          // resident CTAs per SM the sum/mean kernel is compiled for
#endif

// ---------------------------------------------------------------------------
// shared-memory layout
// ---------------------------------------------------------------------------
// A window of W = THREADS << VLOG base pairs is owned by one CTA; thread t owns
// the V = 2^VLOG consecutive positions [t*V, t*V+V).  Per position the window
// keeps a packed cover delta, the fixed-point value delta in 32-bit limbs and a
// byte flag "an interval starts or ends here".  The 32-bit arrays are read back
// by their owner with 128-bit loads; phys() rotates each thread's vectors so
// that those loads (and the scattered atomics) are bank-conflict free.
template <int VLOG>
__device__ __forceinline__ int phys(int p) {
	if (VLOG < 2) return p;
	constexpr int LV = VLOG - 2, NV = 1 << (VLOG >= 2 ? VLOG - 2 : 0);
	const int t = p >> VLOG, o = p & ((1 << VLOG) - 1);
	const int q = (o >> 2) ^ ((t >> (3 - LV)) & (NV - 1));
	return (t << VLOG) | (q << 2) | (o & 3);
}

struct Misc {                        // scalars, warp totals
	int next_bp;                     // first breakpoint at or after the window end
	int base_cov;                    // packed cover carried into the window
	int kept;                        // kept segments of the window
	int nbp;                         // breakpoints of the window
	unsigned long long base_val;     // fixed-point value carried into the window
	long long out_base;              // output offset of the window
	int ticket, ticket1, ticket2;    // window indices drawn by this CTA (current, next, after next)
	int pad0;
	long long wsum[WARPS];           // warp totals (value scan / slow counter)
	long long wsum2[WARPS];          // warp totals (sum-of-squares scan)
	unsigned long long base_val2;    // carried-in sum of squares
	int wcov[WARPS];                 // warp totals (cover scan)
	int wkeep[WARPS];                // warp totals (counts scan)
	double wauc[WARPS];
	double wmom[WARPS][6];           // Pearson moments per warp
	unsigned long long lb_sum[WARPS]; // look-back: per-warp chunk sums
	int lb_found[WARPS];             // look-back: chunk holds an inclusive prefix
};
constexpr int MISC_BYTES = 2048;
constexpr int MISC_SCRATCH = 1024;      // the second half of the slot: 1 KB of scratch for whoever needs it between two barriers
static_assert(sizeof(Misc) <= MISC_SCRATCH, "Misc must fit its shared-memory slot");

constexpr int WALK_MAX_TRACKS = 512;    // order-statistic sweep: N keys per warp in shared memory
constexpr uint16_t BK_KEEP = 0x8000, BK_SLOW = 0x4000, BK_IDX = 0x3fff;

struct Layout {
	size_t off_cur, off_cur2, off_nxt, off_tn, off_toff, off_tmd, off_cov, off_limb, off_ev, off_rows, total;
	// after the scan the position arrays are dead; the dense breakpoint list aliases them:
	size_t off_bpos, off_bkid, off_bval, off_bval2, off_bcnt;
};

// limbs = total number of 32-bit limb arrays (accumulators x limbs per accumulator)
__host__ __device__ inline Layout make_layout(int wlog, int N, bool emit, int limbs, int rowbuf) {
	Layout L;
	const size_t W = (size_t) 1 << wlog;
	size_t o = MISC_BYTES;
	const size_t nn = (size_t) (N + 4) / 4 * 16;
	L.off_cur = o;  o += nn;
	L.off_cur2 = o; o += nn;
	L.off_nxt = o;  o += nn;
	L.off_tn = o;   o += nn;
	L.off_toff = o; o += 2 * nn;
	L.off_tmd = o;  o += emit ? 2 * nn : 0;
	L.off_cov = o;
	L.off_limb = o + 4 * W;
	L.off_ev = L.off_limb + (emit ? (size_t) limbs * 4 * W : 0);
	const size_t pos_bytes = (L.off_ev + W + 16) - L.off_cov;
	// dense list: sum[W] i64 | sum2[W] i64 (moments) | pos[W+8] u16 | kid[W] u16 | count[W] u16 (moments)
	const bool two = limbs > 3;
	L.off_bval = L.off_cov;
	L.off_bval2 = L.off_bval + 8 * W;
	L.off_bpos = L.off_bval2 + (two ? 8 * W : 0);
	L.off_bkid = L.off_bpos + 2 * W + 16;
	L.off_bcnt = L.off_bkid + 2 * W;
	const size_t dense_bytes = (L.off_bcnt + (two ? 2 * W : 0)) - L.off_cov;
	o += pos_bytes > dense_bytes ? pos_bytes : dense_bytes;
	L.off_rows = (o + 15) / 16 * 16; o = L.off_rows + (size_t) rowbuf;
	L.total = (o + 15) / 16 * 16;
	return L;
}

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------

// first j in [0,n) with a[j] >= key (a ascending)
__device__ __forceinline__ int lower_bound_ge(const int32_t *a, int n, long long key) {
	int lo = 0, hi = n;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if ((long long) a[mid] < key) lo = mid + 1; else hi = mid;
	}
	return lo;
}


// Student-t upper tail, Q(t; nu) = 1/2 I_{nu/(nu+t^2)}(nu/2, 1/2): stands where the
// reference calls gsl_cdf_tdist_Q (src/setComparisons.c:117).  Continued
// fraction of the regularised incomplete beta function (modified Lentz).
__host__ __device__ inline double betacf(double a, double b, double x) {
	const double tiny = 1e-300;
	double qab = a + b, qap = a + 1.0, qam = a - 1.0;
	double c = 1.0, d = 1.0 - qab * x / qap;
	if (fabs(d) < tiny) d = tiny;
	d = 1.0 / d;
	double h = d;
	for (int m = 1; m <= 10000; m++) {
		int m2 = 2 * m;
		double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
		d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
		c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
		d = 1.0 / d;
		h *= d * c;
		aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
		d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
		c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
		d = 1.0 / d;
		double del = d * c;
		h *= del;
		if (fabs(del - 1.0) < 1e-16) break;
	}
	return h;
}


// ln B(a,b) = lgamma(a)+lgamma(b)-lgamma(a+b).  With b = 1/2 and large a the
 // lgamma difference cancels catastrophically; use the asymptotic series of
 // ln[Gamma(a+1/2)/Gamma(a)] there.
__host__ __device__ inline double wt_lbeta(double a, double b) {
	if (b == 0.5 && a >= 20.0) {
		double r = 1.0 / a, r2 = r * r;
		double ratio = 0.5 * log(a) - r / 8 + r * r2 / 192 - r * r2 * r2 / 640 + 17 * r * r2 * r2 * r2 / 14336;
		return 0.5723649429247001 - ratio; // ln sqrt(pi)
	}
	if (a == 0.5 && b >= 20.0) {
		double r = 1.0 / b, r2 = r * r;
		double ratio = 0.5 * log(b) - r / 8 + r * r2 / 192 - r * r2 * r2 / 640 + 17 * r * r2 * r2 * r2 / 14336;
		return 0.5723649429247001 - ratio;
	}
	return lgamma(a) + lgamma(b) - lgamma(a + b);
}

__host__ __device__ inline double incbeta(double a, double b, double x, double xc) {
	if (x <= 0.0) return 0.0;
	if (xc <= 0.0) return 1.0;
	double bt = exp(-wt_lbeta(a, b) + a * (xc < 0.5 ? log1p(-xc) : log(x)) + b * (x < 0.5 ? log1p(-x) : log(xc)));
	if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
	return 1.0 - bt * betacf(b, a, xc) / b;
}

__host__ __device__ inline double tdist_Q(double x, double nu) {
	if (isnan(x) || isnan(nu)) return NAN;
	double x2 = x * x;
	double z = nu / (nu + x2), zc = x2 / (nu + x2);
	if (isinf(x)) { z = 0.0; zc = 1.0; }
	double tail = 0.5 * incbeta(0.5 * nu, 0.5, z, zc);
	return x >= 0 ? tail : 1.0 - tail;
}

// ---------------------------------------------------------------------------
// row path: value of track i on the segment starting at position b
// ---------------------------------------------------------------------------
struct Row {
	const Job *job;
	const int32_t *cur, *nxt;
	int b;
	// returns true if track i is in play on [b, ...); x = value or default
	__device__ __forceinline__ bool get(int i, double &x) const {
		const long long o = job->off[i];
		const int n = (int) (job->off[i + 1] - o);
		int lo = cur[i], hi = min(nxt[i], n - 1);
		bool in = false;
		if (lo <= hi) {
			const int32_t *st = job->start + o;
			// last j in [lo,hi] with start[j] <= b
			if (st[lo] <= b) {
				while (lo < hi) {
					int mid = (lo + hi + 1) >> 1;
					if (st[mid] <= b) lo = mid; else hi = mid - 1;
				}
				if (job->end[o + lo] > b) { in = true; x = job->value[o + lo]; }
			}
		}
		if (!in) x = job->dflt[i];
		return in;
	}
};

// Per-segment statistic computed exactly as the reference's pop functions do:
// same loop order over tracks, same float/double conversions, same NaN rules.
__device__ __noinline__ double row_eval(const Row &r) {
	const Job &job = *r.job;
	const int N = job.N;
	double x;
	switch (job.op) {
	case OP_SUM: case OP_MEAN: {         // src/reducers.c:259-292, :367-402
		double acc = 0;
		for (int i = 0; i < N; i++) { r.get(i, x); if (isnan(x)) return NAN; acc += x; }
		return job.op == OP_MEAN ? acc / N : acc;
	}
	case OP_PRODUCT: {                   // :313-346
		double acc = 1;
		for (int i = 0; i < N; i++) { r.get(i, x); if (isnan(x)) return NAN; acc *= x; }
		return acc;
	}
	case OP_MIN: case OP_MAX: {          // :125-168, :192-235 (absent track 0 -> literal 0)
		bool in0 = r.get(0, x);
		double acc = in0 ? x : 0.0;
		if (isnan(acc)) return acc;
		for (int i = 1; i < N; i++) {
			r.get(i, x);
			if (isnan(x)) return NAN;
			if (job.op == OP_MAX ? acc < x : acc > x) acc = x;
		}
		return acc;
	}
	case OP_VAR: {                       // :428-479
		double mean = 0, count = 0;
		for (int i = 0; i < N; i++) {
			r.get(i, x);
			float f = (float) x;
			if (isnan(f)) { mean = NAN; break; }
			mean += f; count++;
		}
		if (count < 2 || isnan(mean)) return NAN;
		mean /= count;
		double acc = 0;
		for (int i = 0; i < N; i++)
			if (r.get(i, x)) { double d = mean - x; acc += d * d; }
		return acc / count;
	}
	case OP_STDDEV: case OP_ENTROPY: {   // :511-563 (entropy registers the stddev pop, :665)
		double mean = 0;
		for (int i = 0; i < N; i++) {
			r.get(i, x);
			float f = (float) x;
			if (isnan(f)) return NAN;
			mean += f;
		}
		mean /= N;
		double acc = 0;
		for (int i = 0; i < N; i++) { r.get(i, x); double d = mean - x; acc += d * d; }
		return sqrt(acc / N);
	}
	case OP_CV: {                        // :672-725
		double mean = 0;
		for (int i = 0; i < N; i++) {
			r.get(i, x);
			float f = (float) x;
			mean += f;
			if (isnan(f)) return NAN;
		}
		mean /= N;
		if (mean == 0) return NAN;
		double acc = 0;
		for (int i = 0; i < N; i++) { r.get(i, x); double d = mean - x; acc += d * d; }
		return sqrt(acc / N) / mean;
	}
	case OP_FILLIN: case OP_TRIMFILL:    // :78-107: track 1's value, or its default where it is absent
		r.get(1, x);
		return x;
	case OP_TTEST: {                     // src/setComparisons.c:56-117
		const int n1 = job.n1, n2 = N - job.n1;
		double sum1 = 0, sum2 = 0, sq1 = 0, sq2 = 0;
		for (int i = 0; i < n1; i++) if (r.get(i, x)) { sum1 += x; sq1 += x * x; }
		for (int i = n1; i < N; i++) if (r.get(i, x)) { sum2 += x; sq2 += x * x; }
		double mean1 = sum1 / n1, mean2 = sum2 / n2;
		double var1 = sq1 / n1 - mean1 * mean1, var2 = sq2 / n2 - mean2 * mean2;
		if (var1 + var2 == 0) return NAN;
		double t = (mean1 - mean2) / sqrt(var1 / n1 + var2 / n2);
		if (t < 0) t = -t;
		double nu = (var1 / n1 + var2 / n2) * (var1 / n1 + var2 / n2)
		          / ((var1 * var1) / (n1 * n1 * (n1 - 1)) + (var2 * var2) / (n2 * n2 * (n2 - 1)));
		return 2 * tdist_Q(t, nu);
	}
	}
	return NAN;
}

// ---------------------------------------------------------------------------
// warp-cooperative sorted-row statistics (median, Mann-Whitney U)
// ---------------------------------------------------------------------------
// One warp per segment: lanes gather the N values into a shared row buffer,
// bitonic-sort it (keys = value, set/index for a stable order) and lane 0
// finishes.  Slow (O(N log^2 N) per segment); a dedicated incremental kernel is
// the planned replacement.
struct SortKey { double v; int ord; int pad; };

__device__ __forceinline__ bool key_less(const SortKey &a, const SortKey &b) {
	if (a.v < b.v) return true;
	if (a.v > b.v) return false;
	return a.ord < b.ord;      // stable: glibc qsort is a merge sort at these sizes
}

__device__ __noinline__ double warp_sorted_eval(const Row &r, SortKey *buf, int P /*pow2 >= N*/) {
	const Job &job = *r.job;
	const int N = job.N, lane = threadIdx.x & 31;
	bool nan = false;
	for (int i = lane; i < P; i += 32) {
		SortKey k;
		if (i < N) { double x; r.get(i, x); k.v = x; k.ord = i; nan |= isnan(x); }
		else { k.v = INFINITY; k.ord = INT_MAX; }
		k.pad = 0;
		buf[i] = k;
	}
	nan = __any_sync(FULL, nan);
	__syncwarp();
	if (nan) return NAN;              // both pops return NaN on any NaN input
	for (int k = 2; k <= P; k <<= 1)
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int t = lane; t < P / 2; t += 32) {
				int i = ((t / j) * 2 * j) + (t % j);   // lower index of the pair
				int l = i + j;
				bool up = ((i & k) == 0);
				SortKey a = buf[i], b = buf[l];
				bool sw = up ? key_less(b, a) : key_less(a, b);
				if (sw) { buf[i] = b; buf[l] = a; }
			}
			__syncwarp();
		}
	double res = 0;
	if (lane == 0) {
		if (job.op == OP_MEDIAN) {
			res = buf[N / 2].v;       // upper median, src/reducers.c:809-810
		} else {
			// Mann-Whitney U with the reference's tie bookkeeping, src/setComparisons.c:327-366
			const int n1 = job.n1;
			double U1 = 0;
			int prev = 0, ties = 0, previousTies = 0;
			for (int i = 0; i < N && prev < n1; i++) {
				if (buf[i].ord < n1) {
					U1 += i - prev;
					if (ties) {
						for (int j = i + 1; j < N && buf[j].v == buf[i].v && buf[j].ord >= n1; j++)
							previousTies++;
						U1 -= previousTies / 2.0;
						U1 += (ties - previousTies) / 2.0;
						if (previousTies == ties) previousTies = ties = 0;
					} else {
						for (int j = i + 1; j < N && buf[j].v == buf[i].v; j++)
							if (buf[j].ord >= n1) ties++;
						if (ties) U1 += ties / 2.0;
					}
					prev++;
				}
			}
			if (U1 > job.mwu_mu) res = 2 * erf((job.mwu_mu - U1) / job.mwu_sigma);
			else res = 2 * erf((U1 - job.mwu_mu) / job.mwu_sigma);
		}
	}
	__syncwarp();
	return __shfl_sync(FULL, res, 0);
}

// ---------------------------------------------------------------------------
// block-wide scans (one value per thread)
// ---------------------------------------------------------------------------
// Exclusive scan of an (int, long long) pair per thread; totals returned to all.
__device__ __forceinline__ void block_scan2(Misc *misc, int c, long long a, int &ec, long long &ea, int &tc, long long &ta) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int ic = c; long long ia = a;
	for (int d = 1; d < 32; d <<= 1) {
		int vc = __shfl_up_sync(FULL, ic, d);
		long long va = __shfl_up_sync(FULL, ia, d);
		if (lane >= d) { ic += vc; ia += va; }
	}
	if (lane == 31) { misc->wcov[warp] = ic; misc->wsum[warp] = ia; }
	__syncthreads();
	int bc = 0; long long ba = 0; tc = 0; ta = 0;
#pragma unroll
	for (int k = 0; k < WARPS; k++) {
		int kc = misc->wcov[k]; long long ka = misc->wsum[k];
		if (k < warp) { bc += kc; ba += ka; }
		tc += kc; ta += ka;
	}
	ec = bc + ic - c; ea = ba + ia - a;
}

// Same with a second 64-bit value (sum and sum of squares).
__device__ __forceinline__ void block_scan3(Misc *misc, int c, long long a, long long b, int &ec, long long &ea, long long &eb) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int ic = c; long long ia = a, ib = b;
	for (int d = 1; d < 32; d <<= 1) {
		int vc = __shfl_up_sync(FULL, ic, d);
		long long va = __shfl_up_sync(FULL, ia, d), vb = __shfl_up_sync(FULL, ib, d);
		if (lane >= d) { ic += vc; ia += va; ib += vb; }
	}
	if (lane == 31) { misc->wcov[warp] = ic; misc->wsum[warp] = ia; misc->wsum2[warp] = ib; }
	__syncthreads();
	int bc = 0; long long ba = 0, bb = 0;
#pragma unroll
	for (int k = 0; k < WARPS; k++)
		if (k < warp) { bc += misc->wcov[k]; ba += misc->wsum[k]; bb += misc->wsum2[k]; }
	ec = bc + ic - c; ea = ba + ia - a; eb = bb + ib - b;
}

__device__ __forceinline__ void block_scan1(Misc *misc, int c, int &ec, int &tc) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int ic = c;
	for (int d = 1; d < 32; d <<= 1) { int vc = __shfl_up_sync(FULL, ic, d); if (lane >= d) ic += vc; }
	if (lane == 31) misc->wkeep[warp] = ic;
	__syncthreads();
	int bc = 0; tc = 0;
#pragma unroll
	for (int k = 0; k < WARPS; k++) { int kc = misc->wkeep[k]; if (k < warp) bc += kc; tc += kc; }
	ec = bc + ic - c;
}


// A segment is emitted when enough tracks cover it: any track (gaps are skipped,
// src/multiplexer.c:112-115), all N under `strict` (:123-128), or at least one of
// each set for the set comparisons (src/setComparisons.c:48-54).  thr = 1, N, 0;
// thr = -1: at least one of either set (fillIn over two tracks, src/reducers.c:78-107).
__device__ __forceinline__ bool keep_rule(int thr, int c) {
	const int covA = c & COV_MASK;
	if (thr > 0) return covA >= thr;
	const int covB = (c >> COV_B_SHIFT) & COV_MASK;
	if (thr == 0) return covA > 0 && covB > 0;
	return (covA | covB) != 0;                          // thr < 0: any track of either set (fillIn)
}

// ---------------------------------------------------------------------------
// order statistics by an incremental sweep (min, max, median)
// ---------------------------------------------------------------------------
// The reference sorts (median: qsort, src/reducers.c:809) or scans (min/max,
// :125-235) all N values again for every segment.  Consecutive segments differ
// in the one or two tracks that start or end at the breakpoint between them, so
// each warp keeps the N values of its current segment SORTED in shared memory
// (key = value, track index: the stable order of the reference's merge sort)
// and, walking its share of the window's breakpoints, replaces only the keys of
// the tracks that change: O(N/32) per change instead of O(N log N) per segment.
//   min = arr[0], max = arr[N-1], median = arr[N/2] (upper median, :810).
// Per track the warp remembers the interval that covers / follows the current
// position and the position of its next event; "which tracks change here" is
// one compare per track and a ballot.
struct TrackState { int j; int nextpos; };


// The window's intervals, staged once in shared memory (start/end/value of
// interval j of track i at tbase[i] + j - cur[i]); falls back to global memory
// when a window holds more intervals than the staging area.
struct IvAccess {
	const Job *job;
	const int32_t *cur;
	const int32_t *tbase;      // null: not staged
	const int32_t *ss, *se;
	const double *sv;
	__device__ __forceinline__ int start(int i, long long o, int j) const { return tbase ? ss[tbase[i] + j - cur[i]] : job->start[o + j]; }
	__device__ __forceinline__ int end(int i, long long o, int j) const { return tbase ? se[tbase[i] + j - cur[i]] : job->end[o + j]; }
	__device__ __forceinline__ double value(int i, long long o, int j) const { return tbase ? sv[tbase[i] + j - cur[i]] : job->value[o + j]; }
};

// Position a track at coordinate b: interval index j = first interval with
// end > b inside the window's cursor range, value-or-default, next event.
__device__ __forceinline__ bool walk_locate(const Job &job, const IvAccess &iv, const int32_t *cur, const int32_t *nxt, int i, int b,
                                            double dflt, TrackState &ts, double &x) {
	const long long o = job.off[i];
	const int n = (int) (job.off[i + 1] - o);
	int lo = cur[i], hi = min(nxt[i], n - 1) + 1;         // search [lo, hi): ends ascending
	while (lo < hi) {
		const int mid = (lo + hi) >> 1;
		if (iv.end(i, o, mid) > b) hi = mid; else lo = mid + 1;
	}
	ts.j = lo;
	if (lo < n && lo <= nxt[i]) {
		const int st = iv.start(i, o, lo);
		if (st <= b) { x = iv.value(i, o, lo); ts.nextpos = iv.end(i, o, lo); return true; }
		x = dflt; ts.nextpos = st;
	} else if (lo < n) { x = dflt; ts.nextpos = job.start[o + lo]; }
	else { x = dflt; ts.nextpos = INT_MAX; }
	return false;
}

// One warp: segments [r0, r1) of the window's dense breakpoint list.
__device__ void order_walk(const Job &job, const int32_t *cur, const int32_t *nxt, long long w0, int r0, int r1, int nbp,
                           int next_bp, const uint16_t *bpos, const uint16_t *bkid, long long out_base,
                           SortKey *arr, TrackState *tstate, int P, const IvAccess &iv, double &auc, unsigned long long &slow) {
	const int N = job.N, lane = threadIdx.x & 31;
	const bool minmax = (job.op == OP_MIN || job.op == OP_MAX);
	const int sel = job.op == OP_MIN ? 0 : job.op == OP_MAX ? N - 1 : N / 2;
	const int b0 = (int) (w0 + bpos[r0]);
	int nan_cnt = 0;
	for (int i = lane; i < P; i += 32) {
		SortKey k;
		k.pad = 0;
		if (i < N) {
			// min/max read a literal 0 for an absent track 0, not its default (reducers.c:142-145,209-212)
			const double d = (minmax && i == 0) ? 0.0 : job.dflt[i];
			TrackState ts; double x;
			walk_locate(job, iv, cur, nxt, i, b0, d, ts, x);
			tstate[i] = ts;
			if (isnan(x)) { nan_cnt++; x = INFINITY; }
			k.v = x; k.ord = i;
		} else { k.v = INFINITY; k.ord = INT_MAX; }
		arr[i] = k;
	}
	nan_cnt = __reduce_add_sync(FULL, nan_cnt);
	__syncwarp();
	for (int k = 2; k <= P; k <<= 1)
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int t = lane; t < P / 2; t += 32) {
				const int i = ((t / j) * 2 * j) + (t % j), l = i + j;
				const bool up = ((i & k) == 0);
				const SortKey a = arr[i], b = arr[l];
				if (up ? key_less(b, a) : key_less(a, b)) { arr[i] = b; arr[l] = a; }
			}
			__syncwarp();
		}
	for (int r = r0; r < r1; r++) {
		const int p = (int) (w0 + bpos[r]);
		if (r > r0) {
			// tracks with an event at p: replace their key in the sorted array
			for (int base = 0; base < N; base += 32) {
				const int i = base + lane;
				unsigned bal = __ballot_sync(FULL, i < N && tstate[i].nextpos == p);
				while (bal) {
					const int src = __ffs(bal) - 1;
					bal &= bal - 1;
					double oldx = 0, newx = 0;
					if (lane == src) {
						const long long o = job.off[i];
						const int n = (int) (job.off[i + 1] - o);
						const double d = (minmax && i == 0) ? 0.0 : job.dflt[i];
						TrackState ts = tstate[i];
						// (intervals up to nxt[i] are staged; the sweep never moves past them inside the window)
						const bool was_in = iv.start(i, o, ts.j) < p;      // else its interval starts at p
						oldx = was_in ? iv.value(i, o, ts.j) : d;
						if (was_in) ts.j++;
						const bool have = ts.j < n && ts.j <= nxt[i];
						const int ns = have ? iv.start(i, o, ts.j) : (ts.j < n ? job.start[o + ts.j] : INT_MAX);
						if (have && ns <= p) { newx = iv.value(i, o, ts.j); ts.nextpos = iv.end(i, o, ts.j); }
						else { newx = d; ts.nextpos = ns; }
						tstate[i] = ts;
					}
					oldx = __shfl_sync(FULL, oldx, src);
					newx = __shfl_sync(FULL, newx, src);
					const int t = base + src;
					nan_cnt += (isnan(newx) ? 1 : 0) - (isnan(oldx) ? 1 : 0);
					SortKey ko, kn;
					ko.v = isnan(oldx) ? INFINITY : oldx; ko.ord = t; ko.pad = 0;
					kn.v = isnan(newx) ? INFINITY : newx; kn.ord = t; kn.pad = 0;
					// ranks of the old and the new key among the N keys
					int lo = 0, ln = 0;
					for (int q = lane; q < N; q += 32) {
						const SortKey a = arr[q];
						lo += key_less(a, ko) ? 1 : 0;
						ln += key_less(a, kn) ? 1 : 0;
					}
					__syncwarp();                                    // every lane has read its keys before any is moved
					const int io = __reduce_add_sync(FULL, lo);
					int in = __reduce_add_sync(FULL, ln);
					if (key_less(ko, kn)) in--;                      // the old key itself was counted below the new one
					if (in > io) {
						for (int q = io; q < in; q += 32) {              // close the gap: shift (io, in] down
							SortKey a; const bool act = q + lane < in;
							if (act) a = arr[q + lane + 1];
							__syncwarp();
							if (act) arr[q + lane] = a;
							__syncwarp();
						}
					} else if (in < io) {
						for (int q = io; q > in; q -= 32) {              // open a gap: shift [in, io) up, from the top
							SortKey a; const int idx = q - 1 - lane; const bool act = idx >= in;
							if (act) a = arr[idx];
							__syncwarp();
							if (act) arr[idx + 1] = a;
							__syncwarp();
						}
					}
					if (lane == 0) arr[in] = kn;
					__syncwarp();
				}
			}
		}
		const uint16_t kd = bkid[r];
		if ((kd & BK_KEEP) && lane == 0) {
			const int be = (r + 1 < nbp) ? (int) (w0 + bpos[r + 1]) : next_bp;
			const double v = nan_cnt ? NAN : arr[sel].v;
			const long long oi = out_base + (kd & BK_IDX);
			if (oi < job.out_cap) {
				job.out_start[oi] = p;
				job.out_end[oi] = be;
				job.out_value[oi] = v;
			}
			if (!isnan(v)) auc += (double) (be - p) * v;
			slow++;
		}
	}
}

// Shared-memory carve-up of the sweep kernels' row buffer (host and device agree).
#ifndef WB_SEL_TICKET_AT_END
#define WB_SEL_TICKET_AT_END 1           // per-segment modes (select / walk / row): draw the next window when the current one is done (0: when its emit phase starts)
#endif
#ifndef WB_SEL_MIN_N
#define WB_SEL_MIN_N 24                  // up to this many tracks: a sorted array per warp (order_walk) instead of sorting the window
                                         // (measured, median, kernel ms on a tenth of the genome: 16 tracks 13.3 walk / 18.3 select, 24: 18.6 / 18.8,
                                         // 32: 22.7 / 18.9; 64 tracks 6.2 -> 12.9 Gbp/s, 128: 2.0 -> 8.1)
#endif
constexpr int SEL_PS = 4096;             // bucketed selection: largest universe (64-bit keys), intervals + gaps
struct WalkLayout {
	int P;            // pow2 >= N: keys per warp of the sorted-array sweep
	int P2;           // universe capacity of the rank-bitset sweep (window intervals + N defaults, 16-byte keys)
	int cap;          // staged intervals per window
	int warps;        // warps that run the serial sweeps (each needs a sorted array / a state row of its own)
	size_t off_tstate, off_keys, off_rank, off_stage, off_tbase, off_bits, off_rse, total;
	size_t cnt_bytes; // bucketed selection: room for the bucket counters (over stage | tstate | bits, dead by then)
};
__host__ __device__ inline WalkLayout walk_layout(int N) {
	WalkLayout w;
	w.P = 1; while (w.P < N) w.P <<= 1;
	const bool sel = N > WB_SEL_MIN_N;                         // the window-sorting sweeps are in use
	w.P2 = (N <= 64 && !sel) ? 1024 : 2048;                    // sel: the key area also holds up to SEL_PS 64-bit keys
	// staged intervals: what a 1024-bp window of dense coverage data holds (about two intervals per track and
	// 512 bp, plus the carried-in and the look-ahead one of every track); more than that: order_walk
	w.cap = !sel ? w.P2 - N : N <= 256 ? (SEL_PS - N) / 2 : 7 * N;
	size_t o = 0;
	// more than 256 tracks: the serial sweeps are the fallbacks of the bucketed selection, and half the warps
	// run them -- their rows would otherwise cost the common path a resident CTA
	w.warps = (sel && N > 256) ? WARPS / 2 : WARPS;
	const size_t keys = (size_t) w.P2 * sizeof(SortKey), arrs = (size_t) w.warps * w.P * sizeof(SortKey);
	w.off_keys = o;   o += keys > arrs ? keys : arrs;          // sorted universe (16-byte or 64-bit keys) | per-warp sorted arrays
	// serial sweeps: rank of every element, state of every track per warp | bucketed: (start, end) of every rank
	const size_t serial = (size_t) w.P2 * 2 + (size_t) w.warps * N * sizeof(TrackState), rse = sel ? (size_t) SEL_PS * 4 : 0;
	w.off_rank = o;   w.off_rse = o;  w.off_tstate = o + (size_t) w.P2 * 2;
	o += serial > rse ? serial : rse;
	w.off_tbase = o;  o += (size_t) (N + 4) / 4 * 16;
	w.off_stage = o;  o += (size_t) w.cap * 16 + 16;
	w.off_bits = o;   o += (size_t) WARPS * (w.P2 / 32) * 4 > (size_t) N * 16 ? (size_t) WARPS * (w.P2 / 32) * 4 : (size_t) N * 16;   // | bucketed: cbase, fa, defaults, gcount
	w.cnt_bytes = o - w.off_stage;
	w.total = (o + 15) / 16 * 16;
	return w;
}

// Rank-bitset sweep (min / max / median).  The window's values -- every staged
// interval plus one default per track -- are sorted ONCE per window by the whole
// CTA; an interval's rank in that order is then a fixed number.  A warp keeps the
// N values of its current segment as N set bits over the ranks: a track that
// changes clears one bit and sets another (O(1)), and the k-th smallest value is
// a popcount prefix over the words plus a find-n-th-set-bit.  Ties order by track
// index, as in the reference's stable sort.
__device__ void bitset_walk(const Job &job, const int32_t *cur, const int32_t *nxt, long long w0, int r0, int r1, int nbp,
                            int next_bp, const uint16_t *bpos, const uint16_t *bkid, long long out_base,
                            const SortKey *keys, const uint16_t *rank_of, uint32_t *bits, int nwords, int T,
                            TrackState *tstate, const IvAccess &iv, double &auc, unsigned long long &slow) {
	const int N = job.N, lane = threadIdx.x & 31;
	const bool minmax = (job.op == OP_MIN || job.op == OP_MAX);
	const int sel = job.op == OP_MIN ? 0 : job.op == OP_MAX ? N - 1 : N / 2;
	const int nwl = (nwords + 31) / 32;                 // words owned by each lane, contiguous
	const int b0 = (int) (w0 + bpos[r0]);
	for (int q = lane; q < nwords; q += 32) bits[q] = 0;
	__syncwarp();
	int nan_cnt = 0;
	for (int i = lane; i < N; i += 32) {
		const double d = (minmax && i == 0) ? 0.0 : job.dflt[i];   // reducers.c:142-145,209-212
		TrackState ts; double x;
		const bool in = walk_locate(job, iv, cur, nxt, i, b0, d, ts, x);
		tstate[i] = ts;
		if (isnan(x)) nan_cnt++;
		const int rk = rank_of[in ? iv.tbase[i] + ts.j - cur[i] : T + i];
		atomicOr(&bits[rk >> 5], 1u << (rk & 31));
	}
	nan_cnt = __reduce_add_sync(FULL, nan_cnt);
	__syncwarp();
	for (int r = r0; r < r1; r++) {
		const int p = (int) (w0 + bpos[r]);
		if (r > r0) {
			for (int base = 0; base < N; base += 32) {
				const int i = base + lane;
				unsigned bal = __ballot_sync(FULL, i < N && tstate[i].nextpos == p);
				while (bal) {
					const int src = __ffs(bal) - 1;
					bal &= bal - 1;
					int dn = 0;
					if (lane == src) {
						const long long o = job.off[i];
						const int n = (int) (job.off[i + 1] - o);
						const double d = (minmax && i == 0) ? 0.0 : job.dflt[i];
						TrackState ts = tstate[i];
						const bool was_in = iv.start(i, o, ts.j) < p;      // else its interval starts at p
						const int uo = was_in ? iv.tbase[i] + ts.j - cur[i] : T + i;
						const bool oldnan = isnan(was_in ? iv.value(i, o, ts.j) : d);
						if (was_in) ts.j++;
						const bool have = ts.j < n && ts.j <= nxt[i];
						const int ns = have ? iv.start(i, o, ts.j) : (ts.j < n ? job.start[o + ts.j] : INT_MAX);
						int un; bool newnan;
						if (have && ns <= p) { un = iv.tbase[i] + ts.j - cur[i]; newnan = isnan(iv.value(i, o, ts.j)); ts.nextpos = iv.end(i, o, ts.j); }
						else { un = T + i; newnan = isnan(d); ts.nextpos = ns; }
						tstate[i] = ts;
						const int ro = rank_of[uo], rn = rank_of[un];
						bits[ro >> 5] &= ~(1u << (ro & 31));
						bits[rn >> 5] |= 1u << (rn & 31);
						dn = (newnan ? 1 : 0) - (oldnan ? 1 : 0);
					}
					nan_cnt += __shfl_sync(FULL, dn, src);
					__syncwarp();
				}
			}
		}
		const uint16_t kd = bkid[r];
		if (kd & BK_KEEP) {
			// k-th set bit: popcount prefix over the lanes' words, then find-n-th inside one word
			int cnt = 0;
			for (int q = 0; q < nwl; q++) { const int wi = lane * nwl + q; if (wi < nwords) cnt += __popc(bits[wi]); }
			int inc = cnt;
			for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += v; }
			const unsigned hit = __ballot_sync(FULL, inc > sel);
			const int Lh = __ffs(hit) - 1;
			double v = 0;
			if (lane == Lh) {
				int k = sel - (inc - cnt);
				for (int q = 0; q < nwl && lane * nwl + q < nwords; q++) {
					const uint32_t wd = bits[lane * nwl + q];
					const int c = __popc(wd);
					if (k < c) { v = keys[(lane * nwl + q) * 32 + __fns(wd, 0, k + 1)].v; break; }
					k -= c;
				}
			}
			v = __shfl_sync(FULL, v, Lh);
			if (nan_cnt) v = NAN;
			if (lane == 0) {
				const int be = (r + 1 < nbp) ? (int) (w0 + bpos[r + 1]) : next_bp;
				const long long oi = out_base + (kd & BK_IDX);
				if (oi < job.out_cap) {
					job.out_start[oi] = p;
					job.out_end[oi] = be;
					job.out_value[oi] = v;
				}
				if (!isnan(v)) auc += (double) (be - p) * v;
				slow++;
			}
		}
	}
}

// Bucketed selection (min / max / median), the whole CTA at once, every breakpoint on its own.
//
// Universe of a window: its staged intervals plus, per track, the GAPS between them (and before
// the first / after the last one), each carrying the track's default -- so that at every position
// exactly one element per track is in play and every element is a plain [start, end).  When all
// values are float32-exact (what a BigWig / bedGraph reader delivers; no NaN, no -0) an element is
// one 64-bit key  order-preserving float bits | track | element index  and the universe is sorted
// once per window with a register-blocked bitonic network.  The ranks are cut into SEL_NB buckets,
// and a byte counter per (position, bucket) says how many elements of the bucket are in play
// there: +1 at an element's start, -1 at its end, then a prefix sum along the positions.  Four
// byte counters share a 32-bit word: the shared atomics add or subtract a shifted one, and since
// every true count lies in [0, ranks per bucket] <= 255 the 32-bit prefix sums decode byte by
// byte.  A kept breakpoint finds its bucket from the SEL_NB counters of its position and then the
// k-th element in play among the bucket's ranks, in rank order.  Ties order by track index, as in
// the reference's stable sort (src/reducers.c:140-207: qsort of the N current values).
constexpr int SEL_NB = 32;                                   // buckets (8 words per position)
// word c of position p.  The prefix pass walks runs of 2^lsh positions, four runs per warp: the rows of positions
// one run apart are rotated into different bank groups (lsh >= 2).
__device__ __forceinline__ int sel_word(int p, int c, int lsh) { return (p ^ ((p >> lsh) & 3)) * (SEL_NB / 4) + (c ^ (((p >> 2) & 1) << 2)); }

// float32 bits <-> unsigned keys in the same order
__device__ __forceinline__ uint32_t f32_key(float f) { const uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float key_f32(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
__device__ __forceinline__ bool f32_exact(double v) { return (double) (float) v == v && !(v == 0.0 && signbit(v)); }   // NaN fails too

// Bitonic sort of Ps 64-bit keys in shared memory (Ps a power of two >= E), ascending by their
// high words, by the whole CTA.  A (virtual) thread holds E = 2^LE keys in registers and runs up
// to LE levels of the network on them between two trips to shared memory: the keys whose
// indices differ in bits b0 .. b0+LE-1 only.  Index -> address is xor-swizzled so that every
// such gather is spread over the banks.
// (le = log2 of the keys per thread the array is sorted with: conflict-free gathers for every trip, checked by enumeration)
__device__ __forceinline__ int sort_phys(int idx, int le) { return idx ^ ((idx >> le) & 15); }
// compare-exchange on the high words (the value): which of two equal values comes first is immaterial
template <bool UP>
__device__ __forceinline__ void sort_ce(unsigned long long &a, unsigned long long &b) {
	const uint32_t ha = (uint32_t) (a >> 32), hb = (uint32_t) (b >> 32);
	const bool sw = UP ? (ha > hb) : (ha < hb);
	const unsigned long long x = sw ? b : a, y = sw ? a : b;
	a = x; b = y;
}
template <bool UP, int LE>
__device__ __forceinline__ void sort_levels(unsigned long long (&r)[1 << LE], int b0, int b) {
#pragma unroll
	for (int sft = LE - 1; sft >= 0; sft--)
		if (b0 + sft <= b) {
#pragma unroll
			for (int m = 0; m < (1 << LE); m++)
				if ((m & (1 << sft)) == 0) sort_ce<UP>(r[m], r[m | (1 << sft)]);
		}
}
template <int LE>
__device__ void bitonic_u64(unsigned long long *keys, int Ps) {
	constexpr int E = 1 << LE;
	const int tid = threadIdx.x, nvt = Ps >> LE;
	// k = 2 .. E: E consecutive keys per thread
	for (int vt = tid; vt < nvt; vt += THREADS) {
		unsigned long long r[E];
		const int base = vt << LE;
#pragma unroll
		for (int m = 0; m < E; m++) r[m] = keys[sort_phys(base | m, LE)];
#pragma unroll
		for (int k = 2; k <= E; k <<= 1)
#pragma unroll
			for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
				for (int m = 0; m < E; m++)
					if ((m & j) == 0) {
						// k < E: the direction is a compile-time function of m; k = E: of the thread
						if (k < E) { if ((m & k) == 0) sort_ce<true>(r[m], r[m | j]); else sort_ce<false>(r[m], r[m | j]); }
						else if ((base & E) == 0) sort_ce<true>(r[m], r[m | j]); else sort_ce<false>(r[m], r[m | j]);
					}
#pragma unroll
		for (int m = 0; m < E; m++) keys[sort_phys(base | m, LE)] = r[m];
	}
	__syncthreads();
	for (int lk = LE + 1; (1 << lk) <= Ps; lk++) {
		int b = lk - 1;                                       // levels b .. 0 of this k, LE at a time
		while (b >= 0) {
			const int b0 = max(b - (LE - 1), 0);
			for (int vt = tid; vt < nvt; vt += THREADS) {
				unsigned long long r[E];
				const int base = ((vt >> b0) << (b0 + LE)) | (vt & ((1 << b0) - 1));
				const bool up = (base & (1 << lk)) == 0;          // bit lk lies above the levels of this trip
#pragma unroll
				for (int m = 0; m < E; m++) r[m] = keys[sort_phys(base | (m << b0), LE)];
				if (up) sort_levels<true, LE>(r, b0, b); else sort_levels<false, LE>(r, b0, b);
#pragma unroll
				for (int m = 0; m < E; m++) keys[sort_phys(base | (m << b0), LE)] = r[m];
			}
			__syncthreads();
			b = b0 - 1;
		}
	}
}

// One range [h0, h0 + ws) of the window's positions.  val: value of every rank; rse: (start, end) of
// every rank, window-relative and clipped to the range, (W, W) for the padding; cntw: ws x SEL_NB
// byte counters, cleared.
__device__ void bucket_select(const Job &job, Misc *misc, long long w0, int W, int h0, int ws, int nbp, int next_bp,
                              const uint16_t *bpos, const uint16_t *bkid, long long out_base, const float *val,
                              const ushort2 *rse, uint32_t *cntw, int bs, double &auc, unsigned long long &slow WB_DIAG_PARAM) {
	constexpr int NBW = SEL_NB / 4;
	const int N = job.N, tid = threadIdx.x;
	const int sel = job.op == OP_MIN ? 0 : job.op == OP_MAX ? N - 1 : N / 2;
	const int Ps = bs * SEL_NB;                               // bs ranks per bucket (a multiple of four, <= 255)
	constexpr int RUNS = THREADS / NBW;                       // prefix pass: runs of max(1, ws / RUNS) positions
	int lsh = 2;
	while ((RUNS << lsh) < ws) lsh++;
	// events, one rank per thread and step
	for (int rk = tid; rk < Ps; rk += THREADS) {
		const ushort2 se = rse[rk];
		if ((int) se.x >= W) continue;
		const int b = rk / bs;
		atomicAdd(&cntw[sel_word(se.x - h0, b >> 2, lsh)], 1u << ((b & 3) << 3));
		if ((int) se.y < h0 + ws) atomicSub(&cntw[sel_word(se.y - h0, b >> 2, lsh)], 1u << ((b & 3) << 3));
	}
	__syncthreads();
	// prefix sums along the positions: NBW word columns x runs of positions
	{
		const int len = max(1, ws / RUNS), runs = ws / len;     // ws is a power of two
		const int c = tid % NBW, run = tid / NBW, pa = run * len;
		uint32_t tot = 0;
		if (run < runs) {
#pragma unroll 8
			for (int p = pa; p < pa + len; p++) tot += cntw[sel_word(p, c, lsh)];
		}
		uint32_t *rt = (uint32_t *) ((unsigned char *) misc + MISC_SCRATCH);   // RUNS x NBW run totals
		rt[run * NBW + c] = tot;
		__syncthreads();
		if (run < runs) {
			uint32_t acc = 0;
			for (int k = 0; k < run; k++) acc += rt[k * NBW + c];
#pragma unroll 8
			for (int p = pa; p < pa + len; p++) { const int x = sel_word(p, c, lsh); acc += cntw[x]; cntw[x] = acc; }
		}
	}
	__syncthreads();
	WB_PHASE(10);
	// selection, one kept breakpoint per thread
	for (int r = tid; r < nbp; r += THREADS) {
		const uint16_t kd = bkid[r];
		const int p = bpos[r];
		if (!(kd & BK_KEEP) || p < h0 || p >= h0 + ws) continue;
		int k = sel, bucket = SEL_NB - 1;
		{
			const uint4 a = *(const uint4 *) (cntw + sel_word(p - h0, 0, lsh)), b = *(const uint4 *) (cntw + sel_word(p - h0, 4, lsh));
			const uint32_t wds[NBW] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
			bool found = false;
#pragma unroll
			for (int q = 0; q < NBW; q++) {
				// bytes of one word: whole-word total first, then the byte inside
				const uint32_t wd = wds[q];
				const int tot = (int) ((wd & 0xff) + ((wd >> 8) & 0xff) + ((wd >> 16) & 0xff) + (wd >> 24));
				if (!found) {
					if (k < tot) {
						found = true;
						bool inside = false;
#pragma unroll
						for (int j = 0; j < 4; j++) {
							const int cb = (wd >> (8 * j)) & 0xff;
							if (!inside) { if (k < cb) { bucket = 4 * q + j; inside = true; } else k -= cb; }
						}
					} else k -= tot;
				}
			}
		}
		// the k-th element in play among the bucket's ranks, in rank order
		double v = NAN;
		const int ra = bucket * bs, rb = ra + bs;
#pragma unroll 2
		for (int rk = ra; rk < rb; rk += 4) {                  // four ranks per load
			const uint4 q = *(const uint4 *) (rse + rk);
			const uint32_t e[4] = {q.x, q.y, q.z, q.w};
			const unsigned up = (unsigned) p;
			int in[4], cnt = 0;
#pragma unroll
			for (int j = 0; j < 4; j++) { in[j] = ((e[j] & 0xffffu) <= up && up < (e[j] >> 16)) ? 1 : 0; cnt += in[j]; }
			if (k < cnt) {
				int hit = 0;
#pragma unroll
				for (int j = 0; j < 4; j++) { if (in[j] && k == 0) hit = j; k -= in[j]; }
				v = (double) val[rk + hit];
				break;
			}
			k -= cnt;
		}
		const int bs = (int) (w0 + p);
		const int be = (r + 1 < nbp) ? (int) (w0 + bpos[r + 1]) : next_bp;
		const long long oi = out_base + (kd & BK_IDX);
		if (oi < job.out_cap) {
			job.out_start[oi] = bs;
			job.out_end[oi] = be;
			job.out_value[oi] = v;
		}
		auc += (double) (be - bs) * v;                         // no NaN on this path
		slow++;
	}
}

// One fixed-point delta into the limb arrays of one accumulator (native 32-bit
// shared atomics; the limbs are recombined after the scan).
template <int VLOG, int LIMBS>
__device__ __forceinline__ void add_limbs(uint32_t *limb, int ph, long long d, int lw) {
	constexpr int W = THREADS << VLOG;
	if (LIMBS == 1) { atomicAdd(&limb[ph], (uint32_t) d); return; }     // the delta sums fit 32 bits
	atomicAdd(&limb[ph], (uint32_t) (d & ((1ll << lw) - 1)));
	if (LIMBS == 2) atomicAdd(&limb[W + ph], (uint32_t) (d >> lw));
	else {
		atomicAdd(&limb[W + ph], (uint32_t) ((d >> lw) & ((1ll << lw) - 1)));
		atomicAdd(&limb[2 * W + ph], (uint32_t) (d >> (2 * lw)));
	}
}

// First accumulator (the plain sum) of a position: one 32-bit limb when the job says every prefix fits
// 31 bits (moments path with a narrow value range: one shared atomic per event instead of LIMBS).
template <int VLOG, int LIMBS>
__device__ __forceinline__ void add_sum(const Job &job, uint32_t *limb, int ph, long long d, int lw) {
	if (job.sum1) atomicAdd(&limb[ph], (uint32_t) d);
	else add_limbs<VLOG, LIMBS>(limb, ph, d, lw);
}

// This thread's V positions of one accumulator -> running prefix in registers.
template <int VLOG, int LIMBS>
__device__ __forceinline__ long long load_prefix(const uint32_t *limb, int tid, int lw, long long *run) {
	constexpr int V = 1 << VLOG, W = THREADS << VLOG;
	const int p0 = tid << VLOG;
	uint32_t l0[V], l1[V], l2[V];
	if (VLOG >= 2) {
		constexpr int NV = V / 4, LV = VLOG - 2;
		const int swz = (tid >> (3 - LV)) & (NV - 1);
#pragma unroll
		for (int q = 0; q < NV; q++) {
			const int pq = p0 + ((q ^ swz) << 2);
			const uint4 y = *(const uint4 *) (limb + pq);
			l0[4 * q] = y.x; l0[4 * q + 1] = y.y; l0[4 * q + 2] = y.z; l0[4 * q + 3] = y.w;
			if (LIMBS >= 2) {
				const uint4 z = *(const uint4 *) (limb + W + pq);
				l1[4 * q] = z.x; l1[4 * q + 1] = z.y; l1[4 * q + 2] = z.z; l1[4 * q + 3] = z.w;
			}
			if (LIMBS == 3) {
				const uint4 u = *(const uint4 *) (limb + 2 * W + pq);
				l2[4 * q] = u.x; l2[4 * q + 1] = u.y; l2[4 * q + 2] = u.z; l2[4 * q + 3] = u.w;
			}
		}
	} else {
#pragma unroll
		for (int q = 0; q < V; q++) { l0[q] = limb[p0 + q]; if (LIMBS >= 2) l1[q] = limb[W + p0 + q]; if (LIMBS == 3) l2[q] = limb[2 * W + p0 + q]; }
	}
	long long a = 0;
#pragma unroll
	for (int q = 0; q < V; q++) {
		if (LIMBS == 1) a += (long long) (int32_t) l0[q];
		else if (LIMBS == 2) a += (long long) (unsigned long long) l0[q] + ((long long) (int32_t) l1[q] << lw);
		else a += (long long) (unsigned long long) l0[q] + ((long long) (unsigned long long) l1[q] << lw)
		        + ((long long) (int32_t) l2[q] << (2 * lw));
		run[q] = a;
	}
	return a;
}


enum { MODE_FAST = 1, MODE_ROW = 2, MODE_MOMENTS = 3, MODE_WALK = 4, MODE_SELECT = 5 };

// ---------------------------------------------------------------------------
// events, vectorised (fixed-point modes)
// ---------------------------------------------------------------------------
// 8-lane groups walk one track each; a lane owns FOUR consecutive intervals per
// step, fetched with 128-bit loads from a 16-byte aligned index grid, so one
// step covers 32 intervals of the track -- normally its whole share of the
// window.  The loads of the next round of tracks are issued before the current
// round is consumed (register double buffering): one exposed memory latency per
// window instead of one per 8 intervals.
// Intervals of a track are sorted and disjoint, so no cross-lane bookkeeping is
// needed: an interval adds +delta at its start if that lies in the window (or
// feeds the carried-in state if it began before it), -delta at its end if that
// lies in the window, and otherwise offers its first coordinate at or beyond the
// window end as the window's closing breakpoint.
// Ask for the first chunk of every track of the window ahead of the events phase
// (one request per 16 bytes a lane will load): the lines travel while the
// position arrays are cleared.
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void prefetch_window(const Job &job, const int32_t *cur, const long long *toff, const int32_t *tn) {
	const int gid = threadIdx.x >> 3, gl = threadIdx.x & 7;
	for (int i = gid; i < job.N; i += THREADS / 8) {
		const long long o = toff[i], first = o + cur[i];
		const long long g = (first & ~3ll) + 4 * gl;
		if (g < o + tn[i] && g + 4 <= job.I) {
#if WB_PREF == 2
			prefetch_l1(job.start + g); prefetch_l1(job.end + g); prefetch_l1(job.value + g); prefetch_l1(job.value + g + 2);
#else
			prefetch_l2(job.start + g); prefetch_l2(job.end + g); prefetch_l2(job.value + g); prefetch_l2(job.value + g + 2);
#endif
		}
	}
}

struct Chunk {
	int4 s, e;
	double v0, v1, v2, v3;
	int lo, hi;          // slots lo <= x < hi hold intervals of this track at or after its cursor
};

template <int MODE, int VLOG, int LIMBS>
__device__ __forceinline__ void events_vec(const Job &job, unsigned char *smem, const Layout &L, const int32_t *cur,
                                           long long w0, int lw, double inv_scale, double lim) {
	constexpr int W = THREADS << VLOG;
	constexpr int NACC = MODE == MODE_MOMENTS ? 2 : 1;
	constexpr int G = THREADS / 8;
	Misc *misc = (Misc *) smem;
	const int32_t *tn = (const int32_t *) (smem + L.off_tn);
	const long long *toff = (const long long *) (smem + L.off_toff);
	const long long *tmd = (const long long *) (smem + L.off_tmd);
	int32_t *cov = (int32_t *) (smem + L.off_cov);
	uint32_t *limb = (uint32_t *) (smem + L.off_limb);
	unsigned char *ev = smem + L.off_ev;
	const int tid = threadIdx.x, lane = tid & 31, gid = tid >> 3, gl = tid & 7, gsh = lane & 24;
	const int N = job.N;
	const int w0i = (int) w0;
	int in_cov = 0, my_next = INT_MAX;
	long long in_val = 0, in_val2 = 0;

	auto load = [&](int i, int k, Chunk &c) {
		c.lo = c.hi = 0;
		if (i >= N) return;
		const long long o = toff[i], first = o + cur[i];
		const long long g = (first & ~3ll) + 32 * k + 4 * gl;      // 16-byte aligned index grid
		c.lo = (int) max(first - g, 0ll);
		c.hi = (int) min(o + tn[i] - g, 4ll);
		if (c.hi <= c.lo) { c.lo = c.hi = 0; return; }
		if (g + 4 <= job.I) {
			c.s = *(const int4 *) (job.start + g);
			c.e = *(const int4 *) (job.end + g);
			const double2 a = *(const double2 *) (job.value + g), b = *(const double2 *) (job.value + g + 2);
			c.v0 = a.x; c.v1 = a.y; c.v2 = b.x; c.v3 = b.y;
		} else {                                                   // the last few intervals of the whole array
			c.s = make_int4(0, 0, 0, 0); c.e = c.s; c.v0 = c.v1 = c.v2 = c.v3 = 0;
			if (c.lo <= 0 && 0 < c.hi) { c.s.x = job.start[g]; c.e.x = job.end[g]; c.v0 = job.value[g]; }
			if (c.lo <= 1 && 1 < c.hi) { c.s.y = job.start[g + 1]; c.e.y = job.end[g + 1]; c.v1 = job.value[g + 1]; }
			if (c.lo <= 2 && 2 < c.hi) { c.s.z = job.start[g + 2]; c.e.z = job.end[g + 2]; c.v2 = job.value[g + 2]; }
		}
	};
	auto one = [&](bool valid, int s, int e, double val, long long md) {
		if (!valid) return;
		const unsigned xs = (unsigned) (s - w0i), xe = (unsigned) (e - w0i);
		const bool start_in = xs < (unsigned) W, end_in = xe < (unsigned) W;
		if (!start_in && s >= w0i) { my_next = min(my_next, s); return; }   // begins at or beyond the window end
		int cd = 1;
		long long d = 0, d2 = 0, m;
		if (to_fixed(val, inv_scale, lim, m) && (NACC == 1 || (double) (float) val == val)) {
			d = m - md;                                // moments: the reference rounds values through float
			if (NACC == 2) d2 = m * m;
		} else cd += (1 << COV_P_SHIFT);
		if (start_in) {
			const int ph = phys<VLOG>((int) xs);
			atomicAdd(&cov[ph], cd);
			ev[xs] = 1;
			if (d != 0) {
				add_sum<VLOG, LIMBS>(job, limb, ph, d, lw);
				if (NACC == 2) add_limbs<VLOG, LIMBS>(limb + LIMBS * W, ph, d2, lw);
			}
		} else { in_cov += cd; in_val += d; in_val2 += d2; }              // began before the window
		if (end_in) {
			const int ph = phys<VLOG>((int) xe);
			atomicAdd(&cov[ph], -cd);
			ev[xe] = 1;
			if (d != 0) {
				add_sum<VLOG, LIMBS>(job, limb, ph, -d, lw);
				if (NACC == 2) add_limbs<VLOG, LIMBS>(limb + LIMBS * W, ph, -d2, lw);
			}
		} else my_next = min(my_next, e);
	};
	// returns true when the track may have further intervals in the window
	auto apply = [&](const Chunk &c, long long md) -> bool {
		one(c.lo <= 0 && 0 < c.hi, c.s.x, c.e.x, c.v0, md);
		one(c.lo <= 1 && 1 < c.hi, c.s.y, c.e.y, c.v1, md);
		one(c.lo <= 2 && 2 < c.hi, c.s.z, c.e.z, c.v2, md);
		one(c.lo <= 3 && 3 < c.hi, c.s.w, c.e.w, c.v3, md);
		const bool last_in = c.hi == 4 && (unsigned) (c.e.w - w0i) < (unsigned) W;
		return __shfl_sync(FULL, last_in, gsh + 7);
	};

#if WB_PREF == 1
	// register double buffering: the next round's loads fly while this round is consumed
	Chunk c;
	load(gid, 0, c);
	for (int rb = 0; rb < N; rb += G) {
		const int i = rb + gid;
		Chunk nx;
		load(i + G, 0, nx);
		const long long md = (MODE == MODE_FAST && i < N) ? tmd[i] : 0;
		bool more = apply(c, md);
		for (int k = 1; __any_sync(FULL, more); k++) {   // rare: more than ~30 intervals of one track in the window
			load(more ? i : N, k, c);
			more = apply(c, md) && more;
		}
		c = nx;
	}
#else
	for (int rb = 0; rb < N; rb += G) {
		const int i = rb + gid;
		const long long md = (MODE == MODE_FAST && i < N) ? tmd[i] : 0;
		bool more = true;
		for (int k = 0; k == 0 || __any_sync(FULL, more); k++) {   // k > 0 is rare: more than ~30 intervals of one track in the window
			Chunk c;
			load(more ? i : N, k, c);
			more = apply(c, md) && more;
		}
	}
#endif
	// carried-in state and the look-ahead breakpoint: one shared atomic per warp
	my_next = __reduce_min_sync(FULL, my_next);
	in_cov = __reduce_add_sync(FULL, in_cov);
	for (int d = 16; d > 0; d >>= 1) in_val += __shfl_down_sync(FULL, in_val, d);
	if (NACC == 2) for (int d = 16; d > 0; d >>= 1) in_val2 += __shfl_down_sync(FULL, in_val2, d);
	if (lane == 0) {
		if (my_next != INT_MAX) atomicMin(&misc->next_bp, my_next);
		if (in_cov) atomicAdd(&misc->base_cov, in_cov);
		if (in_val) atomicAdd(&misc->base_val, (unsigned long long) in_val);
		if (NACC == 2 && in_val2) atomicAdd(&misc->base_val2, (unsigned long long) in_val2);
	}
}

// ---------------------------------------------------------------------------
// the window pipeline
// ---------------------------------------------------------------------------
// EMIT = false: count the window's segments (and chain the cursors).
// EMIT = true : same sweep with values, then write the segments.
//
//  1. events : 8-lane groups walk one track each from its cursor; every interval
//              adds +delta at its start position and -delta at its end position
//              (native 32-bit shared atomics: packed cover counters and the
//              fixed-point value in limbs) and flags both positions.  Intervals
//              that began before the window feed the carried-in state.
//  2. scan   : each thread pulls its V positions into registers (128-bit loads),
//              prefix-sums them, and one block scan stitches the threads: the
//              state after every position = cover set + exact value sum of the
//              segment that starts there.
//  3. compact: flagged positions (breakpoints) are written as a dense list over
//              the now-dead position arrays; kept ones carry their output rank.
//  4. emit   : one breakpoint per thread, coalesced 16 B/segment stores.

template <int MODE, int VLOG, int LIMBS>
__device__ void process_window(const Job &job, int w, unsigned char *smem, const Layout &L, int32_t *cur,
                               int lw, double inv_scale, double lim, double scale) {
	constexpr int V = 1 << VLOG, W = THREADS << VLOG;
	Misc *misc = (Misc *) smem;
	int32_t *nxt = (int32_t *) (smem + L.off_nxt);
	const int32_t *tn = (const int32_t *) (smem + L.off_tn);
	const long long *toff = (const long long *) (smem + L.off_toff);
	const long long *tmd = (const long long *) (smem + L.off_tmd);
	int32_t *cov = (int32_t *) (smem + L.off_cov);
	uint32_t *limb = (uint32_t *) (smem + L.off_limb);
	unsigned char *ev = smem + L.off_ev;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = job.N;
	const long long w0 = job.p0 + ((long long) w << job.wlog);
	const long long w1 = w0 + W;
	const bool setop = (job.op == OP_TTEST || job.op == OP_MWU);
	const int thr = setop ? 0 : job.op == OP_FILLIN ? -1 : (job.strict ? N : 1);   // trimFill: set A is track 0 alone, thr 1
	constexpr bool fast = MODE == MODE_FAST || MODE == MODE_MOMENTS;   // fixed-point value path
	constexpr int NACC = MODE == MODE_MOMENTS ? 2 : 1;                 // accumulators: sum (and sum of squares)

#ifdef WB_DIAG_PHASES
	long long diag_t = clock64();
#endif
	if constexpr (fast && WB_PREF >= 2) { if (job.vec) prefetch_window(job, cur, toff, tn); }
	// ---- clear the position arrays -------------------------------------------
	{
		const uint4 z = make_uint4(0, 0, 0, 0);
		uint4 *zc = (uint4 *) cov;
		constexpr int n16 = (fast ? (1 + NACC * LIMBS) : 1) * W / 4;     // cov and limbs are contiguous
		for (int t = tid; t < n16; t += THREADS) zc[t] = z;
		uint4 *ze = (uint4 *) ev;
		for (int t = tid; t < W / 16; t += THREADS) ze[t] = z;
		if (tid == 0) { misc->next_bp = INT_MAX; misc->base_cov = 0; misc->base_val = 0; misc->base_val2 = 0; }
	}
#if WB_LATETICKET
	__pipeline_wait_prior(0);              // this window's cursor row, requested before the clear
#endif
	__syncthreads();
	WB_PHASE(0);

	// MODE_WALK / MODE_SELECT: the window's intervals staged in shared memory -- start / end / value of interval j of
	// track i at tbase[i] + j - cur[i] -- for the sweeps' chains of dependent reads.  One staged slot per thread and
	// step, four steps in flight; the track of every slot is written down first (over the key area, free until
	// the keys are built).  with_events: the cover events of the window are taken from the staged intervals on the
	// way (MODE_SELECT: instead of the track-by-track walk below; needs nxt = cursor row w + 1).
	int st_T = 0;
	bool st_done = false, st_ok = false, st_inexact = false;
	auto stage_window = [&](bool with_events) {
		const WalkLayout wl = walk_layout(N);
		unsigned char *rb = smem + L.off_rows;
		double *sv = (double *) (rb + wl.off_stage);
		int32_t *ss = (int32_t *) (sv + wl.cap), *se = ss + wl.cap;
		int32_t *tbase = (int32_t *) (rb + wl.off_tbase);
		int run = 0;
		for (int base = 0; base < N; base += THREADS) {
			const int i = base + tid;
			const int len = i < N ? max(0, min(nxt[i], tn[i] - 1) - cur[i] + 1) : 0;
			int ec, tc;
			block_scan1(misc, len, ec, tc);
			if (i < N) tbase[i] = run + ec;
			run += tc;
			__syncthreads();
		}
		st_T = run; st_done = true; st_ok = run <= wl.cap; st_inexact = false;
		if (!st_ok) return;
		const int T = run;
		const bool minmax = (job.op == OP_MIN || job.op == OP_MAX);
		uint16_t *owner = (uint16_t *) (rb + wl.off_keys);
		for (int i = tid; i < N; i += THREADS) {
			const int b = tbase[i], len = max(0, min(nxt[i], tn[i] - 1) - cur[i] + 1);
			for (int m = 0; m < len; m++) owner[b + m] = (uint16_t) i;
		}
		__syncthreads();
		int in_cov = 0, my_next = INT_MAX;
		bool inexact = false;
		for (int u0 = 0; u0 < T; u0 += 4 * THREADS) {
			long long g[4]; bool ok[4]; int tr[4];
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const int u = u0 + q * THREADS + tid;
				ok[q] = u < T;
				tr[q] = ok[q] ? owner[u] : 0;
				g[q] = ok[q] ? toff[tr[q]] + cur[tr[q]] + (u - tbase[tr[q]]) : 0;
			}
			int s4[4], e4[4]; double v4[4];
#pragma unroll
			for (int q = 0; q < 4; q++) if (ok[q]) { s4[q] = job.start[g[q]]; e4[q] = job.end[g[q]]; v4[q] = job.value[g[q]]; }
#pragma unroll
			for (int q = 0; q < 4; q++) if (ok[q]) {
				const int u = u0 + q * THREADS + tid;
				ss[u] = s4[q]; se[u] = e4[q]; sv[u] = v4[q];
				inexact |= !f32_exact(v4[q]);
				if (with_events) {
					// same rules as the walk below: a start inside the window, an end inside the window, the carried-in
					// cover, and the first breakpoint beyond the window from the interval that leaves it
					const long long sq = s4[q], eq = e4[q];
					const int f = (tr[q] < job.n1) ? 1 : (1 << COV_B_SHIFT);
					if (eq >= w1) my_next = min(my_next, sq < w1 ? e4[q] : s4[q]);
					if (sq < w1) {
						if (sq >= w0) { const int x = (int) (sq - w0); atomicAdd(&cov[phys<VLOG>(x)], f); ev[x] = 1; }
						else in_cov += f;
						if (eq < w1) { const int x = (int) (eq - w0); atomicAdd(&cov[phys<VLOG>(x)], -f); ev[x] = 1; }
					}
				}
			}
		}
		if (MODE == MODE_SELECT) {
			float *dfl = (float *) (rb + wl.off_bits) + 2 * N;    // bucketed selection: the tracks' defaults as float32
			for (int i = tid; i < N; i += THREADS) {
				const double d = (minmax && i == 0) ? 0.0 : job.dflt[i];     // reducers.c:142-145,209-212
				inexact |= !f32_exact(d);
				dfl[i] = (float) d;
			}
		}
		st_inexact = inexact;
		if (with_events) {
			my_next = __reduce_min_sync(FULL, my_next);
			in_cov = __reduce_add_sync(FULL, in_cov);
			if (lane == 0) {
				if (my_next != INT_MAX) atomicMin(&misc->next_bp, my_next);
				if (in_cov) atomicAdd(&misc->base_cov, in_cov);
			}
		}
	};

	// ---- 1. events --------------------------------------------------------------
	bool vec_done = false;
	if constexpr (MODE == MODE_SELECT) {
		// the track of window w leaves it at cursor[w + 1] (the first interval whose end reaches the next window's
		// start): with that row every track's intervals are known without walking them
		for (int i = tid; i < N; i += THREADS) nxt[i] = job.cursor[(size_t) (w + 1) * job.cstride + i];
		__syncthreads();
		stage_window(true);
		vec_done = st_ok;                                     // more intervals than the staging area holds: the walk below
	}
	if constexpr (fast) {
#ifdef WB_DIAG_SKIP_EVENTS                  // diagnostic builds only (profiles/): time the window skeleton alone
		vec_done = true;
#else
		if (job.vec) { events_vec<MODE, VLOG, LIMBS>(job, smem, L, cur, w0, lw, inv_scale, lim); vec_done = true; }
#endif
	}
	if (!vec_done) {
		const int gid = tid >> 3, gl = tid & 7, gsh = (lane >> 3) << 3;
		int in_cov = 0, my_next = INT_MAX;
		long long in_val = 0, in_val2 = 0;
		for (int rb = 0; rb < N; rb += THREADS / 8) {
			const int i = rb + gid;
			bool done = i >= N;
			long long o = 0, md = 0;
			int n = 0, j0 = 0, f = 0;
			if (!done) {
				o = toff[i]; n = tn[i]; j0 = cur[i];
				f = (i < job.n1) ? 1 : (1 << COV_B_SHIFT);
				if (fast) md = tmd[i];
			}
			while (true) {
				const int j = j0 + gl;
				const bool valid = !done && j < n;
				int s = 0, e = 0;
				double val = 0;
				if (valid) {
					s = job.start[o + j]; e = job.end[o + j];
					if (fast) val = job.value[o + j];
				}
				const bool stop = !valid || (long long) e >= w1;
				const unsigned gb = (__ballot_sync(FULL, stop) >> gsh) & 0xffu;
				const int Lst = gb ? __ffs(gb) - 1 : 8;
				const bool inside = gl < Lst;                        // w0 <= e < w1
				if (valid && (inside || (gl == Lst && (long long) s < w1))) {
					int cd = f;
					long long d = 0, d2 = 0;
					if (fast) {
						long long m;
						if (to_fixed(val, inv_scale, lim, m) && (NACC == 1 || (double) (float) val == val)) {
							d = m - md;                        // moments: the reference rounds values through float
							if (NACC == 2) d2 = m * m;
						}
						else cd += (1 << COV_P_SHIFT);
					}
					if ((long long) s >= w0) {
						const int x = (int) (s - w0), ph = phys<VLOG>(x);
						atomicAdd(&cov[ph], cd);
						ev[x] = 1;
						if (fast && d != 0) {
							add_sum<VLOG, LIMBS>(job, limb, ph, d, lw);
							if (NACC == 2) add_limbs<VLOG, LIMBS>(limb + LIMBS * W, ph, d2, lw);
						}
					} else { in_cov += cd; in_val += d; in_val2 += d2; }   // began before the window
					if (inside) {
						const int x = (int) (e - w0), ph = phys<VLOG>(x);
						atomicAdd(&cov[ph], -cd);
						ev[x] = 1;
						if (fast && d != 0) {
							add_sum<VLOG, LIMBS>(job, limb, ph, -d, lw);
							if (NACC == 2) add_limbs<VLOG, LIMBS>(limb + LIMBS * W, ph, -d2, lw);
						}
					}
				}
				if (!done && gb) {
					// the track leaves the window at interval j0+Lst
					if (gl == Lst && valid) my_next = min(my_next, (long long) s < w1 ? e : s);
					if (gl == 0) nxt[i] = j0 + Lst;
					done = true;
				}
				j0 += 8;
				if (__all_sync(FULL, done)) break;
			}
		}
		// carried-in state and the look-ahead breakpoint: one shared atomic per warp
		my_next = __reduce_min_sync(FULL, my_next);
		in_cov = __reduce_add_sync(FULL, in_cov);
		for (int d = 16; d > 0; d >>= 1) in_val += __shfl_down_sync(FULL, in_val, d);
		if (NACC == 2) for (int d = 16; d > 0; d >>= 1) in_val2 += __shfl_down_sync(FULL, in_val2, d);
		if (lane == 0) {
			if (my_next != INT_MAX) atomicMin(&misc->next_bp, my_next);
			if (in_cov) atomicAdd(&misc->base_cov, in_cov);
			if (fast && in_val) atomicAdd(&misc->base_val, (unsigned long long) in_val);
			if (NACC == 2 && in_val2) atomicAdd(&misc->base_val2, (unsigned long long) in_val2);
		}
	}
	__syncthreads();
	WB_PHASE(1);

	// ---- 2. scan: this thread's V positions, in registers -------------------------
	int c_run[V];
	long long a_run[V], b_run[NACC == 2 ? V : 1];
	unsigned evbits = 0, keepbits = 0;
	{
		const int p0 = tid << VLOG;
		int cv[V];
		if (VLOG >= 2) {
			constexpr int NV = V / 4, LV = VLOG - 2;
			const int swz = (tid >> (3 - LV)) & (NV - 1);
#pragma unroll
			for (int q = 0; q < NV; q++) {
				const int4 x = *(const int4 *) (cov + p0 + ((q ^ swz) << 2));
				cv[4 * q] = x.x; cv[4 * q + 1] = x.y; cv[4 * q + 2] = x.z; cv[4 * q + 3] = x.w;
			}
		} else {
#pragma unroll
			for (int q = 0; q < V; q++) cv[q] = cov[p0 + q];
		}
		if (V >= 4) {
			// flags are 0/1 bytes: gather one bit per byte, four positions per word
#pragma unroll
			for (int q = 0; q < V / 4; q++) {
				const uint32_t wd = ((const uint32_t *) (ev + p0))[q];
				evbits |= (((wd & 0x01010101u) * 0x01020408u) >> 24 & 0xfu) << (4 * q);
			}
		} else {
#pragma unroll
			for (int q = 0; q < V; q++) if (ev[p0 + q]) evbits |= 1u << q;
		}
		int c = 0;
#pragma unroll
		for (int q = 0; q < V; q++) { c += cv[q]; c_run[q] = c; }
		long long a = 0, b = 0;
		if (fast) a = job.sum1 ? load_prefix<VLOG, 1>(limb, tid, lw, a_run) : load_prefix<VLOG, LIMBS>(limb, tid, lw, a_run);
		else {
#pragma unroll
			for (int q = 0; q < V; q++) a_run[q] = 0;
		}
		if (NACC == 2) b = load_prefix<VLOG, LIMBS>(limb + LIMBS * W, tid, lw, b_run);
		int ec; long long ea, eb = 0;
		if (NACC == 2) block_scan3(misc, c, a, b, ec, ea, eb);
		else { int tc; long long ta; block_scan2(misc, c, a, ec, ea, tc, ta); }   // the barrier also fences the loads above
		ec += misc->base_cov;
		ea += (long long) misc->base_val;
		if (NACC == 2) eb += (long long) misc->base_val2;
#pragma unroll
		for (int q = 0; q < V; q++) {
			c_run[q] += ec; a_run[q] += ea;
			if (NACC == 2) b_run[q] += eb;
			if (keep_rule(thr, c_run[q])) keepbits |= 1u << q;
		}
		keepbits &= evbits;
	}
	WB_PHASE(2);

	// breakpoints and kept segments of this thread; one packed block scan for both
	const int nbp_l = __popc(evbits), kept_l = __popc(keepbits);
	int epack, tpack;
	block_scan1(misc, nbp_l | (kept_l << 16), epack, tpack);
	const int nbp = tpack & 0xffff, kept_total = tpack >> 16;
	WB_PHASE(3);

	// publish this window's segment count so that later windows can start their
	// look-back while we compact (single-pass chained scan over windows)
	if (tid == 0) st_volatile(&job.state[w], LB_AGG | (unsigned long long) kept_total);

	// ---- 3. compact the breakpoints over the dead position arrays --------------------
	long long *bval = (long long *) (smem + L.off_bval);   // exact fixed-point sum of the segment
	uint16_t *bpos = (uint16_t *) (smem + L.off_bpos);
	uint16_t *bkid = (uint16_t *) (smem + L.off_bkid);
	long long *bval2 = (long long *) (smem + L.off_bval2);  // moments: exact sum of squares
	uint16_t *bcnt = (uint16_t *) (smem + L.off_bcnt);      // moments: tracks in play
	constexpr bool row_op = MODE == MODE_ROW;
	{
		int rank = epack & 0xffff, kidx = epack >> 16;
		const int p0 = tid << VLOG;
#pragma unroll
		for (int q = 0; q < V; q++) {
			if ((evbits >> q) & 1) {
				bpos[rank] = (uint16_t) (p0 + q);
				uint16_t kd = 0;
				if ((keepbits >> q) & 1) {
					kd = BK_KEEP | (uint16_t) kidx++;
					if (row_op || (c_run[q] >> COV_P_SHIFT) != 0) kd |= BK_SLOW;
				}
				bkid[rank] = kd;
				bval[rank] = a_run[q];
				if (NACC == 2) { bval2[rank] = b_run[q]; bcnt[rank] = (uint16_t) (c_run[q] & COV_MASK); }
				rank++;
			}
		}
	}
	__syncthreads();
	WB_PHASE(4);

	// ---- output offset: decoupled look-back over the preceding windows --------------------
	// Windows are drawn in ascending order, so every predecessor has been started by
	// a resident CTA and will publish.  Warp k inspects windows w-1-32k .. w-32-32k
	// (lane l the l-th of them) and reduces its chunk up to the first inclusive
	// prefix; the chunks are then chained in order.
	const int lbw = job.lb_warps;        // warps taking part in the first look-back step
	if (warp < lbw) {
		const int idx = w - 1 - warp * 32 - lane;
		unsigned long long st;
		while (true) {
			st = idx >= 0 ? ld_volatile(&job.state[idx]) : LB_PREFIX;      // before window 0: prefix 0
			// only the states up to the nearest published prefix are summed
			const unsigned nr = __ballot_sync(FULL, (st >> 62) == 0), pq = __ballot_sync(FULL, (st >> 62) == 2);
			if (nr == 0 || (pq != 0 && (pq & -pq) < (nr & -nr))) break;
			__nanosleep(40);
		}
		const unsigned pm = __ballot_sync(FULL, (st >> 62) == 2);
		const int first = pm ? __ffs(pm) - 1 : 32;
		unsigned long long v = (lane <= first) ? (st & LB_VALUE) : 0;
		for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(FULL, v, d);
		if (lane == 0) { misc->lb_sum[warp] = v; misc->lb_found[warp] = pm != 0; }
	}
	__syncthreads();
	if (warp == 0) {
		unsigned long long sum = 0;
		bool found = false;
		for (int k = 0; k < lbw && !found; k++) { sum += misc->lb_sum[k]; found = misc->lb_found[k] != 0; }
		int p = w - 1 - lbw * 32;
		while (!found) {                                                     // rare: deeper than 32*WARPS windows
			const int idx = p - lane;
			unsigned long long st;
			while (true) {
				st = idx >= 0 ? ld_volatile(&job.state[idx]) : LB_PREFIX;
				const unsigned nr = __ballot_sync(FULL, (st >> 62) == 0), pq = __ballot_sync(FULL, (st >> 62) == 2);
				if (nr == 0 || (pq != 0 && (pq & -pq) < (nr & -nr))) break;
				__nanosleep(40);
			}
			const unsigned pm = __ballot_sync(FULL, (st >> 62) == 2);
			const int first = pm ? __ffs(pm) - 1 : 32;
			unsigned long long v = (lane <= first) ? (st & LB_VALUE) : 0;
			for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(FULL, v, d);
			sum += __shfl_sync(FULL, v, 0);
			found = pm != 0;
			p -= 32;
		}
		if (lane == 0) {
			misc->out_base = (long long) sum;
			st_volatile(&job.state[w], LB_PREFIX | (sum + (unsigned long long) kept_total));
			if (w == job.nw - 1) *job.total = (long long) (sum + (unsigned long long) kept_total);
		}
	}
	__syncthreads();
	WB_PHASE(5);
#if WB_LATETICKET
	// The look-back of a window waits for every window below it, so windows should start in
	// the order they were drawn: the next one is drawn only now (the atomic lands during the
	// emit phase), not two windows ahead (measured in fastsum.cu: look-back waits 8.5k -> 5k cycles).
	unsigned next_ticket = 0;
	constexpr bool long_emit = (MODE == MODE_SELECT || MODE == MODE_WALK || MODE == MODE_ROW) && WB_SEL_TICKET_AT_END;
	if (tid == 0 && !long_emit) next_ticket = atomicAdd(job.ticket, 1u) + gridDim.x;
	if (MODE == MODE_SELECT && !WB_SEL_TICKET_AT_END) {
		// the selection takes a while: pull the next window's intervals (about eight per track from its cursor
		// on) into L2 meanwhile -- both its events phase and its staging read them with dependent loads
		if (tid == 0) misc->ticket1 = (int) next_ticket;
		__syncthreads();
		const int wn = misc->ticket1;
		if (wn < job.nw)
			for (int i = tid; i < N; i += THREADS) {
				const long long g = (toff[i] + job.cursor[(size_t) wn * job.cstride + i]) & ~3ll;
				if (g + 8 <= job.I) {
					asm volatile("prefetch.global.L2 [%0];" :: "l"(job.start + g)); asm volatile("prefetch.global.L2 [%0];" :: "l"(job.end + g));
					asm volatile("prefetch.global.L2 [%0];" :: "l"(job.value + g)); asm volatile("prefetch.global.L2 [%0];" :: "l"(job.value + g + 4));
				}
			}
	}
#endif

	// ---- 4. emit, one breakpoint per thread (coalesced stores) ---------------------------
	const long long out_base = misc->out_base;
	const int next_bp = misc->next_bp;
	unsigned long long slow = 0;
	double auc = 0;

	if (MODE == MODE_FAST || MODE == MODE_MOMENTS) {
		const bool npow2 = (N & (N - 1)) == 0;
		const double inv_n = 1.0 / N;
		for (int base = 0; base < nbp; base += THREADS) {
			const int r = base + tid;
			if (r >= nbp) break;
			const uint16_t kd = bkid[r];
			if (!(kd & BK_KEEP)) continue;
			const long long oi = out_base + (kd & BK_IDX);
			const int bs = (int) (w0 + bpos[r]);
			const int be = (r + 1 < nbp) ? (int) (w0 + bpos[r + 1]) : next_bp;
			double v;
			if (MODE == MODE_MOMENTS) v = moments_stat(job.op == OP_ENTROPY ? OP_STDDEV : job.op, N, bcnt[r], bval[r], bval2[r], scale);
			else {
				v = (double) (bval[r] + job.fx_base) * scale;      // correctly rounded exact sum
				if (job.op == OP_MEAN) v = npow2 ? v * inv_n : v / N;  // 1/N is exact for a power of two
			}
			if (oi < job.out_cap) {
				job.out_start[oi] = bs;
				job.out_end[oi] = be;
				job.out_value[oi] = v;
				if (kd & BK_SLOW) {
					// an input was not representable: the fix-up kernel recomputes this
					// segment with the reference's own operation order
					const unsigned long long k = atomicAdd(job.slow_count, 1ull);
					if (k < (unsigned long long) job.slow_cap) job.slow_list[k] = oi;
				}
			}
			if (!(kd & BK_SLOW) && !isnan(v)) auc += (double) (be - bs) * v;
		}
	} else if (MODE == MODE_WALK || MODE == MODE_SELECT) {
		if (kept_total > 0) {                                 // (a window without a kept segment has nothing to select)
			const WalkLayout wl = walk_layout(N);
			unsigned char *rb = smem + L.off_rows;
			TrackState *tstate = (TrackState *) (rb + wl.off_tstate) + (size_t) warp * N;
			SortKey *keys = (SortKey *) (rb + wl.off_keys);
			uint16_t *rank_of = (uint16_t *) (rb + wl.off_rank);
			double *sv = (double *) (rb + wl.off_stage);
			int32_t *ss = (int32_t *) (sv + wl.cap), *se = ss + wl.cap;
			int32_t *tbase = (int32_t *) (rb + wl.off_tbase);
			// the window's intervals, staged in shared memory (MODE_SELECT: already there since the events phase)
			if (!st_done) stage_window(false);
			const int T = st_T;
			WB_PHASE(11);
			const bool staged = st_ok;
			const int U = T + N;                              // universe of the bitset sweep: intervals + one default per track
			int Ps = 64; while (Ps < U) Ps <<= 1;
			const bool minmax = (job.op == OP_MIN || job.op == OP_MAX);
			const bool sorting = MODE == MODE_SELECT && staged;   // the window gets sorted (else: a sorted array per warp)
			const bool inexact = st_inexact;                  // a value that is not a float32 (NaN, -0 included)
			float *dfl = (float *) (rb + wl.off_bits) + 2 * N;    // bucketed selection: the tracks' defaults as float32
			// bucketed selection (every breakpoint on its own): float32 values, the universe of intervals and gaps fits
			WB_PHASE(14);
			const bool bucketed = sorting && W <= 4096 && !__syncthreads_or(inexact);
			const bool bitset = sorting && !bucketed && U <= wl.P2;
			__syncthreads();
			if (bucketed) {
				// The window's positions are handled in ranges [h0, h0 + ws): the whole window when its elements
				// fit the sorted universe, else halves, quarters ... (a range of one position always fits).
				// Elements of track i in a range: its staged intervals that overlap it -- fa[i] the first of
				// them, cbase[i] = how many the tracks before i have -- and the gaps around them.
				unsigned long long *k64 = (unsigned long long *) (rb + wl.off_keys);
				const size_t keys_bytes = wl.off_rank - wl.off_keys;
				float *val = (float *) k64;                       // after the sort: value of every rank, over the keys
				ushort2 *rse = (ushort2 *) (rb + wl.off_rse);
				int32_t *cbase = (int32_t *) (rb + wl.off_bits), *fa = cbase + N;   // the bitset sweep's rows: not in use on this path
				int32_t *gcount = fa + N + N;                     // (cbase, fa, defaults, gcount: N words each)
				int h0 = 0, ws = W;
				while (h0 < W) {
					int Tp = 0, Pb = 0, Pm = 0;
					bool in_stage = false;
					while (true) {
						const long long ra = w0 + h0, rbnd = ra + ws;
						Tp = 0;
						for (int base = 0; base < N; base += THREADS) {
							const int i = base + tid;
							int first = 0, cnt = 0;
							if (i < N) {
								const int b = tbase[i], L = max(0, min(nxt[i], tn[i] - 1) - cur[i] + 1);
								if (ws == W) {
									// the whole window: everything staged but the look-ahead interval (an interval that ends
									// exactly at the window start stays in as an element that is never in play)
									cnt = L - ((L > 0 && (long long) ss[b + L - 1] >= rbnd) ? 1 : 0);
								} else {
									int lo = 0, hi = L;                  // first staged interval with end > ra
									while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long) se[b + mid] > ra) hi = mid; else lo = mid + 1; }
									first = lo;
									hi = L;                              // first staged interval with start >= rbnd
									while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long) ss[b + mid] >= rbnd) hi = mid; else lo = mid + 1; }
									cnt = lo - first;
								}
							}
							int ec, tc;
							block_scan1(misc, cnt, ec, tc);
							if (i < N) { cbase[i] = Tp + ec; fa[i] = first; gcount[i] = 0; }
							Tp += tc;
							__syncthreads();
						}
						// sorted: the Tp intervals and one marker per track (its default); ranked: intervals and gaps
						Pm = 256; while (Pm < Tp + N) Pm <<= 1;
						Pb = 256; while (Pb < 2 * Tp + N) Pb <<= 1;
						// the counters go over the staging area when this is the last range, else behind the values
						in_stage = h0 + ws == W && (size_t) ws * SEL_NB <= wl.cnt_bytes;
						const bool fits = 2 * Tp + N <= SEL_PS && (size_t) Pm * 8 <= keys_bytes
						                  && (size_t) Pb * 4 + (in_stage ? 0 : (size_t) ws * SEL_NB) <= keys_bytes;
						if (fits || ws == 1) break;
						ws >>= 1;
					}
					const int le = Pm >= 4096 ? 4 : 3;                // keys per thread of the sort: 16 or 8
					const long long ra = w0 + h0, rbnd = ra + ws;
					auto count_of = [&](int i) -> int { return (i + 1 < N ? cbase[i + 1] : Tp) - cbase[i]; };
					// gap m of track i (m = 0 .. count): from the end of interval m - 1 (the range start) to the start of
					// interval m (the range end); b = the track's first staged interval of the range
					auto gap = [&](int b, int len, int m, long long &ga, long long &gb) {
						ga = m == 0 ? ra : max((long long) se[b + m - 1], ra);
						gb = m == len ? rbnd : min((long long) ss[b + m], rbnd);
					};
					const int gid = tid >> 3, gl = tid & 7;
					for (int i = gid; i < N; i += THREADS / 8) {
						const int len = count_of(i), b = tbase[i] + fa[i], cb = cbase[i];
						int ng = 0;
						for (int m = gl; m <= len; m += 8) {
							if (m < len) {
								const int u = cb + m;
								k64[sort_phys(u, le)] = ((unsigned long long) f32_key((float) sv[b + m]) << 32) | ((unsigned long long) i << 16) | (unsigned) u;
							}
							long long ga, gb;
							gap(b, len, m, ga, gb);
							ng += ga < gb ? 1 : 0;
						}
						if (ng) atomicAdd(&gcount[i], ng);              // non-empty gaps of the track
						if (gl == 0) k64[sort_phys(Tp + i, le)] = ((unsigned long long) f32_key(dfl[i]) << 32) | ((unsigned long long) i << 16) | (unsigned) (Tp + i);
					}
					for (int u = Tp + N + tid; u < Pm; u += THREADS) k64[sort_phys(u, le)] = ~0ull;
					__syncthreads();
					WB_PHASE(15);
					if (le == 4) bitonic_u64<4>(k64, Pm); else bitonic_u64<3>(k64, Pm);
					WB_PHASE(12);
					// ranks: an interval takes one, a marker one per non-empty gap of its track.  Every thread takes
					// Pm / THREADS consecutive sorted keys, a block scan of the multiplicities gives their first rank.
					constexpr int KPT = SEL_PS / THREADS;
					unsigned long long myk[KPT];
					const int kpt = Pm / THREADS;                     // 1 .. KPT
					int mine = 0;
#pragma unroll
					for (int q = 0; q < KPT; q++) {
						myk[q] = ~0ull;
						if (q < kpt) myk[q] = k64[sort_phys(tid * kpt + q, le)];
						if (myk[q] != ~0ull) { const int u = (int) (myk[q] & 0xffff); mine += u < Tp ? 1 : gcount[(myk[q] >> 16) & 0xffff]; }
					}
					int fr, R;
					block_scan1(misc, mine, fr, R);                   // (ends with every key in registers: the values may go over them)
					__syncthreads();
#pragma unroll
					for (int q = 0; q < KPT; q++) {
						if (myk[q] == ~0ull) continue;
						const int u = (int) (myk[q] & 0xffff), i = (int) ((myk[q] >> 16) & 0xffff);
						const float v = key_f32((uint32_t) (myk[q] >> 32));
						const int b = tbase[i] + fa[i];
						if (u < Tp) {
							const int m = u - cbase[i];
							ushort2 x;
							x.x = (uint16_t) (min(max((long long) ss[b + m], ra), rbnd) - w0);   // began before the range: in play from its start
							x.y = (uint16_t) (min(max((long long) se[b + m], ra), rbnd) - w0);
							rse[fr] = x; val[fr] = v; fr++;
						} else { const int g = gcount[i]; gcount[i] = fr; fr += g; }   // a marker: its gaps' first rank, for the pass below
					}
					__syncthreads();
					// the gaps of every track at the ranks of its marker: eight lanes per track
					for (int i0 = 0; i0 < N; i0 += THREADS / 8) {
						const int i = i0 + gid;
						const bool on = i < N;
						const int len = on ? count_of(i) : -1, b = on ? tbase[i] + fa[i] : 0;
						int at = on ? gcount[i] : 0;
						const float v = on ? dfl[i] : 0.f;
						const int mmax = __reduce_max_sync(FULL, len);
						for (int m0 = 0; m0 <= mmax; m0 += 8) {
							const int m = m0 + gl;
							long long ga = 0, gb = 0;
							if (m <= len) gap(b, len, m, ga, gb);
							const unsigned grp = (__ballot_sync(FULL, ga < gb) >> (lane & 24)) & 0xffu;   // this track's lanes
							if (ga < gb) {
								const int k = at + __popc(grp & ((1u << gl) - 1));
								ushort2 x; x.x = (uint16_t) (ga - w0); x.y = (uint16_t) (gb - w0);
								rse[k] = x; val[k] = v;
							}
							at += __popc(grp);
						}
					}
					const int bs = max(4, ((R + SEL_NB - 1) / SEL_NB + 3) & ~3);   // ranks per bucket, a multiple of four; R <= 4096: <= 128
					const int Pr = bs * SEL_NB;
					for (int rk = R + tid; rk < Pr; rk += THREADS) { ushort2 x; x.x = (uint16_t) W; x.y = (uint16_t) W; rse[rk] = x; }   // padding: never in play
					__syncthreads();                                  // (the staging area is dead from here on if this is the last range)
					WB_PHASE(9);
					uint32_t *cntw = in_stage ? (uint32_t *) (rb + wl.off_stage) : (uint32_t *) (rb + wl.off_keys + (size_t) Pb * 4);
					{
						const uint4 z = make_uint4(0, 0, 0, 0);
						uint4 *c4 = (uint4 *) cntw;
						for (int t = tid; t < max(1, ws * SEL_NB / 16); t += THREADS) c4[t] = z;
					}
					__syncthreads();
					bucket_select(job, misc, w0, W, h0, ws, nbp, next_bp, bpos, bkid, out_base, val, rse, cntw, bs, auc, slow WB_DIAG_ARG);
					__syncthreads();
					WB_PHASE(13);
					h0 += ws;
				}
			} else if (bitset) {
				for (int u = tid; u < Ps; u += THREADS) {
					SortKey k;
					if (u < T) {
						// the track of staged interval u: the last one whose base is <= u
						int lo = 0, hi = N - 1;
						while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tbase[mid] <= u) lo = mid; else hi = mid - 1; }
						const double v = sv[u];
						k.v = isnan(v) ? INFINITY : v; k.ord = lo; k.pad = u;
					} else if (u < U) {
						const int i = u - T;
						const double d = (minmax && i == 0) ? 0.0 : job.dflt[i];
						k.v = isnan(d) ? INFINITY : d; k.ord = i; k.pad = u;
					} else { k.v = INFINITY; k.ord = INT_MAX; k.pad = INT_MAX; }
					keys[u] = k;
				}
				__syncthreads();
				for (int k = 2; k <= Ps; k <<= 1)
					for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 4
						for (int t = tid; t < Ps / 2; t += THREADS) {
							const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i + j;   // j is a power of two
							const bool up = ((i & k) == 0);
							const SortKey a = keys[i], b = keys[l];
							if (up ? key_less(b, a) : key_less(a, b)) { keys[i] = b; keys[l] = a; }
						}
						__syncthreads();
					}
				for (int pos = tid; pos < Ps; pos += THREADS) { const int u = keys[pos].pad; if (u < U) rank_of[u] = (uint16_t) pos; }
				__syncthreads();
			}
			IvAccess iv; iv.job = &job; iv.cur = cur; iv.tbase = staged ? tbase : nullptr; iv.ss = ss; iv.se = se; iv.sv = sv;
			// each warp sweeps a contiguous share of the window's breakpoints
			const int chunk = (nbp + wl.warps - 1) / wl.warps;
			const int r0 = warp * chunk, r1 = min(nbp, r0 + chunk);
			if (r0 < r1 && !bucketed && warp < wl.warps) {
				if (bitset) {
					uint32_t *bits = (uint32_t *) (rb + wl.off_bits) + (size_t) warp * (wl.P2 / 32);
					bitset_walk(job, cur, nxt, w0, r0, r1, nbp, next_bp, bpos, bkid, out_base, keys, rank_of, bits, Ps / 32, T,
					            tstate, iv, auc, slow);
				} else {
					// few tracks (a sorted array per warp is cheaper than sorting the window), or more
					// intervals than the staging area holds: per-warp sorted arrays
					SortKey *arr = keys + (size_t) warp * wl.P;
					order_walk(job, cur, nxt, w0, r0, r1, nbp, next_bp, bpos, bkid, out_base, arr, tstate, wl.P, iv, auc, slow);
				}
			}
		}
	} else {
		const bool sorted_op = (job.op == OP_MEDIAN || job.op == OP_MWU);
		const bool pearson = (job.op == OP_PEARSON);
		Row row; row.job = &job; row.cur = cur; row.nxt = nxt; row.b = 0;
		Mom mom = {0, 0, 0, 0, 0, 0};
		for (int base = 0; base < nbp; base += THREADS) {
			const int r = base + tid;
			uint16_t kd = 0;
			if (r < nbp) kd = bkid[r];
			const bool keep = (kd & BK_KEEP) != 0;
			if (pearson) {
				if (keep) {
					const int bs = (int) (w0 + bpos[r]);
					const int be = (r + 1 < nbp) ? (int) (w0 + bpos[r + 1]) : next_bp;
					double X, Y;
					row.b = bs;
					row.get(job.px, X); row.get(job.py, Y);
					const double len = (double) (be - bs);
					Mom one = {len, X * len, Y * len, 0, 0, 0};
					merge_mom(mom, one);
				}
			} else if (!sorted_op) {
				if (keep) {
					const long long oi = out_base + (kd & BK_IDX);
					const int bs = (int) (w0 + bpos[r]);
					const int be = (r + 1 < nbp) ? (int) (w0 + bpos[r + 1]) : next_bp;
					row.b = bs;
					const double v = row_eval(row);
					slow++;
					if (oi < job.out_cap) {
						job.out_start[oi] = bs;
						job.out_end[oi] = be;
						job.out_value[oi] = v;
					}
					if (!isnan(v)) auc += (double) (be - bs) * v;
				}
			} else {
				// one warp per kept segment, in lane order
				int P = 1; while (P < N) P <<= 1;
				SortKey *buf = (SortKey *) (smem + L.off_rows) + (size_t) warp * P;
				unsigned todo = __ballot_sync(FULL, keep);
				while (todo) {
					const int src = __ffs(todo) - 1;
					todo &= todo - 1;
					const int rr = base + warp * 32 + src;
					const int bs = (int) (w0 + bpos[rr]);
					const int be = (rr + 1 < nbp) ? (int) (w0 + bpos[rr + 1]) : next_bp;
					row.b = bs;
					const double v = warp_sorted_eval(row, buf, P);
					if (lane == 0) {
						const long long oi = out_base + (bkid[rr] & BK_IDX);
						if (oi < job.out_cap) {
							job.out_start[oi] = bs;
							job.out_end[oi] = be;
							job.out_value[oi] = v;
						}
						if (!isnan(v)) auc += (double) (be - bs) * v;
						slow++;
					}
				}
			}
		}
		if (pearson) {
			for (int d = 16; d > 0; d >>= 1) {
				Mom other;
				other.n = __shfl_down_sync(FULL, mom.n, d); other.sx = __shfl_down_sync(FULL, mom.sx, d);
				other.sy = __shfl_down_sync(FULL, mom.sy, d); other.txx = __shfl_down_sync(FULL, mom.txx, d);
				other.txy = __shfl_down_sync(FULL, mom.txy, d); other.tyy = __shfl_down_sync(FULL, mom.tyy, d);
				if (lane + d < 32) merge_mom(mom, other);
			}
			if (lane == 0) {
				double *wm = misc->wmom[warp];
				wm[0] = mom.n; wm[1] = mom.sx; wm[2] = mom.sy; wm[3] = mom.txx; wm[4] = mom.txy; wm[5] = mom.tyy;
			}
		}
	}

	WB_PHASE(6);
#if WB_LATETICKET
	// (per-segment modes: most of a window is spent after the look-back, so its next window is drawn only here, when
	// this one is done -- windows then start in the order they were drawn and the look-back waits less)
	if (tid == 0 && long_emit) next_ticket = atomicAdd(job.ticket, 1u) + gridDim.x;
	if (tid == 0) misc->ticket2 = (int) next_ticket;
#endif
	// per-warp AUC partials go straight to global memory (fixed order, folded later)
	for (int d = 16; d > 0; d >>= 1) {
		auc += __shfl_down_sync(FULL, auc, d);
		slow += __shfl_down_sync(FULL, slow, d);
	}
	if (lane == 0) {
		if (job.auc_partial) job.auc_partial[(size_t) w * WARPS + warp] = auc;
		if ((MODE == MODE_ROW || MODE == MODE_WALK || MODE == MODE_SELECT) && slow && job.slow_count) atomicAdd(job.slow_count, slow);
	}
	if (MODE == MODE_ROW && job.op == OP_PEARSON) {
		__syncthreads();
		if (tid == 0 && job.moments) {
			double t[6] = {0, 0, 0, 0, 0, 0};
			for (int k = 0; k < WARPS; k++) merge6(t, misc->wmom[k]);
			for (int k = 0; k < 6; k++) job.moments[(size_t) w * 6 + k] = t[k];
		}
	}
}

#ifndef WB_CURSOR_GROUPS
#define WB_CURSOR_GROUPS 2
#endif
constexpr int CURSOR_GROUPS = WB_CURSOR_GROUPS;     // groups of four intervals per thread of the cursor kernel
// cursor kernel: cursor[w][i] = first interval of track i whose end reaches the
// start of window w.  One thread per interval: interval j is that cursor for every
// window whose start lies in (end[j-1], end[j]] -- usually none or one.  Reads the
// end coordinates once (4 B/interval); row nw closes the table.
__global__ void __launch_bounds__(256) cursor_kernel(const Job job) {
	const int i = blockIdx.y;
	const long long o = job.off[i];
	const int n = (int) (job.off[i + 1] - o);
	const int32_t *end = job.end + o, *start = job.start + o;
	// cnt(e) = number of windows that start at or before e (capped at the table's row count); interval j is
	// the cursor of the windows cnt(end[j-1]) .. cnt(end[j]) - 1, and j = n (one past the last) closes the table
	const int p0 = (int) job.p0, wlog = job.wlog, rows = job.nw + 1;
	auto cnt = [&](int e) -> int { return e < p0 ? 0 : (int) min((((unsigned) e - (unsigned) p0) >> wlog) + 1u, (unsigned) rows); };
	// Four consecutive intervals per group on a 16-byte aligned grid: one 128-bit load per array; a thread takes
	// CURSOR_GROUPS groups one block width apart and has the loads of all of them in flight before it looks at
	// any (the kernel is a latency chain of one load pair per thread otherwise: 35 % of the DRAM bandwidth).
	int ev[CURSOR_GROUPS][4], sv[CURSOR_GROUPS][4], j0[CURSOR_GROUPS];
#pragma unroll
	for (int g = 0; g < CURSOR_GROUPS; g++) {
		// j0: track-relative index of the group's first interval (negative in the track's first, partial group)
		j0[g] = 4 * (int) ((blockIdx.x * CURSOR_GROUPS + g) * blockDim.x + threadIdx.x) - (int) (o & 3);
#pragma unroll
		for (int k = 0; k < 4; k++) { ev[g][k] = 0; sv[g][k] = 0; }
		if (j0[g] <= n) {
			if (job.vec && j0[g] >= 0 && j0[g] + 4 <= n) {
				const int4 v = *(const int4 *) (end + j0[g]), u = *(const int4 *) (start + j0[g]);
				ev[g][0] = v.x; ev[g][1] = v.y; ev[g][2] = v.z; ev[g][3] = v.w;
				sv[g][0] = u.x; sv[g][1] = u.y; sv[g][2] = u.z; sv[g][3] = u.w;
			} else {
#pragma unroll
				for (int k = 0; k < 4; k++) if (j0[g] + k >= 0 && j0[g] + k < n) { ev[g][k] = end[j0[g] + k]; sv[g][k] = start[j0[g] + k]; }
			}
		}
	}
	int32_t *col = job.cursor + i;
	const size_t stride = (size_t) job.cstride;
	int bad = 0;
#pragma unroll
	for (int g = 0; g < CURSOR_GROUPS; g++) {
		// the end before this group sits in the previous lane; only lane 0 loads it
		int e_prev = __shfl_up_sync(FULL, ev[g][3], 1);
		if (j0[g] > n) continue;
		if ((threadIdx.x & 31) == 0 && j0[g] > 0) e_prev = end[j0[g] - 1];
		int c_prev = j0[g] > 0 ? cnt(e_prev) : 0;
		// input contract (include/wiggletools_b200.h): within a track start < end and end[j-1] <= start[j]
		// (sorted, disjoint), all inside [p0, pos_hi] -- what NonOverlappingWiggleIterator guarantees the
		// reference's sweep (multiplexer.c:163).  Violations raise a flag the host turns into WB200_EINVAL.
		int pe = j0[g] > 0 ? e_prev : INT_MIN;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int j = j0[g] + k;
			if (j >= 0 && j < n) {
				if (sv[g][k] >= ev[g][k] || pe > sv[g][k]) bad |= 1;
				if (sv[g][k] < p0 || ev[g][k] > job.pos_hi) bad |= 2;
				pe = ev[g][k];
				const int c = cnt(ev[g][k]);
				if (c != c_prev) {                     // at most a few windows start inside one interval's reach
#pragma unroll 1
					for (int w = c_prev; w < c; w++) col[(size_t) w * stride] = j;
					c_prev = c;
				}
			}
		}
		if (n < j0[g] + 4) {                           // this group holds index n: the closing rows
#pragma unroll 1
			for (int w = c_prev; w < rows; w++) col[(size_t) w * stride] = n;
		}
	}
	if (bad) atomicOr(job.flags, bad);
}

template <int MODE, int VLOG, int LIMBS>
__global__ void __launch_bounds__(THREADS, MODE == MODE_FAST ? (THREADS == 512 ? 2 : WB_FAST_MINB) : MODE == MODE_MOMENTS ? 2 : MODE == MODE_WALK ? 3 : MODE == MODE_SELECT ? 2 : 1) emit_kernel(const Job job, int lw, double inv_scale, double lim, double scale, int rowbuf) {
	extern __shared__ __align__(16) unsigned char smem[];
	const Layout L = make_layout(job.wlog, job.N, true, (MODE == MODE_MOMENTS ? 2 : 1) * LIMBS, rowbuf);
	int32_t *cur = (int32_t *) (smem + L.off_cur);
	int32_t *tn = (int32_t *) (smem + L.off_tn);
	long long *toff = (long long *) (smem + L.off_toff);
	long long *tmd = (long long *) (smem + L.off_tmd);
	for (int i = threadIdx.x; i < job.N; i += THREADS) {
		const long long o = job.off[i];
		toff[i] = o;
		tn[i] = (int) (job.off[i + 1] - o);
		tmd[i] = MODE == MODE_FAST ? job.dflt_fx[i] : 0;   // moments path: all defaults are 0
	}
	Misc *misc = (Misc *) smem;
	int32_t *ca = cur, *cb = (int32_t *) (smem + L.off_cur2);   // current / next cursor rows
	// Windows are handed out in order by a ticket counter: the look-back only ever
	// waits on windows that a resident CTA has already drawn.  The first window of
	// every CTA is assigned statically (the whole grid is resident: it is sized by
	// occupancy); each later one is drawn when the previous window's emit phase starts,
	// so that windows start in (nearly) the order they were drawn, and its cursor row
	// travels by cp.async while the window arrays are cleared.  (WB_LATETICKET=0: the
	// earlier scheme, two tickets ahead with the next cursor row prefetched.)
#if WB_LATETICKET
	int w = blockIdx.x;
	if (w < job.nw)
		for (int i = threadIdx.x; i < job.N; i += THREADS) ca[i] = job.cursor[(size_t) w * job.cstride + i];
	__syncthreads();
	while (w < job.nw) {
#ifdef WB_DIAG_PHASES
		const long long diag_t0 = clock64();
#endif
		process_window<MODE, VLOG, LIMBS>(job, w, smem, L, ca, lw, inv_scale, lim, scale);
		__syncthreads();      // dense list consumed, ticket visible
#ifdef WB_DIAG_PHASES
		if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&job.diag[7], (unsigned long long) (t_ - diag_t0)); atomicAdd(&job.diag[8], 1ull); }
#endif
		w = misc->ticket2;
		// the drawn window's cursor row travels while the window arrays are cleared
		if (w < job.nw)
			for (int i = threadIdx.x; i < job.N; i += THREADS) __pipeline_memcpy_async(&ca[i], &job.cursor[(size_t) w * job.cstride + i], 4);
		__pipeline_commit();
	}
	(void) cb;
}
#else
	int w = blockIdx.x, w1 = blockIdx.x + gridDim.x;
	if (w < job.nw)
		for (int i = threadIdx.x; i < job.N; i += THREADS) ca[i] = job.cursor[(size_t) w * job.cstride + i];
	__syncthreads();
	while (w < job.nw) {
#ifdef WB_DIAG_PHASES
		const long long diag_t0 = clock64();
#endif
		unsigned t2 = 0;
		if (threadIdx.x == 0) t2 = atomicAdd(job.ticket, 1u) + 2u * gridDim.x;   // consumed two windows from now
		if (w1 < job.nw)
			for (int i = threadIdx.x; i < job.N; i += THREADS)
				__pipeline_memcpy_async(&cb[i], &job.cursor[(size_t) w1 * job.cstride + i], 4);
		__pipeline_commit();
		process_window<MODE, VLOG, LIMBS>(job, w, smem, L, ca, lw, inv_scale, lim, scale);
		__pipeline_wait_prior(0);
		if (threadIdx.x == 0) misc->ticket2 = (int) t2;
		__syncthreads();      // dense list consumed, next cursor row landed, ticket visible
#ifdef WB_DIAG_PHASES
		if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&job.diag[7], (unsigned long long) (t_ - diag_t0)); atomicAdd(&job.diag[8], 1ull); }
#endif
		w = w1; w1 = misc->ticket2;
		{ int32_t *t = ca; ca = cb; cb = t; }
	}
}
#endif

// max |v| over a strided sample of the values, and the finest significant bit seen:
// they choose the fixed-point scale.  out2[0] = bits of the largest finite |v|;
// out2[1] = 2048 - (exponent of the lowest set significand bit), 0 if no non-zero value.
__global__ void sample_maxabs_kernel(const double *value, int64_t n, int64_t stride, double *out2) {
	double m = 0;
	long long key = 0;
	for (int64_t i = (int64_t) (blockIdx.x * blockDim.x + threadIdx.x) * stride; i < n; i += (int64_t) gridDim.x * blockDim.x * stride) {
		const double v = fabs(value[i]);
		if (v < DBL_MAX && v > m) m = v;   // skips NaN and Inf
		if (v < DBL_MAX && v >= DBL_MIN) {
			const long long bits = __double_as_longlong(v);
			const long long mant = (bits & ((1ll << 52) - 1)) | (1ll << 52);
			const int lsb = (int) (bits >> 52) - 1075 + (__ffsll(mant) - 1);
			key = max(key, (long long) (2048 - lsb));
		}
	}
	for (int d = 16; d > 0; d >>= 1) {
		m = fmax(m, __shfl_down_sync(FULL, m, d));
		key = max(key, __shfl_down_sync(FULL, key, d));
	}
	if ((threadIdx.x & 31) == 0) {
		if (m > 0) atomicMax((unsigned long long *) out2, (unsigned long long) __double_as_longlong(m)); // positive doubles order like integers
		if (key > 0) atomicMax((unsigned long long *) out2 + 1, (unsigned long long) key);
	}
}

__global__ void __launch_bounds__(1024) reduce_auc_kernel(const double *partial, int nw, double *out) {
	__shared__ double wtot[32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	double acc = 0;
	for (int i = tid; i < nw; i += 1024) acc += partial[i];
	for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(FULL, acc, d);
	if (lane == 0) wtot[warp] = acc;
	__syncthreads();
	if (tid == 0) { double t = 0; for (int k = 0; k < 32; k++) t += wtot[k]; *out = t; }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static int row_buffer_bytes(int N, int op) {
	const bool order = (op == OP_MEDIAN || op == OP_MIN || op == OP_MAX);
	if (order && N <= WALK_MAX_TRACKS) return (int) walk_layout(N).total;          // sweep kernels
	if (op != OP_MEDIAN && op != OP_MWU) return 0;
	int P = 1; while (P < N) P <<= 1;
	return P * (int) sizeof(SortKey) * WARPS;                                      // one sort buffer per warp
}

// shared memory of one emit_kernel CTA: limb_arrays = 32-bit limb arrays per position of the kernel
// that will run (sum / mean: 1..3; var / stddev / CV: 4 or 6; every other statistic: 2)
size_t window_smem_bytes(int wlog, int N, bool emit, int op, int limb_arrays) {
	return make_layout(wlog, N, emit, limb_arrays, emit ? row_buffer_bytes(N, op) : 0).total;
}

cudaError_t launch_cursor(const Job &job, int64_t max_track_len, cudaStream_t s) {
	// max_track_len + 1 intervals (the closing one) + up to 3 slots of alignment, 4 per thread
	dim3 grid((unsigned) (((max_track_len + 8) / 4 + 256 * CURSOR_GROUPS - 1) / (256 * CURSOR_GROUPS)), (unsigned) job.N);
	cursor_kernel<<<grid, 256, 0, s>>>(job);
	return cudaGetLastError();
}

template <int MODE, int VLOG, int LIMBS>
static cudaError_t launch_emit_v(const Job &job, int lw, double inv_scale, double lim, double scale, int rowbuf, cudaStream_t s) {
	size_t smem = make_layout(job.wlog, job.N, true, (MODE == MODE_MOMENTS ? 2 : 1) * LIMBS, rowbuf).total;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaError_t e = cudaFuncSetAttribute(emit_kernel<MODE, VLOG, LIMBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (e != cudaSuccess) return e;
	int per = 1;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, emit_kernel<MODE, VLOG, LIMBS>, THREADS, smem);
	if (getenv("WB200_FS_VERBOSE")) fprintf(stderr, "emit_kernel<%d,%d,%d>: %zu B smem, %d CTAs/SM\n", MODE, VLOG, LIMBS, smem, per);
	int grid = min(job.nw, sms * max(per, 1));
	emit_kernel<MODE, VLOG, LIMBS><<<grid, THREADS, smem, s>>>(job, lw, inv_scale, lim, scale, rowbuf);
	return cudaGetLastError();
}

cudaError_t launch_emit(const Job &job, cudaStream_t s, int *auc_per_window) {
	*auc_per_window = WARPS;
	// fixed-point format: |m| < 2^mbits, limbs of lw bits
	int l2n = 1; while ((1 << l2n) < 2 * job.N) l2n++;
	int lw = 32 - l2n;
	int limbs = 2;
	// widest delta: |m| < 2^mbits (+1 for m - default), or m^2 < 2^(2 mbits) on the moments path
	const int dbits = job.fast == 2 ? 2 * job.fx_mbits : job.fx_mbits + 1;
	if (job.fast && dbits > 30 - l2n + lw) limbs = 3;
	if (job.fast == 1) limbs = job.fx_limbs;      // sum / mean: chosen by the host from the sampled value range
	double inv_scale = ldexp(1.0, -job.fx_q), scale = ldexp(1.0, job.fx_q), lim = ldexp(1.0, job.fx_mbits);
	{
		const char *e = getenv("WB200_FASTSUM");       // development: 0 keeps sum / mean on emit_kernel<MODE_FAST>
		if (!(e && atoi(e) == 0) && fastsum_supported(job)) {
			*auc_per_window = 1;
			return launch_fastsum(job, lw, inv_scale, lim, scale, s);
		}
	}
	int rowbuf = row_buffer_bytes(job.N, job.op);
	const bool walk = (job.op == OP_MEDIAN || job.op == OP_MIN || job.op == OP_MAX) && job.N <= WALK_MAX_TRACKS;
#define WB_EMIT(V) (walk ? (job.N > WB_SEL_MIN_N ? launch_emit_v<MODE_SELECT, V, 2>(job, lw, inv_scale, lim, scale, rowbuf, s) \
                                               : launch_emit_v<MODE_WALK, V, 2>(job, lw, inv_scale, lim, scale, rowbuf, s)) \
                  : !job.fast ? launch_emit_v<MODE_ROW, V, 2>(job, lw, inv_scale, lim, scale, rowbuf, s) \
                  : job.fast == 2 ? (limbs == 2 ? launch_emit_v<MODE_MOMENTS, V, 2>(job, lw, inv_scale, lim, scale, rowbuf, s) \
                                               : launch_emit_v<MODE_MOMENTS, V, 3>(job, lw, inv_scale, lim, scale, rowbuf, s)) \
                  : limbs == 1 ? launch_emit_v<MODE_FAST, V, 1>(job, lw, inv_scale, lim, scale, rowbuf, s) \
                  : limbs == 2 ? launch_emit_v<MODE_FAST, V, 2>(job, lw, inv_scale, lim, scale, rowbuf, s) \
                               : launch_emit_v<MODE_FAST, V, 3>(job, lw, inv_scale, lim, scale, rowbuf, s))
	switch (job.wlog - TLOG) {       // positions per thread = 2^VLOG
	case 2: return WB_EMIT(2);
	case 3: return WB_EMIT(3);
	case 4: return WB_EMIT(4);
	}
#undef WB_EMIT
	return cudaErrorInvalidValue;
}

// Segments of the fast path whose inputs were not representable in fixed point
// (NaN, Inf, out of range): recompute them exactly as the reference does.  The
// cursor rows of the window that holds the segment start bound the search.
__global__ void fixup_kernel(const Job job, long long count) {
	const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= count) return;
	const long long oi = job.slow_list[k];
	const int bs = job.out_start[oi];
	const int w = (int) (((long long) bs - job.p0) >> job.wlog);
	Row row;
	row.job = &job;
	row.cur = job.cursor + (size_t) w * job.cstride;
	row.nxt = job.cursor + (size_t) (w + 1) * job.cstride;
	row.b = bs;
	const double v = row_eval(row);
	job.out_value[oi] = v;
	if (!isnan(v) && job.auc_fix) atomicAdd(job.auc_fix, (double) (job.out_end[oi] - bs) * v);
}

cudaError_t launch_fixup(const Job &job, long long count, cudaStream_t s) {
	if (count <= 0) return cudaSuccess;
	fixup_kernel<<<(unsigned) ((count + 127) / 128), 128, 0, s>>>(job, count);
	return cudaGetLastError();
}

// Raw multiplexer rows: what popCoreMultiplexer2 leaves in Multiplexer.values[] / inplay[]
// (/root/reference/src/multiplexer.c:76-96) for the segments [first, first + count) of the last
// reduction -- one thread per (segment, track); the covering interval is searched between the
// cursor rows of the window that holds the segment start.  values: the track's value where it
// is in play, else its default.
__global__ void rows_kernel(const Job job, long long first, long long count, double *values, unsigned char *inplay) {
	const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= count * job.N) return;
	const long long k = idx / job.N;
	const int i = (int) (idx - k * job.N);
	const int pos = job.out_start[first + k];
	const int w = (int) (((long long) pos - job.p0) >> job.wlog);
	const long long o = job.off[i];
	const int n = (int) (job.off[i + 1] - o);
	int lo = job.cursor[(size_t) w * job.cstride + i], hi = min(job.cursor[(size_t) (w + 1) * job.cstride + i], n - 1);
	int ans = -1;                        // first interval in [lo, hi] whose end lies beyond pos (ends ascend)
	while (lo <= hi) {
		const int mid = (lo + hi) >> 1;
		if (job.end[o + mid] > pos) { ans = mid; hi = mid - 1; } else lo = mid + 1;
	}
	const bool in = ans >= 0 && job.start[o + ans] <= pos;
	values[idx] = in ? job.value[o + ans] : job.dflt[i];
	inplay[idx] = in ? 1 : 0;
}

cudaError_t launch_rows(const Job &job, long long first, long long count, double *values, unsigned char *inplay, cudaStream_t s) {
	if (count <= 0) return cudaSuccess;
	const long long total = count * job.N;
	rows_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, s>>>(job, first, count, values, inplay);
	return cudaGetLastError();
}

cudaError_t launch_sample_maxabs(const double *value, int64_t n, int64_t stride, double *out2, cudaStream_t s) {
	cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(double), s);
	if (e != cudaSuccess) return e;
	if (n > 0) sample_maxabs_kernel<<<64, 256, 0, s>>>(value, n, stride, out2);
	return cudaGetLastError();
}

// fold the per-window Pearson moments in window order (deterministic)
__global__ void __launch_bounds__(1024) reduce_moments_kernel(const double *moments, int nw, double *out6) {
	__shared__ double part[1024][6];
	const int tid = threadIdx.x;
	const int per = (nw + 1023) / 1024;
	double m[6] = {0, 0, 0, 0, 0, 0};
	for (int i = tid * per; i < min(nw, (tid + 1) * per); i++) merge6(m, moments + (size_t) i * 6);
	for (int k = 0; k < 6; k++) part[tid][k] = m[k];
	__syncthreads();
	for (int d = 1; d < 1024; d <<= 1) {
		if ((tid & (2 * d - 1)) == 0 && tid + d < 1024) merge6(part[tid], part[tid + d]);
		__syncthreads();
	}
	if (tid == 0) for (int k = 0; k < 6; k++) out6[k] = part[0][k];
}

cudaError_t launch_reduce_moments(const double *moments, int nw, double *out6, cudaStream_t s) {
	reduce_moments_kernel<<<1, 1024, 0, s>>>(moments, nw, out6);
	return cudaGetLastError();
}

cudaError_t launch_reduce_auc(const double *partial, int count, double *out, cudaStream_t s) {
	reduce_auc_kernel<<<1, 1024, 0, s>>>(partial, count, out);
	return cudaGetLastError();
}

// the DEVICE build of tdist_Q on n argument pairs (tests pin it on a high-precision table)
__global__ void tdist_kernel(const double *t, const double *nu, int n, double *out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = tdist_Q(t[i], nu[i]);
}
cudaError_t launch_tdist(const double *t, const double *nu, int n, double *out, cudaStream_t s) {
	if (n <= 0) return cudaSuccess;
	tdist_kernel<<<(n + 127) / 128, 128, 0, s>>>(t, nu, n, out);
	return cudaGetLastError();
}

} // namespace wb

// exported for tests: the host build of the same routine the kernels use
extern "C" double wb200_tdist_Q(double t, double nu) { return wb::tdist_Q(t, nu); }
