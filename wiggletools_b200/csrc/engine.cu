// wiggletools_b200 -- sm_100a kernels of the interval-stream reduction engine.
//
// What the reference does one segment at a time on one CPU thread
// (popCoreMultiplexer2, /root/reference/src/multiplexer.c:98-121, followed by
// an O(N) reducer loop such as MeanReductionPop, src/reducers.c:367-402), this
// file does per genomic WINDOW of 2^wlog bp, one CTA per window:
//
//   1. mark   : every interval start/end that falls in the window sets a bit of
//               a shared-memory bitmap  -> the window's union breakpoints
//               (replaces the two Fibonacci heaps of multiplexer.c:37-96);
//   2. rank   : popcount prefix over the bitmap words -> slot of each breakpoint;
//   3. events : each interval adds +delta at slot(start) and -delta at
//               slot(end): a packed cover counter (how many tracks of set A/B
//               are in play) and, for sum/mean, the value as a fixed-point
//               integer split in 32-bit limbs (native shared-memory atomics;
//               integer adds commute, so the result is order-independent and
//               exact -- no floating-point atomics anywhere);
//   4. scan   : prefix sum over slots = state of the sweep after each
//               breakpoint = cover set + exact sum of the segment that starts
//               there (the reference recomputes this with an O(N) loop per
//               segment; here it is O(1) per segment);
//   5. emit   : segments with a non-empty cover (or a full one under `strict`,
//               multiplexer.c:123-128) are compacted and written coalesced.
//
// Statistics that are not sums (min/max/product/var/...) and segments whose
// inputs do not fit the fixed-point format (NaN, Inf, huge dynamic range) are
// evaluated by row_eval(): a per-segment loop over the N tracks in index order
// that mirrors the reference's arithmetic operation for operation.
//
// Two kernels per chromosome: count_kernel (steps 1-4, segment counts and the
// per-window cursor table) and emit_kernel (steps 1-5 with values), with a
// scan of the per-window counts in between.
#include "engine.cuh"
#include <cfloat>
#include <climits>
#include <cmath>

namespace wb {

#define FULL 0xffffffffu

// ---------------------------------------------------------------------------
// shared-memory layouts
// ---------------------------------------------------------------------------
// Arrays scanned by thread-owned contiguous chunks (V = 2^k entries per thread)
// are stored with one pad word per 32: pidx(p) = p + (p >> 5), which makes the
// chunked accesses bank-conflict free for every power-of-two V.
__host__ __device__ __forceinline__ int pidx(int p) { return p + (p >> 5); }

constexpr int ITRACK_CAP = 2048;     // flat (track,interval) items per window handled in one pass

struct Misc {                        // scalars and per-thread scan bases
	int nb;                          // breakpoints in the window
	int next_bp;                     // first breakpoint at or after the window end
	int base_cov;                    // count kernel: cover carried into the window
	int total_items;                 // flat item count T
	int kept;                        // count kernel: kept segments
	int pad[3];
	long long wsum[WARPS];           // warp totals (value scan / slow counter)
	int wcov[WARPS];                 // warp totals (cover scan)
	int wkeep[WARPS];                // warp totals (compaction)
	double wauc[WARPS];
	double wmom[WARPS][6];           // Pearson moments per warp
	int tb_cov[THREADS];             // per-thread exclusive scan bases
	int kbase[THREADS];
	long long tb_val[THREADS];
};
constexpr int MISC_BYTES = 5120;
static_assert(sizeof(Misc) <= MISC_BYTES, "Misc must fit its shared-memory slot");

struct CountLayout { size_t off_cur, off_nxt, off_toff, off_tn, off_covp, off_ev, total; };

__host__ __device__ inline CountLayout make_count_layout(int wlog, int N) {
	CountLayout L;
	const int W = 1 << wlog;
	size_t o = MISC_BYTES;
	const size_t nn = (size_t) (N + 3) / 4 * 16;
	L.off_cur = o;  o += nn;
	L.off_nxt = o;  o += nn;
	L.off_tn = o;   o += nn;
	L.off_toff = o; o += 2 * nn;
	L.off_covp = o; o += (size_t) (W + W / 32 + 4) * 4;
	L.off_ev = o;   o += (size_t) W;
	L.total = (o + 15) / 16 * 16;
	return L;
}

struct EmitLayout {
	int words, slots, pslots;        // bitmap words; W+1 slots; padded slot count
	size_t off_cur, off_nxt, off_pre, off_toff, off_tn, off_tmd, off_bitmap, off_wprefix, off_cov, off_limb,
	       off_klocal, off_bpos, off_itrack, off_rows, total;
};

__host__ __device__ inline EmitLayout make_emit_layout(int wlog, int N, int limbs, int rowbuf) {
	EmitLayout L;
	const int W = 1 << wlog;
	L.words = W / 32;
	L.slots = W + 1;
	L.pslots = (pidx(W + 1) + 8) / 4 * 4;
	size_t o = MISC_BYTES;
	const size_t nn = (size_t) (N + 4) / 4 * 16;
	L.off_cur = o;     o += nn;
	L.off_nxt = o;     o += nn;
	L.off_pre = o;     o += nn;
	L.off_tn = o;      o += nn;
	L.off_toff = o;    o += 2 * nn;
	L.off_tmd = o;     o += 2 * nn;
	L.off_bitmap = o;  o += (size_t) L.words * 4;
	L.off_wprefix = o; o += (size_t) L.words * 4;
	L.off_cov = o;     o += (size_t) L.pslots * 4;
	L.off_limb = o;    o += (size_t) limbs * L.pslots * 4;
	L.off_klocal = o;  o += (size_t) L.pslots * 2;
	L.off_bpos = o;    o += (size_t) (L.slots + 8) / 8 * 16;
	L.off_itrack = o;  o += (size_t) ITRACK_CAP * 2;
	L.off_rows = (o + 15) / 16 * 16; o = L.off_rows + (size_t) rowbuf;
	L.total = (o + 15) / 16 * 16;
	return L;
}

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int slot_of(const uint32_t *bitmap, const uint32_t *wprefix, int x) {
	uint32_t w = bitmap[x >> 5];
	return (int) wprefix[x >> 5] + __popc(w & ((1u << (x & 31)) - 1u));
}

// first j in [0,n) with a[j] >= key (a ascending)
__device__ __forceinline__ int lower_bound_ge(const int32_t *a, int n, long long key) {
	int lo = 0, hi = n;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if ((long long) a[mid] < key) lo = mid + 1; else hi = mid;
	}
	return lo;
}

__device__ __forceinline__ bool to_fixed(double v, double inv_scale, double lim, long long &m) {
	double t = v * inv_scale;
	if (!(fabs(t) < lim)) return false;          // NaN, Inf, too large
	if (t != rint(t)) return false;              // bits below 2^q
	if (t == 0.0 && v != 0.0) return false;      // underflow
	m = (long long) t;
	return true;
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
__device__ double row_eval(const Row &r) {
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

__device__ double warp_sorted_eval(const Row &r, SortKey *buf, int P /*pow2 >= N*/) {
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
	__syncthreads();
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
	__syncthreads();
}

__device__ __forceinline__ bool keep_rule(const Job &job, bool setop, int c) {
	const int covA = c & COV_MASK, covB = (c >> COV_B_SHIFT) & COV_MASK;
	if (setop) return covA > 0 && covB > 0;            // src/setComparisons.c:48-54
	if (job.strict) return covA == job.N;              // src/multiplexer.c:123-128
	return covA > 0;                                   // gaps are skipped, :112-115
}

// ---------------------------------------------------------------------------
// count kernel: segments per window + the per-window cursor table
// ---------------------------------------------------------------------------
// Position-indexed: one packed cover delta per base pair of the window
// (atomicAdd, native 32-bit) and a byte flag "an interval starts or ends here".
// One pass over the window's intervals, then a chunked prefix scan over the W
// positions.  One CTA walks a strip of consecutive windows so that each track's
// cursor is chained from window to window (binary search only at strip start).
template <int VLOG>
__device__ void count_window(const Job &job, int w, unsigned char *smem, const CountLayout &L) {
	constexpr int V = 1 << VLOG, W = THREADS << VLOG;
	Misc *misc = (Misc *) smem;
	int32_t *cur = (int32_t *) (smem + L.off_cur);
	int32_t *nxt = (int32_t *) (smem + L.off_nxt);
	const int32_t *tn = (const int32_t *) (smem + L.off_tn);
	const long long *toff = (const long long *) (smem + L.off_toff);
	int32_t *covp = (int32_t *) (smem + L.off_covp);
	unsigned char *ev = smem + L.off_ev;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = job.N;
	const long long w0 = job.p0 + ((long long) w << job.wlog);
	const long long w1 = w0 + W;
	const bool setop = (job.op == OP_TTEST || job.op == OP_MWU);

	{
		uint4 z = make_uint4(0, 0, 0, 0);
		uint4 *zc = (uint4 *) covp;
		for (int t = tid; t < (W + W / 32 + 4) / 4; t += THREADS) zc[t] = z;
		uint4 *ze = (uint4 *) ev;
		for (int t = tid; t < W / 16; t += THREADS) ze[t] = z;
		if (tid == 0) { misc->base_cov = 0; misc->kept = 0; }
	}
	__syncthreads();

	for (int i = warp; i < N; i += WARPS) {
		const long long o = toff[i];
		const int n = tn[i];
		const int f = (i < job.n1) ? 1 : (1 << COV_B_SHIFT);
		int j0 = cur[i];
		while (true) {
			const int j = j0 + lane;
			const bool valid = j < n;
			int s = 0, e = 0;
			if (valid) { s = job.start[o + j]; e = job.end[o + j]; }
			const bool stop = !valid || (long long) e >= w1;
			const unsigned bal = __ballot_sync(FULL, stop);
			const int Lst = bal ? __ffs(bal) - 1 : 32;
			if (lane < Lst || (lane == Lst && valid && (long long) s < w1)) {
				if ((long long) s >= w0) { const int x = (int) (s - w0); atomicAdd(&covp[pidx(x)], f); ev[x] = 1; }
				else atomicAdd(&misc->base_cov, f);                       // spans the window start
				if (lane < Lst) { const int x = (int) (e - w0); atomicAdd(&covp[pidx(x)], -f); ev[x] = 1; }
			}
			if (bal) { if (lane == 0) nxt[i] = j0 + Lst; break; }
			j0 += 32;
		}
	}
	__syncthreads();

	const int p0 = tid * V;
	int local = 0;
#pragma unroll
	for (int q = 0; q < V; q++) local += covp[pidx(p0 + q)];
	int ec, tc;
	block_scan1(misc, local, ec, tc);
	int running = misc->base_cov + ec, kept = 0;
#pragma unroll
	for (int q = 0; q < V; q++) {
		running += covp[pidx(p0 + q)];
		if (ev[p0 + q]) kept += keep_rule(job, setop, running) ? 1 : 0;
	}
	for (int d = 16; d > 0; d >>= 1) kept += __shfl_down_sync(FULL, kept, d);
	if (lane == 0 && kept) atomicAdd(&misc->kept, kept);
	__syncthreads();
	if (tid == 0) job.wcount[w] = (uint32_t) misc->kept;
	for (int i = tid; i < N; i += THREADS) job.cursor[(size_t) w * N + i] = cur[i];
}

template <int VLOG>
__global__ void __launch_bounds__(THREADS) count_kernel(const Job job) {
	extern __shared__ __align__(16) unsigned char smem[];
	const CountLayout L = make_count_layout(job.wlog, job.N);
	int32_t *cur = (int32_t *) (smem + L.off_cur);
	int32_t *nxt = (int32_t *) (smem + L.off_nxt);
	int32_t *tn = (int32_t *) (smem + L.off_tn);
	long long *toff = (long long *) (smem + L.off_toff);
	const int wbeg = blockIdx.x * job.strip;
	const int wend = min(job.nw, wbeg + job.strip);
	if (wbeg >= wend) return;
	const long long w0 = job.p0 + ((long long) wbeg << job.wlog);
	for (int i = threadIdx.x; i < job.N; i += THREADS) {
		const long long o = job.off[i];
		const int n = (int) (job.off[i + 1] - o);
		toff[i] = o; tn[i] = n;
		cur[i] = lower_bound_ge(job.end + o, n, w0);
	}
	__syncthreads();
	for (int w = wbeg; w < wend; w++) {
		count_window<VLOG>(job, w, smem, L);
		__syncthreads();
		for (int i = threadIdx.x; i < job.N; i += THREADS) cur[i] = nxt[i];
		__syncthreads();
	}
	// the cursor row after the last window (row nw) closes the table for emit
	if (wend == job.nw)
		for (int i = threadIdx.x; i < job.N; i += THREADS) job.cursor[(size_t) job.nw * job.N + i] = cur[i];
}

// ---------------------------------------------------------------------------
// emit kernel: the window pipeline with values
// ---------------------------------------------------------------------------
template <int LIMBS>
__device__ void emit_window(const Job &job, int w, unsigned char *smem, const EmitLayout &L,
                            int lw, double inv_scale, double lim, double scale) {
	Misc *misc = (Misc *) smem;
	int32_t *cur = (int32_t *) (smem + L.off_cur);
	int32_t *nxt = (int32_t *) (smem + L.off_nxt);
	int32_t *pre = (int32_t *) (smem + L.off_pre);
	const int32_t *tn = (const int32_t *) (smem + L.off_tn);
	const long long *toff = (const long long *) (smem + L.off_toff);
	const long long *tmd = (const long long *) (smem + L.off_tmd);
	uint32_t *bitmap = (uint32_t *) (smem + L.off_bitmap);
	uint32_t *wprefix = (uint32_t *) (smem + L.off_wprefix);
	int32_t *cov = (int32_t *) (smem + L.off_cov);
	uint32_t *limb = (uint32_t *) (smem + L.off_limb);
	uint16_t *klocal = (uint16_t *) (smem + L.off_klocal);
	uint16_t *bpos = (uint16_t *) (smem + L.off_bpos);
	uint16_t *itrack = (uint16_t *) (smem + L.off_itrack);
	const int ls = L.pslots;                      // limb array stride

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = job.N;
	const int W = 1 << job.wlog;
	const long long w0 = job.p0 + ((long long) w << job.wlog);
	const long long w1 = w0 + W;
	const bool setop = (job.op == OP_TTEST || job.op == OP_MWU);
	const bool fast = job.fast != 0;

	// ---- cursors of this window and the next; flat item index ----------------
	int run = 0;   // carry of the per-track length scan
	for (int base = 0; base < N; base += THREADS) {
		const int i = base + tid;
		int len = 0;
		if (i < N) {
			const int c = job.cursor[(size_t) w * N + i];
			const int x = job.cursor[(size_t) (w + 1) * N + i];
			cur[i] = c; nxt[i] = x;
			len = max(0, min(x, tn[i] - 1) - c + 1);
		}
		int ec, tc;
		block_scan1(misc, len, ec, tc);
		if (i < N) pre[i] = run + ec;
		run += tc;
	}
	const int T = run;
	const bool flat = T <= ITRACK_CAP;
	for (int t = tid; t < L.words; t += THREADS) bitmap[t] = 0;
	if (tid == 0) misc->next_bp = INT_MAX;
	__syncthreads();
	if (flat)
		for (int i = tid; i < N; i += THREADS) {
			const int b = pre[i], len = max(0, min(nxt[i], tn[i] - 1) - cur[i] + 1);
			for (int q = 0; q < len; q++) itrack[b + q] = (uint16_t) i;
		}
	__syncthreads();

	// ---- phase 1: mark breakpoints ------------------------------------------
	auto mark_item = [&](int i, int j) {
		const long long o = toff[i];
		const int s = job.start[o + j], e = job.end[o + j];
		if (j < nxt[i]) {
			// interval ends inside the window (w0 <= e < w1)
			if ((long long) s >= w0) { const int x = (int) (s - w0); atomicOr(&bitmap[x >> 5], 1u << (x & 31)); }
			const int x = (int) (e - w0); atomicOr(&bitmap[x >> 5], 1u << (x & 31));
		} else if ((long long) s < w1) {
			// first interval reaching the window end; its start is still ours
			if ((long long) s >= w0) { const int x = (int) (s - w0); atomicOr(&bitmap[x >> 5], 1u << (x & 31)); }
			atomicMin(&misc->next_bp, e);
		} else
			atomicMin(&misc->next_bp, s);
	};
	if (flat) {
		for (int k = tid; k < T; k += THREADS) { const int i = itrack[k]; mark_item(i, cur[i] + k - pre[i]); }
	} else {
		for (int i = warp; i < N; i += WARPS) {
			const int jl = min(nxt[i], tn[i] - 1);
			for (int j = cur[i] + lane; j <= jl; j += 32) mark_item(i, j);
		}
	}
	__syncthreads();

	// ---- phase 2: rank the breakpoints ----------------------------------------
	{
		const int pc = (tid < L.words) ? __popc(bitmap[tid]) : 0;   // L.words <= THREADS
		int ec, tc;
		block_scan1(misc, pc, ec, tc);
		if (tid < L.words) wprefix[tid] = 1 + ec;                   // slot of the word's first set bit
		if (tid == 0) misc->nb = tc;
		__syncthreads();
	}
	const int nb = misc->nb;
	int vlog = 0;
	while ((THREADS << vlog) < nb + 1) vlog++;
	const int V = 1 << vlog;
	{
		const int top = pidx(nb) + 1;
		for (int t = tid; t < top; t += THREADS) {
			cov[t] = 0;
#pragma unroll
			for (int l = 0; l < LIMBS; l++) limb[l * ls + t] = 0;
		}
		if (tid < L.words) {
			uint32_t bits = bitmap[tid];
			int slot = (int) wprefix[tid];
			while (bits) {
				const int bit = __ffs(bits) - 1;
				bits &= bits - 1;
				bpos[slot++] = (uint16_t) (tid * 32 + bit);
			}
		}
	}
	__syncthreads();

	// ---- phase 3: events -> cover deltas and fixed-point value deltas ---------
	auto add_value = [&](int slot, long long d) {
		const int ps = pidx(slot);
		if (LIMBS == 2) {
			atomicAdd(&limb[ps], (uint32_t) (d & ((1ll << lw) - 1)));
			atomicAdd(&limb[ls + ps], (uint32_t) (d >> lw));
		} else {
			atomicAdd(&limb[ps], (uint32_t) (d & ((1ll << lw) - 1)));
			atomicAdd(&limb[ls + ps], (uint32_t) ((d >> lw) & ((1ll << lw) - 1)));
			atomicAdd(&limb[2 * ls + ps], (uint32_t) (d >> (2 * lw)));
		}
	};
	auto event_item = [&](int i, int j) {
		const long long o = toff[i];
		const int s = job.start[o + j], e = job.end[o + j];
		const bool inside = j < nxt[i];
		if (!inside && (long long) s >= w1) return;
		int cd = (i < job.n1) ? 1 : (1 << COV_B_SHIFT);
		long long d = 0;
		if (fast) {
			long long m;
			if (to_fixed(job.value[o + j], inv_scale, lim, m)) d = m - tmd[i];
			else cd += (1 << COV_P_SHIFT);
		}
		const int ss = ((long long) s >= w0) ? slot_of(bitmap, wprefix, (int) (s - w0)) : 0;
		atomicAdd(&cov[pidx(ss)], cd);
		if (fast && d != 0) add_value(ss, d);
		if (inside) {
			const int se = slot_of(bitmap, wprefix, (int) (e - w0));
			atomicAdd(&cov[pidx(se)], -cd);
			if (fast && d != 0) add_value(se, -d);
		}
	};
	if (flat) {
		for (int k = tid; k < T; k += THREADS) { const int i = itrack[k]; event_item(i, cur[i] + k - pre[i]); }
	} else {
		for (int i = warp; i < N; i += WARPS) {
			const int jl = min(nxt[i], tn[i] - 1);
			for (int j = cur[i] + lane; j <= jl; j += 32) event_item(i, j);
		}
	}
	__syncthreads();

	// ---- phase 4: chunked prefix scan over the slots --------------------------
	// thread t owns slots [t*V, t*V+V); local inclusive prefixes are written back
	// in place, the per-thread bases stay in shared memory (tb_cov / tb_val).
	auto load_val = [&](int ps) -> long long {
		if (LIMBS == 2)
			return (long long) (unsigned long long) limb[ps] + ((long long) (int32_t) limb[ls + ps] << lw);
		return (long long) (unsigned long long) limb[ps] + ((long long) (unsigned long long) limb[ls + ps] << lw)
		     + ((long long) (int32_t) limb[2 * ls + ps] << (2 * lw));
	};
	const int r0 = tid << vlog;
	{
		int c = 0; long long a = 0;
		for (int q = 0; q < V; q++) {
			const int r = r0 + q;
			if (r > nb) break;
			const int ps = pidx(r);
			c += cov[ps];
			cov[ps] = c;
			if (fast) {
				a += load_val(ps);
				limb[ps] = (uint32_t) a;              // from here on limb[0..1] hold the raw int64 prefix
				limb[ls + ps] = (uint32_t) ((unsigned long long) a >> 32);
			}
		}
		int ec, tc; long long ea, ta;
		block_scan2(misc, c, a, ec, ea, tc, ta);
		misc->tb_cov[tid] = ec;
		misc->tb_val[tid] = ea;
		int kept = 0;
		for (int q = 0; q < V; q++) {
			const int r = r0 + q;
			if (r > nb) break;
			const int ps = pidx(r);
			klocal[ps] = (uint16_t) kept;
			if (r >= 1 && keep_rule(job, setop, cov[ps] + ec)) kept++;
		}
		int ek, tk;
		block_scan1(misc, kept, ek, tk);
		misc->kbase[tid] = ek;
	}
	__syncthreads();

	// ---- phase 5: emit, one slot per thread (coalesced stores) -----------------
	const bool sorted_op = (job.op == OP_MEDIAN || job.op == OP_MWU);
	const bool pearson = (job.op == OP_PEARSON);
	const long long out_base = job.woff[w];
	const int next_bp = misc->next_bp;
	Row row; row.job = &job; row.cur = cur; row.nxt = nxt; row.b = 0;
	unsigned long long slow = 0;
	double auc = 0;
	double mom[6] = {0, 0, 0, 0, 0, 0};

	for (int base = 1; base <= nb; base += THREADS) {
		const int r = base + tid;
		bool keep = false;
		int c = 0, ps = 0, owner = 0;
		if (r <= nb) {
			ps = pidx(r);
			owner = r >> vlog;
			c = cov[ps] + misc->tb_cov[owner];
			keep = keep_rule(job, setop, c);
		}
		const unsigned kb = __ballot_sync(FULL, keep);
		if (pearson) {
			if (keep) {
				const int bs = (int) (w0 + bpos[r]);
				const int be = (r < nb) ? (int) (w0 + bpos[r + 1]) : next_bp;
				double X, Y;
				row.b = bs;
				row.get(0, X); row.get(1, Y);
				const double len = (double) (be - bs);
				double one[6] = {len, X * len, Y * len, 0, 0, 0};
				merge6(mom, one);
			}
		} else if (!sorted_op) {
			if (keep) {
				const long long oi = out_base + misc->kbase[owner] + klocal[ps];
				const int bs = (int) (w0 + bpos[r]);
				const int be = (r < nb) ? (int) (w0 + bpos[r + 1]) : next_bp;
				double v;
				if (fast && (c >> COV_P_SHIFT) == 0) {
					const long long a = (long long) (((unsigned long long) limb[ls + ps] << 32) | limb[ps]) + misc->tb_val[owner];
					v = (double) (a + job.fx_base) * scale;
					if (job.op == OP_MEAN) v = v / N;
				} else {
					row.b = bs;
					v = row_eval(row);
					slow++;
				}
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
			unsigned todo = kb;
			while (todo) {
				const int src = __ffs(todo) - 1;
				todo &= todo - 1;
				const int rr = base + warp * 32 + src;
				const int bs = (int) (w0 + bpos[rr]);
				const int be = (rr < nb) ? (int) (w0 + bpos[rr + 1]) : next_bp;
				row.b = bs;
				const double v = warp_sorted_eval(row, buf, P);
				if (lane == 0) {
					const long long oi = out_base + misc->kbase[rr >> vlog] + klocal[pidx(rr)];
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

	// deterministic block reduction of the AUC partial / slow-path count / moments
	for (int d = 16; d > 0; d >>= 1) {
		auc += __shfl_down_sync(FULL, auc, d);
		slow += __shfl_down_sync(FULL, slow, d);
	}
	if (pearson) {
		for (int d = 16; d > 0; d >>= 1) {
			double other[6];
			for (int k = 0; k < 6; k++) other[k] = __shfl_down_sync(FULL, mom[k], d);
			if (lane + d < 32) merge6(mom, other);
		}
		if (lane == 0) for (int k = 0; k < 6; k++) misc->wmom[warp][k] = mom[k];
	}
	__syncthreads();   // all reads of tb_*/kbase done before misc->wsum is reused
	if (lane == 0) { misc->wauc[warp] = auc; misc->wsum[warp] = (long long) slow; }
	__syncthreads();
	if (tid == 0) {
		if (pearson && job.moments) {
			double t[6] = {0, 0, 0, 0, 0, 0};
			for (int k = 0; k < WARPS; k++) merge6(t, misc->wmom[k]);
			for (int k = 0; k < 6; k++) job.moments[(size_t) w * 6 + k] = t[k];
		}
		double t = 0; long long sl = 0;
		for (int k = 0; k < WARPS; k++) { t += misc->wauc[k]; sl += misc->wsum[k]; }
		if (job.auc_partial) job.auc_partial[w] = t;
		if (sl && job.slow_count) atomicAdd(job.slow_count, (unsigned long long) sl);
	}
	__syncthreads();
}

template <int LIMBS>
__global__ void __launch_bounds__(THREADS) emit_kernel(const Job job, int lw, double inv_scale, double lim, double scale, int rowbuf) {
	extern __shared__ __align__(16) unsigned char smem[];
	const EmitLayout L = make_emit_layout(job.wlog, job.N, LIMBS, rowbuf);
	int32_t *tn = (int32_t *) (smem + L.off_tn);
	long long *toff = (long long *) (smem + L.off_toff);
	long long *tmd = (long long *) (smem + L.off_tmd);
	for (int i = threadIdx.x; i < job.N; i += THREADS) {
		const long long o = job.off[i];
		toff[i] = o;
		tn[i] = (int) (job.off[i + 1] - o);
		tmd[i] = job.fast ? job.dflt_fx[i] : 0;
	}
	__syncthreads();
	for (int w = blockIdx.x; w < job.nw; w += gridDim.x)
		emit_window<LIMBS>(job, w, smem, L, lw, inv_scale, lim, scale);
}

// exclusive scan of the per-window counts (single CTA; nw is small)
__global__ void __launch_bounds__(1024) scan_counts_kernel(const uint32_t *wcount, int64_t *woff, int nw) {
	__shared__ long long wtot[32];
	__shared__ long long carry_s;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) carry_s = 0;
	__syncthreads();
	for (int base = 0; base < nw; base += 1024) {
		int i = base + tid;
		long long v = (i < nw) ? wcount[i] : 0, inc = v;
		for (int d = 1; d < 32; d <<= 1) { long long t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
		if (lane == 31) wtot[warp] = inc;
		__syncthreads();
		long long b = carry_s;
		for (int k = 0; k < warp; k++) b += wtot[k];
		if (i < nw) woff[i] = b + inc - v;
		__syncthreads();
		if (tid == 1023) carry_s = b + inc;
		__syncthreads();
	}
	if (tid == 0) woff[nw] = carry_s;
}

// max |v| over a strided sample of the values, and the largest exponent seen:
// chooses the fixed-point scale.  out2[0] = max finite |v|.
__global__ void sample_maxabs_kernel(const double *value, int64_t n, int64_t stride, double *out2) {
	double m = 0;
	for (int64_t i = (int64_t) (blockIdx.x * blockDim.x + threadIdx.x) * stride; i < n; i += (int64_t) gridDim.x * blockDim.x * stride) {
		double v = fabs(value[i]);
		if (v < DBL_MAX && v > m) m = v;   // skips NaN and Inf
	}
	for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_down_sync(FULL, m, d));
	if ((threadIdx.x & 31) == 0 && m > 0)
		atomicMax((unsigned long long *) out2, (unsigned long long) __double_as_longlong(m)); // positive doubles order like integers
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
	if (op != OP_MEDIAN && op != OP_MWU) return 0;
	int P = 1; while (P < N) P <<= 1;
	return P * (int) sizeof(SortKey) * WARPS;
}

size_t window_smem_bytes(int wlog, int N, bool emit, int op) {
	if (!emit) return make_count_layout(wlog, N).total;
	return make_emit_layout(wlog, N, 3, row_buffer_bytes(N, op)).total;
}

template <int VLOG>
static cudaError_t launch_count_v(const Job &job, cudaStream_t s) {
	size_t smem = make_count_layout(job.wlog, job.N).total;
	cudaError_t e = cudaFuncSetAttribute(count_kernel<VLOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (e != cudaSuccess) return e;
	int strips = (job.nw + job.strip - 1) / job.strip;
	count_kernel<VLOG><<<strips, THREADS, smem, s>>>(job);
	return cudaGetLastError();
}

cudaError_t launch_count(const Job &job, cudaStream_t s) {
	switch (job.wlog) {           // window = THREADS << VLOG base pairs
	case 8: return launch_count_v<0>(job, s);
	case 9: return launch_count_v<1>(job, s);
	case 10: return launch_count_v<2>(job, s);
	case 11: return launch_count_v<3>(job, s);
	case 12: return launch_count_v<4>(job, s);
	case 13: return launch_count_v<5>(job, s);
	}
	return cudaErrorInvalidValue;
}

cudaError_t launch_scan_counts(const uint32_t *wcount, int64_t *woff, int nw, cudaStream_t s) {
	scan_counts_kernel<<<1, 1024, 0, s>>>(wcount, woff, nw);
	return cudaGetLastError();
}

cudaError_t launch_emit(const Job &job, cudaStream_t s) {
	// fixed-point format: |m| < 2^mbits, limbs of lw bits
	int l2n = 1; while ((1 << l2n) < 2 * job.N) l2n++;
	int lw = 32 - l2n;
	int limbs = 2;
	if (job.fast && job.fx_mbits > 29 - l2n + lw) limbs = 3;
	double inv_scale = ldexp(1.0, -job.fx_q), scale = ldexp(1.0, job.fx_q), lim = ldexp(1.0, job.fx_mbits);
	int rowbuf = row_buffer_bytes(job.N, job.op);
	size_t smem = make_emit_layout(job.wlog, job.N, limbs, rowbuf).total;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaError_t e;
	int grid;
	if (limbs == 2) {
		e = cudaFuncSetAttribute(emit_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if (e != cudaSuccess) return e;
		int per = 1;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, emit_kernel<2>, THREADS, smem);
		grid = min(job.nw, sms * max(per, 1));
		emit_kernel<2><<<grid, THREADS, smem, s>>>(job, lw, inv_scale, lim, scale, rowbuf);
	} else {
		e = cudaFuncSetAttribute(emit_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if (e != cudaSuccess) return e;
		int per = 1;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, emit_kernel<3>, THREADS, smem);
		grid = min(job.nw, sms * max(per, 1));
		emit_kernel<3><<<grid, THREADS, smem, s>>>(job, lw, inv_scale, lim, scale, rowbuf);
	}
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

cudaError_t launch_reduce_auc(const double *partial, int nw, double *out, cudaStream_t s) {
	reduce_auc_kernel<<<1, 1024, 0, s>>>(partial, nw, out);
	return cudaGetLastError();
}

} // namespace wb

// exported for tests: the host build of the same routine the kernels use
extern "C" double wb200_tdist_Q(double t, double nu) { return wb::tdist_Q(t, nu); }
