// wiggletools_b200 -- host side of the C ABI declared in include/wiggletools_b200.h.
// Owns the CUDA context objects (stream, growable device/pinned buffers),
// stages host tracks to HBM, sequences the kernels of engine.cu and returns
// segments.  No CPU implementation of the path lives here: without a device
// wb200_create() fails.
#include "../../include/wiggletools_b200.h"
#include "engine.cuh"

#include <algorithm>
#include <cctype>
#include <sched.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

using namespace wb;

namespace {

std::string g_last_error;  // for failures before a context exists

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		cudaError_t e = cudaMalloc(&p, want);
		if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <class T> T *as() const { return (T *) p; }
};

struct HostBuf {
	void *p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFreeHost(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
	template <class T> T *as() const { return (T *) p; }
};

} // namespace

struct wb200_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = true;
	int sms = 148;
	int window_bp = 0;
	std::string err;
	wb200_stats stats{};
	cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
	DevBuf in_start, in_end, in_value, off, dflt, dflt_fx, cursor, state, aucp, small, moments, slow_list;
	DevBuf out_start, out_end, out_value;
	HostBuf h_small, h_meta, h_start, h_end, h_value;
	// batch pipeline (wb200_reduce_host_batch)
	cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
	DevBuf bin_start[2], bin_end[2], bin_value[2], bout_start[2], bout_end[2], bout_value[2];
	HostBuf h_arena;
	std::vector<cudaEvent_t> pipe_events;
	unsigned long long diag[16] = {0};      // diagnostic builds: accumulated phase counters
	// narrow transport: float32 value staging and breakpoint-form read-back (wb200_reduce_host_batch)
	bool breakpoint_out = false;
	DevBuf bin_value32[2], ends_val[2], ends_idx[2];
	// raw multiplexer rows of the last single-shard reduction (wb200_rows_host)
	Job last_job{};
	int64_t last_S = -1;
	DevBuf rows_v, rows_p;
	// regional integrators (wb200_apply_*)
	DevBuf ap_rs, ap_re, ap_codes, ap_out;
	HostBuf hap_out;
	// batch device entry (wb200_reduce_device_batch): per-shard tables and result slots
	DevBuf b_tables, b_small;
	HostBuf hb_tables, hb_small;
	std::vector<cudaEvent_t> batch_events;
	// text sink (wb200_format_*)
	DevBuf w_certain, w_heads, w_hp, w_tile_count, w_tile_base, w_small, w_bbytes, w_boff, w_text;
	DevBuf w_in_start, w_in_end, w_in_value;
	HostBuf hw_text, hw_blocks, hw_small;
	// element-wise input maps (wb200_set_input_maps): chains consumed by the next staging call
	std::vector<MapOp> map_chains;
	std::vector<int> map_nchain;
	DevBuf m_chains, m_nchain, m_keep, m_hp, m_tile_count, m_tile_base, m_small, m_newoff, m_off, in2_start, in2_end, in2_value;
	HostBuf hm_small;
};

namespace {

int fail(wb200_ctx *ctx, int code, const std::string &msg) {
	if (ctx) ctx->err = msg; else g_last_error = msg;
	return code;
}

int cuda_fail(wb200_ctx *ctx, cudaError_t e, const char *what) {
	return fail(ctx, e == cudaErrorMemoryAllocation ? WB200_ENOMEM : WB200_ECUDA,
	            std::string(what) + ": " + cudaGetErrorString(e));
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); } while (0)

// Derived default_value of each reducer's output: what the reference's
// constructors compute from the inputs' defaults (src/reducers.c:170-186,
// 237-253, 294-307, 348-361, 404-422, 481-505, 565-589, 640-666, 727-750,
// 815-834; set comparisons return NaN, src/setComparisons.c:130,389).
double derived_default(int op, int n, const double *d) {
	double sum = 0;
	switch (op) {
	case OP_SUM:
		for (int i = 0; i < n; i++) { if (std::isnan(d[i])) return NAN; sum += d[i]; }
		return sum;
	case OP_PRODUCT: {
		double p = 1;
		for (int i = 0; i < n; i++) { if (std::isnan(d[i])) return NAN; p *= d[i]; }
		return p;
	}
	case OP_MEAN: {
		for (int i = 0; i < n; i++) { if (std::isnan(d[i])) return NAN; sum += d[i]; }
		float f = (float) (sum / n);
		return f;
	}
	case OP_VAR: case OP_STDDEV: {
		for (int i = 0; i < n; i++) { if (std::isnan(d[i])) return NAN; sum += d[i]; }
		double mean = sum / n, err = 0;
		for (int i = 0; i < n; i++) err += (d[i] - mean) * (d[i] - mean);
		return op == OP_VAR ? err / n : std::sqrt(err / n);
	}
	case OP_ENTROPY: {
		int count = 0;
		for (int i = 0; i < n; i++) { if (std::isnan(d[i])) return NAN; if (d[i] != 0) count++; }
		double p = count / n;   // integer division, as in the reference
		return p ? -p * std::log(p) - (1 - p) * std::log(1 - p) : 0;
	}
	case OP_CV: {
		for (int i = 0; i < n; i++) { if (std::isnan(d[i])) return NAN; sum += d[i]; }
		double mean = sum / n, err = 0;
		for (int i = 0; i < n; i++) err += (mean - d[i]) * (mean - d[i]);
		float f = (float) (std::sqrt(err / n) / mean);
		return f;
	}
	case OP_MAX: case OP_MIN: {
		double m = d[0];
		if (std::isnan(m)) return m;
		for (int i = 1; i < n; i++) {
			if (std::isnan(d[i])) return NAN;
			if (op == OP_MAX ? d[i] > m : d[i] < m) m = d[i];
		}
		return m;
	}
	case OP_FILLIN: case OP_TRIMFILL:
		return d[1];                 // reducers.c:117
	case OP_MEDIAN: {
		std::vector<double> v(d, d + n);
		for (int i = 0; i < n; i++) if (std::isnan(v[i])) return NAN;
		std::sort(v.begin(), v.end());
		float f = (float) v[n / 2];
		return f;
	}
	}
	return NAN;
}

bool valid_op(int op) {
	return (op >= OP_SUM && op <= OP_TRIMFILL) || op == OP_TTEST || op == OP_MWU;
}

int ceil_log2(long long v) { int l = 0; while ((1ll << l) < v) l++; return l; }

// Host-side plan of one chromosome shard: argument checks (the reference's arity rules), window
// size, job geometry.  Buffers are attached by the caller.
int plan_shard(wb200_ctx *ctx, int op, int N, int &n1, int &strict, const int64_t *h_off, const double *h_dflt,
               const int32_t *d_start, const int32_t *d_end, const double *d_value,
               int32_t pos_lo, int32_t pos_hi, bool moments, int px, int py, Job &job) {
	if (!valid_op(op) && !(op == OP_PEARSON && moments && N >= 2)) return fail(ctx, WB200_EINVAL, "unknown op");
	if (N < 1) return fail(ctx, WB200_EINVAL, "need at least one track");
	const bool setop = (op == OP_TTEST || op == OP_MWU);
	const bool fill = (op == OP_FILLIN || op == OP_TRIMFILL);
	if (fill && N != 2) return fail(ctx, WB200_EINVAL, "The fill in operator can only work on 2 iterators!");
	if (fill) { n1 = 1; strict = 0; }      // cover counters: set A = track 0, set B = track 1
	else if (!setop) n1 = N;
	const int n2 = N - n1;
	if (n1 > MAX_TRACKS || n2 > MAX_TRACKS) return fail(ctx, WB200_EUNSUPPORTED, "more than 1023 tracks in a set");
	// the reference's arity checks (it prints and exit(1)s; here: an error code)
	if (op == OP_TTEST && (n1 < 3 || n2 < 3))
		return fail(ctx, WB200_EINVAL, "The t-test function only works for two sets with enough elements to compute variance");
	if (op == OP_MWU && (n1 < 1 || n2 < 1))
		return fail(ctx, WB200_EINVAL, "The Mann-Whitney U function only works for two non-empty sets");
	if (pos_hi < pos_lo) pos_hi = pos_lo;
	const int64_t I = h_off[N];
	for (int i = 0; i < N; i++)
		if (h_off[i + 1] < h_off[i] || h_off[i + 1] - h_off[i] > INT32_MAX)
			return fail(ctx, WB200_EINVAL, "bad track offsets");

	// window size
	int wlog = 12;
	if (ctx->window_bp > 0) wlog = ceil_log2(ctx->window_bp);
	else {
		const char *env = getenv("WB200_WINDOW");
		if (env && atoi(env) > 0) wlog = ceil_log2(atoi(env));
		else {
			double density = 2.0 * (double) I / ((double) pos_hi - pos_lo + 1.0);
			// dense data: smaller windows for the kernels that keep a window in registers; the sum / mean
			// sweep (fastsum.cu) is fastest with the largest window at any density
			if (density > 0.6 && !(op == OP_SUM || op == OP_MEAN)) wlog = 11;
			// var / stddev / CV on the sweep kernel: 16 B of shared memory per position, so the smaller window keeps
			// more windows in flight per SM (measured: N = 64, density 0.5: 6 instead of 3 CTAs per SM)
			if (density > 0.25 && (op == OP_VAR || op == OP_STDDEV || op == OP_ENTROPY || op == OP_CV)) wlog = 11;
			// the order-statistic sweep is a serial chain per warp: smaller windows, more CTAs per SM
			// (more than 24 tracks: the bucketed selection sorts a window's intervals once and keeps 32 byte counters
			// per position -- 1024-bp windows)
			if ((op == OP_MEDIAN || op == OP_MIN || op == OP_MAX) && N <= 512) wlog = N > 24 ? 10 : 11;
		}
	}
	wlog = std::min(TLOG + 4, std::max(TLOG + 2, wlog));
	{
		// the window must fit shared memory with the kernel that will actually run: sized with that kernel's
		// limb arrays (var / stddev / CV: up to 6; sum / mean outside the sweep kernel: up to 3; all others: 2)
		// and shrunk step by step before giving up (median / min / max / wilcoxon carry per-warp sort rows)
		const bool moment_like = (op == OP_VAR || op == OP_STDDEV || op == OP_ENTROPY || op == OP_CV);
		const int arrays = moment_like ? 6 : (op == OP_SUM || op == OP_MEAN) ? 3 : 2;
		while (wlog > TLOG + 2 && window_smem_bytes(wlog, N, true, op, arrays) > 227 * 1024) wlog--;
		if (window_smem_bytes(wlog, N, true, op, arrays) > 227 * 1024)
			return fail(ctx, WB200_EUNSUPPORTED, op == OP_MWU ? "wilcoxon over more than 1024 tracks in total does not fit shared memory (per-warp sort rows)"
			                                                  : "window does not fit shared memory for this many tracks");
	}
	const int64_t span = (int64_t) pos_hi - pos_lo;
	const int nw = (int) (span >> wlog) + 1;

	job = Job{};
	job.start = d_start; job.end = d_end; job.value = d_value;
	job.I = I;
	{
		const char *e = getenv("WB200_VEC");
		const bool aligned = (((uintptr_t) d_start | (uintptr_t) d_end | (uintptr_t) d_value) & 15) == 0;
		job.vec = aligned && !(e && atoi(e) == 0);
	}
	job.N = N; job.n1 = n1; job.strict = strict ? 1 : 0; job.op = op;
	job.p0 = pos_lo; job.wlog = wlog; job.nw = nw; job.pos_hi = pos_hi;
	job.cstride = (N + 3) & ~3;
	job.px = px; job.py = py;
	{ const char *e = getenv("WB200_LB_WARPS"); job.lb_warps = e ? std::max(1, std::min(WARPS, atoi(e))) : 1; }
	if (op == OP_MWU) {
		// integer division in both moments, src/setComparisons.c:386-387
		job.mwu_mu = (double) (n1 * n2 / 2);
		job.mwu_sigma = std::sqrt((double) (n1 * n2 * (n1 + n2 + 1) / 12));
	}
	return WB200_OK;
}

// whether the statistic has an exact fixed-point path for these defaults (then the values are sampled)
bool wants_format(int op, int N, int64_t I, const double *h_dflt, bool force_row) {
	const bool sum_like = (op == OP_SUM || op == OP_MEAN);
	const bool moment_like = (op == OP_VAR || op == OP_STDDEV || op == OP_ENTROPY || op == OP_CV);
	if (!(sum_like || moment_like) || I <= 0 || force_row) return false;
	for (int i = 0; i < N; i++) if (!std::isfinite(h_dflt[i]) || (moment_like && h_dflt[i] != 0)) return false;
	return true;
}

// Fixed-point format of the exact paths from the sampled value range: sets job.fast / fx_*,
// fills hm_fx[N] (fixed-point defaults) when the sum path is chosen.
void choose_format(Job &job, int op, int N, const double *h_dflt, double maxabs, long long lsb_key, int min_limbs, int64_t *hm_fx) {
	const bool sum_like = (op == OP_SUM || op == OP_MEAN);
	job.fast = 0;
	for (int i = 0; i < N; i++) maxabs = std::max(maxabs, std::fabs(h_dflt[i]));
	int emax = 0;
	if (maxabs > 0) { std::frexp(maxabs, &emax); }   // maxabs < 2^emax
	const int l2n = ceil_log2(2 * (int64_t) N);
	const int lw = 32 - l2n;
	if (sum_like) {
		// format: |m| < 2^mbits in one, two or three 32-bit limbs per position.
		//  one  : the sampled values have no bit below 2^q1, q1 = emax + 2 - (29 - l2n): a single
		//         32-bit accumulator holds every per-position delta sum (integer-like and
		//         coarse fixed-point data, e.g. read counts);
		//  two  : float32-grade inputs (24-bit mantissas) over a moderate exponent range;
		//  three: anything wider.
		// A value finer than the chosen 2^q poisons its interval; those segments are recomputed
		// by the fix-up kernel, and a run with too many of them is repeated one format wider.
		const int cap1 = 29 - l2n;
		const int cap2 = 29 - l2n + lw;
		const int cap3 = std::min(29 - l2n + 2 * lw, 62 - ceil_log2(N + 1));
		const char *env = getenv("WB200_LIMBS");
		const int forced = env ? atoi(env) : 0;
		const int lsb_q = lsb_key > 0 ? (int) (2048 - lsb_key) : 4096;          // finest significant bit seen
		int first = std::max(min_limbs, (forced >= 1 && forced <= 3) ? forced : (cap1 >= 8 && lsb_q >= emax + 2 - cap1 ? 1 : (cap2 >= 40 ? 2 : 3)));
		bool ok = false;
		for (int limbs = first; limbs <= 3 && !ok; limbs++) {
			const int mbits = limbs == 1 ? cap1 : limbs == 2 ? cap2 : cap3;
			if (mbits < 4) continue;
			job.fx_mbits = mbits;
			job.fx_limbs = limbs;
			job.fx_q = emax + 2 - mbits;           // two bits of headroom above the sampled maximum
			job.fx_q = std::max(-1000, std::min(1000, job.fx_q));
			const double inv = std::ldexp(1.0, -job.fx_q), lim = std::ldexp(1.0, mbits);
			long long base = 0;
			ok = true;
			for (int i = 0; i < N && ok; i++) {
				double t = h_dflt[i] * inv;
				if (!(std::fabs(t) < lim) || t != std::rint(t)) ok = false;
				else { hm_fx[i] = (int64_t) t; base += hm_fx[i]; }
			}
			if (ok) { job.fast = 1; job.fx_base = base; }
		}
	} else {
		// squares must stay exact in 64 bits: m = v 2^-q with q the finest bit seen in
		// the sample, |m| < 2^mbits, N m^2 < 2^62.  Finer or larger values poison.
		const int capm = std::min(30, (61 - ceil_log2(N + 1)) / 2);
		const int q = lsb_key > 0 ? (int) (2048 - lsb_key) : 0;
		if (emax - q <= capm && 2 * capm <= 29 - l2n + 2 * lw) {
			job.fast = 2;
			job.fx_q = std::max(-1000, std::min(1000, q));
			// |m| < 2^mbits: two bits of headroom above the sampled range instead of the widest format the
			// arithmetic allows -- the squares then usually fit two limbs instead of three (4 shared atomics
			// per event instead of 6); a value beyond the range poisons its interval and is fixed up exactly
			job.fx_mbits = std::max(4, std::min(capm, emax - q + 2));
			job.fx_base = 0;
			// every prefix of the plain sum then fits 31 bits: one limb (one shared atomic per event) for it
			job.sum1 = job.fx_mbits + ceil_log2(N) <= 30 ? 1 : 0;
		}
	}
}

// Core: inputs already resident on the device.
int run_device(wb200_ctx *ctx, int op, int N, int n1, int strict,
               const int32_t *d_start, const int32_t *d_end, const double *d_value,
               const int64_t *h_off, const double *h_dflt, int32_t pos_lo, int32_t pos_hi,
               int32_t *d_out_start, int32_t *d_out_end, double *d_out_value, int64_t out_cap,
               int64_t *n_out, double *default_out, double *auc_out, double *moments_out = nullptr, bool force_row = false,
               int px = 0, int py = 1, int min_limbs = 1) {
	Job job{};
	{
		int rc = plan_shard(ctx, op, N, n1, strict, h_off, h_dflt, d_start, d_end, d_value, pos_lo, pos_hi, moments_out != nullptr, px, py, job);
		if (rc != WB200_OK) return rc;
	}
	ctx->last_S = -1;
	const bool setop = (op == OP_TTEST || op == OP_MWU);
	const int64_t I = job.I;
	const int nw = job.nw, wlog = job.wlog;
	job.out_start = d_out_start; job.out_end = d_out_end; job.out_value = d_out_value; job.out_cap = out_cap;

	cudaStream_t s = ctx->stream;
	wb200_stats &st = ctx->stats;
	st = wb200_stats{};
	st.intervals = I;

	// small tables
	CK(ctx->off.ensure((N + 1) * sizeof(int64_t)));
	CK(ctx->dflt.ensure(N * sizeof(double)));
	CK(ctx->dflt_fx.ensure(N * sizeof(int64_t)));
	CK(ctx->h_meta.ensure((size_t) (N + 1) * 8 + (size_t) N * 16));
	int64_t *hm_off = ctx->h_meta.as<int64_t>();
	double *hm_dflt = (double *) (hm_off + N + 1);
	int64_t *hm_fx = (int64_t *) (hm_dflt + N);
	memcpy(hm_off, h_off, (N + 1) * sizeof(int64_t));
	memcpy(hm_dflt, h_dflt, N * sizeof(double));
	CK(ctx->cursor.ensure((size_t) (nw + 1) * job.cstride * sizeof(int32_t)));
	CK(ctx->state.ensure((size_t) nw * sizeof(unsigned long long)));
	CK(ctx->aucp.ensure((size_t) nw * WARPS * sizeof(double)));
	CK(ctx->small.ensure(256));
	CK(ctx->h_small.ensure(256));
	if (moments_out) { CK(ctx->moments.ensure((size_t) nw * 6 * sizeof(double))); job.moments = ctx->moments.as<double>(); }
	job.off = ctx->off.as<int64_t>(); job.dflt = ctx->dflt.as<double>(); job.dflt_fx = ctx->dflt_fx.as<int64_t>();
	job.cursor = ctx->cursor.as<int32_t>(); job.state = ctx->state.as<unsigned long long>();
	job.auc_partial = ctx->aucp.as<double>();
	// small: [0..1] sample result, [2] auc, [3] slow count
	double *d_small = ctx->small.as<double>();
	job.slow_count = (unsigned long long *) (d_small + 3);
	job.auc_fix = d_small + 10;
	job.total = (long long *) (d_small + 11);
	job.ticket = (unsigned int *) (d_small + 12);
	job.flags = (int *) (d_small + 13);
	job.diag = (unsigned long long *) (d_small + 16);

	CK(cudaEventRecord(ctx->ev[0], s));
	CK(cudaMemcpyAsync(ctx->off.p, hm_off, (N + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
	CK(cudaMemcpyAsync(ctx->dflt.p, hm_dflt, N * sizeof(double), cudaMemcpyHostToDevice, s));
	CK(cudaMemsetAsync(d_small, 0, 256, s));
	st.kernels = 0;

	// the cursor table does not depend on the accumulator format: it is built while the host
	// waits for the value sample below
	{
		CK(cudaEventRecord(ctx->ev[1], s));
		int64_t max_len = 0;
		for (int i = 0; i < N; i++) max_len = std::max<int64_t>(max_len, h_off[i + 1] - h_off[i]);
		CK(cudaMemsetAsync(ctx->state.p, 0, (size_t) nw * sizeof(unsigned long long), s));
		CK(launch_cursor(job, max_len, s));
		CK(cudaEventRecord(ctx->ev[2], s));
	}

	// fixed-point fast paths: sum / mean (one exact accumulator) and var / stddev /
	// CV (exact sum and sum of squares)
	job.fast = 0;
	if (wants_format(op, N, I, h_dflt, force_row)) {
		int64_t stride = std::max<int64_t>(1, I / (1 << 16));
		CK(launch_sample_maxabs(d_value, I, stride, d_small, s));
		st.kernels++;
		CK(cudaMemcpyAsync(ctx->h_small.p, d_small, 16, cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		choose_format(job, op, N, h_dflt, ctx->h_small.as<double>()[0], ctx->h_small.as<long long>()[1], min_limbs, hm_fx);
		if (job.fast == 1) CK(cudaMemcpyAsync(ctx->dflt_fx.p, hm_fx, N * sizeof(int64_t), cudaMemcpyHostToDevice, s));
	}
	if (job.fast) {
		job.slow_cap = std::max<int64_t>(1 << 16, out_cap / 8);
		CK(ctx->slow_list.ensure((size_t) job.slow_cap * sizeof(long long)));
		job.slow_list = ctx->slow_list.as<long long>();
	}
	(void) wlog;      // (shared-memory fit was settled by plan_shard)

	CK(cudaEventRecord(ctx->ev[4], s));
	int auc_per_window = WARPS;
	CK(launch_emit(job, s, &auc_per_window));
	CK(launch_reduce_auc(job.auc_partial, nw * auc_per_window, d_small + 2, s));
	if (moments_out) { CK(launch_reduce_moments(job.moments, nw, d_small + 4, s)); st.kernels++; }
	CK(cudaEventRecord(ctx->ev[3], s));
	st.kernels += 3;

	// S, AUC, slow-path count
	int64_t *h_small = ctx->h_small.as<int64_t>();
	CK(cudaMemcpyAsync(h_small, d_small, 256, cudaMemcpyDeviceToHost, s));     // [2] auc [3] slow [4..9] moments [10] auc_fix [11] S [13] flags [16..] diag
	CK(cudaStreamSynchronize(s));
	const int64_t S = h_small[11];
	for (int k = 0; k < 16; k++) ctx->diag[k] += (unsigned long long) h_small[16 + k];   // zeros unless built with -DWB_DIAG_PHASES
	if (const int fl = (int) (h_small[13] & 3)) {
		ctx->last_S = -1;
		return fail(ctx, WB200_EINVAL, (fl & 1) ? "a track is not sorted, overlaps itself or holds an empty interval (start >= end)"
		                                        : "an interval lies outside [pos_lo, pos_hi]");
	}
	if (job.fast && h_small[3] > 0) {
		// segments the fixed-point path could not represent exactly
		// too many for the fix-up kernel: repeat one accumulator format wider (the sample missed
		// values with finer bits), and on the row path when the widest format is not enough
		const bool narrow = job.fast == 1 && job.fx_limbs < 3;
		if (narrow && S <= out_cap && h_small[3] > std::max<int64_t>(1 << 12, S >> 8))
			return run_device(ctx, op, N, n1, strict, d_start, d_end, d_value, h_off, h_dflt, pos_lo, pos_hi,
			                  d_out_start, d_out_end, d_out_value, out_cap, n_out, default_out, auc_out, moments_out, false, px, py,
			                  job.fx_limbs + 1);
		if (h_small[3] > job.slow_cap || S > out_cap)
			return run_device(ctx, op, N, n1, strict, d_start, d_end, d_value, h_off, h_dflt, pos_lo, pos_hi,
			                  d_out_start, d_out_end, d_out_value, out_cap, n_out, default_out, auc_out, moments_out, true, px, py);
		CK(launch_fixup(job, h_small[3], s));
		st.kernels++;
		CK(cudaMemcpyAsync(h_small + 10, d_small + 10, 8, cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		double a, f;
		memcpy(&a, h_small + 2, 8); memcpy(&f, h_small + 10, 8);
		a += f;
		memcpy(h_small + 2, &a, 8);
	}
	st.segments = S;
	st.windows = nw;
	st.fast_path = job.fast;
	st.limbs = job.fast == 1 ? job.fx_limbs : 0;
	st.sweep_kernel = auc_per_window == 1 ? 1 : 0;
	st.slow_segments = h_small[3];
	cudaEventElapsedTime(&st.ms_count, ctx->ev[1], ctx->ev[2]);
	cudaEventElapsedTime(&st.ms_emit, ctx->ev[4], ctx->ev[3]);
	cudaEventElapsedTime(&st.ms_total, ctx->ev[0], ctx->ev[3]);
	ctx->last_job = job; ctx->last_S = (S > out_cap || !d_out_start) ? -1 : S;      // inputs, cursor table and segments stay resident until the next call
	if (n_out) *n_out = S;
	if (auc_out) memcpy(auc_out, h_small + 2, 8);
	if (default_out) *default_out = setop ? NAN : derived_default(op, N, h_dflt);
	if (moments_out) { memcpy(moments_out, h_small + 4, 6 * sizeof(double)); return WB200_OK; }
	if (S > out_cap) return fail(ctx, WB200_ECAPACITY, "output capacity too small");
	return WB200_OK;
}

} // namespace

extern "C" {

wb200_ctx *wb200_create(int device) {
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		g_last_error = std::string("wiggletools_b200: no usable CUDA device (") +
		               (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
		               "); this engine has no CPU fallback";
		return nullptr;
	}
	if (device < 0 || device >= count) { g_last_error = "wiggletools_b200: bad device ordinal"; return nullptr; }
	if ((e = cudaSetDevice(device)) != cudaSuccess) { g_last_error = cudaGetErrorString(e); return nullptr; }
	wb200_ctx *ctx = new wb200_ctx();
	ctx->device = device;
	cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device);
	if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
		g_last_error = cudaGetErrorString(e);
		delete ctx;
		return nullptr;
	}
	for (auto &ev : ctx->ev) cudaEventCreate(&ev);
	return ctx;
}

void wb200_destroy(wb200_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	DevBuf *d[] = {&ctx->in_start, &ctx->in_end, &ctx->in_value, &ctx->off, &ctx->dflt, &ctx->dflt_fx, &ctx->cursor,
	               &ctx->state, &ctx->aucp, &ctx->small, &ctx->moments, &ctx->slow_list, &ctx->out_start, &ctx->out_end, &ctx->out_value};
	for (auto *b : d) b->release();
	DevBuf *wd[] = {&ctx->w_certain, &ctx->w_heads, &ctx->w_hp, &ctx->w_tile_count, &ctx->w_tile_base, &ctx->w_small, &ctx->w_bbytes,
	                &ctx->w_boff, &ctx->w_text, &ctx->w_in_start, &ctx->w_in_end, &ctx->w_in_value};
	for (auto *b : wd) b->release();
	ctx->hw_text.release(); ctx->hw_blocks.release(); ctx->hw_small.release(); ctx->hm_small.release();
	DevBuf *md[] = {&ctx->m_chains, &ctx->m_nchain, &ctx->m_keep, &ctx->m_hp, &ctx->m_tile_count, &ctx->m_tile_base, &ctx->m_small,
	                &ctx->m_newoff, &ctx->m_off, &ctx->in2_start, &ctx->in2_end, &ctx->in2_value};
	for (auto *b : md) b->release();
	for (int k = 0; k < 2; k++) {
		ctx->bin_start[k].release(); ctx->bin_end[k].release(); ctx->bin_value[k].release();
		ctx->bout_start[k].release(); ctx->bout_end[k].release(); ctx->bout_value[k].release();
	}
	for (auto ev : ctx->pipe_events) cudaEventDestroy(ev);
	for (auto ev : ctx->batch_events) cudaEventDestroy(ev);
	ctx->b_tables.release(); ctx->b_small.release(); ctx->hb_tables.release(); ctx->hb_small.release();
	ctx->rows_v.release(); ctx->rows_p.release();
	for (int k = 0; k < 2; k++) { ctx->bin_value32[k].release(); ctx->ends_val[k].release(); ctx->ends_idx[k].release(); }
	ctx->ap_rs.release(); ctx->ap_re.release(); ctx->ap_codes.release(); ctx->ap_out.release(); ctx->hap_out.release();
	if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
	if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
	HostBuf *h[] = {&ctx->h_small, &ctx->h_meta, &ctx->h_start, &ctx->h_end, &ctx->h_value, &ctx->h_arena};
	for (auto *b : h) b->release();
	for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
	if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

const char *wb200_last_error(const wb200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }
const wb200_stats *wb200_last_stats(const wb200_ctx *ctx) { return &ctx->stats; }

int wb200_set_stream(wb200_ctx *ctx, void *cuda_stream) {
	if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
	ctx->stream = (cudaStream_t) cuda_stream;
	ctx->own_stream = false;
	return WB200_OK;
}

// -- text sink ------------------------------------------------------------
// Replaces, for one chromosome of segments: CompressionWiggleIterator (src/unaryOps.c:235-263,
// `write` only) + the Tee block hand-over (src/wigWriter.c:159-200) + printBlock (:55-119).
int wb200_format_device(wb200_ctx *ctx, int mode, const char *chrom,
                        const int32_t *d_start, const int32_t *d_end, const double *d_value, int64_t n,
                        wb200_sink_state *state, const char **text, int64_t *nbytes) {
	if (!ctx || !chrom || !state || !text || !nbytes || n < 0 || n > INT32_MAX - 1) return WB200_EINVAL;
	if (mode != WB200_SINK_WIG && mode != WB200_SINK_BEDGRAPH) return fail(ctx, WB200_EINVAL, "unknown sink mode");
	if (strlen(chrom) > 63) return fail(ctx, WB200_EUNSUPPORTED, "chromosome name longer than 63 bytes");
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	*text = "";
	*nbytes = 0;
	if (n == 0) return WB200_OK;
	const int wig = mode == WB200_SINK_WIG;
	const int BR = sink_block_records();
	CK(ctx->w_small.ensure(64));
	CK(ctx->hw_small.ensure(64));
	int *d_flags = ctx->w_small.as<int>();                       // [0] unsupported value, [1] state after the last record
	long long *d_R = (long long *) (ctx->w_small.as<char>() + 16);
	CK(cudaMemsetAsync(ctx->w_small.p, 0, 64, s));
	// 1. records: every segment (write_bg) or the heads of the compress rule (write)
	int64_t R = n;
	const int32_t *hp = nullptr;
	if (wig) {
		const int ntiles = sink_tiles(n);
		CK(ctx->w_certain.ensure((size_t) n)); CK(ctx->w_heads.ensure((size_t) n));
		CK(ctx->w_hp.ensure((size_t) (n + 1) * sizeof(int32_t)));
		CK(ctx->w_tile_count.ensure((size_t) ntiles * sizeof(int))); CK(ctx->w_tile_base.ensure((size_t) ntiles * sizeof(long long)));
		CK(sink_heads(d_start, d_end, d_value, n, ctx->w_certain.as<unsigned char>(), ctx->w_heads.as<unsigned char>(),
		              ctx->w_tile_count.as<int>(), ctx->w_tile_base.as<long long>(), d_R, ctx->w_hp.as<int32_t>(), s));
		CK(cudaMemcpyAsync(ctx->hw_small.p, d_R, 8, cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		R = ctx->hw_small.as<long long>()[0];
		hp = ctx->w_hp.as<int32_t>();
	}
	// 2. measuring pass: bytes per writer block of 10000 records (blocks are counted over the whole sink)
	const int64_t base = state->records;
	const int nblocks = (int) ((base % BR + R + BR - 1) / BR);
	CK(ctx->w_bbytes.ensure((size_t) nblocks * sizeof(int))); CK(ctx->w_boff.ensure((size_t) nblocks * sizeof(long long)));
	CK(ctx->hw_blocks.ensure((size_t) nblocks * (sizeof(int) + sizeof(long long))));
	CK(sink_format(false, d_start, d_end, d_value, hp, R, base, wig, state->point_by_point, chrom,
	               ctx->w_bbytes.as<int>(), nullptr, nullptr, d_flags, nblocks, s));
	long long *h_off = ctx->hw_blocks.as<long long>();
	int *h_bytes = (int *) (h_off + nblocks);
	CK(cudaMemcpyAsync(h_bytes, ctx->w_bbytes.p, (size_t) nblocks * sizeof(int), cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(ctx->hw_small.as<char>() + 8, d_flags, 8, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	const int *h_flags = (const int *) (ctx->hw_small.as<char>() + 8);
	if (h_flags[0])
		return fail(ctx, WB200_EUNSUPPORTED, "a value of 2^63 or more in magnitude: outside the device formatter's range");
	long long total = 0;
	for (int b = 0; b < nblocks; b++) { h_off[b] = total; total += h_bytes[b]; }
	// 3. writing pass
	CK(ctx->w_text.ensure((size_t) total + 32));
	CK(ctx->hw_text.ensure((size_t) total + 32));
	CK(cudaMemcpyAsync(ctx->w_boff.p, h_off, (size_t) nblocks * sizeof(long long), cudaMemcpyHostToDevice, s));
	CK(sink_format(true, d_start, d_end, d_value, hp, R, base, wig, state->point_by_point, chrom,
	               nullptr, ctx->w_boff.as<long long>(), ctx->w_text.as<char>(), d_flags, nblocks, s));
	CK(cudaMemcpyAsync(ctx->hw_text.p, ctx->w_text.p, (size_t) total, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	ctx->hw_text.as<char>()[total] = 0;
	state->records = base + R;
	state->point_by_point = h_flags[1];
	*text = ctx->hw_text.as<char>();
	*nbytes = total;
	ctx->stats.d2h_bytes = total;
	ctx->stats.kernels = wig ? 7 : 2;
	return WB200_OK;
}

int wb200_format_host(wb200_ctx *ctx, int mode, const char *chrom,
                      const int32_t *start, const int32_t *end, const double *value, int64_t n,
                      wb200_sink_state *state, const char **text, int64_t *nbytes) {
	if (!ctx || n < 0 || (n > 0 && (!start || !end || !value))) return WB200_EINVAL;
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	const size_t m = (size_t) std::max<int64_t>(1, n);
	CK(ctx->w_in_start.ensure(m * sizeof(int32_t))); CK(ctx->w_in_end.ensure(m * sizeof(int32_t))); CK(ctx->w_in_value.ensure(m * sizeof(double)));
	if (n > 0) {
		CK(cudaMemcpyAsync(ctx->w_in_start.p, start, (size_t) n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(ctx->w_in_end.p, end, (size_t) n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(ctx->w_in_value.p, value, (size_t) n * sizeof(double), cudaMemcpyHostToDevice, s));
	}
	return wb200_format_device(ctx, mode, chrom, ctx->w_in_start.as<int32_t>(), ctx->w_in_end.as<int32_t>(),
	                           ctx->w_in_value.as<double>(), n, state, text, nbytes);
}

namespace {
int reduce_host_to_device(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict, const wb200_track *tracks,
                          int64_t *S_out, double *dv_out, double *auc_out);
}

// Reduction and sink in one call: the segments of the chromosome never leave HBM.
int wb200_reduce_format_host(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict, const wb200_track *tracks,
                             int mode, const char *chrom, wb200_sink_state *state,
                             const char **text, int64_t *nbytes, int64_t *n_segments, double *auc_out) {
	if (!ctx || !tracks || !state || !text || !nbytes) return WB200_EINVAL;
	int64_t S = 0;
	double dv = NAN, auc = 0;
	int rc = reduce_host_to_device(ctx, op, ntracks, nset1, strict, tracks, &S, &dv, &auc);
	if (rc != WB200_OK) return rc;
	const int64_t h2d = ctx->stats.h2d_bytes;
	rc = wb200_format_device(ctx, mode, chrom, ctx->out_start.as<int32_t>(), ctx->out_end.as<int32_t>(),
	                         ctx->out_value.as<double>(), S, state, text, nbytes);
	ctx->stats.h2d_bytes = h2d;
	if (n_segments) *n_segments = S;
	if (auc_out) *auc_out = auc;
	return rc;
}

// Host build of the device formatter's "%lf" (same source function): for CPU-side fuzzing against printf.
int wb200_debug_format_value(double v, char *out32) { return out32 ? sink_debug_format_value(v, out32) : WB200_EINVAL; }

// -- element-wise maps on the inputs ---------------------------------------
int wb200_set_input_maps(wb200_ctx *ctx, int ntracks, const wb200_map *const *maps, const int *nmaps) {
	if (!ctx || ntracks < 0) return WB200_EINVAL;
	ctx->map_chains.clear(); ctx->map_nchain.clear();
	if (ntracks == 0 || !maps || !nmaps) return WB200_OK;
	bool any = false;
	ctx->map_chains.assign((size_t) ntracks * MAX_MAP_CHAIN, MapOp{0, 0, 0.0});
	ctx->map_nchain.assign(ntracks, 0);
	for (int i = 0; i < ntracks; i++) {
		if (nmaps[i] < 0 || nmaps[i] > MAX_MAP_CHAIN) { ctx->map_chains.clear(); ctx->map_nchain.clear(); return fail(ctx, WB200_EUNSUPPORTED, "more than 8 maps on one track"); }
		for (int c = 0; c < nmaps[i]; c++) {
			const wb200_map &m = maps[i][c];
			if (m.op < WB200_MAP_SCALE || m.op > WB200_MAP_ABS) { ctx->map_chains.clear(); ctx->map_nchain.clear(); return fail(ctx, WB200_EINVAL, "unknown map"); }
			ctx->map_chains[(size_t) i * MAX_MAP_CHAIN + c] = MapOp{m.op, 0, m.arg};
			any = true;
		}
		ctx->map_nchain[i] = nmaps[i];
	}
	if (!any) { ctx->map_chains.clear(); ctx->map_nchain.clear(); }
	return WB200_OK;
}

double wb200_map_default(int op, double arg, double default_value) {
	// the constructors' derived default_value (src/unaryOps.c:673-682, 736-745, 787-811, 843-866, 891-901, 951-960)
	const double d = default_value;
	switch (op) {
	case WB200_MAP_SCALE: return std::isnan(d) ? NAN : (double) (float) (d * arg);
	case WB200_MAP_OFFSET: return std::isnan(d) ? NAN : (double) (float) (d + arg);
	case WB200_MAP_LOG: return (!std::isnan(d) && d > 0) ? std::log(d) / arg : NAN;
	case WB200_MAP_EXP: return std::isnan(d) ? NAN : (double) (float) std::exp(d * arg);
	case WB200_MAP_POW: return (!std::isnan(d) && (d > 0 || arg > 0)) ? std::pow(d, arg) : NAN;
	case WB200_MAP_ABS: return std::isnan(d) ? NAN : std::fabs(d);
	}
	return NAN;
}

// The Student-t upper tail as the DEVICE evaluates it inside the t-test kernel (same source function
// as wb200_tdist_Q, compiled for sm_100a): out[i] = Q(t[i], nu[i]).  For tests that pin the device
// arithmetic on a high-precision table (the reference calls gsl_cdf_tdist_Q, setComparisons.c:117).
// Shared memory one CTA of the window kernel asks for (statistics outside the sum / mean / moments sweep kernel):
// host arithmetic only, no device needed -- the CPU tests hold the budgets of DESIGN 3b against it.
long long wb200_debug_window_smem_bytes(int op, int n_tracks, int window_bp) {
	if (n_tracks < 1 || window_bp < 1) return WB200_EINVAL;
	const bool moment_like = (op == OP_VAR || op == OP_STDDEV || op == OP_ENTROPY || op == OP_CV);
	const int arrays = moment_like ? 6 : (op == OP_SUM || op == OP_MEAN) ? 3 : 2;
	return (long long) window_smem_bytes(ceil_log2(window_bp), n_tracks, true, op, arrays);
}

int wb200_debug_tdist_device(wb200_ctx *ctx, int n, const double *t, const double *nu, double *out) {
	if (!ctx || n < 0 || (n > 0 && (!t || !nu || !out))) return WB200_EINVAL;
	if (n == 0) return WB200_OK;
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	CK(ctx->rows_v.ensure((size_t) 3 * n * sizeof(double)));
	double *d = ctx->rows_v.as<double>();
	CK(cudaMemcpyAsync(d, t, n * sizeof(double), cudaMemcpyHostToDevice, s));
	CK(cudaMemcpyAsync(d + n, nu, n * sizeof(double), cudaMemcpyHostToDevice, s));
	CK(launch_tdist(d, d + n, n, d + 2 * n, s));
	CK(cudaMemcpyAsync(out, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	return WB200_OK;
}

// Diagnostic builds (-DWB_DIAG_PHASES): cycles per pipeline phase summed over windows since the
// last call; all zero in the shipped build.
void wb200_debug_phases(wb200_ctx *ctx, uint64_t out[16]) {
	for (int k = 0; k < 16; k++) { out[k] = ctx->diag[k]; ctx->diag[k] = 0; }
}

// Pin the calling host thread to the CPUs that are local to the context's GPU (the NUMA node
// its PCIe root hangs off, /sys/bus/pci/devices/<id>/local_cpulist): pinned staging buffers
// allocated and touched afterwards land in that node's memory.  Returns the number of CPUs in
// the mask, 0 when the topology is not exposed (nothing changed), < 0 on error.
int wb200_bind_host_thread(wb200_ctx *ctx) {
	if (!ctx) return WB200_EINVAL;
	char bus[32] = {0};
	if (cudaDeviceGetPCIBusId(bus, sizeof(bus), ctx->device) != cudaSuccess) return 0;
	for (char *c = bus; *c; c++) *c = (char) tolower(*c);
	char path[128];
	snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
	FILE *f = fopen(path, "r");
	if (!f) return 0;
	char line[4096] = {0};
	const bool got = fgets(line, sizeof(line), f) != nullptr;
	fclose(f);
	if (!got) return 0;
	cpu_set_t set;
	CPU_ZERO(&set);
	int count = 0;
	for (char *tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
		int a = 0, b = 0;
		const int k = sscanf(tok, "%d-%d", &a, &b);
		if (k < 1) continue;
		if (k == 1) b = a;
		for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); count++; }
	}
	if (count == 0) return 0;
	if (sched_setaffinity(0, sizeof(set), &set) != 0) return fail(ctx, WB200_EINVAL, "sched_setaffinity failed");
	return count;
}

int wb200_set_breakpoint_output(wb200_ctx *ctx, int on) {
	if (!ctx) return WB200_EINVAL;
	ctx->breakpoint_out = on != 0;
	return WB200_OK;
}

int wb200_set_window(wb200_ctx *ctx, int window_bp) {
	if (window_bp < 0 || (window_bp & (window_bp - 1))) return fail(ctx, WB200_EINVAL, "window must be a power of two");
	ctx->window_bp = window_bp;
	return WB200_OK;
}

void *wb200_host_alloc(int64_t bytes) {
	void *p = nullptr;
	if (cudaHostAlloc(&p, (size_t) bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
	return p;
}

void wb200_host_free(void *p) { if (p) cudaFreeHost(p); }

int wb200_reduce_device(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict,
                        const int32_t *d_start, const int32_t *d_end, const double *d_value,
                        const int64_t *h_off, const double *h_defaults, int32_t pos_lo, int32_t pos_hi,
                        int32_t *d_out_start, int32_t *d_out_end, double *d_out_value,
                        int64_t out_cap, int64_t *n_out, double *default_out, double *auc_out) {
	if (!ctx) return WB200_EINVAL;
	cudaSetDevice(ctx->device);
	return run_device(ctx, op, ntracks, nset1, strict, d_start, d_end, d_value, h_off, h_defaults, pos_lo, pos_hi,
	                  d_out_start, d_out_end, d_out_value, out_cap, n_out, default_out, auc_out);
}

// Several chromosome shards resident in HBM, one call: the value samples of all shards are read
// with ONE synchronisation (one fixed-point format for the whole batch), then cursor kernel and
// sweep of every shard are queued back to back on the stream and all results are read at the end.
// Shards the optimistic path cannot finish (segments to fix up, capacity) are repeated one by one.
int wb200_reduce_device_batch(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict, const double *h_defaults,
                              int nshards, wb200_device_shard *shards, double *default_out) {
	if (!ctx || !shards || !h_defaults || nshards < 0) return WB200_EINVAL;
	cudaSetDevice(ctx->device);
	if (!ctx->map_nchain.empty()) {
		ctx->map_chains.clear(); ctx->map_nchain.clear();
		return fail(ctx, WB200_EUNSUPPORTED, "input maps are not applied by the batch entry points");
	}
	const int N = ntracks;
	cudaStream_t s = ctx->stream;
	// development: WB200_BATCH_TIMING=1 prints where the host spends a call (ms since entry: planned, samples read,
	// everything queued, results read)
	static const bool timing = getenv("WB200_BATCH_TIMING") != nullptr;
	const auto t_in = std::chrono::steady_clock::now();
	auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_in).count(); };
	double t_plan = 0, t_sample = 0, t_queued = 0;
	std::vector<Job> jobs(nshards);
	int64_t max_cursor = 1, max_nw = 1, max_slow = 0;
	bool sample = nshards > 0;
	for (int c = 0; c < nshards; c++) {
		wb200_device_shard &h = shards[c];
		int n1 = nset1, st_ = strict;
		int rc = plan_shard(ctx, op, N, n1, st_, h.h_off, h_defaults, h.d_start, h.d_end, h.d_value, h.pos_lo, h.pos_hi, false, 0, 1, jobs[c]);
		if (rc != WB200_OK) return rc;
		Job &job = jobs[c];
		job.out_start = h.d_out_start; job.out_end = h.d_out_end; job.out_value = h.d_out_value; job.out_cap = h.out_cap;
		max_cursor = std::max<int64_t>(max_cursor, (int64_t) (job.nw + 1) * job.cstride);
		max_nw = std::max<int64_t>(max_nw, job.nw);
		max_slow = std::max<int64_t>(max_slow, std::max<int64_t>(1 << 16, h.out_cap / 8));
		if (!wants_format(op, N, job.I, h_defaults, false)) sample = false;
		h.n_out = 0; h.auc = 0; h.slow_segments = 0; h.ms_cursor = h.ms_sweep = 0; h.limbs = 0; h.sweep_kernel = 0;
	}
	wb200_stats total{};
	ctx->last_S = -1;
	if (nshards == 0) { ctx->stats = total; return WB200_OK; }

	// per-shard tables (offsets | defaults | fixed-point defaults) and result slots, pooled
	const size_t tbl = ((size_t) (N + 1) * 8 + (size_t) N * 16 + 15) / 16 * 16;
	CK(ctx->b_tables.ensure(tbl * nshards)); CK(ctx->hb_tables.ensure(tbl * nshards));
	CK(ctx->b_small.ensure((size_t) 256 * nshards)); CK(ctx->hb_small.ensure((size_t) 256 * nshards));
	CK(ctx->cursor.ensure((size_t) max_cursor * sizeof(int32_t)));
	CK(ctx->state.ensure((size_t) max_nw * sizeof(unsigned long long)));
	CK(ctx->aucp.ensure((size_t) max_nw * WARPS * sizeof(double)));
	while ((int) ctx->batch_events.size() < 3 * nshards) {
		cudaEvent_t e;
		CK(cudaEventCreate(&e));
		ctx->batch_events.push_back(e);
	}
	char *ht = ctx->hb_tables.as<char>(), *dt = ctx->b_tables.as<char>();
	double *d_small = ctx->b_small.as<double>();
	CK(cudaMemsetAsync(d_small, 0, (size_t) 256 * nshards, s));

	t_plan = since();
	// 1. one value sample per shard, one synchronisation for all of them
	double maxabs = 0;
	long long lsb_key = 0;
	if (sample) {
		for (int c = 0; c < nshards; c++) {
			const int64_t I = jobs[c].I, stride = std::max<int64_t>(1, I / (1 << 16));
			CK(launch_sample_maxabs(jobs[c].value, I, stride, d_small + 32 * c, s));
			total.kernels++;
		}
		CK(cudaMemcpyAsync(ctx->hb_small.p, d_small, (size_t) 256 * nshards, cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		for (int c = 0; c < nshards; c++) {
			maxabs = std::max(maxabs, ctx->hb_small.as<double>()[32 * c]);
			lsb_key = std::max(lsb_key, ctx->hb_small.as<long long>()[32 * c + 1]);
		}
	}
	t_sample = since();
	// 2. tables of every shard in one upload
	for (int c = 0; c < nshards; c++) {
		Job &job = jobs[c];
		int64_t *hm_off = (int64_t *) (ht + tbl * c);
		double *hm_dflt = (double *) (hm_off + N + 1);
		int64_t *hm_fx = (int64_t *) (hm_dflt + N);
		memcpy(hm_off, shards[c].h_off, (N + 1) * sizeof(int64_t));
		memcpy(hm_dflt, h_defaults, N * sizeof(double));
		memset(hm_fx, 0, N * sizeof(int64_t));
		if (sample) choose_format(job, op, N, h_defaults, maxabs, lsb_key, 1, hm_fx);
		job.off = (const int64_t *) (dt + tbl * c);
		job.dflt = (const double *) (job.off + N + 1);
		job.dflt_fx = (const int64_t *) (job.dflt + N);
		job.cursor = ctx->cursor.as<int32_t>(); job.state = ctx->state.as<unsigned long long>();
		job.auc_partial = ctx->aucp.as<double>();
		double *sm = d_small + 32 * c;
		job.slow_count = (unsigned long long *) (sm + 3);
		job.auc_fix = sm + 10;
		job.total = (long long *) (sm + 11);
		job.ticket = (unsigned int *) (sm + 12);
		job.flags = (int *) (sm + 13);
		job.diag = (unsigned long long *) (sm + 16);
		if (job.fast) job.slow_cap = max_slow;
	}
	if (jobs[0].fast) {
		CK(ctx->slow_list.ensure((size_t) max_slow * sizeof(long long)));
		for (auto &job : jobs) job.slow_list = ctx->slow_list.as<long long>();
	}
	CK(cudaMemcpyAsync(dt, ht, tbl * nshards, cudaMemcpyHostToDevice, s));
	// 3. cursor kernel + sweep of every shard, back to back (the scratch tables are reused in stream order)
	for (int c = 0; c < nshards; c++) {
		Job &job = jobs[c];
		cudaEvent_t *ev = ctx->batch_events.data() + 3 * c;
		int64_t max_len = 0;
		for (int i = 0; i < N; i++) max_len = std::max<int64_t>(max_len, shards[c].h_off[i + 1] - shards[c].h_off[i]);
		CK(cudaMemsetAsync(ctx->state.p, 0, (size_t) job.nw * sizeof(unsigned long long), s));
		CK(cudaEventRecord(ev[0], s));
		CK(launch_cursor(job, max_len, s));
		CK(cudaEventRecord(ev[1], s));
		int auc_per_window = WARPS;
		CK(launch_emit(job, s, &auc_per_window));
		CK(cudaEventRecord(ev[2], s));
		CK(launch_reduce_auc(job.auc_partial, job.nw * auc_per_window, d_small + 32 * c + 2, s));
		total.kernels += 3;
		shards[c].sweep_kernel = auc_per_window == 1 ? 1 : 0;
	}
	// 4. all results, one synchronisation
	CK(cudaMemcpyAsync(ctx->hb_small.p, d_small, (size_t) 256 * nshards, cudaMemcpyDeviceToHost, s));
	t_queued = since();
	CK(cudaStreamSynchronize(s));
	if (timing) fprintf(stderr, "wb200_reduce_device_batch: planned %.3f, samples read %.3f, queued %.3f, results read %.3f ms\n", t_plan, t_sample, t_queued, since());
	for (int c = 0; c < nshards; c++) {
		wb200_device_shard &h = shards[c];
		const Job &job = jobs[c];
		const int64_t *hs = ctx->hb_small.as<int64_t>() + 32 * c;
		if (const int fl = (int) (hs[13] & 3))
			return fail(ctx, WB200_EINVAL, (fl & 1) ? "a track is not sorted, overlaps itself or holds an empty interval (start >= end)"
			                                        : "an interval lies outside [pos_lo, pos_hi]");
		h.n_out = hs[11];
		memcpy(&h.auc, hs + 2, 8);
		h.slow_segments = hs[3];
		h.limbs = job.fast == 1 ? job.fx_limbs : 0;
		cudaEvent_t *ev = ctx->batch_events.data() + 3 * c;
		cudaEventElapsedTime(&h.ms_cursor, ev[0], ev[1]);
		cudaEventElapsedTime(&h.ms_sweep, ev[1], ev[2]);
		for (int k = 0; k < 16; k++) ctx->diag[k] += (unsigned long long) hs[16 + k];   // zeros unless built with -DWB_DIAG_PHASES
	}
	for (int c = 0; c < nshards; c++) {
		wb200_device_shard &h = shards[c];
		const Job &job = jobs[c];
		if ((job.fast && h.slow_segments > 0) || h.n_out > h.out_cap) {
			// rare: values outside the fixed-point format or an undersized output -- the single-shard path
			// (fix-up kernel, wider formats, capacity error) finishes this shard
			const float msc = h.ms_cursor, mss = h.ms_sweep;
			double dv = 0;
			int rc = run_device(ctx, op, N, nset1, strict, h.d_start, h.d_end, h.d_value, h.h_off, h_defaults, h.pos_lo, h.pos_hi,
			                    h.d_out_start, h.d_out_end, h.d_out_value, h.out_cap, &h.n_out, &dv, &h.auc);
			if (rc != WB200_OK) return rc;
			h.slow_segments = ctx->stats.slow_segments; h.limbs = ctx->stats.limbs;
			h.ms_cursor = msc + ctx->stats.ms_count; h.ms_sweep = mss + ctx->stats.ms_emit;
			total.kernels += ctx->stats.kernels;
		}
		total.intervals += job.I; total.segments += h.n_out; total.windows += job.nw;
		total.slow_segments += h.slow_segments; total.limbs = std::max(total.limbs, h.limbs);
		total.ms_count += h.ms_cursor; total.ms_emit += h.ms_sweep;
		total.fast_path = job.fast; total.sweep_kernel = h.sweep_kernel;
	}
	total.ms_total = total.ms_count + total.ms_emit;
	ctx->stats = total;
	const bool setop = (op == OP_TTEST || op == OP_MWU);
	if (default_out) *default_out = setop ? NAN : derived_default(op, N, h_defaults);
	return WB200_OK;
}

namespace {
// The chains set by wb200_set_input_maps, applied to the staged tracks in ctx->in_*: values are
// rewritten in place; if a log dropped records the survivors are compacted into ctx->in2_* and
// `off` is rebuilt.  On return *s / *e / *v point at the arrays to reduce.
int apply_input_maps(wb200_ctx *ctx, int N, std::vector<int64_t> &off, const int32_t **s_io, const int32_t **e_io, const double **v_io) {
	if ((int) ctx->map_nchain.size() != N) return fail(ctx, WB200_EINVAL, "wb200_set_input_maps was called for a different number of tracks");
	const int64_t I = off[N];
	if (I == 0) return WB200_OK;
	cudaStream_t st = ctx->stream;
	int64_t max_len = 0;
	for (int i = 0; i < N; i++) max_len = std::max(max_len, off[i + 1] - off[i]);
	CK(ctx->m_chains.ensure((size_t) N * MAX_MAP_CHAIN * sizeof(MapOp))); CK(ctx->m_nchain.ensure((size_t) N * sizeof(int)));
	CK(ctx->m_off.ensure((size_t) (N + 1) * sizeof(int64_t))); CK(ctx->m_newoff.ensure((size_t) (N + 1) * sizeof(int64_t)));
	CK(ctx->m_keep.ensure((size_t) I)); CK(ctx->m_small.ensure(64)); CK(ctx->hm_small.ensure(64 + (size_t) (N + 1) * sizeof(int64_t)));
	CK(cudaMemcpyAsync(ctx->m_chains.p, ctx->map_chains.data(), (size_t) N * MAX_MAP_CHAIN * sizeof(MapOp), cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->m_nchain.p, ctx->map_nchain.data(), (size_t) N * sizeof(int), cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->m_off.p, off.data(), (size_t) (N + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
	CK(cudaMemsetAsync(ctx->m_small.p, 0, 64, st));
	int *d_any = ctx->m_small.as<int>();
	long long *d_R = (long long *) (ctx->m_small.as<char>() + 16);
	CK(launch_map(const_cast<double *>(*v_io), ctx->m_off.as<int64_t>(), ctx->m_chains.as<MapOp>(), ctx->m_nchain.as<int>(), N, max_len,
	              ctx->m_keep.as<unsigned char>(), d_any, st));
	CK(cudaMemcpyAsync(ctx->hm_small.p, d_any, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));      // pageable sources of the chains / offsets are free again, and: anything dropped?
	if (ctx->hm_small.as<int>()[0] == 0) return WB200_OK;
	// compaction of the surviving records
	const int ntiles = sink_tiles(I);
	CK(ctx->m_hp.ensure((size_t) (I + 1) * sizeof(int32_t)));
	CK(ctx->m_tile_count.ensure((size_t) ntiles * sizeof(int))); CK(ctx->m_tile_base.ensure((size_t) ntiles * sizeof(long long)));
	CK(sink_compact(ctx->m_keep.as<unsigned char>(), I, ctx->m_tile_count.as<int>(), ctx->m_tile_base.as<long long>(), d_R, ctx->m_hp.as<int32_t>(), st));
	CK(cudaMemcpyAsync(ctx->hm_small.as<char>() + 8, d_R, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	const long long R = *(const long long *) (ctx->hm_small.as<char>() + 8);
	const size_t m = (size_t) std::max<long long>(1, R);
	CK(ctx->in2_start.ensure(m * sizeof(int32_t))); CK(ctx->in2_end.ensure(m * sizeof(int32_t))); CK(ctx->in2_value.ensure(m * sizeof(double)));
	CK(launch_map_gather(ctx->m_hp.as<int32_t>(), d_R, R, *s_io, *e_io, *v_io, ctx->in2_start.as<int32_t>(), ctx->in2_end.as<int32_t>(),
	                     ctx->in2_value.as<double>(), ctx->m_off.as<int64_t>(), N, ctx->m_newoff.as<int64_t>(), st));
	int64_t *h_new = (int64_t *) (ctx->hm_small.as<char>() + 64);
	CK(cudaMemcpyAsync(h_new, ctx->m_newoff.p, (size_t) (N + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	for (int i = 0; i <= N; i++) off[i] = h_new[i];
	*s_io = ctx->in2_start.as<int32_t>(); *e_io = ctx->in2_end.as<int32_t>(); *v_io = ctx->in2_value.as<double>();
	return WB200_OK;
}
} // namespace

namespace {
// host tracks of one chromosome -> HBM -> kernels; the segments stay in ctx->out_*
int reduce_host_to_device(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict, const wb200_track *tracks,
                          int64_t *S_out, double *dv_out, double *auc_out) {
	cudaSetDevice(ctx->device);
	const int N = ntracks;
	if (N < 1) return fail(ctx, WB200_EINVAL, "need at least one track");
	std::vector<int64_t> off(N + 1, 0);
	std::vector<double> dflt(N);
	int64_t lo = INT32_MAX, hi = 1;
	for (int i = 0; i < N; i++) {
		off[i + 1] = off[i] + tracks[i].n;
		dflt[i] = tracks[i].default_value;
		if (tracks[i].n > 0) {
			lo = std::min<int64_t>(lo, tracks[i].start[0]);
			hi = std::max<int64_t>(hi, tracks[i].end[tracks[i].n - 1]);
		}
	}
	const int64_t I = off[N];
	if (lo > hi) lo = hi;
	cudaStream_t s = ctx->stream;
	CK(ctx->in_start.ensure(std::max<int64_t>(1, I) * sizeof(int32_t)));
	CK(ctx->in_end.ensure(std::max<int64_t>(1, I) * sizeof(int32_t)));
	CK(ctx->in_value.ensure(std::max<int64_t>(1, I) * sizeof(double)));
	for (int i = 0; i < N; i++) {
		if (!tracks[i].n) continue;
		CK(cudaMemcpyAsync(ctx->in_start.as<int32_t>() + off[i], tracks[i].start, tracks[i].n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(ctx->in_end.as<int32_t>() + off[i], tracks[i].end, tracks[i].n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
		if (tracks[i].value)
			CK(cudaMemcpyAsync(ctx->in_value.as<double>() + off[i], tracks[i].value, tracks[i].n * sizeof(double), cudaMemcpyHostToDevice, s));
		else if (tracks[i].value32) {
			// float32 values travel as they are (4 bytes each) into a scratch array and are widened on the device
			CK(ctx->bin_value32[0].ensure(std::max<int64_t>(1, I) * sizeof(float)));
			CK(cudaMemcpyAsync(ctx->bin_value32[0].as<float>() + off[i], tracks[i].value32, tracks[i].n * sizeof(float), cudaMemcpyHostToDevice, s));
			CK(launch_widen(ctx->bin_value32[0].as<float>() + off[i], ctx->in_value.as<double>() + off[i], tracks[i].n, s));
		} else
			return fail(ctx, WB200_EINVAL, "a track has neither value nor value32");
	}
	const int64_t cap = std::max<int64_t>(1, std::min<int64_t>(2 * I + 1, hi - lo + 1));
	CK(ctx->out_start.ensure(cap * sizeof(int32_t)));
	CK(ctx->out_end.ensure(cap * sizeof(int32_t)));
	CK(ctx->out_value.ensure(cap * sizeof(double)));
	const int32_t *in_s = ctx->in_start.as<int32_t>(), *in_e = ctx->in_end.as<int32_t>();
	const double *in_v = ctx->in_value.as<double>();
	if (!ctx->map_nchain.empty()) {
		int rc = apply_input_maps(ctx, N, off, &in_s, &in_e, &in_v);
		ctx->map_chains.clear(); ctx->map_nchain.clear();
		if (rc != WB200_OK) return rc;
	}
	int rc = run_device(ctx, op, N, nset1, strict, in_s, in_e,
	                    in_v, off.data(), dflt.data(), (int32_t) lo, (int32_t) hi,
	                    ctx->out_start.as<int32_t>(), ctx->out_end.as<int32_t>(), ctx->out_value.as<double>(),
	                    cap, S_out, dv_out, auc_out);
	if (rc == WB200_OK) ctx->stats.h2d_bytes = I * 16;
	return rc;
}
} // namespace

// -- regional integrators ---------------------------------------------------
// Replaces, per chromosome: ApplyMultiplexerPop / computeApplyValues (src/apply.c:259-321) with the
// integrator chain AUCIntegrator ... SpanIntegrator (src/statistics.c:62-326) over a buffered region.
int wb200_apply_device(wb200_ctx *ctx, const int32_t *d_start, const int32_t *d_end, const double *d_value, int64_t nseg,
                       double default_value, int64_t nreg, const int32_t *rstart, const int32_t *rend,
                       int nstats, const int *codes, int strict, double *out) {
	if (!ctx || nseg < 0 || nreg < 0 || nstats < 1 || nstats > 64 || !codes || (nreg > 0 && (!rstart || !rend || !out))) return WB200_EINVAL;
	for (int k = 0; k < nstats; k++) if (codes[k] < WB200_STAT_AUC || codes[k] > WB200_STAT_SPAN) return fail(ctx, WB200_EINVAL, "unknown regional statistic");
	if (nreg == 0) return WB200_OK;
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	const size_t n = (size_t) nreg;
	CK(ctx->ap_rs.ensure(n * sizeof(int32_t))); CK(ctx->ap_re.ensure(n * sizeof(int32_t)));
	CK(ctx->ap_codes.ensure(64 * sizeof(int))); CK(ctx->ap_out.ensure(n * nstats * sizeof(double)));
	CK(ctx->hap_out.ensure(n * nstats * sizeof(double)));
	CK(cudaMemcpyAsync(ctx->ap_rs.p, rstart, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
	CK(cudaMemcpyAsync(ctx->ap_re.p, rend, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
	CK(cudaMemcpyAsync(ctx->ap_codes.p, codes, nstats * sizeof(int), cudaMemcpyHostToDevice, s));
	// the buffer's default goes through float (createBufferedWiggleIteratorData keeps it in a float-typed field, apply.c:58-75)
	const double dflt = (double) (float) default_value;
	bool var_family = false;
	for (int k = 0; k < nstats; k++) var_family = var_family || (codes[k] >= WB200_STAT_VAR && codes[k] <= WB200_STAT_CV);
	CK(launch_apply(d_start, d_end, d_value, nseg, dflt, strict ? 1 : 0, ctx->ap_rs.as<int32_t>(), ctx->ap_re.as<int32_t>(), nreg,
	                nstats, ctx->ap_codes.as<int>(), ctx->ap_out.as<double>(), var_family, s));
	CK(cudaMemcpyAsync(ctx->hap_out.p, ctx->ap_out.p, n * nstats * sizeof(double), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));      // also: the (possibly pageable) region arrays are free again
	memcpy(out, ctx->hap_out.p, n * nstats * sizeof(double));
	ctx->stats = wb200_stats{};
	ctx->stats.segments = nseg; ctx->stats.kernels = 1;
	ctx->stats.h2d_bytes = (int64_t) n * 8; ctx->stats.d2h_bytes = (int64_t) n * nstats * 8;
	return WB200_OK;
}

int wb200_apply_host(wb200_ctx *ctx, const int32_t *start, const int32_t *end, const double *value, int64_t nseg,
                     double default_value, int64_t nreg, const int32_t *rstart, const int32_t *rend,
                     int nstats, const int *codes, int strict, double *out) {
	if (!ctx || nseg < 0 || (nseg > 0 && (!start || !end || !value))) return WB200_EINVAL;
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	const size_t m = (size_t) std::max<int64_t>(1, nseg);
	CK(ctx->w_in_start.ensure(m * sizeof(int32_t))); CK(ctx->w_in_end.ensure(m * sizeof(int32_t))); CK(ctx->w_in_value.ensure(m * sizeof(double)));
	if (nseg > 0) {
		CK(cudaMemcpyAsync(ctx->w_in_start.p, start, (size_t) nseg * sizeof(int32_t), cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(ctx->w_in_end.p, end, (size_t) nseg * sizeof(int32_t), cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(ctx->w_in_value.p, value, (size_t) nseg * sizeof(double), cudaMemcpyHostToDevice, s));
	}
	int rc = wb200_apply_device(ctx, ctx->w_in_start.as<int32_t>(), ctx->w_in_end.as<int32_t>(), ctx->w_in_value.as<double>(), nseg,
	                            default_value, nreg, rstart, rend, nstats, codes, strict, out);
	if (rc == WB200_OK) ctx->stats.h2d_bytes += nseg * 16;
	return rc;
}

// the segments of the last wb200_reduce_host() call on the context are the dataset: they never left HBM
int wb200_apply_last(wb200_ctx *ctx, double default_value, int64_t nreg, const int32_t *rstart, const int32_t *rend,
                     int nstats, const int *codes, int strict, double *out) {
	if (!ctx) return WB200_EINVAL;
	if (ctx->last_S < 0) return fail(ctx, WB200_EINVAL, "wb200_apply_last: no single-chromosome reduction is resident on this context");
	const Job job = ctx->last_job;
	const int64_t S = ctx->last_S;
	int rc = wb200_apply_device(ctx, job.out_start, job.out_end, job.out_value, S, default_value, nreg, rstart, rend, nstats, codes, strict, out);
	return rc;
}

// Raw multiplexer rows of the chromosome the last wb200_reduce_host() call on the context reduced
// (its staged inputs, cursor table and segments are still resident in HBM): for segments
// [first, first + count), values[(k - first) * ntracks + i] = track i's value where it is in play, else
// its default, and inplay[...] = 1 / 0 -- Multiplexer.values[] / inplay[] of the reference's sweep
// (src/multiplexer.c:76-96), which mWigWriter.c:179-225 and SelectReductionPop (reducers.c:40-65) read.
int wb200_rows_host(wb200_ctx *ctx, int64_t first, int64_t count, double *values, unsigned char *inplay) {
	if (!ctx || !values || !inplay || first < 0 || count < 0) return WB200_EINVAL;
	if (ctx->last_S < 0) return fail(ctx, WB200_EINVAL, "wb200_rows_host: no single-chromosome reduction is resident on this context");
	if (first + count > ctx->last_S) return fail(ctx, WB200_EINVAL, "wb200_rows_host: rows beyond the last reduction's segments");
	if (count == 0) return WB200_OK;
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	const size_t n = (size_t) count * ctx->last_job.N;
	CK(ctx->rows_v.ensure(n * sizeof(double)));
	CK(ctx->rows_p.ensure(n));
	CK(launch_rows(ctx->last_job, first, count, ctx->rows_v.as<double>(), ctx->rows_p.as<unsigned char>(), s));
	CK(cudaMemcpyAsync(values, ctx->rows_v.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(inplay, ctx->rows_p.p, n, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	return WB200_OK;
}

int wb200_reduce_host(wb200_ctx *ctx, int op, int ntracks, int nset1, int strict,
                      const wb200_track *tracks, wb200_segments *out) {
	if (!ctx || !tracks || !out) return WB200_EINVAL;
	int64_t S = 0;
	double dv = NAN, auc = 0;
	int rc = reduce_host_to_device(ctx, op, ntracks, nset1, strict, tracks, &S, &dv, &auc);
	if (rc != WB200_OK) return rc;
	cudaStream_t s = ctx->stream;
	CK(ctx->h_start.ensure(std::max<int64_t>(1, S) * sizeof(int32_t)));
	CK(ctx->h_end.ensure(std::max<int64_t>(1, S) * sizeof(int32_t)));
	CK(ctx->h_value.ensure(std::max<int64_t>(1, S) * sizeof(double)));
	if (S > 0) {
		CK(cudaMemcpyAsync(ctx->h_start.p, ctx->out_start.p, S * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
		CK(cudaMemcpyAsync(ctx->h_end.p, ctx->out_end.p, S * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
		CK(cudaMemcpyAsync(ctx->h_value.p, ctx->out_value.p, S * sizeof(double), cudaMemcpyDeviceToHost, s));
	}
	CK(cudaStreamSynchronize(s));
	ctx->stats.d2h_bytes = S * 16;
	out->start = ctx->h_start.as<int32_t>();
	out->end = ctx->h_end.as<int32_t>();
	out->value = ctx->h_value.as<double>();
	out->n = S;
	out->default_value = dv;
	out->auc = auc;
	out->n_ends = 0; out->end_index = nullptr; out->end_value = nullptr;
	return WB200_OK;
}



int wb200_reduce_host_batch(wb200_ctx *ctx, int op, int nshards, int ntracks, int nset1, int strict,
                            const wb200_track *tracks, wb200_segments *out) {
	if (!ctx || !tracks || !out || nshards < 0) return WB200_EINVAL;
	cudaSetDevice(ctx->device);
	if (!ctx->map_nchain.empty()) {       // chains armed by wb200_set_input_maps: never left pending for an unrelated call
		ctx->map_chains.clear(); ctx->map_nchain.clear();
		return fail(ctx, WB200_EUNSUPPORTED, "input maps are applied by wb200_reduce_host / wb200_reduce_format_host only");
	}
	const int N = ntracks;
	if (N < 1) return fail(ctx, WB200_EINVAL, "need at least one track");
	if (!ctx->s_h2d) CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
	if (!ctx->s_d2h) CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
	while ((int) ctx->pipe_events.size() < 3 * nshards) {
		cudaEvent_t e;
		CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		ctx->pipe_events.push_back(e);
	}
	cudaEvent_t *ev_h2d = ctx->pipe_events.data(), *ev_comp = ev_h2d + nshards, *ev_d2h = ev_comp + nshards;
	cudaStream_t s = ctx->stream;

	struct Shard { std::vector<int64_t> off; std::vector<double> dflt; int64_t I, lo, hi, cap, S, arena_off, h2d = 0, n_ends = -1; double dv, auc; };
	std::vector<Shard> sh(nshards);
	int64_t est = 0;
	for (int c = 0; c < nshards; c++) {
		Shard &h = sh[c];
		h.off.assign(N + 1, 0); h.dflt.resize(N);
		h.lo = INT32_MAX; h.hi = 1;
		for (int i = 0; i < N; i++) {
			const wb200_track &t = tracks[(size_t) c * N + i];
			h.off[i + 1] = h.off[i] + t.n;
			h.dflt[i] = t.default_value;
			if (t.n > 0) { h.lo = std::min<int64_t>(h.lo, t.start[0]); h.hi = std::max<int64_t>(h.hi, t.end[t.n - 1]); }
		}
		h.I = h.off[N];
		if (h.lo > h.hi) h.lo = h.hi;
		h.cap = std::max<int64_t>(1, std::min<int64_t>(2 * h.I + 1, h.hi - h.lo + 1));
		h.h2d = h.I * 16;
		est += h.I + h.I / 2 + 1024;
	}
	CK(ctx->h_arena.ensure((size_t) est * 16));     // first guess S ~ 1.5 I; grown below if needed
	bool any32 = false;
	for (int k = 0; k < nshards * N && !any32; k++) any32 = tracks[k].n > 0 && !tracks[k].value && tracks[k].value32;

	// Pipeline order: the first H2D and the last D2H are the only transfers nothing overlaps, so
	// the small shards go to the two ends and the large ones to the middle.  Results are
	// returned per shard, in the caller's order.
	std::vector<int> order(nshards);
	{
		std::vector<int> by_size(nshards);
		for (int c = 0; c < nshards; c++) by_size[c] = c;
		std::stable_sort(by_size.begin(), by_size.end(), [&](int x, int y) { return sh[x].I < sh[y].I; });
		int lo = 0, hi = nshards - 1;
		for (int r = 0; r < nshards; r++) { if (r & 1) order[hi--] = by_size[r]; else order[lo++] = by_size[r]; }
	}
	auto stage = [&](int k) -> int {
		const int c = order[k];
		Shard &h = sh[c];
		const int b = k & 1;
		CK(ctx->bin_start[b].ensure(std::max<int64_t>(1, h.I) * sizeof(int32_t)));
		CK(ctx->bin_end[b].ensure(std::max<int64_t>(1, h.I) * sizeof(int32_t)));
		CK(ctx->bin_value[b].ensure(std::max<int64_t>(1, h.I) * sizeof(double)));
		if (k >= 2) CK(cudaStreamWaitEvent(ctx->s_h2d, ev_comp[k - 2], 0));   // input set b is free again
		// Tracks that lie back to back in host memory (the layout the BigWig batch decoder and any
		// SoA producer hand over) travel as one copy per array: measured on this box, 1 MiB
		// transfers reach 46 GB/s and 63 GB/s with both directions busy, 64 MiB ones 55 and 97
		// (profiles/pcie_probe.py).
		for (int i = 0; i < N; ) {
			const wb200_track &t = tracks[(size_t) c * N + i];
			if (!t.n) { i++; continue; }
			int64_t n = t.n;
			int j = i + 1;
			for (; j < N; j++) {
				const wb200_track &u = tracks[(size_t) c * N + j];
				if (!u.n) continue;
				if (u.start != t.start + n || u.end != t.end + n) break;
				if (t.value ? u.value != t.value + n : (u.value != nullptr || u.value32 != t.value32 + n)) break;
				n += u.n;
			}
			CK(cudaMemcpyAsync(ctx->bin_start[b].as<int32_t>() + h.off[i], t.start, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_h2d));
			CK(cudaMemcpyAsync(ctx->bin_end[b].as<int32_t>() + h.off[i], t.end, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_h2d));
			if (t.value)
				CK(cudaMemcpyAsync(ctx->bin_value[b].as<double>() + h.off[i], t.value, n * sizeof(double), cudaMemcpyHostToDevice, ctx->s_h2d));
			else if (t.value32) {
				// narrow transport: 4 bytes per value over PCIe, widened to fp64 on the device (on the copy stream)
				CK(cudaMemcpyAsync(ctx->bin_value32[b].as<float>() + h.off[i], t.value32, n * sizeof(float), cudaMemcpyHostToDevice, ctx->s_h2d));
				CK(launch_widen(ctx->bin_value32[b].as<float>() + h.off[i], ctx->bin_value[b].as<double>() + h.off[i], n, ctx->s_h2d));
				h.h2d -= n * 4;
			} else
				return fail(ctx, WB200_EINVAL, "a track has neither value nor value32");
			i = j;
		}
		CK(cudaEventRecord(ev_h2d[k], ctx->s_h2d));
		return WB200_OK;
	};

	// DevBuf::ensure may reallocate: size both buffer sets for the largest shard up front
	int64_t maxI = 1, maxcap = 1;
	for (auto &h : sh) { maxI = std::max(maxI, h.I); maxcap = std::max(maxcap, h.cap); }
	for (int b = 0; b < 2; b++) {
		CK(ctx->bin_start[b].ensure(maxI * sizeof(int32_t))); CK(ctx->bin_end[b].ensure(maxI * sizeof(int32_t)));
		CK(ctx->bin_value[b].ensure(maxI * sizeof(double)));
		CK(ctx->bout_start[b].ensure(maxcap * sizeof(int32_t))); CK(ctx->bout_end[b].ensure(maxcap * sizeof(int32_t)));
		CK(ctx->bout_value[b].ensure(maxcap * sizeof(double)));
		if (any32) CK(ctx->bin_value32[b].ensure(maxI * sizeof(float)));
		if (ctx->breakpoint_out) {
			CK(ctx->ends_val[b].ensure((size_t) (maxcap / 4 + 1024) * sizeof(int32_t)));
			CK(ctx->ends_idx[b].ensure((size_t) (maxcap / 4 + 1024) * sizeof(int32_t)));
		}
	}
	if (ctx->breakpoint_out) {
		// scratch of the explicit-end compaction (the text sink's count / scan / scatter buffers)
		const int ntiles = sink_tiles(maxcap);
		CK(ctx->w_heads.ensure((size_t) maxcap)); CK(ctx->w_hp.ensure((size_t) (maxcap + 1) * sizeof(int32_t)));
		CK(ctx->w_tile_count.ensure((size_t) ntiles * sizeof(int))); CK(ctx->w_tile_base.ensure((size_t) ntiles * sizeof(long long)));
		CK(ctx->w_small.ensure(64)); CK(ctx->hw_small.ensure(64));
	}

	int64_t arena_used = 0, h2d = 0, d2h = 0;
	wb200_stats total{};
	if (nshards > 0) { int rc = stage(0); if (rc) return rc; }
	for (int k = 0; k < nshards; k++) {
		const int c = order[k];
		Shard &h = sh[c];
		const int b = k & 1;
		if (k + 1 < nshards) { int rc = stage(k + 1); if (rc) return rc; }
		CK(cudaStreamWaitEvent(s, ev_h2d[k], 0));
		if (k >= 2) CK(cudaStreamWaitEvent(s, ev_d2h[k - 2], 0));              // output set b has been drained
		int rc = run_device(ctx, op, N, nset1, strict, ctx->bin_start[b].as<int32_t>(), ctx->bin_end[b].as<int32_t>(),
		                    ctx->bin_value[b].as<double>(), h.off.data(), h.dflt.data(), (int32_t) h.lo, (int32_t) h.hi,
		                    ctx->bout_start[b].as<int32_t>(), ctx->bout_end[b].as<int32_t>(), ctx->bout_value[b].as<double>(),
		                    h.cap, &h.S, &h.dv, &h.auc);
		if (rc != WB200_OK) return rc;
		CK(cudaEventRecord(ev_comp[k], s));
		// accumulate stats
		total.intervals += ctx->stats.intervals; total.segments += ctx->stats.segments; total.windows += ctx->stats.windows;
		total.slow_segments += ctx->stats.slow_segments; total.kernels += ctx->stats.kernels; total.fast_path = ctx->stats.fast_path;
		total.limbs = std::max(total.limbs, ctx->stats.limbs); total.sweep_kernel = ctx->stats.sweep_kernel;
		total.ms_count += ctx->stats.ms_count; total.ms_emit += ctx->stats.ms_emit; total.ms_total += ctx->stats.ms_total;
		// breakpoint form: which segments need their end sent (the last one, and those followed by a gap)
		bool narrow = false;
		if (ctx->breakpoint_out && h.S > 0) {
			long long *d_R = (long long *) (ctx->w_small.as<char>() + 16);
			CK(launch_explicit_ends(ctx->bout_start[b].as<int32_t>(), ctx->bout_end[b].as<int32_t>(), h.S, ctx->w_heads.as<unsigned char>(), s));
			CK(sink_compact(ctx->w_heads.as<unsigned char>(), h.S, ctx->w_tile_count.as<int>(), ctx->w_tile_base.as<long long>(), d_R,
			                ctx->w_hp.as<int32_t>(), s));
			CK(cudaMemcpyAsync(ctx->hw_small.p, d_R, 8, cudaMemcpyDeviceToHost, s));
			CK(cudaStreamSynchronize(s));
			h.n_ends = ctx->hw_small.as<long long>()[0];
			narrow = h.n_ends <= h.S / 4 + 1;
			if (narrow) {
				CK(launch_gather_ends(ctx->w_hp.as<int32_t>(), h.n_ends, ctx->bout_end[b].as<int32_t>(), ctx->ends_val[b].as<int32_t>(), s));
				// the positions leave the shared compaction scratch for this buffer set's own list
				CK(cudaMemcpyAsync(ctx->ends_idx[b].p, ctx->w_hp.p, (size_t) h.n_ends * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
			} else h.n_ends = -1;
			CK(cudaEventRecord(ev_comp[k], s));
			total.kernels += narrow ? 5 : 4;
		}
		// host space for this shard's segments
		const int64_t need = (h.S + 2) * 16 + (narrow ? (h.n_ends + 2) * 8 : 0);
		if ((size_t) (arena_used + need) > ctx->h_arena.cap) {
			CK(cudaStreamSynchronize(ctx->s_d2h));
			HostBuf bigger;
			CK(bigger.ensure((size_t) ((arena_used + need) * 2)));
			memcpy(bigger.p, ctx->h_arena.p, (size_t) arena_used);
			ctx->h_arena.release();
			ctx->h_arena = bigger;
		}
		h.arena_off = arena_used;
		arena_used += need;
		char *base = (char *) ctx->h_arena.p + h.arena_off;
		CK(cudaStreamWaitEvent(ctx->s_d2h, ev_comp[k], 0));
		if (h.S > 0) {
			CK(cudaMemcpyAsync(base, ctx->bout_value[b].p, h.S * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_d2h));
			CK(cudaMemcpyAsync(base + (h.S + 1) * 8, ctx->bout_start[b].p, h.S * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
			if (!narrow)
				CK(cudaMemcpyAsync(base + (h.S + 1) * 12, ctx->bout_end[b].p, h.S * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
			else {
				CK(cudaMemcpyAsync(base + (h.S + 2) * 16, ctx->ends_idx[b].p, h.n_ends * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
				CK(cudaMemcpyAsync(base + (h.S + 2) * 16 + (h.n_ends + 1) * 4, ctx->ends_val[b].p, h.n_ends * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
			}
		}
		CK(cudaEventRecord(ev_d2h[k], ctx->s_d2h));
		h2d += h.h2d; d2h += narrow ? h.S * 12 + h.n_ends * 8 : h.S * 16;
	}
	CK(cudaStreamSynchronize(ctx->s_d2h));
	CK(cudaStreamSynchronize(ctx->s_h2d));
	for (int c = 0; c < nshards; c++) {
		Shard &h = sh[c];
		char *base = (char *) ctx->h_arena.p + h.arena_off;
		out[c].value = (double *) base;
		out[c].start = (int32_t *) (base + (h.S + 1) * 8);
		const bool narrow = h.n_ends >= 0;
		out[c].end = narrow ? nullptr : (int32_t *) (base + (h.S + 1) * 12);
		out[c].n_ends = narrow ? h.n_ends : 0;
		out[c].end_index = narrow ? (const int32_t *) (base + (h.S + 2) * 16) : nullptr;
		out[c].end_value = narrow ? (const int32_t *) (base + (h.S + 2) * 16 + (h.n_ends + 1) * 4) : nullptr;
		out[c].n = h.S;
		out[c].default_value = h.dv;
		out[c].auc = h.auc;
	}
	ctx->stats = total;
	ctx->stats.h2d_bytes = h2d;
	ctx->stats.d2h_bytes = d2h;
	return WB200_OK;
}

// moments of tracks px and py over the union cover of all ntracks tracks
static int pearson_shard(wb200_ctx *ctx, int ntracks, const wb200_track *tracks, int px, int py, double moments[6]) {
	if (!ctx->map_nchain.empty()) {
		ctx->map_chains.clear(); ctx->map_nchain.clear();
		return fail(ctx, WB200_EUNSUPPORTED, "input maps are applied by wb200_reduce_host / wb200_reduce_format_host only");
	}
	const int N = ntracks;
	std::vector<int64_t> off(N + 1, 0);
	std::vector<double> dflt(N);
	int64_t lo = INT32_MAX, hi = 1;
	for (int i = 0; i < N; i++) {
		off[i + 1] = off[i] + tracks[i].n;
		dflt[i] = tracks[i].default_value;
		if (tracks[i].n > 0) {
			lo = std::min<int64_t>(lo, tracks[i].start[0]);
			hi = std::max<int64_t>(hi, tracks[i].end[tracks[i].n - 1]);
		}
	}
	if (lo > hi) lo = hi;
	const int64_t I = off[N];
	cudaStream_t s = ctx->stream;
	CK(ctx->in_start.ensure(std::max<int64_t>(1, I) * sizeof(int32_t)));
	CK(ctx->in_end.ensure(std::max<int64_t>(1, I) * sizeof(int32_t)));
	CK(ctx->in_value.ensure(std::max<int64_t>(1, I) * sizeof(double)));
	for (int i = 0; i < N; i++) {
		if (!tracks[i].n) continue;
		CK(cudaMemcpyAsync(ctx->in_start.as<int32_t>() + off[i], tracks[i].start, tracks[i].n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(ctx->in_end.as<int32_t>() + off[i], tracks[i].end, tracks[i].n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
		if (tracks[i].value)
			CK(cudaMemcpyAsync(ctx->in_value.as<double>() + off[i], tracks[i].value, tracks[i].n * sizeof(double), cudaMemcpyHostToDevice, s));
		else if (tracks[i].value32) {
			CK(ctx->bin_value32[0].ensure(std::max<int64_t>(1, I) * sizeof(float)));
			CK(cudaMemcpyAsync(ctx->bin_value32[0].as<float>() + off[i], tracks[i].value32, tracks[i].n * sizeof(float), cudaMemcpyHostToDevice, s));
			CK(launch_widen(ctx->bin_value32[0].as<float>() + off[i], ctx->in_value.as<double>() + off[i], tracks[i].n, s));
		} else
			return fail(ctx, WB200_EINVAL, "a track has neither value nor value32");
	}
	int rc = run_device(ctx, OP_PEARSON, N, N, 0, ctx->in_start.as<int32_t>(), ctx->in_end.as<int32_t>(),
	                    ctx->in_value.as<double>(), off.data(), dflt.data(), (int32_t) lo, (int32_t) hi,
	                    nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, moments, false, px, py);
	ctx->stats.h2d_bytes = I * 16;
	ctx->stats.d2h_bytes = 48;
	return rc;
}

int wb200_pearson_host(wb200_ctx *ctx, const wb200_track *x, const wb200_track *y, double moments[6]) {
	if (!ctx || !x || !y || !moments) return WB200_EINVAL;
	cudaSetDevice(ctx->device);
	const wb200_track two[2] = {*x, *y};
	return pearson_shard(ctx, 2, two, 0, 1, moments);
}

int wb200_pearson_pair_host(wb200_ctx *ctx, int ntracks, const wb200_track *tracks, int px, int py, double moments[6]) {
	if (!ctx || !tracks || !moments || ntracks < 2 || px < 0 || py < 0 || px >= ntracks || py >= ntracks) return WB200_EINVAL;
	cudaSetDevice(ctx->device);
	return pearson_shard(ctx, ntracks, tracks, px, py, moments);
}

// Shard merge for the Pearson moments {n, sum_X, sum_Y, T_XX, T_XY, T_YY}: the
// pairwise form of the reference's sequential update (src/statistics.c:443-456,
// derivation in technical-supplement.tex:14-74) applied to two pre-reduced
// shards: T(a+b) = T(a) + T(b) + n_a n_b/(n_a+n_b) * dX * dY.
void wb200_pearson_merge(double a[6], const double b[6]) { merge6(a, b); }

// PearsonPop's final expression (src/statistics.c:420-421): NaN unless T_XX*T_YY != 0
double wb200_pearson_result(const double m[6]) {
	if (m[3] * m[5] != 0) return m[4] / std::sqrt(m[3] * m[5]);
	return NAN;
}

} // extern "C"
