// wiggletools_b200 -- small device helpers shared by the window kernels
// (engine.cu, fastsum.cu).
#pragma once
#include <cstdint>
#include <cmath>

namespace wb {

#define FULL 0xffffffffu

// Diagnostic builds (profiles/): per-phase cycle counters of the window pipeline,
// accumulated by thread 0 of every CTA into job.diag[phase].
#ifdef WB_DIAG_PHASES
#define WB_PHASE(k) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&job.diag[k], (unsigned long long) (t_ - diag_t)); diag_t = t_; } } while (0)
#define WB_DIAG_PARAM , long long &diag_t          // for device functions that mark phases of their own
#define WB_DIAG_ARG , diag_t
#else
#define WB_PHASE(k) do { } while (0)
#define WB_DIAG_PARAM
#define WB_DIAG_ARG
#endif

// value -> fixed point: m = v * 2^-q exactly, |m| < lim; false for NaN, Inf, values out of
// range or with bits below 2^q (those intervals take the row path)
__device__ __forceinline__ bool to_fixed(double v, double inv_scale, double lim, long long &m) {
	double t = v * inv_scale;
	if (!(fabs(t) < lim)) return false;          // NaN, Inf, too large
	if (t != rint(t)) return false;              // bits below 2^q
	if (t == 0.0 && v != 0.0) return false;      // underflow
	m = (long long) t;
	return true;
}

// look-back words of the single-pass chained scan over windows: 2 flag bits | segment count
constexpr unsigned long long LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_VALUE = (1ull << 62) - 1;
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long *p) { return *(const volatile unsigned long long *) p; }
__device__ __forceinline__ void st_volatile(unsigned long long *p, unsigned long long v) { *(volatile unsigned long long *) p = v; }

// one fixed-point delta into one or two 32-bit limb arrays (native 32-bit shared atomics)
template <int LIMBS, int W>
__device__ __forceinline__ void add_limbs_fs(uint32_t *limb, int ph, long long d, int lw) {
	if (LIMBS == 1) { atomicAdd(&limb[ph], (uint32_t) d); return; }
	atomicAdd(&limb[ph], (uint32_t) (d & ((1ll << lw) - 1)));
	atomicAdd(&limb[W + ph], (uint32_t) (d >> lw));
}

// var / stddev / CV of one segment from exact integer moments of the in-play
// values: c tracks in play, Sx = sum m, Sxx = sum m^2 with value = m * 2^q.
// Reference arithmetic being matched (src/reducers.c:428-479, :511-563, :672-725):
//   mean = (sum over all N of float(x)) / N            -- absent tracks contribute 0 here
//   var    = sum over in-play (mean - v)^2 / N
//   stddev = sqrt(sum over all N (mean - x)^2 / N),  CV = stddev / mean
// The sum of squared deviations is evaluated without cancellation:
//   sum_C (mean - v)^2 = [ (c Sxx - Sx^2) 2^2q + (c mean - X)^2 ] / c ,  X = Sx 2^q
// where c Sxx - Sx^2 >= 0 is exact in 128-bit integers.
// WIDE: the integer type c Sxx and Sx^2 are formed in -- __int128 in general, 64 bits when the
// caller knows N 2^mbits <= 2^30 (both products are then below 2^60); the result is the same.
template <typename WIDE = __int128>
__device__ __forceinline__ double moments_stat(int op, int N, int c, long long Sx, long long Sxx, double scale) {
	if (op == OP_VAR && N < 2) return NAN;
	const double X = (double) Sx * scale;                  // exact: |Sx| < 2^53
	const double mean = X / N;
	if (op == OP_CV && mean == 0) return NAN;
	const WIDE d1 = (WIDE) c * (WIDE) Sxx - (WIDE) Sx * (WIDE) Sx;
	const double D1 = (double) d1 * scale * scale;
	const double r = fma((double) c, mean, -X);
	const double EC = (D1 + r * r) / c;
	if (op == OP_VAR) return EC / N;
	const double Eall = EC + (double) (N - c) * mean * mean;
	const double sd = sqrt(Eall / N);
	return op == OP_CV ? sd / mean : sd;
}

} // namespace wb
