// wiggletools_b200 -- element-wise maps on the staged inputs.
//
// `mean map ln a b c`, `sum scale 2 x`, `pow 2 x` ...: the reference wraps every source in a
// unary iterator that rewrites one record at a time on the host
//   ScaleWiggleIteratorPop   /root/reference/src/unaryOps.c:650-665   NaN -> NaN, else s * v
//   ShiftWiggleIteratorPop   :722-734                                 s + v
//   LogWiggleIteratorPop     :760-779    records with v <= 0 are DROPPED, NaN stays, else log(v) / log(base)
//   ExpWiggleIteratorPop     :823-835                                 exp(v * log(radix))
//   PowerWiggleIteratorPop   :873-889    (s < 0 and v <= 0) or NaN -> NaN, else pow(v, s)
//   AbsWiggleIteratorPop     :934-949                                 NaN -> NaN, else |v|
// Here the chain of each track is applied to its values in HBM right after staging, before
// the sweep: map_kernel rewrites the values in place and flags dropped records; when
// anything was dropped the survivors are compacted (positions by the sink's count / scan /
// scatter kernels, then one gather) and the per-track offsets are rebuilt on the device.
#include "engine.cuh"
#include "device_util.cuh"

namespace wb {
namespace maps {

__device__ __forceinline__ double nan_value() { return __longlong_as_double(0x7ff8000000000000ll); }   // NAN of <math.h>

// grid: (blocks over the longest track, tracks)
__global__ void __launch_bounds__(256) map_kernel(double *value, const int64_t *off, const MapOp *chains, const int *nchain,
                                                  unsigned char *keep, int *any_dropped) {
	const int i = blockIdx.y;
	const int nc = nchain[i];
	const MapOp *ch = chains + (size_t) i * MAX_MAP_CHAIN;
	const int64_t o = off[i], n = off[i + 1] - o;
	bool dropped_any = false;
	for (int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; j < n; j += (int64_t) gridDim.x * blockDim.x) {
		double v = value[o + j];
		bool kept = true;
		for (int c = 0; c < nc && kept; c++) {
			const double a = ch[c].arg;
			switch (ch[c].op) {
			case MAP_SCALE: v = isnan(v) ? nan_value() : a * v; break;
			case MAP_OFFSET: v = a + v; break;
			case MAP_LOG:
				if (v <= 0) kept = false;                                  // NaN <= 0 is false: NaN records stay
				else v = isnan(v) ? nan_value() : log(v) / a;
				break;
			case MAP_EXP: v = exp(v * a); break;
			case MAP_POW: v = ((a < 0 && v <= 0) || isnan(v)) ? nan_value() : pow(v, a); break;
			case MAP_ABS: v = isnan(v) ? nan_value() : fabs(v); break;
			default: break;
			}
		}
		value[o + j] = v;
		keep[o + j] = kept ? 1 : 0;
		dropped_any |= !kept;
	}
	if (__any_sync(FULL, dropped_any) && (threadIdx.x & 31) == 0) *any_dropped = 1;
}

__global__ void gather_kernel(const int32_t *hp, long long R, const int32_t *s_in, const int32_t *e_in, const double *v_in,
                              int32_t *s_out, int32_t *e_out, double *v_out) {
	for (long long r = blockIdx.x * (long long) blockDim.x + threadIdx.x; r < R; r += (long long) gridDim.x * blockDim.x) {
		const int32_t k = hp[r];
		s_out[r] = s_in[k]; e_out[r] = e_in[k]; v_out[r] = v_in[k];
	}
}

// new_off[i] = number of kept records before the old start of track i
__global__ void bounds_kernel(const int32_t *hp, long long R, const int64_t *off, int N, int64_t *new_off) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > N) return;
	const int64_t target = off[i];
	long long lo = 0, hi = R;                                              // first r with hp[r] >= target
	while (lo < hi) { const long long mid = (lo + hi) >> 1; if (hp[mid] < target) lo = mid + 1; else hi = mid; }
	new_off[i] = lo;
}

// -- narrow transport over PCIe ---------------------------------------------------------------
// float32 values staged from the host (BigWig payloads are float32) widened to the fp64 the kernels read
__global__ void widen_kernel(const float *in, double *out, long long n) {
	for (long long j = blockIdx.x * (long long) blockDim.x + threadIdx.x; j < n; j += (long long) gridDim.x * blockDim.x) out[j] = (double) in[j];
}
// H[k] = 1 where segment k's end must be sent explicitly: the last segment, and every segment that is
// followed by a gap (end[k] != start[k+1]); everywhere else end[k] is start[k+1]
__global__ void explicit_end_kernel(const int32_t *start, const int32_t *end, long long S, unsigned char *H) {
	for (long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x; k < S; k += (long long) gridDim.x * blockDim.x)
		H[k] = (k + 1 == S || end[k] != start[k + 1]) ? 1 : 0;
}
__global__ void gather_end_kernel(const int32_t *hp, long long R, const int32_t *end, int32_t *out) {
	for (long long r = blockIdx.x * (long long) blockDim.x + threadIdx.x; r < R; r += (long long) gridDim.x * blockDim.x) out[r] = end[hp[r]];
}

} // namespace maps

cudaError_t launch_widen(const float *in, double *out, long long n, cudaStream_t s) {
	if (n <= 0) return cudaSuccess;
	maps::widen_kernel<<<(int) std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, s>>>(in, out, n);
	return cudaGetLastError();
}
cudaError_t launch_explicit_ends(const int32_t *start, const int32_t *end, long long S, unsigned char *H, cudaStream_t s) {
	if (S <= 0) return cudaSuccess;
	maps::explicit_end_kernel<<<(int) std::min<long long>((S + 255) / 256, 148 * 16), 256, 0, s>>>(start, end, S, H);
	return cudaGetLastError();
}
cudaError_t launch_gather_ends(const int32_t *hp, long long R, const int32_t *end, int32_t *out, cudaStream_t s) {
	if (R <= 0) return cudaSuccess;
	maps::gather_end_kernel<<<(int) std::min<long long>((R + 255) / 256, 148 * 16), 256, 0, s>>>(hp, R, end, out);
	return cudaGetLastError();
}

cudaError_t launch_map(double *value, const int64_t *d_off, const MapOp *d_chains, const int *d_nchain, int N, int64_t max_len,
                       unsigned char *keep, int *any_dropped, cudaStream_t s) {
	if (N <= 0 || max_len <= 0) return cudaSuccess;
	dim3 grid((unsigned) std::min<int64_t>((max_len + 255) / 256, 1024), (unsigned) N);
	maps::map_kernel<<<grid, 256, 0, s>>>(value, d_off, d_chains, d_nchain, keep, any_dropped);
	return cudaGetLastError();
}

cudaError_t launch_map_gather(const int32_t *hp, const long long *d_R, long long R_upper, const int32_t *s_in, const int32_t *e_in,
                              const double *v_in, int32_t *s_out, int32_t *e_out, double *v_out,
                              const int64_t *d_off, int N, int64_t *d_new_off, cudaStream_t s) {
	(void) d_R;
	if (R_upper > 0) {
		const int grid = (int) std::min<long long>((R_upper + 255) / 256, 148 * 16);
		maps::gather_kernel<<<grid, 256, 0, s>>>(hp, R_upper, s_in, e_in, v_in, s_out, e_out, v_out);
	}
	maps::bounds_kernel<<<(N + 1 + 255) / 256, 256, 0, s>>>(hp, R_upper, d_off, N, d_new_off);
	return cudaGetLastError();
}

} // namespace wb
