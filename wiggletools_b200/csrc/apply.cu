// wiggletools_b200 -- regional integrators on the device.
//
// `apply` / `apply_paste` (/root/reference/src/apply.c:259-353) evaluate, for every region of a
// sorted BED-like list, a chain of integrators (src/statistics.c:62-326: AUC, meanI, varI,
// stddevI, CVI, maxI, minI, spanI) over the dataset restricted to the region.  The reference
// copies the region into a per-base buffer and feeds it back one base at a time
// (createBufferedWiggleIteratorData / BufferedWiggleIterator, apply.c:58-160), with the runs the
// dataset does not cover served as one record of its default value unless `strict`.
//
// Here the dataset is the segment list of one chromosome resident in HBM (typically the output
// of a reduction that never left the device) and a region is a segmented reduction over the
// segments that overlap it: one warp per region, binary search for the first overlapping
// segment, then 32 segments per step; every lane folds its clipped segments (and the default
// runs between them) into running sums, extrema and a (count, mean, M2) triple, and the lanes
// are merged pairwise (Chan et al.) -- the same quantities the reference's online updates
// (statistics.c:209-260) maintain, without its base-by-base loop.  Regions may overlap or nest:
// each is independent.  NaN records are skipped by every integrator, as in the reference.
// AUC / meanI / maxI / minI / spanI come from this kernel (bit-identical on exactly representable
// data, within a few ulps otherwise); varI / stddevI / CVI are overwritten by apply_var_kernel below.
#include "engine.cuh"
#include "device_util.cuh"
#include <cmath>

namespace wb {

struct RegAcc {
	double sum;      // sum of L * v            (AUC, numerator of meanI)
	double span;     // sum of L                (spanI, denominator of meanI)
	double mx, mn;   // extrema, NaN while empty
	double n, mean, m2;   // count of bases, their mean, sum of squared deviations (varI / stddevI / CVI)
};

__device__ __forceinline__ void acc_add(RegAcc &a, int L, double v) {
	if (isnan(v) || L <= 0) return;
	const double l = (double) L;
	a.sum += l * v;
	a.span += l;
	if (isnan(a.mx) || v > a.mx) a.mx = v;
	if (isnan(a.mn) || v < a.mn) a.mn = v;
	// L bases of value v joined to the running (n, mean, M2)
	const double n = a.n + l, d = v - a.mean;
	a.m2 += d * d * (a.n * l / n);
	a.mean += d * (l / n);
	a.n = n;
}

__device__ __forceinline__ void acc_merge(RegAcc &a, const RegAcc &b) {
	a.sum += b.sum;
	a.span += b.span;
	if (isnan(a.mx) || (!isnan(b.mx) && b.mx > a.mx)) a.mx = b.mx;
	if (isnan(a.mn) || (!isnan(b.mn) && b.mn < a.mn)) a.mn = b.mn;
	if (b.n > 0) {
		if (a.n == 0) { a.n = b.n; a.mean = b.mean; a.m2 = b.m2; }
		else {
			const double n = a.n + b.n, d = b.mean - a.mean;
			a.m2 += b.m2 + d * d * (a.n * b.n / n);
			a.mean += d * (b.n / n);
			a.n = n;
		}
	}
}

// first index in [lo, hi) with a[idx] > key (a ascending); hi if none
__device__ __forceinline__ long long upper_bound_i32(const int32_t *a, long long lo, long long hi, int key) {
	while (lo < hi) {
		const long long mid = (lo + hi) >> 1;
		if (a[mid] > key) hi = mid; else lo = mid + 1;
	}
	return lo;
}

__global__ void __launch_bounds__(128) apply_kernel(const int32_t *start, const int32_t *end, const double *value, long long nseg,
                                                    double dflt, int strict, const int32_t *rstart, const int32_t *rend, long long nreg,
                                                    int nstats, const int *codes, double *out) {
	const long long r = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (r >= nreg) return;
	const int rs = rstart[r], re = rend[r];
	RegAcc a;
	a.sum = 0; a.span = 0; a.mx = NAN; a.mn = NAN; a.n = 0; a.mean = 0; a.m2 = 0;
	if (re > rs) {
		// segments k0 .. kend-1 overlap [rs, re): end > rs and start < re (segments are sorted and disjoint)
		const long long k0 = upper_bound_i32(end, 0, nseg, rs);
		const long long kend = upper_bound_i32(start, k0, nseg, re - 1);
		// record k (k0 <= k <= kend): the default run between segment k-1 and segment k (the region's
		// edges standing in at both ends), then segment k itself clipped to the region
		for (long long k = k0 + lane; k <= kend; k += 32) {
			const int prev_end = k == k0 ? rs : max(end[k - 1], rs);
			const int cur_start = k < kend ? max(start[k], rs) : re;
			if (!strict) acc_add(a, cur_start - prev_end, dflt);
			if (k < kend) acc_add(a, min(end[k], re) - cur_start, value[k]);
		}
	}
	// lanes -> lane 0, pairwise
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		RegAcc b;
		b.sum = __shfl_down_sync(FULL, a.sum, d); b.span = __shfl_down_sync(FULL, a.span, d);
		b.mx = __shfl_down_sync(FULL, a.mx, d); b.mn = __shfl_down_sync(FULL, a.mn, d);
		b.n = __shfl_down_sync(FULL, a.n, d); b.mean = __shfl_down_sync(FULL, a.mean, d); b.m2 = __shfl_down_sync(FULL, a.m2, d);
		acc_merge(a, b);
	}
	if (lane == 0) {
		// final expressions of the integrators (statistics.c:62-326); count - 1 = 0 and count = 0 give the
		// same NaN / -0 the reference's divisions give
		const double var = a.m2 / (a.n - 1.0);
		for (int k = 0; k < nstats; k++) {
			double res = NAN;
			switch (codes[k]) {
			case 1: res = a.sum; break;                                    // AUC
			case 2: res = a.span > 0 ? a.sum / a.span : NAN; break;          // meanI
			case 3: res = var; break;                                      // varI
			case 4: res = sqrt(var); break;                                // stddevI
			case 5: res = sqrt(var) / (a.sum / a.n); break;                // CVI
			case 6: res = a.mx; break;                                     // maxI
			case 7: res = a.mn; break;                                     // minI
			case 8: res = a.span; break;                                   // spanI
			}
			out[r * nstats + k] = res;
		}
	}
}

// varI / stddevI / CVI.  The reference's online update (VarianceCorePop, statistics.c:238-255) is not
// a stable one: over a run of equal values it cancels catastrophically and leaves T at -1e-15
// instead of 0, so its stddevI is NaN where an accurate algorithm says 0 -- noise that no
// reassociated sum can reproduce.  These three statistics therefore replay the reference's own
// recurrence: one thread per region, records in position order, covered bases one at a time (the
// buffer hands them over as 1-bp records, apply.c:118-135), uncovered runs as one record of the
// default value, the same fp64 operations in the same order (the build has no FMA contraction).
// Results are bit-identical to the reference's; the cost is O(region length) per thread.
__global__ void __launch_bounds__(128) apply_var_kernel(const int32_t *start, const int32_t *end, const double *value, long long nseg,
                                                        double dflt, int strict, const int32_t *rstart, const int32_t *rend, long long nreg,
                                                        int nstats, const int *codes, double *out) {
	const long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nreg) return;
	const int rs = rstart[r], re = rend[r];
	double T = 0, vsum = 0;
	long long vcount = 0;
	auto record = [&](int L, double v) {
		if (isnan(v)) return;
		if (vcount) {
			const double old_mean = vsum / vcount;
			const double new_mean = vsum / (vcount + L);
			const double delta_T = old_mean * new_mean - new_mean * 2 * v + ((double) vcount / (vcount + L)) * v * v;
			T += delta_T * L;
		}
		vcount += L;
		vsum += L * v;
	};
	if (re > rs) {
		const long long k0 = upper_bound_i32(end, 0, nseg, rs);
		int pos = rs;                                   // next base to account for
		for (long long k = k0; k < nseg && start[k] < re; k++) {
			const int a = max(start[k], rs), b = min(end[k], re);
			if (a > pos && !strict) record(a - pos, dflt);
			const double v = value[k];
			for (int p = a; p < b; p++) record(1, v);
			pos = max(pos, b);
		}
		if (re > pos && !strict) record(re - pos, dflt);
	}
	const double var = T / (vcount - 1);
	for (int k = 0; k < nstats; k++) {
		if (codes[k] == 3) out[r * nstats + k] = var;
		else if (codes[k] == 4) out[r * nstats + k] = sqrt(var);
		else if (codes[k] == 5) out[r * nstats + k] = sqrt(var) / (vsum / vcount);
	}
}

cudaError_t launch_apply(const int32_t *start, const int32_t *end, const double *value, long long nseg, double dflt, int strict,
                         const int32_t *rstart, const int32_t *rend, long long nreg, int nstats, const int *codes, double *out,
                         bool var_family, cudaStream_t s) {
	if (nreg <= 0) return cudaSuccess;
	const long long threads = nreg * 32;
	apply_kernel<<<(unsigned) ((threads + 127) / 128), 128, 0, s>>>(start, end, value, nseg, dflt, strict, rstart, rend, nreg, nstats, codes, out);
	if (var_family)
		apply_var_kernel<<<(unsigned) ((nreg + 127) / 128), 128, 0, s>>>(start, end, value, nseg, dflt, strict, rstart, rend, nreg, nstats, codes, out);
	return cudaGetLastError();
}

} // namespace wb
