// wiggletools_b200 -- the text sink on the device.
//
// Stands where the reference's default sink runs on one host thread:
//   CompressionWiggleIteratorPop   /root/reference/src/unaryOps.c:235-253  (`write` only)
//   TeeWiggleIteratorPop           /root/reference/src/wigWriter.c:159-200 (10000-record blocks)
//   printBlock                     /root/reference/src/wigWriter.c:55-119  (fixedStep / bedGraph lines, "%lf")
// The segments of one chromosome are already in HBM (the reducer's output); this file
// turns them into the exact bytes the reference prints, so that only text crosses PCIe
// and the host does one fwrite per chromosome.
//
//  heads_kernel     H[k] = 1 where segment k certainly starts a new record: not adjacent to
//                   k-1, NaN-ness differs, or the values are >= 3e-6 apart.  (The reference
//                   compares with the value of the record's FIRST segment, tolerance 1e-6;
//                   anything within 1e-6 of that head is within 2e-6 of its neighbour, so a
//                   3e-6 step always opens a record.)
//                   D keeps these certain heads, H starts as a copy and receives the rest.
//  resolve_kernel   the remaining chains of near-equal neighbours, one warp per chain: 32
//                   members at a time are compared with the current head, the first one
//                   that is not within 1e-6 becomes the next head -- the reference's
//                   sequential rule, evaluated exactly.
//  count/scan/scatter  positions of the heads, hp[r] (3 small kernels).
//  format_kernel<WRITE>  one CTA per writer block of 10000 records (the reference resets
//                   its point-by-point state at every block, so blocks are independent):
//                   256 records at a time -- mode scan (set / clear / keep), header rule,
//                   exact "%lf" (round-half-even on the binary value, integer arithmetic),
//                   byte offsets by a block scan; the measuring pass returns the bytes per
//                   block, the writing pass formats into shared memory and copies out with
//                   128-bit stores.
#include "engine.cuh"
#include "device_util.cuh"
#include <climits>
#include <cstring>

namespace wb {
namespace sink {

constexpr int BLOCK_RECORDS = 10000;     // BLOCK_LENGTH, src/wigWriter.c:29
constexpr int TH = 256;
constexpr int TILE = 4096;               // elements per CTA in the head compaction
constexpr int OP_KEEP = 0, OP_SET = 1, OP_CLEAR = 2;

__device__ __forceinline__ bool close_v(double a, double b) {
	return (isnan(a) && isnan(b)) || fabs(a - b) < 0.000001;              // src/unaryOps.c:248
}

__global__ void heads_kernel(const int32_t *start, const int32_t *end, const double *value, int64_t n, unsigned char *D, unsigned char *H) {
	for (int64_t k = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; k < n; k += (int64_t) gridDim.x * blockDim.x) {
		bool brk = true;
		if (k > 0 && start[k] == end[k - 1]) {
			const double a = value[k], b = value[k - 1];
			const bool na = isnan(a), nb = isnan(b);
			brk = (na != nb) || (!na && !(fabs(a - b) < 0.000003));
		}
		D[k] = H[k] = brk ? 1 : 0;
	}
}

// one warp per tile of 1024 segments; a chain belongs to the tile its first segment lies in
__global__ void resolve_kernel(const double *value, int64_t n, const unsigned char *D, unsigned char *H) {
	const int lane = threadIdx.x & 31;
	const int64_t warp = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	const int64_t t0 = warp * 1024;
	if (t0 >= n) return;
	for (int g = 0; g < 32; g++) {
		const int64_t k = t0 + 32 * g + lane;
		const bool starts = k + 1 < n && D[k] != 0 && D[k + 1] == 0;
		unsigned todo = __ballot_sync(FULL, starts);
		while (todo) {
			const int q = __ffs(todo) - 1;
			todo &= todo - 1;
			int64_t head = t0 + 32 * g + q, pos = head + 1;
			double hv = value[head];
			while (true) {
				const int64_t j = pos + lane;
				const bool inb = j < n;
				const bool stop = inb ? D[j] != 0 : true;             // a certain head (or the end of the data) ends the chain
				const double v = inb ? value[j] : 0.0;
				const unsigned endmask = __ballot_sync(FULL, stop);
				const unsigned far = __ballot_sync(FULL, !close_v(v, hv));
				const unsigned members = endmask ? ((1u << (__ffs(endmask) - 1)) - 1u) : FULL;
				const unsigned nm = far & members;
				if (nm) {
					const int j0 = __ffs(nm) - 1;
					head = pos + j0;
					hv = __shfl_sync(FULL, v, j0);
					if (lane == 0) H[head] = 1;
					pos = head + 1;
					continue;
				}
				if (endmask) break;
				pos += 32;
			}
			__syncwarp();
		}
	}
}

__global__ void __launch_bounds__(TH) count_kernel(const unsigned char *H, int64_t n, int *tile_count) {
	__shared__ int wsum[TH / 32];
	const int64_t base = (int64_t) blockIdx.x * TILE;
	int c = 0;
	for (int i = threadIdx.x; i < TILE; i += TH) { const int64_t k = base + i; if (k < n && H[k]) c++; }
	c = __reduce_add_sync(FULL, c);
	if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < TH / 32; w++) t += wsum[w]; tile_count[blockIdx.x] = t; }
}

// exclusive scan of the tile counts (one CTA); total -> *R
__global__ void __launch_bounds__(1024) scan_kernel(const int *tile_count, int ntiles, long long *tile_base, long long *R) {
	__shared__ long long wsum[32];
	__shared__ long long carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int b = 0; b < ntiles; b += 1024) {
		const int i = b + threadIdx.x;
		const long long v = i < ntiles ? tile_count[i] : 0;
		long long inc = v;
		for (int d = 1; d < 32; d <<= 1) { const long long t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
		if (lane == 31) wsum[warp] = inc;
		__syncthreads();
		long long before = carry;
		for (int w = 0; w < warp; w++) before += wsum[w];
		if (i < ntiles) tile_base[i] = before + inc - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry = before + inc;
		__syncthreads();
	}
	if (threadIdx.x == 0) *R = carry;
}

__global__ void __launch_bounds__(TH) scatter_kernel(const unsigned char *H, int64_t n, const long long *tile_base, const long long *R, int32_t *hp) {
	__shared__ int wsum[TH / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t base = (int64_t) blockIdx.x * TILE;
	long long run = tile_base[blockIdx.x];
	for (int i0 = 0; i0 < TILE; i0 += TH) {
		const int64_t k = base + i0 + threadIdx.x;
		const bool h = k < n && H[k];
		const unsigned m = __ballot_sync(FULL, h);
		if (lane == 0) wsum[warp] = __popc(m);
		__syncthreads();
		int before = 0, total = 0;
		for (int w = 0; w < TH / 32; w++) { if (w < warp) before += wsum[w]; total += wsum[w]; }
		if (h) hp[run + before + __popc(m & ((1u << lane) - 1u))] = (int32_t) k;
		run += total;
		__syncthreads();
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) hp[*R] = (int32_t) n;          // sentinel: one past the last record
}

// ---- formatting ---------------------------------------------------------------------------
struct FormatJob {
	const int32_t *start, *end;
	const double *value;
	const int32_t *hp;          // [R+1] first segment of every record, or null: every segment is a record (bedGraph)
	long long R;                // records of this call
	long long rec_base;         // records handed to the sink before this call
	int wig;                    // 1: `write` (point-by-point allowed), 0: `write_bg`
	int pbp_in;                 // point-by-point state carried in (only used when the first block is a continued one)
	int chrom_len;
	char chrom[64];
	int *block_bytes;           // [nblocks] measuring pass
	const long long *block_off; // [nblocks] writing pass
	char *text;
	int *flags;                 // [0] a value the device formatter does not cover; [1] state after the last record
};

struct Dec { unsigned long long ip; unsigned frac; int neg, special; };   // special: 0 finite, 1 nan, 2 inf

// exact decimal of a double with 6 fractional digits, round-half-even on the binary value
// (what glibc's "%lf" prints); *bad is set for |v| >= 2^63
__host__ __device__ inline unsigned long long mulhi_u64(unsigned long long a, unsigned long long b) {
#ifdef __CUDA_ARCH__
	return __umul64hi(a, b);
#else
	return (unsigned long long) (((unsigned __int128) a * b) >> 64);
#endif
}
__host__ __device__ inline unsigned long long bits_of(double v) {
#ifdef __CUDA_ARCH__
	return (unsigned long long) __double_as_longlong(v);
#else
	unsigned long long b; memcpy(&b, &v, 8); return b;
#endif
}

// (also compiled for the host: wb200_debug_format_value lets a CPU test fuzz this very code against printf)
__host__ __device__ Dec to_dec(double v, int *bad) {
	Dec d;
	const unsigned long long bits = bits_of(v);
	d.neg = (int) (bits >> 63);
	const int ex = (int) ((bits >> 52) & 0x7ff);
	const unsigned long long man = bits & ((1ull << 52) - 1);
	d.ip = 0; d.frac = 0; d.special = 0;
	if (ex == 0x7ff) { d.special = man ? 1 : 2; return d; }
	const unsigned long long m = ex ? (man | (1ull << 52)) : man;
	const int e = (ex ? ex : 1) - 1075;
	if (e >= 0) {
		if (e > 10) { *bad = 1; return d; }
		d.ip = m << e;
		return d;
	}
	const int k = -e;
	unsigned long long fm = m;
	if (k < 64) { d.ip = m >> k; fm = m & ((1ull << k) - 1); }
	const unsigned long long lo = fm * 1000000ull, hi = mulhi_u64(fm, 1000000ull);   // fm * 10^6 < 2^73
	if (k >= 128) return d;                                              // below half of the last digit
	unsigned long long q, rem_hi, rem_lo, half_hi, half_lo;
	if (k < 64) {
		q = (lo >> k) | (hi << (64 - k));
		rem_hi = 0; rem_lo = lo & ((1ull << k) - 1);
		half_hi = 0; half_lo = 1ull << (k - 1);
	} else if (k == 64) {
		q = hi; rem_hi = 0; rem_lo = lo; half_hi = 0; half_lo = 1ull << 63;
	} else {
		const int kk = k - 64;                                           // 1..63
		q = hi >> kk;
		rem_hi = hi & ((1ull << kk) - 1); rem_lo = lo;
		half_hi = 1ull << (kk - 1); half_lo = 0;
	}
	const bool gt = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
	const bool eq = rem_hi == half_hi && rem_lo == half_lo;
	if (gt || (eq && (q & 1))) q++;
	if (q >= 1000000ull) { q -= 1000000ull; d.ip++; }
	d.frac = (unsigned) q;
	return d;
}

__host__ __device__ __forceinline__ int ndigits_u64(unsigned long long x) {
	int n = 1;
	while (x >= 10) { x /= 10; n++; }
	return n;
}
__host__ __device__ __forceinline__ int len_int(int x) {
	const unsigned a = x < 0 ? 0u - (unsigned) x : (unsigned) x;
	return (x < 0 ? 1 : 0) + ndigits_u64(a);
}
__host__ __device__ __forceinline__ int len_value(const Dec &d) {
	return d.neg + (d.special ? 3 : ndigits_u64(d.ip) + 7);
}
__host__ __device__ __forceinline__ char *put_u64(char *p, unsigned long long x) {
	const int n = ndigits_u64(x);
	for (int i = n - 1; i >= 0; i--) { p[i] = (char) ('0' + (int) (x % 10)); x /= 10; }
	return p + n;
}
__host__ __device__ __forceinline__ char *put_int(char *p, int x) {
	if (x < 0) *p++ = '-';
	return put_u64(p, x < 0 ? 0u - (unsigned) x : (unsigned) x);
}
__host__ __device__ __forceinline__ char *put_value(char *p, const Dec &d) {
	if (d.neg) *p++ = '-';
	if (d.special == 1) { *p++ = 'n'; *p++ = 'a'; *p++ = 'n'; return p; }
	if (d.special == 2) { *p++ = 'i'; *p++ = 'n'; *p++ = 'f'; return p; }
	p = put_u64(p, d.ip);
	*p++ = '.';
	unsigned f = d.frac;
	for (int i = 5; i >= 0; i--) { p[i] = (char) ('0' + (int) (f % 10)); f /= 10; }
	return p + 6;
}
__host__ __device__ __forceinline__ char *put_str(char *p, const char *s, int n) {
	for (int i = 0; i < n; i++) p[i] = s[i];
	return p + n;
}

// bytes one record can take at most: header + five value lines (a longer record leaves the mode)
__host__ __device__ inline int max_record_bytes(int chrom_len) { return 16 + chrom_len + 7 + 11 + 8 + 5 * 29 + 8; }

template <bool WRITE>
__global__ void __launch_bounds__(TH) format_kernel(const FormatJob job) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ int s_fin[TH];
	__shared__ int s_wop[TH / 32], s_wbytes[TH / 32];
	__shared__ int s_pbp, s_prevf, s_bad;
	char *stage = (char *) smem_raw;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const long long b0 = job.rec_base / BLOCK_RECORDS;
	const long long g_lo = (b0 + blockIdx.x) * BLOCK_RECORDS;
	const long long lo = max(g_lo, job.rec_base) - job.rec_base;
	const long long hi = min(g_lo + BLOCK_RECORDS, job.rec_base + job.R) - job.rec_base;
	if (tid == 0) {
		s_pbp = (blockIdx.x == 0 && job.rec_base % BLOCK_RECORDS != 0) ? job.pbp_in : 0;
		s_prevf = -1;
		s_bad = 0;
	}
	__syncthreads();
	long long out = WRITE ? job.block_off[blockIdx.x] : 0;
	int total_block = 0;
	const int C = job.chrom_len;
	for (long long c0 = lo; c0 < hi; c0 += TH) {
		const long long r = c0 + tid;
		const bool valid = r < hi;
		int s = 0, f = 0;
		double v = 0;
		if (valid) {
			const long long k = job.hp ? job.hp[r] : r;
			const long long kl = job.hp ? (long long) job.hp[r + 1] - 1 : r;
			s = job.start[k]; f = job.end[kl]; v = job.value[k];
		}
		const int len = f - s;
		// mode change, src/wigWriter.c:69-74: set below two positions, cleared above five
		const int op = !valid ? OP_KEEP : (job.wig && len < 2) ? OP_SET : (len > 5 ? OP_CLEAR : OP_KEEP);
		int inc = op;                                                  // inclusive scan: the last non-KEEP wins
		for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, inc, d); if (lane >= d && inc == OP_KEEP) inc = t; }
		if (lane == 31) s_wop[warp] = inc;
		s_fin[tid] = f;
		__syncthreads();
		int before = __shfl_up_sync(FULL, inc, 1);
		if (lane == 0) before = OP_KEEP;
		for (int w = warp - 1; w >= 0 && before == OP_KEEP; w--) before = s_wop[w];
		const int carry_pbp = s_pbp, carry_f = s_prevf;
		const int prev = before == OP_KEEP ? carry_pbp : (before == OP_SET ? 1 : 0);
		const int now = op == OP_KEEP ? prev : (op == OP_SET ? 1 : 0);
		const bool make_header = op == OP_SET && !prev;
		// the first record of a block has no predecessor, the first of a call follows another chromosome
		const bool forced = r == lo;
		const int pf = tid > 0 ? s_fin[tid - 1] : carry_f;
		const bool header = now && (make_header || forced || s > pf);       // src/wigWriter.c:77-78
		Dec d;
		int bytes = 0, vl = 0;
		if (valid) {
			int bad = 0;
			d = to_dec(v, &bad);
			if (bad) s_bad = 1;
			vl = len_value(d);
			if (now) bytes = (header ? 16 + C + 7 + len_int(s) + 8 : 0) + (len > 0 ? len : 0) * (vl + 1);
			else bytes = C + 1 + len_int(s - 1) + 1 + len_int(f - 1) + 1 + vl + 1;
		}
		int binc = bytes;
		for (int dd = 1; dd < 32; dd <<= 1) { const int t = __shfl_up_sync(FULL, binc, dd); if (lane >= dd) binc += t; }
		if (lane == 31) s_wbytes[warp] = binc;
		__syncthreads();
		int off = binc - bytes, chunk_total = 0;
		for (int w = 0; w < TH / 32; w++) { if (w < warp) off += s_wbytes[w]; chunk_total += s_wbytes[w]; }
		if (WRITE) {
			// the staging buffer is shifted so that it is 16-byte congruent with the destination
			const int shift = (int) ((unsigned long long) (job.text + out) & 15ull);
			if (valid) {
				char *p = stage + shift + off;
				if (now) {
					if (header) {
						p = put_str(p, "fixedStep chrom=", 16);
						p = put_str(p, job.chrom, C);
						p = put_str(p, " start=", 7);
						p = put_int(p, s);
						p = put_str(p, " step=1\n", 8);
					}
					for (int j = 0; j < len; j++) { p = put_value(p, d); *p++ = '\n'; }
				} else {
					p = put_str(p, job.chrom, C);
					*p++ = '\t';
					p = put_int(p, s - 1);                                 // bedGraph lines are 0-based, src/wigWriter.c:83-84
					*p++ = '\t';
					p = put_int(p, f - 1);
					*p++ = '\t';
					p = put_value(p, d);
					*p++ = '\n';
				}
			}
			__syncthreads();
			char *dst = job.text + out - shift;                            // 16-byte aligned
			const int first = shift, last = shift + chunk_total;           // staged bytes [first, last)
			const int a0 = (first + 15) & ~15, a1 = last & ~15;
			if (a0 <= a1) {
				for (int i = first + tid; i < a0; i += TH) dst[i] = stage[i];
				for (int i = a0 + 16 * tid; i < a1; i += 16 * TH) *(uint4 *) (dst + i) = *(const uint4 *) (stage + i);
				for (int i = a1 + tid; i < last; i += TH) dst[i] = stage[i];
			} else {
				for (int i = first + tid; i < last; i += TH) dst[i] = stage[i];
			}
			out += chunk_total;
		}
		total_block += chunk_total;
		// carry to the next 256 records
		__syncthreads();
		const long long last_r = min(hi, c0 + TH) - 1;
		if (r == last_r) { s_pbp = now; s_prevf = f; }
		__syncthreads();
	}
	if (tid == 0) {
		if (!WRITE) {
			job.block_bytes[blockIdx.x] = total_block;
			if (s_bad) job.flags[0] = 1;
			if (blockIdx.x == gridDim.x - 1) job.flags[1] = s_pbp;
		}
	}
}

} // namespace sink

// ---- launch helpers (declared in engine.cuh) ------------------------------------------------
cudaError_t sink_heads(const int32_t *start, const int32_t *end, const double *value, int64_t n, unsigned char *D, unsigned char *H,
                       int *tile_count, long long *tile_base, long long *R, int32_t *hp, cudaStream_t s) {
	if (n <= 0) return cudaSuccess;
	const int grid = (int) std::min<int64_t>((n + 255) / 256, 148 * 16);
	sink::heads_kernel<<<grid, 256, 0, s>>>(start, end, value, n, D, H);
	const int64_t warps = (n + 1023) / 1024;
	sink::resolve_kernel<<<(unsigned) ((warps + 3) / 4), 128, 0, s>>>(value, n, D, H);
	return sink_compact(H, n, tile_count, tile_base, R, hp, s);
}

cudaError_t sink_compact(const unsigned char *H, int64_t n, int *tile_count, long long *tile_base, long long *R, int32_t *hp, cudaStream_t s) {
	if (n <= 0) return cudaSuccess;
	const int ntiles = (int) ((n + sink::TILE - 1) / sink::TILE);
	sink::count_kernel<<<ntiles, sink::TH, 0, s>>>(H, n, tile_count);
	sink::scan_kernel<<<1, 1024, 0, s>>>(tile_count, ntiles, tile_base, R);
	sink::scatter_kernel<<<ntiles, sink::TH, 0, s>>>(H, n, tile_base, R, hp);
	return cudaGetLastError();
}

// host build of the device formatter's value path: "%lf" of one double into out (>= 32 bytes), returns
// the length, -1 for a value outside the formatter's range (tests/test_writer.py fuzzes it against printf)
int sink_debug_format_value(double v, char *out) {
	int bad = 0;
	const sink::Dec d = sink::to_dec(v, &bad);
	if (bad) return -1;
	char *p = sink::put_value(out, d);
	*p = 0;
	return (int) (p - out);
}

int sink_tiles(int64_t n) { return (int) ((n + sink::TILE - 1) / sink::TILE); }
int sink_block_records() { return sink::BLOCK_RECORDS; }

cudaError_t sink_format(bool write, const int32_t *start, const int32_t *end, const double *value, const int32_t *hp,
                        long long R, long long rec_base, int wig, int pbp_in, const char *chrom,
                        int *block_bytes, const long long *block_off, char *text, int *flags, int nblocks, cudaStream_t s) {
	if (nblocks <= 0) return cudaSuccess;
	sink::FormatJob job{};
	job.start = start; job.end = end; job.value = value; job.hp = hp;
	job.R = R; job.rec_base = rec_base; job.wig = wig; job.pbp_in = pbp_in;
	int cl = 0;
	while (chrom[cl] && cl < 63) { job.chrom[cl] = chrom[cl]; cl++; }
	job.chrom_len = cl;
	job.block_bytes = block_bytes; job.block_off = block_off; job.text = text; job.flags = flags;
	const size_t smem = (size_t) sink::TH * sink::max_record_bytes(cl) + 32;
	if (write) {
		cudaError_t e = cudaFuncSetAttribute(sink::format_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if (e != cudaSuccess) return e;
		sink::format_kernel<true><<<nblocks, sink::TH, smem, s>>>(job);
	} else {
		sink::format_kernel<false><<<nblocks, sink::TH, 0, s>>>(job);
	}
	return cudaGetLastError();
}

} // namespace wb
