// wiggletools_b200 -- the sum / mean sweep kernel (BASELINE configs[1]: `mean`
// over N tracks; reference: popCoreMultiplexer2 + MeanReductionPop / SumReductionPop,
// /root/reference/src/multiplexer.c:98-121, src/reducers.c:259-292, 367-402).
//
// Same algorithm as emit_kernel<MODE_FAST> in engine.cu (position-indexed event
// deltas, exact fixed-point prefix sums, decoupled look-back, coalesced emit) laid
// out for latency hiding: measured per-window time on a CTA is a chain of short,
// latency-bound phases whose length hardly depends on the window's content, so
// throughput is (windows in flight per SM) x (window size) / latency.  This
// kernel therefore uses
//   * 128 threads per window (32 positions per thread), 64 registers;
//   * 6 bytes of shared memory per position (N <= 255, one-limb values) -> 27 KB per
//     window, 8 windows in flight per SM:
//       cover field : starts in the low byte, ends in the high byte of a 16-bit field,
//                     two positions per 32-bit word (the atomics are 32-bit adds of a
//                     shifted one) -- a position is a breakpoint iff its field is
//                     non-zero, so there is no flag array.  N > 255: 32-bit fields
//                     (11 + 11 bits);
//       value       : one or two 32-bit limbs (host-chosen from the value range);
//       poison      : NaN / non-representable values are rare and go to a short
//                     side list instead of a field in every word;
//   * a scan that never holds the window in registers: pass 1 sums each thread's
//     32 positions, one block scan stitches the threads, pass 2 re-reads them and
//     writes the exact prefix back IN PLACE; the dense breakpoint list is one
//     16-bit entry per breakpoint (position | keep | slow) over the dead cover
//     array, the output rank of an entry is a per-32-entries base plus a ballot
//     popcount, and the emit phase fetches the value by position;
//   * a flat, load-balanced events phase: the intervals of a window are read as
//     16-byte aligned quads (128-bit loads) shared out evenly over the threads by a
//     block scan + binary search, whatever their distribution over the tracks;
//   * bulk-async (TMA engine) tiles for the per-window metadata: the two cursor rows
//     that bound a window's intervals arrive as one cp.async.bulk global->shared copy
//     signalled on an mbarrier (UBLKCP / SYNCS), requested when the window is drawn;
//     the interval slices of the window some CTA will draw next are pulled into L2
//     by warps 1..3 while warp 0 runs the look-back.
// Windows are drawn from a ticket counter from the first one on, so the look-back
// only ever waits on windows a running CTA owns: no assumption that the whole grid
// is resident.
// Output is bit-identical to emit_kernel<MODE_FAST>; anything this kernel does
// not cover (three limbs, unaligned inputs, other window sizes) stays there.
#include "engine.cuh"
#include "device_util.cuh"
#include <climits>
#include <type_traits>
#include <cstdlib>
#include <cstdio>

namespace wb {
namespace fs {

constexpr int TH = 128, NWARP = TH / 32;
#ifndef WB_FS_MINB
#define WB_FS_MINB 8                  // resident CTAs per SM the 6 B/position variant is compiled for
#endif
#ifndef WB_FS_MOM_MINB
#define WB_FS_MOM_MINB 6              // var / stddev / CV, N <= 255, 2048-bp windows: resident CTAs per SM compiled for
#endif
#ifndef WB_FS_LBK
#define WB_FS_LBK 1
#endif
#ifndef WB_FS_LBSLEEP
#define WB_FS_LBSLEEP 200             // ns between two polls of the look-back (first wait; doubles up to 8x)
#endif
#ifndef WB_FS_PREF
#define WB_FS_PREF 1                  // next window's intervals: 0 no prefetch, 1 prefetch into L2
#endif
#ifndef WB_FS_COV16
#define WB_FS_COV16 1                 // 16-bit cover fields (6 B/position) when N <= 255
#endif
#ifndef WB_FS_TICKET_AT_END
#define WB_FS_TICKET_AT_END 1         // the next window is drawn when this one is completely done (0: when its emit phase starts)
#endif
#ifndef WB_FS_EMIT2
#define WB_FS_EMIT2 1                 // emit: two groups of 32 dense entries per warp and step
#endif
#ifndef WB_FS_MERGE
#define WB_FS_MERGE 0                 // events: an interval that abuts its predecessor shares its atomics with that one's end.
                                      // Measured (profiles/r02_kexp8_*): a quarter fewer shared atomics on the generator, and SLOWER --
                                      // N=64 +8.6 %, N=512 +19 %, var N=512 +20 % kernel time: the held-back end event serialises the
                                      // four intervals of a quad.  Off; kept for gap-free inputs where half of all events would merge.
#endif
#if defined(WB_DIAG_NO_COV_ATOM) || defined(WB_DIAG_NO_LIMB_ATOM)
#undef WB_FS_MERGE
#define WB_FS_MERGE 0
#endif
#ifndef WB_FS_EVPIPE
#define WB_FS_EVPIPE 0                // events: software-pipelined interval loads (two chunks in registers)
#endif
constexpr int END_SHIFT = 11;            // 32-bit cover field: starts | ends << 11 (each <= 1023 per position)
constexpr uint32_t FIELD = (1u << END_SHIFT) - 1;
constexpr int LBK = WB_FS_LBK;           // look-back: 32 * LBK predecessors per step
constexpr int POIS_CAP = 120;            // poison events kept per window; more -> the whole window takes the fix-up path
constexpr unsigned DN_KEEP = 0x8000, DN_SLOW = 0x4000, DN_POS = 0x1fff;   // dense entry: position | flags

struct PoisEv { uint16_t pos; int16_t delta; };

struct Misc {
	int next_bp;                     // first breakpoint at or after the window end
	int base_cov;                    // tracks in play at the window start
	int base_pois;                   // of those, tracks whose value is not representable
	int npois;                       // poison events inside the window
	unsigned long long base_val;     // fixed-point sum carried into the window
	unsigned long long base_val2;    // moments: fixed-point sum of squares carried into the window
	long long out_base;              // output offset of the window
	int ticket2;
	int pad;
	unsigned long long bar_cur;      // mbarrier: cursor row of the window being started
	unsigned long long bar_nxt;      // mbarrier: cursor row of the window prefetched for
	long long wsum[NWARP];
	long long wsum2[NWARP];
	int wcov[NWARP];
	int wpack[NWARP];
	double wauc[NWARP];
	uint16_t kbase[128];             // kept segments before every 32nd dense entry (W <= 4096)
	PoisEv pois[POIS_CAP];
};
constexpr int MISC_BYTES = 1024;
static_assert(sizeof(Misc) <= MISC_BYTES, "Misc must fit its slot");

struct Layout { size_t off_cur, off_cur2, off_pre, off_toff, off_tmd, off_cov, off_limb, off_cnt, total; };
__host__ __device__ inline int quad_table_size(int N) { int p = 2; while (p <= N) p <<= 1; return p; }   // power of two > N
__host__ __device__ inline Layout make_layout(int wlog, int N, int limbs, bool cov16, bool mom) {
	Layout L;
	const size_t W = (size_t) 1 << wlog, nn = (size_t) (N + 4) / 4 * 16;
	size_t o = MISC_BYTES;
	L.off_cur = o;  o += 2 * nn;                 // cursor rows w and w + 1 (one bulk copy)
	L.off_cur2 = o; o += nn;                     // cursor row of the window prefetched for
	L.off_pre = o;  o += (4 * (size_t) quad_table_size(N) + 15) / 16 * 16;   // exclusive prefix of the tracks' quad counts
	L.off_toff = o; o += 2 * nn;                 // N + 1 track offsets (int64)
	L.off_tmd = o;  o += 2 * nn;
	L.off_cov = o;  o += (cov16 ? 2 : 4) * W;
	L.off_limb = o; o += (size_t) (mom ? 3 : limbs) * 4 * W;   // moments: sum (one limb), sum of squares (two)
	L.off_cnt = o;  o += mom ? 2 * W : 0;        // moments: tracks in play at every breakpoint (16 bits)
	L.total = (o + 15) / 16 * 16;
	return L;
}

// ---- bulk-async copies and their mbarriers (PTX; SASS: UBLKCP / SYNCS) ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred P1;\n"
		"LAB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		"@P1 bra DONE;\n"
		"bra LAB_WAIT;\n"
		"DONE:\n"
		"}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one contiguous tile global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Thread t owns the V = 2^VLOG consecutive positions [t*V, t*V+V) and reads them back
// with 128-bit loads; its NV 16-byte vectors are rotated by a function of t so that the 8
// lanes of a quarter warp hit 8 different 16-byte bank groups.
template <int NV>
__device__ __forceinline__ int swz(int t) {
	return NV >= 8 ? (t & (NV - 1)) : NV == 4 ? ((t >> 1) & 3) : NV == 2 ? ((t >> 2) & 1) : 0;
}
// the same rotation applied to a position: index of its limb word / of its 16-bit cover field
template <int VLOG> __device__ __forceinline__ int phys32(int p) { return p ^ ((p >> 3) & (VLOG == 5 ? 0x1c : 0xc)); }
template <int VLOG> __device__ __forceinline__ int phys16(int p) { return p ^ ((p >> 3) & (VLOG == 5 ? 0x18 : 0x8)); }

template <int VLOG, bool COV16>
__device__ __forceinline__ void cov_add(uint32_t *cov, int p, bool is_end) {
#ifdef WB_DIAG_NO_COV_ATOM                 // diagnostic builds only (profiles/): cost of the cover atomics
	if (is_end || p != 0) return;
#endif
	if (COV16) atomicAdd(&cov[phys16<VLOG>(p) >> 1], (is_end ? 0x100u : 1u) << ((p & 1) << 4));
	else atomicAdd(&cov[phys32<VLOG>(p)], is_end ? (1u << END_SHIFT) : 1u);
}

// an end inside the window and an abutting start: one atomic for both counts
template <int VLOG, bool COV16>
__device__ __forceinline__ void cov_add_both(uint32_t *cov, int p) {
	if (COV16) atomicAdd(&cov[phys16<VLOG>(p) >> 1], 0x101u << ((p & 1) << 4));
	else atomicAdd(&cov[phys32<VLOG>(p)], (1u << END_SHIFT) | 1u);
}

struct Chunk {
	int4 s, e;
	double v0, v1, v2, v3;
	int lo, hi;          // slots lo <= x < hi hold intervals of this track at or after its cursor
};

// ---- events: +1 start / +1 end per interval in the cover field, +-value in the limbs ----
// The intervals of track i that can touch window w are [cursor[w][i], cursor[w+1][i]] (the last one
// may begin beyond the window: it only yields the look-ahead breakpoint).  They are read as QUADS,
// 16-byte aligned groups of four intervals (128-bit loads of start, end and twice value); the
// quads of all tracks form one flat list that the threads share out evenly: a block scan over
// the tracks' quad counts, then a binary search per quad.  No lane idles on a track that has
// fewer intervals than its neighbours, and a track with thousands of intervals in the window
// costs no more per interval than any other.
template <int VLOG, int LIMBS, bool COV16, bool MOM>
__device__ __forceinline__ void events(const Job &job, unsigned char *smem, const Layout &L, const int32_t *cur,
                                       int w0i, int lw, double inv_scale, double lim) {
	constexpr int W = TH << VLOG;
	Misc *misc = (Misc *) smem;
	const long long *toff = (const long long *) (smem + L.off_toff);
	const long long *tmd = (const long long *) (smem + L.off_tmd);
	int32_t *pre = (int32_t *) (smem + L.off_pre);
	auto tn = [&](int i) -> int { return (int) (toff[i + 1] - toff[i]); };     // intervals of track i
	uint32_t *cov = (uint32_t *) (smem + L.off_cov);
	uint32_t *limb = (uint32_t *) (smem + L.off_limb);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = job.N;
	const int32_t *nxt = cur + job.cstride;                      // cursor row of window w + 1
	using acc_t = typename std::conditional<LIMBS == 1, int, long long>::type;
	int in_cov = 0, in_pois = 0, my_next = INT_MAX;
	acc_t in_val = 0;
	long long in_val2 = 0;                                       // moments: squares

	// quads of track i: its intervals first..last on the 16-byte aligned index grid
	auto quads = [&](int i) -> int {
		const int first = cur[i], last = min(nxt[i], tn(i) - 1);
		if (last < first) return 0;
		return (((int) (toff[i] + first) & 3) + (last - first + 1) + 3) >> 2;
	};
	const int P2 = quad_table_size(N);
	{
		const int K = (N + TH - 1) / TH;                         // consecutive tracks per thread
		int mine = 0;
		for (int k = 0; k < K; k++) { const int i = tid * K + k; if (i < N) mine += quads(i); }
		int inc = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += v; }
		if (lane == 31) misc->wpack[warp] = inc;
		__syncthreads();
		int run = inc - mine;
#pragma unroll
		for (int k = 0; k < NWARP; k++) if (k < warp) run += misc->wpack[k];
		for (int k = 0; k < K; k++) { const int i = tid * K + k; if (i < N) { pre[i] = run; run += quads(i); } }
		if (tid == TH - 1) pre[N] = run;                         // Q, the number of quads of the window
		for (int i = N + 1 + tid; i < P2; i += TH) pre[i] = INT_MAX;
		__syncthreads();
	}
	const int Q = pre[N];

	auto poison = [&](unsigned x, int delta) {
		const int k = atomicAdd(&misc->npois, 1);
		if (k < POIS_CAP) { PoisEv p; p.pos = (uint16_t) x; p.delta = (int16_t) delta; misc->pois[k] = p; }
	};
	// The end event of an interval is held back until the next interval of the quad is seen: where the two abut
	// (binned coverage, bedGraph runs) the position gets ONE cover atomic (an end and a start) and ONE value atomic
	// (the difference of the two deltas; none at all when the values are equal) instead of two of each.
	bool pend = false;
	int pend_x = 0;
	acc_t pend_d = 0;
	long long pend_d2 = 0;
	auto flush_end = [&]() {
		if (!pend) return;
		pend = false;
		cov_add<VLOG, COV16>(cov, pend_x, true);
#ifndef WB_DIAG_NO_LIMB_ATOM
		if (pend_d != 0) {
			add_limbs_fs<LIMBS, W>(limb, phys32<VLOG>(pend_x), -pend_d, lw);
			if (MOM) add_limbs_fs<2, W>(limb + W, phys32<VLOG>(pend_x), -pend_d2, lw);
		}
#endif
	};
	auto one = [&](bool valid, int s, int e, double val, long long md) {
		if (!valid) return;
		const int xs = s - w0i;
		const unsigned xe = (unsigned) (e - w0i);
		if (xs >= W) { my_next = min(my_next, s); return; }         // begins at or beyond the window end
		const bool start_in = xs >= 0, end_in = xe < (unsigned) W;
		acc_t d = 0;
		bool bad;
		if (LIMBS == 1) {
			// one limb: |m| and the defaults are below 2^29 / 2N (capi.cu), so the whole
			// delta arithmetic is 32-bit; same acceptance rule as to_fixed()
			const double t = val * inv_scale;
			const int mi = __double2int_rn(t);
			bad = !((double) mi == t && fabs(t) < lim && (mi != 0 || val == 0.0));
			if (MOM) bad = bad || (double) (float) val != val;      // var / stddev / CV: the reference rounds values through float
			if (!bad) d = (acc_t) (mi - (int) md);
		} else {
			long long m;
			bad = !to_fixed(val, inv_scale, lim, m);
			if (!bad) d = (acc_t) (m - md);
		}
		const long long d2 = MOM ? (long long) d * (long long) d : 0;        // moments: all defaults are 0, d is the value
		bool merged = false;
		if (WB_FS_MERGE && pend) {
			if (start_in && xs == pend_x) {
				pend = false;
				merged = true;
				cov_add_both<VLOG, COV16>(cov, xs);
				const acc_t dd = d - pend_d;
				if (dd != 0) {
					add_limbs_fs<LIMBS, W>(limb, phys32<VLOG>(xs), dd, lw);
					if (MOM && d2 != pend_d2) add_limbs_fs<2, W>(limb + W, phys32<VLOG>(xs), d2 - pend_d2, lw);
				}
			} else flush_end();
		}
		if (start_in) {
			if (!merged) {
				cov_add<VLOG, COV16>(cov, xs, false);
#ifndef WB_DIAG_NO_LIMB_ATOM               // diagnostic builds only (profiles/): cost of the value atomics
				if (d != 0) {
					add_limbs_fs<LIMBS, W>(limb, phys32<VLOG>(xs), d, lw);
					if (MOM) add_limbs_fs<2, W>(limb + W, phys32<VLOG>(xs), d2, lw);
				}
#endif
			}
			if (bad) poison((unsigned) xs, 1);
		} else { in_cov++; in_val += d; in_val2 += d2; in_pois += bad ? 1 : 0; }           // began before the window
		if (end_in) {
			pend = true; pend_x = (int) xe; pend_d = d; pend_d2 = d2;
			if (!WB_FS_MERGE) flush_end();
			if (bad) poison(xe, -1);
		} else my_next = min(my_next, e);
	};
	// quad x of the flat list: its track (largest i with pre[i] <= x) and its intervals
	auto load = [&](int x, Chunk &c, long long &md) {
		c.lo = c.hi = 0;
		if (x >= Q) return;
		int i = 0;
		for (int st = P2 >> 1; st > 0; st >>= 1) if (pre[i + st] <= x) i += st;
		const long long o = toff[i], first = o + cur[i], end1 = o + min(nxt[i], tn(i) - 1) + 1;
		const long long g = (first & ~3ll) + 4ll * (x - pre[i]);    // 16-byte aligned index grid
		md = tmd[i];
		c.lo = (int) max(first - g, 0ll);
		c.hi = (int) min(end1 - g, 4ll);
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
	auto apply = [&](const Chunk &c, long long md) {
		one(c.lo <= 0 && 0 < c.hi, c.s.x, c.e.x, c.v0, md);
		one(c.lo <= 1 && 1 < c.hi, c.s.y, c.e.y, c.v1, md);
		one(c.lo <= 2 && 2 < c.hi, c.s.z, c.e.z, c.v2, md);
		one(c.lo <= 3 && 3 < c.hi, c.s.w, c.e.w, c.v3, md);
		flush_end();
	};
#if WB_FS_EVPIPE
	// two quads per thread and step: both sets of loads are in flight before the first is applied
	for (int x = tid; x < Q; x += 2 * TH) {
		Chunk c0, c1;
		long long md0 = 0, md1 = 0;
		load(x, c0, md0);
		load(x + TH, c1, md1);
		apply(c0, md0);
		apply(c1, md1);
	}
#else
	for (int x = tid; x < Q; x += TH) {
		Chunk c;
		long long md = 0;
		load(x, c, md);
		apply(c, md);
	}
#endif
	// carried-in state and the look-ahead breakpoint: one shared atomic per warp
	my_next = __reduce_min_sync(FULL, my_next);
	in_cov = __reduce_add_sync(FULL, in_cov);
	in_pois = __reduce_add_sync(FULL, in_pois);
	if (LIMBS == 1) in_val = (acc_t) __reduce_add_sync(FULL, (int) in_val);
	else for (int d = 16; d > 0; d >>= 1) in_val += __shfl_down_sync(FULL, in_val, d);
	if (MOM) for (int d = 16; d > 0; d >>= 1) in_val2 += __shfl_down_sync(FULL, in_val2, d);
	if (lane == 0) {
		if (my_next != INT_MAX) atomicMin(&misc->next_bp, my_next);
		if (in_cov) atomicAdd(&misc->base_cov, in_cov);
		if (in_pois) atomicAdd(&misc->base_pois, in_pois);
		if (in_val) atomicAdd(&misc->base_val, (unsigned long long) (long long) in_val);
		if (MOM && in_val2) atomicAdd(&misc->base_val2, (unsigned long long) in_val2);
	}
}

// the cover fields of four consecutive positions: the 16-byte vector (or half of it) they live in
template <bool COV16>
__device__ __forceinline__ void cover4(const uint4 &cv, int half, int st[4], int en[4]) {
	if (COV16) {
		const uint32_t a = half ? cv.z : cv.x, b = half ? cv.w : cv.y;
		st[0] = a & 0xff; en[0] = (a >> 8) & 0xff; st[1] = (a >> 16) & 0xff; en[1] = a >> 24;
		st[2] = b & 0xff; en[2] = (b >> 8) & 0xff; st[3] = (b >> 16) & 0xff; en[3] = b >> 24;
	} else {
		const uint32_t c[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
		for (int j = 0; j < 4; j++) { st[j] = c[j] & FIELD; en[j] = c[j] >> END_SHIFT; }
	}
}

// ---- one window ------------------------------------------------------------------------
template <int VLOG, int LIMBS, bool COV16, bool MOM>
__device__ void process_window(const Job &job, int w, unsigned char *smem, const Layout &L, int32_t *cur,
                               int32_t *nxt_buf, unsigned par_cur, unsigned par_nxt, int lw, double inv_scale, double lim, double scale) {
	constexpr int V = 1 << VLOG, W = TH << VLOG, NVL = V / 4;
	constexpr int NVC = COV16 ? V / 8 : V / 4;                // 16-byte cover vectors per thread
	static_assert(V <= 32, "one 32-bit mask per thread");
	Misc *misc = (Misc *) smem;
	uint32_t *cov = (uint32_t *) (smem + L.off_cov);
	uint32_t *limb = (uint32_t *) (smem + L.off_limb);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = job.N;
	const long long w0 = job.p0 + ((long long) w << job.wlog);
	const int w0i = (int) w0;
	const int thr = job.strict ? N : 1;
#ifdef WB_DIAG_PHASES
	long long diag_t = clock64();
#endif

	// ---- clear ------------------------------------------------------------------------
	{
		const uint4 z = make_uint4(0, 0, 0, 0);
		uint4 *zc = (uint4 *) cov;                                   // cov and limbs are contiguous
		constexpr int n16 = (COV16 ? W / 8 : W / 4) + (MOM ? 3 : LIMBS) * W / 4;
#pragma unroll 4
		for (int t = tid; t < n16; t += TH) zc[t] = z;
		if (tid == 0) { misc->next_bp = INT_MAX; misc->base_cov = 0; misc->base_pois = 0; misc->npois = 0; misc->base_val = 0; misc->base_val2 = 0; }
	}
	mbar_wait(&misc->bar_cur, par_cur);                          // this window's cursor row (bulk copy requested when it was drawn)
	__syncthreads();
	WB_PHASE(0);

	// ---- 1. events ---------------------------------------------------------------------
#ifndef WB_DIAG_SKIP_EVENTS
	events<VLOG, LIMBS, COV16, MOM>(job, smem, L, cur, w0i, lw, inv_scale, lim);
#endif
	__syncthreads();
	WB_PHASE(1);

	// ---- 2. scan, pass 1: totals of this thread's V positions -------------------------------
	const int p0 = tid << VLOG, swl = swz<NVL>(tid), swc = swz<NVC>(tid);
	const uint32_t *covt = cov + (COV16 ? p0 >> 1 : p0);         // this thread's cover words
	using acc_t = typename std::conditional<LIMBS == 1, int, long long>::type;   // one limb: every prefix fits 32 bits
	int ctot = 0;
	acc_t atot = 0;
	long long btot = 0;                                          // moments: squares
	uint32_t *sq = limb + W;                                     // their two limbs: sq[0..W) low, sq[W..2W) high
	{
		if (COV16) {
			// byte fields summed two positions at a time in 16-bit lanes (<= 16 * 255 each)
			uint32_t ss = 0, se = 0;
#pragma unroll
			for (int k = 0; k < NVC; k++) {
				const uint4 c = *(const uint4 *) (covt + ((k ^ swc) << 2));
				ss += (c.x & 0x00ff00ffu) + (c.y & 0x00ff00ffu) + (c.z & 0x00ff00ffu) + (c.w & 0x00ff00ffu);
				se += ((c.x >> 8) & 0x00ff00ffu) + ((c.y >> 8) & 0x00ff00ffu) + ((c.z >> 8) & 0x00ff00ffu) + ((c.w >> 8) & 0x00ff00ffu);
			}
			ctot = (int) ((ss & 0xffff) + (ss >> 16)) - (int) ((se & 0xffff) + (se >> 16));
		} else {
#pragma unroll
			for (int k = 0; k < NVC; k++) {
				const uint4 c = *(const uint4 *) (covt + ((k ^ swc) << 2));
				ctot += (int) ((c.x & FIELD) + (c.y & FIELD) + (c.z & FIELD) + (c.w & FIELD))
				      - (int) ((c.x >> END_SHIFT) + (c.y >> END_SHIFT) + (c.z >> END_SHIFT) + (c.w >> END_SHIFT));
			}
		}
#pragma unroll
		for (int k = 0; k < NVL; k++) {
			const int pv = p0 + ((k ^ swl) << 2);
			const uint4 l = *(const uint4 *) (limb + pv);
			if (LIMBS == 1) atot += (acc_t) (l.x + l.y + l.z + l.w);
			else {
				const uint4 h = *(const uint4 *) (limb + W + pv);
				atot += (acc_t) ((long long) ((unsigned long long) l.x + l.y + l.z + l.w)
				      + (((long long) (int32_t) h.x + (int32_t) h.y + (int32_t) h.z + (int32_t) h.w) << lw));
			}
			if (MOM) {
				const uint4 ql = *(const uint4 *) (sq + pv), qh = *(const uint4 *) (sq + W + pv);
				btot += (long long) ((unsigned long long) ql.x + ql.y + ql.z + ql.w)
				      + (((long long) (int32_t) qh.x + (int32_t) qh.y + (int32_t) qh.z + (int32_t) qh.w) << lw);
			}
		}
	}
	// exclusive block scan of (cover, value)
	int ec; acc_t ea; long long eb = 0;
	{
		int ic = ctot; acc_t ia = atot; long long ib = btot;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int vc = __shfl_up_sync(FULL, ic, d);
			const acc_t va = __shfl_up_sync(FULL, ia, d);
			if (lane >= d) { ic += vc; ia += va; }
			if (MOM) { const long long vb = __shfl_up_sync(FULL, ib, d); if (lane >= d) ib += vb; }
		}
		if (lane == 31) { misc->wcov[warp] = ic; misc->wsum[warp] = (long long) ia; if (MOM) misc->wsum2[warp] = ib; }
		__syncthreads();
		int bc = 0; acc_t ba = 0; long long bb = 0;
#pragma unroll
		for (int k = 0; k < NWARP; k++) if (k < warp) { bc += misc->wcov[k]; ba += (acc_t) misc->wsum[k]; if (MOM) bb += misc->wsum2[k]; }
		ec = bc + ic - ctot + misc->base_cov;
		ea = ba + ia - atot + (acc_t) (long long) misc->base_val;
		if (MOM) eb = bb + ib - btot + (long long) misc->base_val2;
	}
	WB_PHASE(2);

	// ---- pass 2: running state after every position; the exact prefix goes back in place -----
	unsigned evbits = 0, keepbits = 0;
	{
		int c = ec; acc_t a = ea; long long b = eb;
		uint16_t *cnt = (uint16_t *) (smem + L.off_cnt);
		uint4 cv = make_uint4(0, 0, 0, 0);
#pragma unroll
		for (int k = 0; k < NVL; k++) {
			const int pv = p0 + ((k ^ swl) << 2);
			if (COV16) { if ((k & 1) == 0) cv = *(const uint4 *) (covt + (((k >> 1) ^ swc) << 2)); }
			else cv = *(const uint4 *) (covt + ((k ^ swc) << 2));
			int st[4], en[4];
			cover4<COV16>(cv, k & 1, st, en);
			uint4 l = *(const uint4 *) (limb + pv);
			uint32_t lo[4] = {l.x, l.y, l.z, l.w}, hi[4] = {0, 0, 0, 0};
			if (LIMBS == 2) { const uint4 h = *(const uint4 *) (limb + W + pv); hi[0] = h.x; hi[1] = h.y; hi[2] = h.z; hi[3] = h.w; }
			uint32_t ql[4] = {0, 0, 0, 0}, qh[4] = {0, 0, 0, 0};
			if (MOM) {
				const uint4 x = *(const uint4 *) (sq + pv), y = *(const uint4 *) (sq + W + pv);
				ql[0] = x.x; ql[1] = x.y; ql[2] = x.z; ql[3] = x.w; qh[0] = y.x; qh[1] = y.y; qh[2] = y.z; qh[3] = y.w;
			}
#pragma unroll
			for (int j = 0; j < 4; j++) {
				c += st[j] - en[j];
				if (LIMBS == 1) a += (acc_t) lo[j];
				else a += (acc_t) ((long long) (unsigned long long) lo[j] + ((long long) (int32_t) hi[j] << lw));
				if (MOM) b += (long long) (unsigned long long) ql[j] + ((long long) (int32_t) qh[j] << lw);
				if (st[j] | en[j]) {
					evbits |= 1u << (4 * k + j);
					if (c >= thr) keepbits |= 1u << (4 * k + j);
					if (MOM) cnt[pv + j] = (uint16_t) c;                   // the statistic needs the tracks in play
				}
				lo[j] = (uint32_t) a;                                  // LIMBS == 1: the sum itself fits 32 bits
				if (LIMBS == 2) hi[j] = (uint32_t) ((unsigned long long) a >> 32);
				if (MOM) { ql[j] = (uint32_t) b; qh[j] = (uint32_t) ((unsigned long long) b >> 32); }
			}
			*(uint4 *) (limb + pv) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
			if (LIMBS == 2) *(uint4 *) (limb + W + pv) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
			if (MOM) {
				*(uint4 *) (sq + pv) = make_uint4(ql[0], ql[1], ql[2], ql[3]);
				*(uint4 *) (sq + W + pv) = make_uint4(qh[0], qh[1], qh[2], qh[3]);
			}
		}
	}
	// rare: segments covering an interval whose value is not representable (fix-up kernel)
	unsigned slowbits = 0;
	{
		const int np = misc->npois, bp = misc->base_pois;
		if (np | bp) {
			if (np > POIS_CAP) slowbits = ~0u;
			else {
				int run = bp;
				for (int k = 0; k < np; k++) if ((int) misc->pois[k].pos < p0) run += misc->pois[k].delta;
				for (int q = 0; q < V; q++) {
					for (int k = 0; k < np; k++) if ((int) misc->pois[k].pos == p0 + q) run += misc->pois[k].delta;
					if (run > 0) slowbits |= 1u << q;
				}
			}
		}
	}

	// breakpoints and kept segments of this thread; one packed block scan for both
	int rank, kidx, nbp, kept_total;
	{
		const int mine = __popc(evbits) | (__popc(keepbits) << 16);
		int ip = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, ip, d); if (lane >= d) ip += v; }
		if (lane == 31) misc->wpack[warp] = ip;
		__syncthreads();                                       // every thread is past its reads of cov
		int bp = 0, tp = 0;
#pragma unroll
		for (int k = 0; k < NWARP; k++) { const int kp = misc->wpack[k]; if (k < warp) bp += kp; tp += kp; }
		const int ep = bp + ip - mine;
		rank = ep & 0xffff; kidx = ep >> 16;
		nbp = tp & 0xffff; kept_total = tp >> 16;
	}
	// publish this window's segment count so that later windows can start their look-back
	if (tid == 0) st_volatile(&job.state[w], LB_AGG | (unsigned long long) kept_total);
	// Windows are handed out in order, so window w + grid starts about one window time after
	// this one did: its cursor row is fetched now and its intervals are pulled into L2 (below)
	// for whichever CTA draws it.
	const int w1 = w + (int) gridDim.x;
	const bool pref = WB_FS_PREF && w1 < job.nw;
	if (pref && tid == 0) {
		const uint32_t bytes = (uint32_t) job.cstride * 4;
		mbar_expect_tx(&misc->bar_nxt, bytes);
		bulk_g2s(nxt_buf, job.cursor + (size_t) w1 * job.cstride, bytes, &misc->bar_nxt);
	}
	WB_PHASE(3);

	// ---- 3. dense breakpoint list over the dead cover array ---------------------------------
	uint16_t *dense = (uint16_t *) cov;
	{
		unsigned bits = evbits;
		while (bits) {
			const int q = __ffs(bits) - 1;
			bits &= bits - 1;
			if ((rank & 31) == 0) misc->kbase[rank >> 5] = (uint16_t) kidx;
			unsigned d = (unsigned) (p0 + q);
			if ((keepbits >> q) & 1) {
				d |= DN_KEEP;
				kidx++;
				if ((slowbits >> q) & 1) d |= DN_SLOW;
			}
			dense[rank++] = (uint16_t) d;
		}
	}
	__syncthreads();
	WB_PHASE(4);

	if (warp != 0) {
		// ---- the next window's intervals: L2 prefetch, 8 lanes per track (16 bytes of start and end,
		// 32 of value each), while warp 0 runs the look-back.  (cp.async.bulk.prefetch.L2 was measured
		// here: the bulk instruction is issued once per LANE -- a loop over the warp's active lanes --
		// and cost 10 % of the kernel's instructions; profiles/README.md.)
		if (pref) {
			mbar_wait(&misc->bar_nxt, par_nxt);
			const long long *toff = (const long long *) (smem + L.off_toff);
			const int gid = (tid - 32) >> 3, gl = tid & 7;
			for (int i = gid; i < N; i += (TH - 32) / 8) {
				const long long o = toff[i], first = o + nxt_buf[i];
				const long long g = (first & ~3ll) + 4 * gl;
				if (g < toff[i + 1] && g + 4 <= job.I) {
					asm volatile("prefetch.global.L2 [%0];" :: "l"(job.start + g)); asm volatile("prefetch.global.L2 [%0];" :: "l"(job.end + g));
					asm volatile("prefetch.global.L2 [%0];" :: "l"(job.value + g)); asm volatile("prefetch.global.L2 [%0];" :: "l"(job.value + g + 2));
				}
			}
		}
	} else {
		// ---- output offset: decoupled look-back over the preceding windows --------------------
		// LBK x 32 predecessors per step (the loads of one step are independent: one L2 round
		// trip), summed back to the nearest window that already knows its inclusive prefix
		unsigned long long sum = 0;
		int p = w - 1;
		bool found = false;
#ifdef WB_DIAG_PHASES
		unsigned long long diag_polls = 0, diag_steps = 0;
#endif
		while (!found) {
			unsigned long long st[LBK];
			int use = LBK;                                         // groups of 32 states summed in this step
			unsigned ns = WB_FS_LBSLEEP;
			while (true) {
#pragma unroll
				for (int j = 0; j < LBK; j++) {
					const int idx = p - 32 * j - lane;
					st[j] = idx >= 0 ? ld_volatile(&job.state[idx]) : LB_PREFIX;      // before window 0: prefix 0
				}
#ifdef WB_DIAG_PHASES
				diag_polls++;
#endif
				// only the states between this window and the nearest known prefix are summed: the
				// ones behind that prefix need not have been published yet
				bool wait = false;
				use = LBK;
#pragma unroll
				for (int j = 0; j < LBK; j++) {
					if (j >= use || wait) break;
					const unsigned nr = __ballot_sync(FULL, (st[j] >> 62) == 0), pm = __ballot_sync(FULL, (st[j] >> 62) == 2);
					if (pm != 0 && (nr == 0 || (pm & -pm) < (nr & -nr))) use = j + 1;
					else if (nr != 0) wait = true;
				}
				if (!wait) break;
				__nanosleep(ns);
				if (ns < 8 * WB_FS_LBSLEEP) ns *= 2;
			}
#ifdef WB_DIAG_PHASES
			diag_steps++;
#endif
#pragma unroll
			for (int j = 0; j < LBK; j++) {
				if (j >= use) break;
				const unsigned pm = __ballot_sync(FULL, (st[j] >> 62) == 2);
				const int first = pm ? __ffs(pm) - 1 : 32;
				unsigned long long v = (lane <= first) ? (st[j] & LB_VALUE) : 0;
				for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(FULL, v, d);
				sum += __shfl_sync(FULL, v, 0);
				found = pm != 0;
			}
			p -= 32 * LBK;
		}
#ifdef WB_DIAG_PHASES
		if (lane == 0) { atomicAdd(&job.diag[9], diag_polls); atomicAdd(&job.diag[10], diag_steps); }
#endif
		if (lane == 0) {
			misc->out_base = (long long) sum;
			st_volatile(&job.state[w], LB_PREFIX | (sum + (unsigned long long) kept_total));
			if (w == job.nw - 1) *job.total = (long long) (sum + (unsigned long long) kept_total);
		}
	}
	__syncthreads();
	WB_PHASE(5);
#if WB_FS_TICKET_AT_END == 0
	// (earlier scheme) the next window is drawn now, and its cursor rows land while this window is emitted
	if (tid == 0) {
		const int nw1 = (int) atomicAdd(job.ticket, 1u);
		misc->ticket2 = nw1;
		if (nw1 < job.nw) {
			const uint32_t bytes = (uint32_t) job.cstride * 8;       // rows nw1 and nw1 + 1
			mbar_expect_tx(&misc->bar_cur, bytes);
			bulk_g2s(cur, job.cursor + (size_t) nw1 * job.cstride, bytes, &misc->bar_cur);
		}
	}
#endif

	// ---- 4. emit, one breakpoint per thread (coalesced stores) -------------------------------
	const long long out_base = misc->out_base;
	const int next_bp = misc->next_bp;
	int32_t *os = job.out_start + out_base, *oe = job.out_end + out_base;
	double *ov = job.out_value + out_base;
	const long long room = job.out_cap - out_base;             // output ranks below this are in range
	const bool npow2 = (N & (N - 1)) == 0;
	const bool mean = job.op == OP_MEAN;
	// sum: x 2^q; mean over a power-of-two N: x 2^q / N is one exact power of two as well
	const double mul = (mean && npow2) ? scale * (1.0 / N) : scale;
	const bool div = mean && !npow2;
	const int32_t fx_base32 = (int32_t) job.fx_base;
	const int mom_op = job.op == OP_ENTROPY ? OP_STDDEV : job.op;
	const uint16_t *cnt16 = (const uint16_t *) (smem + L.off_cnt);
	double auc = 0;
	// the three output pointers of this window, opaque to the optimiser: one 32-bit
	// index per store instead of a re-derived 64-bit (out_base + rank) chain
	asm volatile("" : "+l"(os), "+l"(oe), "+l"(ov));
	const int room32 = (int) min(room, (long long) INT_MAX);
	const unsigned lt = (1u << lane) - 1;
	auto emit_loop = [&](auto div_tag) {
		constexpr bool DIV = decltype(div_tag)::value;      // mean over a non-power-of-two N: one fp64 division per segment
		// value of the segment that starts at position `pos`: correctly rounded exact sum
		auto value_at = [&](int pos) -> double {
			const int ph = phys32<VLOG>(pos);
			double v;
			if (MOM) {
				// var / stddev / CV from the exact sums (engine.cu: emit_kernel<MODE_MOMENTS> evaluates the same expression)
				const long long sxx = (long long) (((unsigned long long) sq[W + ph] << 32) | sq[ph]);
				return moments_stat<long long>(mom_op, N, (int) cnt16[ph], (long long) (int32_t) limb[ph], sxx, scale);   // one-limb sums: 64 bits hold the products
			}
			if (LIMBS == 1) v = (double) ((int32_t) limb[ph] + fx_base32) * mul;     // |sum + defaults| < 2^30: one 32-bit conversion
			else v = (double) ((long long) (((unsigned long long) limb[W + ph] << 32) | limb[ph]) + job.fx_base) * mul;
			if (DIV) v = v / N;
			return v;
		};
		auto store = [&](unsigned d, int kid, int bs, int be, double v) {
			if (kid < room32) {
				os[kid] = bs;
				oe[kid] = be;
				ov[kid] = v;
				if (d & DN_SLOW) {
					// an input was not representable: the fix-up kernel recomputes this
					// segment with the reference's own operation order
					const unsigned long long k = atomicAdd(job.slow_count, 1ull);
					if (k < (unsigned long long) job.slow_cap) job.slow_list[k] = out_base + kid;
				}
			}
			if (!(d & DN_SLOW) && !(MOM && isnan(v))) auc += (double) (be - bs) * v;   // sum / mean: v is finite here
		};
#if WB_FS_EMIT2
		// two groups of 32 dense entries per warp and step: the shared-memory reads of both are in flight together
		for (int r0 = warp * 32; r0 < nbp; r0 += 2 * TH) {
			const int ra = r0 + lane, rb = ra + TH;
			const unsigned da = ra < nbp ? dense[ra] : 0u, db = rb < nbp ? dense[rb] : 0u;
			const bool ka = (da & DN_KEEP) != 0, kb = (db & DN_KEEP) != 0;
			const unsigned na = (ka && ra + 1 < nbp) ? dense[ra + 1] : 0u, nb = (kb && rb + 1 < nbp) ? dense[rb + 1] : 0u;
			const unsigned kma = __ballot_sync(FULL, ka), kmb = __ballot_sync(FULL, kb);
			const int kida = (int) misc->kbase[r0 >> 5] + __popc(kma & lt);
			const int kidb = (r0 + TH < nbp ? (int) misc->kbase[(r0 + TH) >> 5] : 0) + __popc(kmb & lt);
			const int posa = (int) (da & DN_POS), posb = (int) (db & DN_POS);
			double va = 0, vb = 0;
			if (ka) va = value_at(posa);
			if (kb) vb = value_at(posb);
			if (ka) store(da, kida, w0i + posa, (ra + 1 < nbp) ? w0i + (int) (na & DN_POS) : next_bp, va);
			if (kb) store(db, kidb, w0i + posb, (rb + 1 < nbp) ? w0i + (int) (nb & DN_POS) : next_bp, vb);
		}
#else
		for (int r0 = warp * 32; r0 < nbp; r0 += TH) {
			const int r = r0 + lane;
			const unsigned d = r < nbp ? dense[r] : 0u;
			const bool keep = (d & DN_KEEP) != 0;
			const unsigned km = __ballot_sync(FULL, keep);
			if (!keep) continue;
			const int kid = (int) misc->kbase[r0 >> 5] + __popc(km & lt);
			const int pos = (int) (d & DN_POS);
			const int be = (r + 1 < nbp) ? w0i + (int) (dense[r + 1] & DN_POS) : next_bp;
			store(d, kid, w0i + pos, be, value_at(pos));
		}
#endif
	};
	if (div) emit_loop(std::true_type{}); else emit_loop(std::false_type{});
	for (int d = 16; d > 0; d >>= 1) auc += __shfl_down_sync(FULL, auc, d);
	if (lane == 0) misc->wauc[warp] = auc;
	WB_PHASE(6);
}

template <int VLOG, int LIMBS, bool COV16, bool MOM>
__global__ void __launch_bounds__(TH, MOM ? (COV16 && VLOG == 4 ? WB_FS_MOM_MINB : 4) : (LIMBS == 1 && COV16) ? WB_FS_MINB : LIMBS == 1 ? 6 : COV16 ? 5 : 4)
fastsum_kernel(const Job job, int lw, double inv_scale, double lim, double scale) {
	extern __shared__ __align__(16) unsigned char smem[];
	const Layout L = make_layout(job.wlog, job.N, LIMBS, COV16, MOM);
	long long *toff = (long long *) (smem + L.off_toff);
	long long *tmd = (long long *) (smem + L.off_tmd);
	for (int i = threadIdx.x; i <= job.N; i += TH) {
		toff[i] = job.off[i];
		if (i < job.N) tmd[i] = MOM ? 0 : job.dflt_fx[i];          // moments: all defaults are 0 (capi.cu: wants_format)
	}
	Misc *misc = (Misc *) smem;
	int32_t *ca = (int32_t *) (smem + L.off_cur), *cb = (int32_t *) (smem + L.off_cur2);   // cursor rows: window being processed / prefetched for
	// Windows are handed out in ascending order by a ticket counter, the first one included: a CTA
	// that holds window w is running, and so is (or was) the holder of every window below w.
	if (threadIdx.x == 0) {
		mbar_init(&misc->bar_cur, 1);
		mbar_init(&misc->bar_nxt, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		misc->ticket2 = (int) atomicAdd(job.ticket, 1u);
	}
	__syncthreads();
	int w = misc->ticket2;
	if (threadIdx.x == 0 && w < job.nw) {
		const uint32_t bytes = (uint32_t) job.cstride * 8;           // rows w and w + 1
		mbar_expect_tx(&misc->bar_cur, bytes);
		bulk_g2s(ca, job.cursor + (size_t) w * job.cstride, bytes, &misc->bar_cur);
	}
	unsigned par_cur = 0, par_nxt = 0;
	while (w < job.nw) {
#ifdef WB_DIAG_PHASES
		const long long diag_t0 = clock64();
#endif
		const bool pref = WB_FS_PREF && w + (int) gridDim.x < job.nw;
		process_window<VLOG, LIMBS, COV16, MOM>(job, w, smem, L, ca, cb, par_cur, par_nxt, lw, inv_scale, lim, scale);
		par_cur ^= 1;
		if (pref) par_nxt ^= 1;
		__syncthreads();      // dense list consumed, ticket and AUC partials visible
		if (threadIdx.x == 0 && job.auc_partial) {
			double t = 0;
			for (int k = 0; k < NWARP; k++) t += misc->wauc[k];
			job.auc_partial[w] = t;
		}
#ifdef WB_DIAG_PHASES
		if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&job.diag[7], (unsigned long long) (t_ - diag_t0)); atomicAdd(&job.diag[8], 1ull); }
#endif
#if WB_FS_TICKET_AT_END
		// The look-back of window w waits for every window below w, so the less time passes between drawing a
		// window and starting it, the less the windows in flight run out of order: the next window is drawn only
		// now, when this one is completely done (its cursor rows arrive while the window arrays are cleared).
		// Measured against drawing at the start of the emit phase: look-back 10.7 k -> 9.0 k cycles per window, -2.7 %.
		if (threadIdx.x == 0) {
			const int nw1 = (int) atomicAdd(job.ticket, 1u);
			misc->ticket2 = nw1;
			if (nw1 < job.nw) {
				const uint32_t bytes = (uint32_t) job.cstride * 8;
				mbar_expect_tx(&misc->bar_cur, bytes);
				bulk_g2s(ca, job.cursor + (size_t) nw1 * job.cstride, bytes, &misc->bar_cur);
			}
		}
		__syncthreads();
#endif
		w = misc->ticket2;
	}
}

template <int VLOG, int LIMBS, bool COV16, bool MOM>
static cudaError_t launch_v(const Job &job, int lw, double inv_scale, double lim, double scale, cudaStream_t s) {
	size_t smem = make_layout(job.wlog, job.N, LIMBS, COV16, MOM).total;
	{ const char *e = getenv("WB200_FS_PAD"); if (e) smem += (size_t) atoi(e) * 1024; }   // development: cap the CTAs per SM
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaError_t e = cudaFuncSetAttribute(fastsum_kernel<VLOG, LIMBS, COV16, MOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (e != cudaSuccess) return e;
	int per = 1;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fastsum_kernel<VLOG, LIMBS, COV16, MOM>, TH, smem);
	if (getenv("WB200_FS_VERBOSE")) fprintf(stderr, "fastsum_kernel<%d,%d,%d,%d>: %zu B smem, %d CTAs/SM\n", VLOG, LIMBS, (int) COV16, (int) MOM, smem, per);
	const int grid = min(job.nw, sms * max(per, 1));
	fastsum_kernel<VLOG, LIMBS, COV16, MOM><<<grid, TH, smem, s>>>(job, lw, inv_scale, lim, scale);
	return cudaGetLastError();
}

static bool use_cov16(const Job &job) {
	const char *e = getenv("WB200_FS_COV16");           // development override of the compiled default
	return job.N <= 255 && (e ? atoi(e) != 0 : WB_FS_COV16 != 0);
}

} // namespace fs

// Whether the sum / mean sweep kernel covers this job (else: emit_kernel<MODE_FAST>).
bool fastsum_supported(const Job &job) {
	if (!job.vec || job.N > MAX_TRACKS) return false;
	const int vlog = job.wlog - 7;
	if (vlog != 4 && vlog != 5) return false;
	if (job.fast == 2) {
		// var / stddev / CV: the plain sum in one limb, the squares in two (else emit_kernel<MODE_MOMENTS>)
		int l2n = 0;
		while ((1ll << l2n) < 2ll * job.N) l2n++;
		const int lw = 32 - l2n;
		const char *e = getenv("WB200_FS_MOMENTS");      // development: 0 keeps them on emit_kernel<MODE_MOMENTS>
		if ((e && atoi(e) == 0) || !job.sum1 || 2 * job.fx_mbits > 30 - l2n + lw || job.N > 65535) return false;
		return fs::make_layout(job.wlog, job.N, 1, fs::use_cov16(job), true).total <= 200 * 1024;
	}
	if (job.fast != 1 || job.fx_limbs > 2) return false;
	return fs::make_layout(job.wlog, job.N, job.fx_limbs, fs::use_cov16(job), false).total <= 200 * 1024;
}

cudaError_t launch_fastsum(const Job &job, int lw, double inv_scale, double lim, double scale, cudaStream_t s) {
	const int vlog = job.wlog - 7;
	const bool mom = job.fast == 2, one = job.fx_limbs == 1, c16 = fs::use_cov16(job);
#define WB_FS(V) (mom ? (c16 ? fs::launch_v<V, 1, true, true>(job, lw, inv_scale, lim, scale, s) : fs::launch_v<V, 1, false, true>(job, lw, inv_scale, lim, scale, s)) \
                : one ? (c16 ? fs::launch_v<V, 1, true, false>(job, lw, inv_scale, lim, scale, s) : fs::launch_v<V, 1, false, false>(job, lw, inv_scale, lim, scale, s)) \
                      : (c16 ? fs::launch_v<V, 2, true, false>(job, lw, inv_scale, lim, scale, s) : fs::launch_v<V, 2, false, false>(job, lw, inv_scale, lim, scale, s)))
	if (vlog == 4) return WB_FS(4);
	if (vlog == 5) return WB_FS(5);
#undef WB_FS
	return cudaErrorInvalidValue;
}

} // namespace wb
