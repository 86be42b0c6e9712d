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
//   * 128 threads per window (32 positions per thread) -> 6 CTAs per SM instead
//     of 3 at the same register budget;
//   * 8 or 12 bytes of shared memory per position instead of 13:
//       cover word  : starts in bits 0..10, ends in bits 11..21 of one int32 -- a
//                     position is a breakpoint iff the word is non-zero, so the
//                     separate flag array (and its stores) is gone;
//       value       : one or two 32-bit limbs (host-chosen from the value range);
//       poison      : NaN / non-representable values are rare and go to a short
//                     side list instead of a field in every word;
//   * a scan that never holds the window in registers: pass 1 sums each thread's
//     32 positions, one block scan stitches the threads, pass 2 re-reads them and
//     writes the exact prefix back IN PLACE, so the dense breakpoint list is one
//     32-bit word per breakpoint (position | output rank | flags) over the dead
//     cover array and the emit phase fetches the value by position.
// Output is bit-identical to emit_kernel<MODE_FAST>; anything this kernel does
// not cover (three limbs, unaligned inputs, other window sizes) stays there.
#include "engine.cuh"
#include "device_util.cuh"
#include <cuda_pipeline.h>
#include <climits>
#include <type_traits>
#include <cstdlib>

namespace wb {
namespace fs {

constexpr int MAXWARP = 8;               // CTAs of 128 or 256 threads (one window per CTA at a time)
#ifndef WB_FS_MINB
#define WB_FS_MINB 4                  // resident CTAs per SM the 256-thread variant is compiled for
#endif
constexpr int END_SHIFT = 11;            // cover word: starts | ends << 11 (each <= 1023 per position)
constexpr uint32_t FIELD = (1u << END_SHIFT) - 1;
#ifndef WB_FS_LBK
#define WB_FS_LBK 1
#endif
#ifndef WB_FS_LBRELAX
#define WB_FS_LBRELAX 1               // look-back waits only for the predecessors it sums
#endif
#ifndef WB_FS_LATETICKET
#define WB_FS_LATETICKET 2            // next window drawn: 0 two windows ahead, 1 after the events phase, 2 when the emit phase starts
#endif
#ifndef WB_FS_EVPIPE
#define WB_FS_EVPIPE 1                // events: software-pipelined interval loads
#endif
#ifndef WB_FS_LBSLEEP
#define WB_FS_LBSLEEP 40
#endif
#ifndef WB_FS_PREF
#define WB_FS_PREF 3                  // next window's intervals: 0 no prefetch, 2 prefetch.L1, 3 prefetch.L2
#endif
constexpr int LBK = WB_FS_LBK;           // look-back: 32 * LBK predecessors per step
constexpr int POIS_CAP = 120;            // poison events kept per window; more -> the whole window takes the fix-up path
constexpr uint32_t D_KEEP = 1u << 26, D_SLOW = 1u << 27;
constexpr int D_KID_SHIFT = 13;          // dense word: position (13 bits) | output rank (13 bits) | flags

struct PoisEv { uint16_t pos; int16_t delta; };

struct Misc {
	int next_bp;                     // first breakpoint at or after the window end
	int base_cov;                    // tracks in play at the window start
	int base_pois;                   // of those, tracks whose value is not representable
	int npois;                       // poison events inside the window
	unsigned long long base_val;     // fixed-point sum carried into the window
	long long out_base;              // output offset of the window
	int ticket2;
	int pad;
	long long wsum[MAXWARP];
	int wcov[MAXWARP];
	int wpack[MAXWARP];
	double wauc[MAXWARP];
	PoisEv pois[POIS_CAP];
};
constexpr int MISC_BYTES = 1024;
static_assert(sizeof(Misc) <= MISC_BYTES, "Misc must fit its slot");

struct Layout { size_t off_cur, off_cur2, off_tn, off_toff, off_tmd, off_cov, off_limb, total; };
__host__ __device__ inline Layout make_layout(int wlog, int N, int limbs) {
	Layout L;
	const size_t W = (size_t) 1 << wlog, nn = (size_t) (N + 4) / 4 * 16;
	size_t o = MISC_BYTES;
	L.off_cur = o;  o += nn;
	L.off_cur2 = o; o += nn;
	L.off_tn = o;   o += nn;
	L.off_toff = o; o += 2 * nn;
	L.off_tmd = o;  o += 2 * nn;
	L.off_cov = o;  o += 4 * W;
	L.off_limb = o; o += (size_t) limbs * 4 * W;
	L.total = (o + 15) / 16 * 16;
	return L;
}

// Thread t owns the V = 2^VLOG consecutive positions [t*V, t*V+V) and reads them back
// with 128-bit loads; its V/4 vectors are rotated by a function of t so that the 8
// lanes of a quarter warp hit 8 different 16-byte bank groups.
template <int VLOG>
__device__ __forceinline__ int swz(int t) {
	constexpr int NV = 1 << (VLOG - 2);
	return NV >= 8 ? (t & (NV - 1)) : ((t >> (5 - VLOG)) & (NV - 1));
}
template <int VLOG>
__device__ __forceinline__ int phys(int p) {
	const int t = p >> VLOG, o = p & ((1 << VLOG) - 1);
	return (t << VLOG) | (((o >> 2) ^ swz<VLOG>(t)) << 2) | (o & 3);
}

struct Chunk {
	int4 s, e;
	double v0, v1, v2, v3;
	int lo, hi;          // slots lo <= x < hi hold intervals of this track at or after its cursor
};

// ---- events: +1 start / +1 end per interval in the cover word, +-value in the limbs ----
template <int TH, int VLOG, int LIMBS>
__device__ __forceinline__ void events(const Job &job, unsigned char *smem, const Layout &L, const int32_t *cur,
                                       long long w0, int lw, double inv_scale, double lim) {
	constexpr int W = TH << VLOG, G = TH / 8;
	Misc *misc = (Misc *) smem;
	const int32_t *tn = (const int32_t *) (smem + L.off_tn);
	const long long *toff = (const long long *) (smem + L.off_toff);
	const long long *tmd = (const long long *) (smem + L.off_tmd);
	uint32_t *cov = (uint32_t *) (smem + L.off_cov);
	uint32_t *limb = (uint32_t *) (smem + L.off_limb);
	const int tid = threadIdx.x, lane = tid & 31, gid = tid >> 3, gl = tid & 7, gsh = lane & 24;
	const int N = job.N;
	const int w0i = (int) w0;
	using acc_t = typename std::conditional<LIMBS == 1, int, long long>::type;
	int in_cov = 0, in_pois = 0, my_next = INT_MAX;
	acc_t in_val = 0;

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
	auto poison = [&](unsigned x, int delta) {
		const int k = atomicAdd(&misc->npois, 1);
		if (k < POIS_CAP) { PoisEv p; p.pos = (uint16_t) x; p.delta = (int16_t) delta; misc->pois[k] = p; }
	};
	auto one = [&](bool valid, int s, int e, double val, long long md) {
		if (!valid) return;
		const unsigned xs = (unsigned) (s - w0i), xe = (unsigned) (e - w0i);
		const bool start_in = xs < (unsigned) W, end_in = xe < (unsigned) W;
		if (!start_in && s >= w0i) { my_next = min(my_next, s); return; }   // begins at or beyond the window end
		acc_t d = 0;
		bool bad;
		if (LIMBS == 1) {
			// one limb: |m| and the defaults are below 2^29 / 2N (capi.cu), so the whole
			// delta arithmetic is 32-bit; same acceptance rule as to_fixed()
			const double t = val * inv_scale;
			const int mi = __double2int_rn(t);
			bad = !((double) mi == t && fabs(t) < lim && (mi != 0 || val == 0.0));
			if (!bad) d = (acc_t) (mi - (int) md);
		} else {
			long long m;
			bad = !to_fixed(val, inv_scale, lim, m);
			if (!bad) d = (acc_t) (m - md);
		}
		if (start_in) {
			const int ph = phys<VLOG>((int) xs);
			atomicAdd(&cov[ph], 1u);
			if (d != 0) add_limbs_fs<LIMBS, W>(limb, ph, d, lw);
			if (bad) poison(xs, 1);
		} else { in_cov++; in_val += d; in_pois += bad ? 1 : 0; }           // began before the window
		if (end_in) {
			const int ph = phys<VLOG>((int) xe);
			atomicAdd(&cov[ph], 1u << END_SHIFT);
			if (d != 0) add_limbs_fs<LIMBS, W>(limb, ph, -d, lw);
			if (bad) poison(xe, -1);
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
#if WB_FS_EVPIPE
	// the first chunk of the next round of tracks is requested before this round's events are
	// applied: one exposed load latency per window instead of one per round
	Chunk nx;
	load(gid, 0, nx);
	for (int rb = 0; rb < N; rb += G) {
		const int i = rb + gid;
		const long long md = i < N ? tmd[i] : 0;
		const Chunk c0 = nx;
		load(i + G, 0, nx);
		bool more = apply(c0, md);
		for (int k = 1; __any_sync(FULL, more); k++) {             // rare: more than ~30 intervals of one track in the window
			Chunk c;
			load(more ? i : N, k, c);
			more = apply(c, md) && more;
		}
	}
#else
	for (int rb = 0; rb < N; rb += G) {
		const int i = rb + gid;
		const long long md = i < N ? tmd[i] : 0;
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
	in_pois = __reduce_add_sync(FULL, in_pois);
	if (LIMBS == 1) in_val = (acc_t) __reduce_add_sync(FULL, (int) in_val);
	else for (int d = 16; d > 0; d >>= 1) in_val += __shfl_down_sync(FULL, in_val, d);
	if (lane == 0) {
		if (my_next != INT_MAX) atomicMin(&misc->next_bp, my_next);
		if (in_cov) atomicAdd(&misc->base_cov, in_cov);
		if (in_pois) atomicAdd(&misc->base_pois, in_pois);
		if (in_val) atomicAdd(&misc->base_val, (unsigned long long) (long long) in_val);
	}
}

// ---- one window ------------------------------------------------------------------------
template <int TH, int VLOG, int LIMBS>
__device__ void process_window(const Job &job, int w, unsigned char *smem, const Layout &L, const int32_t *cur,
                               int32_t *nxt_buf, int w1_early, int lw, double inv_scale, double lim, double scale) {
	constexpr int V = 1 << VLOG, W = TH << VLOG, NV = V / 4, NWARP = TH / 32;
	static_assert(V <= 32, "one 32-bit mask per thread");
	Misc *misc = (Misc *) smem;
	uint32_t *cov = (uint32_t *) (smem + L.off_cov);
	uint32_t *limb = (uint32_t *) (smem + L.off_limb);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = job.N;
	const long long w0 = job.p0 + ((long long) w << job.wlog);
	const int thr = job.strict ? N : 1;
#ifdef WB_DIAG_PHASES
	long long diag_t = clock64();
#endif

	// ---- clear ------------------------------------------------------------------------
	{
		const uint4 z = make_uint4(0, 0, 0, 0);
		uint4 *zc = (uint4 *) cov;                                   // cov and limbs are contiguous
		constexpr int n16 = (1 + LIMBS) * W / 4;
#pragma unroll 4
		for (int t = tid; t < n16; t += TH) zc[t] = z;
		if (tid == 0) { misc->next_bp = INT_MAX; misc->base_cov = 0; misc->base_pois = 0; misc->npois = 0; misc->base_val = 0; }
	}
#if WB_FS_LATETICKET == 2
	__pipeline_wait_prior(0);                                  // this window's cursor row, requested before the clear
#endif
	__syncthreads();
	WB_PHASE(0);

	// ---- 1. events ---------------------------------------------------------------------
#ifndef WB_DIAG_SKIP_EVENTS
	events<TH, VLOG, LIMBS>(job, smem, L, cur, w0, lw, inv_scale, lim);
#endif
	__syncthreads();
	WB_PHASE(1);
	// The look-back of window w waits for every window below w, so the less time passes between
	// drawing a window and starting it, the less the windows in flight run out of order.
#if WB_FS_LATETICKET == 1
	// as late as the draw's own cursor row and interval prefetch allow
	unsigned next_ticket = 0;
	if (tid == 0) next_ticket = atomicAdd(job.ticket, 1u) + gridDim.x;
#endif

	// ---- 2. scan, pass 1: totals of this thread's V positions -------------------------------
	const int p0 = tid << VLOG, sw = swz<VLOG>(tid);
	using acc_t = typename std::conditional<LIMBS == 1, int, long long>::type;   // one limb: every prefix fits 32 bits
	unsigned evbits = 0;
	int ctot = 0;
	acc_t atot = 0;
#pragma unroll
	for (int k = 0; k < NV; k++) {
		const int pv = p0 + ((k ^ sw) << 2);
		const uint4 c = *(const uint4 *) (cov + pv);
		const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
		for (int j = 0; j < 4; j++) {
			ctot += (int) (cw[j] & FIELD) - (int) (cw[j] >> END_SHIFT);
			if (cw[j]) evbits |= 1u << (4 * k + j);
		}
		const uint4 l = *(const uint4 *) (limb + pv);
		if (LIMBS == 1) atot += (acc_t) (l.x + l.y + l.z + l.w);
		else {
			const uint4 h = *(const uint4 *) (limb + W + pv);
			atot += (acc_t) ((long long) ((unsigned long long) l.x + l.y + l.z + l.w)
			      + (((long long) (int32_t) h.x + (int32_t) h.y + (int32_t) h.z + (int32_t) h.w) << lw));
		}
	}
	// exclusive block scan of (cover, value)
	int ec; acc_t ea;
	{
		int ic = ctot; acc_t ia = atot;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int vc = __shfl_up_sync(FULL, ic, d);
			const acc_t va = __shfl_up_sync(FULL, ia, d);
			if (lane >= d) { ic += vc; ia += va; }
		}
		if (lane == 31) { misc->wcov[warp] = ic; misc->wsum[warp] = (long long) ia; }
		__syncthreads();
		int bc = 0; acc_t ba = 0;
#pragma unroll
		for (int k = 0; k < NWARP; k++) if (k < warp) { bc += misc->wcov[k]; ba += (acc_t) misc->wsum[k]; }
		ec = bc + ic - ctot + misc->base_cov;
		ea = ba + ia - atot + (acc_t) (long long) misc->base_val;
	}
	WB_PHASE(2);

	// ---- pass 2: running state after every position; the exact prefix goes back in place -----
	unsigned keepbits = 0;
	{
		int c = ec; acc_t a = ea;
#pragma unroll
		for (int k = 0; k < NV; k++) {
			const int pv = p0 + ((k ^ sw) << 2);
			const uint4 cv = *(const uint4 *) (cov + pv);
			const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w};
			uint4 l = *(const uint4 *) (limb + pv);
			uint32_t lo[4] = {l.x, l.y, l.z, l.w}, hi[4] = {0, 0, 0, 0};
			if (LIMBS == 2) { const uint4 h = *(const uint4 *) (limb + W + pv); hi[0] = h.x; hi[1] = h.y; hi[2] = h.z; hi[3] = h.w; }
#pragma unroll
			for (int j = 0; j < 4; j++) {
				c += (int) (cw[j] & FIELD) - (int) (cw[j] >> END_SHIFT);
				if (LIMBS == 1) a += (acc_t) lo[j];
				else a += (acc_t) ((long long) (unsigned long long) lo[j] + ((long long) (int32_t) hi[j] << lw));
				if (c >= thr) keepbits |= 1u << (4 * k + j);
				lo[j] = (uint32_t) a;                                  // LIMBS == 1: the sum itself fits 32 bits
				if (LIMBS == 2) hi[j] = (uint32_t) ((unsigned long long) a >> 32);
			}
			*(uint4 *) (limb + pv) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
			if (LIMBS == 2) *(uint4 *) (limb + W + pv) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
		}
		keepbits &= evbits;
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
#if WB_FS_LATETICKET == 1
		if (tid == 0) misc->ticket2 = (int) next_ticket;
#endif
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
#if WB_FS_LATETICKET
	// mode 2: the draw happens after the look-back, too late to prefetch for; windows are handed
	// out in order, so window w + grid starts about one window time after this one did -- its
	// intervals are pulled into L2 for whichever CTA draws it
	const int w1 = WB_FS_LATETICKET == 2 ? w + (int) gridDim.x : misc->ticket2;
	if (w1 < job.nw)
		for (int i = tid; i < N; i += TH) __pipeline_memcpy_async(&nxt_buf[i], &job.cursor[(size_t) w1 * N + i], 4);
	__pipeline_commit();
#else
	const int w1 = w1_early;
#endif
	const int32_t *nxt_cur = w1 < job.nw ? nxt_buf : nullptr;
	WB_PHASE(3);

	// ---- 3. dense breakpoint list over the dead cover array ---------------------------------
	uint32_t *dense = cov;
	{
		unsigned bits = evbits;
		while (bits) {
			const int q = __ffs(bits) - 1;
			bits &= bits - 1;
			uint32_t d = (uint32_t) (p0 + q);
			if ((keepbits >> q) & 1) {
				d |= D_KEEP | ((uint32_t) kidx++ << D_KID_SHIFT);
				if ((slowbits >> q) & 1) d |= D_SLOW;
			}
			dense[rank++] = d;
		}
	}
	__pipeline_wait_prior(0);                                  // this thread's part of the next cursor row has landed
	__syncthreads();
	WB_PHASE(4);
#if WB_FS_PREF >= 2
	// ask for the next window's intervals now: they travel while this window is emitted
	if (nxt_cur) {
		const long long *toff = (const long long *) (smem + L.off_toff);
		const int32_t *tn = (const int32_t *) (smem + L.off_tn);
		const int gid = tid >> 3, gl = tid & 7;
		for (int i = gid; i < N; i += TH / 8) {
			const long long o = toff[i], first = o + nxt_cur[i];
			const long long g = (first & ~3ll) + 4 * gl;
			if (g < o + tn[i] && g + 4 <= job.I) {
#if WB_FS_PREF == 2
				asm volatile("prefetch.global.L1 [%0];" :: "l"(job.start + g)); asm volatile("prefetch.global.L1 [%0];" :: "l"(job.end + g));
				asm volatile("prefetch.global.L1 [%0];" :: "l"(job.value + g)); asm volatile("prefetch.global.L1 [%0];" :: "l"(job.value + g + 2));
#else
				asm volatile("prefetch.global.L2 [%0];" :: "l"(job.start + g)); asm volatile("prefetch.global.L2 [%0];" :: "l"(job.end + g));
				asm volatile("prefetch.global.L2 [%0];" :: "l"(job.value + g)); asm volatile("prefetch.global.L2 [%0];" :: "l"(job.value + g + 2));
#endif
			}
		}
	}
#endif

	// ---- output offset: decoupled look-back over the preceding windows ----------------------
	if (warp == 0) {
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
#if WB_FS_LBRELAX
					if (pm != 0 && (nr == 0 || (pm & -pm) < (nr & -nr))) use = j + 1;
					else if (nr != 0) wait = true;
#else
					if (nr != 0) wait = true;
					else if (pm != 0) use = j + 1;
#endif
				}
				if (!wait) break;
				__nanosleep(WB_FS_LBSLEEP);
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
#if WB_FS_LATETICKET == 2
	unsigned next_ticket = 0;
	if (tid == 0) next_ticket = atomicAdd(job.ticket, 1u) + gridDim.x;      // lands while this window is emitted
#endif

	// ---- 4. emit, one breakpoint per thread (coalesced stores) -------------------------------
	const long long out_base = misc->out_base;
	const int next_bp = misc->next_bp;
	const int w0i = (int) w0;
	int32_t *os = job.out_start + out_base, *oe = job.out_end + out_base;
	double *ov = job.out_value + out_base;
	const long long room = job.out_cap - out_base;             // output ranks below this are in range
	const bool npow2 = (N & (N - 1)) == 0;
	const bool mean = job.op == OP_MEAN;
	// sum: x 2^q; mean over a power-of-two N: x 2^q / N is one exact power of two as well
	const double mul = (mean && npow2) ? scale * (1.0 / N) : scale;
	const bool div = mean && !npow2;
	constexpr uint32_t PMASK = (1u << D_KID_SHIFT) - 1;
	const int32_t fx_base32 = (int32_t) job.fx_base;
	double auc = 0;
	// the three output pointers of this window, opaque to the optimiser: one 32-bit
	// index per store instead of a re-derived 64-bit (out_base + rank) chain
	asm volatile("" : "+l"(os), "+l"(oe), "+l"(ov));
	const int room32 = (int) min(room, (long long) INT_MAX);
	auto emit_loop = [&](auto div_tag) {
		constexpr bool DIV = decltype(div_tag)::value;      // mean over a non-power-of-two N: one fp64 division per segment
#pragma unroll 2
		for (int r = tid; r < nbp; r += TH) {
			const uint32_t d = dense[r];
			if (!(d & D_KEEP)) continue;
			const int pos = (int) (d & PMASK), kid = (int) ((d >> D_KID_SHIFT) & PMASK);
			const int bs = w0i + pos;
			const int be = (r + 1 < nbp) ? w0i + (int) (dense[r + 1] & PMASK) : next_bp;
			const int ph = phys<VLOG>(pos);
			double v;                                                // correctly rounded exact sum
			if (LIMBS == 1) v = (double) ((int32_t) limb[ph] + fx_base32) * mul;     // |sum + defaults| < 2^30: one 32-bit conversion
			else v = (double) ((long long) (((unsigned long long) limb[W + ph] << 32) | limb[ph]) + job.fx_base) * mul;
			if (DIV) v = v / N;
			if (kid < room32) {
				os[kid] = bs;
				oe[kid] = be;
				ov[kid] = v;
				if (d & D_SLOW) {
					// an input was not representable: the fix-up kernel recomputes this
					// segment with the reference's own operation order
					const unsigned long long k = atomicAdd(job.slow_count, 1ull);
					if (k < (unsigned long long) job.slow_cap) job.slow_list[k] = out_base + kid;
				}
			}
			if (!(d & D_SLOW)) auc += (double) (be - bs) * v;        // v is finite here
		}
	};
	if (div) emit_loop(std::true_type{}); else emit_loop(std::false_type{});
	for (int d = 16; d > 0; d >>= 1) auc += __shfl_down_sync(FULL, auc, d);
	if (lane == 0) misc->wauc[warp] = auc;
#if WB_FS_LATETICKET == 2
	if (tid == 0) misc->ticket2 = (int) next_ticket;
#endif
	WB_PHASE(6);
}

template <int TH, int VLOG, int LIMBS>
__global__ void __launch_bounds__(TH, TH == 128 ? 6 : WB_FS_MINB) fastsum_kernel(const Job job, int lw, double inv_scale, double lim, double scale) {
	extern __shared__ __align__(16) unsigned char smem[];
	const Layout L = make_layout(job.wlog, job.N, LIMBS);
	int32_t *tn = (int32_t *) (smem + L.off_tn);
	long long *toff = (long long *) (smem + L.off_toff);
	long long *tmd = (long long *) (smem + L.off_tmd);
	for (int i = threadIdx.x; i < job.N; i += TH) {
		const long long o = job.off[i];
		toff[i] = o;
		tn[i] = (int) (job.off[i + 1] - o);
		tmd[i] = job.dflt_fx[i];
	}
	Misc *misc = (Misc *) smem;
	int32_t *ca = (int32_t *) (smem + L.off_cur), *cb = (int32_t *) (smem + L.off_cur2);   // current / next cursor rows
	// Windows are handed out in ascending order (the first two per CTA statically, then a
	// ticket counter): the look-back only ever waits on windows a resident CTA has drawn.
#if WB_FS_LATETICKET
	int w = blockIdx.x;
	if (w < job.nw)
		for (int i = threadIdx.x; i < job.N; i += TH) ca[i] = job.cursor[(size_t) w * job.N + i];
	__syncthreads();
	while (w < job.nw) {
#ifdef WB_DIAG_PHASES
		const long long diag_t0 = clock64();
#endif
		process_window<TH, VLOG, LIMBS>(job, w, smem, L, ca, cb, 0, lw, inv_scale, lim, scale);
		__syncthreads();      // dense list consumed, next cursor row landed, AUC partials visible
		if (threadIdx.x == 0 && job.auc_partial) {
			double t = 0;
			for (int k = 0; k < TH / 32; k++) t += misc->wauc[k];
			job.auc_partial[w] = t;
		}
#ifdef WB_DIAG_PHASES
		if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&job.diag[7], (unsigned long long) (t_ - diag_t0)); atomicAdd(&job.diag[8], 1ull); }
#endif
		w = misc->ticket2;
#if WB_FS_LATETICKET == 2
		// cb only ever holds the row of the window being prefetched for; the row of the window
		// just drawn travels into ca while the window arrays are cleared
		if (w < job.nw)
			for (int i = threadIdx.x; i < job.N; i += TH) __pipeline_memcpy_async(&ca[i], &job.cursor[(size_t) w * job.N + i], 4);
		__pipeline_commit();
#else
		{ int32_t *t = ca; ca = cb; cb = t; }
#endif
	}
#else
	int w = blockIdx.x, w1 = blockIdx.x + gridDim.x;
	if (w < job.nw)
		for (int i = threadIdx.x; i < job.N; i += TH) ca[i] = job.cursor[(size_t) w * job.N + i];
	__syncthreads();
	while (w < job.nw) {
#ifdef WB_DIAG_PHASES
		const long long diag_t0 = clock64();
#endif
		unsigned t2 = 0;
		if (threadIdx.x == 0) t2 = atomicAdd(job.ticket, 1u) + 2u * gridDim.x;   // consumed two windows from now
		if (w1 < job.nw)
			for (int i = threadIdx.x; i < job.N; i += TH)
				__pipeline_memcpy_async(&cb[i], &job.cursor[(size_t) w1 * job.N + i], 4);
		__pipeline_commit();
		process_window<TH, VLOG, LIMBS>(job, w, smem, L, ca, cb, w1, lw, inv_scale, lim, scale);
		if (threadIdx.x == 0) misc->ticket2 = (int) t2;
		__syncthreads();      // dense list consumed, next cursor row landed, ticket and AUC partials visible
		if (threadIdx.x == 0 && job.auc_partial) {
			double t = 0;
			for (int k = 0; k < TH / 32; k++) t += misc->wauc[k];
			job.auc_partial[w] = t;
		}
#ifdef WB_DIAG_PHASES
		if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&job.diag[7], (unsigned long long) (t_ - diag_t0)); atomicAdd(&job.diag[8], 1ull); }
#endif
		w = w1; w1 = misc->ticket2;
		{ int32_t *t = ca; ca = cb; cb = t; }
	}
#endif
}

template <int TH, int VLOG, int LIMBS>
static cudaError_t launch_v(const Job &job, int lw, double inv_scale, double lim, double scale, cudaStream_t s) {
	size_t smem = make_layout(job.wlog, job.N, LIMBS).total;
	{ const char *e = getenv("WB200_FS_PAD"); if (e) smem += (size_t) atoi(e) * 1024; }   // development: cap the CTAs per SM
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaError_t e = cudaFuncSetAttribute(fastsum_kernel<TH, VLOG, LIMBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (e != cudaSuccess) return e;
	int per = 1;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fastsum_kernel<TH, VLOG, LIMBS>, TH, smem);
	const int grid = min(job.nw, sms * max(per, 1));
	fastsum_kernel<TH, VLOG, LIMBS><<<grid, TH, smem, s>>>(job, lw, inv_scale, lim, scale);
	return cudaGetLastError();
}

} // namespace fs

// threads per window (development override: WB200_FS_THREADS)
static int fastsum_threads(const Job &job) {
	const char *e = getenv("WB200_FS_THREADS");
	const int th = e ? atoi(e) : 0;
	return (th == 128 || th == 256) ? th : 128;
}

// Whether the sum / mean sweep kernel covers this job (else: emit_kernel<MODE_FAST>).
bool fastsum_supported(const Job &job) {
	if (job.fast != 1 || !job.vec || job.fx_limbs > 2 || job.N > MAX_TRACKS) return false;
	const int vlog = job.wlog - (fastsum_threads(job) == 256 ? 8 : 7);
	if (vlog != 4 && vlog != 5) return false;
	return fs::make_layout(job.wlog, job.N, job.fx_limbs).total <= 200 * 1024;
}

cudaError_t launch_fastsum(const Job &job, int lw, double inv_scale, double lim, double scale, cudaStream_t s) {
	const bool big = fastsum_threads(job) == 256;
	const int vlog = job.wlog - (big ? 8 : 7);
	const bool one = job.fx_limbs == 1;
#define WB_FS(TH, V) (one ? fs::launch_v<TH, V, 1>(job, lw, inv_scale, lim, scale, s) : fs::launch_v<TH, V, 2>(job, lw, inv_scale, lim, scale, s))
	if (big && vlog == 4) return WB_FS(256, 4);
	if (!big && vlog == 4) return WB_FS(128, 4);
	if (!big && vlog == 5) return WB_FS(128, 5);
#undef WB_FS
	return cudaErrorInvalidValue;
}

} // namespace wb
