"""CPU model of the selection kernel's index arithmetic (csrc/engine.cu: bitonic_u64, sort_phys, sel_word).

The device code claims three things "by enumeration"; this file is that enumeration:
  * the register-blocked trip schedule (E = 8 or 16 keys per thread, up to log2(E) levels per trip) is a
    sorting network for every power-of-two size the kernel uses;
  * the xor swizzle `idx ^ ((idx >> log2(E)) & 15)` makes every gather of a trip bank-conflict free for
    64-bit shared-memory accesses (each half-warp touches 16 distinct 8-byte bank pairs);
  * the counter rows `p ^ ((p >> lsh) & 3)` are a permutation of the positions and put the four runs a warp
    walks in the prefix pass on different bank groups.
No GPU involved: the parity of the kernel itself is tests/test_orderstat.py."""
import random

import pytest


def trips(le, ps):
    """(b0, b, lk) of every trip after the in-register phase, as bitonic_u64<LE> runs them."""
    out = []
    lk = le + 1
    while (1 << lk) <= ps:
        b = lk - 1
        while b >= 0:
            b0 = max(b - (le - 1), 0)
            out.append((b0, b, lk))
            b = b0 - 1
        lk += 1
    return out


def run_network(keys, le):
    e, ps = 1 << le, len(keys)
    a = list(keys)

    def ce(r, x, y, up):
        if (r[x] > r[y]) if up else (r[x] < r[y]):
            r[x], r[y] = r[y], r[x]

    for vt in range(ps >> le):
        base = vt << le
        r = a[base:base + e]
        k = 2
        while k <= e:
            j = k >> 1
            while j:
                for m in range(e):
                    if not m & j:
                        ce(r, m, m | j, (m & k) == 0 if k < e else (base & e) == 0)
                j >>= 1
            k <<= 1
        a[base:base + e] = r
    for b0, b, lk in trips(le, ps):
        for vt in range(ps >> le):
            base = ((vt >> b0) << (b0 + le)) | (vt & ((1 << b0) - 1))
            up = (base & (1 << lk)) == 0
            idx = [base | (m << b0) for m in range(e)]
            r = [a[i] for i in idx]
            for sft in range(le - 1, -1, -1):
                if b0 + sft <= b:
                    for m in range(e):
                        if not m & (1 << sft):
                            ce(r, m, m | (1 << sft), up)
            for i, v in zip(idx, r):
                a[i] = v
    return a


@pytest.mark.parametrize("le,ps", [(3, 256), (3, 512), (3, 1024), (3, 2048), (4, 4096)])
def test_trip_schedule_sorts(le, ps):
    rng = random.Random(ps + le)
    for _ in range(3):
        keys = [rng.getrandbits(32) for _ in range(ps)]
        for _ in range(ps // 16):
            keys[rng.randrange(ps)] = 0xFFFFFFFF           # the padding key, many times
        assert run_network(keys, le) == sorted(keys)


def sort_phys(idx, le):
    return idx ^ ((idx >> le) & 15)


@pytest.mark.parametrize("le,ps", [(3, 512), (3, 2048), (4, 4096)])
def test_swizzle_is_conflict_free(le, ps):
    e = 1 << le
    assert sorted(sort_phys(i, le) for i in range(ps)) == list(range(ps))          # a permutation
    schedule = [0] + [b0 for b0, _, _ in trips(le, ps)]
    for b0 in schedule:
        for warp in range(min((ps >> le) // 32, 8)):
            for m in range(e):
                addr = []
                for lane in range(32):
                    vt = warp * 32 + lane
                    base = ((vt >> b0) << (b0 + le)) | (vt & ((1 << b0) - 1))
                    addr.append(sort_phys(base | (m << b0), le))
                for half in (addr[:16], addr[16:]):                                # 64-bit accesses go by half-warps
                    assert len({x % 16 for x in half}) == 16, (le, b0, m, half)


@pytest.mark.parametrize("ws", [1, 2, 8, 32, 128, 512, 1024, 4096])
def test_counter_rows(ws):
    runs = 32                                               # THREADS / (SEL_NB / 4)
    lsh = 2
    while (runs << lsh) < ws:
        lsh += 1
    rows = [p ^ ((p >> lsh) & 3) for p in range(ws)]
    assert sorted(rows) == list(range(ws))                  # every position keeps a row of its own inside the range
    if ws >= 128:
        length = ws // runs
        for first in range(0, runs, 4):                     # the four runs one warp walks in the prefix pass
            for step in range(length):
                groups = {(rows[(first + j) * length + step]) & 3 for j in range(4)}
                assert len(groups) == 4, (ws, first, step)


# ---- the selection itself, as the kernel organises it ------------------------------------------------------------
def select_model(tracks, defaults, w0, width, k_sel, nb=32):
    """Median-like selection at every position of [w0, w0 + width) the way emit_kernel<MODE_SELECT> does it:
    elements = intervals overlapping the range + the gaps around them (one element per track in play at every
    position), ranks from one sort by value, byte counters per (position, bucket) packed four to a 32-bit word
    and accumulated with wrap-around arithmetic, then bucket search + k-th element in play inside the bucket."""
    import numpy as np
    ra, rb = w0, w0 + width
    elems = []                                                  # (value, start, end) clipped to the range
    for (s, e, v), d in zip(tracks, defaults):
        lo = int(np.searchsorted(e, ra, side="right"))          # first interval with end > ra
        hi = int(np.searchsorted(s, rb, side="left"))           # first interval with start >= rb
        prev = ra
        for j in range(lo, hi):
            a, b = max(int(s[j]), ra), min(int(e[j]), rb)
            if prev < a:
                elems.append((float(d), prev, a))               # the gap before interval j
            elems.append((float(v[j]), a, b))
            prev = b
        if prev < rb:
            elems.append((float(d), prev, rb))
    elems.sort(key=lambda t: t[0])                              # ties: any order gives the same value
    r = len(elems)
    bs = max(4, (-(-r // nb) + 3) & ~3)
    assert bs <= 255
    words = np.zeros((width, nb // 4), dtype=np.uint32)
    with np.errstate(over="ignore", under="ignore"):            # the wrap-around is the point
        for rank, (_, a, b) in enumerate(elems):
            bk = rank // bs
            one = np.uint32(1 << (8 * (bk & 3)))
            words[a - ra, bk >> 2] += one                       # +1 at the start ...
            if b < rb:
                words[b - ra, bk >> 2] -= one                   # ... -1 at the end: uint32 wrap-around, as the atomics do
        words = np.cumsum(words, axis=0, dtype=np.uint32)       # prefix along the positions, still modulo 2^32
    out = np.empty(width)
    for p in range(width):
        counts = [(int(words[p, b >> 2]) >> (8 * (b & 3))) & 0xFF for b in range(nb)]
        assert sum(counts) == len(tracks)                       # exactly one element per track in play
        k, bk = k_sel, 0
        while k >= counts[bk]:
            k -= counts[bk]
            bk += 1
        for val, a, b in elems[bk * bs:(bk + 1) * bs]:
            if a <= ra + p < b:
                if k == 0:
                    out[p] = val
                    break
                k -= 1
    return out


@pytest.mark.parametrize("n,seed", [(25, 1), (64, 2), (200, 3)])
def test_selection_model_against_brute_force(n, seed):
    import numpy as np
    from wiggletools_b200.synth import make_tracks
    tr = make_tracks(n, [3_000], seed=seed, vmax=24, max_len=60, max_gap=40)
    rng = np.random.default_rng(seed)
    defaults = rng.integers(-2, 5, n).astype(float)
    tracks = [(t["start"], t["end"], t["value"]) for t in tr]
    w0, width = 1024, 512
    got = select_model(tracks, defaults, w0, width, n // 2)
    dense = np.tile(defaults[:, None], (1, width))              # brute force: every track's value at every position
    for i, (s, e, v) in enumerate(tracks):
        for a, b, x in zip(s, e, v):
            a, b = max(int(a), w0), min(int(b), w0 + width)
            if a < b:
                dense[i, a - w0:b - w0] = x
    want = np.sort(dense, axis=0)[n // 2]                       # upper median (reducers.c:769-813)
    assert np.array_equal(got, want)
