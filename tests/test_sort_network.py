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
