"""GPU parity tests of the sum / mean sweep kernel (csrc/fastsum.cu) and of the
host logic that picks its accumulator format: every case is compared with the
CPU oracle (reference restatement: src/multiplexer.c:98-121 +
src/reducers.c:259-292, 367-402) on the same inputs -- coordinates bit for bit,
values bit for bit unless the reference's own left-to-right fp64 rounding can
differ from the exact sum (then RTOL = 1e-10)."""
import ctypes as C

import numpy as np
import pytest

from wiggletools_b200 import api
from wiggletools_b200.synth import make_tracks
from oracle.pyoracle import Track as OTrack
from util import RTOL, assert_segments, same_float

pytestmark = pytest.mark.gpu
NAMES = ["chr1", "chr2"]


def both(tr, defaults=None):
    o = [OTrack(t["chrom"], t["start"], t["end"], t["value"], default=(defaults[i] if defaults else 0.0)) for i, t in enumerate(tr)]
    g = [api.Track(t["chrom"], t["start"], t["end"], t["value"], default=(defaults[i] if defaults else 0.0)) for i, t in enumerate(tr)]
    return o, g


def check(engine, port, tr, defaults=None, exact=True, names=NAMES):
    o, g = both(tr, defaults)
    for op in ("sum", "mean"):
        for strict in (False, True):
            want = port.reduce(op, o, names, strict)
            got = engine.reduce(op, g, names, strict=strict)
            assert_segments(got, want.chrom, want.start, want.end, want.value, exact=exact, rtol=RTOL,
                            what="%s strict=%d" % (op, strict))
            assert same_float(got.default, want.default)
    return engine.last_stats


def test_integer_counts_take_one_limb(engine, port):
    """Read-count-like data: one 32-bit accumulator per position, sweep kernel, nothing on the slow path."""
    tr = make_tracks(20, [60000, 9000], seed=5, max_len=50, max_gap=30)
    for t in tr:
        t["value"] = np.floor(t["value"])
    stats = check(engine, port, tr)
    assert all(s["fast_path"] == 1 and s["limbs"] == 1 and s["sweep_kernel"] == 1 and s["slow_segments"] == 0 for s in stats), stats


def test_float32_mantissas_take_two_limbs(engine, port):
    """Full 24-bit mantissas do not fit one limb: two limbs, still exact (and the
    reference's fp64 running sum of such values is exact too: bit-identical)."""
    rng = np.random.default_rng(11)
    tr = make_tracks(20, [50000], seed=6, max_len=60, max_gap=40)
    for t in tr:
        t["value"] = (1.0 + rng.random(len(t["value"]))).astype(np.float32).astype(np.float64)
    stats = check(engine, port, tr, names=["chr1"])
    assert all(s["limbs"] == 2 and s["sweep_kernel"] == 1 and s["slow_segments"] == 0 for s in stats), stats


def test_sample_misses_fine_values_and_the_run_is_repeated_wider(engine, port):
    """The format is chosen from a strided sample of the values.  Here every sampled
    value is an integer while a tenth of the unsampled ones carry a bit at 2^-20: the
    one-limb run marks those intervals for the fix-up path, the host sees too many of
    them and repeats the chromosome with two limbs.  Results must not change."""
    n_tracks = 8
    tr = make_tracks(n_tracks, [5_200_000], seed=8, max_len=400, max_gap=200)
    total = sum(len(t["start"]) for t in tr)
    stride = total // (1 << 16)
    assert stride >= 2, "test needs more than 131072 intervals (has %d)" % total
    rng = np.random.default_rng(3)
    off = 0
    for t in tr:
        v = np.floor(t["value"])
        idx = off + np.arange(len(v))
        fine = (idx % stride != 0) & (rng.random(len(v)) < 0.1)
        v[fine] += 2.0 ** -20
        t["value"] = v
        off += len(v)
    o, g = both(tr)
    want = port.reduce("mean", o, ["chr1"])
    got = engine.reduce("mean", g, ["chr1"])
    assert_segments(got, want.chrom, want.start, want.end, want.value, rtol=RTOL, what="mean")
    st = engine.last_stats[-1]
    assert st["limbs"] == 2 and st["slow_segments"] == 0 and st["sweep_kernel"] == 1, st


def test_nan_storm_overflows_the_poison_list(engine, port):
    """Hundreds of NaN intervals inside one window: more poison events than the
    kernel's side list holds, so the whole window goes through the fix-up kernel."""
    s = np.arange(10, 10 + 2 * 300, 2, dtype=np.int32)
    nan_track = dict(chrom=np.zeros(len(s), np.int32), start=s, end=s + 1, value=np.full(len(s), np.nan))
    tr = make_tracks(6, [20000], seed=12, max_len=30, max_gap=10)
    tr.append(nan_track)
    stats = check(engine, port, tr, names=["chr1"])
    assert any(s["slow_segments"] > 0 for s in stats)


def test_sparse_nans_and_nonzero_defaults(engine, port):
    tr = make_tracks(9, [50000, 12000], seed=21, max_len=80, max_gap=60, nan_prob=0.01, neg=True)
    check(engine, port, tr, defaults=[float(i % 4) - 1.0 for i in range(9)])


@pytest.mark.parametrize("n", [3, 100, 300])
def test_track_counts(engine, port, n):
    """Below, around and above 256 tracks (the cover word keeps two 11-bit counts per position)."""
    tr = make_tracks(n, [16000, 3000], seed=40 + n, max_len=120, max_gap=200)
    stats = check(engine, port, tr)
    assert all(s["sweep_kernel"] == 1 for s in stats)


def test_every_track_starts_and_ends_at_the_same_position(engine, port):
    """N starts and N ends on single positions: the per-position counters at their maximum."""
    n = 200
    tr = [dict(chrom=np.zeros(2, np.int32), start=np.array([100, 5000], np.int32), end=np.array([5000, 9000], np.int32),
               value=np.array([float(i), 1.0])) for i in range(n)]
    check(engine, port, tr, names=["chr1"])


def test_long_runs_of_one_track_in_a_window(engine, port):
    """One track with 1-bp intervals (thousands per window: many 32-interval steps) beside sparse ones."""
    s = np.arange(1, 9001, dtype=np.int32)
    dense = dict(chrom=np.zeros(len(s), np.int32), start=s, end=s + 1, value=(s % 7).astype(np.float64))
    tr = make_tracks(5, [12000], seed=77, max_len=500, max_gap=900)
    tr.insert(2, dense)
    check(engine, port, tr, names=["chr1"])


def test_unaligned_device_arrays_fall_back(engine):
    """wb200_reduce_device on arrays that are not 16-byte aligned: the 128-bit loads are off,
    the call runs on emit_kernel and returns the same segments."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    tr = make_tracks(7, [30000], seed=90, max_len=70, max_gap=50)
    n = len(tr)
    off = np.concatenate(([0], np.cumsum([len(t["start"]) for t in tr]))).astype(np.int64)
    I = int(off[-1])
    hs = np.concatenate([t["start"] for t in tr]).astype(np.int32)
    he = np.concatenate([t["end"] for t in tr]).astype(np.int32)
    hv = np.concatenate([t["value"] for t in tr]).astype(np.float64)
    dflt = np.zeros(n)
    cap = 2 * I + 1
    lib, ctx = engine.lib, engine.ctx
    results = []
    for shift in (0, 1):
        ds = torch.zeros(I + 4, dtype=torch.int32, device=dev); ds[shift:shift + I] = torch.from_numpy(hs).to(dev)
        de = torch.zeros(I + 4, dtype=torch.int32, device=dev); de[shift:shift + I] = torch.from_numpy(he).to(dev)
        dv = torch.zeros(I + 4, dtype=torch.float64, device=dev); dv[shift:shift + I] = torch.from_numpy(hv).to(dev)
        os_, oe_ = torch.zeros(cap, dtype=torch.int32, device=dev), torch.zeros(cap, dtype=torch.int32, device=dev)
        ov_ = torch.zeros(cap, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        nout, dflt_out, auc = C.c_int64(0), C.c_double(0), C.c_double(0)
        rc = lib.wb200_reduce_device(ctx, 2, n, n, 0, ds.data_ptr() + 4 * shift, de.data_ptr() + 4 * shift, dv.data_ptr() + 8 * shift,
                                     off.ctypes.data, dflt.ctypes.data, 1, 30001, os_.data_ptr(), oe_.data_ptr(), ov_.data_ptr(),
                                     cap, C.addressof(nout), C.addressof(dflt_out), C.addressof(auc))
        assert rc == 0, lib.wb200_last_error(ctx)
        st = engine.stats()
        assert st["sweep_kernel"] == (1 if shift == 0 else 0), st
        S = nout.value
        results.append((os_[:S].cpu().numpy(), oe_[:S].cpu().numpy(), ov_[:S].cpu().numpy(), auc.value))
    a, b = results
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.isclose(a[3], b[3], rtol=1e-12)


def test_narrow_transport_is_lossless(engine, port):
    """float32 value inputs (widened on the device) and breakpoint-form read-back (end[] only where it is
    not the next start) through wb200_reduce_host_batch: the rebuilt arrays equal the wide transport's, for
    gap-free data (few explicit ends), data with gaps (some) and very sparse data (falls back to the full form)."""
    import ctypes as C
    from wiggletools_b200 import _lib
    from wiggletools_b200.synth import make_tracks
    lib = engine.lib
    for seed, kw in ((1, dict(max_len=400, max_gap=200)), (2, dict(max_len=30, max_gap=4000, gap_prob=0.9)), (3, dict(max_len=200, max_gap=1, gap_prob=0.0))):
        n = 6
        tr = make_tracks(n, [60_000, 20_000], seed=900 + seed, **kw)
        shards = []
        for c in range(2):
            per = []
            for t in tr:
                sel = t["chrom"] == c
                per.append((np.ascontiguousarray(t["start"][sel]), np.ascontiguousarray(t["end"][sel]), np.ascontiguousarray(t["value"][sel])))
            shards.append(per)
        want = engine.reduce_shards("mean", [[(s, e, v, 0.0) for s, e, v in per] for per in shards])
        keep = []
        arr = (_lib.CTrack * (2 * n))()
        for c, per in enumerate(shards):
            for i, (s, e, v) in enumerate(per):
                v32 = np.ascontiguousarray(v, np.float32)
                assert np.array_equal(v32.astype(np.float64), v)
                keep.append(v32)
                arr[c * n + i] = _lib.CTrack(s.ctypes.data, e.ctypes.data, None, len(s), 0.0, v32.ctypes.data)
        out = (_lib.CSegments * 2)()
        lib.wb200_set_breakpoint_output(engine.ctx, 1)
        try:
            engine._check(lib.wb200_reduce_host_batch(engine.ctx, 2, 2, n, n, 0, C.addressof(arr), C.addressof(out)))
        finally:
            lib.wb200_set_breakpoint_output(engine.ctx, 0)
        forms = []
        for c in range(2):
            S = int(out[c].n)
            assert S == len(want[c][0])
            gs = np.ctypeslib.as_array(C.cast(out[c].start, C.POINTER(C.c_int32)), (S,))
            gv = np.ctypeslib.as_array(C.cast(out[c].value, C.POINTER(C.c_double)), (S,))
            if out[c].end:
                ge = np.ctypeslib.as_array(C.cast(out[c].end, C.POINTER(C.c_int32)), (S,)).copy()
                forms.append("full")
            else:
                k = int(out[c].n_ends)
                idx = np.ctypeslib.as_array(C.cast(out[c].end_index, C.POINTER(C.c_int32)), (k,))
                assert np.all(np.diff(idx) > 0) and idx[-1] == S - 1
                ge = np.empty(S, np.int32); ge[:-1] = gs[1:]
                ge[idx] = np.ctypeslib.as_array(C.cast(out[c].end_value, C.POINTER(C.c_int32)), (k,))
                forms.append("breakpoint")
            assert np.array_equal(gs, want[c][0]) and np.array_equal(ge, want[c][1]) and np.array_equal(gv, want[c][2])
        if seed == 2:
            assert "full" in forms          # explicit ends on more than a quarter of the segments: sent in the full form
        if seed == 3:
            assert forms == ["breakpoint", "breakpoint"]
