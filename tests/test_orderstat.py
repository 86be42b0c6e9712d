"""min / max / median over more than 24 tracks: the window-sorting sweeps of emit_kernel<MODE_SELECT>
(csrc/engine.cu).  Reference: MedianReductionPop / MinReductionPop / MaxReductionPop,
/root/reference/src/reducers.c:131-257 (qsort of the N current values, upper median, track 0's
default read as 0 by min / max).

The bucketed selection handles windows whose values are all float32-exact (one 64-bit key per
interval and per track default); anything else -- doubles with more mantissa, NaN, -0 -- takes the serial
rank-bitset sweep.  Every value class below must give the oracle's output bit for bit, whichever
sweep ends up running, including chromosomes where the classes are mixed window by window."""
import numpy as np
import pytest

from wiggletools_b200 import api
from wiggletools_b200.synth import make_tracks
from oracle.pyoracle import Track as OTrack
from util import assert_segments

pytestmark = pytest.mark.gpu

NAMES = ["chr1", "chr10"]


def both(tr, defaults=None):
    d = defaults if defaults is not None else [0.0] * len(tr)
    o = [OTrack(t["chrom"], t["start"], t["end"], t["value"], default=float(d[k])) for k, t in enumerate(tr)]
    g = [api.Track(t["chrom"], t["start"], t["end"], t["value"], default=float(d[k])) for k, t in enumerate(tr)]
    return o, g


def check(engine, port, tr, ops, what, defaults=None):
    o, g = both(tr, defaults)
    for op in ops:
        want = port.reduce(op, o, NAMES)
        got = engine.reduce(op, g, NAMES)
        assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="%s/%s" % (op, what))


@pytest.mark.parametrize("n", [25, 33, 64, 129, 200, 256, 300, 512])
def test_float32_values(engine, port, n):
    """Signed, fractional float32 values: the bucketed selection."""
    tr = make_tracks(n, [6_000, 2_500], seed=31 + n)
    rng = np.random.default_rng(n)
    for t in tr:
        t["value"] = rng.normal(0, 50, len(t["value"])).astype(np.float32).astype(np.float64)
    check(engine, port, tr, ("median", "min", "max"), "f32/%d" % n)


def test_float32_values_with_defaults(engine, port):
    """Non-zero defaults (one per track, some equal to data values): gap elements carry them; min / max
    read track 0's default as 0 (reducers.c:142-145, 209-212)."""
    n = 200
    tr = make_tracks(n, [5_000, 3_000], seed=77, vmax=16)
    rng = np.random.default_rng(5)
    defaults = rng.integers(-3, 12, n).astype(np.float64)
    defaults[0] = 7.5
    check(engine, port, tr, ("median", "min", "max"), "defaults", defaults)


def test_heavy_ties_and_abutting_intervals(engine, port):
    """Two distinct values only, and tracks whose intervals abut (empty gaps between them)."""
    n = 256
    tr = make_tracks(n, [5_000, 2_000], seed=3, vmax=2, max_gap=1)
    check(engine, port, tr, ("median", "max"), "ties")


@pytest.mark.parametrize("n", [48, 160, 300])
def test_double_values_take_the_serial_sweep(engine, port, n):
    """Values with more mantissa than a float32: same answers from the rank-bitset sweep (above 256 tracks
    half the warps run it)."""
    tr = make_tracks(n, [4_000, 1_500], seed=9)
    rng = np.random.default_rng(2)
    for t in tr:
        t["value"] = rng.random(len(t["value"])) * 1e-3 + 1.0 / 3.0
    check(engine, port, tr, ("median", "min"), "f64/%d" % n)


def test_mixed_windows_nan_and_negative_zero(engine, port):
    """One chromosome, value class changing along it: float32 stretches, a stretch with doubles, NaN
    values (the segment's result is NaN) and -0.0 among zeros (sign kept as the reference's stable
    sort leaves it)."""
    n = 192
    tr = make_tracks(n, [12_000, 1_000], seed=21, vmax=4)
    rng = np.random.default_rng(8)
    for k, t in enumerate(tr):
        v = t["value"].astype(np.float64)
        s = t["start"]
        on_chr1 = t["chrom"] == 0
        v[on_chr1 & (s > 3_000) & (s < 5_000)] += 1e-9 * (k + 1)          # doubles
        if k % 17 == 0:
            v[on_chr1 & (s > 6_000) & (s < 7_000)] = np.nan
        zero = on_chr1 & (s > 8_000) & (s < 10_000)
        v[zero] = np.where(rng.random(int(zero.sum())) < 0.5, -0.0, 0.0)
        t["value"] = v
    check(engine, port, tr, ("median", "min", "max"), "mixed")


@pytest.mark.parametrize("n", [48, 160, 300])
def test_dense_window_falls_back(engine, port, n):
    """More intervals per window than the staging area holds (1-bp intervals on every track):
    the per-warp sorted arrays take over; results unchanged."""
    tr = make_tracks(n, [1_500], seed=4, max_len=1, max_gap=1, vmax=9)
    o, g = both(tr)
    want = port.reduce("median", o, ["chr1"])
    got = engine.reduce("median", g, ["chr1"])
    assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="median/dense")


def test_strict_and_empty_windows(engine, port):
    """`strict` keeps only the segments every track covers (most windows then have nothing to select and are
    skipped), and sparse tracks leave whole windows without a breakpoint."""
    n = 48
    tr = make_tracks(n, [20_000, 3_000], seed=15, vmax=32, max_len=900, max_gap=120)
    o, g = both(tr)
    for strict in (False, True):
        want = port.reduce("median", o, NAMES, strict)
        got = engine.reduce("median", g, NAMES, strict=strict)
        assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="median/strict=%d" % strict)
    sparse = make_tracks(40, [60_000], seed=16, vmax=8, max_len=30, max_gap=6000, gap_prob=0.9)
    o, g = both(sparse)
    want = port.reduce("max", o, ["chr1"])
    got = engine.reduce("max", g, ["chr1"])
    assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="max/sparse")


def test_infinite_values_are_float32_too(engine, port):
    """+inf / -inf are float32 values: they stay on the bucketed selection (largest / smallest keys below the
    padding) and come out as the reference's sort leaves them."""
    n = 40
    tr = make_tracks(n, [6_000], seed=23, vmax=16)
    rng = np.random.default_rng(4)
    for t in tr:
        v = t["value"].copy()
        r = rng.random(len(v))
        v[r < 0.05] = np.inf
        v[r > 0.96] = -np.inf
        t["value"] = v
    o, g = both(tr)
    for op in ("median", "min", "max"):
        want = port.reduce(op, o, ["chr1"])
        got = engine.reduce(op, g, ["chr1"])
        assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what=op + "/inf")
