"""GPU parity tests: the CUDA engine, called through the C ABI
(wb200_reduce_host / wb200_pearson_host), against the CPU oracle on the same
seeded inputs and against the committed fixtures generated from the compiled
reference.  Coordinates must be bit-identical; fp64 values within RTOL=1e-10
relative (most are bit-identical and asserted so where the algorithm guarantees
it)."""
import math

import numpy as np
import pytest

from wiggletools_b200 import api
from wiggletools_b200.synth import make_tracks
from oracle.pyoracle import Track as OTrack
from util import GOLDEN_NAMES, REDUCE_OPS, RTOL, assert_segments, golden_tracks, load_golden, same_float

pytestmark = pytest.mark.gpu

NAMES = ["chr1", "chr10", "chr2"]
# var / stddev / entropy / CV come from exact integer moments (one final rounding) when the
# inputs allow it: equal to the reference within a few ulps, not always bit for bit
MOMENT_OPS = ("var", "stddev", "entropy", "cv")


def both(tr, defaults=None):
    o = [OTrack(t["chrom"], t["start"], t["end"], t["value"], default=(defaults[i] if defaults else 0.0)) for i, t in enumerate(tr)]
    g = [api.Track(t["chrom"], t["start"], t["end"], t["value"], default=(defaults[i] if defaults else 0.0)) for i, t in enumerate(tr)]
    return o, g


def test_committed_golden_fixtures(engine):
    z, cases = load_golden()
    for case in cases:
        tracks = golden_tracks(z, case, cls=api.Track)
        for op in REDUCE_OPS:
            for strict in (0, 1):
                key = "%s/%s/%d" % (case, op, strict)
                r = engine.reduce(op, tracks, GOLDEN_NAMES, strict=bool(strict))
                assert_segments(r, z[key + "/chrom"], z[key + "/start"], z[key + "/end"], z[key + "/value"], what=key)
                assert same_float(r.default, float(z[key + "/default"])), key
        want = float(z[case + "/auc_mean"])
        got = engine.auc("mean", tracks, GOLDEN_NAMES)
        assert math.isclose(got, want, rel_tol=RTOL), (case, got, want)
        if len(tracks) >= 2:
            got, want = engine.pearson(tracks[0], tracks[1], GOLDEN_NAMES), float(z[case + "/pearson01"])
            assert same_float(got, want) or math.isclose(got, want, rel_tol=1e-9), (case, got, want)
        if len(tracks) >= 6:
            for op in ("ttest", "wilcoxon"):
                key = "%s/%s" % (case, op)
                r = engine.setcompare(op, tracks[:3], tracks[3:], GOLDEN_NAMES)
                assert_segments(r, z[key + "/chrom"], z[key + "/start"], z[key + "/end"], z[key + "/value"], what=key)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("n", [1, 2, 5, 33])
def test_random_vs_oracle(engine, port, seed, n):
    kw = dict(nan_prob=0.02 if seed % 2 else 0.0, neg=(seed == 2), align=(1 if seed < 3 else 10))
    tr = make_tracks(n, [30000, 9000, 18000], seed=200 + seed, max_len=40, max_gap=30, **kw)
    defaults = None if seed < 2 else [float(i % 3) for i in range(n)]
    o, g = both(tr, defaults)
    for op in REDUCE_OPS:
        for strict in (False, True):
            want = port.reduce(op, o, NAMES, strict)
            got = engine.reduce(op, g, NAMES, strict=strict)
            # every statistic here follows the reference's operation order exactly
            # (row path), or is an exact integer sum (fast path): bit-identical
            assert_segments(got, want.chrom, want.start, want.end, want.value, exact=(op not in MOMENT_OPS), what="%s strict=%d" % (op, strict))
            assert same_float(got.default, want.default)
    if n >= 6:
        for op in ("ttest", "wilcoxon"):
            want = port.setcompare(op, o[:3], o[3:], NAMES)
            got = engine.setcompare(op, g[:3], g[3:], NAMES)
            assert_segments(got, want.chrom, want.start, want.end, want.value, what=op)
    if n >= 2:
        want, got = port.pearson(o[0], o[1], NAMES), engine.pearson(g[0], g[1], NAMES)
        assert same_float(got, want) or math.isclose(got, want, rel_tol=1e-9), (got, want)
    want, got = port.auc("mean", o, NAMES), engine.auc("mean", g, NAMES)
    assert same_float(got, want) or math.isclose(got, want, rel_tol=RTOL)


def test_general_doubles_within_tolerance(engine, port):
    """Values with full 53-bit mantissas and a wide dynamic range: the
    fixed-point path must fall back per segment where it cannot be exact."""
    rng = np.random.default_rng(7)
    tr = make_tracks(12, [40000], seed=9, max_len=60, max_gap=20)
    for i, t in enumerate(tr):
        t["value"] = rng.lognormal(0.0, 3.0 if i % 2 else 0.5, len(t["value"]))
    tr[3]["value"][::17] = np.inf
    tr[4]["value"][::13] = 1e300
    tr[5]["value"][::11] = 1e-300
    o, g = both(tr)
    for op in ("sum", "mean"):
        want = port.reduce(op, o, ["chr1"])
        got = engine.reduce(op, g, ["chr1"])
        assert_segments(got, want.chrom, want.start, want.end, want.value, rtol=RTOL, what=op)
    st = engine.last_stats[-1]
    assert st["fast_path"] == 1 and st["slow_segments"] > 0


@pytest.mark.parametrize("window", [1024, 2048, 4096])
def test_window_size_does_not_change_results(port, window):
    tr = make_tracks(9, [70000, 5000], seed=31, max_len=300, max_gap=2500)
    # long intervals spanning many windows, and long empty stretches
    tr[0]["start"] = np.array([10, 30000], np.int32); tr[0]["end"] = np.array([25000, 69000], np.int32)
    tr[0]["value"] = np.array([2.5, 7.0]); tr[0]["chrom"] = np.zeros(2, np.int32)
    o, g = both(tr)
    eng = api.Engine(0, window=window)
    for op in ("mean", "max", "var", "median"):
        for strict in (False, True):
            want = port.reduce(op, o, ["chr1", "chr2"], strict)
            got = eng.reduce(op, g, ["chr1", "chr2"], strict=strict)
            assert_segments(got, want.chrom, want.start, want.end, want.value, exact=(op not in MOMENT_OPS), what="%s w=%d" % (op, window))
    eng.close()


def test_edge_cases(engine, port):
    names = ["chr1", "chr2"]
    e = np.empty(0)
    cases = {
        "all_empty": [dict(chrom=e, start=e, end=e, value=e)] * 3,
        "one_empty": [dict(chrom=[0, 0], start=[5, 9], end=[7, 12], value=[1.0, 2.0]), dict(chrom=e, start=e, end=e, value=e)],
        "abutting": [dict(chrom=[0, 0, 0], start=[1, 4, 9], end=[4, 9, 20], value=[1.0, 1.0, 3.0]),
                     dict(chrom=[0, 0], start=[4, 9], end=[9, 10], value=[5.0, 6.0])],
        "identical": [dict(chrom=[0, 1], start=[100, 7], end=[200, 9], value=[2.0, 4.0])] * 4,
        "disjoint_chroms": [dict(chrom=[0], start=[3], end=[8], value=[1.5]), dict(chrom=[1], start=[3], end=[8], value=[2.5])],
        "window_edges": [dict(chrom=[0, 0, 0], start=[1, 2049, 8193], end=[2049, 8193, 8194], value=[1.0, 2.0, 3.0]),
                         dict(chrom=[0, 0], start=[4096, 4098], end=[4097, 12289], value=[4.0, 8.0])],
        "big_coords": [dict(chrom=[0, 0], start=[2147400000, 2147480000], end=[2147450000, 2147483000], value=[1.0, 2.0]),
                       dict(chrom=[0], start=[2147449999], end=[2147480001], value=[0.5])],
    }
    for name, tr in cases.items():
        o, g = both(tr)
        for op in ("sum", "mean", "min", "var", "median"):
            for strict in (False, True):
                want = port.reduce(op, o, names, strict)
                got = engine.reduce(op, g, names, strict=strict)
                assert_segments(got, want.chrom, want.start, want.end, want.value, exact=(op not in MOMENT_OPS), what="%s %s" % (name, op))
                assert same_float(got.default, want.default)


def test_arity_errors(engine):
    t = api.Track([0], [1], [2], [1.0])
    with pytest.raises(api.WB200Error):
        engine.setcompare("ttest", [t, t], [t, t, t])
    with pytest.raises(api.WB200Error):
        engine.reduce("mean", [])


def test_many_tracks(engine, port):
    tr = make_tracks(300, [6000], seed=77, max_len=50, max_gap=40)
    o, g = both(tr)
    for op in ("mean", "stddev", "median", "max"):
        want = port.reduce(op, o, ["chr1"])
        got = engine.reduce(op, g, ["chr1"])
        assert_segments(got, want.chrom, want.start, want.end, want.value, exact=(op not in MOMENT_OPS), what=op)
    want = port.setcompare("wilcoxon", o[:120], o[120:], ["chr1"])
    got = engine.setcompare("wilcoxon", g[:120], g[120:], ["chr1"])
    assert_segments(got, want.chrom, want.start, want.end, want.value, what="wilcoxon")
    want = port.setcompare("ttest", o[:120], o[120:], ["chr1"])
    got = engine.setcompare("ttest", g[:120], g[120:], ["chr1"])
    assert_segments(got, want.chrom, want.start, want.end, want.value, what="ttest")


def test_mirror_protocol(engine, port):
    """The Python mirror of newMultiplexer / MeanReduction / pop()."""
    tr = make_tracks(4, [3000], seed=5, max_len=30, max_gap=30)
    o, g = both(tr)
    want = port.reduce("mean", o, ["chr1"])
    wi = api.MeanReduction(api.Multiplexer(g, names=["chr1"]), engine=engine)
    k = 0
    while not wi.done:
        assert (wi.chrom, wi.start, wi.finish, wi.value) == ("chr1", want.start[k], want.end[k], want.value[k])
        wi.pop()
        k += 1
    assert k == len(want)


def test_full_scale_properties(engine):
    """Size-independent properties at a size the oracle cannot check quickly:
    linearity (sum(x,x) == 2*x on x's own segmentation), mean*N == sum, segment
    tiling, and idempotence of mean over identical tracks."""
    tr = make_tracks(6, [3_000_000], seed=123)
    g = [api.Track(t["chrom"], t["start"], t["end"], t["value"]) for t in tr]
    s = engine.reduce("sum", g, ["chr1"])
    m = engine.reduce("mean", g, ["chr1"])
    assert np.array_equal(s.start, m.start) and np.array_equal(s.end, m.end)
    assert np.array_equal(m.value * 6, s.value)
    assert (s.end > s.start).all() and (s.start[1:] >= s.end[:-1]).all()
    one = engine.reduce("sum", [g[0], g[0]], ["chr1"])
    assert np.array_equal(one.start, g[0].start) and np.array_equal(one.value, 2 * g[0].value)
    same = engine.reduce("mean", [g[1]] * 5, ["chr1"])
    assert np.array_equal(same.value, g[1].value)
    mx, mn = engine.reduce("max", g, ["chr1"]), engine.reduce("min", g, ["chr1"])
    assert (mx.value >= mn.value).all() and (mx.value * 6 >= s.value).all()
    # AUC of the sum == sum of the tracks' own AUCs (integers here: exact)
    total = sum(float(((t.end - t.start) * t.value).sum()) for t in g)
    assert math.isclose(s.auc, total, rel_tol=1e-12)


# ---------------------------------------------------------------------------
# the drop-in boundary: the reference's own API exported by the GPU library
# ---------------------------------------------------------------------------
def test_dropin_api_matches_oracle(port):
    """newMultiplexer / <op>Reduction / newMultiset / TTest / MWU / AUCIntegrator /
    PearsonIntegrator of libwiggletools_b200.so, driven by the same C harness
    that drives the compiled reference (oracle/ref_harness.c)."""
    import os
    from oracle.pyoracle import Oracle
    if not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libwtdropin.so")):
        pytest.skip("oracle/_ref/libwtdropin.so not built")
    drop = Oracle("dropin")
    for seed, n in ((0, 3), (1, 7), (3, 12)):
        kw = dict(nan_prob=0.02 if seed % 2 else 0.0, align=(1 if seed < 3 else 10))
        tr = make_tracks(n, [20000, 6000, 9000], seed=300 + seed, max_len=40, max_gap=30, **kw)
        defaults = None if seed < 1 else [float(i % 3) for i in range(n)]
        o, _ = both(tr, defaults)
        for op in REDUCE_OPS:
            for strict in (False, True):
                want, got = port.reduce(op, o, NAMES, strict), drop.reduce(op, o, NAMES, strict)
                assert_segments(got, want.chrom, want.start, want.end, want.value, exact=(op not in MOMENT_OPS), what="dropin %s" % op)
                assert same_float(got.default, want.default), op
        if n >= 6:
            for op in ("ttest", "wilcoxon"):
                want, got = port.setcompare(op, o[:3], o[3:], NAMES), drop.setcompare(op, o[:3], o[3:], NAMES)
                assert_segments(got, want.chrom, want.start, want.end, want.value, what="dropin " + op)
        want, got = port.pearson(o[0], o[1], NAMES), drop.pearson(o[0], o[1], NAMES)
        assert same_float(got, want) or math.isclose(got, want, rel_tol=1e-9)
        want, got = port.auc("mean", o, NAMES), drop.auc("mean", o, NAMES)
        assert same_float(got, want) or math.isclose(got, want, rel_tol=RTOL)


def test_unsorted_input_is_refused(engine):
    """include/wiggletools_b200.h promises WB200_EINVAL for input that breaks the contract (sorted,
    disjoint, non-empty intervals per track): the cursor kernel checks every interval it reads
    (such input silently produced wrong segments)."""
    good = api.Track(np.zeros(4, np.int32), np.array([1, 10, 20, 30], np.int32), np.array([5, 15, 25, 35], np.int32), np.ones(4))
    other = api.Track(np.zeros(2, np.int32), np.array([3, 40], np.int32), np.array([8, 44], np.int32), np.ones(2))
    assert len(engine.reduce("sum", [good, other], ["chr1"])) > 0
    cases = {
        "unsorted": (np.array([1, 20, 10, 30], np.int32), np.array([5, 25, 15, 35], np.int32)),
        "overlap": (np.array([1, 10, 14, 30], np.int32), np.array([5, 15, 25, 35], np.int32)),
        "empty": (np.array([1, 10, 20, 30], np.int32), np.array([5, 10, 25, 35], np.int32)),
    }
    for name, (s_, e_) in cases.items():
        bad = api.Track(np.zeros(4, np.int32), s_, e_, np.ones(4))
        for op in ("sum", "median", "var"):
            with pytest.raises(api.WB200Error, match="not sorted|empty interval"):
                engine.reduce(op, [bad, other], ["chr1"])
    # the engine is still usable afterwards
    assert len(engine.reduce("sum", [good, other], ["chr1"])) > 0


def test_device_tdist_on_committed_table(engine):
    """The t-test p-value is 2 * Q(t, nu) (src/setComparisons.c:117, gsl_cdf_tdist_Q in the
    reference; GSL is absent here).  The DEVICE build of the engine's tdist_Q -- the function
    the t-test kernel calls -- is pinned on a committed mpmath table (tests/golden/tdist_table.json),
    not on the oracle's copy of the same continued fraction."""
    import ctypes as C
    import json
    import os
    rows = np.array(json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tdist_table.json")))["rows"])
    t, nu, p = (np.ascontiguousarray(rows[:, k]) for k in range(3))
    out = np.empty(len(t))
    engine._check(engine.lib.wb200_debug_tdist_device(engine.ctx, len(t), t.ctypes.data, nu.ctypes.data, out.ctypes.data))
    got = 2 * out
    rel = np.abs(got - p) / np.maximum(p, 1e-300)
    assert rel.max() <= 1e-10, (rel.max(), t[rel.argmax()], nu[rel.argmax()], got[rel.argmax()], p[rel.argmax()])
    # the host build of the same source function (what CPU-side tests see) agrees with the device build
    # up to the two libms' last bits through the continued fraction
    host = np.array([engine.lib.wb200_tdist_Q(float(a), float(b)) for a, b in zip(t, nu)])
    assert np.all(np.abs(host - out) <= 1e-11 * np.maximum(np.abs(host), 1e-300) + 1e-300)


def test_dropin_multiplexer_rows(port):
    """Raw multiplexer rows through the drop-in: newMultiplexer shows its first row, popMultiplexer
    serves chrom/start/finish + values[] + inplay[] per segment (what mWigWriter.c:179-225 and
    SelectReductionPop read) -- bit-identical to the reference's sweep (oracle port pinned on it,
    tests/test_oracle.py), strict and not, with defaults, NaNs and > 10 000 rows per chromosome
    (several row blocks)."""
    import os
    from oracle.pyoracle import Oracle
    if not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libwtdropin.so")):
        pytest.skip("oracle/_ref/libwtdropin.so not built")
    drop = Oracle("dropin")
    for seed, n, lens in ((0, 2, [3000, 800]), (1, 5, [60000, 9000, 30000]), (2, 9, [20000, 500, 7000])):
        kw = dict(nan_prob=0.03 if seed == 1 else 0.0, align=(1 if seed < 2 else 10))
        tr = make_tracks(n, lens, seed=700 + seed, max_len=40, max_gap=30, **kw)
        defaults = None if seed == 0 else [float(i % 3) - 1.0 for i in range(n)]
        o, _ = both(tr, defaults)
        names = NAMES[:len(lens)]
        for strict in (False, True):
            (ws, wv, wi), (gs, gv, gi) = port.rows(o, names, strict), drop.rows(o, names, strict)
            assert len(gs) == len(ws), (seed, strict, len(gs), len(ws))
            assert np.array_equal(gs.chrom, ws.chrom) and np.array_equal(gs.start, ws.start) and np.array_equal(gs.end, ws.end)
            assert np.array_equal(gi, wi), (seed, strict)
            assert np.array_equal(gv, wv, equal_nan=True), (seed, strict)
    # a reduction built on a strict multiplexer takes over the chromosome staged for its first row
    tr = make_tracks(4, [5000, 2000], seed=731, max_len=30, max_gap=40)
    o, _ = both(tr)
    want, got = port.reduce("mean", o, NAMES[:2], True), drop.reduce("mean", o, NAMES[:2], True)
    assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="strict mean over a staged multiplexer")


def test_dropin_cli_multi_writer(tmp_path):
    """`mwrite_bg`, `mwrite` and `print select`-style consumers of a raw Multiplexer through the
    reference's unmodified mWigWriter.c linked against the library: byte for byte what the
    unmodified reference CLI prints (test/test.py does the same for mwrite_bg on its fixtures)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(__file__), "..", "integration", "_build", "wiggletools_b200")
    ref = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "wiggletools")
    if not (os.path.exists(exe) and os.path.exists(ref)):
        pytest.skip("drop-in / reference CLI not built (needs /root/reference at build time)")
    (tmp_path / "fixedStep.wig").write_text("fixedStep chrom=chr1 start=1 step=1 span=1\n" + "".join("%d\n" % i for i in range(10)))
    (tmp_path / "variableStep.wig").write_text("variableStep chrom=chr1 span=1\n" + "".join("%d\t%d\n" % (p, x) for p, x in zip((1, 2, 4, 6, 8), (1, 2, 3, 4, 5))))
    (tmp_path / "overlapping.bed").write_text("chr1\t2\t5\nchr1\t4\t7\nchr2\t1\t4\n")
    rng = np.random.default_rng(5)
    for t in range(3):                                   # three bedGraph tracks, ~25 000 rows in the union: several row blocks
        pos, lines = 0, []
        for chrom in ("chr1", "chr2"):
            pos = 0
            for _ in range(6000):
                pos += int(rng.integers(0, 5)); ln = int(rng.integers(1, 9))
                lines.append("%s\t%d\t%d\t%g\n" % (chrom, pos, pos + ln, rng.integers(0, 50) / 4.0))
                pos += ln
        (tmp_path / ("t%d.bg" % t)).write_text("".join(lines))
    progs = [["mwrite_bg", "-", "fixedStep.wig", "variableStep.wig", "overlapping.bed"],
             ["mwrite", "-", "fixedStep.wig", "variableStep.wig"],
             ["mwrite_bg", "-", "strict", "fixedStep.wig", "variableStep.wig"],
             ["mwrite_bg", "-", "t0.bg", "t1.bg", "t2.bg"],
             ["mwrite_bg", "-", "strict", "t0.bg", "t1.bg", "t2.bg"],
             ["mwrite_bg", "-", "map", "scale", "2", "t0.bg", "t1.bg"]]
    for prog in progs:
        want = subprocess.run([ref] + prog, capture_output=True, text=True, check=True, cwd=tmp_path).stdout
        got = subprocess.run([exe] + prog, capture_output=True, text=True, check=True, cwd=tmp_path).stdout
        assert got == want, prog
        assert len(want) > 0, prog


def test_dropin_cli_config1(tmp_path):
    """BASELINE.json configs[0]: `wiggletools mean fixedStep variableStep` through
    the reference's unmodified parser/readers/writer linked against the GPU
    library (integration/_build/wiggletools_b200).  The .wig twins of the .bw fixtures
    are written here from their values (test/fixedStep.wig, test/variableStep.wig)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(__file__), "..", "integration", "_build", "wiggletools_b200")
    if not os.path.exists(exe):
        pytest.skip("integration/_build/wiggletools_b200 not built (needs /root/reference at build time)")
    f, v = tmp_path / "fixedStep.wig", tmp_path / "variableStep.wig"
    f.write_text("fixedStep chrom=chr1 start=1 step=1 span=1\n" + "".join("%d\n" % i for i in range(10)))
    v.write_text("variableStep chrom=chr1 span=1\n" + "".join("%d\t%d\n" % (p, x) for p, x in zip((1, 2, 4, 6, 8), (1, 2, 3, 4, 5))))

    def run(*args):
        return subprocess.run([exe] + list(args), capture_output=True, text=True, check=True, cwd=tmp_path).stdout

    out = run("mean", "fixedStep.wig", "variableStep.wig").split("\n")
    assert out[0] == "fixedStep chrom=chr1 start=1 step=1"
    assert [float(x) for x in out[1:11]] == [.5, 1.5, 1, 3, 2, 4.5, 3, 6, 4, 4.5]
    bg = run("write_bg", "-", "var", "fixedStep.wig", "variableStep.wig").strip().split("\n")
    assert [float(l.split("\t")[3]) for l in bg] == [.25, .25, .5, 0, 2, .25, 4.5, 1, 8, 10.125]
    assert run("AUC", "mean", "fixedStep.wig", "variableStep.wig").strip() == "30.000000"
    assert run("pearson", "fixedStep.wig", "variableStep.wig").strip() == "-0.028968"   # test/expected/pearson.txt
    # BASELINE.json configs[0], literally: the BigWig twins through the library's own
    # decoder (BigWiggleReader of the drop-in), test/test.py:28,52,76 style identities
    import shutil
    for f in ("fixedStep.bw", "variableStep.bw"):
        shutil.copy(os.path.join(os.path.dirname(__file__), "golden", f), tmp_path / f)
    out = run("mean", "fixedStep.bw", "variableStep.bw").split("\n")
    assert out[0] == "fixedStep chrom=chr1 start=1 step=1"
    assert [float(x) for x in out[1:11]] == [.5, 1.5, 1, 3, 2, 4.5, 3, 6, 4, 4.5]
    # ... and byte for byte what the REFERENCE prints for the same command line (its own BigWig reader
    # compiled over oracle/bigwig_shim.c, its own multiplexer, reducer and writer)
    ref_bw = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "wiggletools_bw")
    if os.path.exists(ref_bw):
        for prog in (["mean", "fixedStep.bw", "variableStep.bw"], ["write_bg", "-", "var", "fixedStep.bw", "variableStep.bw"],
                     ["pearson", "fixedStep.bw", "variableStep.bw"]):
            want = subprocess.run([ref_bw] + prog, capture_output=True, text=True, check=True, cwd=tmp_path).stdout
            assert run(*prog) == want, prog
    for args in (["do", "isZero", "diff", "fixedStep.bw", "fixedStep.wig"],
                 ["do", "isZero", "diff", "variableStep.bw", "variableStep.wig"],
                 ["do", "isZero", "diff", "sum", "fixedStep.bw", "fixedStep.bw", ":", "scale", "2", "fixedStep.bw"],
                 ["do", "isZero", "seek", "chr1", "2", "6", "diff", "variableStep.bw", "variableStep.wig"]):
        assert subprocess.run([exe] + args, cwd=tmp_path).returncode == 0, args
    assert subprocess.run([exe, "do", "isZero", "diff", "fixedStep.bw", "variableStep.wig"], cwd=tmp_path).returncode == 1  # negative control, test/test.py:22
    # seek below an integrator / at the top: the reduction's integral restarts with the sought region
    # (the eager first pop had already added the whole first chromosome) -- same text as the
    # unmodified reference CLI
    ref_cli = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "wiggletools")
    if os.path.exists(ref_cli):
        (tmp_path / "g.wig").write_text("fixedStep chrom=chr1 start=3 step=1 span=2\n5\n")
        for prog in (["AUC", "seek", "chr1", "3", "8", "mean", "fixedStep.wig", "fixedStep.wig"],
                     ["AUC", "seek", "chr1", "3", "8", "sum", "fixedStep.wig", "g.wig"],
                     ["AUC", "seek", "chr1", "1", "4", "sum", "fixedStep.wig", "variableStep.wig"],
                     ["seek", "chr1", "3", "8", "sum", "fixedStep.wig", "g.wig"],
                     ["AUC", "mean", "fixedStep.wig", "variableStep.wig"]):
            want = subprocess.run([ref_cli] + prog, capture_output=True, text=True, check=True, cwd=tmp_path).stdout
            assert run(*prog) == want, prog
    # Where a track is NOT in play at the start of the sought region the reference shows quirk Q10
    # (seekCoreMultiplexer keeps stale values[], multiplexer.c:130-141): it prints 20.5 here, counting
    # variableStep's value from before the region at position 3.  The engine does not reproduce that
    # staleness: the integral of the mean over [3, 9) is 19.5.
    assert run("AUC", "seek", "chr1", "3", "8", "mean", "fixedStep.wig", "variableStep.wig").strip() == "19.500000"
    # sum x x == scale 2 x (test/test.py:76), via the reference's own isZero/diff chain
    assert subprocess.run([exe, "do", "isZero", "diff", "sum", "fixedStep.wig", "fixedStep.wig", ":", "scale", "2", "fixedStep.wig"],
                          cwd=tmp_path).returncode == 0


def test_fast_path_overflow_reruns_as_rows(engine, port):
    """More unrepresentable segments than the fix-up list holds: the chromosome is
    re-run on the row path; results still match the oracle exactly."""
    tr = make_tracks(3, [2_000_000], seed=41, max_len=12, max_gap=6)
    for t in tr:
        t["value"][::2] = np.nan            # half of all intervals poison their segments
    o, g = both(tr)
    want = port.reduce("mean", o, ["chr1"])
    got = engine.reduce("mean", g, ["chr1"])
    assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="nan-heavy mean")
    assert engine.last_stats[-1]["fast_path"] == 0 and engine.last_stats[-1]["slow_segments"] > 65536


def test_long_intervals_and_sparse_windows(engine, port):
    """Intervals spanning hundreds of windows, windows with no breakpoint at all,
    and a track that is one single interval."""
    L = 3_000_000
    tr = make_tracks(5, [L], seed=52, max_len=50, max_gap=200000)      # very sparse
    tr.append(dict(chrom=np.zeros(1, np.int32), start=np.array([7], np.int32), end=np.array([L - 3], np.int32), value=np.array([0.5])))
    tr.append(dict(chrom=np.zeros(3, np.int32), start=np.array([1, 1_000_000, 2_999_000], np.int32),
                   end=np.array([999_999, 2_998_000, L + 1], np.int32), value=np.array([1.0, 2.0, 4.0])))
    o, g = both(tr)
    for op in ("sum", "mean", "max", "stddev"):
        for strict in (False, True):
            want = port.reduce(op, o, ["chr1"], strict)
            got = engine.reduce(op, g, ["chr1"], strict=strict)
            assert_segments(got, want.chrom, want.start, want.end, want.value, exact=(op not in MOMENT_OPS), what="sparse %s" % op)


def test_dense_every_base_a_breakpoint(engine, port):
    """Unit-length abutting intervals on many tracks: every base pair is a breakpoint
    (the regime of the 512-track configuration)."""
    n, L = 40, 30000
    rng = np.random.default_rng(3)
    tr = []
    for i in range(n):
        s = np.arange(1 + i % 3, L, 1 + i % 2, dtype=np.int32)
        tr.append(dict(chrom=np.zeros(len(s), np.int32), start=s, end=s + 1 + i % 2, value=rng.integers(0, 999, len(s)) / 8.0))
    o, g = both(tr)
    for op in ("mean", "min", "var"):
        want = port.reduce(op, o, ["chr1"])
        got = engine.reduce(op, g, ["chr1"])
        assert_segments(got, want.chrom, want.start, want.end, want.value, exact=(op not in MOMENT_OPS), what="dense %s" % op)
    st = engine.last_stats[-1]
    assert st["segments"] >= L - 5


def test_fillin(engine, port):
    """fillIn / trimFill (FillInReduction, src/reducers.c:78-119) through the engine and the drop-in API."""
    import os
    from oracle.pyoracle import Oracle
    tr = make_tracks(2, [30000, 8000], seed=77, max_len=40, max_gap=60, nan_prob=0.01)
    o, g = both(tr, [0.0, 2.5])
    for op in ("fillIn", "trimFill"):
        want = port.reduce(op, o, NAMES)
        got = engine.reduce(op, g, NAMES)
        assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what=op)
        assert same_float(got.default, want.default)
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libwtdropin.so")):
        drop = Oracle("dropin")
        for op in ("fillIn", "trimFill"):
            want, got = port.reduce(op, o, NAMES), drop.reduce(op, o, NAMES)
            assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="dropin " + op)
    with pytest.raises(api.WB200Error):
        engine.reduce("fillIn", g + g, NAMES)


def test_moments_fast_path_is_used_and_tight(engine, port):
    """var / stddev / CV on float32-grade data run on the exact-moments path (no row
    loop) and stay within a few ulps of the reference; data that does not fit the
    format falls back to the row path per segment and stays exact."""
    tr = make_tracks(48, [400000], seed=19)
    o, g = both(tr)
    for op in ("var", "stddev", "cv"):
        want = port.reduce(op, o, ["chr1"])
        got = engine.reduce(op, g, ["chr1"])
        assert engine.last_stats[-1]["fast_path"] == 2 and engine.last_stats[-1]["slow_segments"] == 0
        assert_segments(got, want.chrom, want.start, want.end, want.value, rtol=1e-13, what=op)
    # full-mantissa doubles: not float-representable -> every segment recomputed on the row path
    rng = np.random.default_rng(1)
    for t in tr[:6]:
        t["value"] = rng.random(len(t["value"])) * 100
    o, g = both(tr[:6])
    want = port.reduce("stddev", o, ["chr1"])
    got = engine.reduce("stddev", g, ["chr1"])
    assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="stddev doubles")


def test_randomised_differential(port):
    """Many small random configurations (track counts, densities, window sizes,
    defaults, NaNs, strictness, every statistic) against the oracle."""
    rng = np.random.default_rng(20260925)
    engines = {w: api.Engine(0, window=w) for w in (1024, 2048, 4096)}
    names = ["chr1", "chr10", "chr2"]
    ops = REDUCE_OPS + ["fillIn", "trimFill"]
    for case in range(300):
        n = int(rng.integers(1, 41))
        lengths = [int(rng.integers(50, 40000)) for _ in range(int(rng.integers(1, 4)))]
        kw = dict(max_len=int(rng.integers(1, 400)), max_gap=int(rng.integers(1, 3000)),
                  gap_prob=float(rng.choice([0.0, 0.3, 0.9])), vmax=int(rng.choice([2, 50, 100000])),
                  align=int(rng.choice([1, 1, 7, 64])), nan_prob=float(rng.choice([0.0, 0.0, 0.05])),
                  neg=bool(rng.integers(0, 2)))
        tr = make_tracks(n, lengths, seed=int(rng.integers(0, 1 << 30)), **kw)
        defaults = None if rng.random() < 0.6 else [float(rng.choice([0.0, 1.0, -2.5, 0.125])) for _ in range(n)]
        o, g = both(tr, defaults)
        eng = engines[int(rng.choice([1024, 2048, 4096]))]
        for op in rng.choice(ops, size=3, replace=False):
            op = str(op)
            strict = bool(rng.integers(0, 2))
            if op in ("fillIn", "trimFill"):
                if n < 2:
                    continue
                oo, gg, strict = o[:2], g[:2], False
            else:
                oo, gg = o, g
            want = port.reduce(op, oo, names[:len(lengths)], strict)
            got = eng.reduce(op, gg, names[:len(lengths)], strict=strict)
            assert_segments(got, want.chrom, want.start, want.end, want.value, exact=(op not in MOMENT_OPS),
                            what="case %d %s n=%d strict=%d %s" % (case, op, n, strict, kw))
            assert same_float(got.default, want.default)
        if n >= 6 and case % 4 == 0:
            k = int(rng.integers(3, n - 2))
            for op in ("ttest", "wilcoxon"):
                want = port.setcompare(op, o[:k], o[k:], names[:len(lengths)])
                got = eng.setcompare(op, g[:k], g[k:], names[:len(lengths)])
                assert_segments(got, want.chrom, want.start, want.end, want.value, what="case %d %s" % (case, op))
    for e in engines.values():
        e.close()


def test_ndpearson(engine, port):
    """NDPearsonIntegrator through the engine and through the drop-in API."""
    import os
    from oracle.pyoracle import Oracle
    tr = make_tracks(8, [40000, 9000], seed=88, max_len=60, max_gap=400)
    o, g = both(tr, [float(i % 2) for i in range(8)])
    want = port.ndpearson(o[:4], o[4:], NAMES)
    got = engine.ndpearson(g[:4], g[4:], NAMES)
    assert math.isclose(got, want, rel_tol=1e-9), (got, want)
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libwtdropin.so")):
        got = Oracle("dropin").ndpearson(o[:4], o[4:], NAMES)
        assert math.isclose(got, want, rel_tol=1e-9), (got, want)
