"""CPU tests: pin the oracle (C restatement) against (1) the reference's own
golden vectors, (2) outputs captured from the compiled reference (SURVEY.md
section 8(c)), (3) the committed fixtures generated from the compiled reference,
(4) the compiled reference itself when oracle/_ref was built here."""
import math

import numpy as np
import pytest

from oracle.pyoracle import Track
from wiggletools_b200.synth import make_tracks
from util import GOLDEN_NAMES, REDUCE_OPS, assert_segments, golden_tracks, load_golden, same_float

NAMES = ["chr1", "chr2"]


def fixed_step():
    # /root/reference/test/fixedStep.wig: values 0..9 at 1-based positions 1..10
    p = np.arange(1, 11)
    return Track(np.zeros(10), p, p + 1, np.arange(10.0))


def variable_step():
    # /root/reference/test/variableStep.wig: 1,2,3,4,5 at positions 1,2,4,6,8
    p = np.array([1, 2, 4, 6, 8])
    return Track(np.zeros(5), p, p + 1, np.arange(1.0, 6.0))


def overlapping_coverage():
    # /root/reference/test/overlapping_coverage.wig (coverage of overlapping.bed)
    c = np.array([0] * 6 + [1] * 3)
    s = np.array([3, 4, 5, 6, 7, 8, 2, 3, 4])
    v = np.array([1, 2, 2, 2, 1, 1, 1, 1, 1.0])
    return Track(c, s, s + 1, v)


# outputs of `write_bg - <op> fixedStep.wig variableStep.wig` from the compiled
# reference; one value per position 1..10
CAPTURED = {
    "sum": [1, 3, 2, 6, 4, 9, 6, 12, 8, 9],
    "mean": [.5, 1.5, 1, 3, 2, 4.5, 3, 6, 4, 4.5],
    "var": [.25, .25, .5, 0, 2, .25, 4.5, 1, 8, 10.125],
    "stddev": [.5, .5, 1, 0, 2, .5, 3, 1, 4, 4.5],
    "entropy": [.5, .5, 1, 0, 2, .5, 3, 1, 4, 4.5],
    "median": [1, 2, 2, 3, 4, 5, 6, 7, 8, 9],
    "max": [1, 2, 2, 3, 4, 5, 6, 7, 8, 9],
    "min": [0, 1, 0, 3, 0, 4, 0, 5, 0, 0],
    "cv": [1, 1 / 3, 1, 0, 1, 1 / 9, 1, 1 / 6, 1, 1],
    "mult": [0, 2, 0, 9, 0, 20, 0, 35, 0, 0],
}


@pytest.mark.parametrize("op", sorted(CAPTURED))
def test_captured_reference_outputs(port, op):
    r = port.reduce(op, [fixed_step(), variable_step()], NAMES)
    assert list(r.start) == list(range(1, 11)) and list(r.end) == list(range(2, 12))
    np.testing.assert_allclose(r.value, CAPTURED[op], rtol=1e-15, atol=0)


def test_captured_strict_and_wilcoxon(port):
    r = port.reduce("sum", [fixed_step(), variable_step()], NAMES, strict=True)
    assert list(r.start) == [1, 2, 4, 6, 8] and list(r.value) == [1, 3, 6, 9, 12]
    f, v, oc = fixed_step(), variable_step(), overlapping_coverage()
    r = port.setcompare("wilcoxon", [f, v, f], [f, v, oc], NAMES)
    want = [-1.314437, -0.496341, -0.496341, 0, -0.496341, -1.314437, -0.496341, -1.314437, -1.588194, -1.588194]
    np.testing.assert_allclose(r.value, want, atol=5e-7)


def test_reference_golden_pearson(port):
    # /root/reference/test/expected/pearson.txt, produced by test/test.py:104
    assert "%f" % port.pearson(fixed_step(), variable_step(), NAMES) == "-0.028968"


def test_reference_golden_auc(port):
    # test/test.py:95-98: AUC lt/lte/gte/gt 4 fixedStep.wig == 4/5/6/5; the
    # comparison filters are unary ops outside the path, so apply them here and
    # push the 0/1 track through the reduction + AUC integrator.
    f = fixed_step()
    for pred, want in ((f.value < 4, 4), (f.value <= 4, 5), (f.value >= 4, 6), (f.value > 4, 5)):
        t = Track(f.chrom, f.start, f.end, pred.astype(float))
        assert port.auc("sum", [t], NAMES) == want
    assert port.auc("mean", [fixed_step(), variable_step()], NAMES) == 30


def test_reference_identity_sum_scale(port):
    # test/test.py:76: sum x x == scale 2 x
    f = fixed_step()
    r = port.reduce("sum", [f, f], NAMES)
    assert np.array_equal(r.value, 2 * f.value) and np.array_equal(r.start, f.start)


def test_committed_golden_fixtures(port):
    z, cases = load_golden()
    for case in cases:
        tracks = golden_tracks(z, case)
        for op in REDUCE_OPS:
            for strict in (0, 1):
                key = "%s/%s/%d" % (case, op, strict)
                r = port.reduce(op, tracks, GOLDEN_NAMES, bool(strict))
                assert_segments(r, z[key + "/chrom"], z[key + "/start"], z[key + "/end"], z[key + "/value"], exact=True, what=key)
                assert same_float(r.default, float(z[key + "/default"])), key
        assert port.auc("mean", tracks, GOLDEN_NAMES) == float(z[case + "/auc_mean"])
        if len(tracks) >= 2:
            assert same_float(port.pearson(tracks[0], tracks[1], GOLDEN_NAMES), float(z[case + "/pearson01"]))
        if len(tracks) >= 6:
            for op in ("ttest", "wilcoxon"):
                key = "%s/%s" % (case, op)
                r = port.setcompare(op, tracks[:3], tracks[3:], GOLDEN_NAMES)
                assert_segments(r, z[key + "/chrom"], z[key + "/start"], z[key + "/end"], z[key + "/value"], exact=True, what=key)


@pytest.mark.parametrize("seed", range(4))
def test_port_equals_compiled_reference(port, reference, seed):
    names = ["chr1", "chr10", "chr2"]
    for n in (1, 2, 5, 12):
        kw = dict(nan_prob=0.02 if seed % 2 else 0.0, neg=(seed == 2), align=(1 if seed < 3 else 10))
        tr = make_tracks(n, [2500, 900, 1800], seed=100 + seed, max_len=40, max_gap=30, **kw)
        tracks = [Track(t["chrom"], t["start"], t["end"], t["value"], default=(0.0 if seed < 2 else float(i % 3)))
                  for i, t in enumerate(tr)]
        for op in REDUCE_OPS:
            for strict in (False, True):
                a, b = port.reduce(op, tracks, names, strict), reference.reduce(op, tracks, names, strict)
                assert_segments(a, b.chrom, b.start, b.end, b.value, exact=True, what="%s n=%d" % (op, n))
                assert same_float(a.default, b.default)
        if n >= 6:
            for op in ("ttest", "wilcoxon"):
                a, b = port.setcompare(op, tracks[:3], tracks[3:], names), reference.setcompare(op, tracks[:3], tracks[3:], names)
                assert_segments(a, b.chrom, b.start, b.end, b.value, exact=True, what=op)
        if n >= 2:
            assert same_float(port.pearson(tracks[0], tracks[1], names), reference.pearson(tracks[0], tracks[1], names))
        assert port.auc("mean", tracks, names) == reference.auc("mean", tracks, names)
        if n >= 6:
            assert same_float(port.ndpearson(tracks[:3], tracks[3:6], names), reference.ndpearson(tracks[:3], tracks[3:6], names))
        if n >= 2:
            for op in ("fillIn", "trimFill"):
                a, b = port.reduce(op, tracks[:2], names), reference.reduce(op, tracks[:2], names)
                assert_segments(a, b.chrom, b.start, b.end, b.value, exact=True, what=op)
        ra, rb = port.rows(tracks, names), reference.rows(tracks, names)
        assert np.array_equal(ra[1], rb[1], equal_nan=True) and np.array_equal(ra[2], rb[2])


def _tdist_Q_exact(t, nu):
    import mpmath
    mpmath.mp.dps = 40
    t, nu = mpmath.mpf(t), mpmath.mpf(nu)
    return float(mpmath.betainc(nu / 2, mpmath.mpf(1) / 2, 0, nu / (nu + t * t), regularized=True) / 2)


def test_tdist_closed_forms(port):
    # the GSL stand-in (parity unpinned vs GSL): nu=1 Cauchy, nu=2
    for t in (0.0, 0.3, 1.0, 2.5, 7.0, 40.0):
        assert math.isclose(port.tdist_Q(t, 1.0), 0.5 - math.atan(t) / math.pi, rel_tol=1e-12, abs_tol=1e-15)
        assert math.isclose(port.tdist_Q(t, 2.0), 0.5 - t / (2 * math.sqrt(2 + t * t)), rel_tol=1e-11, abs_tol=1e-15)
    assert math.isclose(port.tdist_Q(-1.0, 3.0), 1 - port.tdist_Q(1.0, 3.0), rel_tol=1e-14)


def test_tdist_against_high_precision(port):
    pytest.importorskip("mpmath")
    for nu in (1.0, 2.5, 4.0, 7.3, 30.0, 254.0, 1022.0, 1e5):
        for t in (1e-3, 0.1, 0.7, 1.96, 3.3, 8.0, 25.0):
            assert math.isclose(port.tdist_Q(t, nu), _tdist_Q_exact(t, nu), rel_tol=5e-12, abs_tol=1e-300), (t, nu)


def test_tdist_oracle_on_committed_table(port):
    """The oracle's Student-t tail on the committed high-precision table (tests/golden/tdist_table.json,
    mpmath 50 digits, generator beside it): the same table the DEVICE function is pinned on in
    tests/test_gpu_parity.py, so the t-test p-value is never compared with a copy of itself."""
    import json
    import os
    rows = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tdist_table.json")))["rows"]
    assert len(rows) > 300
    for t, nu, p in rows:
        assert math.isclose(2 * port.tdist_Q(t, nu), p, rel_tol=1e-10, abs_tol=1e-300), (t, nu)
