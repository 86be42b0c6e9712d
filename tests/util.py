"""Shared helpers for the parity tests."""
import os

import numpy as np

from oracle.pyoracle import Track

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_NAMES = ["chr1", "chr10", "chr2", "chrX"]
REDUCE_OPS = ["sum", "mult", "mean", "var", "stddev", "entropy", "cv", "min", "max", "median"]

# fp64 reductions must agree with the reference within this RELATIVE tolerance
# (BASELINE.json north_star); coordinates must agree bit for bit.
RTOL = 1e-10


def load_golden():
    z = np.load(os.path.join(HERE, "golden", "ref_cases.npz"))
    cases = sorted({k.split("/")[0] for k in z.files})
    return z, cases


def golden_tracks(z, case, cls=Track):
    n = int(z[case + "/n"])
    return [cls(z["%s/t%d/chrom" % (case, i)], z["%s/t%d/start" % (case, i)], z["%s/t%d/end" % (case, i)],
                z["%s/t%d/value" % (case, i)], default=float(z["%s/t%d/default" % (case, i)])) for i in range(n)]


def same_float(a, b):
    return (a == b) or (np.isnan(a) and np.isnan(b))


def assert_segments(got, want_chrom, want_start, want_end, want_value, rtol=RTOL, exact=False, what=""):
    assert len(got.start) == len(want_start), "%s: %d segments, expected %d" % (what, len(got.start), len(want_start))
    assert np.array_equal(got.chrom, want_chrom), what + ": chrom differs"
    assert np.array_equal(got.start, want_start), what + ": start differs"
    assert np.array_equal(got.end, want_end), what + ": end differs"
    a, b = np.asarray(got.value), np.asarray(want_value)
    assert np.array_equal(np.isnan(a), np.isnan(b)), what + ": NaN pattern differs"
    if exact:
        assert np.array_equal(a, b, equal_nan=True), what + ": values not bit-identical (max |d| %g)" % np.nanmax(np.abs(a - b))
    else:
        ok = np.isnan(a) | (a == b) | (np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)))
        assert ok.all(), "%s: %d values beyond rtol %g (worst |d| %g)" % (what, (~ok).sum(), rtol, np.nanmax(np.abs(a - b)))
