"""Generate tests/golden/ref_cases.npz from the UNMODIFIED reference compiled
into oracle/_ref/libwtref.so (run here, in the build container, where
/root/reference exists):   python tests/golden/make_golden.py

Each case stores its inputs (so the fixture does not depend on the generator)
and the reference's output for every statistic of the hot path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, Track  # noqa: E402
from wiggletools_b200.synth import make_tracks  # noqa: E402

NAMES = ["chr1", "chr10", "chr2", "chrX"]
OPS = ["sum", "mult", "mean", "var", "stddev", "entropy", "cv", "min", "max", "median"]

CASES = [
    # name, N, lengths, kwargs, defaults
    ("plain8", 8, [4000, 1500, 2500], dict(max_len=40, max_gap=30), None),
    ("nan5", 5, [3000, 800], dict(max_len=25, max_gap=60, nan_prob=0.03), None),
    ("neg7_defaults", 7, [2500, 2500, 100], dict(max_len=30, max_gap=10, neg=True), [0.0, 1.5, -2.0, 0.25, 0.0, 3.0, 1.0]),
    ("binned16", 16, [6000, 3000], dict(max_len=8, max_gap=4, align=25), None),
    ("ties9", 9, [5000], dict(max_len=60, max_gap=20, vmax=4), None),
    ("single1", 1, [2000, 500], dict(max_len=50, max_gap=50), None),
    ("dense40", 40, [1200], dict(max_len=20, max_gap=6), None),
]


def main():
    ref = Oracle("reference")
    out = {}
    for name, n, lengths, kw, defaults in CASES:
        tr = make_tracks(n, lengths, seed=len(name) * 7 + n, **kw)
        tracks = [Track(t["chrom"], t["start"], t["end"], t["value"],
                        default=(defaults[i] if defaults else 0.0)) for i, t in enumerate(tr)]
        out[name + "/n"] = np.int32(n)
        for i, t in enumerate(tracks):
            for f in ("chrom", "start", "end", "value"):
                out["%s/t%d/%s" % (name, i, f)] = getattr(t, f)
            out["%s/t%d/default" % (name, i)] = np.float64(t.default)
        for op in OPS:
            for strict in (0, 1):
                r = ref.reduce(op, tracks, NAMES, bool(strict))
                key = "%s/%s/%d" % (name, op, strict)
                out[key + "/chrom"], out[key + "/start"], out[key + "/end"], out[key + "/value"] = r.chrom, r.start, r.end, r.value
                out[key + "/default"] = np.float64(r.default)
        out[name + "/auc_mean"] = np.float64(ref.auc("mean", tracks, NAMES))
        if n >= 2:
            out[name + "/pearson01"] = np.float64(ref.pearson(tracks[0], tracks[1], NAMES))
        if n >= 6:
            for op in ("ttest", "wilcoxon"):
                r = ref.setcompare(op, tracks[:3], tracks[3:], NAMES)
                key = "%s/%s" % (name, op)
                out[key + "/chrom"], out[key + "/start"], out[key + "/end"], out[key + "/value"] = r.chrom, r.start, r.end, r.value
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
