import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wiggletools_b200 import api
from wiggletools_b200.synth import make_tracks
eng = api.Engine(0)
tr = make_tracks(12, [30000, 9000], seed=3, max_len=60, max_gap=40, nan_prob=0.01)
g = [api.Track(t["chrom"], t["start"], t["end"], t["value"]) for t in tr]
for op in ("mean", "var", "median", "max", "mult", "wilcoxon", "ttest"):
    if op in ("wilcoxon", "ttest"):
        r = eng.setcompare(op, g[:5], g[5:], ["chr1", "chr2"])
    else:
        r = eng.reduce(op, g, ["chr1", "chr2"])
    print(op, len(r), float(np.nansum(r.value)))
print("pearson", eng.pearson(g[0], g[1], ["chr1", "chr2"]))
# the sum / mean sweep kernel on a larger chromosome (many windows per CTA: tickets, look-back), one and two limbs
tr2 = make_tracks(6, [900000], seed=5)
g2 = [api.Track(t["chrom"], t["start"], t["end"], t["value"]) for t in tr2]
r = eng.reduce("mean", g2, ["chr1"]); print("mean (sweep kernel)", len(r), float(np.nansum(r.value)), eng.last_stats[-1]["sweep_kernel"], eng.last_stats[-1]["limbs"])
for t in g2: t.value[:] = t.value * 1.0000001
r = eng.reduce("sum", g2, ["chr1"]); print("sum (two limbs)", len(r), float(np.nansum(r.value)), eng.last_stats[-1]["sweep_kernel"], eng.last_stats[-1]["limbs"])
# the text sink: compress chains, writer blocks, both modes
rng = np.random.default_rng(1)
n = 30000
start = np.cumsum(rng.integers(1, 4, n)).astype(np.int32)
seg = api.Segments(np.zeros(n, np.int32), start, (start + rng.integers(1, 3, n)).astype(np.int32).clip(max=np.r_[start[1:], start[-1] + 2]),
                   np.round(rng.random(n), 1) + rng.integers(-3, 4, n) * 2e-7)
for bg in (False, True):
    print("sink", bg, len(eng.format_segments(seg, ["chr1"], bg)))
# element-wise maps: in-place map, a log that drops records (compaction + rebuilt offsets), nested chains
r = eng.reduce("mean", g, ["chr1", "chr2"], maps=[[("abs",), ("offset", 0.5), ("ln",)] if i % 2 else [("scale", -1.0), ("log", 2.0)] for i in range(len(g))])
print("maps", len(r), float(np.nansum(r.value)))
# round 2: the selection kernel (order statistics over more than 24 tracks: bucketed path, ranges of a window, the
# bitset and sorted-array fallbacks) and the moments variant of the sweep kernel
for n, kw, what in ((48, dict(), "f32"), (200, dict(), "f32"), (300, dict(max_len=6, max_gap=4), "dense ranges"), (160, dict(max_len=1, max_gap=1), "staging overflow")):
    trn = make_tracks(n, [5000], seed=7, vmax=64, **kw)
    gn = [api.Track(t["chrom"], t["start"], t["end"], t["value"]) for t in trn]
    for op in ("median", "max"):
        r = eng.reduce(op, gn, ["chr1"]); print(op, n, what, len(r), float(np.nansum(r.value)))
trn = make_tracks(160, [5000], seed=8)
gn = [api.Track(t["chrom"], t["start"], t["end"], t["value"] + 1e-9) for t in trn]
r = eng.reduce("median", gn, ["chr1"]); print("median 160 doubles (bitset sweep)", len(r), float(np.nansum(r.value)))
for n in (48, 300):
    trn = make_tracks(n, [40000], seed=9, vmax=4096)
    gn = [api.Track(t["chrom"], t["start"], t["end"], t["value"]) for t in trn]
    for op in ("var", "stddev", "cv"):
        r = eng.reduce(op, gn, ["chr1"]); print(op, n, len(r), float(np.nansum(r.value)), "sweep kernel" if eng.last_stats[-1]["sweep_kernel"] else "emit kernel")
