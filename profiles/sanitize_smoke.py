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
