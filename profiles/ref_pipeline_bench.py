#!/usr/bin/env python
"""SURVEY 8(d) CPU baseline (B): the reference's own pthread pipeline from BigWig files -- N reader
threads (src/bigWiggleReader.c + src/bufferedReader.c, unmodified, over the libBigWig-interface shim
of oracle/bigwig_shim.c), the main-thread heap sweep + MeanReduction, and the writer thread
(`write_bg`) -- on this machine's host cores.  No GPU involved.  Note that the shim decodes a
chromosome once and serves the 10000-bp stretch queries by binary search, which is cheaper than
libBigWig's per-query R-tree walk and inflate: the number flatters the reference's reader.
usage: python profiles/ref_pipeline_bench.py [tracks=16] [chrom_mbp=20] -> one JSON line"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bwwrite import write_bigwig  # noqa: E402
from wiggletools_b200.synth import make_chrom_track  # noqa: E402

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
length = int(float(sys.argv[2]) * 1e6) if len(sys.argv) > 2 else 20_000_000
exe = os.path.join(ROOT, "oracle", "_ref", "wiggletools_bw")
with tempfile.TemporaryDirectory() as d:
    files = []
    total = 0
    for t in range(n_tracks):
        rng = np.random.default_rng([7, t])
        s, e, v = make_chrom_track(rng, length)                       # same generator as bench.py (1-based half-open)
        items = list(zip((s - 1).tolist(), (e - 1).tolist(), v.astype(np.float32).tolist()))
        sections = [(0, 1, items[k:k + 512], 0, 0) for k in range(0, len(items), 512)]
        path = os.path.join(d, "t%d.bw" % t)
        write_bigwig(path, [("chr1", length + 1)], sections, compress=True, rtree_block=256)
        files.append(path)
        total += len(items)
    out = {"workload": "wiggletools write_bg /dev/null mean <%d BigWig tracks>, one %.0f Mbp chromosome, %d intervals" % (n_tracks, length / 1e6, total),
           "host_cores": os.cpu_count()}
    for name, prog in (("write_bg", ["write_bg", os.path.join(d, "out.bg"), "mean"] + files), ("AUC (no writer)", ["AUC", "mean"] + files)):
        best = None
        for rep in range(2):
            if os.path.exists(os.path.join(d, "out.bg")):
                os.remove(os.path.join(d, "out.bg"))
            t0 = time.perf_counter()
            subprocess.run([exe] + prog, check=True, stdout=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out[name] = {"seconds": round(best, 2), "Mbp_per_s": round(length / 1e6 / best, 2)}
    print(json.dumps(out))
