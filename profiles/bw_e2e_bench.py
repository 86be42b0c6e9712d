#!/usr/bin/env python
"""End to end from BigWig files (SURVEY 8(d) baseline B beside the drop-in): the same command line,
`write_bg <out> mean t0.bw ... tN.bw`, through
  (ref)     oracle/_ref/wiggletools_bw: the reference's own pthread pipeline -- N reader threads
            (src/bigWiggleReader.c + src/bufferedReader.c, unmodified, over the libBigWig-interface shim),
            the main-thread heap sweep + MeanReduction, the writer thread;
  (dropin)  integration/_build/wiggletools_b200: the reference's front end linked against the library --
            batch BigWig decode, reduction and text formatting on the GPU, fwrite.
Both outputs must be byte-identical.  WB200_PROFILE=1 makes the drop-in report where its time goes.
usage: python profiles/bw_e2e_bench.py [tracks=8] [chrom_mbp=20] -> one JSON line"""
import hashlib
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

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
length = int(float(sys.argv[2]) * 1e6) if len(sys.argv) > 2 else 20_000_000
ref = os.path.join(ROOT, "oracle", "_ref", "wiggletools_bw")
drop = os.path.join(ROOT, "integration", "_build", "wiggletools_b200")
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    files, total = [], 0
    for t in range(n_tracks):
        rng = np.random.default_rng([7, t])
        s, e, v = make_chrom_track(rng, length)                       # same generator as bench.py (1-based half-open)
        items = np.empty(len(s), dtype=[("s", "<u4"), ("e", "<u4"), ("v", "<f4")])
        items["s"], items["e"], items["v"] = s - 1, e - 1, v
        sections = [(0, 1, items[k:k + 512], 0, 0) for k in range(0, len(items), 512)]
        path = os.path.join(d, "t%d.bw" % t)
        write_bigwig(path, [("chr1", length + 1)], sections, compress=True, rtree_block=256)
        files.append(path)
        total += len(items)
    out = {"workload": "write_bg <file> mean <%d BigWig tracks>, one %.0f Mbp chromosome, %d intervals" % (n_tracks, length / 1e6, total),
           "host_cores": os.cpu_count()}
    digests = {}
    for name, exe in (("reference", ref), ("dropin", drop)):
        if not os.path.exists(exe):
            out[name] = "not built"
            continue
        best, prof = None, ""
        target = os.path.join(d, name + ".bg")
        for rep in range(2):
            if os.path.exists(target):
                os.remove(target)
            t0 = time.perf_counter()
            r = subprocess.run([exe, "write_bg", target, "mean"] + files, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                               text=True, env=dict(os.environ, WB200_PROFILE="1"))
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, prof = dt, r.stderr.strip()
        h = hashlib.sha256()
        with open(target, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        digests[name] = h.hexdigest()
        out[name] = {"seconds": round(best, 3), "Mbp_per_s": round(length / 1e6 / best, 2), "output_bytes": os.path.getsize(target)}
        if prof:
            out[name]["profile"] = prof.split("\n")[-1]
    if len(digests) == 2:
        out["outputs_identical"] = digests["reference"] == digests["dropin"]
        out["speedup"] = round(out["reference"]["seconds"] / out["dropin"]["seconds"], 2)
    print(json.dumps(out))
