#!/usr/bin/env python
"""Host BigWig batch decoder: throughput of wb200_bw_decode_chrom on a
synthetic single-chromosome BigWig (bedGraph sections of 512 items, zlib-compressed, the layout
UCSC's tools write), on this machine's host cores.  No GPU involved.
usage: python profiles/bw_decode_bench.py [intervals=2e6] -> one JSON line"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bwwrite import write_bigwig  # noqa: E402
import ctypes as C  # noqa: E402
from wiggletools_b200 import _lib, bigwig  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(1)
lens = rng.integers(1, 400, n)
gaps = rng.integers(0, 100, n) * rng.integers(0, 2, n)
start = np.cumsum(gaps + np.r_[0, lens[:-1]])
vals = (rng.integers(0, 100000, n) / 64.0).astype(np.float32)
items = list(zip(start.tolist(), (start + lens).tolist(), vals.tolist()))
sections = [(0, 1, items[k:k + 512], 0, 0) for k in range(0, n, 512)]
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "bench.bw")
    write_bigwig(path, [("chr1", int(start[-1] + lens[-1] + 10))], sections, compress=True, rtree_block=256)
    size = os.path.getsize(path)
    for stretch, tag in ((10000, "reference stretch semantics (10000-bp boxing, quirk Q9)"), (0, "plain records")):
        best = None
        bw = bigwig.BigWig(path, stretch=stretch)
        for rep in range(5):                                                  # rep 0 also allocates and first-touches the output arrays
            t = _lib.CTrack()
            t0 = time.perf_counter()
            rc = bw.lib.wb200_bw_decode_chrom(bw.h, 0, C.addressof(t))        # the C call alone: arrays stay in the handle
            dt = time.perf_counter() - t0
            assert rc == 0
            s = range(t.n)
            if rep == 0:
                first = dt
            elif best is None or dt < best:
                best = dt
        bw.close()
        print(json.dumps({"decoder": tag, "file_bytes": size, "intervals_in_file": n, "records_out": int(len(s)),
                          "ms": round(1e3 * best, 2), "first_call_ms": round(1e3 * first, 2), "records_per_s": round(len(s) / best), "file_MBps": round(size / best / 1e6, 1),
                          "soa_out_MBps": round(16 * len(s) / best / 1e6, 1), "host_cores": os.cpu_count()}))
