#!/usr/bin/env python
"""Device text sink beside the reference's own sink.

Workload: `mean` over N synthetic tracks of one chromosome (same generator as bench.py), then
  * device: wb200_format_device on the HBM-resident segments (write_bg and write), CUDA-event
    time of the kernels, and the whole call including the D2H of the text;
  * CPU:    the UNMODIFIED reference sink (oracle/_ref: TeeWiggleIterator + printBlock) on a
    bounded sample of the same segments, written to /dev/null-like tmp.
usage: python profiles/sink_bench.py [tracks=64] [chrom_bp=60e6] -> one JSON line"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wiggletools_b200 import _lib  # noqa: E402
from wiggletools_b200.synth import make_tracks  # noqa: E402

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
chrom_bp = int(float(sys.argv[2])) if len(sys.argv) > 2 else 60_000_000
lib = _lib.load()
ctx = lib.wb200_create(0)
assert ctx, lib.wb200_last_error(None)
dev = torch.device("cuda:0")

tracks = make_tracks(n_tracks, [chrom_bp], seed=7)
off = np.zeros(n_tracks + 1, np.int64)
for i, t in enumerate(tracks):
    off[i + 1] = off[i] + len(t["start"])
cat = lambda k, dt: torch.from_numpy(np.concatenate([t[k] for t in tracks]).astype(dt)).to(dev)
d_s, d_e, d_v = cat("start", np.int32), cat("end", np.int32), cat("value", np.float64)
I = int(off[-1])
cap = min(2 * I + 1, chrom_bp + 2)
o_s = torch.empty(cap, dtype=torch.int32, device=dev); o_e = torch.empty(cap, dtype=torch.int32, device=dev)
o_v = torch.empty(cap, dtype=torch.float64, device=dev)
dflt = np.zeros(n_tracks)
n_out, dv, auc = C.c_int64(0), C.c_double(0), C.c_double(0)
rc = lib.wb200_reduce_device(ctx, 2, n_tracks, n_tracks, 0, d_s.data_ptr(), d_e.data_ptr(), d_v.data_ptr(), off.ctypes.data,
                             dflt.ctypes.data, 1, chrom_bp + 1, o_s.data_ptr(), o_e.data_ptr(), o_v.data_ptr(), cap,
                             C.addressof(n_out), C.addressof(dv), C.addressof(auc))
assert rc == 0, lib.wb200_last_error(ctx)
S = n_out.value
out = {"workload": "mean over %d tracks, one %.0f Mbp chromosome, then the default sink" % (n_tracks, chrom_bp / 1e6),
       "segments": S}

for mode, name in ((1, "write_bg"), (0, "write")):
    best = None
    for rep in range(4):
        state = _lib.CSinkState(0, 0, 0)
        text, nbytes = C.c_char_p(), C.c_int64(0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = lib.wb200_format_device(ctx, mode, b"chr1", o_s.data_ptr(), o_e.data_ptr(), o_v.data_ptr(), S,
                                     C.addressof(state), C.addressof(text), C.addressof(nbytes))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        assert rc == 0, lib.wb200_last_error(ctx)
        if rep > 0 and (best is None or dt < best):
            best = dt
    out[name] = {"records": int(state.records), "text_bytes": int(nbytes.value), "call_ms": round(1e3 * best, 2),
                 "records_per_s": round(state.records / best), "text_GBps_incl_d2h": round(nbytes.value / best / 1e9, 2),
                 "algorithmic_bytes": 16 * S + int(nbytes.value)}

# CPU: the reference's own sink on a sample of the same segments
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle  # noqa: E402
try:
    ref = Oracle("reference"); kind = "reference"
except FileNotFoundError:
    ref = Oracle("port"); kind = "port"
m = min(S, 4_000_000)
hs, he, hv = o_s[:m].cpu().numpy(), o_e[:m].cpu().numpy(), o_v[:m].cpu().numpy()
for bg, name in ((True, "write_bg"), (False, "write")):
    t0 = time.perf_counter()
    text = ref.write(np.zeros(m, np.int32), hs, he, hv, ("chr1",), bg)
    dt = time.perf_counter() - t0
    out[name]["cpu_" + kind + "_records_per_s"] = round(m / dt)
    out[name]["cpu_sample"] = "%d segments, %.2f s (includes reading the file back)" % (m, dt)
    out[name]["speedup_vs_cpu"] = round(out[name]["records_per_s"] / (m / dt), 1)
print(json.dumps(out))
lib.wb200_destroy(ctx)
