#!/usr/bin/env python
"""PCIe ceiling for the staged (end-to-end) number: pinned host <-> HBM copy bandwidth, one
direction at a time and both at once (two streams), large and per-track-sized transfers.
usage: python profiles/pcie_probe.py  -> one JSON line"""
import json, time, torch
dev = torch.device("cuda:0")
GB = 1 << 30
def bw(nbytes, chunk, h2d, d2h, reps=3):
    n = nbytes // chunk
    hs = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); ds = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    hd = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); dd = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 0.0
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(n):
            a, b = k * chunk, (k + 1) * chunk
            if h2d:
                with torch.cuda.stream(s1): ds[a:b].copy_(hs[a:b], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): hd[a:b].copy_(dd[a:b], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = max(best, nbytes * (int(h2d) + int(d2h)) / dt / 1e9)
    return round(best, 1)
out = {}
for chunk in (64 << 20, 1 << 20):
    tag = "%dMiB" % (chunk >> 20)
    out["h2d_" + tag] = bw(2 * GB, chunk, True, False)
    out["d2h_" + tag] = bw(2 * GB, chunk, False, True)
    out["both_" + tag] = bw(2 * GB, chunk, True, True)
print(json.dumps({"pcie_GBps": out}))
