"""Several GPUs from one process through the C ABI (wb200_multi_*, csrc/multi.cpp): chromosome
shards packed onto the visible devices, one host thread per device, the AUC scalar all-reduced and
the Pearson moments all-gathered over NCCL.  Runs on however many GPUs the box has (one included:
the NCCL calls then have a single rank); results must equal the single-context ones -- segments bit
for bit (no data crosses GPUs), scalars within 1e-12 (the partials are folded in a different order)."""
import ctypes as C
import math

import numpy as np
import pytest

from wiggletools_b200 import _lib, api
from wiggletools_b200.synth import make_tracks

pytestmark = pytest.mark.gpu

NAMES = ["chr1", "chr10", "chr2", "chr21", "chrX"]


def test_multi_reduce_and_pearson(engine):
    lib = _lib.load()
    m = lib.wb200_multi_create(0, None)
    assert m, lib.wb200_multi_last_error(None).decode()
    try:
        nd = lib.wb200_multi_ndev(m)
        assert nd >= 1
        lengths = [40_000, 12_000, 25_000, 6_000, 18_000]
        n = 6
        tr = make_tracks(n, lengths, seed=51)
        shards, keep = [], []
        for c in range(len(lengths)):
            per = []
            for t in tr:
                sel = t["chrom"] == c
                per.append((np.ascontiguousarray(t["start"][sel]), np.ascontiguousarray(t["end"][sel]), np.ascontiguousarray(t["value"][sel]), 0.0))
            shards.append(per)
        want = engine.reduce_shards("mean", shards)
        arr = (_lib.CTrack * (len(shards) * n))()
        for c, per in enumerate(shards):
            for i, (s, e, v, d) in enumerate(per):
                keep.append((s, e, v))
                arr[c * n + i] = _lib.CTrack(s.ctypes.data, e.ctypes.data, v.ctypes.data, len(s), d)
        out = (_lib.CSegments * len(shards))()
        auc = C.c_double(0)
        rc = lib.wb200_multi_reduce_host(m, 2, len(shards), n, n, 0, C.addressof(arr), C.addressof(out), C.addressof(auc))
        assert rc == 0, lib.wb200_multi_last_error(m).decode()
        owners = {lib.wb200_multi_owner(m, c) for c in range(len(shards))}
        assert owners <= set(range(nd)) and (nd == 1 or len(owners) > 1)
        for c in range(len(shards)):
            k = out[c].n
            assert k == len(want[c][0])
            gs = np.ctypeslib.as_array(C.cast(out[c].start, C.POINTER(C.c_int32)), (k,))
            ge = np.ctypeslib.as_array(C.cast(out[c].end, C.POINTER(C.c_int32)), (k,))
            gv = np.ctypeslib.as_array(C.cast(out[c].value, C.POINTER(C.c_double)), (k,))
            assert np.array_equal(gs, want[c][0]) and np.array_equal(ge, want[c][1]) and np.array_equal(gv, want[c][2])
        assert math.isclose(auc.value, sum(w[4] for w in want), rel_tol=1e-12)
        # Pearson: moments of two tracks, shards in chromosome order
        xs = (_lib.CTrack * len(shards))(*[_lib.CTrack(p[0][0].ctypes.data, p[0][1].ctypes.data, p[0][2].ctypes.data, len(p[0][0]), 0.0) for p in shards])
        ys = (_lib.CTrack * len(shards))(*[_lib.CTrack(p[1][0].ctypes.data, p[1][1].ctypes.data, p[1][2].ctypes.data, len(p[1][0]), 0.0) for p in shards])
        mom = (C.c_double * 6)()
        rc = lib.wb200_multi_pearson_host(m, len(shards), C.addressof(xs), C.addressof(ys), C.addressof(mom))
        assert rc == 0, lib.wb200_multi_last_error(m).decode()
        g = [api.Track(t["chrom"], t["start"], t["end"], t["value"]) for t in tr[:2]]
        # chromosome indices are visited in list order here, so name them in an order strcmp keeps
        r_single = engine.pearson(g[0], g[1], ["c%d" % k for k in range(len(lengths))])
        r_multi = lib.wb200_pearson_result(C.addressof(mom))
        assert math.isclose(r_multi, r_single, rel_tol=1e-10), (r_multi, r_single)
    finally:
        lib.wb200_multi_destroy(m)
