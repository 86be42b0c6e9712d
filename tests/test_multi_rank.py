"""CPU tests of the N>1 host logic with world_size=2 over gloo: chromosome
sharding is a partition every rank agrees on, and the whole-genome scalars
(AUC all-reduce, Pearson moment all-gather + merge) equal the single-process
answers of the oracle."""
import math
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["chr1", "chr10", "chr2", "chrX"]
LENGTHS = [6000, 2500, 4000, 1500]


def _moments(x, y):
    """Per-shard Pearson moments of two step functions, by the defining sums."""
    n = float(len(x))
    sx, sy = x.sum(), y.sum()
    return [n, sx, sy, ((x - sx / n) ** 2).sum(), ((x - sx / n) * (y - sy / n)).sum(), ((y - sy / n) ** 2).sum()]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.pyoracle import Oracle, Track
    from wiggletools_b200 import shard
    from wiggletools_b200.synth import make_tracks
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    orc = Oracle("port")
    tr = make_tracks(4, LENGTHS, seed=11, max_len=40, max_gap=30)
    owner = shard.lpt_assign(LENGTHS, world)
    mine = [c for c in range(len(LENGTHS)) if owner[c] == rank]
    # this rank's partial AUC of `mean` over its chromosomes
    part = 0.0
    mom = [0.0] * 6
    for c in mine:
        sub = []
        for t in tr:
            m = t["chrom"] == c
            sub.append(Track(t["chrom"][m], t["start"][m], t["end"][m], t["value"][m]))
        part += orc.auc("mean", sub, NAMES)
        # dense per-base expansion of tracks 0 and 1 on the union support -> moments
        seg, vals, inplay = orc.rows(sub[:2], NAMES)
        L = (seg.end - seg.start)
        x = np.repeat(vals[:, 0], L); y = np.repeat(vals[:, 1], L)
        if len(x):
            import ctypes as C
            from wiggletools_b200 import _lib
            lib = _lib.load()
            a = (C.c_double * 6)(*mom); b = (C.c_double * 6)(*_moments(x, y))
            lib.wb200_pearson_merge(C.addressof(a), C.addressof(b))
            mom = list(a)
    total = shard.allreduce_auc(part, dist)
    _, r = shard.allgather_pearson(mom, dist)
    q.put((rank, owner, total, r))
    dist.destroy_process_group()


def test_world2_sharding_and_scalars():
    sys.path.insert(0, ROOT)
    from oracle.pyoracle import Oracle, Track
    from wiggletools_b200.synth import make_tracks
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, owner0, auc0, r0), (_, owner1, auc1, r1) = res
    assert owner0 == owner1 and sorted(set(owner0)) == [0, 1]          # same partition on every rank, both used
    tr = make_tracks(4, LENGTHS, seed=11, max_len=40, max_gap=30)
    tracks = [Track(t["chrom"], t["start"], t["end"], t["value"]) for t in tr]
    orc = Oracle("port")
    want_auc = orc.auc("mean", tracks, NAMES)
    assert auc0 == auc1 and math.isclose(auc0, want_auc, rel_tol=1e-12)
    want_r = orc.pearson(tracks[0], tracks[1], NAMES)
    assert r0 == r1 and math.isclose(r0, want_r, rel_tol=1e-9)


def test_lpt_balance():
    from wiggletools_b200.shard import lpt_assign
    from wiggletools_b200.synth import sorted_chroms
    _, lengths = sorted_chroms()
    for world in (1, 2, 4, 8):
        owner = lpt_assign(lengths, world)
        load = [sum(l for l, o in zip(lengths, owner) if o == r) for r in range(world)]
        assert sum(load) == sum(lengths)
        assert max(load) <= 1.08 * sum(lengths) / world
