"""Multi-GPU host logic: chromosome sharding and the whole-genome scalars.

The sweep never crosses a chromosome (multiplexer.c:50-74), so chromosomes are
the shard unit: one process per GPU, each owning a set of chromosomes, no
data-path collective.  Only the scalars cross ranks: the AUC partial sum
(all_reduce) and the six Pearson moments (all_gather + pairwise merge)."""
import ctypes as C


def lpt_assign(weights, world):
    """Longest-processing-time bin packing: owner rank of each item."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    load = [0] * world
    owner = [0] * len(weights)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += weights[i]
    return owner


def allreduce_auc(partial, dist=None, device=None):
    """Sum of per-rank AUC partials (ncclSum / gloo sum on one fp64)."""
    import torch
    t = torch.tensor([partial], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return float(t.item())


def allgather_pearson(moments, dist=None, device=None):
    """Gather every rank's {n, sum_X, sum_Y, T_XX, T_XY, T_YY} and fold them in
    rank order with the engine's pairwise merge; returns (moments, r)."""
    import torch
    from . import _lib
    lib = _lib.load()
    t = torch.tensor(list(moments), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, t)
    else:
        parts = [t]
    total = (C.c_double * 6)(0, 0, 0, 0, 0, 0)
    for p in parts:
        m = (C.c_double * 6)(*[float(x) for x in p.cpu()])
        lib.wb200_pearson_merge(C.addressof(total), C.addressof(m))
    return list(total), lib.wb200_pearson_result(C.addressof(total))
