#!/usr/bin/env python
"""Benchmark of the merge+reduce hot path: Mbp/s for `mean` over N synthetic
genome tracks (BASELINE.json), on 1..8 B200s, with the HBM roofline of the
dominant kernel, an in-bench parity gate against the compiled reference, and the
reference's CPU path timed beside it.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference        # the reference's CPU path on host cores

A step = one pass of the hot path over the whole synthetic genome (24 hg38-like
chromosomes, 3.09 Gbp): ONE wb200_reduce_device_batch call (per chromosome a
cursor kernel and a sweep kernel, queued back to back) on inputs resident in HBM.
With N GPUs the chromosomes are sharded across ranks (no data-path collective);
the whole-genome AUC scalar is all-reduced over NCCL.

The headline line is BASELINE configs[1] (`mean`, 64 tracks).  The same run also
measures the other BASELINE configurations on track-prefix views of one resident
512-track data set and reports them under "extra": mean over 8 and 512 tracks,
var + stddev over 512, median over 256, and mean over 64 tracks whose values carry
full float32 mantissas (the two-limb accumulator).  Every workload is parity-gated:
after its timed region the segments it produced are compared with the unmodified
reference (oracle/_ref, else the oracle port) on whole chromosomes / regions --
segment count, bit-identical (start, end), values within 1e-10 relative -- and the
process exits non-zero on a mismatch.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from wiggletools_b200.synth import sorted_chroms  # noqa: E402
from wiggletools_b200.shard import lpt_assign  # noqa: E402

SCALE = float(os.environ.get("WB200_BENCH_SCALE", 1.0))   # <1 shrinks every chromosome (development only)
MAX_LEN, MAX_GAP, GAP_PROB, VMAX = 400, 200, 0.5, 100000
MEAN_STEP = (MAX_LEN + 1) / 2 + GAP_PROB * (MAX_GAP + 1) / 2
OP_CODE = {"sum": 0, "mult": 1, "mean": 2, "var": 3, "stddev": 4, "cv": 6, "min": 7, "max": 8, "median": 9}
PAD = 1024          # parity regions: intervals are at most MAX_LEN long, so every segment that starts in a region ends within PAD of it
REL_TOL = 1e-10     # north_star: fp64 reductions within 1e-10 relative of the reference


def workloads(world):
    """The headline (first) and the extra workloads of this run."""
    dev_n, dev_op = os.environ.get("WB200_BENCH_TRACKS"), os.environ.get("WB200_BENCH_OP")
    if dev_n or dev_op:     # development: one workload, no extras
        return [dict(key="dev", op=dev_op or "mean", N=int(dev_n or 64), f32=False, e2e=True)]
    wl = [dict(key="mean64", op="mean", N=64, f32=False, e2e=True, config="configs[1]")]
    if os.environ.get("WB200_BENCH_EXTRAS", "1") == "0":
        return wl
    wl += [dict(key="mean8", op="mean", N=8, f32=False, e2e=world == 1, config="north_star N=8"),
           dict(key="mean512", op="mean", N=512, f32=False, e2e=False, config="configs[4] / north_star N=512")]
    if world == 1:
        wl += [dict(key="var512", op="var", N=512, f32=False, e2e=False, config="configs[2]"),
               dict(key="stddev512", op="stddev", N=512, f32=False, e2e=False, config="configs[2]"),
               dict(key="median256", op="median", N=256, f32=False, e2e=False, config="configs[3]"),
               dict(key="mean64_f32", op="mean", N=64, f32=True, e2e=False, config="configs[1] with float32-mantissa values (two-limb accumulator)")]
    if os.environ.get("WB200_BENCH_MEDIAN_CHROMS"):     # development: the median line on a few chromosomes only
        for w in wl:
            if w["key"] == "median256":
                w["chroms"] = os.environ["WB200_BENCH_MEDIAN_CHROMS"].split(",")
    return wl


def genome():
    names, lengths = sorted_chroms()
    lengths = [max(10000, int(l * SCALE)) for l in lengths]
    return names, lengths


# ---------------------------------------------------------------------------
# synthetic data on the device (torch is plumbing: memory + RNG + NCCL)
# ---------------------------------------------------------------------------
def gen_chrom_device(torch, dev, n_tracks, length, seed, n_f32=0):
    """All tracks of one chromosome, concatenated track-major, on the GPU.  Values are
    multiples of 1/64 below 1563; the first n_f32 tracks also get a second
    value array with full float32 mantissas."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = int(length / MEAN_STEP * 1.06) + 64
    starts, ends, values, counts = [], [], [], []
    chunk = max(1, min(n_tracks, int(1.2e8 // n)))
    for t0 in range(0, n_tracks, chunk):
        k = min(chunk, n_tracks - t0)
        lens = torch.randint(1, MAX_LEN + 1, (k, n), device=dev, generator=g, dtype=torch.int64)
        gaps = torch.randint(1, MAX_GAP + 1, (k, n), device=dev, generator=g, dtype=torch.int64)
        gaps = gaps * (torch.rand((k, n), device=dev, generator=g) < GAP_PROB)
        step = gaps + torch.cat([torch.zeros((k, 1), device=dev, dtype=torch.int64), lens[:, :-1]], 1)
        s = 1 + torch.cumsum(step, 1)
        e = s + lens
        keep = e <= length + 1
        counts.append(keep.sum(1))
        starts.append(s[keep].to(torch.int32))
        ends.append(e[keep].to(torch.int32))
        v = torch.randint(0, VMAX, (k, n), device=dev, generator=g, dtype=torch.int64)
        values.append((v[keep].to(torch.float64)) / 64.0)
        del lens, gaps, step, s, e, keep, v
    counts = torch.cat(counts).cpu().numpy().astype(np.int64)
    off = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    v32 = None
    if n_f32:
        m = int(off[min(n_f32, n_tracks)])
        v32 = (torch.rand(m, device=dev, generator=g, dtype=torch.float32) * (VMAX / 64.0)).to(torch.float64)
    return torch.cat(starts), torch.cat(ends), torch.cat(values), off, v32


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML, every 20 ms;
    falls back to polling nvidia-smi)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is not None:
            nv = self.nv
            bits = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                    ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                    ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
            while not self.stop_flag:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append([str(sm), str(self.max_sm)] + ["Active" if r & b else "Not Active" for _, b in bits])
                except Exception:
                    pass
                time.sleep(0.02)     # NVML queries share driver locks with the launching thread: keep them sparse
            return
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for j, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[2 + j].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------
# the reference's CPU path: the parity checker of every workload, and the CPU baseline
# ---------------------------------------------------------------------------
def oracle_kind():
    return "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libwtref.so")) else "port"


def _region_worker(args):
    """One parity region [a, b) of an exported shard: the reference (oracle/_ref) reduces the
    intervals that overlap [a, b + PAD); the segments that start inside [a, b) come back."""
    kind, path, n_tracks, op, a, b = args
    # every breakpoint in [a, b + PAD) and every interval covering a position in it is kept, so the
    # segments that start in [a, b) come out exactly as in the whole-chromosome problem
    from oracle.pyoracle import Oracle, Track
    s = np.load(os.path.join(path, "s.npy"), mmap_mode="r")
    e = np.load(os.path.join(path, "e.npy"), mmap_mode="r")
    v = np.load(os.path.join(path, "v.npy"), mmap_mode="r")
    off = np.load(os.path.join(path, "off.npy"))
    tracks = []
    for i in range(n_tracks):
        si, ei = s[off[i]:off[i + 1]], e[off[i]:off[i + 1]]
        lo = int(np.searchsorted(ei, a, side="left"))            # first interval that ends at or after a (an end AT a is a breakpoint at a)
        hi = int(np.searchsorted(si, b + PAD, side="left"))      # first interval that starts at or after b + PAD
        hi = max(hi, lo)
        tracks.append(Track(np.zeros(hi - lo, np.int32), np.array(si[lo:hi]), np.array(ei[lo:hi]),
                            np.array(v[off[i] + lo:off[i] + hi])))
    orc = Oracle(kind)
    t0 = time.perf_counter()
    seg = orc.reduce(op, tracks, ["chr1"])
    dt = time.perf_counter() - t0
    m = (seg.start >= a) & (seg.start < b)
    return a, b, dt, seg.start[m], seg.end[m], seg.value[m]


def _cpu_worker(args):
    kind, n_tracks, length, seed = args
    from oracle.pyoracle import Oracle, Track
    from wiggletools_b200.synth import make_chrom_track
    orc = Oracle(kind)
    tracks = []
    for t in range(n_tracks):
        rng = np.random.default_rng([seed, t])
        s, e, v = make_chrom_track(rng, length, max_len=MAX_LEN, max_gap=MAX_GAP, gap_prob=GAP_PROB, vmax=VMAX)
        tracks.append(Track(np.zeros(len(s), np.int32), s, e, v))
    t0 = time.perf_counter()
    seg = orc.reduce("mean", tracks, ["chr1"])
    dt = time.perf_counter() - t0
    return length, dt, len(seg)


def cpu_reference_sample(n_tracks, mbp_per_core, cores):
    """Time the reference CPU path (`mean`) on `cores` independent regions of
    mbp_per_core Mbp each (the reference's own scale-out model: one process per
    genome region, python/wiggletools/parallelWiggleTools.py:66-68)."""
    import multiprocessing as mp
    kind = oracle_kind()
    length = int(mbp_per_core * 1e6)
    jobs = [(kind, n_tracks, length, 1000 + i) for i in range(cores)]
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = max(r[1] for r in res)     # slowest shard, generation excluded
    total_bp = sum(r[0] for r in res)
    return {"value": total_bp / 1e6 / wall, "unit": "Mbp/s", "cores": cores, "kind": kind,
            "sample": "mean over %d tracks, %d region(s) of %.1f Mbp (same generator as the GPU workload), %.1f s" %
                      (n_tracks, cores, mbp_per_core, wall)}


def run_reference_arm(args):
    """`--impl reference`: the UNMODIFIED reference (oracle/_ref) on this box's host
    cores, with the reference's own scale-out model (independent processes on
    independent genome regions).  The process count that gives the reference its
    best throughput is picked during the warm-up steps (memory bandwidth makes
    "all cores" slower than a subset on large hosts)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n_tracks = int(os.environ.get("WB200_BENCH_TRACKS", 64))
    ncpu = os.cpu_count() or 1
    mbp = 8.0 * (64.0 / n_tracks) ** 2          # ~1 s per step per core at N=64
    candidates = sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4)}, reverse=True)
    best, tried = None, []
    for i in range(max(1, args.warmup)):
        c = candidates[min(i, len(candidates) - 1)]
        r = cpu_reference_sample(n_tracks, mbp, c)
        tried.append((c, r["value"]))
        if best is None or r["value"] > best[1]:
            best = (c, r["value"])
    cores = best[0]
    times = [cpu_reference_sample(n_tracks, mbp, cores) for _ in range(args.steps)]
    value = float(np.mean([t["value"] for t in times]))
    cb = dict(times[-1]); cb["value"] = value
    cb["sample"] += "; process counts tried in warm-up: " + ", ".join("%d -> %.0f Mbp/s" % t for t in tried)
    names, lengths = genome()
    line = {"impl": "reference", "metric": "Mbp/s, mean over %d tracks" % n_tracks, "value": value, "unit": "Mbp/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (cores * mbp) / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config("mean", n_tracks, lengths, args.gpus, False), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(op, n_tracks, lengths, gpus, f32):
    return {"workload": "%s over %d synthetic tracks, %d hg38-like chromosomes (%.2f Gbp), ~%.1e intervals/track%s"
                        % (op, n_tracks, len(lengths), sum(lengths) / 1e9, sum(lengths) / MEAN_STEP,
                           ", float32-mantissa values" if f32 else ""),
            "tracks": n_tracks, "genome_bp": int(sum(lengths)), "sharding": "chromosomes over %d GPU(s), LPT" % gpus,
            "l2": "inputs (%.1f GB/step) larger than L2; no flush needed" % (n_tracks * sum(lengths) / MEAN_STEP * 16 / 1e9)}


# ---------------------------------------------------------------------------
# the parity gate
# ---------------------------------------------------------------------------
class ParityChecker:
    """Compares segments left in HBM by a timed workload with the reference CPU path on the
    same inputs, region by region, on a pool of host processes (which also yields the CPU
    baseline of the workload: the reference timed on this box's cores)."""

    def __init__(self, torch, procs):
        import multiprocessing as mp
        self.torch = torch
        self.procs = max(1, procs)
        self.pool = mp.get_context("spawn").Pool(self.procs)
        self.dir = tempfile.mkdtemp(prefix="wb200_parity_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        self.exports = {}
        self.kind = oracle_kind()

    def close(self):
        self.pool.close()
        self.pool.join()
        shutil.rmtree(self.dir, ignore_errors=True)

    def export(self, sh, a, b, f32):
        """Intervals of every track of shard `sh` overlapping [a, b + PAD), as .npy files the workers map."""
        key = (sh["ci"], a, b, bool(f32))
        if key in self.exports:
            return self.exports[key]
        torch = self.torch
        path = os.path.join(self.dir, "x%d" % len(self.exports))
        os.makedirs(path)
        off = sh["off"]
        ntr = len(off) - 1 if not f32 else sh["n_f32"]
        vals = sh["v32"] if f32 else sh["v"]
        if a <= 1 and b >= sh["L"] + 1:
            m = int(off[ntr])
            s_h, e_h, v_h, off_h = sh["s"][:m].cpu().numpy(), sh["e"][:m].cpu().numpy(), vals[:m].cpu().numpy(), off[:ntr + 1].copy()
        else:
            ps, pe, pv, cnt = [], [], [], []
            for i in range(ntr):
                o0, o1 = int(off[i]), int(off[i + 1])
                lo = int(torch.searchsorted(sh["e"][o0:o1], a))
                hi = max(lo, int(torch.searchsorted(sh["s"][o0:o1], b + PAD)))
                ps.append(sh["s"][o0 + lo:o0 + hi]); pe.append(sh["e"][o0 + lo:o0 + hi]); pv.append(vals[o0 + lo:o0 + hi])
                cnt.append(hi - lo)
            s_h, e_h, v_h = torch.cat(ps).cpu().numpy(), torch.cat(pe).cpu().numpy(), torch.cat(pv).cpu().numpy()
            off_h = np.concatenate(([0], np.cumsum(cnt))).astype(np.int64)
        for name, arr in (("s", s_h), ("e", e_h), ("v", v_h), ("off", off_h)):
            np.save(os.path.join(path, name + ".npy"), arr)
        self.exports[key] = path
        return path

    def check(self, wl, regions):
        """regions: [(shard, a, b, out_start, out_end, out_value (device, S valid entries), S)].
        Every region is cut into chunks for the process pool."""
        torch = self.torch
        jobs, meta = [], []
        per_mbp = {"mean": 0.0016, "sum": 0.0016, "var": 0.005, "stddev": 0.005, "median": 0.025}.get(wl["op"], 0.01) * wl["N"]  # ~core-seconds per Mbp
        for ri, (sh, a, b, os_, oe_, ov_, S) in enumerate(regions):
            path = self.export(sh, a, b, wl["f32"])
            # chunks of ~0.5 core-seconds, at least one per process
            nchunk = int(max(self.procs if (b - a) > 2e5 else 1, min(4096, per_mbp * (b - a) / 1e6 / 0.5)))
            edges = np.linspace(a, b, nchunk + 1).astype(np.int64)
            for k in range(nchunk):
                if edges[k + 1] > edges[k]:
                    jobs.append((self.kind, path, wl["N"], wl["op"], int(edges[k]), int(edges[k + 1])))
                    meta.append(ri)
        t0 = time.perf_counter()
        res = self.pool.map(_region_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        ok, nseg, bad_coord, max_rel, n_bits, cpu_s, bp = True, 0, 0, 0.0, 0, 0.0, 0
        detail = []
        host = {}
        for ri, (a, b, dt, rs, re_, rv) in zip(meta, res):
            sh, ra, rb, os_, oe_, ov_, S = regions[ri]
            if ri not in host:
                host[ri] = (os_[:S].cpu().numpy(), oe_[:S].cpu().numpy(), ov_[:S].cpu().numpy())
            hs, he, hv = host[ri]
            i0, i1 = int(np.searchsorted(hs, a, side="left")), int(np.searchsorted(hs, b, side="left"))
            gs, ge, gv = hs[i0:i1], he[i0:i1], hv[i0:i1]
            cpu_s += dt; bp += b - a
            if len(gs) != len(rs):
                ok = False
                detail.append("region [%d,%d) of shard %d: %d segments, reference has %d" % (a, b, sh["ci"], len(gs), len(rs)))
                continue
            nseg += len(gs)
            if not (np.array_equal(gs, rs) and np.array_equal(ge, re_)):
                ok = False
                bad_coord += int(np.count_nonzero((gs != rs) | (ge != re_)))
                detail.append("region [%d,%d) of shard %d: coordinates differ" % (a, b, sh["ci"]))
                continue
            both_nan = np.isnan(gv) & np.isnan(rv)
            with np.errstate(invalid="ignore", divide="ignore"):
                rel = np.abs(gv - rv) / np.maximum(np.abs(gv), np.abs(rv))
            rel[(gv == rv) | both_nan] = 0.0
            rel[np.isnan(rel)] = np.inf                     # NaN on one side only
            if len(rel):
                max_rel = max(max_rel, float(rel.max()))
            n_bits += int(np.count_nonzero((gv == rv) | both_nan))
            if len(rel) and rel.max() > REL_TOL:
                ok = False
                detail.append("region [%d,%d) of shard %d: value error %.3g" % (a, b, sh["ci"], float(rel.max())))
        out = {"ok": bool(ok), "oracle": "%s (%s)" % (self.kind, "oracle/_ref/libwtref.so: the unmodified reference" if self.kind == "reference" else "oracle/liboracle.so"),
               "regions": ["shard %d [%d,%d)" % (r[0]["ci"], r[1], r[2]) for r in regions], "bp_checked": int(bp),
               "segments_checked": int(nseg), "coords": "bit-identical" if bad_coord == 0 and ok or bad_coord == 0 and not detail else "%d differ" % bad_coord,
               "values_max_rel_err": max_rel, "values_bit_identical": int(n_bits), "tolerance": REL_TOL}
        if detail:
            out["mismatch"] = detail[:8]
        cpu = {"value": bp / 1e6 / wall if wall > 0 else 0.0, "unit": "Mbp/s", "cores": self.procs, "kind": self.kind,
               "sample": "%s over %d tracks: the parity regions (%.1f Mbp in %d chunks on %d processes, %.1f core-seconds of reference time, %.1f s wall)"
                         % (wl["op"], wl["N"], bp / 1e6, len(jobs), self.procs, cpu_s, wall)}
        return out, cpu


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the parity gate / CPU baseline (development A/B runs only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from wiggletools_b200 import _lib
    from wiggletools_b200._lib import CDeviceShard

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lib = _lib.load()
    ctx = lib.wb200_create(local)
    if not ctx:
        raise SystemExit(lib.wb200_last_error(None).decode())
    stream = torch.cuda.current_stream()
    lib.wb200_set_stream(ctx, C.c_void_p(stream.cuda_stream))
    # one process per GPU: keep this rank's host thread (and the pinned buffers it allocates from
    # here on) on the NUMA node of its GPU
    local_cpus = 0 if os.environ.get("WB200_NO_BIND") else lib.wb200_bind_host_thread(ctx)

    names, lengths = genome()
    owner = lpt_assign(lengths, world)
    mine = [i for i in range(len(lengths)) if owner[i] == rank]
    wls = workloads(world)
    n_max = max(w["N"] for w in wls)
    n_f32 = max([w["N"] for w in wls if w["f32"]] + [0])

    # resident inputs for this rank's chromosomes: one data set of n_max tracks; every workload
    # reads a track-prefix view of it (tracks are concatenated track-major per chromosome)
    shards = []
    for ci in mine:
        s, e, v, off, v32 = gen_chrom_device(torch, dev, n_max, lengths[ci], seed=7919 * (ci + 1), n_f32=n_f32)
        shards.append(dict(ci=ci, s=s, e=e, v=v, v32=v32, n_f32=n_f32, off=off, L=lengths[ci]))
    torch.cuda.synchronize()

    checker = None
    if not args.no_cpu:
        ncpu = os.cpu_count() or 1
        checker = ParityChecker(torch, max(1, min(32, ncpu // (2 * world))))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic_ratio = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes / algorithmic bytes from the committed ncu --set full capture
    if os.path.exists(tpath):
        traffic_ratio = json.load(open(tpath)).get("emit_kernel_dram_bytes_per_algorithmic_byte")

    def run_workload(wl, headline):
        N, op = wl["N"], wl["op"]
        use = [sh for sh in shards if "chroms" not in wl or names[sh["ci"]] in wl["chroms"]]
        if SCALE < 1.0 and "chroms" in wl:
            use = shards
        wl_lengths = [lengths[i] for i in range(len(lengths)) if "chroms" not in wl or SCALE < 1.0 or names[i] in wl["chroms"]]
        genome_bp = float(sum(wl_lengths))
        # outputs: one arena per workload
        caps = [int(min(2 * int(sh["off"][N]) + 1, sh["L"] + 1)) for sh in use]
        tot = sum(caps) + 16
        arena_s = torch.empty(tot, dtype=torch.int32, device=dev)
        arena_e = torch.empty(tot, dtype=torch.int32, device=dev)
        arena_v = torch.empty(tot, dtype=torch.float64, device=dev)
        arr = (CDeviceShard * max(1, len(use)))()
        offs, outs, o = [], [], 0
        for k, sh in enumerate(use):
            off = np.ascontiguousarray(sh["off"][:N + 1])
            offs.append(off)
            o = (o + 3) & ~3                       # 16-byte aligned output slices
            outs.append((arena_s[o:o + caps[k]], arena_e[o:o + caps[k]], arena_v[o:o + caps[k]]))
            vals = sh["v32"] if wl["f32"] else sh["v"]
            arr[k] = CDeviceShard(sh["s"].data_ptr(), sh["e"].data_ptr(), vals.data_ptr(), off.ctypes.data, 1, sh["L"] + 1,
                                  outs[k][0].data_ptr(), outs[k][1].data_ptr(), outs[k][2].data_ptr(), caps[k])
            o += caps[k]
        dflt = np.zeros(N, np.float64)
        dv = C.c_double(0)
        per_launch = []
        kernels = 0

        def step(collect):
            nonlocal kernels
            total_auc = 0.0
            if use:
                rc = lib.wb200_reduce_device_batch(ctx, OP_CODE[op], N, N, 0, dflt.ctypes.data, len(use), C.addressof(arr), C.addressof(dv))
                if rc != 0:
                    raise SystemExit("wb200_reduce_device_batch: %s" % lib.wb200_last_error(ctx).decode())
                for k in range(len(use)):
                    total_auc += arr[k].auc
                    if collect:
                        per_launch.append((int(offs[k][N]), int(arr[k].n_out), arr[k].ms_cursor, arr[k].ms_sweep, int(arr[k].slow_segments), int(arr[k].limbs), int(arr[k].sweep_kernel)))
                if collect:
                    kernels += lib.wb200_last_stats(ctx).contents.kernels
            t = torch.tensor([total_auc], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t)   # whole-genome AUC scalar: the path's only exchange step
            return t

        K = args.steps if headline else max(3, min(args.steps, 5))
        Wm = max(3, args.warmup) if not headline else args.warmup
        for _ in range(Wm):
            step(False)
        barrier()
        sampler = ClockSampler(local) if headline else None
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(K):
            auc_t = step(True)
        ev1.record()
        barrier()
        if sampler:
            sampler.stop_flag = True
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item()) / K
        value = genome_bp / 1e6 / (ms_per_step / 1e3)

        # roofline of the dominant kernel: algorithmic bytes 16 B per input interval read + 16 B per
        # segment written, per launch; duration from CUDA events the library records around each
        # launch on the launching stream.
        by = sum(16.0 * i + 16.0 * s for i, s, *_ in per_launch)
        t_emit = sum(r[3] for r in per_launch) / 1e3
        t_count = sum(r[2] for r in per_launch) / 1e3
        achieved = by / t_emit / 1e9 if t_emit > 0 else 0.0
        # which kernel ran is the library's own report per launch (sum / mean / var / stddev / CV run on
        # fastsum_kernel when their fixed-point format allows it, everything else on emit_kernel)
        kernel = "fastsum_kernel" if per_launch and all(r[6] for r in per_launch) else "emit_kernel"
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak,
                    "traffic": (traffic_ratio * by / len(per_launch)) if traffic_ratio and per_launch and kernel == "fastsum_kernel" and op in ("sum", "mean") else None,
                    "traffic_source": "ncu --set full capture of one launch (profiles/traffic.json): measured dram bytes / algorithmic bytes, times this run's algorithmic bytes per launch",
                    "peak_source": peak_src, "algorithmic_bytes_per_step": by / K,
                    "algorithmic_bytes_per_launch": by / max(1, len(per_launch)),
                    "emit_ms_per_step": 1e3 * t_emit / K, "cursor_ms_per_step": 1e3 * t_count / K,
                    "path_frac": (by / (t_emit + t_count) / 1e9) / peak if t_emit + t_count > 0 else 0.0,
                    # exact-sum format the library chose from the value range: 32-bit limbs per position (1: values like
                    # this generator's; 2: full float32 mantissas; 3: wider), 0 for the statistics without it
                    "accumulator_limbs": max((r[5] for r in per_launch), default=0)}
        seg_local = sum(int(arr[k].n_out) for k in range(len(use)))
        int_local = sum(int(offs[k][N]) for k in range(len(use)))
        slow = sum(r[4] for r in per_launch) // max(1, K)
        tot_t = torch.tensor([seg_local, int_local, slow, kernels], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot_t)

        # parity gate on what the timed region left in HBM: this rank's smallest shard (whole for
        # N <= 64, its first 8 Mbp otherwise) and, for N <= 64, the last 2 Mbp of its largest one
        parity, cpu = None, None
        if checker is not None and use:
            order = sorted(range(len(use)), key=lambda k: use[k]["L"])
            regions = []
            k0 = order[0]
            L0 = use[k0]["L"]
            b0 = L0 + 1 if N <= 64 else min(L0 + 1, 1 + int(8e6))
            regions.append((use[k0], 1, b0, outs[k0][0], outs[k0][1], outs[k0][2], int(arr[k0].n_out)))
            if N <= 64 and len(use) > 1:
                k1 = order[-1]
                L1 = use[k1]["L"]
                regions.append((use[k1], max(1, L1 + 1 - int(2e6)), L1 + 1, outs[k1][0], outs[k1][1], outs[k1][2], int(arr[k1].n_out)))
            parity, cpu = checker.check(wl, regions)
        pk = torch.tensor([1.0 if (parity is None or parity["ok"]) else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(pk, op=dist.ReduceOp.MIN)
        if parity is not None:
            parity["all_ranks_ok"] = bool(pk.item() > 0.5)

        line = {"metric": "Mbp/s, %s over %d tracks" % (op, N), "value": value, "unit": "Mbp/s", "n_gpus": world,
                "steps": K, "warmup": Wm, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(op, N, wl_lengths, world, wl["f32"]), "roofline": roofline, "parity": parity,
                "cpu_baseline": cpu, "gpu_launches": int(tot_t[3].item()),
                "segments": int(tot_t[0].item()), "intervals": int(tot_t[1].item()), "slow_segments": int(tot_t[2].item()),
                "auc": float(auc_t.item())}
        if "config" in wl:
            line["config"]["baseline_config"] = wl["config"]
        if sampler:
            line["clocks"] = sampler.summary()
        state = dict(use=use, outs=outs, arr=arr, offs=offs, genome_bp=genome_bp, ok=bool(pk.item() > 0.5))
        return line, state

    lines, all_ok = [], True
    for wi, wl in enumerate(wls):
        headline = wi == 0
        line, st = run_workload(wl, headline)
        all_ok = all_ok and st["ok"]
        # end to end through the host-buffer entry point: pinned host tracks -> H2D -> kernels -> D2H of
        # the segments, every step, timed with the wall clock around synchronous calls (max over ranks)
        if wl["e2e"] and not args.no_e2e:
            # `e2e`: the library's narrow PCIe transport (float32 values in, breakpoint-form segments out); `e2e_wide`: fp64 values
            # in, all three arrays out -- same call, same results (checked against the resident run), 25 % fewer bytes
            ksteps = max(1, args.steps if headline else min(args.steps, 3))
            e2e = run_e2e(torch, dist, lib, ctx, wl, st, dev, world, ksteps, narrow=True)
            e2e["host_thread_cpus"] = int(local_cpus)     # CPUs local to this rank's GPU the host thread was pinned to (0: not pinned)
            line["e2e"] = e2e
            if headline:
                line["e2e_wide"] = run_e2e(torch, dist, lib, ctx, wl, st, dev, world, ksteps, narrow=False)
            all_ok = all_ok and e2e["segments_match_resident_run"]
        elif not headline:
            line["e2e"] = None
        # the device text sink on this rank's smallest shard, outside the timed region:
        # `write_bg` of the HBM-resident segments, D2H of the text included
        if headline and rank == 0 and st["use"] and not args.no_e2e and wl["op"] in ("sum", "mean"):
            line["sink"] = run_sink(torch, lib, ctx, st, names)
        if headline:
            ph = (C.c_uint64 * 16)()
            lib.wb200_debug_phases(ctx, ph)
            if any(ph):                                   # diagnostic builds only
                names_ = ["clear", "events", "scan", "count", "compact", "lookback", "emit", "window_total", "windows", "lookback_polls|order_ranks", "lookback_steps|order_counters",
                          "order_offsets", "order_sort", "order_select", "order_stage", "order_keys"]
                line["phase_cycles_per_window"] = {n: ph[k] / max(1, ph[8]) for k, n in enumerate(names_)}
            if os.environ.get("WB200_BENCH_CHECKSUM"):      # development: order-sensitive digest of every output array
                acc = torch.zeros((), dtype=torch.int64, device=dev)
                for k in range(len(st["use"])):
                    S = int(st["arr"][k].n_out)
                    idx = torch.arange(1, S + 1, device=dev, dtype=torch.int64)
                    acc += (st["outs"][k][0][:S].long() * idx).sum() * 3 + (st["outs"][k][1][:S].long() * idx).sum() * 5
                    acc += (st["outs"][k][2][:S].view(torch.int64) * idx).sum()
                    del idx
                line["checksum"] = int(acc.item())
        lines.append(line)
        del st
        torch.cuda.empty_cache()

    if checker is not None:
        checker.close()
    if rank == 0:
        line = lines[0]
        if "e2e" not in line:
            line["e2e"] = None
        if len(lines) > 1:
            line["extra"] = lines[1:]
        line["parity_gate"] = "passed" if all_ok else "FAILED"
        print(json.dumps(line))
    lib.wb200_destroy(ctx)
    if world > 1:
        dist.destroy_process_group()
    if not all_ok:
        sys.exit(1)


def run_sink(torch, lib, ctx, st, names):
    from wiggletools_b200._lib import CSinkState
    k = min(range(len(st["use"])), key=lambda j: int(st["arr"][j].n_out))
    sh, S = st["use"][k], int(st["arr"][k].n_out)
    best = None
    for rep in range(3):
        state = CSinkState(0, 0, 0)
        text, nbytes = C.c_char_p(), C.c_int64(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.wb200_format_device(ctx, 1, names[sh["ci"]].encode(), st["outs"][k][0].data_ptr(), st["outs"][k][1].data_ptr(),
                                     st["outs"][k][2].data_ptr(), S, C.addressof(state), C.addressof(text), C.addressof(nbytes))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rc != 0:
            raise SystemExit("wb200_format_device: %s" % lib.wb200_last_error(ctx).decode())
        if rep and (best is None or dt < best):
            best = dt
    return {"mode": "write_bg", "shard": names[sh["ci"]], "records": int(state.records), "text_bytes": int(nbytes.value),
            "ms": 1e3 * best, "records_per_s": state.records / best, "d2h_GBps": nbytes.value / best / 1e9,
            "note": "wb200_format_device on HBM-resident segments, text copied to pinned host memory inside the call"}


def run_e2e(torch, dist, lib, ctx, wl, st, dev, world, steps, narrow):
    """The call a user of the library makes: wb200_reduce_host_batch on pinned host SoA tracks ->
    pinned host segments (H2D, kernels, D2H all inside the timed region).
    narrow=True is the library's narrow PCIe transport: float32 values in (what BigWig files hold; widened to
    fp64 on the device) and segments back in breakpoint form (start + value in full, end only where it is
    not the next start) -- 12 + 12 instead of 16 + 16 bytes per interval / segment; narrow=False moves fp64
    values in and all three arrays out."""
    from wiggletools_b200._lib import CTrack, CSegments
    use, n = st["use"], wl["N"]
    nsh = len(use)
    arr = (CTrack * max(1, nsh * n))()
    host = []
    for c, sh in enumerate(use):
        off = st["offs"][c]
        m = int(off[n])
        vals = sh["v32"] if wl["f32"] else sh["v"]
        hs = sh["s"][:m].cpu().pin_memory(); he = sh["e"][:m].cpu().pin_memory()
        if narrow:
            v32 = vals[:m].to(torch.float32)
            assert bool((v32.to(torch.float64) == vals[:m]).all()), "values are not float32-representable"
            hv = v32.cpu().pin_memory()
        else:
            hv = vals[:m].cpu().pin_memory()
        for i in range(n):
            a, b = int(off[i]), int(off[i + 1])
            if narrow:
                arr[c * n + i] = CTrack(hs.data_ptr() + 4 * a, he.data_ptr() + 4 * a, None, b - a, 0.0, hv.data_ptr() + 4 * a)
            else:
                arr[c * n + i] = CTrack(hs.data_ptr() + 4 * a, he.data_ptr() + 4 * a, hv.data_ptr() + 8 * a, b - a, 0.0)
        host.append((hs, he, hv))
    out = (CSegments * max(1, nsh))()
    lib.wb200_set_breakpoint_output(ctx, 1 if narrow else 0)

    def one():
        if not nsh:
            return
        rc = lib.wb200_reduce_host_batch(ctx, OP_CODE[wl["op"]], nsh, n, n, 0, C.addressof(arr), C.addressof(out))
        if rc != 0:
            raise SystemExit("wb200_reduce_host_batch: %s" % lib.wb200_last_error(ctx).decode())

    one()   # first call sizes the pinned result arena
    one()
    # what came back over PCIe is what the resident run left in HBM: counts for every shard, and all three
    # arrays (end rebuilt from the breakpoint form) bit for bit on this rank's smallest shard
    same = all(int(out[c].n) == int(st["arr"][c].n_out) for c in range(nsh))
    if nsh:
        c = min(range(nsh), key=lambda j: int(out[j].n))
        S = int(out[c].n)
        hs_ = np.ctypeslib.as_array(C.cast(out[c].start, C.POINTER(C.c_int32)), (S,))
        hv_ = np.ctypeslib.as_array(C.cast(out[c].value, C.POINTER(C.c_double)), (S,))
        if out[c].end:
            he_ = np.ctypeslib.as_array(C.cast(out[c].end, C.POINTER(C.c_int32)), (S,))
        else:
            k = int(out[c].n_ends)
            he_ = np.empty(S, np.int32)
            he_[:-1] = hs_[1:]
            idx = np.ctypeslib.as_array(C.cast(out[c].end_index, C.POINTER(C.c_int32)), (k,))
            he_[idx] = np.ctypeslib.as_array(C.cast(out[c].end_value, C.POINTER(C.c_int32)), (k,))
        os_, oe_, ov_ = (t[:S].cpu().numpy() for t in st["outs"][c])
        same = same and np.array_equal(hs_, os_) and np.array_equal(he_, oe_) and np.array_equal(hv_, ov_, equal_nan=True)
    stats = lib.wb200_last_stats(ctx).contents
    h2d, d2h = int(stats.h2d_bytes), int(stats.d2h_bytes)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    lib.wb200_set_breakpoint_output(ctx, 0)
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    tb = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tb)
    dtm = float(t.item())
    res = {"value": st["genome_bp"] / 1e6 / (dtm / steps), "unit": "Mbp/s", "h2d_bytes_per_step": int(tb[0].item()),
           "d2h_bytes_per_step": int(tb[1].item()), "steps": steps, "ms_per_step": 1e3 * dtm / steps,
           "segments_match_resident_run": bool(same),
           "this_rank_copy_GBps": (h2d + d2h) / (dt / steps) / 1e9 if dt > 0 else 0.0,
           "transport": ("narrow: int32 start/end + float32 value in (widened on the device), start + value + explicit ends only where a gap follows out"
                         if narrow else "wide: int32 start/end + fp64 value in, int32 start/end + fp64 value out"),
           "api": "wb200_reduce_host_batch (pinned host SoA in, pinned host segments out; staging, kernels and read-back pipelined per chromosome)"}
    del host
    return res


if __name__ == "__main__":
    main()
