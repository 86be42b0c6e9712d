#!/usr/bin/env python
"""Benchmark of the merge+reduce hot path: Mbp/s for `mean` over N synthetic
genome tracks (BASELINE.json), on 1..8 B200s, with the HBM roofline of the
dominant kernel and the reference's CPU path timed beside it.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference        # the reference's CPU path on host cores

A step = one pass of the hot path over the whole synthetic genome (24 hg38-like
chromosomes, 3.1 Gbp, 64 tracks of ~1.2e7 intervals): per chromosome one
wb200_reduce_device call (cursor kernel, single-pass emit kernel) on inputs resident in
HBM.  With N GPUs the chromosomes are sharded across ranks (no data-path
collective); the whole-genome AUC scalar is all-reduced over NCCL.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from wiggletools_b200.synth import sorted_chroms  # noqa: E402
from wiggletools_b200.shard import lpt_assign  # noqa: E402

N_TRACKS = int(os.environ.get("WB200_BENCH_TRACKS", 64))
SCALE = float(os.environ.get("WB200_BENCH_SCALE", 1.0))   # <1 shrinks every chromosome (development only)
MAX_LEN, MAX_GAP, GAP_PROB, VMAX = 400, 200, 0.5, 100000
BENCH_OP = os.environ.get("WB200_BENCH_OP", "mean")       # development: time another statistic of the path
# the kernel the roofline line is about: sum / mean run on fastsum_kernel (csrc/fastsum.cu), the rest on emit_kernel
DOMINANT_KERNEL = "fastsum_kernel" if BENCH_OP in ("sum", "mean") and os.environ.get("WB200_FASTSUM") != "0" else "emit_kernel"
OP_CODE = {"sum": 0, "mult": 1, "mean": 2, "var": 3, "stddev": 4, "cv": 6, "min": 7, "max": 8, "median": 9}[BENCH_OP]


def genome():
    names, lengths = sorted_chroms()
    lengths = [max(10000, int(l * SCALE)) for l in lengths]
    return names, lengths


# ---------------------------------------------------------------------------
# synthetic data on the device (torch is plumbing: memory + RNG + NCCL)
# ---------------------------------------------------------------------------
def gen_chrom_device(torch, dev, n_tracks, length, seed):
    """All tracks of one chromosome, concatenated track-major, on the GPU."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    mean_step = (MAX_LEN + 1) / 2 + GAP_PROB * (MAX_GAP + 1) / 2
    n = int(length / mean_step * 1.06) + 64
    starts, ends, values, counts = [], [], [], []
    chunk = max(1, min(n_tracks, int(2.5e8 // n)))
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
    return torch.cat(starts), torch.cat(ends), torch.cat(values), off


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML, every few ms;
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
                time.sleep(0.004)
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
# the reference's CPU path on a bounded sample (oracle/_ref or the oracle port)
# ---------------------------------------------------------------------------
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
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libwtref.so")) else "port"
    length = int(mbp_per_core * 1e6)
    jobs = [(kind, n_tracks, length, 1000 + i) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            t0 = time.perf_counter()
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
    ncpu = os.cpu_count() or 1
    mbp = 8.0 * (64.0 / N_TRACKS) ** 2          # ~1 s per step per core at N=64
    candidates = sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4)}, reverse=True)
    best, tried = None, []
    for i in range(max(1, args.warmup)):
        c = candidates[min(i, len(candidates) - 1)]
        r = cpu_reference_sample(N_TRACKS, mbp, c)
        tried.append((c, r["value"]))
        if best is None or r["value"] > best[1]:
            best = (c, r["value"])
    cores = best[0]
    times = [cpu_reference_sample(N_TRACKS, mbp, cores) for _ in range(args.steps)]
    value = float(np.mean([t["value"] for t in times]))
    cb = dict(times[-1]); cb["value"] = value
    cb["sample"] += "; process counts tried in warm-up: " + ", ".join("%d -> %.0f Mbp/s" % t for t in tried)
    names, lengths = genome()
    line = {"impl": "reference", "metric": "Mbp/s, mean over %d tracks" % N_TRACKS, "value": value, "unit": "Mbp/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (cores * mbp) / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(lengths, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(lengths, gpus):
    return {"workload": "mean over %d synthetic tracks, %d hg38-like chromosomes (%.2f Gbp), ~%.1e intervals/track"
                        % (N_TRACKS, len(lengths), sum(lengths) / 1e9, sum(lengths) / 250.75),
            "tracks": N_TRACKS, "genome_bp": int(sum(lengths)), "sharding": "chromosomes over %d GPU(s), LPT" % gpus,
            "l2": "inputs (%.1f GB/step) larger than L2; no flush needed" % (N_TRACKS * sum(lengths) / 250.75 * 16 / 1e9)}


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from wiggletools_b200 import _lib

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

    # resident inputs and outputs for this rank's chromosomes
    shards = []
    for ci in mine:
        s, e, v, off = gen_chrom_device(torch, dev, N_TRACKS, lengths[ci], seed=7919 * (ci + 1))
        I = int(off[-1])
        cap = min(2 * I + 1, lengths[ci] + 1)
        out_s = torch.empty(cap, dtype=torch.int32, device=dev)
        out_e = torch.empty(cap, dtype=torch.int32, device=dev)
        out_v = torch.empty(cap, dtype=torch.float64, device=dev)
        shards.append(dict(ci=ci, s=s, e=e, v=v, off=off, I=I, cap=cap, out=(out_s, out_e, out_v), L=lengths[ci]))
    torch.cuda.synchronize()
    dflt = np.zeros(N_TRACKS, np.float64)
    nout = C.c_int64(0)
    auc = C.c_double(0)
    dv = C.c_double(0)

    def step(collect=None):
        total_auc = 0.0
        for sh in shards:
            rc = lib.wb200_reduce_device(ctx, OP_CODE, N_TRACKS, N_TRACKS, 0,
                                         sh["s"].data_ptr(), sh["e"].data_ptr(), sh["v"].data_ptr(),
                                         sh["off"].ctypes.data, dflt.ctypes.data, 1, sh["L"] + 1,
                                         sh["out"][0].data_ptr(), sh["out"][1].data_ptr(), sh["out"][2].data_ptr(),
                                         sh["cap"], C.addressof(nout), C.addressof(dv), C.addressof(auc))
            if rc != 0:
                raise SystemExit("wb200_reduce_device: %s" % lib.wb200_last_error(ctx).decode())
            sh["S"] = nout.value
            total_auc += auc.value
            if collect is not None:
                st = lib.wb200_last_stats(ctx).contents
                collect.append((sh["I"], sh["S"], st.ms_count, st.ms_emit, st.kernels, st.slow_segments, st.limbs))
        t = torch.tensor([total_auc], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)   # whole-genome AUC scalar: the path's only exchange step
        return t

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    per_launch = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        auc_t = step(per_launch)
    ev1.record()
    barrier()
    sampler.stop_flag = True
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    genome_bp = float(sum(lengths))
    value = genome_bp / 1e6 / (ms_per_step / 1e3)

    # roofline of the dominant kernel (emit_kernel): algorithmic bytes 16 B per
    # input interval read + 16 B per segment written, per launch; duration from
    # CUDA events recorded around the launches on the launching stream.
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    by = sum(16.0 * i + 16.0 * s for i, s, *_ in per_launch)
    t_emit = sum(r[3] for r in per_launch) / 1e3
    t_count = sum(r[2] for r in per_launch) / 1e3
    achieved = by / t_emit / 1e9 if t_emit > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("emit_kernel_dram_bytes_per_algorithmic_byte")
    roofline = {"bound": "hbm", "kernel": DOMINANT_KERNEL, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": (traffic * by / len(per_launch)) if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_step": by / args.steps,
                "algorithmic_bytes_per_launch": by / max(1, len(per_launch)),
                "emit_ms_per_step": 1e3 * t_emit / args.steps, "cursor_ms_per_step": 1e3 * t_count / args.steps,
                "path_frac": (by / (t_emit + t_count) / 1e9) / peak if t_emit + t_count > 0 else 0.0,
                # exact-sum format the library chose from the value range: 32-bit limbs per position (1: values like
                # this generator's; 2: full float32 mantissas; 3: wider), 0 for the statistics without it
                "accumulator_limbs": max((r[6] for r in per_launch), default=0)}
    launches = sum(r[4] for r in per_launch)
    seg_local = sum(sh["S"] for sh in shards)
    int_local = sum(sh["I"] for sh in shards)
    tot = torch.tensor([seg_local, int_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    slow = sum(r[5] for r in per_launch)
    checksum = None
    if os.environ.get("WB200_BENCH_CHECKSUM"):      # development: order-sensitive digest of every output array
        acc = torch.zeros((), dtype=torch.int64, device=dev)
        for sh in shards:
            S = sh["S"]
            k = torch.arange(1, S + 1, device=dev, dtype=torch.int64)
            acc += (sh["out"][0][:S].long() * k).sum() * 3 + (sh["out"][1][:S].long() * k).sum() * 5
            acc += (sh["out"][2][:S].view(torch.int64) * k).sum()
            del k
        checksum = int(acc.item())

    # end to end through the host-buffer entry point: pinned host tracks ->
    # H2D -> kernels -> D2H of the segments, every step, timed with the wall
    # clock around synchronous calls (max over ranks).
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, dist, lib, ctx, shards, dev, world, genome_bp, max(1, min(args.steps, 3)))
        e2e["host_thread_cpus"] = int(local_cpus)     # CPUs local to this rank's GPU the host thread was pinned to (0: not pinned)

    # the device text sink on this rank's smallest shard, outside the timed region:
    # `write_bg` of the HBM-resident segments, D2H of the text included
    sink = None
    if rank == 0 and shards and not args.no_e2e and BENCH_OP in ("sum", "mean"):
        from wiggletools_b200._lib import CSinkState
        sh = min(shards, key=lambda x: x["S"])
        best = None
        for rep in range(3):
            state = CSinkState(0, 0, 0)
            text, nbytes = C.c_char_p(), C.c_int64(0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = lib.wb200_format_device(ctx, 1, names[sh["ci"]].encode(), sh["out"][0].data_ptr(), sh["out"][1].data_ptr(),
                                         sh["out"][2].data_ptr(), sh["S"], C.addressof(state), C.addressof(text), C.addressof(nbytes))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rc != 0:
                raise SystemExit("wb200_format_device: %s" % lib.wb200_last_error(ctx).decode())
            if rep and (best is None or dt < best):
                best = dt
        sink = {"mode": "write_bg", "shard": names[sh["ci"]], "records": int(state.records), "text_bytes": int(nbytes.value),
                "ms": 1e3 * best, "records_per_s": state.records / best, "d2h_GBps": nbytes.value / best / 1e9,
                "note": "wb200_format_device on HBM-resident segments, text copied to pinned host memory inside the call"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = max(1, (os.cpu_count() or 1) // 4)     # the reference scales best on a subset of a large host
        cpu = cpu_reference_sample(N_TRACKS, 12.0 * (64.0 / N_TRACKS) ** 2, cores)

    if rank == 0:
        line = {"metric": "Mbp/s, %s over %d tracks" % (BENCH_OP, N_TRACKS), "value": value, "unit": "Mbp/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(lengths, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": sampler.summary(),
                "segments": int(tot[0].item()), "intervals": int(tot[1].item()), "slow_segments": int(slow),
                "auc": float(auc_t.item())}
        if sink is not None:
            line["sink"] = sink
        if checksum is not None:
            line["checksum"] = checksum
        ph = (C.c_uint64 * 16)()
        lib.wb200_debug_phases(ctx, ph)
        if any(ph):                                   # diagnostic builds only
            names_ = ["clear", "events", "scan", "count", "compact", "lookback", "emit", "window_total", "windows", "lookback_polls", "lookback_steps"]
            line["phase_cycles_per_window"] = {n: ph[k] / max(1, ph[8]) for k, n in enumerate(names_)}
        print(json.dumps(line))
    lib.wb200_destroy(ctx)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(torch, dist, lib, ctx, shards, dev, world, genome_bp, steps):
    """The call a user of the library makes: wb200_reduce_host_batch on pinned host
    SoA tracks -> pinned host segments (H2D, kernels, D2H all inside the timed region)."""
    from wiggletools_b200._lib import CTrack, CSegments
    nsh = len(shards)
    n = len(shards[0]["off"]) - 1 if nsh else 0
    arr = (CTrack * max(1, nsh * n))()
    host = []
    h2d = 0
    for c, sh in enumerate(shards):
        hs = sh["s"].cpu().pin_memory(); he = sh["e"].cpu().pin_memory(); hv = sh["v"].cpu().pin_memory()
        for i in range(n):
            a, b = int(sh["off"][i]), int(sh["off"][i + 1])
            arr[c * n + i] = CTrack(hs.data_ptr() + 4 * a, he.data_ptr() + 4 * a, hv.data_ptr() + 8 * a, b - a, 0.0)
        host.append((hs, he, hv))
        h2d += 16 * sh["I"]
    out = (CSegments * max(1, nsh))()
    d2h = 0

    def one():
        nonlocal d2h
        if not nsh:
            return
        rc = lib.wb200_reduce_host_batch(ctx, 2, nsh, n, n, 0, C.addressof(arr), C.addressof(out))
        if rc != 0:
            raise SystemExit("wb200_reduce_host_batch: %s" % lib.wb200_last_error(ctx).decode())
        d2h = sum(16 * out[c].n for c in range(nsh))

    one()   # first call sizes the pinned result arena
    one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    tb = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tb)
    dt = float(t.item())
    return {"value": genome_bp / 1e6 / (dt / steps), "unit": "Mbp/s", "h2d_bytes_per_step": int(tb[0].item()),
            "d2h_bytes_per_step": int(tb[1].item()), "steps": steps, "ms_per_step": 1e3 * dt / steps,
            "api": "wb200_reduce_host_batch (pinned host SoA in, pinned host segments out; staging, kernels and read-back pipelined per chromosome)"}


if __name__ == "__main__":
    main()
