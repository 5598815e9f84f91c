#!/usr/bin/env python
"""Benchmark of the paint -> FFT -> P(k) multipoles hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at every N: BASELINE.json configs[1] ("C2") -- a lognormal mock of 1e8 particles in
redshift space, TSC on a 512^3 mesh, P0/P2/P4 in kF-wide bins -- one independent realisation
per GPU (weak scaling, no data-path collective; realisations are the unit the path shards on,
as in configs[4]).  A "step" is one full pass particles -> mesh -> delta_k -> multipoles.

  value  : whole-job Gparticles/s with the catalogue already resident in HBM
  e2e    : the same through the public API with HOST (pinned) catalogues: H2D copy of the
           particles and D2H read of the multipoles inside the timed region, every step
  roofline / kernels : per-kernel device time from CUDA events recorded by the library around
           each of its launches (second pass over the same K steps), algorithmic bytes from
           DESIGN.md, peak from MEASURED_PEAKS.json
  cpu_baseline : the C/NumPy restatement of the reference (oracle/) on the host cores, on a
           bounded sample of the same workload, rank 0 only

--impl reference times that CPU restatement as its own arm (JAX, and therefore the
reference itself, cannot run offline on this image).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import socket
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "paint+FFT+P(k) multipoles end-to-end throughput"
UNIT = "Gparticles/s"

# ---- workload C2 (BASELINE.json configs[1])
C2 = dict(name="C2", n_part=100_000_000, n_mesh=512, box=2000.0, order=3, seed=5, gen_grid=256)
# ---- workload C4 (BASELINE.json configs[3]): ONE mesh slab-sharded over the GPUs (strong scaling)
C4 = dict(name="C4", n_part=1_000_000_000, n_mesh=2048, box=2000.0, order=4, seed=42)


def k_edges_for(box, n_mesh):
    kF = 2.0 * math.pi / box
    return np.arange(kF, math.pi * n_mesh / box, kF).astype(np.float32)


def config_dict(a, wl, n_gpus):
    return {
        "workload": (f"{wl['name']}: lognormal mock, {wl['n_part']:.3g} particles in redshift space (LOS z), "
                     f"{ {2: 'CIC', 3: 'TSC', 4: 'PCS'}[wl['order']] } on {wl['n_mesh']}^3, box {wl['box']:g} Mpc/h, "
                     f"P0/P2/P4 in kF-wide bins up to k_Nyquist"),
        "n_part_per_gpu": wl["n_part"], "n_mesh": wl["n_mesh"], "box_size": wl["box"],
        "mas_order": wl["order"], "n_kbins": int(len(k_edges_for(wl["box"], wl["n_mesh"])) - 1),
        "paint_method": a.method,
        "particle_order": "random (catalogue shuffled)",
        "cache": f"inputs {12 * wl['n_part'] / 1e6:.0f} MB + mesh {4 * wl['n_mesh'] ** 3 / 1e6:.0f} MB per step, larger than the 126 MB L2 (no flush needed)",
        "parallelism": f"{n_gpus} independent realisation(s), one per GPU; no data-path collective",
    }


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """NVML sampling (5 ms period, own thread) of SM clock, power and throttle reasons DURING the
    timed region -- the same fields as the nvidia-smi line in B200_PROFILING.md."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:                      # pragma: no cover
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, pw, rs))
            except Exception as e:                  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(0.005)

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        reasons = sorted({k for _, _, rs in self.rows for k, b in bits.items() if rs & b})
        sm = [r[0] for r in self.rows]
        pw = [r[1] for r in self.rows]
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm   # samples taken under load
        return {"sm_mhz": statistics.median(load), "sm_mhz_min": min(load), "sm_max_mhz": self.max_sm,
                "power_w_max": max(pw), "samples": len(sm), "samples_under_load": len(load), "reasons": reasons}


# ------------------------------------------------------------------------------- CPU legs
def cpu_reference_step(sample, wl, k_edges):
    """One pass of the CPU restatement on `sample` particles; returns stage times (s)."""
    from oracle import cport
    x, y, z = sample
    n, box = wl["n_mesh"], wl["box"]
    t0 = time.perf_counter()
    rho = cport.paint(np.zeros((n, n, n), np.float32), x, y, z, None, 0.0, 0.0, 0.0, box, n, True,
                      order=wl["order"], compat="fixed")
    t1 = time.perf_counter()
    delta = rho / rho.mean() - np.float32(1.0)           # tests/correlations.py:49-50
    t2 = time.perf_counter()
    k3d, pk, nm = cport.powspec(delta, box, k_edges, mas_order=wl["order"], workers=-1)
    t3 = time.perf_counter()
    return {"paint_s": t1 - t0, "contrast_s": t2 - t1, "fft_bin_s": t3 - t2}, pk


def cpu_sample_catalog(wl, n_sample):
    import torch
    from jax_powspec_b200.mocks import lognormal_catalog
    x, y, z = lognormal_catalog(n_sample, wl["box"], n_grid=128, seed=wl["seed"], device="cpu")
    return tuple(t.numpy() for t in (x, y, z))


def cpu_throughput(times, wl, n_sample):
    """Whole-workload Gparticles/s extrapolated from the bounded sample: painting scales with the
    particle count, density contrast + FFT + binning are per-mesh costs paid once."""
    full = times["paint_s"] * (wl["n_part"] / n_sample) + times["contrast_s"] + times["fft_bin_s"]
    return wl["n_part"] / full / 1e9, full


def run_reference_arm(a, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import build as obuild
    obuild.build()
    k_edges = k_edges_for(wl["box"], wl["n_mesh"])
    n_sample = a.cpu_sample or 2_000_000
    sample = cpu_sample_catalog(wl, n_sample)
    for _ in range(a.warmup):
        cpu_reference_step(sample, wl, k_edges)
    acc = {"paint_s": 0.0, "contrast_s": 0.0, "fft_bin_s": 0.0}
    t0 = time.perf_counter()
    for _ in range(a.steps):
        t, _ = cpu_reference_step(sample, wl, k_edges)
        for k in acc:
            acc[k] += t[k] / a.steps
    wall = (time.perf_counter() - t0) / a.steps
    value, full_s = cpu_throughput(acc, wl, n_sample)
    cores = len(os.sched_getaffinity(0))
    sample_desc = (f"{n_sample:.3g} of {wl['n_part']:.3g} particles painted (serial float32 scatter, as XLA-CPU), "
                   f"full {wl['n_mesh']}^3 density contrast + scipy rfftn ({cores} threads) + serial binning; "
                   f"value = n_part / (paint_s * n_part/n_sample + contrast_s + fft_bin_s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(a, wl, a.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc,
                         "stage_s": acc, "extrapolated_full_workload_s": full_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "JAX is not installable offline, so the reference's own JAX-on-CPU path cannot run; this arm "
                "times oracle/ (C + NumPy restatement, pinned to the reference source run under oracle/jaxshim.py)",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------- GPU arm
ALGO_BYTES = {
    # kernel name -> algorithmic bytes per launch (DESIGN.md "Kernels"); np_ = particles, n = mesh side
    "paint_atomic": lambda np_, n, w: np_ * (12 + 4 * w) + 4 * n ** 3,
    "paint_tile": lambda np_, n, w: np_ * 16 + 4 * n ** 3,
    "bucket_count": lambda np_, n, w: np_ * 12,
    "bucket_scatter": lambda np_, n, w: np_ * (12 + 4 * w) + np_ * 16,
    "bucket_fine": lambda np_, n, w: np_ * 16 + np_ * 16,
    "pk_fold_bin": lambda np_, n, w: 8 * n * n * (n // 2 + 1),
    "cufft_r2c": lambda np_, n, w: 24 * n ** 3,
    "memset": lambda np_, n, w: 4 * n ** 3,
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of
# this same command (profiles/r1_ncu_full_summary.md); only meaningful for the default C2 configuration.
NCU_KERNEL = {"paint_tile": "paint_tile_fx_kernel", "bucket_scatter": "coarse_scatter_kernel",
              "bucket_fine": "fine_scatter_kernel", "bucket_count": "bucket_count_smem_kernel",
              "pk_fold_bin": "pk_fold_bin_kernel"}


def ncu_traffic(kernel, wl):
    if (wl["n_part"], wl["n_mesh"], wl["order"]) != (C2["n_part"], C2["n_mesh"], C2["order"]):
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_dram_traffic_c2.json")) as f:
            return json.load(f).get(NCU_KERNEL.get(kernel, ""))
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_gpu_arm(a, wl):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import jax_powspec_b200 as jps
    from jax_powspec_b200 import _lib
    from jax_powspec_b200.dist import bind_near_gpu
    from jax_powspec_b200.mocks import lognormal_catalog
    all_cpus = os.sched_getaffinity(0)
    numa = bind_near_gpu(local) if not a.no_numa_bind else {"bound": False, "disabled": True}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, box, npart, order = wl["n_mesh"], wl["box"], wl["n_part"], wl["order"]
    k_edges = k_edges_for(box, n)
    x, y, z = lognormal_catalog(npart, box, n_grid=wl["gen_grid"], seed=wl["seed"] + rank, device=dev)
    torch.cuda.synchronize()
    pipe = jps.PaintPowspec(n, box, k_edges, order=order, compat="fixed", method=a.method,
                            n_part_max=npart, device=dev)

    def step():
        return pipe(x, y, z)

    # ---- device-resident throughput
    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.profile_enable(False)
    _lib.profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.nvtx.range_push("jps_timed")
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    torch.cuda.nvtx.range_pop()
    ms = max_over_ranks(e0.elapsed_time(e1) / a.steps)
    clocks = sampler.stop() if rank == 0 else None
    counts = _lib.profile_snapshot()
    launches = sum(c for k, (c, _) in counts.items() if k not in _lib.LIBRARY_KERNELS)
    result_pk = pipe.pk.clone()

    if a.quick:                                   # profiler runs: only the timed region matters
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms, "gpu_launches": int(launches)}), flush=True)
        return 0

    # ---- per-kernel device time, same K steps, events recorded by the library around each launch
    _lib.profile_reset()
    _lib.profile_enable(True)
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    prof = _lib.profile_snapshot()
    _lib.profile_enable(False)

    if a.quick_kernels:                            # parameter sweeps: skip the e2e / cpu legs
        if rank == 0:
            print(json.dumps({"ms_per_step": ms, "kernels": {k: {"ms_per_launch": t / c} for k, (c, t) in prof.items()}}), flush=True)
        return 0

    # ---- end to end through the public API with host buffers
    xh, yh, zh = (t.cpu().pin_memory() for t in (x, y, z))
    del x, y, z
    torch.cuda.empty_cache()
    host = jps.HostPipeline(pipe, npart)
    for _ in range(2):
        host(xh, yh, zh)
    barrier()
    e0.record()
    for _ in range(a.steps):
        out = host(xh, yh, zh)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / a.steps)
    h2d = 3 * npart * 4
    d2h = int(sum(np.asarray(o).nbytes for o in out))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    kernels = {}
    for name, (cnt, tot_ms) in prof.items():
        per = tot_ms / cnt
        entry = {"launches_per_step": cnt / a.steps, "ms_per_launch": per, "share_of_step": tot_ms / a.steps / ms}
        if name in ALGO_BYTES:
            b = ALGO_BYTES[name](npart, n, 0)
            entry["algorithmic_bytes"] = b
            entry["ncu_dram_bytes"] = ncu_traffic(name, wl)
            entry["achieved_gbs"] = b / per / 1e6
            entry["frac_of_peak"] = b / per / 1e6 / peak
        kernels[name] = entry
    ours = {k: v for k, v in kernels.items() if k not in _lib.LIBRARY_KERNELS and "algorithmic_bytes" in v}
    dom = max(ours, key=lambda k: ours[k]["ms_per_launch"] * ours[k]["launches_per_step"]) if ours else None
    roofline = None
    if dom:
        d = ours[dom]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": d["frac_of_peak"], "traffic": ncu_traffic(dom, wl),
                    "traffic_source": "profiles/r1_ncu_full_summary.md (ncu --set full, same command)",
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": d["algorithmic_bytes"], "ms_per_launch": d["ms_per_launch"]}

    cpu = None
    if world == 1 and not a.no_cpu:
        os.sched_setaffinity(0, all_cpus)          # the CPU baseline may use every host core
        from oracle import build as obuild
        obuild.build()
        n_sample = a.cpu_sample or 10_000_000
        sample = tuple(t[:n_sample].numpy() for t in (xh, yh, zh))
        times, pk_cpu = cpu_reference_step(sample, wl, k_edges)
        v, full_s = cpu_throughput(times, wl, n_sample)
        cores = len(os.sched_getaffinity(0))
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": (f"first {n_sample:.3g} of {npart:.3g} particles painted with the serial float32 C port "
                          f"(1 thread, as XLA-CPU scatter), full {n}^3 density contrast + scipy rfftn ({cores} threads) "
                          f"+ serial binning (1 thread); value extrapolates painting linearly to the full catalogue"),
               "stage_s": times, "extrapolated_full_workload_s": full_s}

    line = {
        "metric": METRIC, "value": world * npart / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(a, wl, world),
        "clocks": clocks, "host_affinity": numa,
        "e2e": {"value": world * npart / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels": kernels,
        "cpu_baseline": cpu,
        "check": {"P0_first_bins": [float(v) for v in result_pk[:3, 0].cpu()]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_slab_arm(a, wl):
    """--workload c4: 1e9 uniform particles, PCS on 2048^3, ONE mesh sharded in x-slabs over the ranks
    (jax_powspec_b200/slab.py): halo exchange + all-to-all transpose + allreduce are inside the step."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from jax_powspec_b200 import _lib
    from jax_powspec_b200.slab import SlabPipeline, halo_exchange_add, transpose_all_to_all
    n, box, order = wl["n_mesh"], wl["box"], wl["order"]
    nloc = wl["n_part"] // world
    g = torch.Generator(device=dev)
    g.manual_seed(wl["seed"] + rank)
    w_slab = box / world
    hi = float(np.nextafter(np.float32((rank + 1) * w_slab), np.float32(0)))
    x = (torch.rand(nloc, generator=g, device=dev) * w_slab + rank * w_slab).clamp_(max=hi)
    y = torch.rand(nloc, generator=g, device=dev) * box
    z = torch.rand(nloc, generator=g, device=dev) * box
    for t in (y, z):
        t[t >= box] = 0.0
    pipe = SlabPipeline(n, box, k_edges_for(box, n), order=order, compat="fixed", method=a.method, transport=a.transport,
                        overlap=not a.no_overlap, layout=a.layout)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        pipe(x, y, z)
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.profile_enable(False); _lib.profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(a.steps):
        k3d, pk, nm = pipe(x, y, z)
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    counts = _lib.profile_snapshot()
    launches = sum(c for k, (c, _) in counts.items() if k not in _lib.LIBRARY_KERNELS and k != "misc")
    stages = {}

    def timed(name, fn):
        sync()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); fn(); s1.record(); sync()
        tt = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        stages[name] = float(tt.item())

    timed("paint", lambda: pipe.stage_paint(x, y, z))
    if world > 1:
        timed("halo_exchange", lambda: halo_exchange_add(pipe.mesh, pipe.nxl))
    timed("fft_yz_pack", pipe.stage_fft_yz_pack)
    if world > 1 and pipe.transport == "nccl":
        timed("all_to_all", lambda: transpose_all_to_all(pipe.buf_b, pipe.buf_a))
    timed("fft_x", pipe.stage_fft_x)
    timed("bin_partial", lambda: pipe.stage_partial(True))
    _lib.profile_reset(); _lib.profile_enable(True)
    pipe(x, y, z)
    torch.cuda.synchronize()
    prof = _lib.profile_snapshot()
    _lib.profile_enable(False)
    if rank == 0:
        peak, peak_src = measured_peak()
        a2a = (world - 1) / world ** 2 * 8 * n * n * (n // 2 + 1)
        paint_bytes = 12 * nloc + 4 * n ** 3 / world
        line = {
            "metric": METRIC, "value": wl["n_part"] / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C4: {wl['n_part']:.3g} uniform particles generated per rank inside its x-slab, "
                                   f"{ {2: 'CIC', 3: 'TSC', 4: 'PCS'}[order] } on {n}^3, slab-sharded paint + distributed R2C FFT "
                                   f"(2-D local, NCCL all-to-all, 1-D local) + multipoles",
                       "n_part_total": wl["n_part"], "n_mesh": n, "box_size": box, "mas_order": order,
                       "parallelism": f"{world} x-slabs, halo ring exchange + all-to-all + allreduce per step"},
            "clocks": clocks, "gpu_launches": int(launches), "stages_ms_max_over_ranks": stages,
            "kernels_ms_rank0": {k: round(ms_ / max(c, 1), 4) for k, (c, ms_) in prof.items()},
            "transport": pipe.transport, "layout": "xfast" if pipe.xfast else "xslow",
            "all_to_all": {"bytes_per_rank": a2a,
                           "achieved_gbs_per_rank": (a2a / stages["all_to_all"] / 1e6) if "all_to_all" in stages else None,
                           "note": "p2p transport: the transfer is inside fft_yz_pack (one fused pack + peer-store kernel)"
                                   if pipe.transport == "p2p" else "NCCL all_to_all_single",
                           "nvlink_peer_copy_ref_gbs": 770.0},
            "roofline": {"kernel": "paint (bucket + tile deposit)", "bound": "hbm", "achieved": paint_bytes / stages["paint"] / 1e6,
                         "peak": peak, "unit": "GB/s", "frac": paint_bytes / stages["paint"] / 1e6 / peak, "traffic": None,
                         "peak_source": peak_src},
            "check": {"P0_first_bins": [float(v) for v in pk[:3, 0].cpu()], "Nmodes_first_bins": [float(v) for v in nm[:3].cpu()]},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", default="auto", choices=["auto", "atomic", "sorted"])
    ap.add_argument("--n-part", type=float, default=None, help="override particles per GPU (smoke runs)")
    ap.add_argument("--n-mesh", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="c4: do not overlap the 2-D FFT with the peer transfer")
    ap.add_argument("--layout", default="auto", choices=["auto", "xfast", "xslow"], help="c4: layout of the transposed shard")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to the CPUs nearest its GPU")
    ap.add_argument("--quick", action="store_true", help="stop after the device-resident timed region (ncu runs)")
    ap.add_argument("--quick-kernels", action="store_true", help="stop after the per-kernel pass (sweeps)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2: one realisation per GPU (default, weak scaling); c4: one slab-sharded mesh (strong scaling)")
    ap.add_argument("--order", type=int, default=None)
    ap.add_argument("--transport", default="auto", choices=["auto", "p2p", "nccl"],
                    help="c4: how the slab transpose crosses GPUs (fused peer-store kernel or NCCL all-to-all)")
    a = ap.parse_args()
    wl = dict(C4 if a.workload == "c4" else C2)
    if a.order:
        wl["order"] = a.order
    if a.n_part:
        wl["n_part"] = int(a.n_part)
    if a.n_mesh:
        wl["n_mesh"] = a.n_mesh
    if a.impl == "reference":
        return run_reference_arm(a, wl)
    if a.gpus > 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_slab_arm(a, wl) if a.workload == "c4" else run_gpu_arm(a, wl)


if __name__ == "__main__":
    sys.exit(main())
