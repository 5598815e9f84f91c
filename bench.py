#!/usr/bin/env python
"""Benchmark of the paint -> FFT -> P(k) multipoles hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c2]

Default workload at every N: BASELINE.json configs[3] ("C4", the north-star configuration) --
1e9 uniform particles, PCS on a 2048^3 mesh, P0/P2/P4 in kF-wide bins -- as ONE problem sharded
over the N GPUs in x-slabs (strong scaling): every rank paints its slab's particles, ghost planes
go round a ring, the R2C FFT is 2-D local + fused transposing peer-store over NVLink + 1-D local,
the binned sums are all-reduced.  At N=1 the whole 2048^3 mesh lives on the one GPU (147 GB) and the
step is the single-GPU pipeline with the pencil FFT plan (three contiguous 1-D passes).  A "step" is one full pass
particles -> mesh -> delta_k -> multipoles.

  value  : whole-job Gparticles/s (1e9 / step time, max over ranks) with the catalogue resident in HBM
  e2e    : the same through the public API with HOST (pinned) catalogues: H2D copy of every rank's
           particles and D2H read of the multipoles inside the timed region, every step
  roofline / kernels : per-kernel device time from CUDA events the library records around each of
           its launches (second pass over the same K steps, rank 0), algorithmic bytes from DESIGN.md,
           peak from MEASURED_PEAKS.json
  stages / transpose : per-stage device time (max over ranks); NVLink GB/s of the transposing
           peer-store against the box's own peer-copy rate (straight copy kernel, same bytes)
  check  : parity inside the run -- the sharded pipeline on a 512^3 replica of the workload (same
           particles per cell) against the single-rank pipeline: mode counts equal, max |dP|/P0
  cpu_baseline : the C/NumPy restatement of the reference (oracle/) on the host cores, bounded sample,
           rank 0 at N=1 only; serial (as XLA-CPU's scatter) and best-effort multi-core legs

`--workload c2` keeps round 1's line: BASELINE.json configs[1] (1e8 lognormal particles, TSC, 512^3),
one independent realisation per GPU (weak scaling, no data-path collective).

--impl reference times the CPU restatement as its own arm (JAX, and therefore the reference itself,
cannot run offline on this image).  That arm imports nothing of the product.
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
MAS = {2: "CIC", 3: "TSC", 4: "PCS"}

# ---- workload C2 (BASELINE.json configs[1])
C2 = dict(name="C2", n_part=100_000_000, n_mesh=512, box=2000.0, order=3, seed=5, gen_grid=256)
# ---- workload C4 (BASELINE.json configs[3]): ONE mesh slab-sharded over the GPUs (strong scaling)
C4 = dict(name="C4", n_part=1_000_000_000, n_mesh=2048, box=2000.0, order=4, seed=42)
# 512^3 replica of C4 with the same particles per cell: the in-run parity check
# and the mesh the bounded CPU sample is timed on
REPLICA = dict(n_mesh=512, n_part=15_625_000)
CPU_SAMPLE = 1_000_000          # particles painted per CPU step (both CPU legs use this same sample)


def k_edges_for(box, n_mesh):
    kF = 2.0 * math.pi / box
    return np.arange(kF, math.pi * n_mesh / box, kF).astype(np.float32)


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
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        reasons = sorted({k for _, _, rs in self.rows for k, b in bits.items() if rs & b})
        sm = [r[0] for r in self.rows]
        pw = [r[1] for r in self.rows]
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm   # samples taken under load
        return {"sm_mhz": statistics.median(load), "sm_mhz_min": min(load), "sm_max_mhz": self.max_sm,
                "power_w_max": max(pw), "samples": len(sm), "samples_under_load": len(load), "reasons": reasons}


# ------------------------------------------------------------------------------- CPU legs
# Nothing below imports the product: only NumPy, SciPy and oracle/ (the C + NumPy restatement).
def cpu_workload(wl):
    """(mesh side the CPU sample runs on, factor from that mesh to the workload's).  C4's 2048^3 float32
    mesh + its transform do not fit a bounded CPU step, so the sample runs on the 512^3 replica and the
    per-mesh costs are scaled: FFT by N^3 log N, density contrast and binning by N^3."""
    if wl["n_mesh"] <= 512:
        return wl["n_mesh"], 1.0, 1.0
    n_s, n = REPLICA["n_mesh"], wl["n_mesh"]
    vol = (n / n_s) ** 3
    return n_s, vol, vol * math.log2(n) / math.log2(n_s)


def cpu_sample(wl, n_sample):
    """The bounded sample of the workload's catalogue BOTH CPU legs paint (seeded NumPy, no product code):
    C4 uniform in the box; C2 half uniform, half in Gaussian blobs (the CPU painter's cost per particle
    does not depend on the clustering model: every deposit misses the caches on a >= 512^3 mesh)."""
    rng = np.random.default_rng(wl["seed"])
    box = np.float32(wl["box"])
    if wl["name"] == "C4":
        p = rng.random((n_sample, 3), dtype=np.float32) * box
    else:
        nu = n_sample // 2
        centres = rng.random((64, 3)) * wl["box"]
        blob = centres[rng.integers(0, 64, n_sample - nu)] + rng.normal(size=(n_sample - nu, 3)) * 0.02 * wl["box"]
        p = (np.concatenate([rng.random((nu, 3)) * wl["box"], blob]) % wl["box"]).astype(np.float32)
    p[p >= box] = 0.0
    return tuple(np.ascontiguousarray(p[:, i]) for i in range(3))


def cpu_reference_step(sample, wl, multicore=False):
    """One pass of the CPU restatement on `sample`: paint, density contrast, FFT, binning (stage seconds)."""
    from oracle import cport
    x, y, z = sample
    n, _, _ = cpu_workload(wl)
    box = wl["box"]
    k_edges = k_edges_for(box, n)
    mesh0 = np.zeros((n, n, n), np.float32)
    mesh0.fill(0.0)                                       # touch the pages outside the timed region
    paint = ((lambda *c: cport.paint_mt(mesh0, *c, None, 0.0, 0.0, 0.0, box, n, True, order=wl["order"])) if multicore else
             (lambda *c: cport.paint(mesh0, *c, None, 0.0, 0.0, 0.0, box, n, True, order=wl["order"], compat="fixed")))
    # the painters are functional like the reference (copy of the input mesh first): that copy is per-MESH work,
    # timed alone with an empty catalogue so that only the per-particle part is extrapolated to the catalogue
    empty = np.zeros(0, np.float32)
    tc = time.perf_counter()
    paint(empty, empty, empty)
    t0 = time.perf_counter()
    rho = paint(x, y, z)
    t1 = time.perf_counter()
    delta = rho / rho.mean() - np.float32(1.0)           # tests/correlations.py:49-50
    t2 = time.perf_counter()
    tt = {}
    k3d, pk, nm = cport.powspec(delta, box, k_edges, mas_order=wl["order"], workers=-1, times=tt)
    copy_s = t0 - tc
    return {"paint_s": max(t1 - t0 - copy_s, 1e-9), "contrast_s": t2 - t1 + copy_s, "fft_s": tt["fft_s"], "bin_s": tt["bin_s"]}, pk


def cpu_throughput(times, wl, n_sample):
    """Whole-workload Gparticles/s extrapolated from the bounded sample: painting scales with the
    particle count; density contrast, FFT and binning are per-mesh costs (scaled from the sample mesh to
    the workload's mesh by cpu_workload())."""
    _, vol, fft = cpu_workload(wl)
    full = times["paint_s"] * (wl["n_part"] / n_sample) + (times["contrast_s"] + times["bin_s"]) * vol + times["fft_s"] * fft
    return wl["n_part"] / full / 1e9, full


def cpu_sample_desc(wl, n_sample, cores, multicore):
    n_s, vol, fft = cpu_workload(wl)
    painter = (f"multi-core float32 C painter (OpenMP atomics, {cores} threads)" if multicore
               else "serial float32 C painter (1 thread, as XLA-CPU's scatter)")
    mesh = (f"a {n_s}^3 replica mesh (per-mesh costs scaled to {wl['n_mesh']}^3: contrast and binning x{vol:g}, FFT x{fft:.1f})"
            if vol != 1.0 else f"the full {n_s}^3 mesh")
    return (f"{n_sample:.3g} of {wl['n_part']:.3g} particles painted ({MAS[wl['order']]}) with the {painter} on {mesh}, "
            f"density contrast + scipy rfftn ({cores} threads) + serial float32 binning; "
            f"value = n_part / (paint_s * n_part/n_sample + per-mesh seconds)")


def cpu_baseline_block(wl, n_sample):
    """cpu_baseline of the GPU arm's line: the faithful serial leg (value) and the multi-core leg."""
    from oracle import build as obuild
    from oracle import cport
    obuild.build()
    sample = cpu_sample(wl, n_sample)
    cores = len(os.sched_getaffinity(0))
    times, _ = cpu_reference_step(sample, wl)
    v, full_s = cpu_throughput(times, wl, n_sample)
    times_mt, _ = cpu_reference_step(sample, wl, multicore=True)
    v_mt, full_mt = cpu_throughput(times_mt, wl, n_sample)
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_desc(wl, n_sample, cores, False),
            "stage_s": times, "extrapolated_full_workload_s": full_s,
            "multicore": {"value": v_mt, "unit": UNIT, "cores": cport.num_threads(),
                          "sample": cpu_sample_desc(wl, n_sample, cport.num_threads(), True), "stage_s": times_mt,
                          "extrapolated_full_workload_s": full_mt},
            "note": "value = the faithful leg (XLA-CPU's scatter is single-threaded, so this is what the reference's "
                    "JAX-on-CPU path does); multicore = best-effort threaded painter, reported next to it"}


def workload_text(wl, world):
    if wl["name"] == "C4":
        return (f"C4: {wl['n_part']:.3g} uniform particles (every rank generates its x-slab's share), {MAS[wl['order']]} on "
                f"{wl['n_mesh']}^3, box {wl['box']:g} Mpc/h, slab-sharded paint + distributed R2C FFT + P0/P2/P4 in kF-wide "
                f"bins up to k_Nyquist" + (" (one rank: whole mesh on one GPU, pencil FFT plan)" if world == 1 else ""))
    return (f"C2: lognormal mock, {wl['n_part']:.3g} particles in redshift space (LOS z), {MAS[wl['order']]} on "
            f"{wl['n_mesh']}^3, box {wl['box']:g} Mpc/h, P0/P2/P4 in kF-wide bins up to k_Nyquist")


def parallelism_text(wl, world):
    if wl["name"] == "C4":
        return (f"{world} x-slab(s) of ONE mesh: ring halo exchange of the ghost planes + fused transposing peer-store "
                f"(all-to-all over NVLink) + allreduce of the binned sums, every step" if world > 1
                else "1 rank: whole mesh on one GPU, no exchange")
    return f"{world} independent realisation(s), one per GPU; no data-path collective"


def run_reference_arm(a, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import build as obuild
    obuild.build()
    n_sample = a.cpu_sample or CPU_SAMPLE
    sample = cpu_sample(wl, n_sample)
    for _ in range(a.warmup):
        cpu_reference_step(sample, wl)
    acc = {"paint_s": 0.0, "contrast_s": 0.0, "fft_s": 0.0, "bin_s": 0.0}
    t0 = time.perf_counter()
    for _ in range(a.steps):
        t, _ = cpu_reference_step(sample, wl)
        for k in acc:
            acc[k] += t[k] / a.steps
    wall = (time.perf_counter() - t0) / a.steps
    value, full_s = cpu_throughput(acc, wl, n_sample)
    cores = len(os.sched_getaffinity(0))
    t_mt, _ = cpu_reference_step(sample, wl, multicore=True)
    v_mt, full_mt = cpu_throughput(t_mt, wl, n_sample)
    strong = wl["name"] == "C4"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(wl, a.gpus), "n_part_total": wl["n_part"], "n_mesh": wl["n_mesh"],
                   "box_size": wl["box"], "mas_order": wl["order"],
                   "parallelism": "CPU: one process on rank 0's host cores (the other ranks exit)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": cpu_sample_desc(wl, n_sample, cores, False),
                         "stage_s": acc, "extrapolated_full_workload_s": full_s,
                         "multicore": {"value": v_mt, "unit": UNIT, "stage_s": t_mt, "extrapolated_full_workload_s": full_mt,
                                       "sample": cpu_sample_desc(wl, n_sample, cores, True)}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "JAX is not installable offline, so the reference's own JAX-on-CPU path cannot run; this arm times "
                "oracle/ (C + NumPy restatement, pinned to the reference source run under oracle/jaxshim.py). value is the "
                "FAITHFUL leg: serial float32 scatter (XLA-CPU's scatter is single-threaded) + FFT on all cores; "
                "cpu_baseline.multicore is the best-effort threaded painter, not used in value",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------- GPU arms
ALGO_BYTES = {
    # kernel name -> algorithmic bytes per launch (DESIGN.md "Kernels"); np_ = particles of the launch,
    # cells = mesh cells of the launch (one rank's share), modes = stored modes of the launch
    "paint_atomic": lambda np_, cells, modes: np_ * 12 + 4 * cells,
    "paint_tile": lambda np_, cells, modes: np_ * 16 + 4 * cells,
    "bucket_count": lambda np_, cells, modes: np_ * 12,
    "bucket_scatter": lambda np_, cells, modes: np_ * 12 + np_ * 16,
    "bucket_fine": lambda np_, cells, modes: np_ * 16 + np_ * 16,
    "pk_fold_bin": lambda np_, cells, modes: 8 * modes,
    "memset": lambda np_, cells, modes: 4 * cells,
}

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
NCU_KERNEL = {"paint_tile": ("paint_tile_fx_kernel",), "bucket_scatter": ("coarse_scatter_kernel",),
              "bucket_fine": ("fine_staged_kernel", "fine_scatter_kernel"),
              "bucket_count": ("bucket_count_smem_kernel", "bucket_count_global_kernel"),
              "pk_fold_bin": ("pk_fold_bin_kernel", "pk_bin_xfast_kernel")}


def ncu_traffic(kernel, wl, world):
    """Per-launch DRAM bytes from the committed ncu captures: C2 = profiles/r2_ncu_dram_traffic_c2.json, C4 on one GPU
    = profiles/r2_ncu_dram_traffic_c4_1gpu.json, one C4 rank of the 8-GPU decomposition =
    profiles/r2_ncu_dram_traffic_c4_rank.json (when present)."""
    name = None
    if wl["name"] == "C2" and (wl["n_part"], wl["n_mesh"], wl["order"]) == (C2["n_part"], C2["n_mesh"], C2["order"]):
        name = "r2_ncu_dram_traffic_c2.json"
        if not os.path.exists(os.path.join(ROOT, "profiles", name)):
            name = "r1_ncu_dram_traffic_c2.json"
    elif wl["name"] == "C4" and world in (1, 8) and (wl["n_part"], wl["n_mesh"], wl["order"]) == (C4["n_part"], C4["n_mesh"], C4["order"]):
        name = "r2_ncu_dram_traffic_c4_rank.json" if world == 8 else "r2_ncu_dram_traffic_c4_1gpu.json"
    if name is None:
        return None, None
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            table = json.load(f)
        for k in NCU_KERNEL.get(kernel, ()):
            if k in table:
                return table[k], "profiles/" + name
        return None, None
    except Exception:
        return None, None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_table(prof, steps, ms_step, np_launch, cells_launch, modes_launch, wl, world):
    from jax_powspec_b200 import _lib
    peak, peak_src = measured_peak()
    kernels = {}
    for name, (cnt, tot_ms) in prof.items():
        per = tot_ms / cnt
        entry = {"launches_per_step": cnt / steps, "ms_per_launch": per, "share_of_step": tot_ms / steps / ms_step}
        if name in ALGO_BYTES:
            b = ALGO_BYTES[name](np_launch, cells_launch, modes_launch)
            entry["algorithmic_bytes"] = b
            entry["ncu_dram_bytes"] = ncu_traffic(name, wl, world)[0]
            entry["achieved_gbs"] = b / per / 1e6
            entry["frac_of_peak"] = b / per / 1e6 / peak
        kernels[name] = entry
    ours = {k: v for k, v in kernels.items() if k not in _lib.LIBRARY_KERNELS and "algorithmic_bytes" in v}
    dom = max(ours, key=lambda k: ours[k]["ms_per_launch"] * ours[k]["launches_per_step"]) if ours else None
    roofline = None
    if dom:
        d = ours[dom]
        traffic, src = ncu_traffic(dom, wl, world)
        roofline = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": d["frac_of_peak"], "traffic": traffic,
                    "traffic_source": (src + " (ncu --set full)") if src else None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": d["algorithmic_bytes"], "ms_per_launch": d["ms_per_launch"]}
    return kernels, roofline


def dist_setup():
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
    return torch, dist, world, rank, local, dev


def run_c2_arm(a, wl):
    """--workload c2: one independent realisation per GPU (weak scaling)."""
    torch, dist, world, rank, local, dev = dist_setup()
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

    kernels, roofline = kernel_table(prof, a.steps, ms, npart, n ** 3, n * n * (n // 2 + 1), wl, world)
    if "cufft_r2c" in kernels:
        kernels["cufft_r2c"].update(algorithmic_bytes=24 * n ** 3, achieved_gbs=24 * n ** 3 / kernels["cufft_r2c"]["ms_per_launch"] / 1e6)
    cpu = None
    if world == 1 and not a.no_cpu:
        os.sched_setaffinity(0, all_cpus)          # the CPU baseline may use every host core
        cpu = cpu_baseline_block(wl, a.cpu_sample or CPU_SAMPLE)

    line = {
        "metric": METRIC, "value": world * npart / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(wl, world), "n_part_per_gpu": npart, "n_mesh": n, "box_size": box,
                   "mas_order": order, "n_kbins": int(len(k_edges) - 1), "paint_method": a.method,
                   "particle_order": "random (catalogue shuffled)",
                   "cache": f"inputs {12 * npart / 1e6:.0f} MB + mesh {4 * n ** 3 / 1e6:.0f} MB per step, larger than the 126 MB L2 (no flush needed)",
                   "parallelism": parallelism_text(wl, world)},
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


def slab_catalog(torch, wl, rank, world, dev, n_total=None, seed=None):
    """Uniform particles of rank `rank`'s x-slab (BASELINE configs[3]: generated in place, no routing)."""
    box = wl["box"]
    nloc = (n_total or wl["n_part"]) // world
    g = torch.Generator(device=dev)
    g.manual_seed((wl["seed"] if seed is None else seed) + rank)
    w_slab = box / world
    hi = float(np.nextafter(np.float32((rank + 1) * w_slab), np.float32(0)))
    x = (torch.rand(nloc, generator=g, device=dev) * w_slab + rank * w_slab).clamp_(max=hi)
    y = torch.rand(nloc, generator=g, device=dev) * box
    z = torch.rand(nloc, generator=g, device=dev) * box
    for t in (y, z):
        t[t >= box] = 0.0
    return x, y, z


def replica_check(torch, dist, a, wl, world, rank, dev):
    """Parity inside the bench run: the SAME pipeline class on a 512^3 replica of the
    workload (same particles per cell, same order), every rank painting its slab's particles, against the
    single-rank pipeline run on rank 0 over the gathered catalogue.  With one rank the comparison is the
    bucketed tile painter against the plain atomic painter (two independent deposits)."""
    import jax_powspec_b200 as jps
    from jax_powspec_b200.slab import SlabPipeline
    n, box, order = REPLICA["n_mesh"], wl["box"], wl["order"]
    rep = dict(wl, n_mesh=n, n_part=REPLICA["n_part"])
    ke = k_edges_for(box, n)
    x, y, z = slab_catalog(torch, rep, rank, world, dev, seed=wl["seed"] + 1000)
    pipe = SlabPipeline(n, box, ke, order=order, compat="fixed", method="sorted", transport=a.transport,
                        overlap=not a.no_overlap, layout=a.layout, pipeline=a.pipeline, fft=a.slab_fft)
    pipe._force_chunks = True                       # take the chunked FFT / transfer overlap path like the big mesh
    k3d, pk, nm = (t.clone() for t in pipe(x, y, z))
    out = {"mesh": n, "n_part": REPLICA["n_part"], "particles_per_cell": REPLICA["n_part"] / n ** 3}
    if world > 1:
        parts = [torch.empty_like(x) for _ in range(world)]
        full = []
        for t in (x, y, z):
            dist.all_gather(parts, t)
            full.append(torch.cat(parts))
    else:
        full = [x, y, z]
    if rank == 0:
        ref = jps.PaintPowspec(n, box, ke, order=order, compat="fixed", method="atomic" if world == 1 else "sorted", device=dev)
        k1, pk1, nm1 = ref(*full)
        err = (pk - pk1).abs() / pk1[:, :1].abs()
        out.update(against=("single-rank pipeline on the gathered catalogue" if world > 1
                            else "plain atomic painter (same pipeline otherwise)"),
                   counts_equal=bool(torch.equal(nm, nm1)), k_equal=bool(torch.equal(k3d, k1)),
                   max_rel_P0=float(err.max().item()), tolerance=1e-5, ok=bool(torch.equal(nm, nm1) and err.max().item() <= 1e-5),
                   modes_total=float(nm.sum().item()))
    pipe.close()
    return out


def run_c4_arm(a, wl):
    """Default: 1e9 uniform particles, PCS on 2048^3, ONE mesh sharded in x-slabs over the ranks
    (jax_powspec_b200/slab.py): halo exchange + transposing peer-store + allreduce are inside the step."""
    torch, dist, world, rank, local, dev = dist_setup()
    from jax_powspec_b200 import _lib
    from jax_powspec_b200.dist import bind_near_gpu
    from jax_powspec_b200.slab import SlabHostPipeline, SlabPipeline, halo_exchange_add, transpose_all_to_all
    all_cpus = os.sched_getaffinity(0)
    numa = bind_near_gpu(local) if not a.no_numa_bind else {"bound": False, "disabled": True}
    n, box, order = wl["n_mesh"], wl["box"], wl["order"]
    nloc = wl["n_part"] // world
    nz = n // 2 + 1
    x, y, z = slab_catalog(torch, wl, rank, world, dev)
    pipe = SlabPipeline(n, box, k_edges_for(box, n), order=order, compat="fixed", method=a.method, transport=a.transport,
                        overlap=not a.no_overlap, layout=a.layout, pipeline=a.pipeline, fft=a.slab_fft)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pipe_pipelined = pipe.local is None and pipe.pipeline and pipe._can_pipeline(nloc)
    for _ in range(max(a.warmup, 3)):
        pipe(x, y, z)
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.profile_enable(False)
    _lib.profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    torch.cuda.nvtx.range_push("jps_timed")
    e0.record()
    for _ in range(a.steps):
        k3d, pk, nm = pipe(x, y, z)
    e1.record()
    sync()
    torch.cuda.nvtx.range_pop()
    ms = max_over_ranks(e0.elapsed_time(e1) / a.steps)
    clocks = sampler.stop() if rank == 0 else None
    counts = _lib.profile_snapshot()
    launches = sum(c for k, (c, _) in counts.items() if k not in _lib.LIBRARY_KERNELS)
    pk_first = [float(v) for v in pk[:3, 0].cpu()]
    nm_first = [float(v) for v in nm[:3].cpu()]
    if a.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms, "gpu_launches": int(launches)}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- per-kernel device time (rank 0's view), same K steps
    _lib.profile_reset()
    _lib.profile_enable(True)
    for _ in range(a.steps):
        pipe(x, y, z)
    torch.cuda.synchronize()
    prof = _lib.profile_snapshot()
    _lib.profile_enable(False)
    sync()

    # ---- per-stage device time, max over ranks (sharded path only: one rank is a single fused call)
    stages, transpose = {}, None
    if pipe.local is None:
        def timed(name, fn):
            sync()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(); fn(); s1.record(); sync()
            stages[name] = max_over_ranks(s0.elapsed_time(s1))

        timed("paint", lambda: pipe.stage_paint(x, y, z))
        if world > 1:
            timed("halo_exchange", lambda: halo_exchange_add(pipe.mesh, pipe.nxl))
        timed("fft_yz_transpose", pipe.stage_fft_yz_pack)
        if world > 1 and pipe.transport == "nccl":
            timed("all_to_all", lambda: transpose_all_to_all(pipe.buf_b, pipe.buf_a))
        timed("fft_x", pipe.stage_fft_x)
        timed("bin_partial", lambda: pipe.stage_partial(True))
        if world > 1:
            timed("allreduce_sums", lambda: dist.all_reduce(pipe.sums))
            a2a = (world - 1) / world ** 2 * 8 * n * n * nz
            transpose = {"bytes_leaving_each_rank": a2a, "transport": pipe.transport, "layout": "xfast" if pipe.xfast else "xslow",
                         "fused_stage_ms": stages["fft_yz_transpose"], "nvlink_nominal_gbs_per_direction": 900.0}
            timed("fft_yz_alone", pipe.stage_fft_yz_only)
            transpose["fft_yz_alone_ms"] = stages.pop("fft_yz_alone")
            if pipe.transport == "p2p":
                t_copy = max_over_ranks(pipe.probe_transpose(False))
                t_store = max_over_ranks(pipe.probe_transpose(True)) if pipe.xfast else t_copy
                transpose.update(
                    peer_copy_ms=t_copy, peer_copy_gbs=a2a / t_copy / 1e6,
                    store_alone_ms=t_store, nvlink_gbs=a2a / t_store / 1e6,
                    frac_of_peer_copy=t_copy / t_store, frac_of_nominal=a2a / t_store / 1e6 / 900.0,
                    hidden_ms=transpose["fft_yz_alone_ms"] + t_store - stages["fft_yz_transpose"],
                    note="peer_copy = the straight contiguous peer-store kernel moving the same bytes (this box's all-to-all "
                         "copy rate); store_alone = the transposing peer-store the pipeline uses, run alone; hidden_ms = "
                         "fft_yz_alone + store_alone - fused stage (what the chunked overlap hides)")
            else:
                transpose.update(nvlink_gbs=a2a / stages["all_to_all"] / 1e6, frac_of_nominal=a2a / stages["all_to_all"] / 1e6 / 900.0)

    if a.quick_kernels:
        if rank == 0:
            print(json.dumps({"ms_per_step": ms, "pipelined": bool(pipe_pipelined), "stages_ms": stages, "transpose": transpose,
                              "kernels": {k: {"ms_per_launch": t / c, "launches_per_step": c / a.steps} for k, (c, t) in prof.items()}}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- end to end through the public API with host buffers (every rank its own slab's catalogue)
    xh, yh, zh = (torch.empty(nloc, dtype=torch.float32, pin_memory=True) for _ in range(3))
    for h, d in ((xh, x), (yh, y), (zh, z)):
        h.copy_(d)
    del x, y, z
    torch.cuda.empty_cache()
    host = SlabHostPipeline(pipe, nloc)
    for _ in range(2):
        host(xh, yh, zh)
    sync()
    e0.record()
    for _ in range(a.steps):
        out = host(xh, yh, zh)
    e1.record()
    sync()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / a.steps)
    h2d = 3 * nloc * 4 * world
    d2h = int(sum(np.asarray(o).nbytes for o in out)) * world
    e2e_pk_first = [float(v) for v in out[1][:3, 0]]
    # the copies ALONE, all ranks at once: the host-side ceiling of this box for the e2e number
    stage = torch.empty(nloc, dtype=torch.float32, device=dev)
    sync()
    e0.record()
    for h in (xh, yh, zh):
        stage.copy_(h, non_blocking=True)
    e1.record()
    sync()
    ms_copy = max_over_ranks(e0.elapsed_time(e1))
    del stage
    del host, xh, yh, zh
    pipe.close()
    del pipe
    torch.cuda.empty_cache()

    # ---- parity inside the run
    check = replica_check(torch, dist, a, wl, world, rank, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    kernels, roofline = kernel_table(prof, a.steps, ms, nloc, n ** 3 // world, n * n * nz // world, wl, world)
    fft_parts = [k for k in ("cufft_r2c", "cufft_c2c_y", "cufft_c2c_x", "fft_transpose") if k in kernels]
    fft = None
    if fft_parts:
        b = 24 * n ** 3 // world
        t = sum(kernels[k]["ms_per_launch"] * kernels[k]["launches_per_step"] for k in fft_parts)
        fft = {"parts": fft_parts, "ms_per_step_rank0": t, "algorithmic_bytes_per_step": b, "achieved_gbs": b / t / 1e6,
               "frac_of_peak": b / t / 1e6 / measured_peak()[0],
               "note": "forward R2C transform of this rank's share against the 3-pass model (3 axes x read + write = 24 B per "
                       "cell); cuFFT launches (library) + our transposing kernels of the pencil plan"}
    paint_ms = sum(kernels[k]["ms_per_launch"] * kernels[k]["launches_per_step"]
                   for k in ("bucket_count", "bucket_scan", "bucket_scatter", "bucket_fine", "paint_tile", "paint_atomic") if k in kernels)
    peak, _ = measured_peak()
    paint_bytes = 12 * nloc + 4 * n ** 3 // world
    painting = {"ms_per_step_rank0": paint_ms, "gparticles_per_s_rank0": nloc / paint_ms / 1e6 if paint_ms else None,
                "algorithmic_bytes": paint_bytes, "frac_of_peak": paint_bytes / paint_ms / 1e6 / peak if paint_ms else None,
                "note": "painting as ONE op (SURVEY 8d: bucketing is overhead, not bytes): 12 B/particle + 4 B/cell over "
                        "count + scans + coarse + fine + deposit"}
    e2e_bytes = 16 * wl["n_part"] + 32 * n ** 3
    cpu = None
    if world == 1 and not a.no_cpu:
        os.sched_setaffinity(0, all_cpus)
        cpu = cpu_baseline_block(wl, a.cpu_sample or CPU_SAMPLE)

    line = {
        "metric": METRIC, "value": wl["n_part"] / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(wl, world), "n_part_total": wl["n_part"], "n_part_per_gpu": nloc, "n_mesh": n,
                   "box_size": box, "mas_order": order, "n_kbins": int(len(k_edges_for(box, n)) - 1), "paint_method": a.method,
                   "particle_order": "random (uniform in the rank's slab)",
                   "cache": f"inputs {12 * nloc / 1e9:.1f} GB + mesh {4 * n ** 3 / world / 1e9:.1f} GB per rank and step, "
                            f"far larger than the 126 MB L2 (no flush needed)",
                   "parallelism": parallelism_text(wl, world) + ("; deposit, halo exchange, 2-D FFT and peer transfer run as one "
                                                                "pipeline over pieces of planes, so the stage table (stages run "
                                                                "back to back, one at a time) adds up to more than the step"
                                                                if pipe_pipelined else "")},
        "clocks": clocks, "host_affinity": numa,
        "e2e": {"value": wl["n_part"] / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_gbs_per_gpu": h2d / world / ms_e2e / 1e6, "P0_first_bins": e2e_pk_first,
                "h2d_copies_alone_ms": ms_copy, "h2d_copies_alone_gbs_per_gpu": h2d / world / ms_copy / 1e6,
                "note": "h2d_copies_alone = the same pinned host -> device copies with nothing else running, all ranks at "
                        "once: the host memory / PCIe ceiling of this box under the e2e step"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "whole_step": {"algorithmic_bytes": e2e_bytes, "frac_of_peak_all_gpus": e2e_bytes / ms / 1e6 / (peak * world)},
        "painting": painting,
        "fft": fft,
        "kernels": kernels,
        "stages_ms_max_over_ranks": stages or None,
        "transpose": transpose,
        "pipelined": bool(pipe_pipelined),
        "cpu_baseline": cpu,
        "check": dict(check, P0_first_bins=pk_first, Nmodes_first_bins=nm_first),
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
    ap.add_argument("--n-part", type=float, default=None, help="override the particle count (smoke runs)")
    ap.add_argument("--n-mesh", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="c4: do not overlap the 2-D FFT with the peer transfer")
    ap.add_argument("--layout", default="auto", choices=["auto", "xfast", "xslow"], help="c4: layout of the transposed shard")
    ap.add_argument("--slab-fft", default="auto", choices=["auto", "pencil", "cufft2d"],
                    help="c4: form of the 2-D transform of a rank's planes (auto: pencil from 6 GB of planes per rank)")
    ap.add_argument("--pipeline", action="store_true", help="c4: deposit, halo, FFT and transfer as one pipeline over pieces of planes")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to the CPUs nearest its GPU")
    ap.add_argument("--quick", action="store_true", help="stop after the device-resident timed region (ncu runs)")
    ap.add_argument("--quick-kernels", action="store_true", help="stop after the per-kernel / per-stage pass (sweeps)")
    ap.add_argument("--workload", default="c4", choices=["c2", "c4"],
                    help="c4 (default): one slab-sharded mesh, strong scaling (north-star config); "
                         "c2: one realisation per GPU, weak scaling")
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
    return run_c4_arm(a, wl) if a.workload == "c4" else run_c2_arm(a, wl)


if __name__ == "__main__":
    sys.exit(main())
