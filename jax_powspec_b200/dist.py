"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The path shards without communication on *realisations* (BASELINE.json configs[4]: a covariance
batch of independent mocks; also what ``bench.py --gpus N`` runs): rank r takes realisations
r, r+W, r+2W, ...; the only exchange is the final gather of the (nbins x 3) multipole rows.
The helpers below hold that host logic; they never touch the compute path, so they are testable
with the gloo backend on CPU (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import os
from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size) from the initialised process group, else (0, 1)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str | None = None, device: torch.device | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank).  No-op for a single process."""
    w = int(os.environ.get("WORLD_SIZE", "1"))
    r = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    if w > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=r, world_size=w, **kw)
    return r, w, lr


def bind_near_gpu(local_rank: int | None = None) -> dict:
    """Pin the calling process to the CPU cores NVML reports as closest to its GPU (same NUMA
    node / PCIe root), so that the pinned host buffers it allocates next are first-touched on that
    node.  With one process per GPU and 8 GPUs on two sockets, unbound ranks stage half of the
    host->device traffic across the socket interconnect.  Best effort: returns what was done and
    never raises (containers with a restricted cpuset keep their affinity)."""
    info = {"bound": False}
    try:
        import pynvml
        if local_rank is None:
            local_rank = torch.cuda.current_device()
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        before = sorted(os.sched_getaffinity(0))
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        ideal = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (int(wd) >> b) & 1}
        target = sorted(ideal & set(before))
        info.update(pci=bus, cpus_before=len(before), cpus_ideal=len(ideal))
        if target and len(target) < len(before):
            os.sched_setaffinity(0, target)
            info.update(bound=True, cpus_after=len(target), first_cpu=target[0])
        else:
            info.update(cpus_after=len(before))
    except Exception as e:  # noqa: BLE001 -- diagnostics only
        info["error"] = f"{type(e).__name__}: {e}"[:200]
    return info


def shard_indices(n_items: int, rank: int | None = None, world_size: int | None = None) -> list[int]:
    """Round-robin shard of range(n_items): rank r owns r, r+W, r+2W, ... (balanced to +-1)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(rank, n_items, world_size))


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """Device-time statistics are reported as the max over ranks (bench.py contract)."""
    _, w = world()
    if w == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local_rows: torch.Tensor, n_items: int) -> torch.Tensor | None:
    """Inverse of shard_indices: every rank passes its [n_local, ...] rows (in its shard order);
    rank 0 gets the [n_items, ...] array in global order, other ranks get None."""
    r, w = world()
    if w == 1:
        return local_rows
    counts = [len(range(q, n_items, w)) for q in range(w)]
    pad = max(counts)
    shape = (pad,) + tuple(local_rows.shape[1:])
    buf = torch.zeros(shape, dtype=local_rows.dtype, device=local_rows.device)
    buf[: local_rows.shape[0]] = local_rows
    out = [torch.empty_like(buf) for _ in range(w)]
    dist.all_gather(out, buf)
    if r != 0:
        return None
    full = torch.empty((n_items,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    for q in range(w):
        full[q::w] = out[q][: counts[q]]
    return full


def sample_covariance(rows: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Mean and unbiased sample covariance of realisation rows [n_real, n_features] (float64)."""
    rows = np.asarray(rows, dtype=np.float64)
    mean = rows.mean(axis=0)
    d = rows - mean
    cov = d.T @ d / max(rows.shape[0] - 1, 1)
    return mean, cov


def covariance_batch(seeds: Sequence[int], measure: Callable[[int], torch.Tensor]):
    """BASELINE.json configs[4]: P(k) of many independent realisations, sharded over ranks.

    ``measure(seed)`` returns that realisation's multipoles as a tensor [nbins, 3] (on any device);
    it is called only for this rank's seeds.  Returns on rank 0 ``(pk_rows [n, nbins, 3] numpy,
    mean [nbins*3], cov [nbins*3, nbins*3])`` and ``None`` elsewhere."""
    seeds = list(seeds)
    mine = shard_indices(len(seeds))
    rows = [measure(seeds[i]).detach().to(torch.float32).reshape(1, -1) for i in mine]
    ncol = None
    local = None
    if rows:
        local = torch.cat(rows, dim=0)
        ncol = local.shape[1]
    r, w = world()
    if w == 1 and local is None:                # no seeds at all: nothing to measure, nothing to gather
        return np.zeros((0, 0, 3), np.float32), np.zeros(0), np.zeros((0, 0))
    if w > 1:                                  # ranks with no work still need the row width
        dev = rows[0].device if rows else (torch.device("cuda", torch.cuda.current_device())
                                          if dist.get_backend() == "nccl" else torch.device("cpu"))
        t = torch.tensor([ncol or 0], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ncol = int(t.item())
        if not rows:
            local = torch.zeros((0, ncol), dtype=torch.float32, device=dev)
    full = gather_rows(local, len(seeds))
    if full is None:
        return None
    arr = full.cpu().numpy()
    mean, cov = sample_covariance(arr)
    return arr.reshape(len(seeds), -1, 3), mean, cov


def sharded_rows(n_items: int, compute: Callable[[list], "torch.Tensor | tuple"]):
    """Units that need no communication (triangle bins of the bispectrum sweep,
    realisations of a batch): rank r evaluates items r, r+W, ... with ONE call ``compute(indices)``, which
    returns a tensor -- or a tuple of tensors -- whose leading dimension is ``len(indices)``; rank 0 gets
    the same structure with ``n_items`` rows in global order, the other ranks ``None``.  The only exchange
    is the final gather.  ``compute`` is not called on a rank whose shard is empty."""
    r, w = world()
    mine = shard_indices(n_items)
    out = compute(mine) if mine else None
    single = isinstance(out, torch.Tensor)
    parts = None if out is None else ([out] if single else list(out))
    if w == 1:
        return out
    # ranks with an empty shard need the row shapes / dtypes / device kind of the others
    meta = None if parts is None else [(tuple(p.shape[1:]), str(p.dtype).replace("torch.", ""), single) for p in parts]
    metas = [None] * w
    dist.all_gather_object(metas, meta)
    ref = next(m for m in metas if m is not None)
    single = ref[0][2]
    if parts is None:
        dev = (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl"
               else torch.device("cpu"))
        parts = [torch.zeros((0,) + shp, dtype=getattr(torch, dt), device=dev) for shp, dt, _ in ref]
    full = [gather_rows(p.contiguous(), n_items) for p in parts]
    if r != 0:
        return None
    return full[0] if single else tuple(full)


def bispec_pairs_sharded(delta, box_size, k1, k2, theta, **kw):
    """BASELINE.json configs[2] on several GPUs: the (k1, k2) pairs of ``correlations.bispec_pairs`` are
    independent, so every rank holds the (small) mesh and evaluates its round-robin share of the pairs;
    rank 0 returns ``(k_all[np, bins+2], Pk[np, bins+2], theta, B[np, bins], Q[np, bins])`` in the order
    of ``k1``/``k2``, the other ranks ``None``.  No mesh sharding, no data-path collective."""
    from .correlations import bispec_pairs
    k1 = np.asarray(k1, dtype=np.float32).ravel()
    k2 = np.asarray(k2, dtype=np.float32).ravel()
    if k1.size != k2.size or k1.size < 1:
        raise ValueError("k1 and k2 must be equally long, non-empty 1-d arrays")
    th = np.asarray(theta, dtype=np.float32).ravel()

    on_nccl = dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"

    def compute(idx):
        k_all, pk, _, B, Q = bispec_pairs(delta, box_size, k1[idx], k2[idx], th, **kw)
        rows = tuple(torch.as_tensor(a) for a in (k_all, pk, B, Q))
        if on_nccl:                                   # NCCL gathers device tensors only (NumPy in -> host rows)
            dev = torch.device("cuda", torch.cuda.current_device())
            rows = tuple(t.to(dev) for t in rows)
        return rows

    res = sharded_rows(k1.size, compute)
    if res is None:
        return None
    k_all, pk, B, Q = res
    # same container kind as the input, like bispec_pairs: NumPy in -> NumPy out; host tensor in -> host tensors
    # (also under NCCL, which gathers on the device); CUDA tensor in -> CUDA tensors.  theta follows.
    if not isinstance(delta, torch.Tensor):
        k_all, pk, B, Q = (t.cpu().numpy() for t in (k_all, pk, B, Q))
        return k_all, pk, th, B, Q
    if not delta.is_cuda:
        k_all, pk, B, Q = (t.cpu() for t in (k_all, pk, B, Q))
    return k_all, pk, torch.from_numpy(th).to(k_all.device), B, Q
