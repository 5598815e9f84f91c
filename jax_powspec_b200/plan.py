"""Plan cache and tensor plumbing (torch is used for device memory and streams only)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.JpsError(
            "jax_powspec_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class Plan:
    """Owns the workspace tensor and the jps_plan_t* built on it."""

    def __init__(self, n_mesh: int, n_shell_fields: int, device: torch.device, flags: int = 0):
        self.n = int(n_mesh)
        self.n_shell_fields = int(n_shell_fields)
        self.device = device
        self.flags = int(flags)
        nbytes = C.c_size_t(0)
        with torch.cuda.device(device):
            check(lib.jps_plan_workspace_bytes(self.n, self.n_shell_fields, self.flags, C.byref(nbytes)),
                  "jps_plan_workspace_bytes")
            self.workspace = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=device)
            base = self.workspace.data_ptr()
            aligned = (base + 255) // 256 * 256
            handle = C.c_void_p(0)
            check(lib.jps_plan_create(self.n, self.n_shell_fields, self.flags, C.c_void_p(aligned),
                                      C.c_size_t(nbytes.value), C.byref(handle)), "jps_plan_create")
        self.handle = handle
        self.workspace_bytes = nbytes.value

    def close(self):
        if getattr(self, "handle", None):
            lib.jps_plan_destroy(self.handle)
            self.handle = None
            self.workspace = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_PLANS: dict = {}


def get_plan(n_mesh: int, device: torch.device, n_shell_fields: int = 0) -> Plan:
    """Plan of the functional API, one per (mesh size, device, CUDA stream): a plan's delta_k buffer and
    accumulators belong to the stream its calls are enqueued on (include/jps.h: "one plan per host
    thread/stream"), so two streams of one GPU never share one.  A plan that is too small (fewer shell
    fields than this call needs) is REPLACED in the cache but never destroyed here: objects that still
    hold it keep a valid plan, and it is freed when the last of them drops it."""
    key = (int(n_mesh), device.index, int(torch.cuda.current_stream(device).cuda_stream))
    p = _PLANS.get(key)
    if p is None or p.handle is None or p.n_shell_fields < n_shell_fields:
        p = Plan(n_mesh, n_shell_fields, device)
        _PLANS[key] = p
    return p


def clear_plans():
    for p in list(_PLANS.values()):
        p.close()
    _PLANS.clear()


# --------------------------------------------------------------------------- array plumbing
class ArrayKind:
    """Remembers what the caller passed so results come back the same way."""

    def __init__(self, like):
        self.is_torch = isinstance(like, torch.Tensor)
        self.on_device = self.is_torch and like.is_cuda

    def out(self, t: torch.Tensor):
        if self.is_torch:
            return t if self.on_device else t.cpu()
        return t.cpu().numpy()


def check_particles(x, y, z, w, device):
    """Fast-path argument check (the C ABI takes raw pointers and reads ``w`` with stride 1): x, y, z
    [, w] must be 1-d float32 CUDA tensors of equal length on ``device``; x, y, z may be equally strided
    column views, ``w`` is made contiguous.  Returns (x, y, z, w, stride).  Raises instead of
    misreading memory."""
    n = None
    for name, t in (("x", x), ("y", y), ("z", z), ("w", w)):
        if t is None and name == "w":
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch tensor (got {type(t).__name__}); use paint()/paint_powspec() for NumPy input")
        if t.dtype != torch.float32:
            raise TypeError(f"{name} must be float32 (got {t.dtype})")
        if not t.is_cuda or (device.index is not None and t.device.index != device.index):
            raise ValueError(f"{name} lives on {t.device}, the pipeline on {device}")
        if t.dim() != 1:
            raise ValueError(f"{name} must be 1-d (got shape {tuple(t.shape)})")
        if n is None:
            n = t.numel()
        elif t.numel() != n:
            raise ValueError(f"{name} has {t.numel()} elements, x has {n}")
    strides = {t.stride(0) if t.numel() > 1 else 1 for t in (x, y, z)}
    if len(strides) > 1 or min(strides) < 1:
        x, y, z = x.contiguous(), y.contiguous(), z.contiguous()
        stride = 1
    else:
        stride = strides.pop()
    if w is not None and w.numel() > 1 and w.stride(0) != 1:
        w = w.contiguous()
    return x, y, z, w, stride


def to_device_f32(a, device, *, allow_strided=False):
    """float32 CUDA tensor for `a` (numpy array, host or device torch tensor).  Host data goes
    through one host->device copy; device data is used in place (no copy) when it is float32
    and contiguous (or a 1-d strided view when allow_strided)."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if not t.is_cuda:
        t = t.to(device, non_blocking=True)
    if not t.is_contiguous() and not (allow_strided and t.dim() == 1 and t.stride(0) >= 1):
        t = t.contiguous()
    return t
