"""Mesh painting with the reference's names and positional signatures.

Drop-in for /root/reference/src/mas.py (``cic_mas_vec`` :88-153, ``cic_mas`` :5-87) plus the
TSC / PCS painters the north star adds.  Arrays may be NumPy arrays, host torch tensors
(copied host->device once) or CUDA torch tensors (used in place); the mesh comes back in the
same kind of container as ``delta``.  Like the jitted reference the call is functional: the
input mesh is not modified unless ``inplace=True``.

All arithmetic runs in libjps.so (hand-written sm_100a kernels); there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib
from .plan import ArrayKind, ptr, require_cuda, stream_ptr, to_device_f32

__all__ = ["cic_mas_vec", "cic_mas", "tsc_mas_vec", "pcs_mas_vec", "paint", "paint_interlaced"]

_WS: dict = {}


def paint_workspace_bytes(n_mesh, n_part, order, method) -> int:
    nbytes = C.c_size_t(0)
    check(lib.jps_paint_workspace_bytes(int(n_mesh), int(n_part), int(order), int(method), C.byref(nbytes)),
          "jps_paint_workspace_bytes")
    return int(nbytes.value)


def new_paint_workspace(n_mesh, n_part, order, method, device):
    """A PRIVATE bucketing workspace (pipelines own theirs: two pipelines on two streams of one GPU
    must not share scratch)."""
    nbytes = paint_workspace_bytes(n_mesh, n_part, order, method)
    if nbytes == 0:
        return None, 0
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def _paint_workspace(n_mesh, n_part, order, method, device):
    """Scratch of the functional API: cached per (device, CUDA stream), grown on demand."""
    nbytes = paint_workspace_bytes(n_mesh, n_part, order, method)
    if nbytes == 0:
        return None, 0
    key = (device.index, int(torch.cuda.current_stream(device).cuda_stream))
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws, nbytes


def _common_stride(x, y, z):
    sx, sy, sz = (t.stride(0) if t.numel() > 1 else 1 for t in (x, y, z))
    if sx == sy == sz:
        return x, y, z, sx
    return x.contiguous(), y.contiguous(), z.contiguous(), 1


def paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap=True, *, order=2,
          compat="reference", variant="vec", method="auto", inplace=False):
    """mesh += deposit(particles); returns the mesh.  order 2/3/4 = CIC/TSC/PCS."""
    device = require_cuda()
    kind = ArrayKind(delta)
    n = int(n_bins)
    mesh = to_device_f32(delta, device)
    if tuple(mesh.shape) != (n, n, n):
        raise ValueError(f"delta has shape {tuple(mesh.shape)}, expected ({n},{n},{n}) for n_bins={n}")
    if not inplace and isinstance(delta, torch.Tensor) and mesh.data_ptr() == delta.data_ptr():
        mesh = mesh.clone()
    xd = to_device_f32(x, device, allow_strided=True)
    yd = to_device_f32(y, device, allow_strided=True)
    zd = to_device_f32(z, device, allow_strided=True)
    if not (xd.dim() == yd.dim() == zd.dim() == 1 and xd.numel() == yd.numel() == zd.numel()):
        raise ValueError("x, y, z must be 1-d arrays of equal length")
    xd, yd, zd, stride = _common_stride(xd, yd, zd)
    wd = None
    if w is not None:
        wd = to_device_f32(w, device)          # contiguous: the C ABI reads w with stride 1
        if wd.dim() != 1 or wd.numel() != xd.numel():
            raise ValueError("w must be a 1-d array as long as x")
    npart = xd.numel()
    meth = _lib.METHOD[method]
    ws, ws_bytes = _paint_workspace(n, npart, order, meth, device)
    check(lib.jps_paint(n, ptr(xd), ptr(yd), ptr(zd), ptr(wd), stride, npart,
                        float(xmin), float(ymin), float(zmin), float(box_size),
                        int(order), int(bool(wrap)), _lib.COMPAT[compat],
                        _lib.VARIANT_SCAN if variant == "scan" else _lib.VARIANT_VEC, meth,
                        ptr(mesh), ptr(ws), ws_bytes, stream_ptr()), "jps_paint")
    return kind.out(mesh)


def cic_mas_vec(delta, x, y, z, w, n_part, xmin, ymin, zmin, box_size, n_bins, wrap, **kw):
    """/root/reference/src/mas.py:89 ``cic_mas_vec``; ``n_part`` is unused there too (Q5)."""
    return paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap, order=2,
                 variant="vec", **kw)


def cic_mas(delta, x, y, z, w, n_part, xmin, ymin, zmin, box_size, n_bins, wrap, **kw):
    """/root/reference/src/mas.py:6 ``cic_mas`` (the lax.scan painter): same deposit, its own
    wrap=False handling (Q2).  Summation order differs (atomics), values agree to float32."""
    return paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap, order=2,
                 variant="scan", **kw)


def tsc_mas_vec(delta, x, y, z, w, n_part, xmin, ymin, zmin, box_size, n_bins, wrap, **kw):
    """Triangular-shaped-cloud painter (absent from the reference; same signature)."""
    kw.setdefault("compat", "fixed")
    return paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap, order=3, **kw)


def pcs_mas_vec(delta, x, y, z, w, n_part, xmin, ymin, zmin, box_size, n_bins, wrap, **kw):
    """Piecewise-cubic-spline painter (absent from the reference; same signature)."""
    kw.setdefault("compat", "fixed")
    return paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap, order=4, **kw)


def paint_interlaced(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, *, order=2, method="auto"):
    """The two meshes of an interlaced estimate (no reference counterpart): the particles painted on the
    grid at ``(xmin, ymin, zmin)`` and on the grid displaced by +half a cell, both periodic, textbook
    weights (``compat="fixed"``).  ``delta`` is the starting mesh of both (zeros, as a rule).  Feed the
    pair to ``powspec_vec(mesh1, ..., delta2=mesh2)``."""
    import numpy as np
    half = np.float32(0.5) * (np.float32(box_size) / np.float32(n_bins))          # float32, as the painters work
    m1 = paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, True, order=order, compat="fixed",
               method=method)
    m2 = paint(delta, x, y, z, w, float(np.float32(xmin) + half), float(np.float32(ymin) + half),
               float(np.float32(zmin) + half), box_size, n_bins, True, order=order, compat="fixed", method=method)
    return m1, m2
