"""ctypes wrapper of the C restatement (oracle/c/jps_oracle.c) -- TEST INFRASTRUCTURE.
Used by tests and by bench.py's CPU legs (cpu_baseline / --impl reference)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.fft as sfft

from . import build as _build
from .correlations import grid_edges, k_fundamental

F32 = np.float32
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB if os.path.exists(_build.LIB) else _build.build()
        _lib = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        _lib.jpso_paint_cic_reference.argtypes = [fp, fp, fp, fp, fp, C.c_int64, C.c_float, C.c_float, C.c_float,
                                                  C.c_float, C.c_int, C.c_int, C.c_int]
        _lib.jpso_paint_bspline.argtypes = [fp, fp, fp, fp, fp, C.c_int64, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_int, C.c_int, C.c_int]
        _lib.jpso_pk_bin.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, fp, fp, fp, fp]
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        _lib.jpso_paint_f64.argtypes = [dp, fp, fp, fp, fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.jpso_pk_bin_f64.argtypes = [dp, C.c_int, fp, C.c_int, C.c_int, dp, dp, dp, ip]
        _lib.jpso_paint_bspline_mt.argtypes = _lib.jpso_paint_bspline.argtypes
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def paint(mesh, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap=True, *, order=2,
          compat="reference", variant="vec"):
    """Serial float32 painter; accumulates into a copy of ``mesh``."""
    n = int(n_bins)
    out = np.ascontiguousarray(mesh, dtype=F32).copy()
    x, y, z = (np.ascontiguousarray(a, dtype=F32) for a in (x, y, z))
    w = None if w is None else np.ascontiguousarray(w, dtype=F32)
    if order == 2 and compat == "reference":
        rc = lib().jpso_paint_cic_reference(_p(out), _p(x), _p(y), _p(z), _p(w), len(x), xmin, ymin, zmin,
                                            box_size, n, int(bool(wrap)), 1 if variant == "scan" else 0)
    else:
        rc = lib().jpso_paint_bspline(_p(out), _p(x), _p(y), _p(z), _p(w), len(x), xmin, ymin, zmin,
                                      box_size, n, int(bool(wrap)), int(order))
    if rc != 0:
        raise RuntimeError("C oracle paint failed")
    return out


def powspec(delta, box_size, k_edges, *, mas_order=2, workers=-1, times=None):
    """rfftn (scipy, all cores) + serial float32 binning.  Returns (k3D, Pk3D f32[nb,3], Nmodes f32).
    ``times`` (a dict) receives the wall time of the two parts: fft_s, bin_s."""
    import time
    delta = np.ascontiguousarray(delta, dtype=F32)
    n = delta.shape[0]
    t0 = time.perf_counter()
    dk = sfft.rfftn(delta, workers=workers).astype(np.complex64, copy=False)
    dk = np.ascontiguousarray(dk)
    t1 = time.perf_counter()
    kedges = grid_edges(k_edges, box_size)
    nb = len(kedges) - 1
    s0, s2, s4, cnt = (np.zeros(nb, F32) for _ in range(4))
    rc = lib().jpso_pk_bin(_p(dk.view(F32)), n, _p(kedges), nb, int(mas_order), _p(s0), _p(s2), _p(s4), _p(cnt))
    if rc != 0:
        raise RuntimeError("C oracle binning failed")
    if times is not None:
        times["fft_s"], times["bin_s"] = t1 - t0, time.perf_counter() - t1
    vol = (F32(box_size) / F32(n * n)) ** 3
    with np.errstate(invalid="ignore", divide="ignore"):
        pk = np.stack([s0 / cnt * vol, s2 / cnt * F32(5.0) * vol, s4 / cnt * F32(9.0) * vol], axis=1)
    k3d = (F32(0.5) * (kedges[1:] + kedges[:-1]) * k_fundamental(box_size)).astype(F32)
    return k3d, pk, cnt


def num_threads():
    return int(lib().jpso_num_threads())


def paint_f64(mesh, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap=True, *, order=2,
              compat="reference", variant="vec"):
    """oracle/mas.py ``paint(precision="f64")`` in C + OpenMP (float32 cell decisions, float64 weights and
    sums) for catalogues of 1e7..1e8 particles; pinned on the NumPy oracle in tests/test_cport_cpu.py.
    Returns a float64 mesh."""
    n = int(n_bins)
    out = np.ascontiguousarray(mesh, dtype=np.float64).copy()
    x, y, z = (np.ascontiguousarray(a, dtype=F32) for a in (x, y, z))
    w = None if w is None else np.ascontiguousarray(w, dtype=F32)
    ref = 1 if (order == 2 and compat == "reference") else 0
    rc = lib().jpso_paint_f64(out.ctypes.data_as(C.POINTER(C.c_double)), _p(x), _p(y), _p(z), _p(w), len(x),
                              xmin, ymin, zmin, box_size, n, int(bool(wrap)), int(order), ref,
                              1 if variant == "scan" else 0)
    if rc != 0:
        raise RuntimeError("C oracle paint_f64 failed")
    return out


def paint_mt(mesh, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap=True, *, order=2):
    """Best-effort multi-core float32 painter (textbook weights); bench.py's cpu_multicore leg."""
    n = int(n_bins)
    out = np.ascontiguousarray(mesh, dtype=F32).copy()
    x, y, z = (np.ascontiguousarray(a, dtype=F32) for a in (x, y, z))
    w = None if w is None else np.ascontiguousarray(w, dtype=F32)
    rc = lib().jpso_paint_bspline_mt(_p(out), _p(x), _p(y), _p(z), _p(w), len(x), xmin, ymin, zmin,
                                     box_size, n, int(bool(wrap)), int(order))
    if rc != 0:
        raise RuntimeError("C oracle paint_mt failed")
    return out


def powspec_f64(delta, box_size, k_edges, *, mas_order=2, workers=-1):
    """oracle/correlations.py ``powspec(precision="f64")`` for big meshes: float64 rfftn (scipy) of
    ``delta`` as given, float32 bin decisions, float64 window / Legendre weights / sums in C + OpenMP.
    Returns (k3D f32, Pk3D float64[nb,3], counts int64)."""
    d = np.ascontiguousarray(delta, dtype=np.float64)
    n = d.shape[0]
    dk = np.ascontiguousarray(sfft.rfftn(d, workers=workers))
    del d
    kedges = grid_edges(k_edges, box_size)
    nb = len(kedges) - 1
    s0, s2, s4 = (np.zeros(nb, np.float64) for _ in range(3))
    cnt = np.zeros(nb, np.int64)
    dp = C.POINTER(C.c_double)
    rc = lib().jpso_pk_bin_f64(dk.view(np.float64).ctypes.data_as(dp), n, _p(kedges), nb, int(mas_order),
                               s0.ctypes.data_as(dp), s2.ctypes.data_as(dp), s4.ctypes.data_as(dp),
                               cnt.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc != 0:
        raise RuntimeError("C oracle pk_bin_f64 failed")
    vol = (np.float64(box_size) / np.float64(n * n)) ** 3
    with np.errstate(invalid="ignore", divide="ignore"):
        pk = np.stack([s0 / cnt * vol, s2 / cnt * 5.0 * vol, s4 / cnt * 9.0 * vol], axis=1)
    k3d = (F32(0.5) * (kedges[1:] + kedges[:-1]) * k_fundamental(box_size)).astype(F32)
    return k3d, pk, cnt
