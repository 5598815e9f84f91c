"""ctypes binding of libjps.so (the C ABI declared in include/jps.h).

There is NO fallback: if the CUDA library is missing or fails to load, importing this
module raises.  Build it with ``python jax_powspec_b200/build.py``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjps.so")

# constants of include/jps.h
ORDER_CIC, ORDER_TSC, ORDER_PCS = 2, 3, 4
COMPAT_REFERENCE, COMPAT_FIXED = 0, 1
VARIANT_VEC, VARIANT_SCAN = 0, 1
PAINT_AUTO, PAINT_ATOMIC, PAINT_SORTED = 0, 1, 2
PK_HERMITIAN = 1
PLAN_TABLES_ONLY, PLAN_FFT_PENCIL = 1, 2
PAINT_PHASE_ALL, PAINT_PHASE_BUCKET, PAINT_PHASE_DEPOSIT = 0, 1, 2

COMPAT = {"reference": COMPAT_REFERENCE, "fixed": COMPAT_FIXED}
METHOD = {"auto": PAINT_AUTO, "atomic": PAINT_ATOMIC, "sorted": PAINT_SORTED}


class JpsError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA library is the product and there is no CPU fallback. "
        "Build it with `python jax_powspec_b200/build.py` (needs nvcc, targets sm_100a)."
    )

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_fp = C.POINTER(C.c_float)

# every symbol include/jps.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "jps_version": (_i, []),
    "jps_last_error": (C.c_char_p, []),
    "jps_profile_enable": (_i, [_i]),
    "jps_profile_reset": (_i, []),
    "jps_profile_num_kernels": (_i, []),
    "jps_profile_get": (_i, [_i, C.POINTER(C.c_char_p), C.POINTER(C.c_ulonglong), C.POINTER(C.c_double)]),
    "jps_plan_workspace_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "jps_plan_create": (_i, [_i, _i, _i, _vp, _sz, C.POINTER(_vp)]),
    "jps_plan_destroy": (_i, [_vp]),
    "jps_paint_workspace_bytes": (_i, [_i, _i64, _i, _i, C.POINTER(_sz)]),
    "jps_paint": (_i, [_i, _vp, _vp, _vp, _vp, _i64, _i64, _f, _f, _f, _f, _i, _i, _i, _i, _i,
                       _vp, _vp, _sz, _vp]),
    "jps_paint_slab": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i64, _i64, _f, _f, _f, _f, _i, _i, _i, _i, _i,
                            _vp, _vp, _sz, _vp]),
    "jps_paint_tile_rows": (_i, [_i]),
    "jps_paint_slab_phase": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i64, _i64, _f, _f, _f, _f, _i, _i, _i, _i, _i,
                                  _vp, _vp, _sz, _i, _i, _i, _vp]),
    "jps_powspec": (_i, [_vp, _vp, _i, _f, _fp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jps_powspec_ex": (_i, [_vp, _vp, _vp, _i, _f, _fp, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jps_fundamental_nbins": (_i, [_i]),
    "jps_powspec_fundamental": (_i, [_vp, _vp, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jps_xi": (_i, [_vp, _vp, _i, _f, _fp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jps_xi_fundamental": (_i, [_vp, _vp, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jps_bispec": (_i, [_vp, _vp, _i, _f, _f, _f, _fp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "jps_bispec_pairs": (_i, [_vp, _vp, _i, _f, _fp, _fp, _i, _fp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "jps_compute_2pt_correlations": (_i, [_vp, _vp, _i, _f, _fp, _i, _fp, _i, _i,
                                          _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jps_compute_all_correlations": (_i, [_vp, _vp, _i, _f, _fp, _i, _fp, _i, _f, _f, _fp, _i, _i,
                                          _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jps_slab_plan_workspace_bytes": (_i, [_i, _i, C.POINTER(_sz)]),
    "jps_slab_plan_create": (_i, [_i, _i, _i, _vp, _sz, C.POINTER(_vp)]),
    "jps_slab_plan_destroy": (_i, [_vp]),
    "jps_slab_fft_yz": (_i, [_vp, _vp, _vp, _vp]),
    "jps_slab_pack": (_i, [_vp, _vp, _vp, _vp]),
    "jps_ipc_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "jps_ipc_close": (_i, [_vp]),
    "jps_enable_peer_access": (_i, [_i]),
    "jps_slab_pack_p2p": (_i, [_vp, _vp, C.POINTER(_vp), _vp]),
    "jps_slab_chunk_planes": (_i, [_vp]),
    "jps_slab_set_layout": (_i, [_vp, _i]),
    "jps_slab_fft_yz_planes": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "jps_slab_pack_p2p_planes": (_i, [_vp, _vp, C.POINTER(_vp), _i, _i, _vp]),
    "jps_slab_fft_x": (_i, [_vp, _vp, _vp]),
    "jps_slab_powspec_partial": (_i, [_vp, _vp, _vp, _i, _f, _fp, _i, _i, _vp, _vp, _vp]),
    "jps_slab_powspec_finalize": (_i, [_vp, _f, _fp, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp]),
    "jps_powspec_grad": (_i, [_vp, _vp, _i, _f, _fp, _i, _i, _vp, _vp, _vp]),
    "jps_paint_grad": (_i, [_i, _vp, _vp, _vp, _vp, _i64, _i64, _f, _f, _f, _f, _i, _i, _i, _i, _vp,
                            _vp, _vp, _vp, _vp, _vp]),
    "jps_paint_powspec": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _f, _f, _f, _f, _i, _i, _i, _i,
                               _fp, _i, _f, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jps_xi_grad": (_i, [_vp, _vp, _f, _fp, _i, _i, _vp, _vp, _vp]),
    "jps_bispec_grad": (_i, [_vp, _vp, _f, _f, _f, _fp, _i, _i, _vp, _vp, _vp, _vp]),
    "jps_text_workspace_bytes": (_sz, [_i64, _i64, _i]),
    "jps_text_count_lines": (_i, [_vp, _i64, _vp, _vp, _sz, _vp]),
    "jps_text_parse": (_i, [_vp, _i64, _i64, _i, _i, C.POINTER(_i), _i, _i, _f, _f, _vp, _vp, _vp, _i64,
                            _vp, _sz, _vp]),
    "jps_mock_field_workspace_bytes": (_sz, [_i]),
    "jps_mock_gaussian_field": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), _i, _i, C.c_ulonglong, _f,
                                     _vp, _vp, _sz, _vp]),
    "jps_mock_populate_workspace_bytes": (_sz, [_i]),
    "jps_mock_populate_counts_offset": (_sz, [_i]),
    "jps_mock_populate_count": (_i, [_vp, _i, _f, _f, _i, _f, C.c_ulonglong, _vp, _sz, _vp, _vp]),
    "jps_mock_populate_fill": (_i, [_i, _f, C.c_ulonglong, _vp, _sz, _i64, _vp, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = library / header mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.jps_last_error()
        raise JpsError(f"{what or 'libjps'} failed ({rc}): {msg.decode() if msg else ''}")


# kernels that are library calls (cuFFT, cudaMemset), not hand-written ones
LIBRARY_KERNELS = ("cufft_r2c", "cufft_c2r", "cufft_c2c_y", "cufft_c2c_x", "memset")


def profile_enable(on: bool) -> None:
    check(lib.jps_profile_enable(int(bool(on))))


def profile_reset() -> None:
    check(lib.jps_profile_reset())


def profile_snapshot() -> dict:
    """{kernel name: (launches, device ms)}; synchronises on the recorded events."""
    out = {}
    for i in range(lib.jps_profile_num_kernels()):
        name, n, ms = C.c_char_p(), C.c_ulonglong(0), C.c_double(0.0)
        check(lib.jps_profile_get(i, C.byref(name), C.byref(n), C.byref(ms)))
        if n.value:
            out[name.value.decode()] = (int(n.value), float(ms.value))
    return out
