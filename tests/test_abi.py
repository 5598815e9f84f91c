"""The C-ABI library loads without a GPU and exports every symbol include/jps.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jax_powspec_b200", "libjps.so")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "jps.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"JPS_API\s+[\w\s\*]+?\b(jps_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import runpy
        runpy.run_path(os.path.join(ROOT, "jax_powspec_b200", "build.py"))["build"]()
    return ctypes.CDLL(LIB)


def test_header_declares_something():
    syms = _declared_symbols()
    assert "jps_paint" in syms and "jps_powspec" in syms and len(syms) >= 10


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/jps.h but not exported: {missing}"


def test_python_binding_covers_header():
    from jax_powspec_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_version_and_error_string(lib):
    lib.jps_version.restype = ctypes.c_int
    assert lib.jps_version() == 100
    lib.jps_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.jps_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # no compute: argument checks fire before any CUDA call
    lib.jps_paint_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_size_t)]
    n = ctypes.c_size_t(0)
    assert lib.jps_paint_workspace_bytes(64, 1000, 7, 1, ctypes.byref(n)) == -1
    lib.jps_last_error.restype = ctypes.c_char_p
    assert b"order" in lib.jps_last_error()
    assert lib.jps_paint_workspace_bytes(64, 1000, 2, 1, ctypes.byref(n)) == 0 and n.value == 0
    lib.jps_fundamental_nbins.argtypes = [ctypes.c_int]
    assert lib.jps_fundamental_nbins(256) == 221      # floor(sqrt(3)*128), SURVEY a-6


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jax_powspec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import jax_powspec_b200 as jps
    with pytest.raises(jps._lib.JpsError):
        jps.powspec_vec(np.zeros((8, 8, 8), np.float32), 100.0, np.arange(0.1, 1, 0.1))
