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
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: not the package, not
    the developer tools, not the FFI shim, not the C ABI header."""
    for top in ("jax_powspec_b200", "tools", "ffi", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".sh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import jax_powspec_b200 as jps
    with pytest.raises(jps._lib.JpsError):
        jps.powspec_vec(np.zeros((8, 8, 8), np.float32), 100.0, np.arange(0.1, 1, 0.1))


def _declared_prototypes():
    """name -> (return type, [parameter types]) parsed from include/jps.h (comments stripped)."""
    text = open(os.path.join(ROOT, "include", "jps.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"JPS_API\s+([\w\s\*]+?)\b(jps_\w+)\s*\((.*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        types = []
        for p in plist:
            p = re.sub(r"\s+", " ", p)
            if "*" in p:
                types.append("ptr")
            else:
                types.append(re.sub(r"\s*\w+$", "", p).replace("const ", "").strip())   # drop the parameter name
        protos[name] = (ret, types)
    return protos


def test_ctypes_signatures_match_the_header():
    """Every binding in _lib.SIGNATURES has the header's arity, scalar kinds and pointer positions:
    a wrong argtypes list corrupts the call silently."""
    import ctypes as C
    from jax_powspec_b200 import _lib
    scalar = {"int": C.c_int, "float": C.c_float, "int64_t": C.c_int64, "size_t": C.c_size_t,
              "unsigned long long": C.c_ulonglong, "double": C.c_double}
    protos = _declared_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, (ret, types) in protos.items():
        res, args = _lib.SIGNATURES[name]
        assert len(args) == len(types), f"{name}: header has {len(types)} parameters, binding {len(args)}"
        for i, (t, a) in enumerate(zip(types, args)):
            if t == "ptr":
                is_ptr = a in (C.c_void_p, C.c_char_p) or hasattr(a, "contents") or issubclass(a, C._Pointer)
                assert is_ptr, f"{name} arg {i}: header pointer, binding {a}"
            else:
                assert t in scalar, f"{name} arg {i}: unknown scalar type '{t}' in the header"
                assert a is scalar[t], f"{name} arg {i}: header {t}, binding {a}"
        if "*" in ret:
            assert res in (C.c_char_p, C.c_void_p), name
        else:
            assert res is scalar[ret.replace("const ", "").strip()], f"{name}: return {ret} vs {res}"


def test_header_is_plain_c_and_links_from_c(tmp_path, lib):
    """include/jps.h is the drop-in boundary: it must be consumable by a C compiler (no C++-isms, no torch
    types) and libjps.so must link from plain C.  The program only calls entry points that need no GPU."""
    import subprocess
    src = tmp_path / "caller.c"
    src.write_text(
        '#include <stdio.h>\n#include "jps.h"\n'
        "int main(void) {\n"
        "  size_t bytes = 123;\n"
        "  if (jps_version() != JPS_VERSION) return 1;\n"
        "  if (jps_paint_workspace_bytes(64, 1000, 2, JPS_PAINT_ATOMIC, &bytes) != JPS_OK || bytes != 0) return 2;\n"
        "  if (jps_paint_workspace_bytes(64, 1000, 9, JPS_PAINT_ATOMIC, &bytes) != JPS_ERR_INVALID) return 3;\n"
        "  if (jps_last_error()[0] == 0) return 4;\n"
        "  if (jps_fundamental_nbins(256) != 221) return 5;\n"
        "  if (jps_mock_field_workspace_bytes(100) < 1600) return 6;\n"
        '  puts("ok");\n  return 0;\n}\n')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(LIB)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-L", libdir, "-ljps", f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout)
