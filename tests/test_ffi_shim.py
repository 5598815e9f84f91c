"""The jax.ffi shim (ffi/jax_ffi_shim.cc) cannot be linked here (no jaxlib), but it must at least parse
and be ARITY-CORRECT: it is compiled against ffi/stub/xla/ffi/api/ffi.h, a stub of the XLA FFI binding
API whose `.To(fn)` refuses a handler whose Bind() chain and implementation disagree (the round-1
`JpsPaint` bound 5 operands while INTEGRATION.md passed 6).  CPU only."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "ffi", "jax_ffi_shim.cc")
CUDA_INC = "/usr/local/cuda/include"


def _compile(src, extra=()):
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-DJPS_WITH_JAX_FFI", "-I", os.path.join(ROOT, "ffi", "stub"),
           "-I", os.path.join(ROOT, "include"), "-I", CUDA_INC, *extra, src]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(CUDA_INC), reason="needs g++ and the CUDA headers")
def test_shim_compiles_against_the_stub_and_binds_every_entry_point():
    r = _compile(SHIM)
    assert r.returncode == 0, r.stdout
    src = open(SHIM).read()
    handlers = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", src))
    assert handlers == {"JpsPaint", "JpsPowspec", "JpsPowspecFundamental", "JpsBispec", "JpsPaintPowspec",
                        "JpsPaintGrad", "JpsPowspecGrad"}
    # the registration snippet of INTEGRATION.md passes exactly the operands JpsPaint binds
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    call = doc[doc.index('ffi_call("jps_paint"'):]
    operands = re.search(r"\)\(\s*(?:#[^\n]*\n\s*)?([^=]*?),\s*xmin=", call, re.S).group(1)
    n_operands = len([t for t in operands.replace("\n", " ").split(",") if t.strip()])
    bind = src[src.index("XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsPaint"):]
    bind = bind[:bind.index(".Ret<")]
    assert n_operands == bind.count(".Arg<") == 6


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(CUDA_INC), reason="needs g++ and the CUDA headers")
def test_stub_rejects_an_arity_mismatch(tmp_path):
    """The check has teeth: dropping one .Arg<> from a Bind() chain must fail to compile."""
    src = open(SHIM).read()
    broken = src.replace(".Arg<ffi::Buffer<ffi::F32>>()   // delta (operand 0, aliased to the result)\n", "", 1)
    assert broken != src
    f = tmp_path / "broken_shim.cc"
    f.write_text(broken)
    r = _compile(str(f))
    assert r.returncode != 0 and "disagree in arity" in r.stdout
