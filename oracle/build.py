"""Build the C restatement (oracle/c/jps_oracle.c) into oracle/_build/libjps_oracle.so with gcc.
TEST INFRASTRUCTURE, NOT PRODUCT.  The reference itself is pure Python: there is nothing to
compile into oracle/_ref/ (see DESIGN.md)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "jps_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libjps_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # -ffp-contract=off: keep float32 products and sums separately rounded, as XLA CPU does
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force=True))
