"""Golden vectors for the mock generator (row f-3) from the UNMODIFIED reference
/root/reference/src/gauss_field.py.  TEST INFRASTRUCTURE, NOT PRODUCT.

gaussian_field() is plain NumPy, so it runs here as is.  Its only non-portable ingredient is the
random stream: it draws np.random.random() twice per mode in loop order (gauss_field.py:49-51).  The
device generator gives every mode its own Philox counter instead, so to compare the ARITHMETIC
(interpolation, amplitude, Hermitian fill order) this script feeds the reference exactly the uniforms
our generator assigns to each mode -- np.random.random is replaced by an iterator over them, nothing
else is touched -- and stores what the reference then returns.

populate_field() is run the same way on a small mesh, in both of its versions: the NumPy twin
(gauss_field.py:90-110) directly, with np.random.poisson returning our per-cell counts (in the
reference's descending-density cell order) and np.random.random our per-particle uniforms; and the JAX
one (src/populate_field.py:11-29) through oracle/jaxshim.py with the same draws injected for jax.random.
Stored are their coordinates re-ordered to C cell order.  The NumPy twin mixes float32 cell centres with
float64 offsets (hence a 2e-7 relative tolerance in the tests); the JAX twin is float32 throughout, which
is the arithmetic the device follows -- its rows are reproduced bit for bit.

    python -m oracle.make_mock_golden        (needs /root/reference; writes tests/golden/ref_mock.npz)
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/src/gauss_field.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_gauss_field", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    sys.path.insert(0, ROOT)
    from tests.test_mockgen_cpu import load_host
    host = load_host()
    ref = load_reference()
    out = {}
    real_random, real_seed, real_poisson = np.random.random, np.random.seed, np.random.poisson
    cases = [(8, 0, 11, 300.0), (12, 1, 12, 500.0), (9, 1, 13, 250.0)]
    try:
        for i, (n, rayleigh, seed, box) in enumerate(cases):
            kf = np.linspace(1e-3, 2.5, 64)
            pkf = 3.0e3 * (kf / 0.05) / (1.0 + (kf / 0.05) ** 2) ** 1.6
            u = np.zeros((n, n, n // 2 + 1, 2), dtype=np.float64)
            host.mock_field_uniforms(n, seed, u.ctypes.data)
            stream = iter(u.ravel())                      # C order, (phase, amplitude) per mode
            np.random.seed = lambda s: None
            np.random.random = lambda: float(next(stream))
            dk = ref.gaussian_field(n, kf, pkf, rayleigh, seed, box)
            assert next(stream, None) is None             # the reference consumed exactly two draws per mode
            out.update({f"gf{i}_n": n, f"gf{i}_rayleigh": rayleigh, f"gf{i}_seed": seed, f"gf{i}_box": box,
                        f"gf{i}_kf": kf, f"gf{i}_pkf": pkf, f"gf{i}_delta_k": dk})
            print(f"gaussian_field case {i}: n={n} rayleigh={rayleigh} |dk|max={np.abs(dk).max():.4g}")

        # ---- populate_field on a small mesh
        n, box, density, seed = 6, 120.0, 0.004, 31
        rng = np.random.default_rng(4)
        rho0 = np.exp(rng.normal(size=(n, n, n))).astype(np.float32)
        s = host.mock_density_sum(rho0.ctypes.data, rho0.size, 0, 0.0, 148 * 8, 256)
        counts = np.zeros(rho0.size, dtype=np.uint32)
        total = host.mock_populate_count(rho0.ctypes.data, n, box, density, 0, 0.0, seed, s, counts.ctypes.data)
        pos = np.zeros((total, 3), dtype=np.float32)
        host.mock_populate_fill(counts.ctypes.data, n, np.float32(box), seed, pos.ctypes.data)
        # our per-particle uniforms, in OUR particle order (cells in C order)
        ctr = np.zeros(4, np.uint32)
        key = np.array([seed & 0xffffffff, seed >> 32], np.uint32)
        o = np.zeros(4, np.uint32)
        uni = np.zeros((total, 3), dtype=np.float64)
        for p in range(total):
            ctr[:] = (p & 0xffffffff, p >> 32, 0, 0x4f464673)          # STREAM_OFFSET
            host.mock_philox(ctr.ctypes.data, key.ctypes.data, o.ctypes.data)
            uni[p] = [(int(v) >> 8) * 2.0 ** -24 for v in o[:3]]
        first = np.concatenate([[0], np.cumsum(counts)[:-1]])
        # the reference visits cells by descending density: hand it counts and uniforms in that order
        rho_scaled = rho0.copy()
        rho_scaled *= (box / n) ** 3 * density / rho_scaled.mean()       # what populate_field does in place (:95)
        order = np.argsort(rho_scaled.ravel())[::-1]
        rows = np.concatenate([np.arange(first[c], first[c] + counts[c]) for c in order]).astype(np.int64)
        np.random.poisson = lambda lam, size=None: counts[order].astype(np.int64)
        np.random.random = lambda size=None: uni[rows]
        coords = ref.populate_field(rho0.copy(), n, box, density, seed)
        back = np.empty_like(coords)
        back[rows] = coords                                               # to C cell order
        out.update({"pf_n": n, "pf_box": box, "pf_density": density, "pf_seed": seed, "pf_rho": rho0,
                    "pf_counts": counts, "pf_coords_ref": back, "pf_coords_host": pos})
        print(f"populate_field: {total} particles, max |ref - host| = {np.abs(back - pos).max():.3g}")

        # ---- the JAX twin, src/populate_field.py:11-29, UNMODIFIED, on NumPy through oracle/jaxshim.py with
        # the same counts and uniforms injected for jax.random (float32 arithmetic throughout, as under JAX)
        np.random.random, np.random.seed, np.random.poisson = real_random, real_seed, real_poisson
        from oracle import jaxshim
        refj = jaxshim.load_reference_module("populate_field")
        jaxshim.RANDOM_FEED["poisson"].append(counts[order])
        jaxshim.RANDOM_FEED["uniform"].append(uni[rows].astype(np.float32))
        key = sys.modules["jax"].random.PRNGKey(seed)
        coords_j = np.asarray(refj.populate_field(jaxshim._wrap(rho0.copy()), n, box, density, key))
        assert coords_j.dtype == np.float32
        back_j = np.empty_like(coords_j)
        back_j[rows] = coords_j
        out["pf_coords_ref_jax"] = back_j
        print(f"populate_field (JAX twin): bit-identical rows {np.mean(np.all(back_j == pos, axis=1)):.6f}, "
              f"max |ref - host| = {np.abs(back_j - pos).max():.3g}")
    finally:
        np.random.random, np.random.seed, np.random.poisson = real_random, real_seed, real_poisson
    np.savez_compressed(os.path.join(GOLD, "ref_mock.npz"), **out)
    print("wrote", os.path.join(GOLD, "ref_mock.npz"))


if __name__ == "__main__":
    main()
