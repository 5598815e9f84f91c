"""Generate the golden vectors under tests/golden/ by executing the UNMODIFIED reference
sources (/root/reference/src/mas.py, /root/reference/src/correlations.py) on NumPy through
oracle/jaxshim.py.  TEST INFRASTRUCTURE, NOT PRODUCT.

Run once in the build container (where /root/reference exists):
    python -m oracle.run_reference
The .npz files it writes are committed; tests never need /root/reference at run time.
Every file stores the inputs next to the outputs so nothing depends on RNG stability.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
F32 = np.float32


def clustered_particles(rng, n_part, box, n_blobs=12):
    """A seeded clustered catalogue: half uniform, half in Gaussian blobs; z-elongated (RSD-like)."""
    nu = n_part // 2
    uni = rng.random((nu, 3)) * box
    centres = rng.random((n_blobs, 3)) * box
    which = rng.integers(0, n_blobs, n_part - nu)
    sig = np.array([0.03, 0.03, 0.06]) * box
    blob = centres[which] + rng.normal(size=(n_part - nu, 3)) * sig
    p = np.concatenate([uni, blob]) % box
    p = p.astype(F32)
    p[p >= F32(box)] = 0.0   # the modulo can round up to box in float32
    return p


def helpers(corr):
    """xi_vec_coords / s_edges_conv (src/correlations.py:262-272): 1-d host helpers of the xi path."""
    out = {}
    cases = [(256, 2500.0, np.arange(1e-3, 200.0, 2.0)), (64, 1000.0, np.arange(0.0, 150.0, 7.5)),
             (15, 250.0, np.linspace(0.5, 90.0, 23))]
    for i, (dims, box, se) in enumerate(cases):
        se = se.astype(F32)
        ke = np.asarray(corr.s_edges_conv(dims, box, se))
        out[f"h{i}_dims"], out[f"h{i}_box"], out[f"h{i}_s_edges"] = dims, box, se
        out[f"h{i}_s_edges_conv"] = ke
        out[f"h{i}_xi_vec_coords"] = np.asarray(corr.xi_vec_coords(dims, box, ke))
    np.savez_compressed(os.path.join(GOLD, "ref_helpers.npz"), **out)
    print("wrote helpers")


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import jaxshim

    mas, corr = jaxshim.load_reference()
    os.makedirs(GOLD, exist_ok=True)
    warnings.simplefilter("ignore")
    helpers(corr)
    if "--helpers-only" in sys.argv:
        return

    # ---------------------------------------------------------------- painting
    rng = np.random.default_rng(20261017)
    for tag, n, box, npart, xmin in (("a", 16, 100.0, 3000, 0.0), ("b", 12, 77.5, 2000, -5.0)):
        p = clustered_particles(rng, npart, box) + F32(xmin)
        w = (0.5 + rng.random(npart)).astype(F32)
        # edge cases: on a node, on the lower edge, just below the upper edge, outside the box
        p[0] = (xmin, xmin, xmin)
        p[1] = (xmin + box * 0.5, xmin + box * 0.25, xmin + box * 0.75)
        p[2] = np.nextafter(F32(xmin + box), F32(-1e30))
        p[3] = (xmin - 0.3 * box / n, xmin + 1.0, xmin + 2.0)          # negative grid coordinate
        p[4] = (xmin + box * (1 + 0.4 / n), xmin + 3.0, xmin + 4.0)     # beyond the box
        p[5] = (xmin + 1.0, xmin - 1.7 * box / n, xmin + box * (1 + 1.2 / n))
        out = {"particles": p, "weights": w, "n": n, "box": box, "xmin": xmin}
        zero = np.zeros((n, n, n), F32)
        pre = rng.random((n, n, n)).astype(F32)
        for wrap in (True, False):
            out[f"vec_wrap{int(wrap)}"] = np.asarray(
                mas.cic_mas_vec(zero, p[:, 0], p[:, 1], p[:, 2], w, npart, xmin, xmin, xmin, box, n, wrap))
            out[f"scan_wrap{int(wrap)}"] = np.asarray(
                mas.cic_mas(zero, p[:, 0], p[:, 1], p[:, 2], w, npart, xmin, xmin, xmin, box, n, wrap))
        out["pre"] = pre
        out["vec_accumulate"] = np.asarray(
            mas.cic_mas_vec(pre, p[:, 0], p[:, 1], p[:, 2], w, npart, xmin, xmin, xmin, box, n, True))
        np.savez_compressed(os.path.join(GOLD, f"ref_paint_{tag}.npz"), **out)

    # ---------------------------------------------------------------- estimators
    rng = np.random.default_rng(5)
    for tag, n, box, npart in (("a", 32, 1000.0, 60000), ("b", 24, 600.0, 30000), ("c", 15, 250.0, 8000)):
        p = clustered_particles(rng, npart, box)
        w = np.ones(npart, F32)
        rho = np.asarray(mas.cic_mas_vec(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, npart,
                                         0.0, 0.0, 0.0, box, n, True))
        delta = (rho / rho.mean() - F32(1.0)).astype(F32)       # tests/correlations.py:49-50
        kF = 2 * np.pi / box
        kny = np.pi * n / box
        edge_sets = {
            "kf": np.arange(kF, kny, kF).astype(F32),                               # C2/C4 style
            "fine": np.arange(1e-4, 2.0 * kny, 0.37 * kF).astype(F32),              # tests/correlations.py:76 style
            "wide": np.arange(0.003, kny, 2.5 * kF).astype(F32),                    # tests/voids.py:55 style
        }
        out = {"delta": delta, "n": n, "box": box}
        for name, ke in edge_sets.items():
            k3d, pk, nm = corr.powspec_vec(delta, box, ke)
            out[f"pk_{name}_edges"] = ke
            out[f"pk_{name}_k3D"] = np.asarray(k3d)
            out[f"pk_{name}_Pk3D"] = np.asarray(pk)
            out[f"pk_{name}_Nmodes3D"] = np.asarray(nm)
        k3d, pk, nm = corr.powspec_vec_fundamental(jaxshim._wrap(delta), box)
        out["pkf_k3D"], out["pkf_Pk3D"], out["pkf_Nmodes3D"] = map(np.asarray, (k3d, pk, nm))

        s_edges0 = np.arange(0.0, 0.45 * box, box / n * 1.7).astype(F32)
        s_edges1 = np.arange(box / n * 0.5, 0.45 * box, box / n * 1.3).astype(F32)
        for name, se in (("s0", s_edges0), ("s1", s_edges1)):
            r3d, xi3d, nmx = corr.xi_vec(delta, box, se)
            out[f"xi_{name}_edges"] = se
            out[f"xi_{name}_r3D"], out[f"xi_{name}_xi3D"], out[f"xi_{name}_Nmodes3D"] = map(np.asarray, (r3d, xi3d, nmx))
        r3d, xi3d, nmx = corr.xi_vec_fundamental(jaxshim._wrap(delta), box)
        out["xif_r3D"], out["xif_xi3D"], out["xif_Nmodes3D"] = map(np.asarray, (r3d, xi3d, nmx))

        k1, k2 = 4.0 * kF, 6.0 * kF
        theta = np.linspace(0.0, np.pi, 7).astype(F32)                # tests/bispec.py:54 style
        res = corr.bispec(delta, box, k1, k2, theta)
        out["bk_k1"], out["bk_k2"], out["bk_theta"] = F32(k1), F32(k2), theta
        for key, v in zip(("k_all", "Pk", "theta_out", "B", "Q"), res):
            out[f"bk_{key}"] = np.asarray(v)

        res = corr.compute_all_correlations(delta, box, s_edges0, edge_sets["kf"], k1, k2, theta)
        for i, v in enumerate(res):
            out[f"all_{i}"] = np.asarray(v)
        res = corr.compute_2pt_correlations(delta, box, s_edges0, edge_sets["kf"])
        for i, v in enumerate(res):
            out[f"twopt_{i}"] = np.asarray(v)
        np.savez_compressed(os.path.join(GOLD, f"ref_corr_{tag}.npz"), **out)
        print("wrote", tag, n, "Nmodes[kf][:4] =", out["pk_kf_Nmodes3D"][:4])


if __name__ == "__main__":
    main()
