"""The NumPy restatement (oracle/) against the golden vectors produced by running the
UNMODIFIED reference sources through oracle/jaxshim.py (oracle/run_reference.py).
CPU only."""
import os

import numpy as np
import pytest

from oracle import correlations as oc
from oracle import mas as om

F32 = np.float32


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("variant", ["vec", "scan"])
@pytest.mark.parametrize("wrap", [True, False])
def test_paint_reference_compat(golden_dir, tag, variant, wrap):
    g = _load(golden_dir, f"ref_paint_{tag}.npz")
    p, w, n, box, xmin = g["particles"], g["weights"], int(g["n"]), float(g["box"]), float(g["xmin"])
    want = g[f"{variant}_wrap{int(wrap)}"]
    zero = np.zeros((n, n, n), F32)
    got32 = om.paint(zero, p[:, 0], p[:, 1], p[:, 2], w, xmin, xmin, xmin, box, n, wrap,
                     order=2, compat="reference", variant=variant, precision="f32")
    # faithful_f32 follows the same serial order: bit-exact against the shim-run reference
    np.testing.assert_array_equal(got32, want)
    got64 = om.paint(zero, p[:, 0], p[:, 1], p[:, 2], w, xmin, xmin, xmin, box, n, wrap,
                     order=2, compat="reference", variant=variant, precision="f64")
    np.testing.assert_allclose(got64, want, rtol=2e-6, atol=2e-6)


def test_paint_accumulates_into_input(golden_dir):
    g = _load(golden_dir, "ref_paint_a.npz")
    p, w, n, box, xmin = g["particles"], g["weights"], int(g["n"]), float(g["box"]), float(g["xmin"])
    got = om.paint(g["pre"], p[:, 0], p[:, 1], p[:, 2], w, xmin, xmin, xmin, box, n, True,
                   precision="f32")
    np.testing.assert_array_equal(got, g["vec_accumulate"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("edges", ["kf", "fine", "wide"])
def test_powspec(golden_dir, tag, edges):
    g = _load(golden_dir, f"ref_corr_{tag}.npz")
    delta, box = g["delta"], float(g["box"])
    ke = g[f"pk_{edges}_edges"]
    want_k, want_pk, want_nm = g[f"pk_{edges}_k3D"], g[f"pk_{edges}_Pk3D"], g[f"pk_{edges}_Nmodes3D"]
    for prec, rtol in (("f32", 2e-6), ("f64", 2e-5)):
        k3d, pk, nm = oc.powspec(delta, box, ke, precision=prec)
        np.testing.assert_array_equal(nm, want_nm.astype(np.int64))        # mode counts bit-exact
        np.testing.assert_array_equal(k3d, want_k)
        ok = want_nm > 0
        assert np.all(np.isnan(pk[~ok])) and np.all(np.isnan(want_pk[~ok]))  # Q11
        scale = np.abs(want_pk[ok][:, :1])
        np.testing.assert_allclose(pk[ok] / scale, want_pk[ok] / scale, rtol=0, atol=rtol * 10)
        np.testing.assert_allclose(pk[ok][:, 0], want_pk[ok][:, 0], rtol=rtol)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_powspec_fundamental(golden_dir, tag):
    g = _load(golden_dir, f"ref_corr_{tag}.npz")
    delta, box = g["delta"], float(g["box"])
    k3d, pk, nm = oc.powspec_fundamental(delta, box, precision="f32", compat="reference")
    np.testing.assert_array_equal(nm, g["pkf_Nmodes3D"].astype(np.int64))
    ok = nm > 0
    np.testing.assert_allclose(k3d[ok], g["pkf_k3D"][ok], rtol=1e-6)
    np.testing.assert_allclose(pk[ok][:, 0], g["pkf_Pk3D"][ok][:, 0], rtol=3e-6)
    k3d, pk, nm = oc.powspec_fundamental(delta, box, precision="f64", compat="reference")
    np.testing.assert_allclose(pk[ok][:, 0], g["pkf_Pk3D"][ok][:, 0], rtol=2e-5)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("name", ["s0", "s1"])
def test_xi(golden_dir, tag, name):
    g = _load(golden_dir, f"ref_corr_{tag}.npz")
    delta, box = g["delta"], float(g["box"])
    se = g[f"xi_{name}_edges"]
    want = g[f"xi_{name}_xi3D"]
    for prec, tol in (("f32", 1e-5), ("f64", 1e-4)):
        r3d, xi3d, nm = oc.xi(delta, box, se, precision=prec, guard_mu=False)
        np.testing.assert_array_equal(nm, g[f"xi_{name}_Nmodes3D"].astype(np.int64))
        np.testing.assert_array_equal(r3d, g[f"xi_{name}_r3D"])
        np.testing.assert_array_equal(np.isnan(xi3d), np.isnan(want))        # Q22
        m = ~np.isnan(want)
        scale = np.abs(want[:, :1]).max()
        np.testing.assert_allclose(xi3d[m] / scale, want[m] / scale, rtol=0, atol=tol)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_bispec(golden_dir, tag):
    g = _load(golden_dir, f"ref_corr_{tag}.npz")
    delta, box = g["delta"], float(g["box"])
    for prec, tol in (("f32", 2e-5), ("f64", 2e-4)):
        k_all, pk, th, B, Q = oc.bispec(delta, box, g["bk_k1"], g["bk_k2"], g["bk_theta"], precision=prec)
        np.testing.assert_array_equal(k_all, g["bk_k_all"])
        m = np.isfinite(g["bk_Pk"])
        np.testing.assert_array_equal(np.isfinite(pk), m)
        np.testing.assert_allclose(pk[m], g["bk_Pk"][m], rtol=tol)
        m = np.isfinite(g["bk_B"])
        np.testing.assert_array_equal(np.isfinite(B), m)
        scale = np.abs(g["bk_B"][m]).max()
        np.testing.assert_allclose(B[m] / scale, g["bk_B"][m] / scale, rtol=0, atol=tol)
        m = np.isfinite(g["bk_Q"])
        np.testing.assert_allclose(Q[m], g["bk_Q"][m], rtol=0, atol=tol * np.abs(g["bk_Q"][m]).max())


@pytest.mark.parametrize("tag", ["a", "c"])
def test_composites(golden_dir, tag):
    g = _load(golden_dir, f"ref_corr_{tag}.npz")
    delta, box = g["delta"], float(g["box"])
    se, ke = g["xi_s0_edges"], g["pk_kf_edges"]
    res = oc.compute_all_correlations(delta, box, se, ke, g["bk_k1"], g["bk_k2"], g["bk_theta"], precision="f32")
    assert len(res) == 11
    for i, got in enumerate(res):
        want = g[f"all_{i}"]
        got = np.asarray(got, dtype=np.float64)
        m = np.isfinite(want)
        np.testing.assert_array_equal(np.isfinite(got), m)
        scale = max(np.abs(want[m]).max(), 1e-30)
        np.testing.assert_allclose(got[m] / scale, want[m] / scale, rtol=0, atol=3e-5)
    res = oc.compute_2pt_correlations(delta, box, se, ke, precision="f32")
    assert len(res) == 5
    for i, got in enumerate(res):
        want = g[f"twopt_{i}"]
        got = np.asarray(got, dtype=np.float64)
        m = np.isfinite(want)
        scale = max(np.abs(want[m]).max(), 1e-30)
        np.testing.assert_allclose(got[m] / scale, want[m] / scale, rtol=0, atol=3e-5)


def test_xi_host_helpers_match_the_shim_run_reference(golden_dir):
    """xi_vec_coords / s_edges_conv (/root/reference/src/correlations.py:262-272): the product's host
    helpers (pure NumPy float32, importable without a GPU) against the shim-run reference, bit for bit."""
    from jax_powspec_b200.correlations import s_edges_conv, xi_vec_coords
    g = np.load(os.path.join(golden_dir, "ref_helpers.npz"))
    for i in range(3):
        dims, box, se = int(g[f"h{i}_dims"]), float(g[f"h{i}_box"]), g[f"h{i}_s_edges"]
        ke = s_edges_conv(dims, box, se)
        assert ke.dtype == np.float32
        np.testing.assert_array_equal(ke.view(np.uint32), g[f"h{i}_s_edges_conv"].view(np.uint32))
        r = np.asarray(xi_vec_coords(dims, box, ke), dtype=np.float32)
        np.testing.assert_array_equal(r.view(np.uint32), g[f"h{i}_xi_vec_coords"].view(np.uint32))
