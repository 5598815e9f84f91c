"""xi(s), bispectrum and the composite calls (CUDA, through the C ABI) against the golden vectors
from the shim-run reference and against the f64 oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import correlations as oc
from oracle import mas as om
from tests.util import F32, clustered_particles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jps():
    import jax_powspec_b200
    return jax_powspec_b200


def _close_scaled(got, want, tol):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    m = np.isfinite(want)
    np.testing.assert_array_equal(np.isfinite(got), m)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    if m.any():
        scale = max(np.abs(want[m]).max(), 1e-300)
        err = np.abs(got[m] - want[m]).max() / scale
        assert err <= tol, f"max scaled error {err:.3e} > {tol}"


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("name", ["s0", "s1"])
def test_golden_xi_vec(jps, golden_dir, tag, name):
    g = np.load(os.path.join(golden_dir, f"ref_corr_{tag}.npz"))
    r3d, xi, nm = jps.xi_vec(g["delta"], float(g["box"]), g[f"xi_{name}_edges"])
    np.testing.assert_array_equal(r3d, g[f"xi_{name}_r3D"])
    np.testing.assert_array_equal(nm, g[f"xi_{name}_Nmodes3D"])           # incl. inf for empty bins (Q11)
    _close_scaled(xi[:, 0], g[f"xi_{name}_xi3D"][:, 0], 2e-5)
    _close_scaled(xi, g[f"xi_{name}_xi3D"], 5e-5)                          # NaN pattern of Q22 included
    _, xi64, _ = oc.xi(g["delta"], float(g["box"]), g[f"xi_{name}_edges"], precision="f64")
    _close_scaled(xi, xi64, 2e-5)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_golden_xi_fundamental(jps, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"ref_corr_{tag}.npz"))
    r3d, xi, nm = jps.xi_vec_fundamental(g["delta"], float(g["box"]))
    np.testing.assert_array_equal(nm, g["xif_Nmodes3D"])
    ok = nm > 0
    np.testing.assert_allclose(r3d[ok], g["xif_r3D"][ok], rtol=3e-5)   # golden sums |r| serially in float32
    _close_scaled(xi[ok], g["xif_xi3D"][ok], 5e-5)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_golden_bispec(jps, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"ref_corr_{tag}.npz"))
    k_all, pk, th, B, Q = jps.bispec(g["delta"], float(g["box"]), float(g["bk_k1"]), float(g["bk_k2"]), g["bk_theta"])
    np.testing.assert_allclose(k_all, g["bk_k_all"], rtol=3e-7)            # host sinf/cosf vs numpy float32
    np.testing.assert_array_equal(th, g["bk_theta"])
    _close_scaled(pk, g["bk_Pk"], 3e-5)
    _close_scaled(B, g["bk_B"], 3e-5)
    _close_scaled(Q, g["bk_Q"], 3e-5)
    # the real bar: exact arithmetic
    _, pk64, _, B64, Q64 = oc.bispec(g["delta"], float(g["box"]), g["bk_k1"], g["bk_k2"], g["bk_theta"], precision="f64")
    _close_scaled(pk, pk64, 1e-5)
    _close_scaled(B, B64, 1e-5)
    _close_scaled(Q, Q64, 1e-5)


@pytest.mark.parametrize("tag", ["a", "c"])
def test_golden_composites(jps, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"ref_corr_{tag}.npz"))
    delta, box = g["delta"], float(g["box"])
    se, ke = g["xi_s0_edges"], g["pk_kf_edges"]
    res = jps.compute_all_correlations(delta, box, se, ke, float(g["bk_k1"]), float(g["bk_k2"]), g["bk_theta"])
    assert len(res) == 11
    for i, got in enumerate(res):
        _close_scaled(got, g[f"all_{i}"], 5e-5)
    np.testing.assert_array_equal(res[2], g["all_2"])                     # Nmodes3D_pk exact
    np.testing.assert_array_equal(res[5], g["all_5"])                     # Nmodes3D_xi exact (inf for empty)
    res2 = jps.compute_2pt_correlations(delta, box, se, ke)
    assert len(res2) == 5
    for i, got in enumerate(res2):
        _close_scaled(got, g[f"twopt_{i}"], 5e-5)
    # sharing one FFT must not change anything: composite == stand-alone calls
    k3d, pk, nm = jps.powspec_vec(delta, box, ke)
    np.testing.assert_allclose(res[1], pk, rtol=2e-6, equal_nan=True)     # accumulation order is not fixed
    r3d, xi, nmx = jps.xi_vec(delta, box, se, guard_mu=True)
    np.testing.assert_allclose(res[4], xi, rtol=1e-5, atol=1e-7)


def test_bispec_n128_reference_call_shape(jps):
    """tests/bispec.py:53-54 style call (20 angles) at N=128 against the f64 oracle, device tensors."""
    n, box, npart = 128, 1000.0, 1_000_000
    p = clustered_particles(31, npart, box)
    rho = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True,
                   order=2, compat="reference", precision="f64")
    delta = (rho / rho.mean() - 1.0).astype(F32)
    theta = np.linspace(0, np.pi, 20).astype(F32)
    d = torch.from_numpy(delta).cuda()
    k_all, pk, th, B, Q = jps.bispec(d, box, 0.1, 0.2, theta)
    assert B.is_cuda and B.shape == (20,) and pk.shape == (22,)
    _, pk64, _, B64, Q64 = oc.bispec(delta, box, 0.1, 0.2, theta, precision="f64")
    _close_scaled(pk.cpu().numpy(), pk64, 1e-5)
    _close_scaled(B.cpu().numpy(), B64, 1e-5)
    _close_scaled(Q.cpu().numpy(), Q64, 1e-5)
    # normalise=True on rho equals the call on delta
    k_all2, pk2, _, B2, Q2 = jps.bispec(torch.from_numpy(rho.astype(F32)).cuda(), box, 0.1, 0.2, theta, normalise=True)
    _close_scaled(B2.cpu().numpy(), B.cpu().numpy(), 1e-5)


def test_xi_n128_oracle(jps):
    n, box, npart = 128, 1000.0, 1_000_000
    p = clustered_particles(32, npart, box)
    rho = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True,
                   order=2, compat="reference", precision="f64")
    delta = (rho / rho.mean() - 1.0).astype(F32)
    se = np.arange(0.0, 200.0, 5.0).astype(F32)
    r3d, xi, nm = jps.xi_vec(delta, box, se, guard_mu=True)
    _, xi64, counts = oc.xi(delta, box, se, precision="f64", guard_mu=True)
    np.testing.assert_array_equal(nm.astype(np.int64), counts)
    _close_scaled(xi, xi64, 1e-5)


def test_bispec_indicator_cache_is_transparent(jps, golden_dir):
    """The sums over the shell indicator fields are data independent and cached on the device after the
    first call: repeated calls, calls with other (k1,k2) in between and calls on other data must agree
    with a cold computation."""
    from jax_powspec_b200.plan import clear_plans
    g = np.load(os.path.join(golden_dir, "ref_corr_a.npz"))
    delta, box = g["delta"], float(g["box"])
    k1, k2, th = float(g["bk_k1"]), float(g["bk_k2"]), g["bk_theta"]
    clear_plans()
    cold = jps.bispec(delta, box, k1, k2, th)
    warm = jps.bispec(delta, box, k1, k2, th)                  # everything cached
    for a, b in zip(cold, warm):
        np.testing.assert_allclose(a, b, rtol=2e-6, equal_nan=True)
    other = jps.bispec(delta, box, 0.8 * k1, k2, th)           # new triples, partly cached pairs
    back = jps.bispec(2.0 * delta, box, k1, k2, th)            # other data, cached geometry
    np.testing.assert_allclose(back[3], 8.0 * cold[3], rtol=1e-5)      # B scales as delta^3
    np.testing.assert_allclose(back[1], 4.0 * cold[1], rtol=1e-5)      # P scales as delta^2
    clear_plans()
    cold_other = jps.bispec(delta, box, 0.8 * k1, k2, th)
    for a, b in zip(other, cold_other):
        np.testing.assert_allclose(a, b, rtol=2e-6, equal_nan=True)
