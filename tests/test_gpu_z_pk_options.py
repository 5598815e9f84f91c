"""The two P(k) estimator options without a reference counterpart: Hermitian
mode weighting (Q7 corrected) and interlacing, through jps_powspec_ex, against oracle/correlations.py
(itself pinned on analytic tests in tests/test_oracle_analytic.py).  Sorts after the older GPU tests."""
import numpy as np
import pytest

from oracle import correlations as oc
from oracle import mas as om
from tests.test_gpu_powspec import _check_pk
from tests.util import F32, clustered_particles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jps():
    import jax_powspec_b200
    return jax_powspec_b200


@pytest.fixture(scope="module")
def cat():
    n, box = 64, 800.0
    return clustered_particles(23, 300_000, box), n, box


def _edges(n, box):
    kF = 2 * np.pi / box
    return np.arange(kF, np.pi * n / box, kF).astype(F32)


def _oracle_mesh(p, n, box, order, xmin=0.0):
    rho = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, xmin, xmin, xmin, box, n, True,
                   order=order, compat="fixed", precision="f64")
    return (rho / rho.mean() - 1.0).astype(F32)


@pytest.mark.parametrize("order", [2, 3])
def test_hermitian_weighting(jps, cat, order):
    p, n, box = cat
    delta = _oracle_mesh(p, n, box, order)
    ke = _edges(n, box)
    k3d, pk, nm, (sums, counts) = jps.powspec_vec(delta, box, ke, mas_order=order, mode_weighting="hermitian",
                                                  return_raw=True)
    k64, pk64, c64 = oc.powspec(delta, box, ke, mas_order=order, precision="f64", mode_weighting="hermitian")
    np.testing.assert_array_equal(counts, c64)                          # exact, also for the doubled planes
    np.testing.assert_array_equal(k3d, k64)
    _check_pk(pk, nm, pk64, c64)
    # the reference's half-space counting is untouched by the new table mode (shared LRU slots)
    _, pk_h, nm_h = jps.powspec_vec(delta, box, ke, mas_order=order)
    _, pk64_h, c64_h = oc.powspec(delta, box, ke, mas_order=order, precision="f64")
    _check_pk(pk_h, nm_h, pk64_h, c64_h)
    assert (c64 > c64_h).all()


@pytest.mark.parametrize("order,method", [(2, "atomic"), (3, "sorted"), (4, "auto")])
def test_interlaced_against_oracle(jps, cat, order, method):
    p, n, box = cat
    m1, m2 = jps.paint_interlaced(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], None, 0.0, 0.0, 0.0, box, n,
                                  order=order, method=method)
    half = F32(0.5) * (F32(box) / F32(n))
    o1 = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, 0.0, 0.0, 0.0, box, n, True,
                  order=order, compat="fixed", precision="f64")
    o2 = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, half, half, half, box, n, True,
                  order=order, compat="fixed", precision="f64")
    scale = max(o1.max(), o1.mean())
    assert np.abs(m1 - o1).max() <= 4e-6 * scale and np.abs(m2 - o2).max() <= 4e-6 * scale
    assert abs(m2.astype(np.float64).sum() - p.shape[0]) < 1e-6 * p.shape[0]
    ke = _edges(n, box)
    d1 = (o1 / o1.mean() - 1.0).astype(F32)
    d2 = (o2 / o2.mean() - 1.0).astype(F32)
    for weighting in ("half", "hermitian"):
        _, pk, nm = jps.powspec_vec(d1, box, ke, mas_order=order, delta2=d2, mode_weighting=weighting)
        _, pk64, c64 = oc.powspec(d1, box, ke, mas_order=order, precision="f64", delta2=d2, mode_weighting=weighting)
        _check_pk(pk, nm, pk64, c64, tol=2e-5)                          # + float32 sincospi of the phase factor
    # normalise=1 on the raw meshes = the density contrast folded into the kernels, interlaced too
    _, pk_raw, nm_raw = jps.powspec_vec(m1, box, ke, mas_order=order, delta2=m2, normalise=True)
    _, pk64, c64 = oc.powspec(d1, box, ke, mas_order=order, precision="f64", delta2=d2)
    _check_pk(pk_raw, nm_raw, pk64, c64, tol=3e-5)


def test_interlacing_leaves_a_band_limited_field_alone(jps):
    n, box = 32, 100.0
    g = np.arange(n, dtype=np.float64)
    gx, gy, gz = g[:, None, None], g[None, :, None], g[None, None, :]

    def field(shift):
        x, y, z = gx + shift, gy + shift, gz + shift
        return (0.3 * np.cos(2 * np.pi * (2 * x - 3 * y + 1 * z) / n + 0.4)
                + 0.2 * np.sin(2 * np.pi * (5 * x + 0 * y + 4 * z) / n)
                + 0.1 * np.cos(2 * np.pi * (-7 * x + 9 * y + 15 * z) / n)).astype(F32)
    d1, d2 = field(0.0), field(0.5)
    ke = _edges(n, box)
    _, plain, _ = jps.powspec_vec(d1, box, ke)
    _, inter, _ = jps.powspec_vec(d1, box, ke, delta2=d2)
    scale = np.nanmax(np.abs(plain))
    assert np.nanmax(np.abs(inter - plain)) < 1e-5 * scale
    _, wrong, _ = jps.powspec_vec(d1, box, ke, delta2=field(-0.5))
    assert np.nanmax(np.abs(wrong - plain)) > 0.1 * scale


def test_options_argument_errors_and_plan_reuse(jps, cat):
    p, n, box = cat
    delta = _oracle_mesh(p, n, box, 2)
    ke = _edges(n, box)
    with pytest.raises(ValueError):
        jps.powspec_vec(delta, box, ke, mode_weighting="full")
    with pytest.raises(ValueError):
        jps.powspec_vec(delta, box, ke, delta2=np.zeros((n, n, n // 2), F32))
    # the plan grows a shell field for the second spectrum; the other estimators keep working on it
    jps.powspec_vec(delta, box, ke, delta2=delta)
    theta = np.linspace(0.2, 2.5, 4).astype(F32)
    kF = 2 * np.pi / box
    _, pkb, _, B, _ = jps.bispec(delta, box, 4 * kF, 6 * kF, theta)
    _, pk64, _, B64, _ = oc.bispec(delta, box, F32(4 * kF), F32(6 * kF), theta, precision="f64")
    assert np.abs(B - B64).max() <= 1e-5 * np.abs(B64).max()
    _, pk, nm = jps.powspec_vec(delta, box, ke)
    _, pk64, c64 = oc.powspec(delta, box, ke, precision="f64")
    _check_pk(pk, nm, pk64, c64)
