"""BASELINE.json configs[2] ("all triangle bins"): the batched (k1, k2)-pair bispectrum call against a
loop of single `bispec` calls and against the f64 oracle.  (File name sorts after the older GPU
tests on purpose: this entry point was added last.)"""
import numpy as np
import pytest

from oracle import correlations as oc
from oracle import mas as om
from tests.util import clustered_particles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jps():
    import jax_powspec_b200
    return jax_powspec_b200


@pytest.fixture(scope="module")
def field():
    n, box = 48, 600.0
    p = clustered_particles(11, 60_000, box)
    rho = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, 0.0, 0.0, 0.0, box, n, True,
                   order=2, compat="reference", precision="f64")
    return (rho / rho.mean() - 1.0).astype(np.float32), box, n


def test_pairs_equal_single_calls(jps, field):
    delta, box, n = field
    kF = 2 * np.pi / box
    centres = (2 * kF * np.arange(1, 6)).astype(np.float32)              # shells centred at 2 kF j
    k1, k2 = jps.triangle_pairs(centres)
    theta = np.linspace(0.0, np.pi, 7).astype(np.float32)
    k_all, pk, th, B, Q = jps.bispec_pairs(delta, box, k1, k2, theta)
    assert k_all.shape == (15, 9) and pk.shape == (15, 9) and B.shape == (15, 7) and Q.shape == (15, 7)
    np.testing.assert_array_equal(th, theta)
    for p in range(k1.size):
        ka1, pk1, _, B1, Q1 = jps.bispec(delta, box, float(k1[p]), float(k2[p]), theta)
        np.testing.assert_array_equal(k_all[p], ka1)
        # same kernels, same inputs; only the float64 atomic order of the block sums may differ
        np.testing.assert_allclose(pk[p], pk1, rtol=1e-6, atol=0, equal_nan=True)
        np.testing.assert_allclose(B[p], B1, rtol=1e-5, atol=1e-6 * np.nanmax(np.abs(B1)), equal_nan=True)
        np.testing.assert_allclose(Q[p], Q1, rtol=1e-5, atol=1e-6 * np.nanmax(np.abs(Q1)), equal_nan=True)


def test_pairs_against_f64_oracle(jps, field):
    delta, box, n = field
    kF = 2 * np.pi / box
    k1 = np.float32([4 * kF, 4 * kF, 6 * kF])
    k2 = np.float32([4 * kF, 8 * kF, 8 * kF])
    theta = np.linspace(0.1, 3.0, 5).astype(np.float32)
    k_all, pk, _, B, Q = jps.bispec_pairs(delta, box, k1, k2, theta)
    for p in range(3):
        ka64, pk64, _, B64, Q64 = oc.bispec(delta, box, k1[p], k2[p], theta, precision="f64")
        np.testing.assert_allclose(k_all[p], ka64, rtol=3e-7)
        m = np.isfinite(pk64)
        assert np.abs(pk[p][m] - pk64[m]).max() <= 1e-5 * np.abs(pk64[m]).max()
        mb = np.isfinite(B64)
        assert np.abs(B[p][mb] - B64[mb]).max() <= 1e-5 * np.abs(B64[mb]).max()
        mq = np.isfinite(Q64)
        assert np.abs(Q[p][mq] - Q64[mq]).max() <= 1e-5 * np.abs(Q64[mq]).max()


def test_pairs_argument_errors(jps, field):
    delta, box, n = field
    with pytest.raises(ValueError):
        jps.bispec_pairs(delta, box, [0.1, 0.2], [0.1], [0.5])
    with pytest.raises(ValueError):
        jps.bispec_pairs(delta, box, [], [], [0.5])
