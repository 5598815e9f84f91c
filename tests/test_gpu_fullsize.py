"""BASELINE.json configs[1] at its FULL size (1e8 particles, TSC, 512^3, kF-wide bins) through
size-independent properties: the oracle cannot run this in seconds, invariants can."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, BOX, NPART = 512, 2000.0, 100_000_000


@pytest.fixture(scope="module")
def c2():
    import jax_powspec_b200 as jps
    from jax_powspec_b200.mocks import lognormal_catalog
    x, y, z = lognormal_catalog(NPART, BOX, n_grid=256, seed=5, device="cuda")
    kF = 2.0 * math.pi / BOX
    ke = np.arange(kF, math.pi * N / BOX, kF).astype(np.float32)
    return jps, x, y, z, ke


def test_mass_linearity_and_permutation(c2):
    jps, x, y, z, ke = c2
    zero = torch.zeros((N, N, N), device="cuda")
    whole = jps.tsc_mas_vec(zero, x, y, z, None, NPART, 0., 0., 0., BOX, N, True)
    # sum of TSC weights is 1 per particle
    assert abs(whole.sum(dtype=torch.float64).item() - NPART) < 1e-6 * NPART
    # deposit is linear in the catalogue: first half, then the second half accumulated on top
    h = NPART // 2
    half = jps.tsc_mas_vec(zero, x[:h], y[:h], z[:h], None, h, 0., 0., 0., BOX, N, True)
    both = jps.tsc_mas_vec(half, x[h:], y[h:], z[h:], None, NPART - h, 0., 0., 0., BOX, N, True)
    tol = 4e-6 * torch.clamp(whole.abs(), min=1.0)
    assert bool(((both - whole).abs() <= tol).all())
    # and does not depend on the order of the particles (fixed-point tile sums; float32 only in the flush)
    perm = torch.randperm(NPART, device="cuda")
    shuffled = jps.tsc_mas_vec(zero, x[perm], y[perm], z[perm], None, NPART, 0., 0., 0., BOX, N, True)
    del perm
    assert bool(((shuffled - whole).abs() <= tol).all())
    # weights: w = 2 everywhere doubles the mesh exactly (powers of two commute with every rounding)
    w2 = torch.full((NPART,), 2.0, device="cuda")
    doubled = jps.tsc_mas_vec(zero, x, y, z, w2, NPART, 0., 0., 0., BOX, N, True)
    assert bool(((doubled - 2.0 * whole).abs() <= 2.0 * tol).all())


def test_fused_pipeline_equals_the_separate_calls_and_counts_every_mode(c2):
    jps, x, y, z, ke = c2
    zero = torch.zeros((N, N, N), device="cuda")
    rho = jps.tsc_mas_vec(zero, x, y, z, None, NPART, 0., 0., 0., BOX, N, True)
    k_a, pk_a, nm_a = jps.powspec_vec(rho / rho.mean() - 1.0, BOX, ke, mas_order=3)
    k_b, pk_b, nm_b = jps.paint_powspec(x, y, z, None, 0., 0., 0., BOX, N, ke, order=3)
    assert torch.equal(nm_a, nm_b) and torch.equal(k_a, k_b)
    # same density to ~1e-7 (mean folded through the DC mode vs divided in float32) -> same multipoles
    scale = pk_a[:, 0].abs()
    assert bool(((pk_a - pk_b).abs() <= 2e-4 * scale[:, None] + 1e-3).all())
    # mode counts are geometry: every stored half-space mode with kF <= |k| < last edge, once (Q7)
    ki = torch.fft.fftfreq(N, d=1.0 / N, device="cuda").to(torch.int64)
    ki[N // 2] = N // 2
    kz = torch.arange(N // 2 + 1, device="cuda", dtype=torch.int64)
    k2 = (ki[:, None, None] ** 2 + ki[None, :, None] ** 2 + kz[None, None, :] ** 2)
    kf = torch.sqrt(k2.to(torch.float32))
    kF = np.float32(2.0 * np.pi) / np.float32(BOX)
    e = torch.as_tensor((ke / kF).astype(np.float32), device="cuda")
    bins = torch.bucketize(kf, e, right=True) - 1
    bins = torch.where(kf == e[-1], torch.full_like(bins, len(e) - 2), bins)
    ok = (bins >= 0) & (bins < len(e) - 1)
    want = torch.bincount(bins[ok], minlength=len(e) - 1).to(torch.float32)
    assert torch.equal(nm_a, want)
    # shot-noise sanity at the smallest scales: P0 ~ 1/nbar within a factor of a few at k_Nyquist/1
    nbar = NPART / BOX ** 3
    assert 0.2 / nbar < float(pk_a[-1, 0]) < 5.0 / nbar
