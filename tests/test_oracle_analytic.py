"""Analytic known answers and independently written second formulations for the oracle (CPU).

The reference has no TSC / PCS painter and no golden vectors, so the rows
that cannot be pinned on it are pinned on mathematics instead: B-spline weights against
scipy.interpolate.BSpline, window exponents against the transform of the assignment kernel,
mode counts against a brute-force triple loop, P(k) against a full complex-FFT formulation and a
single plane wave."""
import numpy as np
import pytest
from scipy.interpolate import BSpline

from oracle import correlations as oc
from oracle import mas as om

F32 = np.float32


# ------------------------------------------------------------------ painting
def cardinal_bspline(order, t):
    """Centred cardinal B-spline of the given order (2 = CIC hat, 3 = TSC, 4 = PCS) at offsets t."""
    knots = np.arange(order + 1) - order / 2.0
    return np.nan_to_num(BSpline.basis_element(knots, extrapolate=False)(t))


@pytest.mark.parametrize("order", [2, 3, 4])
def test_bspline_weights_against_scipy(order):
    rng = np.random.default_rng(order)
    pos = rng.random(2000) * 40.0 + 5.0
    idx, w = om.bspline_axis(pos.astype(F32), order, np.float64)
    p = pos.astype(F32).astype(np.float64)
    want = np.stack([cardinal_bspline(order, idx[s] - p) for s in range(order)])
    np.testing.assert_allclose(w, want, atol=1e-12)
    np.testing.assert_allclose(w.sum(axis=0), 1.0, atol=1e-12)
    # every node outside the listed ones has zero weight: the stencil is complete
    for extra in (idx[0] - 1, idx[-1] + 1):
        assert np.abs(cardinal_bspline(order, extra - p)).max() < 1e-12


@pytest.mark.parametrize("order,centre,side", [(2, 1.0, None), (3, 0.75, 0.125), (4, 2.0 / 3.0, 1.0 / 6.0)])
def test_particle_on_a_node(order, centre, side):
    n, box = 8, 8.0
    mesh = om.paint(np.zeros((n, n, n)), [3.0], [4.0], [5.0], [2.0], 0.0, 0.0, 0.0, box, n, True,
                    order=order, compat="fixed", precision="f64")
    assert mesh[3, 4, 5] == pytest.approx(2.0 * centre ** 3)
    if side is not None:
        assert mesh[2, 4, 5] == pytest.approx(2.0 * side * centre ** 2)
        assert mesh[3, 4, 6] == pytest.approx(2.0 * side * centre ** 2)
        assert mesh[4, 5, 6] == pytest.approx(2.0 * side ** 3)
    assert mesh.sum() == pytest.approx(2.0)
    assert np.count_nonzero(mesh) == (1 if order == 2 else 27)


def test_reference_cic_midcell_quirk():
    """A particle at a cell centre: textbook CIC gives 8 x 1/8; the reference's corner (ix, iy+1, iz+1)
    uses mdx*mdy*ddz (src/mas.py:149, Q1) -- the same 1/8 here, so the quirk only shows off-centre."""
    n, box = 4, 4.0
    ref = om.paint(np.zeros((n, n, n)), [1.5], [1.5], [1.5], None, 0.0, 0.0, 0.0, box, n, True,
                   order=2, compat="reference", precision="f64")
    assert ref[1:3, 1:3, 1:3] == pytest.approx(np.full((2, 2, 2), 0.125))
    off = om.paint(np.zeros((n, n, n)), [1.25], [1.25], [1.75], None, 0.0, 0.0, 0.0, box, n, True,
                   order=2, compat="reference", precision="f64")
    mdx, ddx, mdy, ddy, mdz, ddz = 0.75, 0.25, 0.75, 0.25, 0.25, 0.75
    assert off[1, 2, 2] == pytest.approx(mdx * mdy * ddz)            # Q1: not mdx*ddy*ddz
    assert off.sum() == pytest.approx(1.0 + mdx * ddz * (mdy - ddy))  # SURVEY quirk list, Q1 mass
    fixed = om.paint(np.zeros((n, n, n)), [1.25], [1.25], [1.75], None, 0.0, 0.0, 0.0, box, n, True,
                     order=2, compat="fixed", precision="f64")
    assert fixed[1, 2, 2] == pytest.approx(mdx * ddy * ddz) and fixed.sum() == pytest.approx(1.0)


@pytest.mark.parametrize("order", [2, 3, 4])
def test_mass_conservation_and_translation(order):
    rng = np.random.default_rng(10 + order)
    n, box = 12, 36.0
    p = (rng.random((500, 3)) * box).astype(F32)
    w = rng.random(500).astype(F32) + 0.5
    mesh = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], w, 0.0, 0.0, 0.0, box, n, True,
                    order=order, compat="fixed", precision="f64")
    assert mesh.sum() == pytest.approx(w.astype(np.float64).sum(), rel=1e-12)
    # whole-cell shift = roll of the mesh; dyadic coordinates so that every float32 operation is exact
    n, box = 8, 16.0
    p = (rng.integers(0, 16 * 64, (500, 3)) / 64.0).astype(F32)
    q = p.copy()
    q[:, 0] = (q[:, 0] + 4.0) % box                                      # two cells
    a = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], w, 0.0, 0.0, 0.0, box, n, True,
                 order=order, compat="fixed", precision="f64")
    b = om.paint(np.zeros((n, n, n)), q[:, 0], q[:, 1], q[:, 2], w, 0.0, 0.0, 0.0, box, n, True,
                 order=order, compat="fixed", precision="f64")
    np.testing.assert_allclose(b, np.roll(a, 2, axis=0), atol=1e-12)


@pytest.mark.parametrize("order", [2, 3, 4])
def test_window_is_sinc_to_the_order(order):
    """The Fourier transform of the assignment kernel is sinc(pi k/N)^order per axis: painting ONE particle
    and deconvolving with the oracle's window must leave |delta_k| = 1 on every axis mode below Nyquist."""
    n, box = 16, 16.0
    x0 = 5.3
    mesh = om.paint(np.zeros((n, n, n)), [x0], [7.0], [9.0], None, 0.0, 0.0, 0.0, box, n, True,
                    order=order, compat="fixed", precision="f64")
    line = np.fft.fft(mesh.sum(axis=(1, 2)))                           # kx axis, ky = kz = 0
    c = oc.window_axis(n, order, "f64")
    k = np.arange(1, n // 2)
    # aliasing: sum over images m of sinc(pi (k + m N)/N)^order e^{...}; compare with that exact sum
    m = np.arange(-200, 201)[:, None]
    kk = k[None, :] + m * n
    alias = (np.sinc(kk / n) ** order * np.exp(-2j * np.pi * kk * x0 / n)).sum(axis=0)
    np.testing.assert_allclose(line[k], alias, atol=2e-4)              # the image sum converges like 1/m^order
    np.testing.assert_allclose(np.abs(c[k]), 1.0 / np.sinc(k / n) ** order, rtol=1e-12)


# ------------------------------------------------------------------ P(k)
def test_mode_counts_brute_force():
    n, box = 10, 100.0
    kF = 2 * np.pi / box
    edges = (np.array([0.5, 1.0, 1.5, 2.5, 4.0, 5.0, 9.0]) * kF).astype(F32)
    delta = np.random.default_rng(0).normal(size=(n, n, n)).astype(F32)
    _, _, counts = oc.powspec(delta, box, edges, precision="f64")
    ge = oc.grid_edges(edges, box)
    want = np.zeros(len(edges) - 1, dtype=np.int64)
    mid = n // 2
    for i in range(n):
        for j in range(n):
            for l in range(mid + 1):                                    # half space, every stored cell once (Q7)
                kx, ky, kz = (i - n if i > mid else i), (j - n if j > mid else j), l
                k = np.sqrt(F32(kx * kx + ky * ky + kz * kz))
                b = np.searchsorted(ge, k, side="right") - 1
                if k == ge[-1]:
                    b = len(ge) - 2
                if 0 <= b < len(want) and not (k > ge[-1]):
                    want[b] += 1
    np.testing.assert_array_equal(counts, want)


def test_powspec_against_full_complex_fft():
    """Second formulation: full fftn, then keep the half space -- no rfftn, no shared helper."""
    rng = np.random.default_rng(2)
    n, box = 16, 320.0
    delta = rng.normal(size=(n, n, n)).astype(F32)
    kF = 2 * np.pi / box
    edges = (np.arange(1, 9) * kF).astype(F32)
    k3d, pk, counts = oc.powspec(delta, box, edges, mas_order=2, precision="f64")
    full = np.fft.fftn(delta.astype(np.float64))
    freq = np.array([i - n if i > n // 2 else i for i in range(n)])
    win = 1.0 / np.sinc(freq / n) ** 2
    full = full * win[:, None, None] * win[None, :, None] * win[None, None, :]
    kx, ky, kz = np.meshgrid(freq, freq, freq, indexing="ij")
    half = np.zeros((n, n, n), bool)
    half[:, :, : n // 2 + 1] = True                                     # the rfft half: indices 0..N/2 of the last axis
    kmag = np.sqrt((kx * kx + ky * ky + kz * kz).astype(F32))
    ge = (edges / F32(kF)).astype(F32)
    p2 = np.abs(full) ** 2
    k2i = kx * kx + ky * ky + kz * kz
    mu2 = np.where(k2i > 0, kz * kz / np.where(k2i > 0, k2i, 1).astype(np.float64), 0.0)
    vol = (box / n ** 2) ** 3
    for b in range(len(edges) - 1):
        last = b == len(edges) - 2
        sel = half & (kmag >= ge[b]) & ((kmag <= ge[b + 1]) if last else (kmag < ge[b + 1]))
        assert sel.sum() == counts[b]
        v = p2[sel]
        assert pk[b, 0] == pytest.approx(v.mean() * vol, rel=1e-10)
        assert pk[b, 1] == pytest.approx((v * (3 * mu2[sel] - 1) / 2).mean() * 5 * vol, rel=1e-9, abs=1e-9 * pk[b, 0])
        assert pk[b, 2] == pytest.approx((v * (35 * mu2[sel] ** 2 - 30 * mu2[sel] + 3) / 8).mean() * 9 * vol,
                                         rel=1e-9, abs=1e-9 * pk[b, 0])
    np.testing.assert_allclose(k3d, 0.5 * (edges[1:] + edges[:-1]), rtol=1e-6)   # bin centres, not mean k (Q10)


def test_plane_wave_known_answer():
    """delta = A cos(2 pi m.x / N): all power in the bin holding |m|, amplitude (A N^3/2)^2 per stored mode."""
    n, box, amp = 16, 160.0, 0.3
    m = np.array([2, -1, 2])                                            # |m| = 3
    g = np.arange(n)
    phase = 2 * np.pi * (m[0] * g[:, None, None] + m[1] * g[None, :, None] + m[2] * g[None, None, :]) / n
    delta = (amp * np.cos(phase)).astype(F32)
    kF = 2 * np.pi / box
    edges = (np.array([0.5, 2.5, 3.5, 6.0]) * kF).astype(F32)
    _, pk, counts = oc.powspec(delta, box, edges, mas_order=2, precision="f64")
    assert pk[0, 0] == pytest.approx(0.0, abs=1e-6) and pk[2, 0] == pytest.approx(0.0, abs=1e-6)
    freq = np.array([2, -1, 2])
    win = np.prod(1.0 / np.sinc(freq / n) ** 2)
    # kz = 2 > 0: only +m is stored in the half space (its partner -m has kz = -2)
    want = (amp * n ** 3 / 2) ** 2 * win ** 2 / counts[1] * (box / n ** 2) ** 3
    assert pk[1, 0] == pytest.approx(want, rel=1e-5)
    mu2 = 4.0 / 9.0
    assert pk[1, 1] / pk[1, 0] == pytest.approx(5 * (3 * mu2 - 1) / 2, rel=1e-5)
    assert pk[1, 2] / pk[1, 0] == pytest.approx(9 * (35 * mu2 ** 2 - 30 * mu2 + 3) / 8, rel=1e-5)


def test_bispectrum_indicator_sum_identity():
    """Appendix A of SURVEY.md: sum_x I_j^2 = (# full-space modes in shell j) / N^3, half-space cells with
    0 < kz < N/2 counting twice -- the identity the device cache of indicator sums rests on."""
    n, box = 12, 120.0
    kF = F32(2 * np.pi) / F32(box)
    k_all, lo, hi = oc.bispec_shells(box, 3 * kF, 4 * kF, np.array([0.7, 2.0], dtype=F32))
    kx, ky, kz, k2 = oc._half_grids(n)
    kf = np.sqrt(k2.astype(F32))
    for j in range(len(k_all)):
        msk = (kf >= lo[j]) & (kf < hi[j])
        ind = np.fft.irfftn(msk.astype(np.complex128), s=(n, n, n), axes=(0, 1, 2))
        weight = np.where((kz == 0) | (kz == n // 2), 1, 2) * np.ones_like(k2)
        assert (ind * ind).sum() == pytest.approx((msk * weight).sum() / n ** 3, rel=1e-10)


# ------------------------------------------------------------------ extensions (no reference counterpart)
def test_hermitian_weighting_is_the_full_space_average():
    rng = np.random.default_rng(5)
    n, box = 12, 240.0
    delta = rng.normal(size=(n, n, n)).astype(F32)
    kF = 2 * np.pi / box
    edges = (np.arange(1, 7) * kF).astype(F32)
    _, pk, counts = oc.powspec(delta, box, edges, precision="f64", mode_weighting="hermitian")
    full = np.fft.fftn(delta.astype(np.float64))
    freq = np.array([i - n if i > n // 2 else i for i in range(n)])
    win = 1.0 / np.sinc(freq / n) ** 2
    p2 = np.abs(full * win[:, None, None] * win[None, :, None] * win[None, None, :]) ** 2
    kx, ky, kz = np.meshgrid(freq, freq, freq, indexing="ij")
    kmag = np.sqrt((kx * kx + ky * ky + kz * kz).astype(F32))
    ge = (edges / F32(kF)).astype(F32)
    # the full n^3 grid holds every mode of the half space once plus the mirror images of the interior planes;
    # index N/2 along z is its own mirror, so the full grid IS the Hermitian-weighted half space
    for b in range(len(edges) - 1):
        last = b == len(edges) - 2
        sel = (kmag >= ge[b]) & ((kmag <= ge[b + 1]) if last else (kmag < ge[b + 1]))
        assert sel.sum() == counts[b]
        assert pk[b, 0] == pytest.approx(p2[sel].mean() * (box / n ** 2) ** 3, rel=1e-10)
    half_counts = oc.powspec(delta, box, edges, precision="f64")[2]
    assert (counts > half_counts).all() and (counts <= 2 * half_counts).all()


def test_interlacing_leaves_a_band_limited_field_alone():
    """Sampled (not painted) plane waves on the two grids: no aliases to cancel, so the interlaced spectrum
    equals the plain one -- this fixes the SIGN of the phase factor and of the half-cell displacement."""
    n, box = 16, 100.0
    g = np.arange(n, dtype=np.float64)
    gx, gy, gz = g[:, None, None], g[None, :, None], g[None, None, :]
    def field(shift):
        x, y, z = gx + shift, gy + shift, gz + shift
        return (0.3 * np.cos(2 * np.pi * (2 * x - 3 * y + 1 * z) / n + 0.4)
                + 0.2 * np.sin(2 * np.pi * (5 * x + 0 * y + 4 * z) / n)).astype(np.float64)
    d1, d2 = field(0.0), field(0.5)                                       # grid 2 displaced by +half a cell
    kF = 2 * np.pi / box
    edges = (np.arange(1, 9) * kF).astype(F32)
    _, plain, _ = oc.powspec(d1, box, edges, precision="f64")
    _, inter, _ = oc.powspec(d1, box, edges, precision="f64", delta2=d2)
    np.testing.assert_allclose(inter, plain, rtol=1e-9, atol=1e-9 * np.abs(plain).max())
    # displaced the other way the odd-sum modes flip sign instead: power is NOT preserved
    _, wrong, _ = oc.powspec(d1, box, edges, precision="f64", delta2=field(-0.5))
    assert np.abs(wrong - plain).max() > 0.1 * np.abs(plain).max()


@pytest.mark.parametrize("order", [2, 3])
def test_interlacing_suppresses_aliasing(order):
    """Poisson particles: the painted + deconvolved spectrum rises above the shot-noise level towards
    Nyquist (aliased images); interlacing removes the odd images and brings it back down."""
    rng = np.random.default_rng(17)
    n, box, npart = 24, 240.0, 40_000
    p = (rng.random((npart, 3)) * box).astype(F32)
    cell = F32(box / n)
    half = F32(0.5) * cell
    def paint(xmin):
        rho = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, xmin, xmin, xmin, box, n, True,
                       order=order, compat="fixed", precision="f64")
        return rho / rho.mean() - 1.0
    d1, d2 = paint(F32(0.0)), paint(half)
    kF = 2 * np.pi / box
    edges = (np.arange(1, 16) * kF).astype(F32)                           # through Nyquist (12 kF) into the corners
    shot = box ** 3 / npart
    _, plain, _ = oc.powspec(d1, box, edges, mas_order=order, precision="f64")
    _, inter, _ = oc.powspec(d1, box, edges, mas_order=order, precision="f64", delta2=d2)
    at_nyq = slice(10, 13)                                                # the bins around k_Nyquist
    excess_plain = np.abs(plain[at_nyq, 0] / shot - 1).mean()             # CIC 0.41, TSC 0.25
    excess_inter = np.abs(inter[at_nyq, 0] / shot - 1).mean()             # CIC 0.02, TSC 0.02
    assert excess_plain > 0.15 and excess_inter < 0.25 * excess_plain, (excess_plain, excess_inter)
    # large scales are untouched
    np.testing.assert_allclose(inter[:5, 0], plain[:5, 0], rtol=2e-2)
