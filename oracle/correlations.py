"""CPU oracle for the Fourier-space estimators -- TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, in vectorised NumPy, the arithmetic of /root/reference/src/correlations.py:
  powspec_vec :7-56, powspec_vec_fundamental :60-117, xi_vec :120-187,
  xi_vec_fundamental :191-261, bispec :334-462, compute_all_correlations :464-637,
  compute_2pt_correlations :640-712   (quirks Q7-Q22 of SURVEY.md section 8).

Bin membership is ALWAYS decided in float32 exactly as the reference does
(kedges = k_edges / kF in f32, k = sqrt_f32(kx^2+ky^2+kz^2), searchsorted 'right',
last edge inclusive) because mode counts must match bit for bit.  ``precision``
selects the arithmetic of the values:
  "f64": FFT, window, |delta_k|^2, Legendre weights and sums in float64 -> exact_f64
  "f32": float32 in the reference's operation order, serial float32 sums -> faithful_f32
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft

F32 = np.float32
_PI32 = F32(np.pi)


def k_fundamental(box_size):
    """kF = 2.0*pi/box_size evaluated in float32 (box_size is a traced f32 under jit). :12"""
    return F32(2.0 * np.pi) / F32(box_size)


def k_index(n):
    """ki = i - N if i > N//2 else i  (int32).  correlations.py:25 (Q15)."""
    r = np.arange(n, dtype=np.int32)
    return np.where(r > n // 2, r - n, r).astype(np.int32)


def window_axis(n, mas_order, precision):
    """Per-axis deconvolution factor (1/sinc(pi k/N))**p.  correlations.py:15,20-21,32 (Q13).
    p = 2 is what the reference hard-codes (CIC); p = 3 / 4 generalise it to TSC / PCS."""
    ki = k_index(n)
    p = int(mas_order)
    if precision == "f64":
        x = np.pi * ki.astype(np.float64) / n
        s = np.where(ki == 0, 1.0, np.sin(x) / np.where(ki == 0, 1.0, x))
        return (1.0 / s) ** p
    pref = F32(np.pi / n)                       # jnp.pi / dims is a Python double, cast when used
    x = pref * ki.astype(F32)
    y = x / _PI32                               # jnp.sinc argument
    pix = _PI32 * y
    safe = np.where(y == 0, F32(1.0), pix)
    s = np.where(y == 0, F32(1.0), np.sin(safe) / safe).astype(F32)
    r = F32(1.0) / s
    out = r
    for _ in range(p - 1):
        out = out * r
    return out.astype(F32)


def grid_edges(k_edges, box_size):
    """kedges = k_edges / kF in float32.  correlations.py:42 (Q9)."""
    return (np.asarray(k_edges, dtype=F32) / k_fundamental(box_size)).astype(F32)


def bin_index_from_k(k, kedges):
    """jnp.histogram membership: returns bin in [0, nb) or -1.  (Q9)"""
    idx = np.searchsorted(kedges, k, side="right")
    idx = np.where(k == kedges[-1], len(kedges) - 1, idx)
    b = idx - 1
    return np.where((idx >= 1) & (idx <= len(kedges) - 1), b, -1)


def _half_grids(n):
    ki = k_index(n)
    mid = n // 2
    kx = ki[:, None, None]
    ky = ki[None, :, None]
    kz = ki[None, None, : mid + 1]
    k2 = (kx.astype(np.int64) ** 2 + ky.astype(np.int64) ** 2 + kz.astype(np.int64) ** 2)
    return kx, ky, kz, k2


def deconvolved_dk(delta, mas_order, precision):
    """rfftn + window deconvolution.  correlations.py:22,32,39."""
    n = delta.shape[0]
    mid = n // 2
    c = window_axis(n, mas_order, precision)
    if precision == "f64":
        dk = sfft.rfftn(np.asarray(delta, dtype=np.float64), workers=-1)
        corr = c[:, None, None] * c[None, :, None] * c[None, None, : mid + 1]
        return dk * corr
    dk = sfft.rfftn(np.asarray(delta, dtype=F32), workers=-1).astype(np.complex64, copy=False)
    corr = (c[:, None, None] * c[None, :, None]) * c[None, None, : mid + 1]
    return (dk * corr.astype(F32)).astype(np.complex64)


def _weighted(v, mu2, dtype):
    """v*L2(mu), v*L4(mu) in the reference's operation order:
    v * (3.0*mu2-1.0)/2.0 ; v * (35.0*mu2*mu2 - 30.0*mu2 + 3.0)/8.0   (correlations.py:46-47)."""
    t = dtype
    v = v.astype(t)
    mu2 = mu2.astype(t)
    w2 = v * (t(3.0) * mu2 - t(1.0)) / t(2.0)
    w4 = v * (t(35.0) * mu2 * mu2 - t(30.0) * mu2 + t(3.0)) / t(8.0)
    return w2, w4


def _mu2_half(n, k2, kz, precision):
    kf = np.sqrt(k2.astype(F32))                # jnp.sqrt(int32) -> float32
    if precision == "f64":
        k = np.sqrt(k2.astype(np.float64))
        mu = np.where(k2 == 0, 0.0, kz / np.where(k2 == 0, 1.0, k))
        return kf, mu * mu
    mu = np.where(kf == 0, F32(0.0), kz.astype(F32) / np.where(kf == 0, F32(1.0), kf)).astype(F32)
    return kf, (mu * mu).astype(F32)


def _accumulate(bins, nb, weights, precision):
    ok = bins >= 0
    if precision == "f64":
        return np.bincount(bins[ok], weights=weights[ok], minlength=nb)[:nb]
    out = np.zeros(nb, dtype=F32)
    np.add.at(out, bins[ok], weights[ok].astype(F32))   # serial float32, element order
    return out


def _binned_multipoles(v, kf, mu2, kedges, precision):
    """Sums of v * {1, L2, L4} and mode counts per bin.  correlations.py:45-48 (Q7,Q16)."""
    dt = np.float64 if precision == "f64" else F32
    nb = len(kedges) - 1
    bins = bin_index_from_k(kf.ravel(), kedges)
    v = v.astype(dt).ravel()
    w2, w4 = _weighted(v, mu2.ravel(), dt)
    s0 = _accumulate(bins, nb, v, precision)
    s2 = _accumulate(bins, nb, w2, precision)
    s4 = _accumulate(bins, nb, w4, precision)
    counts = np.bincount(bins[bins >= 0], minlength=nb)[:nb].astype(np.int64)
    return s0, s2, s4, counts


def interlace_phase(n, dtype=np.complex128):
    """exp(-i pi (kx + ky + kz) / N) on the half-space array, frequencies as k_index() (index N/2 is +N/2,
    Q15).  No reference counterpart (the reference has no interlacing): second mesh painted on a grid
    displaced by +half a cell, i.e. with xmin + cell/2 (Sefusatti et al. 2016, eq. 17-18).  Modes on the
    Nyquist planes are their own mirror images up to the sign of N/2, so their phase is ambiguous: as in
    other implementations they are not meaningful in an interlaced estimate."""
    ki = k_index(n).astype(np.float64)
    ph = np.exp(-1j * np.pi * ki / n)
    mid = n // 2
    return (ph[:, None, None] * ph[None, :, None] * ph[None, None, : mid + 1]).astype(dtype)


def powspec(delta, box_size, k_edges, *, mas_order=2, precision="f64", shot_noise=0.0, delta2=None,
            mode_weighting="half"):
    """P0, P2, P4 in user k-bins.  correlations.py:7-56.
    Returns (k3D f32[nb], Pk3D [nb,3], Nmodes int64[nb]).  ``shot_noise`` (absent in the
    reference, Q12) is subtracted from the monopole only.

    Extensions without a reference counterpart (parity unpinned):
    ``delta2`` = the same particles painted on the grid displaced by half a cell -> interlaced
    spectrum (dk1 + dk2 * phase)/2, odd aliases cancel; ``mode_weighting='hermitian'`` counts every
    stored mode with 0 < kz < N/2 twice (the full-space average the reference's half-space sum
    only approximates, Q7)."""
    delta = np.asarray(delta)
    n = delta.shape[0]
    dt = np.float64 if precision == "f64" else F32
    kedges = grid_edges(k_edges, box_size)
    kx, ky, kz, k2 = _half_grids(n)
    dk = deconvolved_dk(delta, mas_order, precision)
    if delta2 is not None:
        dk2 = deconvolved_dk(np.asarray(delta2), mas_order, precision)
        ct = np.complex128 if precision == "f64" else np.complex64
        dk = ((dk + dk2 * interlace_phase(n, ct)) * dt(0.5)).astype(ct)
    d2 = (dk.real * dk.real + dk.imag * dk.imag)
    kzb = np.broadcast_to(kz, k2.shape)
    kf, mu2 = _mu2_half(n, k2, kzb, precision)
    if mode_weighting == "hermitian":
        twice = (kzb > 0) & (2 * kzb != n)
        mult = np.where(twice, 2, 1)
        s0, s2, s4, counts = _binned_multipoles(d2 * mult.astype(d2.dtype), kf, mu2, kedges, precision)
        bins = bin_index_from_k(kf.ravel(), kedges)
        ok = bins >= 0
        counts = np.bincount(bins[ok], weights=mult.ravel()[ok], minlength=len(kedges) - 1)[: len(kedges) - 1].astype(np.int64)
    elif mode_weighting == "half":
        s0, s2, s4, counts = _binned_multipoles(d2, kf, mu2, kedges, precision)
    else:
        raise ValueError("mode_weighting must be 'half' or 'hermitian'")
    vol = (dt(box_size) / dt(n * n)) ** 3 if precision == "f64" else (F32(box_size) / F32(n * n)) ** 3
    nm = counts.astype(dt)
    with np.errstate(invalid="ignore", divide="ignore"):
        pk = np.stack([s0 / nm * vol - dt(shot_noise), s2 / nm * dt(5.0) * vol, s4 / nm * dt(9.0) * vol], axis=1)
    k3d = (F32(0.5) * (kedges[1:] + kedges[:-1]) * k_fundamental(box_size)).astype(F32)
    return k3d, pk, counts


def fundamental_bins(n):
    """bin = int32(k); size kmax+1 with kmax = int32(sqrt(3 * (N//2)**2)).  correlations.py:68,87 (Q18)."""
    mid = n // 2
    kmax = int(np.int32(np.sqrt(F32(3 * mid * mid))))
    return kmax


def powspec_fundamental(delta, box_size, *, mas_order=2, precision="f64", compat="reference"):
    """Same with kF-wide integer bins.  correlations.py:60-117 (Q18).
    compat='reference': k3D[bin] = (k of the LAST stored mode of that bin in C order)/Nmodes*kF
    (the reference uses .set, not .add); compat='fixed': mean k of the bin."""
    delta = np.asarray(delta)
    n = delta.shape[0]
    dt = np.float64 if precision == "f64" else F32
    kmax = fundamental_bins(n)
    kx, ky, kz, k2 = _half_grids(n)
    dk = deconvolved_dk(delta, mas_order, precision)
    d2 = (dk.real * dk.real + dk.imag * dk.imag)
    kf, mu2 = _mu2_half(n, k2, np.broadcast_to(kz, k2.shape), precision)
    bins = kf.astype(np.int32).ravel()           # trunc
    bins = np.where(bins <= kmax, bins, -1)      # scatter drops out-of-range
    d2r = d2.astype(dt).ravel()
    w2, w4 = _weighted(d2r, mu2.ravel(), dt)
    nb = kmax + 1
    s0 = _accumulate(bins, nb, d2r, precision)
    s2 = _accumulate(bins, nb, w2, precision)
    s4 = _accumulate(bins, nb, w4, precision)
    counts = np.bincount(bins[bins >= 0], minlength=nb)[:nb].astype(np.int64)
    ksum = np.zeros(nb, dtype=dt)
    ok = bins >= 0
    if compat == "reference":
        ksum[bins[ok]] = kf.ravel()[ok].astype(dt)        # duplicates: last wins
    else:
        ksum = np.bincount(bins[ok], weights=kf.ravel()[ok].astype(np.float64), minlength=nb)[:nb].astype(dt)
    vol = (dt(box_size) / dt(n * n)) ** 3
    nm = counts[1:].astype(dt)
    with np.errstate(invalid="ignore", divide="ignore"):
        pk = np.stack([s0[1:] / nm * vol, s2[1:] / nm * dt(5.0) * vol, s4[1:] / nm * dt(9.0) * vol], axis=1)
        k3d = ksum[1:] / nm * dt(k_fundamental(box_size))
    return k3d, pk, counts[1:]


# ----------------------------------------------------------------------------- xi(s)
def s_edges_to_grid(s_edges, box_size, n):
    """k_edges = kF*s_edges*dims/box_size ; kedges = k_edges/kF  (float32).  correlations.py:126,170."""
    kF = k_fundamental(box_size)
    ke = kF * np.asarray(s_edges, dtype=F32) * F32(n) / F32(box_size)
    return (ke / kF).astype(F32)


def xi(delta, box_size, s_edges, *, mas_order=2, precision="f64", guard_mu=False):
    """xi_0,2,4(s).  correlations.py:120-187 (guard_mu=False: NaN for the r=0 cell enters
    the first bin of xi2/xi4 when s_edges[0] == 0, Q22) and the xi blocks of the composites
    :522-543 / :689-710 (guard_mu=True)."""
    delta = np.asarray(delta)
    n = delta.shape[0]
    dt = np.float64 if precision == "f64" else F32
    dk = deconvolved_dk(delta, mas_order, precision)
    d2 = (dk.real * dk.real + dk.imag * dk.imag)
    if precision == "f64":
        dxi = sfft.irfftn(d2.astype(np.complex128), s=(n, n, n), workers=-1)
    else:
        dxi = sfft.irfftn(d2.astype(np.complex64), s=(n, n, n), workers=-1).astype(F32)
    ki = k_index(n)
    rx, ry, rz = ki[:, None, None], ki[None, :, None], ki[None, None, :]
    r2 = rx.astype(np.int64) ** 2 + ry.astype(np.int64) ** 2 + rz.astype(np.int64) ** 2
    rf = np.sqrt(r2.astype(F32))
    with np.errstate(invalid="ignore", divide="ignore"):
        if precision == "f64":
            mu = np.broadcast_to(rz, r2.shape) / np.sqrt(r2.astype(np.float64))
        else:
            mu = (np.broadcast_to(rz, r2.shape).astype(F32) / rf).astype(F32)
    if guard_mu:
        mu = np.where(r2 == 0, dt(0.0), mu)
    mu2 = (mu * mu).astype(dt)
    kedges = s_edges_to_grid(s_edges, box_size, n)
    nb = len(kedges) - 1
    bins = bin_index_from_k(rf.ravel(), kedges)
    v = dxi.astype(dt).ravel()
    w2, w4 = _weighted(v, mu2.ravel(), dt)
    s0 = _accumulate(bins, nb, v, precision)
    s2 = _accumulate(bins, nb, w2, precision)
    s4 = _accumulate(bins, nb, w4, precision)
    counts = np.bincount(bins[bins >= 0], minlength=nb)[:nb].astype(np.int64)
    nm = np.where(counts == 0, np.inf, counts.astype(dt))      # Q11
    n3 = dt(n) ** 3
    xi3d = np.stack([s0 / nm / n3, s2 / nm * dt(5.0) / n3, s4 / nm * dt(9.0) / n3], axis=1)
    r3d = (F32(0.5) * (kedges[1:] + kedges[:-1]) * (F32(box_size) * F32(1.0) / F32(n))).astype(F32)
    return r3d, xi3d, counts


# ----------------------------------------------------------------------------- bispectrum
def bispec_shells(box_size, k1, k2, theta):
    """k_all, and shell bounds in grid units (float32).  correlations.py:347-357 (Q19)."""
    theta = np.asarray(theta, dtype=F32)
    k1, k2 = F32(k1), F32(k2)
    kF = k_fundamental(box_size)
    k3 = np.sqrt((k2 * np.sin(theta)) ** 2 + (k2 * np.cos(theta) + k1) ** 2).astype(F32)
    k_all = np.concatenate([[k1, k2], k3]).astype(F32)
    lo = ((k_all - kF) / kF).astype(F32)
    hi = ((k_all + kF) / kF).astype(F32)
    return k_all, lo, hi


def bispec(delta, box_size, k1, k2, theta, *, mas_order=2, precision="f64"):
    """FFT bispectrum.  correlations.py:334-462 (Q19-Q21).
    Returns (k_all, Pk[bins+2], theta, B[bins], Q[bins])."""
    delta = np.asarray(delta)
    n = delta.shape[0]
    dt = np.float64 if precision == "f64" else F32
    ct = np.complex128 if precision == "f64" else np.complex64
    k_all, lo, hi = bispec_shells(box_size, k1, k2, theta)
    kx, ky, kz, k2g = _half_grids(n)
    kf = np.sqrt(k2g.astype(F32))
    dk = deconvolved_dk(delta, mas_order, precision)

    def shell_fields(j):
        m = (kf >= lo[j]) & (kf < hi[j])
        d = sfft.irfftn((m * dk).astype(ct), s=(n, n, n), workers=-1).astype(dt)
        i = sfft.irfftn(m.astype(ct), s=(n, n, n), workers=-1).astype(dt)
        return d, i

    box = dt(box_size)
    vol_p = (box / dt(n * n)) ** 3
    vol_b = (box * box / dt(n) ** 3) ** 3
    nsh = len(k_all)
    pk = np.zeros(nsh, dtype=dt)
    d1, i1 = shell_fields(0)
    d2_, i2 = shell_fields(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        pk[0] = np.nansum(d1 * d1) / (i1 * i1).sum() * vol_p
        pk[1] = np.nansum(d2_ * d2_) / (i2 * i2).sum() * vol_p
        bins = nsh - 2
        B = np.zeros(bins, dtype=dt)
        Q = np.zeros(bins, dtype=dt)
        for b in range(bins):
            d3, i3 = shell_fields(b + 2)
            pk[b + 2] = np.nansum(d3 * d3) / np.nansum(i3 * i3) * vol_p
            B[b] = np.nansum(d1 * d2_ * d3) / (i1 * i2 * i3).sum() * vol_b
            Q[b] = B[b] / (pk[0] * pk[1] + pk[0] * pk[b + 2] + pk[1] * pk[b + 2])
    return k_all, pk, np.asarray(theta, dtype=F32), B, Q


def compute_2pt_correlations(delta, box_size, s_edges, k_edges, **kw):
    """correlations.py:640-712: (k3D, Pk3D, Nmodes3D_pk, r3D, xi3D)."""
    k3d, pk, nm = powspec(delta, box_size, k_edges, **kw)
    r3d, xi3d, _ = xi(delta, box_size, s_edges, guard_mu=True, **kw)
    return k3d, pk, nm, r3d, xi3d


def compute_all_correlations(delta, box_size, s_edges, k_edges, k1, k2, theta, **kw):
    """correlations.py:464-637: 11 outputs."""
    k3d, pk, nm = powspec(delta, box_size, k_edges, **kw)
    r3d, xi3d, nmx = xi(delta, box_size, s_edges, guard_mu=True, **kw)
    k_all, pkb, th, B, Q = bispec(delta, box_size, k1, k2, theta, **kw)
    return k3d, pk, nm, r3d, xi3d, nmx, k_all, pkb, th, B, Q
