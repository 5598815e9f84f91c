"""Seeded synthetic catalogues for benchmarks and tests (torch ops on the chosen device).

Input generation only -- nothing here is on the measured path.  Mirrors what the reference's
mock pipeline produces (/root/reference/tests/create_lognormal.py:44-55: Gaussian field ->
exp(b*delta) -> Poisson sample -> in-cell offsets), with a plane-parallel redshift-space shift
along z (BASELINE.json configs[1]: "lognormal mock ... in redshift space").
"""
from __future__ import annotations

import math

import torch


def _power_spectrum(k, amp=2.0e4, k0=0.02, ns=0.96):
    """Smooth LCDM-like shape: ~k^ns at low k, turnover at k0, ~k^-2.5 tail.  (h/Mpc)^-3."""
    x = k / k0
    return amp * x.pow(ns) / (1.0 + x * x).pow((ns + 2.5) / 2.0)


@torch.no_grad()
def lognormal_catalog(n_part, box_size, *, n_grid=256, seed=5, bias=1.5, growth_rate=0.8,
                      sigma_g=0.7, device="cuda", shuffle=True):
    """Exactly ``n_part`` particles (x, y, z float32 tensors in [0, box)) drawn from a lognormal
    density field on an n_grid^3 lattice (Gaussian rms ``sigma_g`` per lattice cell, lognormal
    bias ``bias``), displaced along z by the linear velocity field."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    n = int(n_grid)
    cell = box_size / n
    white = torch.randn((n, n, n), generator=gen, device=dev, dtype=torch.float32)
    wk = torch.fft.rfftn(white)
    kf = 2.0 * math.pi / box_size
    k1 = torch.fft.fftfreq(n, d=1.0 / n, device=dev) * kf
    kzv = torch.fft.rfftfreq(n, d=1.0 / n, device=dev) * kf
    kx, ky, kz = k1[:, None, None], k1[None, :, None], kzv[None, None, :]
    k2 = kx * kx + ky * ky + kz * kz
    k2[0, 0, 0] = 1.0
    amp = torch.sqrt(_power_spectrum(torch.sqrt(k2)) / cell ** 3)
    amp[0, 0, 0] = 0.0
    dk = wk * amp
    del wk, white
    delta_g = torch.fft.irfftn(dk, s=(n, n, n))
    # linear displacement along the line of sight: psi_z(k) = i kz / k^2 * delta(k)
    psi_z = torch.fft.irfftn(dk * (1j * kz / k2), s=(n, n, n))
    del dk
    # fix the rms of the Gaussian field on the lattice (sets how heavy the lognormal tail is)
    scale = sigma_g / delta_g.std()
    delta_g = delta_g * scale
    psi_z = psi_z * scale
    sigma2 = delta_g.var()
    rho = torch.exp(bias * delta_g - 0.5 * bias * bias * sigma2)
    lam = rho * (n_part / rho.sum())
    counts = torch.poisson(lam, generator=gen).to(torch.int64)
    # hit n_part exactly: add / remove the difference in randomly chosen occupied cells
    total = int(counts.sum().item())
    flat = counts.view(-1)
    if total != n_part:
        diff = n_part - total
        occ = torch.nonzero(flat > (0 if diff > 0 else 1)).view(-1)
        pick = occ[torch.randint(0, occ.numel(), (abs(diff),), generator=gen, device=dev)]
        flat.index_add_(0, pick, torch.full_like(pick, 1 if diff > 0 else -1))
        flat.clamp_(min=0)
    cells = torch.repeat_interleave(torch.arange(flat.numel(), device=dev), flat)
    if cells.numel() > n_part:
        cells = cells[:n_part]
    elif cells.numel() < n_part:
        cells = torch.cat([cells, cells[: n_part - cells.numel()]])
    if shuffle:
        cells = cells[torch.randperm(cells.numel(), generator=gen, device=dev)]
    iz = cells % n
    iy = (cells // n) % n
    ix = cells // (n * n)
    u = torch.rand((3, cells.numel()), generator=gen, device=dev, dtype=torch.float32)
    x = (ix.to(torch.float32) + u[0]) * cell
    y = (iy.to(torch.float32) + u[1]) * cell
    z = (iz.to(torch.float32) + u[2]) * cell + growth_rate * psi_z.view(-1)[cells]
    z = torch.remainder(z, box_size)
    box32 = torch.tensor(box_size, dtype=torch.float32, device=dev)
    for t in (x, y, z):
        t[t >= box32] = 0.0            # float32 rounding can land exactly on the upper edge
    return x.contiguous(), y.contiguous(), z.contiguous()


@torch.no_grad()
def uniform_catalog(n_part, box_size, *, seed=42, device="cuda"):
    """U[0, box)^3 (BASELINE.json configs[3])."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    p = torch.rand((3, n_part), generator=gen, device=dev, dtype=torch.float32) * box_size
    p[p >= box_size] = 0.0
    return p[0].contiguous(), p[1].contiguous(), p[2].contiguous()
