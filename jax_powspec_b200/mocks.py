"""Mock catalogues on the device.

Row f-3 of SURVEY.md section 8 -- the reference's input pipeline behind its own names:

    gaussian_field(grid, kf, Pkf, Rayleigh_sampling, seed, BoxSize)   /root/reference/src/gauss_field.py:5
    populate_field(rho, n_bins, box_size, density, seed)              /root/reference/src/populate_field.py:11
    lognormal_mock(...)                                               /root/reference/tests/create_lognormal.py:44-55

run as CUDA kernels (csrc/mockgen.cu; jps_mock_* in include/jps.h) with a counter-based Philox stream:
the arithmetic is the reference's (bit for bit for gaussian_field when both are fed the same
uniforms, tests/golden/ref_mock.npz), the random draws are not -- NumPy's sequential Mersenne Twister
and jax.random cannot be followed by a parallel generator -- so mocks agree in distribution.

``lognormal_catalog`` / ``uniform_catalog`` below are the older seeded input synthesisers of the
benchmarks and tests (torch ops; nothing there is on a measured path): exactly n_part particles, a
plane-parallel redshift-space shift along z (BASELINE.json configs[1]), optional shuffling.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from ._lib import check, lib
from .plan import ArrayKind, ptr, require_cuda, stream_ptr, to_device_f32

__all__ = ["gaussian_field", "populate_field", "lognormal_mock", "lognormal_catalog", "uniform_catalog"]


def _workspace(nbytes, device):
    """256-byte aligned scratch of at least nbytes: (tensor that owns it, aligned address)."""
    t = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
    return t, (t.data_ptr() + 255) // 256 * 256


def _seed_arg(seed):
    """Python int, NumPy integer, or a 2-word uint32 key as jax.random.PRNGKey produces."""
    if isinstance(seed, torch.Tensor):
        seed = seed.detach().cpu().numpy()
    a = np.asarray(seed)
    if a.ndim == 0:
        return int(a) & 0xFFFFFFFFFFFFFFFF
    a = a.astype(np.uint64).ravel()
    if a.size != 2:
        raise ValueError("seed must be an integer or a 2-word key")
    return (int(a[0]) << 32 | int(a[1])) & 0xFFFFFFFFFFFFFFFF


def gaussian_field(grid, kf, Pkf, Rayleigh_sampling, seed, BoxSize, *, device=None):
    """delta_k of a Gaussian random field: complex64 CUDA tensor ``[grid, grid, grid//2+1]`` with
    ``<|delta_k|^2> = P(|k|) (grid^2/BoxSize)^3``, like /root/reference/src/gauss_field.py:5-80
    (same signature; P(k) interpolated linearly in the table ``(kf, Pkf)``; ``Rayleigh_sampling=0``
    fixes the amplitudes).  ``irfftn`` of the result is the real-space field."""
    cur = require_cuda()
    dev = cur if device is None else torch.device(device)
    n = int(grid)
    k = np.ascontiguousarray(np.asarray(kf, dtype=np.float64).ravel())
    p = np.ascontiguousarray(np.asarray(Pkf, dtype=np.float64).ravel())
    if k.size != p.size or k.size < 2:
        raise ValueError("kf and Pkf must be equally long 1-d tables with at least two points")
    with torch.cuda.device(dev):
        out = torch.empty((n, n, n // 2 + 1), dtype=torch.complex64, device=dev)
        nbytes = lib.jps_mock_field_workspace_bytes(k.size)
        ws, addr = _workspace(nbytes, dev)
        dp = C.POINTER(C.c_double)
        check(lib.jps_mock_gaussian_field(n, k.ctypes.data_as(dp), p.ctypes.data_as(dp), k.size,
                                          int(bool(Rayleigh_sampling)), _seed_arg(seed), float(BoxSize),
                                          ptr(out), C.c_void_p(addr), nbytes, stream_ptr()),
              "jps_mock_gaussian_field")
    return out


def populate_field(rho, n_bins, box_size, density, seed, *, lognormal_bias=None, return_counts=False):
    """Poisson-sample a density mesh into particles: ``(Np, 3)`` float32 positions in ``[0, box)``,
    like /root/reference/src/populate_field.py:11-29 (``seed``: integer or 2-word PRNG key).  The mean
    count of a cell is ``rho * (box/n)^3 * density / mean(rho)``; every particle sits at its cell centre
    plus a triangular offset of up to one cell per axis.  ``rho`` is NOT rescaled in place (the reference's
    ``rho *= ...`` is a rebinding under jit).  With ``lognormal_bias=b`` the mesh is read as a Gaussian
    field g and ``exp(b g)`` is sampled (tests/create_lognormal.py:49-50) without materialising it.
    Particles come back grouped by cell in C order; NumPy in -> NumPy out, CUDA tensor in -> CUDA out."""
    device = require_cuda()
    kind = ArrayKind(rho)
    mesh = to_device_f32(rho, device)
    n = int(n_bins)
    if mesh.dim() != 3 or tuple(mesh.shape) != (n, n, n):
        raise ValueError("rho must be an (n_bins, n_bins, n_bins) mesh")
    nbytes = lib.jps_mock_populate_workspace_bytes(n)
    ws, addr = _workspace(nbytes, device)
    total = torch.zeros(1, dtype=torch.int64, device=device)
    s = _seed_arg(seed)
    logn = lognormal_bias is not None
    check(lib.jps_mock_populate_count(ptr(mesh), n, float(box_size), float(density), int(logn),
                                      float(lognormal_bias or 0.0), s, C.c_void_p(addr), nbytes, ptr(total),
                                      stream_ptr()), "jps_mock_populate_count")
    n_out = int(total.item())                           # the one host read-back: the output size is data dependent
    pos = torch.empty((n_out, 3), dtype=torch.float32, device=device)
    check(lib.jps_mock_populate_fill(n, float(box_size), s, C.c_void_p(addr), nbytes, n_out, ptr(pos),
                                     stream_ptr()), "jps_mock_populate_fill")
    if return_counts:
        off = addr - ws.data_ptr() + lib.jps_mock_populate_counts_offset(n)
        counts = ws[off: off + n ** 3 * 4].view(torch.int32).view(n, n, n).clone()
        return kind.out(pos), kind.out(counts)
    return kind.out(pos)


def lognormal_mock(grid, kf, Pkf, bias, density, seed, BoxSize, *, Rayleigh_sampling=0, rsd_growth_rate=None):
    """The reference's mock recipe (/root/reference/tests/create_lognormal.py:44-55) end to end on the
    device: Gaussian field of the linear P(k) -> real space (cuFFT through torch.fft.irfftn) ->
    Poisson sample of ``exp(bias * g)`` -> ``(Np, 3)`` float32 CUDA tensor.  ``seed`` feeds the field,
    ``seed + 1`` the sampling (as the script's ``PRNGKey(seed + 1)``).  ``rsd_growth_rate=f`` adds the
    plane-parallel linear redshift-space shift ``f * psi_z`` along z (BASELINE.json configs[1])."""
    n = int(grid)
    dk = gaussian_field(n, kf, Pkf, Rayleigh_sampling, seed, BoxSize)
    g = torch.fft.irfftn(dk, s=(n, n, n)).contiguous()
    pos = populate_field(g, n, BoxSize, density, _seed_arg(seed) + 1, lognormal_bias=float(bias))
    if rsd_growth_rate:
        dev = dk.device
        kf1 = 2.0 * math.pi / float(BoxSize)
        k1 = torch.fft.fftfreq(n, d=1.0 / n, device=dev) * kf1
        kzv = torch.fft.rfftfreq(n, d=1.0 / n, device=dev) * kf1
        k2 = k1[:, None, None] ** 2 + k1[None, :, None] ** 2 + kzv[None, None, :] ** 2
        k2[0, 0, 0] = 1.0
        psi_z = torch.fft.irfftn(dk * (1j * kzv[None, None, :] / k2), s=(n, n, n))
        cell = torch.clamp((pos * (n / float(BoxSize))).long(), 0, n - 1)
        shift = psi_z[cell[:, 0], cell[:, 1], cell[:, 2]] * float(rsd_growth_rate)
        z = torch.remainder(pos[:, 2] + shift, float(BoxSize))
        z[z >= float(BoxSize)] = 0.0
        pos[:, 2] = z
    return pos


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
