"""Gradients of paint and P(k) (CUDA kernels behind the C ABI) against torch autograd through a plain
float64 PyTorch restatement of the same arithmetic (the "reference" for a floating-point kernel)."""
import math

import numpy as np
import pytest
import torch

from tests.util import F32, clustered_particles

pytestmark = pytest.mark.gpu


def _weights_torch(pos, order, n):
    """(indices [s,Np] long, weights [s,Np]) of the order-2/3/4 B-spline, float64, differentiable in pos."""
    if order == 2:
        i0 = torch.floor(pos).detach(); d = pos - i0
        w = [1 - d, d]; base = i0.long()
    elif order == 3:
        j0 = torch.floor(pos + 0.5).detach(); d = pos - j0
        w = [0.5 * (0.5 - d) ** 2, 0.75 - d * d, 0.5 * (0.5 + d) ** 2]; base = j0.long() - 1
    else:
        i0 = torch.floor(pos).detach(); d = pos - i0; e = 1 - d
        w = [e ** 3 / 6, (4 - 6 * d * d + 3 * d ** 3) / 6, (4 - 6 * e * e + 3 * e ** 3) / 6, d ** 3 / 6]; base = i0.long() - 1
    idx = torch.stack([(base + s) % n for s in range(order)])
    return idx, torch.stack(w)


def _paint_torch(x, y, z, w, box, n, order):
    inv = n / box
    (ix, wx), (iy, wy), (iz, wz) = (_weights_torch(t * inv, order, n) for t in (x, y, z))
    mesh = torch.zeros(n * n * n, dtype=torch.float64, device=x.device)
    for a in range(order):
        for b in range(order):
            for c in range(order):
                mesh = mesh.index_add(0, (ix[a] * n + iy[b]) * n + iz[c], wx[a] * wy[b] * wz[c] * w)
    return mesh.view(n, n, n)


def _powspec_torch(delta, box, edges_grid, mas_order):
    """Same estimator as the reference in float64 torch ops (half-space, every stored mode once)."""
    n = delta.shape[0]
    dk = torch.fft.rfftn(delta)
    ki = torch.fft.fftfreq(n, d=1.0 / n, device=delta.device)
    ki = torch.where(torch.arange(n, device=delta.device) == n // 2, torch.tensor(n / 2.0, device=delta.device, dtype=ki.dtype), ki) \
        if n % 2 == 0 else ki
    kz = torch.arange(n // 2 + 1, device=delta.device, dtype=torch.float64)
    def corr(k):
        xx = math.pi * k / n
        s = torch.where(k == 0, torch.ones_like(xx), torch.sin(xx) / torch.where(k == 0, torch.ones_like(xx), xx))
        return (1.0 / s) ** mas_order
    c = corr(ki)[:, None, None] * corr(ki)[None, :, None] * corr(kz)[None, None, :]
    d2 = (dk.real ** 2 + dk.imag ** 2) * c * c
    k2 = ki[:, None, None] ** 2 + ki[None, :, None] ** 2 + kz[None, None, :] ** 2
    k = torch.sqrt(k2)
    mu2 = torch.where(k2 == 0, torch.zeros_like(k2), kz[None, None, :] ** 2 / torch.where(k2 == 0, torch.ones_like(k2), k2))
    kf32 = torch.sqrt(k2.to(torch.float32))                        # bin decisions in float32 like the reference
    e = torch.as_tensor(edges_grid, device=delta.device)
    bins = torch.bucketize(kf32, e, right=True) - 1
    bins = torch.where(kf32 == e[-1], torch.full_like(bins, len(e) - 2), bins)
    ok = (bins >= 0) & (bins < len(e) - 1)
    nb = len(e) - 1
    out = []
    cnt = torch.zeros(nb, dtype=torch.float64, device=delta.device).index_add(0, bins[ok], torch.ones_like(d2[ok]))
    for leg, mult in ((torch.ones_like(mu2), 1.0), ((3 * mu2 - 1) / 2, 5.0), ((35 * mu2 ** 2 - 30 * mu2 + 3) / 8, 9.0)):
        s = torch.zeros(nb, dtype=torch.float64, device=delta.device).index_add(0, bins[ok], (d2 * leg)[ok])
        out.append(s / cnt * mult * (box / n ** 2) ** 3)
    return torch.stack(out, dim=1), cnt


@pytest.mark.parametrize("order,compat", [(2, "fixed"), (3, "fixed"), (4, "fixed")])
def test_paint_gradients_match_torch_autograd(order, compat):
    from jax_powspec_b200 import autograd as ja
    n, box, npart = 24, 240.0, 4000
    p = clustered_particles(1, npart, box).astype(np.float64)
    rng = np.random.default_rng(2)
    w0 = rng.uniform(0.5, 1.5, npart)
    gout = torch.from_numpy(rng.standard_normal((n, n, n))).cuda()
    xr, yr, zr, wr = (torch.tensor(a, device="cuda", dtype=torch.float64, requires_grad=True) for a in (p[:, 0], p[:, 1], p[:, 2], w0))
    (_paint_torch(xr, yr, zr, wr, box, n, order) * gout).sum().backward()
    xs, ys, zs, ws = (torch.tensor(a, device="cuda", dtype=torch.float32, requires_grad=True) for a in (p[:, 0], p[:, 1], p[:, 2], w0))
    mesh = ja.paint(torch.zeros((n, n, n), device="cuda"), xs, ys, zs, ws, 0., 0., 0., box, n, True, order=order, compat=compat)
    (mesh * gout.float()).sum().backward()
    for got, want, name in ((xs.grad, xr.grad, "x"), (ys.grad, yr.grad, "y"), (zs.grad, zr.grad, "z"), (ws.grad, wr.grad, "w")):
        scale = want.abs().max().item()
        err = (got.double() - want).abs().max().item() / scale
        assert err < 2e-5, f"d/d{name}: {err:.2e}"


def test_reference_cic_gradient_uses_the_quirk_weights():
    """compat='reference': the gradient is that of the weights the reference really uses (Q1)."""
    from jax_powspec_b200 import autograd as ja
    n, box = 8, 8.0
    x = torch.tensor([1.25], device="cuda", requires_grad=True); y = torch.tensor([1.125], device="cuda", requires_grad=True)
    z = torch.tensor([1.75], device="cuda", requires_grad=True); w = torch.tensor([2.0], device="cuda", requires_grad=True)
    mesh = ja.cic_mas_vec(torch.zeros((n, n, n), device="cuda"), x, y, z, w, 1, 0., 0., 0., box, n, True)
    mesh.sum().backward()
    dx, dy, dz = 0.25, 0.125, 0.75
    # total mass = w * (1 + mdx*ddz*(mdy-ddy)): d/dw, d/dx, d/dy, d/dz analytically
    mdx, mdy = 1 - dx, 1 - dy
    assert abs(w.grad.item() - (1 + mdx * dz * (mdy - dy))) < 1e-6
    assert abs(x.grad.item() - 2.0 * (-dz * (mdy - dy))) < 1e-5
    assert abs(y.grad.item() - 2.0 * (mdx * dz * (-2.0))) < 1e-5
    assert abs(z.grad.item() - 2.0 * (mdx * (mdy - dy))) < 1e-5


@pytest.mark.parametrize("normalise", [False, True])
@pytest.mark.parametrize("mas_order", [2, 3])
def test_powspec_gradient_matches_torch_autograd(normalise, mas_order):
    from jax_powspec_b200 import autograd as ja
    from oracle import correlations as oc
    n, box = 16, 100.0
    rng = np.random.default_rng(4)
    field = (1.0 + 0.3 * rng.standard_normal((n, n, n))) if normalise else 0.3 * rng.standard_normal((n, n, n))
    ke = np.arange(0.05, 0.52, 0.06).astype(F32)
    gpk = rng.standard_normal((len(ke) - 1, 3))
    fr = torch.tensor(field, device="cuda", dtype=torch.float64, requires_grad=True)
    d = fr / fr.mean() - 1 if normalise else fr
    pk_ref, cnt = _powspec_torch(d, box, oc.grid_edges(ke, box), mas_order)
    ok = cnt > 0
    (pk_ref[ok] * torch.tensor(gpk, device="cuda")[ok]).sum().backward()
    fs = torch.tensor(field, device="cuda", dtype=torch.float32, requires_grad=True)
    k3d, pk, nm = ja.powspec_vec(fs, box, ke, mas_order=mas_order, normalise=normalise)
    assert torch.allclose(pk[ok].double(), pk_ref[ok].detach(), rtol=2e-4)
    (pk[ok] * torch.tensor(gpk, device="cuda", dtype=torch.float32)[ok]).sum().backward()
    scale = fr.grad.abs().max().item()
    err = (fs.grad.double() - fr.grad).abs().max().item() / scale
    assert err < 5e-5, f"max gradient error {err:.2e} of the largest component"


def test_end_to_end_chain_like_the_reference_scripts():
    """tests/lognormal.py:61-107 pattern: loss(P(k) of painted particles) differentiated w.r.t. positions."""
    from jax_powspec_b200 import autograd as ja
    n, box, npart = 16, 160.0, 3000
    p = clustered_particles(6, npart, box).astype(np.float64)
    ke = np.arange(0.06, 0.3, 0.05).astype(F32)
    from oracle import correlations as oc
    def loss_ref(x, y, z):
        rho = _paint_torch(x, y, z, torch.ones_like(x), box, n, 2)
        pk, cnt = _powspec_torch(rho / rho.mean() - 1, box, oc.grid_edges(ke, box), 2)
        return torch.log(pk[:, 0]).sum()
    xr, yr, zr = (torch.tensor(a, device="cuda", dtype=torch.float64, requires_grad=True) for a in (p[:, 0], p[:, 1], p[:, 2]))
    loss_ref(xr, yr, zr).backward()
    xs, ys, zs = (torch.tensor(a, device="cuda", dtype=torch.float32, requires_grad=True) for a in (p[:, 0], p[:, 1], p[:, 2]))
    rho = ja.cic_mas_vec(torch.zeros((n, n, n), device="cuda"), xs, ys, zs, torch.ones(npart, device="cuda"), npart,
                         0., 0., 0., box, n, True, compat="fixed")
    k3d, pk, nm = ja.powspec_vec(rho / rho.mean() - 1, box, ke)
    torch.log(pk[:, 0]).sum().backward()
    for got, want in ((xs.grad, xr.grad), (ys.grad, yr.grad), (zs.grad, zr.grad)):
        err = (got.double() - want).abs().max().item() / want.abs().max().item()
        assert err < 5e-4, f"{err:.2e}"
