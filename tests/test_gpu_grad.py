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


# ---------------------------------------------------------------- xi(s) and bispectrum gradients
def _deconvolved_dk_torch(delta, mas_order):
    n = delta.shape[0]
    dk = torch.fft.rfftn(delta)
    ki = torch.fft.fftfreq(n, d=1.0 / n, device=delta.device).to(torch.float64)
    if n % 2 == 0:
        ki = torch.where(torch.arange(n, device=delta.device) == n // 2, torch.full_like(ki, n / 2.0), ki)
    kz = torch.arange(n // 2 + 1, device=delta.device, dtype=torch.float64)

    def corr(k):
        xx = math.pi * k / n
        s = torch.where(k == 0, torch.ones_like(xx), torch.sin(xx) / torch.where(k == 0, torch.ones_like(xx), xx))
        return (1.0 / s) ** mas_order
    c = corr(ki)[:, None, None] * corr(ki)[None, :, None] * corr(kz)[None, None, :]
    k2 = ki[:, None, None] ** 2 + ki[None, :, None] ** 2 + kz[None, None, :] ** 2
    return dk * c, ki, k2


def _bins_f32(val2, edges_grid):
    """searchsorted(edges, sqrt_f32(val2), 'right') - 1 with the last edge inclusive (jnp.histogram)."""
    v = torch.sqrt(val2.to(torch.float32))
    e = torch.as_tensor(edges_grid, device=val2.device)
    bins = torch.bucketize(v, e, right=True) - 1
    bins = torch.where(v == e[-1], torch.full_like(bins, len(e) - 2), bins)
    return bins, (bins >= 0) & (bins < len(e) - 1)


def _xi_torch(delta, box, edges_grid, mas_order):
    """/root/reference/src/correlations.py:120-187 with the composites' mu(r=0)=0 (:527), float64."""
    n = delta.shape[0]
    D, ki, _ = _deconvolved_dk_torch(delta, mas_order)
    X = torch.fft.irfftn((D.real ** 2 + D.imag ** 2).to(torch.complex128), s=(n, n, n))
    r2 = ki[:, None, None] ** 2 + ki[None, :, None] ** 2 + ki[None, None, :] ** 2
    mu2 = torch.where(r2 == 0, torch.zeros_like(r2), (ki[None, None, :] ** 2).expand_as(r2) / torch.where(r2 == 0, torch.ones_like(r2), r2))
    bins, ok = _bins_f32(r2, edges_grid)
    nb = len(edges_grid) - 1
    cnt = torch.zeros(nb, dtype=torch.float64, device=delta.device).index_add(0, bins[ok], torch.ones_like(X[ok]))
    out = []
    for leg, mult in ((torch.ones_like(mu2), 1.0), ((3 * mu2 - 1) / 2, 5.0), ((35 * mu2 ** 2 - 30 * mu2 + 3) / 8, 9.0)):
        s = torch.zeros(nb, dtype=torch.float64, device=delta.device).index_add(0, bins[ok], (X * leg)[ok])
        out.append(s / cnt * mult / n ** 3)
    return torch.stack(out, dim=1), cnt


def _bispec_torch(delta, box, k1, k2, theta, mas_order):
    """/root/reference/src/correlations.py:334-462 in float64 torch ops (shell decisions in float32)."""
    from oracle import correlations as oc
    n = delta.shape[0]
    D, ki, k2g = _deconvolved_dk_torch(delta, mas_order)
    kf = torch.sqrt(k2g.to(torch.float32))
    k_all, lo, hi = oc.bispec_shells(box, k1, k2, theta)

    def fields(j):
        m = ((kf >= float(lo[j])) & (kf < float(hi[j]))).to(torch.float64)
        return torch.fft.irfftn(m * D, s=(n, n, n)), torch.fft.irfftn(m.to(torch.complex128), s=(n, n, n))
    vol_p = (box / n ** 2) ** 3
    vol_b = (box * box / n ** 3) ** 3
    d0, i0 = fields(0)
    d1, i1 = fields(1)
    pk = [(d0 * d0).sum() / (i0 * i0).sum() * vol_p, (d1 * d1).sum() / (i1 * i1).sum() * vol_p]
    B, Q = [], []
    for b in range(len(theta)):
        d3, i3 = fields(b + 2)
        p3 = (d3 * d3).sum() / (i3 * i3).sum() * vol_p
        bb = (d0 * d1 * d3).sum() / (i0 * i1 * i3).sum() * vol_b
        pk.append(p3)
        B.append(bb)
        Q.append(bb / (pk[0] * pk[1] + pk[0] * p3 + pk[1] * p3))
    return torch.stack(pk), torch.stack(B), torch.stack(Q)


def _rel_err(got, want):
    return (got.double() - want).abs().max().item() / want.abs().max().item()


@pytest.mark.parametrize("mas_order", [2, 3])
def test_xi_gradient_matches_torch_autograd(mas_order):
    from jax_powspec_b200 import autograd as ja
    from oracle import correlations as oc
    n, box = 20, 200.0
    rng = np.random.default_rng(8)
    field = 0.4 * rng.standard_normal((n, n, n))
    se = np.arange(0.0, 90.0, 10.0).astype(F32)                    # first edge 0: the r = 0 cell is in bin 0
    gxi = rng.standard_normal((len(se) - 1, 3))
    fr = torch.tensor(field, device="cuda", dtype=torch.float64, requires_grad=True)
    xi_ref, cnt = _xi_torch(fr, box, oc.s_edges_to_grid(se, box, n), mas_order)
    assert (cnt > 0).all()
    (xi_ref * torch.tensor(gxi, device="cuda")).sum().backward()
    fs = torch.tensor(field, device="cuda", dtype=torch.float32, requires_grad=True)
    r3d, xi, nm = ja.xi_vec(fs, box, se, mas_order=mas_order, guard_mu=True)
    assert torch.allclose(xi.double(), xi_ref.detach(), rtol=2e-3, atol=2e-6 * xi_ref.abs().max().item())
    (xi * torch.tensor(gxi, device="cuda", dtype=torch.float32)).sum().backward()
    err = _rel_err(fs.grad, fr.grad)
    assert err < 2e-4, f"max gradient error {err:.2e} of the largest component"


def test_bispec_gradient_matches_torch_autograd():
    from jax_powspec_b200 import autograd as ja
    n, box = 24, 240.0
    rng = np.random.default_rng(9)
    field = 0.4 * rng.standard_normal((n, n, n))
    kF = 2 * math.pi / box
    k1, k2 = 4.2 * kF, 5.6 * kF
    theta = np.linspace(0.2, 2.9, 6).astype(F32)
    gpk, gB, gQ = rng.standard_normal(len(theta) + 2), rng.standard_normal(len(theta)), rng.standard_normal(len(theta))
    fr = torch.tensor(field, device="cuda", dtype=torch.float64, requires_grad=True)
    pk_r, B_r, Q_r = _bispec_torch(fr, box, k1, k2, theta, 2)
    # cotangents scaled so that the three terms contribute comparably
    tg = lambda a, ref: torch.tensor(a, device="cuda") / ref.detach().abs().mean()
    ((pk_r * tg(gpk, pk_r)).sum() + (B_r * tg(gB, B_r)).sum() + (Q_r * tg(gQ, Q_r)).sum()).backward()
    fs = torch.tensor(field, device="cuda", dtype=torch.float32, requires_grad=True)
    k_all, pk, th, B, Q = ja.bispec(fs, box, k1, k2, theta)
    assert torch.allclose(pk.double(), pk_r.detach(), rtol=1e-3)
    assert torch.allclose(B.double(), B_r.detach(), rtol=2e-3, atol=2e-3 * B_r.abs().max().item())
    ((pk * tg(gpk, pk_r).float()).sum() + (B * tg(gB, B_r).float()).sum() + (Q * tg(gQ, Q_r).float()).sum()).backward()
    err = _rel_err(fs.grad, fr.grad)
    assert err < 5e-4, f"max gradient error {err:.2e} of the largest component"
    # second call: the indicator sums come from the cache, the gradient must not change
    fs2 = torch.tensor(field, device="cuda", dtype=torch.float32, requires_grad=True)
    _, pk2, _, B2, Q2 = ja.bispec(fs2, box, k1, k2, theta)
    ((pk2 * tg(gpk, pk_r).float()).sum() + (B2 * tg(gB, B_r).float()).sum() + (Q2 * tg(gQ, Q_r).float()).sum()).backward()
    assert torch.allclose(fs2.grad, fs.grad, rtol=1e-5, atol=1e-6 * fs.grad.abs().max().item())
    # backward on a fresh plan (no cached indicator sums): the library runs the forward pass itself
    from jax_powspec_b200.plan import clear_plans
    fs3 = torch.tensor(field, device="cuda", dtype=torch.float32, requires_grad=True)
    _, pk3, _, B3, Q3 = ja.bispec(fs3, box, k1, k2, theta)
    clear_plans()
    ((pk3 * tg(gpk, pk_r).float()).sum() + (B3 * tg(gB, B_r).float()).sum() + (Q3 * tg(gQ, Q_r).float()).sum()).backward()
    assert torch.allclose(fs3.grad, fs.grad, rtol=1e-4, atol=1e-5 * fs.grad.abs().max().item())


def test_all_correlations_loss_like_lognormal_bispec():
    """/root/reference/tests/lognormal_bispec.py:71-106: one loss on P0, xi0 and B, differentiated
    w.r.t. particle positions through paint -> delta -> compute_all_correlations."""
    from jax_powspec_b200 import autograd as ja
    from oracle import correlations as oc
    n, box, npart = 16, 160.0, 4000
    p = clustered_particles(12, npart, box).astype(np.float64)
    ke = np.arange(0.06, 0.3, 0.05).astype(F32)
    se = np.arange(0.0, 70.0, 10.0).astype(F32)
    kF = 2 * math.pi / box
    k1, k2 = 3.1 * kF, 4.3 * kF
    theta = np.linspace(0.3, 2.8, 4).astype(F32)

    def loss_terms(pk, xi, B, ref_scale=None):
        return (pk[:, 0] ** 2).mean() * 1e-6 + (xi[:, 0] ** 2).mean() * 1e2 + (B ** 2).mean() * 1e-12

    xr, yr, zr = (torch.tensor(a, device="cuda", dtype=torch.float64, requires_grad=True) for a in (p[:, 0], p[:, 1], p[:, 2]))
    rho = _paint_torch(xr, yr, zr, torch.ones_like(xr), box, n, 2)
    d = rho / rho.mean() - 1
    pk_r, _ = _powspec_torch(d, box, oc.grid_edges(ke, box), 2)
    xi_r, _ = _xi_torch(d, box, oc.s_edges_to_grid(se, box, n), 2)
    _, B_r, _ = _bispec_torch(d, box, k1, k2, theta, 2)
    loss_terms(pk_r, xi_r, B_r).backward()
    xs, ys, zs = (torch.tensor(a, device="cuda", dtype=torch.float32, requires_grad=True) for a in (p[:, 0], p[:, 1], p[:, 2]))
    rho = ja.cic_mas_vec(torch.zeros((n, n, n), device="cuda"), xs, ys, zs, torch.ones(npart, device="cuda"), npart,
                         0., 0., 0., box, n, True, compat="fixed")
    out = ja.compute_all_correlations(rho / rho.mean() - 1, box, se, ke, k1, k2, theta)
    assert len(out) == 11
    loss_terms(out[1], out[4], out[9]).backward()
    for got, want in ((xs.grad, xr.grad), (ys.grad, yr.grad), (zs.grad, zr.grad)):
        err = _rel_err(got, want)
        assert err < 2e-3, f"{err:.2e}"
