"""Differentiable versions of the drop-in functions.

The reference is built to be differentiated: its scripts run ``jax.value_and_grad`` through
``cic_mas_vec`` -> ``delta/mean - 1`` -> ``powspec_vec`` (e.g. /root/reference/tests/lognormal.py:99-107,
tests/bias.py:36-66).  Here the same composition works under ``torch.autograd`` on CUDA tensors:

    from jax_powspec_b200.autograd import cic_mas_vec, powspec_vec
    rho = cic_mas_vec(zeros, x, y, z, w, n_part, 0., 0., 0., box, n, True)     # x, y, z, w may require grad
    k, pk, nm = powspec_vec(rho / rho.mean() - 1, box, k_edges)
    loss(pk).backward()

``xi_vec``, ``bispec`` and the composites ``compute_2pt_correlations`` / ``compute_all_correlations``
are differentiable too (the loss of /root/reference/tests/lognormal_bispec.py:71-106 mixes P, xi and B).

Backward passes are hand-written kernels behind the C ABI (``jps_paint_grad``: stencil gather with
B-spline weights and their derivatives; ``jps_powspec_grad``: per-mode cotangent + one C2R;
``jps_xi_grad``: cotangent field on the lag grid -> R2C -> per-mode factor -> C2R; ``jps_bispec_grad``:
3 inverse + 3 forward FFTs + 1 inverse whatever the number of triangle bins).  As under JAX, the
integer cell / bin choice has zero derivative; NaN outputs of empty bins contribute nothing.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, lib
from .correlations import _edge_ptr, _host_edges, _theta_arg
from .correlations import bispec as _bispec
from .correlations import powspec_vec as _powspec_vec
from .correlations import xi_vec as _xi_vec
from .mas import paint as _paint
from .plan import get_plan, ptr, require_cuda, stream_ptr

__all__ = ["paint", "cic_mas_vec", "tsc_mas_vec", "pcs_mas_vec", "powspec_vec", "xi_vec", "bispec",
           "compute_2pt_correlations", "compute_all_correlations"]


class _PaintFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mesh_in, x, y, z, w, cfg):
        out = _paint(mesh_in.detach(), x.detach(), y.detach(), z.detach(), None if w is None else w.detach(),
                     cfg["xmin"], cfg["ymin"], cfg["zmin"], cfg["box_size"], cfg["n_bins"], cfg["wrap"],
                     order=cfg["order"], compat=cfg["compat"], variant=cfg["variant"], method=cfg["method"])
        ctx.cfg = cfg
        ctx.has_w = w is not None
        ctx.save_for_backward(x, y, z, w if w is not None else x.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, g):
        x, y, z, w = ctx.saved_tensors
        cfg = ctx.cfg
        g = g.contiguous().to(torch.float32)
        xs, ys, zs = (t.detach().contiguous().to(torch.float32) for t in (x, y, z))
        ws = w.detach().contiguous().to(torch.float32) if ctx.has_w else None
        n = xs.numel()
        need = ctx.needs_input_grad
        gx = torch.empty(n, dtype=torch.float32, device=g.device) if need[1] else None
        gy = torch.empty(n, dtype=torch.float32, device=g.device) if need[2] else None
        gz = torch.empty(n, dtype=torch.float32, device=g.device) if need[3] else None
        gw = torch.empty(n, dtype=torch.float32, device=g.device) if (ctx.has_w and need[4]) else None
        check(lib.jps_paint_grad(int(cfg["n_bins"]), ptr(xs), ptr(ys), ptr(zs), ptr(ws), 1, n,
                                 float(cfg["xmin"]), float(cfg["ymin"]), float(cfg["zmin"]), float(cfg["box_size"]),
                                 int(cfg["order"]), int(bool(cfg["wrap"])), _lib.COMPAT[cfg["compat"]],
                                 _lib.VARIANT_SCAN if cfg["variant"] == "scan" else _lib.VARIANT_VEC,
                                 ptr(g), ptr(gx), ptr(gy), ptr(gz), ptr(gw), stream_ptr()), "jps_paint_grad")
        return (g if need[0] else None), gx, gy, gz, gw, None


def paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap=True, *, order=2,
          compat="reference", variant="vec", method="auto"):
    """Differentiable ``mas.paint`` for CUDA torch tensors (gradients w.r.t. delta, x, y, z, w)."""
    require_cuda()
    cfg = dict(xmin=float(xmin), ymin=float(ymin), zmin=float(zmin), box_size=float(box_size), n_bins=int(n_bins),
               wrap=bool(wrap), order=int(order), compat=compat, variant=variant, method=method)
    return _PaintFn.apply(delta, x, y, z, w, cfg)


def cic_mas_vec(delta, x, y, z, w, n_part, xmin, ymin, zmin, box_size, n_bins, wrap, **kw):
    return paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap, order=2, **kw)


def tsc_mas_vec(delta, x, y, z, w, n_part, xmin, ymin, zmin, box_size, n_bins, wrap, **kw):
    kw.setdefault("compat", "fixed")
    return paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap, order=3, **kw)


def pcs_mas_vec(delta, x, y, z, w, n_part, xmin, ymin, zmin, box_size, n_bins, wrap, **kw):
    kw.setdefault("compat", "fixed")
    return paint(delta, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap, order=4, **kw)


class _PowspecFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, delta, cfg):
        k3d, pk, nm = _powspec_vec(delta.detach(), cfg["box_size"], cfg["edges"], mas_order=cfg["mas_order"],
                                   shot_noise=cfg["shot_noise"], normalise=cfg["normalise"])
        ctx.cfg = cfg
        ctx.save_for_backward(delta)
        ctx.mark_non_differentiable(k3d, nm)
        return k3d, pk, nm

    @staticmethod
    def backward(ctx, _gk, gpk, _gnm):
        (delta,) = ctx.saved_tensors
        cfg = ctx.cfg
        mesh = delta.detach().contiguous().to(torch.float32)
        n = mesh.shape[0]
        e = cfg["edges"]
        gpk = torch.nan_to_num(gpk.contiguous().to(torch.float32), nan=0.0)      # empty bins carry NaN outputs
        plan = get_plan(n, mesh.device)
        gmesh = torch.empty_like(mesh)
        check(lib.jps_powspec_grad(plan.handle, ptr(mesh), int(bool(cfg["normalise"])), float(cfg["box_size"]),
                                   _edge_ptr(e), e.size - 1, int(cfg["mas_order"]), ptr(gpk), ptr(gmesh),
                                   stream_ptr()), "jps_powspec_grad")
        return gmesh, None


def powspec_vec(delta, box_size, k_edges, *, mas_order=2, shot_noise=0.0, normalise=False):
    """Differentiable ``correlations.powspec_vec`` for a CUDA torch mesh (gradient w.r.t. delta)."""
    require_cuda()
    cfg = dict(box_size=float(box_size), edges=_host_edges(k_edges), mas_order=int(mas_order),
               shot_noise=float(shot_noise), normalise=bool(normalise))
    return _PowspecFn.apply(delta, cfg)


class _XiFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, delta, cfg):
        r3d, xi, nm = _xi_vec(delta.detach(), cfg["box_size"], cfg["edges"], mas_order=cfg["mas_order"],
                              guard_mu=cfg["guard_mu"])
        ctx.cfg = cfg
        ctx.save_for_backward(delta)
        ctx.mark_non_differentiable(r3d, nm)
        return r3d, xi, nm

    @staticmethod
    def backward(ctx, _gr, gxi, _gnm):
        (delta,) = ctx.saved_tensors
        cfg = ctx.cfg
        mesh = delta.detach().contiguous().to(torch.float32)
        e = cfg["edges"]
        gxi = torch.nan_to_num(gxi.contiguous().to(torch.float32), nan=0.0, posinf=0.0, neginf=0.0)
        plan = get_plan(mesh.shape[0], mesh.device, n_shell_fields=1)
        gmesh = torch.empty_like(mesh)
        check(lib.jps_xi_grad(plan.handle, ptr(mesh), float(cfg["box_size"]), _edge_ptr(e), e.size - 1,
                              int(cfg["mas_order"]), ptr(gxi), ptr(gmesh), stream_ptr()), "jps_xi_grad")
        return gmesh, None


def xi_vec(delta, box_size, s_edges, *, mas_order=2, guard_mu=False):
    """Differentiable ``correlations.xi_vec`` for a CUDA torch mesh (gradient w.r.t. delta)."""
    require_cuda()
    cfg = dict(box_size=float(box_size), edges=_host_edges(s_edges), mas_order=int(mas_order), guard_mu=bool(guard_mu))
    return _XiFn.apply(delta, cfg)


class _BispecFn(torch.autograd.Function):
    """Outputs (k_all, Pk, theta, B); Q is formed from Pk and B with torch ops by the caller so that
    autograd folds its cotangent into those of Pk and B."""

    @staticmethod
    def forward(ctx, delta, cfg):
        k_all, pk, th, B, _Q = _bispec(delta.detach(), cfg["box_size"], cfg["k1"], cfg["k2"], cfg["theta"],
                                       mas_order=cfg["mas_order"])
        ctx.cfg = cfg
        ctx.save_for_backward(delta)
        ctx.mark_non_differentiable(k_all, th)
        return k_all, pk, th, B

    @staticmethod
    def backward(ctx, _gk, gpk, _gth, gB):
        (delta,) = ctx.saved_tensors
        cfg = ctx.cfg
        mesh = delta.detach().contiguous().to(torch.float32)
        t = cfg["theta"]
        gpk = torch.nan_to_num(gpk.contiguous().to(torch.float32), nan=0.0, posinf=0.0, neginf=0.0)
        gB = torch.nan_to_num(gB.contiguous().to(torch.float32), nan=0.0, posinf=0.0, neginf=0.0)
        plan = get_plan(mesh.shape[0], mesh.device, n_shell_fields=6)
        gmesh = torch.empty_like(mesh)
        check(lib.jps_bispec_grad(plan.handle, ptr(mesh), float(cfg["box_size"]), float(cfg["k1"]), float(cfg["k2"]),
                                  _edge_ptr(t), t.size, int(cfg["mas_order"]), ptr(gpk), ptr(gB), ptr(gmesh),
                                  stream_ptr()), "jps_bispec_grad")
        return gmesh, None


def bispec(delta, box_size, k1, k2, theta, *, mas_order=2):
    """Differentiable ``correlations.bispec``: ``(k_all, Pk, theta, B, Q)`` with gradients of Pk, B and Q
    w.r.t. delta."""
    require_cuda()
    cfg = dict(box_size=float(box_size), k1=float(k1), k2=float(k2), theta=_theta_arg(theta), mas_order=int(mas_order))
    k_all, pk, th, B = _BispecFn.apply(delta, cfg)
    p0, p1, p3 = pk[0], pk[1], pk[2:]
    Q = B / (p0 * p1 + p0 * p3 + p1 * p3)                  # /root/reference/src/correlations.py:452
    return k_all, pk, th, B, Q


def compute_2pt_correlations(delta, box_size, s_edges, k_edges, *, mas_order=2):
    """Differentiable ``(k3D, Pk3D, Nmodes3D_pk, r3D, xi3D)`` (/root/reference/src/correlations.py:641)."""
    k3d, pk, nmk = powspec_vec(delta, box_size, k_edges, mas_order=mas_order)
    r3d, xi, _ = xi_vec(delta, box_size, s_edges, mas_order=mas_order, guard_mu=True)
    return k3d, pk, nmk, r3d, xi


def compute_all_correlations(delta, box_size, s_edges, k_edges, k1, k2, theta, *, mas_order=2):
    """Differentiable 11-tuple of /root/reference/src/correlations.py:465.  (The forward pass of the
    non-differentiable ``correlations.compute_all_correlations`` shares one FFT between the three
    estimators; here each estimator keeps its own autograd node.)"""
    k3d, pk, nmk = powspec_vec(delta, box_size, k_edges, mas_order=mas_order)
    r3d, xi, nmx = xi_vec(delta, box_size, s_edges, mas_order=mas_order, guard_mu=True)
    k_all, pks, th, B, Q = bispec(delta, box_size, k1, k2, theta, mas_order=mas_order)
    return k3d, pk, nmk, r3d, xi, nmx, k_all, pks, th, B, Q
