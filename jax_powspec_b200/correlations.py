"""Fourier-space estimators with the reference's names and positional signatures.

Drop-in for /root/reference/src/correlations.py: ``powspec_vec`` (:7-56),
``powspec_vec_fundamental`` (:60-117), ``xi_vec`` (:120-187), ``xi_vec_fundamental`` (:191-261),
``bispec`` (:334-462), ``compute_all_correlations`` (:464-637), ``compute_2pt_correlations``
(:640-712), and the fused ``paint_powspec`` path the benchmark times (paint -> density contrast
-> rfftn -> multipoles, tests/correlations.py:41-78).
Results come back in the container kind of ``delta`` (NumPy in -> NumPy out, CUDA tensor in ->
CUDA tensors out), float32 as the reference returns them.

All arithmetic runs in libjps.so (cuFFT + hand-written sm_100a kernels); no fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .mas import new_paint_workspace, paint_workspace_bytes
from .plan import ArrayKind, Plan, check_particles, get_plan, ptr, require_cuda, stream_ptr, to_device_f32

__all__ = ["powspec_vec", "powspec_vec_fundamental", "xi_vec", "xi_vec_fundamental", "xi_vec_coords",
           "s_edges_conv", "bispec", "bispec_pairs", "triangle_pairs", "compute_2pt_correlations", "compute_all_correlations",
           "paint_powspec", "PaintPowspec", "HostPipeline"]


def _host_edges(k_edges):
    if isinstance(k_edges, torch.Tensor):
        k_edges = k_edges.detach().cpu().numpy()
    e = np.ascontiguousarray(np.asarray(k_edges, dtype=np.float32))
    if e.ndim != 1 or e.size < 2:
        raise ValueError("k_edges must be a 1-d array with at least two edges")
    return e


def _edge_ptr(e):
    return e.ctypes.data_as(C.POINTER(C.c_float))


def powspec_vec(delta, box_size, k_edges, *, mas_order=2, shot_noise=0.0, normalise=False,
                return_raw=False, mode_weighting="half", delta2=None):
    """P0, P2, P4 in the user's k bins: returns ``(k3D[nb], Pk3D[nb,3], Nmodes3D[nb])`` float32.

    mas_order / shot_noise / normalise extend the reference (which hard-codes the CIC window,
    never subtracts shot noise and expects ``delta`` already normalised).  With
    ``return_raw=True`` a 4th item ``(sums float64[nb,3], counts int64[nb])`` is appended.

    Two estimator options the reference lacks (defaults = the reference's behaviour):
    ``mode_weighting="hermitian"`` counts every stored mode with 0 < kz < N/2 twice, i.e. the
    full-space shell average instead of the reference's half-space one (Q7);
    ``delta2`` = the same particles painted on the grid displaced by half a cell
    (:func:`jax_powspec_b200.mas.paint_interlaced`) gives the interlaced, alias-suppressed spectrum."""
    device = require_cuda()
    kind = ArrayKind(delta)
    mesh = to_device_f32(delta, device)
    n = mesh.shape[0]
    if mesh.dim() != 3 or tuple(mesh.shape) != (n, n, n):
        raise ValueError("delta must be a cubic 3-d mesh")
    if mode_weighting not in ("half", "hermitian"):
        raise ValueError("mode_weighting must be 'half' or 'hermitian'")
    mesh2 = None
    if delta2 is not None:
        mesh2 = to_device_f32(delta2, device)
        if tuple(mesh2.shape) != (n, n, n):
            raise ValueError("delta2 must have the shape of delta")
    e = _host_edges(k_edges)
    nb = e.size - 1
    extended = mesh2 is not None or mode_weighting != "half"
    plan = get_plan(n, device, n_shell_fields=1 if mesh2 is not None else 0)
    k3d = torch.empty(nb, dtype=torch.float32, device=device)
    pk = torch.empty((nb, 3), dtype=torch.float32, device=device)
    nm = torch.empty(nb, dtype=torch.float32, device=device)
    sums = torch.empty((nb, 3), dtype=torch.float64, device=device) if return_raw else None
    counts = torch.empty(nb, dtype=torch.int64, device=device) if return_raw else None
    if extended:
        flags = _lib.PK_HERMITIAN if mode_weighting == "hermitian" else 0
        check(lib.jps_powspec_ex(plan.handle, ptr(mesh), ptr(mesh2), int(bool(normalise)), float(box_size),
                                 _edge_ptr(e), nb, int(mas_order), float(shot_noise), flags, ptr(k3d), ptr(pk),
                                 ptr(nm), ptr(sums), ptr(counts), stream_ptr()), "jps_powspec_ex")
    else:
        check(lib.jps_powspec(plan.handle, ptr(mesh), int(bool(normalise)), float(box_size), _edge_ptr(e), nb,
                              int(mas_order), float(shot_noise), ptr(k3d), ptr(pk), ptr(nm), ptr(sums),
                              ptr(counts), stream_ptr()), "jps_powspec")
    out = (kind.out(k3d), kind.out(pk), kind.out(nm))
    if return_raw:
        out = out + ((kind.out(sums), kind.out(counts)),)
    return out


def powspec_vec_fundamental(delta, box_size, *, mas_order=2, compat="reference", normalise=False,
                            return_raw=False):
    """kF-wide integer bins, bin 0 dropped: ``(k3D, Pk3D[kmax,3], Nmodes3D)``.
    compat='reference' reproduces the reference's k3D (Q18); 'fixed' returns the mean |k|."""
    device = require_cuda()
    kind = ArrayKind(delta)
    mesh = to_device_f32(delta, device)
    n = mesh.shape[0]
    if mesh.dim() != 3 or tuple(mesh.shape) != (n, n, n):
        raise ValueError("delta must be a cubic 3-d mesh")
    nb = lib.jps_fundamental_nbins(n)
    plan = get_plan(n, device)
    k3d = torch.empty(nb, dtype=torch.float32, device=device)
    pk = torch.empty((nb, 3), dtype=torch.float32, device=device)
    nm = torch.empty(nb, dtype=torch.float32, device=device)
    sums = torch.empty((nb, 3), dtype=torch.float64, device=device) if return_raw else None
    counts = torch.empty(nb, dtype=torch.int64, device=device) if return_raw else None
    check(lib.jps_powspec_fundamental(plan.handle, ptr(mesh), int(bool(normalise)), float(box_size),
                                      int(mas_order), _lib.COMPAT[compat], ptr(k3d), ptr(pk), ptr(nm),
                                      ptr(sums), ptr(counts), stream_ptr()), "jps_powspec_fundamental")
    out = (kind.out(k3d), kind.out(pk), kind.out(nm))
    if return_raw:
        out = out + ((kind.out(sums), kind.out(counts)),)
    return out


def _mesh_arg(delta, device):
    mesh = to_device_f32(delta, device)
    n = mesh.shape[0]
    if mesh.dim() != 3 or tuple(mesh.shape) != (n, n, n):
        raise ValueError("delta must be a cubic 3-d mesh")
    return mesh, n


def _f32(device, *shape):
    return torch.empty(shape, dtype=torch.float32, device=device)


def xi_vec(delta, box_size, s_edges, *, mas_order=2, normalise=False, guard_mu=False):
    """xi_0,2,4(s): ``(r3D[nb], xi3D[nb,3], Nmodes3D[nb])`` like
    /root/reference/src/correlations.py:121 (empty bins: Nmodes = inf, xi = 0; with
    ``s_edges[0] == 0`` the first bin of xi2/xi4 is NaN exactly as in the reference, Q22 --
    pass ``guard_mu=True`` for the composites' behaviour)."""
    device = require_cuda()
    kind = ArrayKind(delta)
    mesh, n = _mesh_arg(delta, device)
    e = _host_edges(s_edges)
    nb = e.size - 1
    plan = get_plan(n, device, n_shell_fields=1)
    r3d, xi, nm = _f32(device, nb), _f32(device, nb, 3), _f32(device, nb)
    check(lib.jps_xi(plan.handle, ptr(mesh), int(bool(normalise)), float(box_size), _edge_ptr(e), nb,
                     int(mas_order), int(bool(guard_mu)), ptr(r3d), ptr(xi), ptr(nm), None, None,
                     stream_ptr()), "jps_xi")
    return kind.out(r3d), kind.out(xi), kind.out(nm)


def xi_vec_fundamental(delta, box_size, *, mas_order=2, normalise=False):
    """Integer-lag bins, bin 0 dropped (/root/reference/src/correlations.py:191)."""
    device = require_cuda()
    kind = ArrayKind(delta)
    mesh, n = _mesh_arg(delta, device)
    nb = lib.jps_fundamental_nbins(n)
    plan = get_plan(n, device, n_shell_fields=1)
    r3d, xi, nm = _f32(device, nb), _f32(device, nb, 3), _f32(device, nb)
    check(lib.jps_xi_fundamental(plan.handle, ptr(mesh), int(bool(normalise)), float(box_size),
                                 int(mas_order), ptr(r3d), ptr(xi), ptr(nm), None, None, stream_ptr()),
          "jps_xi_fundamental")
    return kind.out(r3d), kind.out(xi), kind.out(nm)


def xi_vec_coords(dims, box_size, k_edges):
    """/root/reference/src/correlations.py:263 -- bin centres in Mpc/h for grid-unit edges
    (tiny 1-d host arithmetic, float32 like the reference)."""
    kF = np.float32(2.0 * np.pi) / np.float32(box_size)
    kedges = np.asarray(k_edges, dtype=np.float32) / kF
    return np.float32(0.5) * (kedges[1:] + kedges[:-1]) * (np.float32(box_size) * np.float32(1.0) / np.float32(dims))


def s_edges_conv(dims, box_size, s_edges):
    """/root/reference/src/correlations.py:270."""
    kF = np.float32(2.0 * np.pi) / np.float32(box_size)
    return kF * np.asarray(s_edges, dtype=np.float32) * np.float32(dims) / np.float32(box_size)


def _theta_arg(theta):
    if isinstance(theta, torch.Tensor):
        theta = theta.detach().cpu().numpy()
    t = np.ascontiguousarray(np.asarray(theta, dtype=np.float32))
    if t.ndim != 1 or t.size < 1:
        raise ValueError("theta must be a non-empty 1-d array")
    return t


def bispec(delta, box_size, k1, k2, theta, *, mas_order=2, normalise=False):
    """FFT bispectrum: ``(k_all[bins+2], Pk[bins+2], theta, B[bins], Q[bins])`` like
    /root/reference/src/correlations.py:335."""
    device = require_cuda()
    kind = ArrayKind(delta)
    mesh, n = _mesh_arg(delta, device)
    t = _theta_arg(theta)
    nb = t.size
    plan = get_plan(n, device, n_shell_fields=6)
    k_all, pk, B, Q = _f32(device, nb + 2), _f32(device, nb + 2), _f32(device, nb), _f32(device, nb)
    check(lib.jps_bispec(plan.handle, ptr(mesh), int(bool(normalise)), float(box_size), float(k1), float(k2),
                         _edge_ptr(t), nb, int(mas_order), ptr(k_all), ptr(pk), ptr(B), ptr(Q), stream_ptr()),
          "jps_bispec")
    th = torch.from_numpy(t).to(device) if kind.on_device else (torch.from_numpy(t) if kind.is_torch else t)
    return kind.out(k_all), kind.out(pk), th, kind.out(B), kind.out(Q)


def triangle_pairs(k_centres):
    """All (k1 <= k2) pairs of the shell centres, in row-major order of the upper triangle: the
    "all triangle bins up to k_max" sweep of BASELINE.json configs[2]."""
    kc = np.asarray(k_centres, dtype=np.float32).ravel()
    i, j = np.triu_indices(kc.size)
    return kc[i].copy(), kc[j].copy()


def bispec_pairs(delta, box_size, k1, k2, theta, *, mas_order=2, normalise=False):
    """``bispec`` (/root/reference/src/correlations.py:335) for many (k1, k2) pairs with ONE forward
    FFT: returns ``(k_all[np, bins+2], Pk[np, bins+2], theta, B[np, bins], Q[np, bins])``; row p is
    what ``bispec(delta, box_size, k1[p], k2[p], theta)`` returns."""
    device = require_cuda()
    kind = ArrayKind(delta)
    mesh, n = _mesh_arg(delta, device)
    t = _theta_arg(theta)
    a = np.ascontiguousarray(np.asarray(k1, dtype=np.float32).ravel())
    b = np.ascontiguousarray(np.asarray(k2, dtype=np.float32).ravel())
    if a.size != b.size or a.size < 1:
        raise ValueError("k1 and k2 must be equally long, non-empty 1-d arrays")
    nb, npairs = t.size, a.size
    plan = get_plan(n, device, n_shell_fields=6)
    k_all, pk = _f32(device, npairs, nb + 2), _f32(device, npairs, nb + 2)
    B, Q = _f32(device, npairs, nb), _f32(device, npairs, nb)
    check(lib.jps_bispec_pairs(plan.handle, ptr(mesh), int(bool(normalise)), float(box_size), _edge_ptr(a),
                               _edge_ptr(b), npairs, _edge_ptr(t), nb, int(mas_order), ptr(k_all), ptr(pk),
                               ptr(B), ptr(Q), stream_ptr()), "jps_bispec_pairs")
    th = torch.from_numpy(t).to(device) if kind.on_device else (torch.from_numpy(t) if kind.is_torch else t)
    return kind.out(k_all), kind.out(pk), th, kind.out(B), kind.out(Q)


def compute_2pt_correlations(delta, box_size, s_edges, k_edges, *, mas_order=2, normalise=False):
    """``(k3D, Pk3D, Nmodes3D_pk, r3D, xi3D)`` with ONE forward FFT, like
    /root/reference/src/correlations.py:641."""
    device = require_cuda()
    kind = ArrayKind(delta)
    mesh, n = _mesh_arg(delta, device)
    se, ke = _host_edges(s_edges), _host_edges(k_edges)
    ns, nk = se.size - 1, ke.size - 1
    plan = get_plan(n, device, n_shell_fields=1)
    k3d, pk, nmk = _f32(device, nk), _f32(device, nk, 3), _f32(device, nk)
    r3d, xi, nmx = _f32(device, ns), _f32(device, ns, 3), _f32(device, ns)
    check(lib.jps_compute_2pt_correlations(plan.handle, ptr(mesh), int(bool(normalise)), float(box_size),
                                           _edge_ptr(se), ns, _edge_ptr(ke), nk, int(mas_order),
                                           ptr(k3d), ptr(pk), ptr(nmk), ptr(r3d), ptr(xi), ptr(nmx),
                                           stream_ptr()), "jps_compute_2pt_correlations")
    return tuple(kind.out(t) for t in (k3d, pk, nmk, r3d, xi))


def compute_all_correlations(delta, box_size, s_edges, k_edges, k1, k2, theta, *, mas_order=2,
                             normalise=False):
    """The 11 outputs of /root/reference/src/correlations.py:465 with ONE forward FFT:
    ``(k3D, Pk3D, Nmodes3D_pk, r3D, xi3D, Nmodes3D_xi, k_all, Pk, theta, B, Q)``."""
    device = require_cuda()
    kind = ArrayKind(delta)
    mesh, n = _mesh_arg(delta, device)
    se, ke, t = _host_edges(s_edges), _host_edges(k_edges), _theta_arg(theta)
    ns, nk, nb = se.size - 1, ke.size - 1, t.size
    plan = get_plan(n, device, n_shell_fields=6)
    k3d, pk, nmk = _f32(device, nk), _f32(device, nk, 3), _f32(device, nk)
    r3d, xi, nmx = _f32(device, ns), _f32(device, ns, 3), _f32(device, ns)
    k_all, pks, B, Q = _f32(device, nb + 2), _f32(device, nb + 2), _f32(device, nb), _f32(device, nb)
    check(lib.jps_compute_all_correlations(plan.handle, ptr(mesh), int(bool(normalise)), float(box_size),
                                           _edge_ptr(se), ns, _edge_ptr(ke), nk, float(k1), float(k2),
                                           _edge_ptr(t), nb, int(mas_order),
                                           ptr(k3d), ptr(pk), ptr(nmk), ptr(r3d), ptr(xi), ptr(nmx),
                                           ptr(k_all), ptr(pks), ptr(B), ptr(Q), stream_ptr()),
          "jps_compute_all_correlations")
    th = torch.from_numpy(t).to(device) if kind.on_device else (torch.from_numpy(t) if kind.is_torch else t)
    o = kind.out
    return o(k3d), o(pk), o(nmk), o(r3d), o(xi), o(nmx), o(k_all), o(pks), th, o(B), o(Q)


class PaintPowspec:
    """Reusable end-to-end pipeline (what bench.py times): particles -> mesh -> delta_k ->
    multipoles, one C-ABI call per step, every buffer allocated once.

    Every pipeline OWNS its plan (cuFFT plans, delta_k buffer, accumulators, bin tables) and its
    bucketing workspace: two pipelines may run concurrently on two CUDA streams of one GPU (the
    covariance batch of BASELINE.json configs[4] does), and nothing the functional API does can
    invalidate a live pipeline.  One pipeline must be driven from one stream at a time
    (include/jps.h).  Inputs are validated (float32, CUDA, same device, equal lengths); the
    weights are read with stride 1, strided views are made contiguous."""

    def __init__(self, n_mesh, box_size, k_edges, *, order=2, compat="fixed", method="auto",
                 n_part_max=0, shot_noise=0.0, wrap=True, device=None, plan=None, fft="auto"):
        """fft: "3d" = one monolithic cuFFT 3-D R2C plan; "pencil" = three contiguous batched 1-D passes with
        transposing kernels in between (JPS_PLAN_FFT_PENCIL; a second delta_k-sized buffer); "auto" = pencil
        from 1024^3 up, where the 3-D plan falls to a third of the HBM roofline (env JPS_FFT overrides)."""
        self.device = device or require_cuda()
        self.n = int(n_mesh)
        self.box = float(box_size)
        self.edges = _host_edges(k_edges)
        self.nb = self.edges.size - 1
        self.order, self.compat, self.method = int(order), compat, method
        self.wrap, self.shot_noise = bool(wrap), float(shot_noise)
        import os
        fft = os.environ.get("JPS_FFT", fft)
        if fft not in ("auto", "3d", "pencil"):
            raise ValueError("fft must be 'auto', '3d' or 'pencil'")
        self.fft = ("pencil" if self.n >= 1024 else "3d") if fft == "auto" else fft
        if plan is not None:
            self.plan = plan
            self.fft = "pencil" if (plan.flags & _lib.PLAN_FFT_PENCIL) else "3d"
        else:
            self.plan = Plan(self.n, 0, self.device, _lib.PLAN_FFT_PENCIL if self.fft == "pencil" else 0)
        d = self.device
        self.mesh = torch.empty((self.n,) * 3, dtype=torch.float32, device=d)
        self.k3d = torch.empty(self.nb, dtype=torch.float32, device=d)
        self.pk = torch.empty((self.nb, 3), dtype=torch.float32, device=d)
        self.nm = torch.empty(self.nb, dtype=torch.float32, device=d)
        self.sums = torch.empty((self.nb, 3), dtype=torch.float64, device=d)
        self.counts = torch.empty(self.nb, dtype=torch.int64, device=d)
        self.ws, self.ws_bytes = (None, 0)
        if n_part_max:
            self.reserve(n_part_max)

    def reserve(self, n_part):
        self.ws, self.ws_bytes = new_paint_workspace(self.n, n_part, self.order, _lib.METHOD[self.method], self.device)

    def __call__(self, x, y, z, w=None, xmin=0.0, ymin=0.0, zmin=0.0):
        """x,y,z[,w]: float32 CUDA tensors.  Returns (k3D, Pk3D, Nmodes3D) device tensors that are
        overwritten by the next call."""
        x, y, z, w, stride = check_particles(x, y, z, w, self.device)
        npart = x.numel()
        if paint_workspace_bytes(self.n, npart, self.order, _lib.METHOD[self.method]) > self.ws_bytes:
            self.reserve(npart)
        if self.n >= 512 and npart >= (1 << 18) and self.method != "atomic":
            return self._call_overlapped(x, y, z, w, stride, npart, xmin, ymin, zmin)
        check(lib.jps_paint_powspec(self.plan.handle, ptr(x), ptr(y), ptr(z), ptr(w), stride, npart,
                                    float(xmin), float(ymin), float(zmin), self.box, self.order,
                                    int(self.wrap), _lib.COMPAT[self.compat], _lib.METHOD[self.method],
                                    _edge_ptr(self.edges), self.nb, self.shot_noise, ptr(self.mesh),
                                    ptr(self.ws), self.ws_bytes, ptr(self.k3d), ptr(self.pk), ptr(self.nm),
                                    ptr(self.sums), ptr(self.counts), stream_ptr()), "jps_paint_powspec")
        return self.k3d, self.pk, self.nm

    def _call_overlapped(self, x, y, z, w, stride, npart, xmin, ymin, zmin):
        """Big meshes: the mesh is zeroed on a side stream WHILE the bucketing passes run (they never touch the
        mesh and leave most of the HBM bandwidth idle): 2048^3 = 34 GB = 4.6 ms of memset off the critical path.
        Same C-ABI work as jps_paint_powspec, split at the phase boundary (jps_paint_slab_phase)."""
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_zero_stream", None) is None:
            self._zero_stream = torch.cuda.Stream(self.device)
            self._zero_ev = (torch.cuda.Event(), torch.cuda.Event())
        self._zero_ev[0].record(main)                     # the previous step's transform has consumed the mesh
        self._zero_stream.wait_event(self._zero_ev[0])
        with torch.cuda.stream(self._zero_stream):
            self.mesh.zero_()
            self._zero_ev[1].record(self._zero_stream)
        args = (self.n, 0, self.n, ptr(x), ptr(y), ptr(z), ptr(w), stride, npart, float(xmin), float(ymin), float(zmin),
                self.box, self.order, int(self.wrap), _lib.COMPAT[self.compat], _lib.VARIANT_VEC, _lib.PAINT_SORTED,
                ptr(self.mesh), ptr(self.ws), self.ws_bytes)
        check(lib.jps_paint_slab_phase(*args, _lib.PAINT_PHASE_BUCKET, 0, 0, stream_ptr()), "jps_paint (bucket)")
        main.wait_event(self._zero_ev[1])
        check(lib.jps_paint_slab_phase(*args, _lib.PAINT_PHASE_DEPOSIT, 0, lib.jps_paint_tile_rows(self.n), stream_ptr()),
              "jps_paint (deposit)")
        return self.finish()

    # ---- the same pipeline in pieces (HostPipeline streams the catalogue through these)
    def paint_chunk(self, x, y, z, w=None, xmin=0.0, ymin=0.0, zmin=0.0):
        """self.mesh += deposit of one piece of the catalogue (zero self.mesh before the first piece)."""
        x, y, z, w, stride = check_particles(x, y, z, w, self.device)
        npart = x.numel()
        if paint_workspace_bytes(self.n, npart, self.order, _lib.METHOD[self.method]) > self.ws_bytes:
            self.reserve(npart)
        check(lib.jps_paint(self.n, ptr(x), ptr(y), ptr(z), ptr(w), stride, npart, float(xmin), float(ymin),
                            float(zmin), self.box, self.order, int(self.wrap), _lib.COMPAT[self.compat],
                            _lib.VARIANT_VEC, _lib.METHOD[self.method], ptr(self.mesh), ptr(self.ws), self.ws_bytes,
                            stream_ptr()), "jps_paint")

    def finish(self):
        """FFT + multipoles of self.mesh (density contrast folded in through the DC mode)."""
        check(lib.jps_powspec(self.plan.handle, ptr(self.mesh), 1, self.box, _edge_ptr(self.edges), self.nb,
                              self.order, self.shot_noise, ptr(self.k3d), ptr(self.pk), ptr(self.nm),
                              ptr(self.sums), ptr(self.counts), stream_ptr()), "jps_powspec")
        return self.k3d, self.pk, self.nm


def host_chunks(n_chunks, n_part, n_cells):
    """Pieces a host catalogue is streamed in.  Every piece pays the deposit's per-MESH cost again (zero,
    convert and flush every tile: ~1.2 ps per cell) while it hides a copy that shrinks with the piece, so big
    sparse meshes want few pieces and small dense ones many: measured on one B200, 1e9 particles on 2048^3:
    2 / 4 / 8 / 16 pieces -> 336 / 322 / 368 / 580 ms end to end; 1e8 particles on 512^3: 8 pieces, PCIe bound.
    Rule: copy time of the whole catalogue (55 GB/s) over five times the per-mesh cost, clamped to [2, 8];
    an explicit n_chunks or JPS_HOST_CHUNKS overrides."""
    import os
    env = os.environ.get("JPS_HOST_CHUNKS")
    if env:
        return max(1, int(env))
    if n_chunks is not None:
        return max(1, int(n_chunks))
    copy_s = 12.0 * n_part / 55e9
    fixed_s = 1.2e-12 * n_cells
    return int(min(8, max(2, copy_s / (5.0 * fixed_s))))


class HostPipeline:
    """End-to-end call for catalogues that live in HOST memory (what a user of the reference has
    after np.loadtxt, tests/correlations.py:29-31): host->device copy of x, y, z[, w], paint,
    FFT, multipoles, device->host copy of (k3D, Pk3D, Nmodes3D).

    The catalogue is streamed in ``n_chunks`` pieces: a copy stream moves piece c+1 over PCIe
    while the compute stream buckets and paints piece c into the (accumulating) mesh, so the
    step costs max(H2D, compute) instead of their sum.  Device staging buffers and pinned result
    buffers are allocated once."""

    def __init__(self, pipe: PaintPowspec, n_part_max: int, weighted: bool = False, n_chunks=None):
        self.pipe = pipe
        d = pipe.device
        self.cap = int(n_part_max)
        self.n_chunks = host_chunks(n_chunks, self.cap, pipe.n ** 3)
        self.dev = [torch.empty(self.cap, dtype=torch.float32, device=d) for _ in range(4 if weighted else 3)]
        self.k3d = torch.empty(pipe.nb, dtype=torch.float32).pin_memory()
        self.pk = torch.empty((pipe.nb, 3), dtype=torch.float32).pin_memory()
        self.nm = torch.empty(pipe.nb, dtype=torch.float32).pin_memory()
        self.copy_stream = torch.cuda.Stream(device=d)
        self.paint_done = None                      # event: previous step no longer reads self.dev
        chunk = -(-self.cap // self.n_chunks)
        pipe.reserve(chunk)

    def __call__(self, x, y, z, w=None, xmin=0.0, ymin=0.0, zmin=0.0):
        """x, y, z[, w]: host arrays (NumPy or torch, ideally pinned).  Returns NumPy arrays (copies:
        the pinned staging buffers are reused by the next call)."""
        host = [x, y, z] + ([w] if w is not None else [])
        n = len(x)
        if n > self.cap or len(host) > len(self.dev):
            raise ValueError("HostPipeline: catalogue larger than the buffers it was built for")
        host = [h if isinstance(h, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(h, dtype=np.float32))
                for h in host]
        p = self.pipe
        compute = torch.cuda.current_stream(p.device)
        nchunks = min(self.n_chunks, max(1, n // 65536))
        bounds = [n * c // nchunks for c in range(nchunks + 1)]
        ready = []
        with torch.cuda.stream(self.copy_stream):
            if self.paint_done is not None:
                self.copy_stream.wait_event(self.paint_done)
            for c in range(nchunks):
                lo, hi = bounds[c], bounds[c + 1]
                for h, d in zip(host, self.dev):
                    d[lo:hi].copy_(h[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                ready.append(ev)
        p.mesh.zero_()
        for c in range(nchunks):
            lo, hi = bounds[c], bounds[c + 1]
            compute.wait_event(ready[c])
            wd = self.dev[3][lo:hi] if w is not None else None
            p.paint_chunk(self.dev[0][lo:hi], self.dev[1][lo:hi], self.dev[2][lo:hi], wd, xmin, ymin, zmin)
        self.paint_done = torch.cuda.Event()
        self.paint_done.record(compute)
        k3d, pk, nm = p.finish()
        self.k3d.copy_(k3d, non_blocking=True)
        self.pk.copy_(pk, non_blocking=True)
        self.nm.copy_(nm, non_blocking=True)
        compute.synchronize()
        return self.k3d.numpy().copy(), self.pk.numpy().copy(), self.nm.numpy().copy()


def paint_powspec(x, y, z, w, xmin, ymin, zmin, box_size, n_bins, k_edges, *, order=2,
                  compat="fixed", method="auto", wrap=True, shot_noise=0.0):
    """One-shot convenience wrapper of :class:`PaintPowspec` accepting host or device arrays."""
    device = require_cuda()
    kind = ArrayKind(x)
    xd = to_device_f32(x, device, allow_strided=True)
    yd = to_device_f32(y, device, allow_strided=True)
    zd = to_device_f32(z, device, allow_strided=True)
    wd = None if w is None else to_device_f32(w, device)
    pipe = PaintPowspec(n_bins, box_size, k_edges, order=order, compat=compat, method=method,
                        n_part_max=xd.numel(), shot_noise=shot_noise, wrap=wrap, device=device,
                        plan=get_plan(int(n_bins), device))      # the functional API's per-stream plan
    k3d, pk, nm = pipe(xd, yd, zd, wd, xmin, ymin, zmin)
    return kind.out(k3d.clone()), kind.out(pk.clone()), kind.out(nm.clone())
