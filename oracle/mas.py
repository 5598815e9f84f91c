"""CPU oracle for mesh painting -- TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/__init__.py).

Restates, in vectorised NumPy, the arithmetic of
  * /root/reference/src/mas.py:88-153  ``cic_mas_vec``  (variant="vec")
  * /root/reference/src/mas.py:5-87    ``cic_mas``      (variant="scan")
including the behaviours SURVEY.md section 8 lists as quirks Q1-Q6, and defines
the TSC / PCS painters the north star asks for (absent from the reference ->
parity UNPINNED for order 3 and 4; standard B-spline kernels on integer nodes).

Cell choice and in-cell offsets are ALWAYS computed in float32 exactly as the
reference does (they decide which cells are touched); ``precision`` only selects
how the weights are multiplied and accumulated:
  "f64": products and sums in float64 (bincount)         -> exact_f64 oracle
  "f32": products in float32 in the reference's order, sums serial float32 in the
         reference's scatter order (corner-major for "vec", particle-major for
         "scan")                                         -> faithful_f32 oracle
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def grid_positions(x, xmin, box_size, n_bins):
    """pos = (x - xmin) * inv_bin_size in float32.  /root/reference/src/mas.py:100-105 (Q4)."""
    bin_size = F32(box_size) / F32(n_bins)
    inv = F32(1.0) / bin_size
    return (np.asarray(x, dtype=F32) - F32(xmin)) * inv


def _pymod(a, n):
    return np.mod(a, n)  # numpy mod has the sign of the divisor, like jnp / Python


def _cic_reference_axis(pos, n, wrap, variant):
    """Per-axis base index, +1 index, dd and md for compat='reference'.
    vec : /root/reference/src/mas.py:107-140 ; scan: :47-66 with wrap_fun :20-28, nowrap_fun :30-38."""
    i = np.trunc(pos).astype(np.int32)          # jnp.int32() truncates toward zero (Q3)
    dd = pos - i.astype(F32)
    ip = i + 1
    if variant == "vec":
        if wrap:
            ip = _pymod(ip + n, n)
        else:
            ip = np.where(ip >= n, 0, ip)       # dd is NOT zeroed: the second where never fires (Q2)
        md = F32(1.0) - dd
    elif variant == "scan":
        md = F32(1.0) - dd                      # computed before the cond (mas.py:55-57)
        over = ip >= n
        if wrap:
            ip = np.where(over, ip - n, ip)
        else:
            ip = np.where(over, 0, ip)
            dd = np.where(over, F32(0.0), dd)
    else:
        raise ValueError(variant)
    return i, ip.astype(np.int32), dd.astype(F32), md.astype(F32)


def _scatter_norm(idx, n):
    """JAX .at[] index semantics: negatives wrap once, still-out-of-range dropped (Q3)."""
    idx = np.where(idx < 0, idx + n, idx)
    ok = (idx >= 0) & (idx < n)
    return idx, ok


# corner order and weight factors of the reference's 8 scatters, mas.py:142-151.
# entries: (use ixp?, use iyp?, use izp?, x-factor, y-factor, z-factor) with 'd' = dd, 'm' = md
_REF_CORNERS = (
    (0, 0, 0, "m", "m", "m"),
    (1, 0, 0, "d", "m", "m"),
    (0, 1, 0, "m", "d", "m"),
    (0, 0, 1, "m", "m", "d"),
    (1, 1, 0, "d", "d", "m"),
    (1, 0, 1, "d", "m", "d"),
    (0, 1, 1, "m", "m", "d"),   # Q1: the reference writes mdx*mdy*ddz here (should be mdx*ddy*ddz)
    (1, 1, 1, "d", "d", "d"),
)


def bspline_axis(pos, order, dtype):
    """Node indices (unwrapped, int64) and weights for the order-2/3/4 B-spline on integer
    nodes.  order 2 = CIC, 3 = TSC, 4 = PCS.  Returns (idx[s, Np], wgt[s, Np])."""
    pos = np.asarray(pos, dtype=F32)
    if order == 2:
        i0 = np.floor(pos)
        d = (pos - i0).astype(dtype)
        one = dtype(1.0)
        w = [one - d, d]
        base = i0.astype(np.int64)
    elif order == 3:
        j0 = np.floor(pos + F32(0.5))
        d = (pos - j0).astype(dtype)
        h, tq = dtype(0.5), dtype(0.75)
        w = [h * (h - d) * (h - d), tq - d * d, h * (h + d) * (h + d)]
        base = j0.astype(np.int64) - 1
    elif order == 4:
        i0 = np.floor(pos)
        d = (pos - i0).astype(dtype)
        one, six, four, three = dtype(1.0), dtype(6.0), dtype(4.0), dtype(3.0)
        sixth = one / six
        e = one - d
        w = [
            e * e * e * sixth,
            (four - six * d * d + three * d * d * d) * sixth,
            (four - six * e * e + three * e * e * e) * sixth,
            d * d * d * sixth,
        ]
        base = i0.astype(np.int64) - 1
    else:
        raise ValueError("order must be 2 (CIC), 3 (TSC) or 4 (PCS)")
    idx = np.stack([base + s for s in range(order)])
    return idx, np.stack(w)


def paint(mesh, x, y, z, w, xmin, ymin, zmin, box_size, n_bins, wrap=True, *,
          order=2, compat="reference", variant="vec", precision="f64"):
    """Deposit particles onto ``mesh`` (accumulating, Q5) and return the new mesh.

    Returns float64 for precision='f64', float32 for 'f32'."""
    n = int(n_bins)
    mesh = np.asarray(mesh)
    assert mesh.shape == (n, n, n)
    x = np.asarray(x, dtype=F32); y = np.asarray(y, dtype=F32); z = np.asarray(z, dtype=F32)
    w = np.ones_like(x) if w is None else np.asarray(w, dtype=F32)
    px = grid_positions(x, xmin, box_size, n)
    py = grid_positions(y, ymin, box_size, n)
    pz = grid_positions(z, zmin, box_size, n)
    acc_dtype = np.float64 if precision == "f64" else F32
    out = mesh.astype(acc_dtype).ravel().copy()

    contributions = []  # list of (flat_index int64 [Np], weight [Np], keep [Np]) in reference scatter order
    if order == 2 and compat == "reference":
        ax = [_cic_reference_axis(p, n, bool(wrap), variant) for p in (px, py, pz)]
        for (ux, uy, uz, fx, fy, fz) in _REF_CORNERS:
            sel = []
            for (i, ip, dd, md), use_p, f in zip(ax, (ux, uy, uz), (fx, fy, fz)):
                sel.append((ip if use_p else i, dd if f == "d" else md))
            (jx, wx), (jy, wy), (jz, wz) = sel
            jx, okx = _scatter_norm(jx.astype(np.int64), n)
            jy, oky = _scatter_norm(jy.astype(np.int64), n)
            jz, okz = _scatter_norm(jz.astype(np.int64), n)
            if precision == "f64":
                wt = wx.astype(np.float64) * wy.astype(np.float64) * wz.astype(np.float64) * w.astype(np.float64)
            else:
                wt = ((wx * wy) * wz) * w      # left-to-right float32, as mas.py:142-151
            contributions.append(((jx * n + jy) * n + jz, wt, okx & oky & okz))
    else:
        dt = np.float64 if precision == "f64" else F32
        (ix, wx), (iy, wy), (iz, wz) = (bspline_axis(p, order, dt) for p in (px, py, pz))
        wv = w.astype(dt)
        for a in range(order):
            for b in range(order):
                for c in range(order):
                    jx, jy, jz = ix[a], iy[b], iz[c]
                    if wrap:
                        ok = np.ones(jx.shape, bool)
                        jx, jy, jz = _pymod(jx, n), _pymod(jy, n), _pymod(jz, n)
                    else:
                        ok = (jx >= 0) & (jx < n) & (jy >= 0) & (jy < n) & (jz >= 0) & (jz < n)
                    wt = ((wx[a] * wy[b]) * wz[c]) * wv
                    contributions.append(((jx * n + jy) * n + jz, wt, ok))

    if precision == "f64":
        for flat, wt, ok in contributions:
            out += np.bincount(flat[ok], weights=wt[ok], minlength=out.size)
    elif variant == "scan" and order == 2 and compat == "reference":
        # particle-major serial order (lax.scan, mas.py:40-83)
        flat = np.stack([c[0] for c in contributions], axis=1)
        wt = np.stack([c[1] for c in contributions], axis=1)
        ok = np.stack([c[2] for c in contributions], axis=1)
        # the scan variant's corner order (mas.py:69-78) equals _REF_CORNERS order
        np.add.at(out, flat[ok], wt[ok])
    else:
        for flat, wt, ok in contributions:      # corner-major serial order (Q6)
            np.add.at(out, flat[ok], wt[ok])
    return out.reshape(n, n, n)


def density_contrast(rho):
    """delta = rho/mean(rho) - 1, the caller-side step at /root/reference/tests/correlations.py:49-50."""
    rho = np.asarray(rho)
    return rho / rho.mean(dtype=rho.dtype) - rho.dtype.type(1.0)
