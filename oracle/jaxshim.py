"""NumPy stand-in for the slice of the JAX API that the reference's hot path uses.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  JAX is not installed in this
image and cannot be installed (no network), so the reference cannot run as
shipped.  This module registers fake ``jax``, ``jax.numpy``, ``jax.lax``,
``jax.experimental.loops`` modules (and the removed
``numpy.core.defchararray._join_dispatcher`` symbol that
/root/reference/src/correlations.py:1 imports) so that the UNMODIFIED reference
sources ``/root/reference/src/mas.py`` and ``/root/reference/src/correlations.py``
can be imported and executed line by line on NumPy.  It is how the golden
vectors under tests/golden/ were produced (oracle/run_reference.py).

What this file encodes are JAX *semantics*, not reference code:
  * x64 disabled: every value is at most 32 bit (float32 / int32 / complex64);
    Python scalars are weak-typed; int (op) float -> float32.
  * ``jnp.int32(float)`` truncates toward zero.
  * ``x.at[idx].add/set`` is functional; negative indices wrap once (+size),
    indices still out of range are dropped (scatter mode FILL_OR_DROP);
    duplicate ``add`` updates are applied serially in element order (XLA CPU).
  * ``jnp.histogram`` = searchsorted(side='right'), right-most edge inclusive,
    unweighted counts have the dtype of the data (float32).
  * ``jnp.sinc(x)`` = sin(pi x)/(pi x), 1 at 0;  rfftn unnormalised, irfftn 1/N^3.
  * ``jax.jit`` turns Python scalar arguments into 0-d 32-bit arrays (so e.g.
    ``2*pi/box_size`` is evaluated in float32, as it is under tracing).
"""
from __future__ import annotations

import sys
import types

import numpy as np
import scipy.fft as _sfft

_F32, _I32, _C64 = np.float32, np.int32, np.complex64

_FLOAT_UFUNCS = {
    np.true_divide, np.sqrt, np.sin, np.cos, np.tan, np.exp, np.log, np.arctan2,
    np.arcsin, np.arccos, np.arctan, np.log10, np.log2, np.exp2, np.sinh, np.cosh,
    np.tanh, np.reciprocal,
}


def _demote(a):
    """64-bit -> 32-bit, the x64-disabled canonicalisation."""
    if isinstance(a, np.ndarray) or isinstance(a, np.generic):
        a = np.asarray(a)
        k = a.dtype
        if k == np.float64:
            return a.astype(_F32)
        if k == np.int64 or k == np.uint64:
            return a.astype(_I32)
        if k == np.complex128:
            return a.astype(_C64)
    return a


class JArray(np.ndarray):
    """ndarray with JAX dtype promotion and the functional ``.at`` property."""

    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        ins = []
        for i in inputs:
            if isinstance(i, np.ndarray):
                i = _demote(i.view(np.ndarray))
            elif isinstance(i, np.generic):
                i = _demote(i)
            ins.append(i)
        inexact = any(
            isinstance(i, (float, complex)) or (isinstance(i, np.ndarray) and i.dtype.kind in "fc")
            for i in ins
        )
        if inexact or ufunc in _FLOAT_UFUNCS:
            ins = [
                i.astype(_F32) if isinstance(i, np.ndarray) and i.dtype.kind in "iub" else i
                for i in ins
            ]
            # int python scalars next to float arrays stay weak (numpy NEP 50 does that)
        if out is not None:
            out = tuple(o.view(np.ndarray) if isinstance(o, np.ndarray) else o for o in out)
            kw["out"] = out
        res = getattr(ufunc, method)(*ins, **kw)
        if isinstance(res, tuple):
            return tuple(_wrap(r) for r in res)
        return _wrap(res)

    @property
    def at(self):
        return _At(self)

    # reductions keep float32 (numpy already does for f32 input); wrap results
    def sum(self, *a, **k):
        return _wrap(np.asarray(self).sum(*a, **k))

    def mean(self, *a, **k):
        return _wrap(np.asarray(self).mean(*a, **k))

    def flatten(self, *a, **k):
        return _wrap(np.asarray(self).flatten(*a, **k))


def _wrap(x):
    x = _demote(x)
    if isinstance(x, np.ndarray):
        return x.view(JArray)
    if isinstance(x, np.generic):
        return np.asarray(x).view(JArray)
    return x


def _raw(x):
    if isinstance(x, np.ndarray):
        return x.view(np.ndarray)
    return x


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


def _normalise_index(idx, shape):
    """JAX scatter index semantics for integer-array indices: wrap negatives once,
    then flag what is still out of range (to be dropped).  Slices pass through."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    out, keep = [], None
    for ax, i in enumerate(idx):
        if isinstance(i, slice) or i is None or i is Ellipsis:
            out.append(i)
            continue
        i = np.asarray(_raw(i))
        if i.dtype.kind not in "iu":
            raise TypeError("shim: only integer / slice indices are supported in .at[]")
        n = shape[ax]
        i = np.where(i < 0, i + n, i)
        ok = (i >= 0) & (i < n)
        keep = ok if keep is None else (keep & ok)
        out.append(i)
    return tuple(out), keep


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _prep(self, val):
        base = np.array(_raw(self.arr), copy=True)
        idx, keep = _normalise_index(self.idx, base.shape)
        val = _raw(_demote(np.asarray(_raw(val)))) if not isinstance(val, (int, float)) else val
        if not isinstance(val, (int, float)):
            val = val.astype(base.dtype, copy=False)
        return base, idx, keep, val

    def _filtered(self, base, idx, keep, val):
        if keep is None or np.all(keep):
            return idx, val
        arr_idx = [i for i in idx if isinstance(i, np.ndarray)]
        bshape = np.broadcast(*arr_idx).shape
        keep = np.broadcast_to(keep, bshape)
        new_idx = tuple(
            np.broadcast_to(i, bshape)[keep] if isinstance(i, np.ndarray) else i for i in idx
        )
        if isinstance(val, np.ndarray) and val.shape != ():
            val = np.broadcast_to(val, bshape)[keep]
        return new_idx, val

    def add(self, val):
        base, idx, keep, val = self._prep(val)
        idx, val = self._filtered(base, idx, keep, val)
        np.add.at(base, idx if len(idx) > 1 else idx[0], val)  # serial, in element order
        return _wrap(base)

    def set(self, val):
        base, idx, keep, val = self._prep(val)
        idx, val = self._filtered(base, idx, keep, val)
        base[idx if len(idx) > 1 else idx[0]] = val  # duplicates: last one wins
        return _wrap(base)


# --------------------------------------------------------------------------- jnp
def _asj(x, dtype=None):
    if isinstance(x, (list, tuple)):
        x = np.array([_raw(np.asarray(_raw(e))) for e in x])
    a = np.asarray(_raw(x))
    if dtype is not None:
        a = a.astype(dtype)
    return _wrap(a)


def _zeros(shape, dtype=_F32):
    return _wrap(np.zeros(shape, dtype=dtype))


def _ones(shape, dtype=_F32):
    return _wrap(np.ones(shape, dtype=dtype))


def _arange(*a, dtype=None):
    r = np.arange(*[_raw(v) if not isinstance(v, JArray) else v.item() for v in a])
    if dtype is not None:
        r = r.astype(dtype)
    return _wrap(r)


def _where(c, a, b):
    c = np.asarray(_raw(c))
    wa, wb = isinstance(a, (int, float)), isinstance(b, (int, float))
    A = a if wa else _demote(np.asarray(_raw(a)))
    B = b if wb else _demote(np.asarray(_raw(b)))
    if wa and wb:
        dt = _F32 if isinstance(a, float) or isinstance(b, float) else _I32
    elif wa:
        dt = B.dtype if not (isinstance(a, float) and B.dtype.kind in "iub") else _F32
    elif wb:
        dt = A.dtype if not (isinstance(b, float) and A.dtype.kind in "iub") else _F32
    else:
        dt = np.result_type(A, B)
    return _wrap(np.where(c, np.asarray(A, dtype=dt), np.asarray(B, dtype=dt)))


class _ScalarType:
    """``jnp.float32`` / ``jnp.int32``: usable as a dtype and as a converter."""

    def __init__(self, np_type):
        self.dtype = np.dtype(np_type)

    def __call__(self, x):
        a = np.asarray(_raw(x))
        if self.dtype.kind in "iu" and a.dtype.kind in "fc":
            a = np.trunc(a)  # XLA convert float -> int: round toward zero
        return _wrap(a.astype(self.dtype))


_int32, _float32 = _ScalarType(_I32), _ScalarType(_F32)


def _unary(fn):
    def f(x):
        a = _demote(np.asarray(_raw(x)))
        if a.dtype.kind in "iub":
            a = a.astype(_F32)
        return _wrap(fn(a))
    return f


def _sinc(x):
    x = _demote(np.asarray(_raw(x)))
    if x.dtype.kind in "iub":
        x = x.astype(_F32)
    pix = _F32(np.pi) * x
    safe = np.where(x == 0, _F32(1), pix)
    return _wrap(np.where(x == 0, _F32(1), np.sin(safe) / safe).astype(_F32))


def _histogram(a, bins, weights=None):
    a = _demote(np.asarray(_raw(a)))
    edges = _demote(np.asarray(_raw(bins)))
    idx = np.searchsorted(edges, a, side="right")
    idx = np.where(a == edges[-1], len(edges) - 1, idx)
    w = np.ones_like(a) if weights is None else _demote(np.asarray(_raw(weights)))
    counts = np.zeros(len(edges) + 1, dtype=w.dtype)  # slot len(edges) collects a > last edge
    np.add.at(counts, idx, w)  # serial accumulation in w.dtype, element order
    return _wrap(counts[1:len(edges)]), _wrap(edges)


def _nansum(x, *a, **k):
    return _wrap(np.nansum(np.asarray(_raw(x)), *a, **k))


def _broadcast_to(x, shape):
    return _wrap(np.broadcast_to(np.asarray(_raw(x)), shape))


def _logical_and(a, b):
    return _wrap(np.logical_and(np.asarray(_raw(a)), np.asarray(_raw(b))))


def _rfftn(x, axes=None):
    a = np.asarray(_raw(x)).astype(_F32, copy=False)
    return _wrap(_sfft.rfftn(a, axes=axes).astype(_C64, copy=False))


def _irfftn(x, s=None, axes=None):
    a = np.asarray(_raw(x)).astype(_C64)
    return _wrap(_sfft.irfftn(a, s=tuple(s) if s is not None else None, axes=axes).astype(_F32, copy=False))


def _einsum(*a, **k):
    return _wrap(np.einsum(*[_raw(v) for v in a], **k))


# --------------------------------------------------------------------------- jax.*
def _jit(fn=None, **_kw):
    if fn is None:
        return lambda f: _jit(f)

    def wrapped(*args, **kwargs):
        def conv(v):
            if isinstance(v, bool):
                return _wrap(np.asarray(v))
            if isinstance(v, int):
                return _wrap(np.asarray(v, dtype=_I32))
            if isinstance(v, float):
                return _wrap(np.asarray(v, dtype=_F32))
            if isinstance(v, (np.ndarray, np.generic)):
                return _wrap(np.asarray(v))
            return v
        return fn(*[conv(a) for a in args], **{k: conv(v) for k, v in kwargs.items()})

    wrapped.__wrapped__ = fn
    wrapped.__name__ = getattr(fn, "__name__", "jitted")
    return wrapped


def _cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(np.asarray(_raw(pred))) else false_fun(*operands)


def _tree_index(xs, i):
    if isinstance(xs, (tuple, list)):
        return tuple(_tree_index(x, i) for x in xs)
    return _wrap(np.asarray(_raw(xs))[i])


def _tree_len(xs):
    if isinstance(xs, (tuple, list)):
        return _tree_len(xs[0])
    return len(xs)


def _tree_stack(ys):
    if isinstance(ys[0], (tuple, list)):
        return tuple(_tree_stack([y[j] for y in ys]) for j in range(len(ys[0])))
    return _wrap(np.stack([np.asarray(_raw(y)) for y in ys]))


def _scan(f, init, xs):
    carry, ys = init, []
    for i in range(_tree_len(xs)):
        carry, y = f(carry, _tree_index(xs, i))
        ys.append(y)
    return carry, (_tree_stack(ys) if ys else None)


# --------------------------------------------------------------------------- jax.random (injected draws)
# /root/reference/src/populate_field.py draws from jax.random (threefry).  That stream cannot be
# reproduced without JAX, and the device generator does not try to (it uses Philox counters): what the
# golden vectors pin for that file is its ARITHMETIC.  So the stand-ins below hand out whatever the
# caller queued in RANDOM_FEED -- uniforms and Poisson counts chosen by oracle/make_mock_golden.py --
# with JAX's dtypes (float32 uniforms, int32 counts).
RANDOM_FEED = {"uniform": [], "poisson": []}


class _Key:
    def __init__(self, tag):
        self.tag = tag


def _prng_key(seed):
    return _Key((int(seed),))


def _split(key, num=2):
    return tuple(_Key(key.tag + (i,)) for i in range(num))


def _uniform(key, shape=(), dtype=_F32, minval=0.0, maxval=1.0):
    u = np.asarray(RANDOM_FEED["uniform"].pop(0), dtype=_F32)
    assert u.shape == tuple(shape), (u.shape, shape)
    return _wrap(u)


def _poisson(key, lam, shape=None, dtype=_I32):
    k = np.asarray(RANDOM_FEED["poisson"].pop(0)).astype(_I32)
    assert shape is None or k.shape == tuple(shape), (k.shape, shape)
    return _wrap(k)


def _argsort(a, axis=-1):
    return _wrap(np.argsort(np.asarray(_raw(a)), axis=axis, kind="stable"))


def _unravel_index(idx, shape):
    return tuple(_wrap(i) for i in np.unravel_index(np.asarray(_raw(idx)), tuple(shape)))


def _repeat(a, repeats, axis=None):
    return _wrap(np.repeat(np.asarray(_raw(a)), np.asarray(_raw(repeats)), axis=axis))


def install():
    """Register the fake modules.  Idempotent."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "__is_jps_shim__", False):
        return sys.modules["jax"]
    if "jax" in sys.modules:
        raise RuntimeError("a real jax is already imported; the shim is not needed")

    jax = types.ModuleType("jax")
    jax.__is_jps_shim__ = True
    jnp = types.ModuleType("jax.numpy")
    lax = types.ModuleType("jax.lax")
    exp = types.ModuleType("jax.experimental")
    loops = types.ModuleType("jax.experimental.loops")
    fft = types.ModuleType("jax.numpy.fft")

    jnp.pi, jnp.inf, jnp.nan = float(np.pi), float("inf"), float("nan")
    jnp.float32, jnp.int32 = _float32, _int32
    jnp.zeros, jnp.ones, jnp.arange, jnp.array, jnp.asarray = _zeros, _ones, _arange, _asj, _asj
    jnp.where, jnp.sinc, jnp.histogram, jnp.nansum = _where, _sinc, _histogram, _nansum
    jnp.sqrt, jnp.sin, jnp.cos, jnp.exp = _unary(np.sqrt), _unary(np.sin), _unary(np.cos), _unary(np.exp)
    jnp.broadcast_to, jnp.logical_and, jnp.einsum = _broadcast_to, _logical_and, _einsum
    jnp.argsort, jnp.unravel_index, jnp.repeat = _argsort, _unravel_index, _repeat
    jnp.sign, jnp.abs = _unary(np.sign), _unary(np.abs)
    random = types.ModuleType("jax.random")
    random.PRNGKey, random.split, random.uniform, random.poisson = _prng_key, _split, _uniform, _poisson
    jax.random = random
    fft.rfftn, fft.irfftn = _rfftn, _irfftn
    jnp.fft = fft
    lax.cond, lax.scan = _cond, _scan
    exp.loops = loops
    jax.numpy, jax.lax, jax.experimental, jax.jit = jnp, lax, exp, _jit

    sys.modules.update({
        "jax": jax, "jax.numpy": jnp, "jax.numpy.fft": fft, "jax.lax": lax,
        "jax.experimental": exp, "jax.experimental.loops": loops, "jax.random": random,
    })

    # /root/reference/src/correlations.py:1 imports a private numpy symbol that no
    # longer exists; give it a harmless stub (the reference never uses it).
    try:
        import numpy.core.defchararray as _dc  # noqa: F401  (numpy>=2 warns/fails)
        if not hasattr(_dc, "_join_dispatcher"):
            _dc._join_dispatcher = lambda *a, **k: None
    except Exception:
        core = sys.modules.get("numpy.core") or types.ModuleType("numpy.core")
        dc = types.ModuleType("numpy.core.defchararray")
        dc._join_dispatcher = lambda *a, **k: None
        core.defchararray = dc
        sys.modules.setdefault("numpy.core", core)
        sys.modules["numpy.core.defchararray"] = dc
    return jax


def load_reference(ref_root="/root/reference"):
    """Import the unmodified reference hot-path modules under the shim.
    Returns (mas_module, correlations_module)."""
    import importlib.util
    import os

    install()
    sys.dont_write_bytecode = True
    mods = []
    for name in ("mas", "correlations"):
        path = os.path.join(ref_root, "src", f"{name}.py")
        spec = importlib.util.spec_from_file_location(f"_jps_reference_{name}", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


def load_reference_module(name, ref_root="/root/reference"):
    """Import one more unmodified reference source (e.g. ``populate_field``) under the shim."""
    import importlib.util
    import os

    install()
    sys.dont_write_bytecode = True
    path = os.path.join(ref_root, "src", f"{name}.py")
    spec = importlib.util.spec_from_file_location(f"_jps_reference_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
