"""Catalogue text reader: whitespace-separated ASCII -> float32 rows on the GPU.

Drop-in for what the reference's scripts do on the host before painting
(/root/reference/tests/correlations.py:29-31, tests/positions.py:25-27)::

    particles = np.loadtxt(path, usecols=(0, 1, 2), dtype=np.float32)
    mask = ((particles < box_size) & (particles > 0)).all(axis=1)
    particles = jax.device_put(particles[mask].copy())

becomes ``particles = read_catalog_text(path, usecols=(0, 1, 2), box_size=box_size)``: the file's bytes
are read into pinned memory, copied to the device once and parsed there by libjps.so
(csrc/reader.cu); the values are bit-identical to NumPy's (decimal -> nearest double -> nearest
float32).  The few fields outside the exact envelope of the device converter (more than 19
significant digits, |decimal exponent| > 27, ``nan`` / ``inf``) are re-converted one by one on the host
with Python's ``float`` (the same strtod NumPy uses); a malformed or short row raises
``ValueError`` like ``np.loadtxt``.  There is no CPU parsing path for the bulk of the file.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from ._lib import check, lib
from .plan import require_cuda, stream_ptr

__all__ = ["read_catalog_text", "parse_catalog_bytes"]


def _pinned_file(path):
    nbytes = os.path.getsize(path)
    buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
    view = memoryview(buf.numpy())[:nbytes]
    with open(path, "rb", buffering=0) as f:
        got = 0
        while got < nbytes:
            k = f.readinto(view[got:])
            if not k:
                break
            got += k
    if got != nbytes:
        raise IOError(f"{path}: read {got} of {nbytes} bytes")
    return buf, nbytes


def parse_catalog_bytes(text, nbytes=None, usecols=(0, 1, 2), *, skiprows=0, comments="#", box_size=None,
                        bounds=None, host_bytes=None, return_info=False):
    """Parse ``text`` (uint8 CUDA tensor holding the file content) -> float32 tensor [n_rows, len(usecols)].

    ``box_size`` keeps only the rows with ``0 < value < box_size`` in every requested column (the
    reference's mask); ``bounds=(lo, hi)`` does the same with explicit limits.  ``host_bytes`` (any
    bytes-like copy of the content) is only consulted for rows the device defers to the host.
    """
    device = require_cuda()
    if not (isinstance(text, torch.Tensor) and text.is_cuda and text.dtype == torch.uint8 and text.is_contiguous()):
        raise TypeError("text must be a contiguous uint8 CUDA tensor")
    nbytes = int(text.numel() if nbytes is None else nbytes)
    cols = [int(c) for c in usecols]
    ncols = len(cols)
    if comments is None:
        comment = 0
    elif isinstance(comments, str) and len(comments) == 1:
        comment = ord(comments)
    else:
        raise ValueError("comments must be a single character or None")
    if box_size is not None and bounds is not None:
        raise ValueError("give box_size or bounds, not both")
    if box_size is not None:
        bounds = (0.0, float(box_size))
    filt, lo, hi = (1, float(bounds[0]), float(bounds[1])) if bounds is not None else (0, 0.0, 0.0)
    sp = stream_ptr()

    ws_bytes = lib.jps_text_workspace_bytes(nbytes, 0, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    n_lines_d = torch.zeros(1, dtype=torch.int64, device=device)
    check(lib.jps_text_count_lines(text.data_ptr(), nbytes, n_lines_d.data_ptr(), ws.data_ptr(), ws_bytes, sp),
          "jps_text_count_lines")
    n_lines = int(n_lines_d.item())                      # the one unavoidable sync: sizes the output

    out = torch.empty((max(n_lines, 1), ncols), dtype=torch.float32, device=device)
    counters = torch.zeros(4, dtype=torch.int64, device=device)
    ws_bytes = lib.jps_text_workspace_bytes(nbytes, n_lines, ncols)
    if ws.numel() < ws_bytes:
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    cols_c = (C.c_int * ncols)(*cols)
    slow_cap = 4096
    while True:
        slow = torch.empty((slow_cap, 2), dtype=torch.int64, device=device)
        check(lib.jps_text_parse(text.data_ptr(), nbytes, n_lines, int(skiprows), comment, cols_c, ncols, filt, lo, hi,
                                 out.data_ptr(), counters.data_ptr(), slow.data_ptr(), slow_cap, ws.data_ptr(),
                                 ws_bytes, sp), "jps_text_parse")
        n_rows, n_slow, n_bad, first_bad = (int(v) for v in counters.tolist())
        if n_slow <= slow_cap:
            break
        slow_cap = n_slow
    if n_bad:
        raise ValueError(f"{n_bad} row(s) have a non-numeric or missing column among {tuple(cols)}; "
                         f"first at line {first_bad + 1}")
    out = out[:n_rows]
    if n_slow:
        out = _patch_slow_rows(out, slow[:n_slow].cpu().numpy(), text, nbytes, host_bytes, cols, comment, bounds)
    if return_info:
        return out, {"n_lines": n_lines, "n_rows": int(out.shape[0]), "n_host_rows": n_slow, "nbytes": nbytes}
    return out


def _patch_slow_rows(out, slow, text, nbytes, host_bytes, cols, comment, bounds):
    """Host conversion of the rows the device listed (same strtod as NumPy), then the mask for them."""
    drop = []
    vals = np.empty((len(slow), len(cols)), dtype=np.float32)
    for k, (row, off) in enumerate(slow):
        if host_bytes is not None:
            end = min(off + 4096, nbytes)
            chunk = bytes(host_bytes[off:end])
        else:
            chunk = bytes(text[off:min(off + 4096, nbytes)].cpu().numpy())
        line = chunk.split(b"\n", 1)[0]
        if comment:
            line = line.split(bytes([comment]), 1)[0]
        tok = line.split()
        v = np.array([float(tok[c]) for c in cols], dtype=np.float64).astype(np.float32)
        vals[k] = v
        if bounds is not None and not bool(((v < np.float32(bounds[1])) & (v > np.float32(bounds[0]))).all()):
            drop.append(int(row))
    rows = torch.as_tensor(slow[:, 0].copy(), device=out.device)
    out[rows] = torch.as_tensor(vals, device=out.device)
    if drop:
        keep = torch.ones(out.shape[0], dtype=torch.bool, device=out.device)
        keep[torch.as_tensor(drop, device=out.device)] = False
        out = out[keep]
    return out


def read_catalog_text(path, usecols=(0, 1, 2), *, skiprows=0, comments="#", box_size=None, bounds=None,
                      return_info=False):
    """``np.loadtxt(path, usecols=usecols, dtype=np.float32)`` (+ the reference's box mask) on the GPU.

    Returns a float32 CUDA tensor [n_rows, len(usecols)]; columns are views ``out[:, i]`` with stride
    ``len(usecols)``, which the painters take without a copy.  For pandas-style files with a header
    line pass ``skiprows=1``.
    """
    device = require_cuda()
    host, nbytes = _pinned_file(path)
    text = torch.empty(host.numel(), dtype=torch.uint8, device=device)
    text.copy_(host, non_blocking=True)
    return parse_catalog_bytes(text, nbytes, usecols, skiprows=skiprows, comments=comments, box_size=box_size,
                               bounds=bounds, host_bytes=memoryview(host.numpy()), return_info=return_info)
