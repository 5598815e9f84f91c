"""Shared helpers for the GPU parity tests (seeded synthetic catalogues)."""
import numpy as np

F32 = np.float32


def clustered_particles(seed, n_part, box, n_blobs=40, frac_uniform=0.5):
    rng = np.random.default_rng(seed)
    nu = int(n_part * frac_uniform)
    uni = rng.random((nu, 3)) * box
    centres = rng.random((n_blobs, 3)) * box
    which = rng.integers(0, n_blobs, n_part - nu)
    sig = np.array([0.02, 0.02, 0.05]) * box
    blob = centres[which] + rng.normal(size=(n_part - nu, 3)) * sig
    p = (np.concatenate([uni, blob]) % box).astype(F32)
    p[p >= F32(box)] = 0.0
    rng.shuffle(p, axis=0)
    return p


def rel_to_monopole(pk, ref):
    """|pk - ref| / |ref P0| per bin (the normalisation BASELINE.md section 5 prescribes)."""
    scale = np.abs(ref[:, :1])
    return np.abs(pk - ref) / scale
