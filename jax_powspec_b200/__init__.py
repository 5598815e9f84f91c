"""jax_powspec_b200 -- B200 (sm_100a) implementation of jax-powspec's mesh-painting and
Fourier-space clustering hot path behind the reference's own function names.

    from jax_powspec_b200.mas import cic_mas_vec
    from jax_powspec_b200.correlations import powspec_vec

Importing the package loads libjps.so and raises if it is missing (no CPU fallback).
"""
from . import _lib  # noqa: F401  (loads the CUDA library, fails loudly if absent)
from .correlations import (HostPipeline, PaintPowspec, bispec, bispec_pairs, compute_2pt_correlations,
                           compute_all_correlations, paint_powspec, powspec_vec, powspec_vec_fundamental,
                           s_edges_conv, triangle_pairs, xi_vec, xi_vec_coords, xi_vec_fundamental)
from .mas import cic_mas, cic_mas_vec, paint, paint_interlaced, pcs_mas_vec, tsc_mas_vec
from .reader import parse_catalog_bytes, read_catalog_text

__all__ = [
    "cic_mas", "cic_mas_vec", "tsc_mas_vec", "pcs_mas_vec", "paint", "paint_interlaced",
    "powspec_vec", "powspec_vec_fundamental", "xi_vec", "xi_vec_fundamental", "xi_vec_coords",
    "s_edges_conv", "bispec", "bispec_pairs", "triangle_pairs", "compute_2pt_correlations", "compute_all_correlations", "paint_powspec", "PaintPowspec", "HostPipeline",
    "read_catalog_text", "parse_catalog_bytes",
]
