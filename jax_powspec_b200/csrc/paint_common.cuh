// Per-particle stencil arithmetic shared by the painters.
//
// compat=reference, order 2 follows /root/reference/src/mas.py:100-151 (cic_mas_vec) and
// :40-81 (cic_mas) including quirks Q1-Q4 of SURVEY.md section 8; everything else is the
// textbook B-spline of that order on integer nodes (absent from the reference).
#pragma once

#include "common.cuh"

namespace jps {

struct PaintParams {
  int n;                   // global mesh side
  int x0;                  // global x-plane stored at local plane 0 (0 for a full mesh; may be < 0)
  int nx;                  // allocated x-planes of `mesh` (n for a full mesh; slab + ghosts otherwise)
  int wrap;
  int variant;
  float xmin, ymin, zmin;
  float inv;               // 1 / (box_size / n), float32 (Q4)
  int64_t stride;
  int64_t n_part;
  const float* x;
  const float* y;
  const float* z;
  const float* w;          // may be null
  float* mesh;
};

// pos = (x - xmin) * inv_bin_size, rounded to float32 after each operation exactly as the
// reference's materialised arrays are (src/mas.py:103-105, Q4).  The intrinsics stop nvcc from
// contracting the product into a later `pos - floor(pos)` FMA, which would compute the in-cell
// offset from the UNROUNDED product and shift weights by up to half an ulp of pos (measured:
// 3e-5 relative on a 512^3 mesh).
__device__ __forceinline__ float grid_pos(float x, float xmin, float inv) {
  return __fmul_rn(__fsub_rn(x, xmin), inv);
}

// Python-style a mod n (result in [0, n)).  Integer division by a run-time n costs ~30
// instructions and was the dominant instruction cost of the bucketing passes; particles within one
// box length of the box only need a conditional add and a conditional subtract (branch-free
// selects), and the division lives out of line so that the hot loops stay small.
static __device__ __noinline__ int pymod_slow(int a, int n) {
  const int r = a % n;
  return r < 0 ? r + n : r;
}

__device__ __forceinline__ int wrap_once(int a, int n) {
  a += (a < 0) ? n : 0;
  a -= (a >= n) ? n : 0;
  return a;                                        // in [0, n) iff the input was in [-n, 2n)
}

__device__ __forceinline__ int pymod(int a, int n) {
  a = wrap_once(a, n);
  if ((unsigned)a >= (unsigned)n) a = pymod_slow(a, n);
  return a;
}

// Global (already wrapped / validated, or -1) x-plane -> local plane of a [nx][n][n] slab mesh
// that starts at global plane x0; -1 if the plane is not held locally.  Identity for a full mesh.
__device__ __forceinline__ int local_plane(int gx, int x0, int nx, int n) {
  if (gx < 0) return -1;
  const int l = pymod(gx - x0, n);
  return l < nx ? l : -1;
}


// JAX .at[] scatter index: negatives wrap once, what is still out of range is dropped (-1).
__device__ __forceinline__ int scatter_norm(int i, int n) {
  if (i < 0) i += n;
  return (i >= 0 && i < n) ? i : -1;
}

// Reference CIC, one axis: base index i0, "+1" index i1 (both already in scatter-normalised
// form, -1 = dropped), md = 1-dd and dd.
__device__ __forceinline__ void cic_reference_axis(float pos, int n, int wrap, int variant,
                                                   int& i0, int& i1, float& md, float& dd) {
  int i = (int)pos;                   // cvt.rzi: truncation toward zero (Q3)
  dd = pos - (float)i;
  int ip = i + 1;
  md = 1.0f - dd;                     // both variants: from the un-modified dd
  if (variant == JPS_VARIANT_VEC) {
    if (wrap) ip = pymod(ip + n, n);
    else if (ip >= n) ip = 0;         // dd is NOT zeroed (Q2)
  } else {
    if (ip >= n) {
      if (wrap) ip -= n;
      else { ip = 0; dd = 0.0f; }
    }
  }
  i0 = scatter_norm(i, n);
  i1 = scatter_norm(ip, n);
}

// Textbook B-spline of order ORDER (2 CIC, 3 TSC, 4 PCS) on integer nodes.
// idx[s] = -1 when the contribution is dropped (wrap == 0 and node outside the mesh).
template <int ORDER>
__device__ __forceinline__ void bspline_axis(float pos, int n, int wrap, int (&idx)[ORDER],
                                             float (&w)[ORDER]) {
  int base;
  if (ORDER == 2) {
    float f = floorf(pos);
    float d = pos - f;
    w[0] = 1.0f - d;
    w[1] = d;
    base = (int)f;
  } else if (ORDER == 3) {
    float f = floorf(pos + 0.5f);
    float d = pos - f;
    float a = 0.5f - d, b = 0.5f + d;
    w[0] = 0.5f * a * a;
    w[1] = 0.75f - d * d;
    w[2] = 0.5f * b * b;
    base = (int)f - 1;
  } else {
    float f = floorf(pos);
    float d = pos - f;
    float e = 1.0f - d;
    const float sixth = 1.0f / 6.0f;
    w[0] = e * e * e * sixth;
    w[1] = (4.0f - 6.0f * d * d + 3.0f * d * d * d) * sixth;
    w[2] = (4.0f - 6.0f * e * e + 3.0f * e * e * e) * sixth;
    w[ORDER - 1] = d * d * d * sixth;
    base = (int)f - 1;
  }
#pragma unroll
  for (int s = 0; s < ORDER; ++s) {
    int j = base + s;
    if (wrap) idx[s] = pymod(j, n);
    else idx[s] = (j >= 0 && j < n) ? j : -1;
  }
}

}  // namespace jps
