// Device helpers shared by the Fourier-space kernels: canonical (|kx|,|ky|) pair enumeration,
// the rows of the half-space array that map onto a pair, and the warp segmented reduction.
#pragma once

#include "common.cuh"

namespace jps {

int npairs_for(int n);        // number of canonical pairs a <= b, 0 <= a,b <= n/2

struct PairDecode {
  int a, b;
};

__host__ __device__ inline PairDecode decode_pair(int p) {
  int b = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
  while ((long long)(b + 1) * (b + 2) / 2 <= p) ++b;
  while ((long long)b * (b + 1) / 2 > p) --b;
  PairDecode d;
  d.b = b;
  d.a = p - (int)((long long)b * (b + 1) / 2);
  return d;
}

// rows (ix,iy) of the half-space array whose (|kx|,|ky|) is {a,b} in either order
struct RowSet {
  int nrows;
  int ix[8], iy[8];
};

__host__ __device__ inline RowSet make_rows(int a, int b, int n) {
  RowSet r;
  const int na = (a > 0 && 2 * a != n) ? 2 : 1;
  const int nbb = (b > 0 && 2 * b != n) ? 2 : 1;
  const int ia[2] = {a, n - a};
  const int ib[2] = {b, n - b};
  r.nrows = 0;
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nbb; ++j) {
      r.ix[r.nrows] = ia[i]; r.iy[r.nrows] = ib[j]; ++r.nrows;
    }
  if (a != b) {
    for (int i = 0; i < na; ++i)
      for (int j = 0; j < nbb; ++j) {
        r.ix[r.nrows] = ib[j]; r.iy[r.nrows] = ia[i]; ++r.nrows;
      }
  }
  for (int q = r.nrows; q < 8; ++q) { r.ix[q] = 0; r.iy[q] = 0; }
  return r;
}

// ------------------------------------------------------------------ device helpers
// Segmented suffix sums over lanes; `heads` has a bit set for every lane that starts a segment.
// After the call every head lane holds the sum of its segment.
template <int NV>
__device__ __forceinline__ void segmented_reduce(float (&v)[NV], unsigned heads, int lane) {
  const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const bool ok = (lane + off < 32) && ((above & ((1u << off) - 1u)) == 0u);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float t = __shfl_down_sync(0xffffffffu, v[j], off);
      if (ok) v[j] += t;
    }
  }
}

}  // namespace jps
