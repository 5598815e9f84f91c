// Per-mode / per-cell arithmetic of the mock generator (row f-3 of SURVEY.md section 8), shared by the
// device kernels (mockgen.cu) and by a host build used only by the CPU test-suite
// (tests/helpers/mockgen_host.cpp) -- the same pattern as textparse.cuh.
//
// What is restated from the reference:
//   gaussian_field   /root/reference/src/gauss_field.py:5-80   (a transcription of Pylians3)
//   populate_field   /root/reference/src/populate_field.py:11-29, get_positions :4-9
// What is NOT: the random streams.  The reference draws from NumPy's Mersenne Twister in loop
// order (gauss_field.py:20,49-51) or from jax.random (populate_field.py:22-24); a sequential
// stream cannot be consumed by 10^8 independent threads, so every mode / cell / particle here owns
// a COUNTER of the counter-based Philox4x32-10 generator (Salmon et al. 2011): results do not depend
// on the launch geometry and the host build reproduces the device bit for bit in the integer parts.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define JPS_HD __host__ __device__ __forceinline__
#else
#define JPS_HD inline
#endif

namespace jps {
namespace mock {

// ------------------------------------------------------------------ Philox4x32-10
struct U4 { uint32_t x, y, z, w; };

JPS_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
    U4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    n.w = (uint32_t)p0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// stream tags: one key per use so that the draws of a field, of the cell counts and of the in-cell
// offsets never share a (key, counter) pair even for equal seeds
enum Stream : uint32_t { STREAM_FIELD = 0x46494c44u, STREAM_COUNT = 0x434e5473u, STREAM_OFFSET = 0x4f464673u };

JPS_HD U4 draw(uint64_t seed, uint32_t stream, uint64_t index, uint32_t sub) {
  U4 c;
  c.x = (uint32_t)index;
  c.y = (uint32_t)(index >> 32);
  c.z = sub;
  c.w = stream;
  return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// uniform double strictly inside (0, 1): 52 random bits + half a step (never 0, never 1, so the
// reference's "while amplitude == 0: redraw" (gauss_field.py:51) can never fire)
JPS_HD double uniform_open(uint32_t a, uint32_t b) {
  const uint64_t bits = ((uint64_t)(a >> 6) << 26) | (uint64_t)(b >> 6);      // 26 + 26 bits
  return ((double)bits + 0.5) * (1.0 / 4503599627370496.0);                   // 2^-52
}

// uniform float32 in [0, 1) from 24 bits (what jax.random.uniform / np.float32 draws give)
JPS_HD float uniform_f32(uint32_t a) { return (float)(a >> 8) * (1.0f / 16777216.0f); }

// ------------------------------------------------------------------ Gaussian field, one mode
// P(|k|) by linear interpolation in the caller's table, found by the reference's bisection
// (gauss_field.py:38-45): the bracket is [lmin, lmax] with lmax - lmin == 1 at exit, so values
// outside the table are EXTRAPOLATED along the first / last segment, as the reference does.
JPS_HD double interp_power(const double* kf, const double* pkf, int nk, double kmod) {
  int lmin = 0, lmax = nk - 1;
  while (lmax - lmin > 1) {
    const int l = (lmin + lmax) / 2;
    if (kf[l] < kmod) lmin = l; else lmax = l;
  }
  return (pkf[lmax] - pkf[lmin]) / (kf[lmax] - kf[lmin]) * (kmod - kf[lmin]) + pkf[lmin];
}

JPS_HD int freq_of(int i, int n, int mid) { return i > mid ? i - n : i; }          // gauss_field.py:26
JPS_HD int minus_index(int k, int n) { return k > 0 ? n - k : -k; }               // gauss_field.py:27

// Value of delta_k at (ix, iy, iz) of the [n][n][n/2+1] half-space array.
// Reference fill order (gauss_field.py:25-73): modes are visited in C order; on the planes
// kz == 0 and kz == middle the FIRST visited member of a pair (k, -k) keeps its own draw and its
// partner receives the complex conjugate; a self-conjugate mode is real and equal to its amplitude.
// With counters instead of a stream this becomes: use the draw of the lexicographically smaller of
// (ix, iy) and (-ix, -iy), conjugated when this mode is the larger one.
JPS_HD void gaussian_mode(int ix, int iy, int iz, int n, const double* kf, const double* pkf, int nk,
                          int rayleigh, uint64_t seed, double box_size, float& re, float& im) {
  const int mid = n / 2;
  const int kx = freq_of(ix, n, mid), ky = freq_of(iy, n, mid), kz = freq_of(iz, n, mid);
  int cx = ix, cy = iy;
  bool conj = false, self = false;
  if (kz == 0 || kz == mid) {
    const int mx = minus_index(kx, n), my = minus_index(ky, n);
    self = (mx == ix && my == iy);
    if (mx < ix || (mx == ix && my < iy)) { cx = mx; cy = my; conj = true; }
  }
  if (ix == 0 && iy == 0 && iz == 0) { re = 0.0f; im = 0.0f; return; }            // gauss_field.py:76
  const double kmod = sqrt((double)(kx * kx + ky * ky + kz * kz)) * (2.0 * M_PI / box_size);
  const double g2 = (double)n * (double)n / box_size;
  const double pk = interp_power(kf, pkf, nk, kmod) * (g2 * g2 * g2);             // "remove units", :46
  const uint64_t flat = ((uint64_t)cx * (uint64_t)n + (uint64_t)cy) * (uint64_t)(mid + 1) + (uint64_t)iz;
  const U4 r = draw(seed, STREAM_FIELD, flat, 0u);
  const double phase = 2.0 * M_PI * uniform_open(r.x, r.y);
  double amp = uniform_open(r.z, r.w);
  amp = rayleigh ? sqrt(-log(amp)) : 1.0;
  amp *= sqrt(pk);
  if (self) { re = (float)amp; im = 0.0f; return; }                               // gauss_field.py:72-73
  const double c = cos(phase), s = sin(phase);
  re = (float)(amp * c);
  im = (float)(conj ? -(amp * s) : amp * s);
}

// ------------------------------------------------------------------ Poisson counts, one cell
// lam < 12: inversion by sequential search on ONE uniform (exact in float64 for these rates);
// otherwise Hoermann's transformed rejection (PTRS, 1993) -- the algorithm NumPy uses for
// np.random.poisson (populate_field in gauss_field.py:103) -- with two uniforms per trial.
JPS_HD uint32_t poisson_draw(double lam, uint64_t seed, uint64_t cell) {
  if (!(lam > 0.0)) return 0u;                                                    // also NaN
  if (lam < 12.0) {
    const U4 r = draw(seed, STREAM_COUNT, cell, 0u);
    const double u = uniform_open(r.x, r.y);
    double p = exp(-lam), s = p;
    uint32_t k = 0;
    while (u > s && k < 200u) {
      ++k;
      p *= lam / (double)k;
      s += p;
    }
    return k;
  }
  if (lam > 2.0e9) lam = 2.0e9;
  const double slam = sqrt(lam), loglam = log(lam);
  const double b = 0.931 + 2.53 * slam;
  const double a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
  const double vr = 0.9277 - 3.6224 / (b - 2.0);
  for (uint32_t trial = 0; trial < 1000u; ++trial) {
    const U4 r = draw(seed, STREAM_COUNT, cell, trial);
    const double U = uniform_open(r.x, r.y) - 0.5;
    const double V = uniform_open(r.z, r.w);
    const double us = 0.5 - fabs(U);
    const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return (uint32_t)kf;
    if (kf < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0))
      return (uint32_t)kf;
  }
  return (uint32_t)lam;                                                           // unreachable in practice
}

// rate of one cell: rho * mean_obj_per_cell / mean(rho) (populate_field.py:12-16), or the same for
// rho = exp(bias * g) when the Gaussian field itself is passed (tests/create_lognormal.py:49-50)
JPS_HD double cell_density(float v, int lognormal, double bias) {
  return lognormal ? exp(bias * (double)v) : (double)v;
}

// ------------------------------------------------------------------ one particle
// centre + triangular in-cell offset, wrapped into [0, box): populate_field.py:4-9,20,27-29, float32
// as jnp computes it.  `cell` = C-order flat index, `j` = running index of the particle in the cell.
// separately rounded float32 product / sum (jnp evaluates op by op; nvcc would otherwise contract
// a * b + c into one FMA and the device would differ from the host build in the last bit)
JPS_HD float mulf(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
JPS_HD float addf(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}

JPS_HD float tri_offset(uint32_t bits, float bin_size) {
  const float r = addf(mulf(2.0f, uniform_f32(bits)), -1.0f);
  const float a = fabsf(r);
  const float m = addf(1.0f, -sqrtf(a));
  const float sg = (r > 0.0f) ? 1.0f : (r < 0.0f ? -1.0f : 0.0f);
  return mulf(mulf(sg, m), bin_size);
}

JPS_HD float wrap_coord(float c, float box) {
  float v = fmodf(addf(c, box), box);
  if (v < 0.0f) v += box;                         // jnp's % is a floor-mod
  if (v >= box) v = 0.0f;
  return v;
}

JPS_HD void particle_position(int ix, int iy, int iz, uint64_t particle, uint64_t seed, float bin_size,
                              float box, float& x, float& y, float& z) {
  const U4 r = draw(seed, STREAM_OFFSET, particle, 0u);
  const float half = mulf(0.5f, bin_size);
  x = wrap_coord(addf(addf(mulf((float)ix, bin_size), half), tri_offset(r.x, bin_size)), box);
  y = wrap_coord(addf(addf(mulf((float)iy, bin_size), half), tri_offset(r.y, bin_size)), box);
  z = wrap_coord(addf(addf(mulf((float)iz, bin_size), half), tri_offset(r.z, bin_size)), box);
}

}  // namespace mock
}  // namespace jps
