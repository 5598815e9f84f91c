/* CPU restatement of the jax-powspec hot path in plain C -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Serial, float32, in the order the reference's XLA-CPU program runs (one scatter after the
 * other, one histogram pass), so it doubles as the CPU baseline bench.py reports
 * (cpu_baseline.kind = "port").  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may load it.  Follows:
 *   jpso_paint_cic_reference  /root/reference/src/mas.py:100-151 (vec) and :40-81 (scan), Q1-Q6
 *   jpso_paint_bspline        textbook CIC/TSC/PCS on integer nodes (absent from the reference)
 *   jpso_pk_bin               /root/reference/src/correlations.py:25-48 (window, |dk|^2, mu, histograms)
 *   jpso_paint_f64            oracle/mas.py paint(precision="f64") for catalogues NumPy cannot hold in
 *                             seconds: float32 cell choice / in-cell offsets (mas.py:100-117), weights,
 *                             products and sums in float64; OpenMP over particles (order of float64
 *                             additions is irrelevant at the 1e-6 tolerances it is used for)
 *   jpso_pk_bin_f64           oracle/correlations.py powspec(precision="f64") binning of a complex128
 *                             spectrum: float32 bin decisions, float64 values
 *   jpso_paint_bspline_mt     best-effort multi-core CPU painter (float32, atomic adds), bench.py's
 *                             "cpu_multicore" leg only
 * Build: oracle/build.py (gcc -O2 -fopenmp -shared -fPIC).
 */
#ifdef _OPENMP
#include <omp.h>
#endif
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int pymod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }
static int scatter_norm(int i, int n) { if (i < 0) i += n; return (i >= 0 && i < n) ? i : -1; }

/* variant: 0 = cic_mas_vec (corner-major: 8 scatters over all particles), 1 = cic_mas (scan) */
int jpso_paint_cic_reference(float* mesh, const float* x, const float* y, const float* z,
                             const float* w, int64_t np, float xmin, float ymin, float zmin,
                             float box_size, int n, int wrap, int variant) {
  const float bin_size = box_size / (float)n;
  const float inv = 1.0f / bin_size;
  const size_t n2 = (size_t)n * n;
  /* the vectorised reference materialises the index / weight arrays first */
  int* idx = (int*)malloc(sizeof(int) * 6 * (size_t)(np ? np : 1));
  float* wt = (float*)malloc(sizeof(float) * 6 * (size_t)(np ? np : 1));
  if (!idx || !wt) { free(idx); free(wt); return -1; }
  const float* pos[3] = {x, y, z};
  const float mins[3] = {xmin, ymin, zmin};
  for (int a = 0; a < 3; ++a) {
    int* i0 = idx + (size_t)(2 * a) * np; int* i1 = idx + (size_t)(2 * a + 1) * np;
    float* md = wt + (size_t)(2 * a) * np; float* dd = wt + (size_t)(2 * a + 1) * np;
    for (int64_t p = 0; p < np; ++p) {
      const float g = (pos[a][p] - mins[a]) * inv;
      int i = (int)g;                       /* truncation toward zero (Q3) */
      float d = g - (float)i;
      int ip = i + 1;
      float m = 1.0f - d;
      if (variant == 0) {
        if (wrap) ip = pymod(ip + n, n); else if (ip >= n) ip = 0;     /* Q2 */
      } else if (ip >= n) {
        if (wrap) ip -= n; else { ip = 0; d = 0.0f; }
      }
      i0[p] = scatter_norm(i, n); i1[p] = scatter_norm(ip, n); md[p] = m; dd[p] = d;
    }
  }
  /* corner table of mas.py:142-151: which index (0 = i, 1 = ip) and which factor (0 = md, 1 = dd) */
  static const int cidx[8][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,0,1},{0,1,1},{1,1,1}};
  static const int cwgt[8][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,0,1},{0,0,1},{1,1,1}}; /* row 6: Q1 */
  if (variant == 0) {
    for (int c = 0; c < 8; ++c) {
      const int* ix = idx + (size_t)(0 + cidx[c][0]) * np; const int* iy = idx + (size_t)(2 + cidx[c][1]) * np;
      const int* iz = idx + (size_t)(4 + cidx[c][2]) * np;
      const float* fx = wt + (size_t)(0 + cwgt[c][0]) * np; const float* fy = wt + (size_t)(2 + cwgt[c][1]) * np;
      const float* fz = wt + (size_t)(4 + cwgt[c][2]) * np;
      for (int64_t p = 0; p < np; ++p) {
        if ((ix[p] | iy[p] | iz[p]) < 0) continue;
        mesh[(size_t)ix[p] * n2 + (size_t)iy[p] * n + iz[p]] += ((fx[p] * fy[p]) * fz[p]) * (w ? w[p] : 1.0f);
      }
    }
  } else {
    for (int64_t p = 0; p < np; ++p)
      for (int c = 0; c < 8; ++c) {
        const int ix = idx[(size_t)(0 + cidx[c][0]) * np + p], iy = idx[(size_t)(2 + cidx[c][1]) * np + p],
                  iz = idx[(size_t)(4 + cidx[c][2]) * np + p];
        if ((ix | iy | iz) < 0) continue;
        mesh[(size_t)ix * n2 + (size_t)iy * n + iz] +=
            ((wt[(size_t)(0 + cwgt[c][0]) * np + p] * wt[(size_t)(2 + cwgt[c][1]) * np + p]) *
             wt[(size_t)(4 + cwgt[c][2]) * np + p]) * (w ? w[p] : 1.0f);
      }
  }
  free(idx); free(wt);
  return 0;
}

static void bspline_axis(float pos, int order, int n, int wrap, int* idx, float* w) {
  int base;
  if (order == 2) {
    float f = floorf(pos), d = pos - f; w[0] = 1.0f - d; w[1] = d; base = (int)f;
  } else if (order == 3) {
    float f = floorf(pos + 0.5f), d = pos - f, a = 0.5f - d, b = 0.5f + d;
    w[0] = 0.5f * a * a; w[1] = 0.75f - d * d; w[2] = 0.5f * b * b; base = (int)f - 1;
  } else {
    float f = floorf(pos), d = pos - f, e = 1.0f - d; const float s = 1.0f / 6.0f;
    w[0] = e * e * e * s; w[1] = (4.0f - 6.0f * d * d + 3.0f * d * d * d) * s;
    w[2] = (4.0f - 6.0f * e * e + 3.0f * e * e * e) * s; w[3] = d * d * d * s; base = (int)f - 1;
  }
  for (int s = 0; s < order; ++s) {
    int j = base + s;
    idx[s] = wrap ? pymod(j, n) : ((j >= 0 && j < n) ? j : -1);
  }
}

int jpso_paint_bspline(float* mesh, const float* x, const float* y, const float* z, const float* w,
                       int64_t np, float xmin, float ymin, float zmin, float box_size, int n,
                       int wrap, int order) {
  if (order < 2 || order > 4) return -1;
  const float bin_size = box_size / (float)n;
  const float inv = 1.0f / bin_size;
  const size_t n2 = (size_t)n * n;
  for (int64_t p = 0; p < np; ++p) {
    int ix[4], iy[4], iz[4]; float wx[4], wy[4], wz[4];
    bspline_axis((x[p] - xmin) * inv, order, n, wrap, ix, wx);
    bspline_axis((y[p] - ymin) * inv, order, n, wrap, iy, wy);
    bspline_axis((z[p] - zmin) * inv, order, n, wrap, iz, wz);
    const float wp = w ? w[p] : 1.0f;
    for (int a = 0; a < order; ++a) for (int b = 0; b < order; ++b) {
      if ((ix[a] | iy[b]) < 0) continue;
      float* row = mesh + (size_t)ix[a] * n2 + (size_t)iy[b] * n;
      const float wxy = wx[a] * wy[b];
      for (int c = 0; c < order; ++c) if (iz[c] >= 0) row[iz[c]] += (wxy * wz[c]) * wp;
    }
  }
  return 0;
}

/* window factor per axis, float32 op by op as correlations.py:15,20-21 */
static void window_axis(int n, int p, float* out) {
  const float pref = (float)(M_PI / (double)n), pi32 = (float)M_PI;
  for (int i = 0; i < n; ++i) {
    const int ki = i > n / 2 ? i - n : i;
    const float xx = pref * (float)ki, yy = xx / pi32;
    float s = 1.0f;
    if (yy != 0.0f) { const float pix = pi32 * yy; s = sinf(pix) / pix; }
    const float r = 1.0f / s; float v = r;
    for (int j = 1; j < p; ++j) v = v * r;
    out[i] = v;
  }
}

/* dk: complex64 [n][n][n/2+1] interleaved (NOT yet deconvolved).  kedges: grid units, nb+1.
 * out: s0,s2,s4,cnt float32[nb], accumulated serially in C order like XLA-CPU's scatter. */
int jpso_pk_bin(const float* dk, int n, const float* kedges, int nb, int mas_order,
                float* s0, float* s2, float* s4, float* cnt) {
  const int nz = n / 2 + 1, mid = n / 2;
  float* wl = (float*)malloc(sizeof(float) * (size_t)n);
  if (!wl) return -1;
  window_axis(n, mas_order, wl);
  memset(s0, 0, sizeof(float) * nb); memset(s2, 0, sizeof(float) * nb);
  memset(s4, 0, sizeof(float) * nb); memset(cnt, 0, sizeof(float) * nb);
  for (int ix = 0; ix < n; ++ix) {
    const int kx = ix > mid ? ix - n : ix;
    for (int iy = 0; iy < n; ++iy) {
      const int ky = iy > mid ? iy - n : iy;
      const float cxy = wl[ix] * wl[iy];
      const float* row = dk + 2 * ((size_t)ix * n + iy) * nz;
      for (int kz = 0; kz < nz; ++kz) {
        const float k = sqrtf((float)(kx * kx + ky * ky + kz * kz));
        /* searchsorted(kedges, k, 'right') */
        int lo = 0, hi = nb + 1;
        while (lo < hi) { int m = (lo + hi) / 2; if (kedges[m] <= k) lo = m + 1; else hi = m; }
        int idx = lo;
        if (k == kedges[nb]) idx = nb;
        if (idx < 1 || idx > nb) continue;
        const float c = cxy * wl[kz];
        const float re = row[2 * kz] * c, im = row[2 * kz + 1] * c;
        const float d2 = re * re + im * im;
        const float mu = (k == 0.0f) ? 0.0f : (float)kz / k;
        const float mu2 = mu * mu;
        s0[idx - 1] += d2;
        s2[idx - 1] += d2 * (3.0f * mu2 - 1.0f) / 2.0f;
        s4[idx - 1] += d2 * (35.0f * mu2 * mu2 - 30.0f * mu2 + 3.0f) / 8.0f;
        cnt[idx - 1] += 1.0f;
      }
    }
  }
  free(wl);
  return 0;
}


/* ------------------------------------------------------------------------------------------
 * float64 restatement (oracle/mas.py precision="f64") for full-size catalogues. */
static void bspline_axis_f64(float pos, int order, int n, int wrap, int* idx, double* w) {
  int base;
  if (order == 2) {
    float f = floorf(pos); double d = (double)(pos - f); w[0] = 1.0 - d; w[1] = d; base = (int)f;
  } else if (order == 3) {
    float f = floorf(pos + 0.5f); double d = (double)(pos - f);
    w[0] = 0.5 * (0.5 - d) * (0.5 - d); w[1] = 0.75 - d * d; w[2] = 0.5 * (0.5 + d) * (0.5 + d); base = (int)f - 1;
  } else {
    float f = floorf(pos); double d = (double)(pos - f), e = 1.0 - d; const double s = 1.0 / 6.0;
    w[0] = e * e * e * s; w[1] = (4.0 - 6.0 * d * d + 3.0 * d * d * d) * s;
    w[2] = (4.0 - 6.0 * e * e + 3.0 * e * e * e) * s; w[3] = d * d * d * s; base = (int)f - 1;
  }
  for (int s = 0; s < order; ++s) {
    int j = base + s;
    idx[s] = wrap ? pymod(j, n) : ((j >= 0 && j < n) ? j : -1);
  }
}

/* compat_ref != 0 (order 2 only): the reference's CIC with quirks Q1-Q4, variant as above */
int jpso_paint_f64(double* mesh, const float* x, const float* y, const float* z, const float* w,
                   int64_t np, float xmin, float ymin, float zmin, float box_size, int n, int wrap,
                   int order, int compat_ref, int variant) {
  if (order < 2 || order > 4 || (compat_ref && order != 2)) return -1;
  const float bin_size = box_size / (float)n;
  const float inv = 1.0f / bin_size;
  const size_t n2 = (size_t)n * n;
  static const int cidx[8][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,0,1},{0,1,1},{1,1,1}};
  static const int cwgt[8][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,0,1},{0,0,1},{1,1,1}}; /* row 6: Q1 */
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < np; ++p) {
    const double wp = w ? (double)w[p] : 1.0;
    const float g[3] = {(x[p] - xmin) * inv, (y[p] - ymin) * inv, (z[p] - zmin) * inv};
    if (compat_ref) {
      int id[3][2]; double wt[3][2];
      for (int a = 0; a < 3; ++a) {
        int i = (int)g[a]; float d = g[a] - (float)i; int ip = i + 1; float m = 1.0f - d;
        if (variant == 0) { if (wrap) ip = pymod(ip + n, n); else if (ip >= n) ip = 0; }
        else if (ip >= n) { if (wrap) ip -= n; else { ip = 0; d = 0.0f; } }
        id[a][0] = scatter_norm(i, n); id[a][1] = scatter_norm(ip, n); wt[a][0] = (double)m; wt[a][1] = (double)d;
      }
      for (int c = 0; c < 8; ++c) {
        const int ix = id[0][cidx[c][0]], iy = id[1][cidx[c][1]], iz = id[2][cidx[c][2]];
        if ((ix | iy | iz) < 0) continue;
        const double v = wt[0][cwgt[c][0]] * wt[1][cwgt[c][1]] * wt[2][cwgt[c][2]] * wp;
#pragma omp atomic
        mesh[(size_t)ix * n2 + (size_t)iy * n + iz] += v;
      }
    } else {
      int ix[4], iy[4], iz[4]; double wx[4], wy[4], wz[4];
      bspline_axis_f64(g[0], order, n, wrap, ix, wx);
      bspline_axis_f64(g[1], order, n, wrap, iy, wy);
      bspline_axis_f64(g[2], order, n, wrap, iz, wz);
      for (int a = 0; a < order; ++a) for (int b = 0; b < order; ++b) {
        if ((ix[a] | iy[b]) < 0) continue;
        double* row = mesh + (size_t)ix[a] * n2 + (size_t)iy[b] * n;
        const double wxy = wx[a] * wy[b];
        for (int c = 0; c < order; ++c) if (iz[c] >= 0) {
          const double v = (wxy * wz[c]) * wp;
#pragma omp atomic
          row[iz[c]] += v;
        }
      }
    }
  }
  return 0;
}

/* dk: complex128 [n][n][n/2+1] interleaved, NOT deconvolved.  float32 bin decisions as jpso_pk_bin,
 * float64 window ((1/sinc(pi k/N))^p), |dk|^2, mu^2 and sums (oracle/correlations.py "f64"). */
int jpso_pk_bin_f64(const double* dk, int n, const float* kedges, int nb, int mas_order,
                    double* s0, double* s2, double* s4, int64_t* cnt) {
  const int nz = n / 2 + 1, mid = n / 2;
  double* wl = (double*)malloc(sizeof(double) * (size_t)n);
  if (!wl) return -1;
  for (int i = 0; i < n; ++i) {
    const int ki = i > mid ? i - n : i;
    double s = 1.0;
    if (ki != 0) { const double xx = M_PI * (double)ki / (double)n; s = sin(xx) / xx; }
    wl[i] = pow(1.0 / s, (double)mas_order);
  }
  memset(s0, 0, sizeof(double) * nb); memset(s2, 0, sizeof(double) * nb);
  memset(s4, 0, sizeof(double) * nb); memset(cnt, 0, sizeof(int64_t) * nb);
  int fail = 0;
#pragma omp parallel
  {
    double* t0 = (double*)calloc((size_t)nb * 3, sizeof(double));
    int64_t* tc = (int64_t*)calloc((size_t)nb, sizeof(int64_t));
    if (!t0 || !tc) {
#pragma omp atomic write
      fail = 1;
    } else {
#pragma omp for schedule(static)
      for (int ix = 0; ix < n; ++ix) {
        const int kx = ix > mid ? ix - n : ix;
        for (int iy = 0; iy < n; ++iy) {
          const int ky = iy > mid ? iy - n : iy;
          const double cxy = wl[ix] * wl[iy];
          const double* row = dk + 2 * ((size_t)ix * n + iy) * nz;
          for (int kz = 0; kz < nz; ++kz) {
            const int k2 = kx * kx + ky * ky + kz * kz;
            const float k = sqrtf((float)k2);
            int lo = 0, hi = nb + 1;
            while (lo < hi) { int m = (lo + hi) / 2; if (kedges[m] <= k) lo = m + 1; else hi = m; }
            int idx = lo;
            if (k == kedges[nb]) idx = nb;
            if (idx < 1 || idx > nb) continue;
            const double c = cxy * wl[kz];
            const double re = row[2 * kz] * c, im = row[2 * kz + 1] * c;
            const double d2 = re * re + im * im;
            double mu2 = 0.0;
            if (k2 > 0) { const double mu = (double)kz / sqrt((double)k2); mu2 = mu * mu; }
            t0[(size_t)(idx - 1) * 3 + 0] += d2;
            t0[(size_t)(idx - 1) * 3 + 1] += d2 * (3.0 * mu2 - 1.0) / 2.0;
            t0[(size_t)(idx - 1) * 3 + 2] += d2 * (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0;
            tc[idx - 1] += 1;
          }
        }
      }
#pragma omp critical
      for (int b = 0; b < nb; ++b) {
        s0[b] += t0[(size_t)b * 3]; s2[b] += t0[(size_t)b * 3 + 1]; s4[b] += t0[(size_t)b * 3 + 2]; cnt[b] += tc[b];
      }
    }
    free(t0); free(tc);
  }
  free(wl);
  return fail ? -1 : 0;
}

/* Best-effort multi-core painter (bench.py "cpu_multicore" leg): float32, particles split over
 * the OpenMP threads, atomic adds into the shared mesh. */
int jpso_paint_bspline_mt(float* mesh, const float* x, const float* y, const float* z, const float* w,
                          int64_t np, float xmin, float ymin, float zmin, float box_size, int n,
                          int wrap, int order) {
  if (order < 2 || order > 4) return -1;
  const float bin_size = box_size / (float)n;
  const float inv = 1.0f / bin_size;
  const size_t n2 = (size_t)n * n;
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < np; ++p) {
    int ix[4], iy[4], iz[4]; float wx[4], wy[4], wz[4];
    bspline_axis((x[p] - xmin) * inv, order, n, wrap, ix, wx);
    bspline_axis((y[p] - ymin) * inv, order, n, wrap, iy, wy);
    bspline_axis((z[p] - zmin) * inv, order, n, wrap, iz, wz);
    const float wp = w ? w[p] : 1.0f;
    for (int a = 0; a < order; ++a) for (int b = 0; b < order; ++b) {
      if ((ix[a] | iy[b]) < 0) continue;
      float* row = mesh + (size_t)ix[a] * n2 + (size_t)iy[b] * n;
      const float wxy = wx[a] * wy[b];
      for (int c = 0; c < order; ++c) if (iz[c] >= 0) {
        const float v = (wxy * wz[c]) * wp;
#pragma omp atomic
        row[iz[c]] += v;
      }
    }
  }
  return 0;
}

int jpso_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
