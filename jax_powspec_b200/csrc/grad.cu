// Gradients: the reference exists to be differentiated
// (jax.value_and_grad through cic_mas_vec -> powspec_vec, e.g. tests/lognormal.py:99-107).
//   jps_powspec_grad : cotangent of Pk3D  ->  cotangent of the input mesh
//   jps_paint_grad   : cotangent of the mesh -> cotangents of x, y, z, w
// Adjoint of the binning: every stored mode k gets the real factor
//   g_k = sum_l gS_l[bin(k)] L_l(mu_k) C_k^2 s^2,   gS_l[b] = gPk[b,l] (2l+1) V / Nmodes[b]
// and d/d rho_x sum_k g_k |rho_k|^2 = Re sum_{k stored} 2 g_k rho_k e^{+ikx}: one unnormalised C2R
// of H_k = g_k rho_k (x2 on the kz = 0 and kz = N/2 planes, whose conjugates are stored explicitly and
// counted as separate modes by the reference, Q7).  With normalise (delta = rho/mean - 1 folded in
// through s = N^3/rho_0) the DC mode adds the constant -2/rho_0 sum_{b,l} gS_l[b] S_l[b].
// Adjoint of the deposit: a gather of the mesh cotangent over the particle's stencil with the
// B-spline weights (d/dw) and their derivatives (d/dx, times w and 1/cell).
#include "paint_common.cuh"
#include "fold.cuh"

#include <algorithm>
#include <cmath>

namespace jps {

// gS[c*3 + l] for compact bin c, and the DC constant (in dcterm[0]) when normalise
__global__ void pk_grad_coeff_kernel(int nb, const int32_t* __restrict__ bin_to_compact,
                                     const unsigned long long* __restrict__ cnt,
                                     const double* __restrict__ acc, const float* __restrict__ grad_pk,
                                     double vol, int normalise, const float2* __restrict__ dk,
                                     float* __restrict__ gS, double* __restrict__ dcterm) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb) return;
  const int c = bin_to_compact[j];
  if (c < 0) return;
  const double nm = (double)(float)cnt[c];
  const double mult[3] = {1.0, 5.0, 9.0};
  double dot = 0.0;
  for (int l = 0; l < 3; ++l) {
    const double g = nm > 0.0 ? (double)grad_pk[j * 3 + l] * mult[l] * vol / nm : 0.0;
    gS[c * 3 + l] = (float)g;
    dot += g * acc[(size_t)c * 4 + l];
  }
  if (normalise) atomicAdd(dcterm, -2.0 * dot / (double)dk[0].x);
}

// H_k = fac * g_k * rho_k, written in place over delta_k
__global__ void __launch_bounds__(256) pk_grad_modes_kernel(float2* __restrict__ dk, int n, int nz,
                                                            int pitch, const int32_t* __restrict__ lut,
                                                            const float* __restrict__ wl,
                                                            const float* __restrict__ gS, int normalise) {
  float scale2 = 1.0f;
  if (normalise) {
    const double s = (double)n * (double)n * (double)n / (double)dk[0].x;
    scale2 = (float)(s * s);
  }
  __syncthreads();                                  // every thread has read dk[0] before it is overwritten below
  const int mid = n / 2;
  const long long rows = (long long)n * n;
  // row 0 holds the DC mode every CTA needs: it is processed last, by CTA 0, after a grid-wide
  // hand-shake is unnecessary because scale2 was captured above by all threads of all CTAs only if
  // they started -- so row 0 is excluded here and handled by a second launch (first_row = 0).
  for (long long row = blockIdx.x + 1; row < rows; row += gridDim.x) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const int kx = ix > mid ? ix - n : ix, ky = iy > mid ? iy - n : iy;
    const int k2xy = kx * kx + ky * ky;
    const float wxy = wl[ix] * wl[iy];
    float2* r = dk + (size_t)row * pitch;
    for (int kz = threadIdx.x; kz < nz; kz += blockDim.x) {
      const int k2 = k2xy + kz * kz;
      const int cb = lut[k2];
      float2 h = make_float2(0.0f, 0.0f);
      if (cb >= 0) {
        const float c = wxy * wl[kz];
        const float mu2 = k2 > 0 ? (float)(kz * kz) / (float)k2 : 0.0f;
        const float g = (gS[cb * 3] + gS[cb * 3 + 1] * (3.0f * mu2 - 1.0f) * 0.5f +
                         gS[cb * 3 + 2] * (35.0f * mu2 * mu2 - 30.0f * mu2 + 3.0f) * 0.125f) * (c * c) * scale2;
        const float fac = (kz == 0 || 2 * kz == n) ? 2.0f : 1.0f;
        const float2 d = r[kz];
        h = make_float2(fac * g * d.x, fac * g * d.y);
      }
      r[kz] = h;
    }
  }
}

// same for row 0 (ix = iy = 0), launched after the kernel above with the DC value passed by value
__global__ void pk_grad_row0_kernel(float2* __restrict__ dk, int n, int nz, const int32_t* __restrict__ lut,
                                    const float* __restrict__ wl, const float* __restrict__ gS,
                                    int normalise, const float* __restrict__ dc_saved) {
  float scale2 = 1.0f;
  if (normalise) {
    const double s = (double)n * (double)n * (double)n / (double)dc_saved[0];
    scale2 = (float)(s * s);
  }
  const float wxy = wl[0] * wl[0];
  for (int kz = threadIdx.x; kz < nz; kz += blockDim.x) {
    const int k2 = kz * kz;
    const int cb = lut[k2];
    float2 h = make_float2(0.0f, 0.0f);
    if (cb >= 0 && !(normalise && k2 == 0)) {
      const float c = wxy * wl[kz];
      const float mu2 = k2 > 0 ? 1.0f : 0.0f;
      const float g = (gS[cb * 3] + gS[cb * 3 + 1] * (3.0f * mu2 - 1.0f) * 0.5f +
                       gS[cb * 3 + 2] * (35.0f * mu2 * mu2 - 30.0f * mu2 + 3.0f) * 0.125f) * (c * c) * scale2;
      const float fac = (kz == 0 || 2 * kz == n) ? 2.0f : 1.0f;
      const float2 d = dk[kz];
      h = make_float2(fac * g * d.x, fac * g * d.y);
    }
    dk[kz] = h;
  }
}

__global__ void save_dc_kernel(const float2* __restrict__ dk, float* __restrict__ dc_saved) { dc_saved[0] = dk[0].x; }

// grad_mesh[x] = field[x] (padded rows) + dcterm
__global__ void __launch_bounds__(256) unpad_add_kernel(const float* __restrict__ field, int n, int rowpitch,
                                                        const double* __restrict__ dcterm,
                                                        float* __restrict__ out) {
  const float add = (float)dcterm[0];
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x)
    for (int j = threadIdx.x; j < n; j += blockDim.x)
      out[(size_t)row * n + j] = field[(size_t)row * rowpitch + j] + add;
}

// ---------------------------------------------------------------- deposit adjoint
template <int ORDER>
__device__ __forceinline__ void bspline_axis_grad(float pos, int n, int wrap, int (&idx)[ORDER],
                                                  float (&w)[ORDER], float (&dw)[ORDER]) {
  bspline_axis<ORDER>(pos, n, wrap, idx, w);
  if (ORDER == 2) {
    dw[0] = -1.0f; dw[1] = 1.0f;
  } else if (ORDER == 3) {
    const float d = pos - floorf(pos + 0.5f);
    dw[0] = -(0.5f - d); dw[1] = -2.0f * d; dw[2] = 0.5f + d;
  } else {
    const float d = pos - floorf(pos), e = 1.0f - d;
    dw[0] = -0.5f * e * e;
    dw[1] = -2.0f * d + 1.5f * d * d;
    dw[2] = 2.0f * e - 1.5f * e * e;
    dw[ORDER - 1] = 0.5f * d * d;
  }
}

template <int ORDER, bool REFCIC>
__global__ void __launch_bounds__(256) paint_grad_kernel(PaintParams p, const float* __restrict__ gmesh,
                                                         float* __restrict__ gx, float* __restrict__ gy,
                                                         float* __restrict__ gz, float* __restrict__ gw) {
  const int n = p.n;
  const size_t n2 = (size_t)n * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_part;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float wgt = p.w ? p.w[i] : 1.0f;
    const float px = grid_pos(p.x[i * p.stride], p.xmin, p.inv);
    const float py = grid_pos(p.y[i * p.stride], p.ymin, p.inv);
    const float pz = grid_pos(p.z[i * p.stride], p.zmin, p.inv);
    float ax = 0.0f, ay = 0.0f, az = 0.0f, aw = 0.0f;
    if (REFCIC) {
      int x0, x1, y0, y1, z0, z1;
      float mdx, ddx, mdy, ddy, mdz, ddz;
      cic_reference_axis(px, n, p.wrap, p.variant, x0, x1, mdx, ddx);
      cic_reference_axis(py, n, p.wrap, p.variant, y0, y1, mdy, ddy);
      cic_reference_axis(pz, n, p.wrap, p.variant, z0, z1, mdz, ddz);
      x0 = local_plane(x0, p.x0, p.nx, n); x1 = local_plane(x1, p.x0, p.nx, n);
      // corners of src/mas.py:142-151; factor derivative: d(md)/dpos = -1, d(dd)/dpos = +1
      // (the scan variant's zeroed dd at the box edge has derivative 0 there: ignored, measure-zero set)
#define JPS_GCORNER(ix, iy, iz, fx, fy, fz, sx, sy, sz)                                         \
  if (((ix) | (iy) | (iz)) >= 0) {                                                              \
    const float gc = gmesh[(size_t)(ix) * n2 + (size_t)(iy) * n + (iz)];                        \
    aw += (fx) * (fy) * (fz) * gc;                                                              \
    ax += (sx) * (fy) * (fz) * gc; ay += (fx) * (sy) * (fz) * gc; az += (fx) * (fy) * (sz) * gc; \
  }
      JPS_GCORNER(x0, y0, z0, mdx, mdy, mdz, -1.f, -1.f, -1.f)
      JPS_GCORNER(x1, y0, z0, ddx, mdy, mdz, 1.f, -1.f, -1.f)
      JPS_GCORNER(x0, y1, z0, mdx, ddy, mdz, -1.f, 1.f, -1.f)
      JPS_GCORNER(x0, y0, z1, mdx, mdy, ddz, -1.f, -1.f, 1.f)
      JPS_GCORNER(x1, y1, z0, ddx, ddy, mdz, 1.f, 1.f, -1.f)
      JPS_GCORNER(x1, y0, z1, ddx, mdy, ddz, 1.f, -1.f, 1.f)
      JPS_GCORNER(x0, y1, z1, mdx, mdy, ddz, -1.f, -1.f, 1.f)      // Q1: the weight really is mdx*mdy*ddz
      JPS_GCORNER(x1, y1, z1, ddx, ddy, ddz, 1.f, 1.f, 1.f)
#undef JPS_GCORNER
    } else {
      int ix[ORDER], iy[ORDER], iz[ORDER];
      float wx[ORDER], wy[ORDER], wz[ORDER], dx[ORDER], dy[ORDER], dz[ORDER];
      bspline_axis_grad<ORDER>(px, n, p.wrap, ix, wx, dx);
      bspline_axis_grad<ORDER>(py, n, p.wrap, iy, wy, dy);
      bspline_axis_grad<ORDER>(pz, n, p.wrap, iz, wz, dz);
#pragma unroll
      for (int a = 0; a < ORDER; ++a) {
        const int lx = local_plane(ix[a], p.x0, p.nx, n);
#pragma unroll
        for (int b = 0; b < ORDER; ++b) {
          if ((lx | iy[b]) < 0) continue;
          const float* row = gmesh + (size_t)lx * n2 + (size_t)iy[b] * n;
#pragma unroll
          for (int c = 0; c < ORDER; ++c) {
            if (iz[c] < 0) continue;
            const float gc = row[iz[c]];
            aw += wx[a] * wy[b] * wz[c] * gc;
            ax += dx[a] * wy[b] * wz[c] * gc;
            ay += wx[a] * dy[b] * wz[c] * gc;
            az += wx[a] * wy[b] * dz[c] * gc;
          }
        }
      }
    }
    const float s = wgt * p.inv;
    if (gx) gx[i] = ax * s;
    if (gy) gy[i] = ay * s;
    if (gz) gz[i] = az * s;
    if (gw) gw[i] = aw;
  }
}

// host launcher for other translation units (kernels cannot be launched across TUs without -rdc)
int launch_unpad_add(const float* field, int n, int rowpitch, const double* dcterm, float* out, cudaStream_t s) {
  ScopedLaunch L(K_MISC, s);
  const int blocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 16);
  unpad_add_kernel<<<blocks, 256, 0, s>>>(field, n, rowpitch, dcterm, out);
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

// powspec.cu
int npairs_for(int n);
int bin_from_dk_public(jps_plan* plan, const BinTable& T, int normalise, int mas_order, cudaStream_t s);

}  // namespace jps

using namespace jps;

extern "C" int jps_powspec_grad(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                                const float* k_edges, int nb, int mas_order, const float* grad_pk,
                                float* grad_mesh, void* stream) {
  JPS_REQUIRE(plan && mesh && k_edges && grad_pk && grad_mesh, "jps_powspec_grad: NULL argument");
  JPS_REQUIRE(nb >= 1 && nb <= kMaxUserBins && mas_order >= 2 && mas_order <= 4 && box_size > 0.0f,
              "jps_powspec_grad: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = plan->n;
  int rc = forward_fft(plan, mesh, s);
  if (rc) return rc;
  const float kF = ref_kF(box_size);
  std::vector<float> kg((size_t)nb + 1);
  for (int i = 0; i <= nb; ++i) kg[(size_t)i] = k_edges[i] / kF;
  BinTable* T = nullptr;
  rc = ensure_bin_table(plan, kg.data(), nb, TABLE_PK_EDGES, s, &T);
  if (rc) return rc;
  rc = bin_from_dk_public(plan, *T, normalise, mas_order, s);     // S_l[b] (needed by the DC term)
  if (rc) return rc;
  // scratch inside plan->scal: [0] dcterm (double), [8..) saved DC (float); gS lives after acc's used part
  double* dcterm = plan->scal;
  float* dc_saved = reinterpret_cast<float*>(plan->scal + 8);
  float* gS = reinterpret_cast<float*>(plan->acc + (size_t)std::max(T->nbc, 1) * 4);
  JPS_REQUIRE((size_t)T->nbc * 4 * 8 + (size_t)T->nbc * 3 * 4 <= (size_t)plan->acc_cap * 4 * 8,
              "jps_powspec_grad: too many bins for the plan scratch");
  JPS_CHECK_CUDA(cudaMemsetAsync(dcterm, 0, 8, s));
  JPS_CHECK_CUDA(cudaMemsetAsync(gS, 0, (size_t)std::max(T->nbc, 1) * 3 * 4, s));
  const float* wl = plan->wlut + (size_t)(mas_order - 2) * n;
  {
    ScopedLaunch L(K_MISC, s);
    save_dc_kernel<<<1, 1, 0, s>>>(plan->dk, dc_saved);
    pk_grad_coeff_kernel<<<(nb + 127) / 128, 128, 0, s>>>(nb, T->bin_to_compact, T->cnt, plan->acc, grad_pk,
                                                         ref_volume(box_size, n), normalise, plan->dk, gS, dcterm);
  }
  JPS_CHECK_LAUNCH();
  {
    ScopedLaunch L(K_MISC, s);
    const int blocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 16);
    // all CTAs read dk[0] (row 0) for the scale while other rows are overwritten: row 0 is written by
    // a separate launch afterwards, from the saved DC value
    pk_grad_modes_kernel<<<blocks, 256, 0, s>>>(plan->dk, n, plan->nz, plan->pitch, T->lut, wl, gS, normalise);
    pk_grad_row0_kernel<<<1, 256, 0, s>>>(plan->dk, n, plan->nz, T->lut, wl, gS, normalise, dc_saved);
  }
  JPS_CHECK_LAUNCH();
  JPS_CHECK_CUFFT(cufftSetStream(plan->c2r, s));
  {
    ScopedLaunch L(K_FFT_C2R, s);
    JPS_CHECK_CUFFT(cufftExecC2R(plan->c2r, (cufftComplex*)plan->dk, (cufftReal*)plan->dk));
  }
  {
    ScopedLaunch L(K_MISC, s);
    const int blocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 16);
    unpad_add_kernel<<<blocks, 256, 0, s>>>((const float*)plan->dk, n, 2 * plan->pitch, dcterm, grad_mesh);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

extern "C" int jps_paint_grad(int n_mesh, const float* x, const float* y, const float* z, const float* w,
                              int64_t stride, int64_t n_part, float xmin, float ymin, float zmin,
                              float box_size, int order, int wrap, int compat, int variant,
                              const float* grad_mesh, float* gx, float* gy, float* gz, float* gw,
                              void* stream) {
  JPS_REQUIRE(n_mesh >= 2 && n_mesh <= 4096 && n_part >= 0 && order >= 2 && order <= 4 && stride >= 1 &&
                  box_size > 0.0f && grad_mesh, "jps_paint_grad: bad arguments");
  JPS_REQUIRE(n_part == 0 || (x && y && z), "jps_paint_grad: x/y/z is NULL");
  if (n_part == 0) return JPS_OK;
  PaintParams p;
  p.n = n_mesh; p.x0 = 0; p.nx = n_mesh; p.wrap = wrap ? 1 : 0; p.variant = variant;
  p.xmin = xmin; p.ymin = ymin; p.zmin = zmin;
  const float bin_size = box_size / (float)n_mesh;
  p.inv = 1.0f / bin_size;
  p.stride = stride; p.n_part = n_part; p.x = x; p.y = y; p.z = z; p.w = w; p.mesh = nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>((n_part + 255) / 256, (int64_t)kNumSMs * 32);
  ScopedLaunch L(K_MISC, s);
  if (order == 2 && compat == JPS_COMPAT_REFERENCE) paint_grad_kernel<2, true><<<blocks, 256, 0, s>>>(p, grad_mesh, gx, gy, gz, gw);
  else if (order == 2) paint_grad_kernel<2, false><<<blocks, 256, 0, s>>>(p, grad_mesh, gx, gy, gz, gw);
  else if (order == 3) paint_grad_kernel<3, false><<<blocks, 256, 0, s>>>(p, grad_mesh, gx, gy, gz, gw);
  else paint_grad_kernel<4, false><<<blocks, 256, 0, s>>>(p, grad_mesh, gx, gy, gz, gw);
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}
