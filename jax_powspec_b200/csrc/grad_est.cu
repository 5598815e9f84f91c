// Gradients of the configuration-space multipoles and of the FFT bispectrum with respect to the
// input mesh: the reference's optimisation scripts differentiate a loss on
// P(k), xi(s) AND B(theta) through compute_all_correlations
// (/root/reference/tests/lognormal_bispec.py:71-106, tests/lognormal_xi.py:74-108).
//
//   jps_xi_grad     : cotangent of xi3D [nb][3]              -> cotangent of the mesh
//   jps_bispec_grad : cotangents of Pk [bins+2] and B [bins] -> cotangent of the mesh
//                     (Q = B / (P0 P1 + P0 P3 + P1 P3) is folded into these two by the caller)
//
// Conventions: rho_k = R2C(mesh) (unnormalised), D_k = c_k rho_k (c = window correction), C2R is
// cuFFT's unnormalised inverse: C2R(Y)(r) = sum over STORED modes of fac_k Re(Y_k e^{ikr}), fac = 1
// on the kz = 0 and kz = N/2 planes and 2 elsewhere.  Two adjoint identities carry everything:
//   y = C2R(Y)     ->  Ybar_k = fac_k R2C(ybar)_k
//   Y = R2C(mesh)  ->  meshbar = C2R(Ybar / fac)
// so the fac factors cancel and every backward pass ends in ONE C2R of a half-spectrum Z.
//
// xi (src/correlations.py:120-187):  val = C2R(|D|^2),  xi_l[b] = m_l/(cnt_b N^6) sum_{r in b} L_l(mu_r) val(r)
//   G(r) = sum_l gxi[b(r)][l] m_l L_l(mu_r) / (cnt N^6);   Z_k = 2 c_k^2 Re(R2C(G)_k) rho_k
//
// bispectrum (:334-462), in the reference's normalised fields d_j = irfftn(m_j D) = C2R(m_j D) / N^3:
//   P_j = vol_p sum d_j^2 / sum i_j^2,   B_b = vol_b sum d_0 d_1 d_{b+2} / sum i_0 i_1 i_{b+2}
//   alpha_j = gP_j vol_p / sum i_j^2,    tau_b = gB_b vol_b / sum i_0 i_1 i_{b+2}
//   E = sum_b tau_b d_{b+2} = irfftn(M_E D),  M_E(k) = sum_b tau_b m_{b+2}(k)     (ONE inverse FFT for all bins)
//   Z_k = c_k / N^3 [ 2 M_A(k) D_k + m_0 R2C(d_1 E) + m_1 R2C(d_0 E) + M_E R2C(d_0 d_1) ],  M_A = sum_j alpha_j m_j
// 3 inverse + 3 forward FFTs + the final inverse, whatever the number of theta bins (the forward
// pass needs bins + 2).
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace jps {

// grad.cu
int launch_unpad_add(const float* field, int n, int rowpitch, const double* dcterm, float* out, cudaStream_t s);
// bispec.cu
int launch_shell_filter(jps_plan* plan, int mas_order, int tlo, int thi, float* out, cudaStream_t s);
int bispec_from_dk(jps_plan* plan, int normalise, float box_size, float k1, float k2, const float* theta,
                   int nbins, int mas_order, float* k_all_out, float* pk_out, float* B_out, float* Q_out,
                   cudaStream_t s);
// xi.cu
void xi_grid_edges(const float* s_edges, int nb, float box_size, int n, std::vector<float>& out);

// ---------------------------------------------------------------- xi
__global__ void xi_grad_coeff_kernel(int nb, int first_bin, const int32_t* __restrict__ bin_to_compact,
                                     const unsigned long long* __restrict__ cnt,
                                     const float* __restrict__ grad_xi, double inv_n6, float* __restrict__ gC) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb) return;
  const int c = bin_to_compact[j + first_bin];
  if (c < 0) return;
  const double nm = (double)(float)cnt[c];
  const double mult[3] = {1.0, 5.0, 9.0};
  for (int l = 0; l < 3; ++l) {
    const float g = grad_xi[j * 3 + l];
    gC[c * 3 + l] = (nm > 0.0 && g == g) ? (float)((double)g * mult[l] * inv_n6 / nm) : 0.0f;   // NaN cotangent -> 0
  }
}

// G(r) on the padded real grid [n][n][rowpitch]
__global__ void __launch_bounds__(256) xi_grad_field_kernel(float* __restrict__ field, int n, int rowpitch,
                                                            const int32_t* __restrict__ lut,
                                                            const float* __restrict__ gC) {
  const int mid = n / 2;
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const int rx = ix > mid ? ix - n : ix, ry = iy > mid ? iy - n : iy;
    const int r2xy = rx * rx + ry * ry;
    float* dst = field + (size_t)row * rowpitch;
    for (int iz = threadIdx.x; iz < rowpitch; iz += blockDim.x) {
      float g = 0.0f;
      if (iz < n) {
        const int rz = iz > mid ? iz - n : iz;
        const int r2 = r2xy + rz * rz;
        const int cb = lut[r2];
        if (cb >= 0) {
          g = gC[cb * 3];
          if (r2 > 0) {
            const float mu2 = (float)(rz * rz) / (float)r2;
            g += gC[cb * 3 + 1] * (3.0f * mu2 - 1.0f) * 0.5f +
                 gC[cb * 3 + 2] * (35.0f * mu2 * mu2 - 30.0f * mu2 + 3.0f) * 0.125f;
          } else {
            // r = 0 has no direction: mu = 0 as the composites define it (:527); xi_vec itself yields
            // NaN there (Q22) and its NaN cotangents were zeroed above
            g += gC[cb * 3 + 1] * -0.5f + gC[cb * 3 + 2] * 0.375f;
          }
        }
      }
      dst[iz] = g;
    }
  }
}

// Z_k = 2 c_k^2 Re(Ghat_k) rho_k, in place over delta_k
__global__ void __launch_bounds__(256) xi_grad_modes_kernel(float2* __restrict__ dk, const float2* __restrict__ ghat,
                                                            int n, int nz, int pitch,
                                                            const float* __restrict__ wl) {
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const float wxy = wl[ix] * wl[iy];
    float2* r = dk + (size_t)row * pitch;
    const float2* gh = ghat + (size_t)row * pitch;
    for (int kz = threadIdx.x; kz < nz; kz += blockDim.x) {
      const float c = wxy * wl[kz];
      const float f = 2.0f * (c * c) * gh[kz].x;
      const float2 d = r[kz];
      r[kz] = make_float2(f * d.x, f * d.y);
    }
  }
}

// ---------------------------------------------------------------- bispectrum
struct BispecGradScratch {        // carved out of the 6th shell field
  double* alpha;                  // [nshell]
  double* tau;                    // [nbins]
  int* tlo;                       // [nshell]
  int* thi;                       // [nshell]
  int* slots;                     // [2 nshell]
  float* MA;                      // [k2max + 1]
  float* ME;                      // [k2max + 1]
  float* fwd_out;                 // [4 nshell] dummy outputs of a cache-filling forward pass
};

__global__ void bispec_grad_coeff_kernel(int nshell, const float* __restrict__ grad_pk,
                                         const float* __restrict__ grad_B, const double* __restrict__ isum,
                                         const int* __restrict__ slots, double vol_p, double vol_b, double n3,
                                         double* __restrict__ alpha, double* __restrict__ tau) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nshell) return;
  // isum holds the sums over the UNNORMALISED indicator fields: sum i^2 = isum2 / N^6, sum iii = isum3 / N^9
  const float gp = grad_pk[j];
  const double s2 = isum[slots[j]];
  alpha[j] = (gp == gp && s2 != 0.0) ? (double)gp * vol_p * (n3 * n3) / s2 : 0.0;
  if (j >= 2) {
    const float gb = grad_B[j - 2];
    const double s3 = isum[slots[nshell + j]];
    tau[j - 2] = (gb == gb && s3 != 0.0) ? (double)gb * vol_b * (n3 * n3 * n3) / s3 : 0.0;
  }
}

__global__ void bispec_grad_lut_kernel(int k2max, int nshell, const int* __restrict__ tlo,
                                       const int* __restrict__ thi, const double* __restrict__ alpha,
                                       const double* __restrict__ tau, float* __restrict__ MA,
                                       float* __restrict__ ME) {
  const int k2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (k2 > k2max) return;
  double a = 0.0, e = 0.0;
  for (int j = 0; j < nshell; ++j) {
    if (k2 >= tlo[j] && k2 < thi[j]) {
      a += alpha[j];
      if (j >= 2) e += tau[j - 2];
    }
  }
  MA[k2] = (float)a;
  ME[k2] = (float)e;
}

// out_k = lut[k^2] c_k rho_k  (the E field before its inverse FFT)
__global__ void __launch_bounds__(256) shell_filter_lut_kernel(const float2* __restrict__ dk, int n, int nz, int pitch,
                                                               const float* __restrict__ wl,
                                                               const float* __restrict__ lut,
                                                               float2* __restrict__ out) {
  const int mid = n / 2;
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const int kx = ix > mid ? ix - n : ix, ky = iy > mid ? iy - n : iy;
    const int k2xy = kx * kx + ky * ky;
    const float wxy = wl[ix] * wl[iy];
    const float2* src = dk + (size_t)row * pitch;
    float2* dst = out + (size_t)row * pitch;
    for (int kz = threadIdx.x; kz < pitch; kz += blockDim.x) {
      float2 o = make_float2(0.0f, 0.0f);
      if (kz < nz) {
        const float m = lut[k2xy + kz * kz];
        if (m != 0.0f) {
          const float c = wxy * wl[kz] * m;
          const float2 d = src[kz];
          o = make_float2(d.x * c, d.y * c);
        }
      }
      dst[kz] = o;
    }
  }
}

// p3 = s^2 d0 d1, p1 = s^2 d1 e, e <- s^2 d0 e   (s = 1/N^3: the fields are unnormalised C2R outputs)
__global__ void __launch_bounds__(256) bispec_grad_products_kernel(const float* __restrict__ d0,
                                                                   const float* __restrict__ d1,
                                                                   float* __restrict__ e, float* __restrict__ p3,
                                                                   float* __restrict__ p1, int n, int rowpitch,
                                                                   float s) {
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const size_t base = (size_t)row * rowpitch;
    for (int j = threadIdx.x; j < rowpitch; j += blockDim.x) {
      float a = 0.0f, b = 0.0f, c = 0.0f;
      if (j < n) {
        const float x0 = d0[base + j] * s, x1 = d1[base + j] * s, xe = e[base + j] * s;
        a = x0 * x1;
        b = x1 * xe;
        c = x0 * xe;
      }
      p3[base + j] = a;
      p1[base + j] = b;
      e[base + j] = c;
    }
  }
}

// Z_k = c_k / N^3 [ 2 M_A c_k rho_k + m_0 F1 + m_1 F2 + M_E F3 ], in place over delta_k
__global__ void __launch_bounds__(256) bispec_grad_combine_kernel(float2* __restrict__ dk, const float2* __restrict__ f1,
                                                                  const float2* __restrict__ f2,
                                                                  const float2* __restrict__ f3, int n, int nz,
                                                                  int pitch, const float* __restrict__ wl,
                                                                  const float* __restrict__ MA,
                                                                  const float* __restrict__ ME, int tlo0, int thi0,
                                                                  int tlo1, int thi1, float inv_n3) {
  const int mid = n / 2;
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const int kx = ix > mid ? ix - n : ix, ky = iy > mid ? iy - n : iy;
    const int k2xy = kx * kx + ky * ky;
    const float wxy = wl[ix] * wl[iy];
    const size_t base = (size_t)row * pitch;
    for (int kz = threadIdx.x; kz < nz; kz += blockDim.x) {
      const int k2 = k2xy + kz * kz;
      const float c = wxy * wl[kz];
      const float ma = MA[k2], me = ME[k2];
      const float2 d = dk[base + kz];
      float zx = 2.0f * ma * c * d.x, zy = 2.0f * ma * c * d.y;
      if (k2 >= tlo0 && k2 < thi0) { const float2 v = f1[base + kz]; zx += v.x; zy += v.y; }
      if (k2 >= tlo1 && k2 < thi1) { const float2 v = f2[base + kz]; zx += v.x; zy += v.y; }
      if (me != 0.0f) { const float2 v = f3[base + kz]; zx += me * v.x; zy += me * v.y; }
      const float f = c * inv_n3;
      dk[base + kz] = make_float2(f * zx, f * zy);
    }
  }
}

static int c2r_inplace(jps_plan* plan, float* field, cudaStream_t s) {
  ScopedLaunch L(K_FFT_C2R, s);
  JPS_CHECK_CUFFT(cufftExecC2R(plan->c2r, (cufftComplex*)field, (cufftReal*)field));
  return JPS_OK;
}

static int r2c_inplace(jps_plan* plan, float* field, cudaStream_t s) {
  ScopedLaunch L(K_FFT_R2C, s);
  JPS_CHECK_CUFFT(cufftExecR2C(plan->r2c_ip, (cufftReal*)field, (cufftComplex*)field));
  return JPS_OK;
}

static int finish_with_c2r(jps_plan* plan, float* grad_mesh, cudaStream_t s) {
  const int n = plan->n;
  JPS_CHECK_CUDA(cudaMemsetAsync(plan->scal, 0, 8, s));
  int rc = c2r_inplace(plan, (float*)plan->dk, s);
  if (rc) return rc;
  return launch_unpad_add((const float*)plan->dk, n, 2 * plan->pitch, plan->scal, grad_mesh, s);
}

}  // namespace jps

using namespace jps;

extern "C" int jps_xi_grad(jps_plan_t* plan, const float* mesh, float box_size, const float* s_edges, int nb,
                           int mas_order, const float* grad_xi, float* grad_mesh, void* stream) {
  JPS_REQUIRE(plan && mesh && s_edges && grad_xi && grad_mesh, "jps_xi_grad: NULL argument");
  JPS_REQUIRE(nb >= 1 && nb <= kMaxUserBins && mas_order >= 2 && mas_order <= 4 && box_size > 0.0f,
              "jps_xi_grad: bad arguments");
  JPS_REQUIRE(plan->n_shell_fields >= 1 && plan->r2c_ip_ok, "jps_xi_grad: the plan needs n_shell_fields >= 1");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = plan->n;
  std::vector<float> kg;
  xi_grid_edges(s_edges, nb, box_size, n, kg);
  BinTable* T = nullptr;
  int rc = ensure_bin_table(plan, kg.data(), nb, TABLE_XI_EDGES, s, &T);
  if (rc) return rc;
  rc = forward_fft(plan, mesh, s);
  if (rc) return rc;
  float* gC = reinterpret_cast<float*>(plan->acc);                // [nbc][3] floats inside the accumulator scratch
  JPS_REQUIRE((size_t)std::max(T->nbc, 1) * 3 * 4 <= (size_t)plan->acc_cap * 4 * 8, "jps_xi_grad: too many bins");
  JPS_CHECK_CUDA(cudaMemsetAsync(gC, 0, (size_t)std::max(T->nbc, 1) * 3 * 4, s));
  const double n3 = (double)n * n * n;
  float* G = plan->shell;
  JPS_CHECK_CUFFT(cufftSetStream(plan->c2r, s));
  JPS_CHECK_CUFFT(cufftSetStream(plan->r2c_ip, s));
  const int blocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 16);
  {
    ScopedLaunch L(K_MISC, s);
    xi_grad_coeff_kernel<<<(nb + 127) / 128, 128, 0, s>>>(nb, 0, T->bin_to_compact, T->cnt, grad_xi, 1.0 / n3 / n3, gC);
    xi_grad_field_kernel<<<blocks, 256, 0, s>>>(G, n, 2 * plan->pitch, T->lut, gC);
  }
  JPS_CHECK_LAUNCH();
  rc = r2c_inplace(plan, G, s);
  if (rc) return rc;
  {
    ScopedLaunch L(K_MISC, s);
    xi_grad_modes_kernel<<<blocks, 256, 0, s>>>(plan->dk, (const float2*)G, n, plan->nz, plan->pitch,
                                                plan->wlut + (size_t)(mas_order - 2) * n);
  }
  JPS_CHECK_LAUNCH();
  return finish_with_c2r(plan, grad_mesh, s);
}

extern "C" int jps_bispec_grad(jps_plan_t* plan, const float* mesh, float box_size, float k1, float k2,
                               const float* theta, int nbins, int mas_order, const float* grad_pk,
                               const float* grad_B, float* grad_mesh, void* stream) {
  JPS_REQUIRE(plan && mesh && theta && grad_pk && grad_B && grad_mesh, "jps_bispec_grad: NULL argument");
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4 && box_size > 0.0f && nbins >= 1 && nbins + 2 <= 250,
              "jps_bispec_grad: bad arguments");
  JPS_REQUIRE(plan->n_shell_fields >= 6 && plan->r2c_ip_ok, "jps_bispec_grad: the plan needs n_shell_fields >= 6");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = plan->n, nshell = nbins + 2;
  const size_t field_floats = (size_t)n * n * 2 * plan->pitch;
  float* F[6];
  for (int i = 0; i < 6; ++i) F[i] = plan->shell + (size_t)i * field_floats;
  const int k2max = (int)plan->k2max;
  // scratch inside the 6th shell field
  BispecGradScratch S;
  {
    char* p = reinterpret_cast<char*>(F[5]);
    size_t off = 0;
    auto take = [&](size_t b) { char* q = p + off; off = align_up(off + b, 256); return q; };
    S.alpha = (double*)take((size_t)nshell * 8);
    S.tau = (double*)take((size_t)nshell * 8);
    S.tlo = (int*)take((size_t)nshell * 4);
    S.thi = (int*)take((size_t)nshell * 4);
    S.slots = (int*)take((size_t)2 * nshell * 4);
    S.MA = (float*)take((size_t)(k2max + 1) * 4);
    S.ME = (float*)take((size_t)(k2max + 1) * 4);
    S.fwd_out = (float*)take((size_t)4 * nshell * 4);
    JPS_REQUIRE(off <= field_floats * 4, "jps_bispec_grad: mesh too small for the gradient scratch");
  }
  int rc = forward_fft(plan, mesh, s);
  if (rc) return rc;
  // shells exactly as the forward pass (bispec.cu): k_all in float32, integer thresholds on k^2
  const float kF = ref_kF(box_size);
  std::vector<float> k_all((size_t)nshell);
  k_all[0] = k1; k_all[1] = k2;
  for (int b = 0; b < nbins; ++b) {
    const float sn = k2 * sinf(theta[b]);
    const float cs = k2 * cosf(theta[b]) + k1;
    k_all[(size_t)b + 2] = sqrtf(sn * sn + cs * cs);
  }
  std::vector<int> tlo((size_t)nshell), thi((size_t)nshell);
  for (int j = 0; j < nshell; ++j) {
    const float lo = (k_all[(size_t)j] - kF) / kF, hi = (k_all[(size_t)j] + kF) / kF;
    tlo[(size_t)j] = (int)edge_threshold(lo, false, plan->k2max);
    thi[(size_t)j] = (int)edge_threshold(hi, false, plan->k2max);
  }
  // indicator sums are cached by the forward pass; run it once if this geometry has not been seen
  std::vector<int> slots((size_t)2 * nshell, 0);
  auto lookup = [&]() {
    for (int j = 0; j < nshell; ++j) {
      auto it = plan->isum_slot.find({tlo[(size_t)j], thi[(size_t)j]});
      if (it == plan->isum_slot.end()) return false;
      slots[(size_t)j] = it->second;
      if (j >= 2) {
        auto it3 = plan->isum_slot.find({tlo[0], thi[0], tlo[1], thi[1], tlo[(size_t)j], thi[(size_t)j]});
        if (it3 == plan->isum_slot.end()) return false;
        slots[(size_t)nshell + j] = it3->second;
      }
    }
    return true;
  };
  if (!lookup()) {
    rc = bispec_from_dk(plan, 0, box_size, k1, k2, theta, nbins, mas_order, S.fwd_out, S.fwd_out + nshell,
                        S.fwd_out + 2 * nshell, S.fwd_out + 3 * nshell, s);
    if (rc) return rc;
    JPS_REQUIRE(lookup(), "jps_bispec_grad: indicator sums missing after the forward pass");
  }
  JPS_CHECK_CUDA(cudaMemcpyAsync(S.tlo, tlo.data(), (size_t)nshell * 4, cudaMemcpyHostToDevice, s));
  JPS_CHECK_CUDA(cudaMemcpyAsync(S.thi, thi.data(), (size_t)nshell * 4, cudaMemcpyHostToDevice, s));
  JPS_CHECK_CUDA(cudaMemcpyAsync(S.slots, slots.data(), slots.size() * 4, cudaMemcpyHostToDevice, s));
  const double n3 = (double)n * n * n;
  const float tp = box_size / (float)(n * n);
  const float tb = (box_size * box_size) / (float)((long long)n * n * n);
  const float* wl = plan->wlut + (size_t)(mas_order - 2) * n;
  const int blocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 16);
  JPS_CHECK_CUFFT(cufftSetStream(plan->c2r, s));
  JPS_CHECK_CUFFT(cufftSetStream(plan->r2c_ip, s));
  {
    ScopedLaunch L(K_MISC, s);
    bispec_grad_coeff_kernel<<<(nshell + 127) / 128, 128, 0, s>>>(nshell, grad_pk, grad_B, plan->isum, S.slots,
                                                                 (double)(tp * tp * tp), (double)(tb * tb * tb), n3,
                                                                 S.alpha, S.tau);
    bispec_grad_lut_kernel<<<(k2max + 256) / 256, 256, 0, s>>>(k2max, nshell, S.tlo, S.thi, S.alpha, S.tau, S.MA, S.ME);
  }
  JPS_CHECK_LAUNCH();
  // d0 -> F[0], d1 -> F[1], E -> F[2]  (unnormalised inverse FFTs)
  rc = launch_shell_filter(plan, mas_order, tlo[0], thi[0], F[0], s);
  if (rc) return rc;
  rc = launch_shell_filter(plan, mas_order, tlo[1], thi[1], F[1], s);
  if (rc) return rc;
  {
    ScopedLaunch L(K_SHELL_FILTER, s);
    shell_filter_lut_kernel<<<blocks, 256, 0, s>>>(plan->dk, n, plan->nz, plan->pitch, wl, S.ME, (float2*)F[2]);
  }
  JPS_CHECK_LAUNCH();
  for (int i = 0; i < 3; ++i) {
    rc = c2r_inplace(plan, F[i], s);
    if (rc) return rc;
  }
  {
    ScopedLaunch L(K_TRIPLE_REDUCE, s);
    bispec_grad_products_kernel<<<blocks, 256, 0, s>>>(F[0], F[1], F[2], F[3], F[4], n, 2 * plan->pitch,
                                                       (float)(1.0 / n3));
  }
  JPS_CHECK_LAUNCH();
  // F[4] = d1 E -> F1, F[2] = d0 E -> F2, F[3] = d0 d1 -> F3
  for (int i = 2; i <= 4; ++i) {
    rc = r2c_inplace(plan, F[i], s);
    if (rc) return rc;
  }
  {
    ScopedLaunch L(K_MISC, s);
    bispec_grad_combine_kernel<<<blocks, 256, 0, s>>>(plan->dk, (const float2*)F[4], (const float2*)F[2],
                                                      (const float2*)F[3], n, plan->nz, plan->pitch, wl, S.MA, S.ME,
                                                      tlo[0], thi[0], tlo[1], thi[1], (float)(1.0 / n3));
  }
  JPS_CHECK_LAUNCH();
  return finish_with_c2r(plan, grad_mesh, s);
}
