// Configuration-space multipoles xi_l(s): jps_xi / jps_xi_fundamental and the shared
// "from delta_k" stage used by the composite calls.
//
// Replaces /root/reference/src/correlations.py:120-187 (xi_vec), :191-261 (xi_vec_fundamental)
// and the xi blocks of the composites (:522-543, :689-710):
//   delta_k *= window ; delta2 = |delta_k|^2 ; delta_xi = irfftn(delta2) ; 3 weighted
//   histograms + 1 count histogram over |r| on the full N^3 grid ; normalise.
// Here: one elementwise kernel (|delta_k|^2, HBM stream), cuFFT C2R in place, one binning
// kernel that reads the real field once (4 B/cell) with the same integer-threshold bin lookup
// and warp-segmented reduction as the P(k) kernel.
#include "common.cuh"
#include "fold.cuh"

#include <algorithm>
#include <cmath>

namespace jps {

__global__ void __launch_bounds__(256) xi_power_kernel(const float2* __restrict__ dk,
                                                       float2* __restrict__ out, int n, int nz,
                                                       int pitch, const float* __restrict__ wl,
                                                       int normalise) {
  float scale2 = 1.0f;
  if (normalise) {
    const double dc = (double)dk[0].x;
    const double s = (double)n * (double)n * (double)n / dc;
    scale2 = (float)(s * s);
  }
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const float wxy = wl[ix] * wl[iy];
    const float2* src = dk + (size_t)row * pitch;
    float2* dst = out + (size_t)row * pitch;
    for (int kz = threadIdx.x; kz < pitch; kz += blockDim.x) {
      float2 o = make_float2(0.0f, 0.0f);
      if (kz < nz) {
        const float c = wxy * wl[kz];
        const float2 d = src[kz];
        const float re = d.x * c, im = d.y * c;
        o.x = (re * re + im * im) * scale2;
        if (normalise && row == 0 && kz == 0) o.x = 0.0f;     // delta_0 = 0 after rho/mean - 1
      }
      dst[kz] = o;
    }
  }
}

struct XiParams {
  const float* field;        // [n][n][rowpitch] real, = N^3 * irfftn(|delta_k|^2)
  int n, rowpitch;
  const int32_t* lut;
  int nbc;
  double* acc;               // [nbc][4]
  int guard_mu;              // 1: mu = 0 at r = 0 (composites); 0: NaN as xi_vec does (Q22)
};

// one warp per (ix,iy) row, lanes along z
template <bool SMEM>
__global__ void __launch_bounds__(256) xi_bin_kernel(XiParams P) {
  extern __shared__ float sacc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nacc = P.nbc * 3;
  if (SMEM) {
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) sacc[i] = 0.0f;
    __syncthreads();
  }
  const int n = P.n, mid = n / 2;
  const long long rows = (long long)n * n;
  for (long long row = (long long)blockIdx.x * nwarps + warp; row < rows; row += (long long)gridDim.x * nwarps) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const int rx = ix > mid ? ix - n : ix, ry = iy > mid ? iy - n : iy;
    const int r2xy = rx * rx + ry * ry;
    const float* src = P.field + (size_t)row * P.rowpitch;
    for (int z0 = 0; z0 < n; z0 += 32) {
      const int iz = z0 + lane;
      float v[3] = {0.0f, 0.0f, 0.0f};
      int cb = -2;
      if (iz < n) {
        const int rz = iz > mid ? iz - n : iz;
        const int r2 = r2xy + rz * rz;
        cb = __ldg(P.lut + r2);
        const float val = src[iz];
        float mu2;
        if (r2 > 0) mu2 = (float)(rz * rz) / (float)r2;
        else mu2 = P.guard_mu ? 0.0f : __int_as_float(0x7fc00000);   // 0/0 (Q22)
        v[0] = val;
        v[1] = val * (3.0f * mu2 - 1.0f) * 0.5f;
        v[2] = val * (35.0f * mu2 * mu2 - 30.0f * mu2 + 3.0f) * 0.125f;
      }
      const int prev = __shfl_up_sync(0xffffffffu, cb, 1);
      const bool head = (lane == 0) || (cb != prev);
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      segmented_reduce<3>(v, heads, lane);
      if (head && cb >= 0) {
        if (SMEM) {
          // |r| is not monotone along z (the index wraps to negative lags): the same bin can
          // head two segments of one warp, so these are real atomics (few per warp).
          atomicAdd(sacc + cb * 3 + 0, v[0]);
          atomicAdd(sacc + cb * 3 + 1, v[1]);
          atomicAdd(sacc + cb * 3 + 2, v[2]);
        } else {
          atomicAdd(P.acc + (size_t)cb * 4 + 0, (double)v[0]);
          atomicAdd(P.acc + (size_t)cb * 4 + 1, (double)v[1]);
          atomicAdd(P.acc + (size_t)cb * 4 + 2, (double)v[2]);
        }
      }
    }
  }
  if (SMEM) {
    __syncthreads();
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
      const float s = sacc[i];
      if (s != 0.0f || s != s) atomicAdd(P.acc + (size_t)(i / 3) * 4 + (i % 3), (double)s);
    }
  }
}

struct XiFinalizeParams {
  int nb, first_bin;
  const int32_t* bin_to_compact;
  const float* edges;
  const double* acc;
  const unsigned long long* cnt;
  const double* ksum;
  double inv_n6;              // 1/N^3 (irfftn) * 1/N^3 (the reference's / dims**3)
  float cell;                 // box_size * 1.0 / dims
  int fundamental;
  float* r3d; float* xi3d; float* nmodes; double* sums; int64_t* counts;
};

__global__ void xi_finalize_kernel(XiFinalizeParams F) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= F.nb) return;
  const int bin = j + F.first_bin;
  const int c = F.bin_to_compact[bin];
  double s0 = 0, s2 = 0, s4 = 0, ks = 0;
  unsigned long long cnt = 0;
  if (c >= 0) {
    s0 = F.acc[(size_t)c * 4 + 0]; s2 = F.acc[(size_t)c * 4 + 1]; s4 = F.acc[(size_t)c * 4 + 2];
    cnt = F.cnt[c]; ks = F.ksum[c];
  }
  double nm = (double)(float)cnt;
  float nm_out = (float)cnt;
  if (!F.fundamental && cnt == 0) {             // Nmodes3D = where(Nmodes3D == 0, inf, Nmodes3D) (Q11)
    nm = INFINITY;
    nm_out = INFINITY;
  }
  F.xi3d[j * 3 + 0] = (float)(s0 / nm * F.inv_n6);
  F.xi3d[j * 3 + 1] = (float)(s2 / nm * 5.0 * F.inv_n6);
  F.xi3d[j * 3 + 2] = (float)(s4 / nm * 9.0 * F.inv_n6);
  F.nmodes[j] = nm_out;
  if (F.fundamental) F.r3d[j] = (float)(ks / (double)cnt) * F.cell;      // r3D.at[].add(r) / Nmodes * cell
  else F.r3d[j] = (0.5f * (F.edges[bin + 1] + F.edges[bin])) * F.cell;
  if (F.sums) { F.sums[j * 3 + 0] = s0; F.sums[j * 3 + 1] = s2; F.sums[j * 3 + 2] = s4; }
  if (F.counts) F.counts[j] = (int64_t)cnt;
}

// xi stage given plan->dk (forward FFT already done).  Uses shell field 0 as scratch.
int xi_from_dk(jps_plan* plan, const BinTable& T, int normalise, int mas_order, int guard_mu,
               cudaStream_t s) {
  if (plan->n_shell_fields < 1) {
    set_error("xi needs a plan created with n_shell_fields >= 1 (scratch for the xi(r) field)");
    return JPS_ERR_WORKSPACE;
  }
  float2* buf = (float2*)plan->shell;
  const int n = plan->n;
  {
    ScopedLaunch L(K_SHELL_FILTER, s);
    const int blocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 16);
    xi_power_kernel<<<blocks, 256, 0, s>>>(plan->dk, buf, n, plan->nz, plan->pitch,
                                           plan->wlut + (size_t)(mas_order - 2) * n, normalise);
  }
  JPS_CHECK_LAUNCH();
  JPS_CHECK_CUFFT(cufftSetStream(plan->c2r, s));
  {
    ScopedLaunch L(K_FFT_C2R, s);
    JPS_CHECK_CUFFT(cufftExecC2R(plan->c2r, (cufftComplex*)buf, (cufftReal*)buf));
  }
  {
    ScopedLaunch L(K_MEMSET, s);
    JPS_CHECK_CUDA(cudaMemsetAsync(plan->acc, 0, (size_t)std::max(T.nbc, 1) * 4 * 8, s));
  }
  if (T.nbc == 0) return JPS_OK;
  XiParams P;
  P.field = (const float*)buf; P.n = n; P.rowpitch = 2 * plan->pitch; P.lut = T.lut; P.nbc = T.nbc;
  P.acc = plan->acc; P.guard_mu = guard_mu;
  const size_t smem = (size_t)T.nbc * 3 * sizeof(float);
  const long long want = ((long long)n * n + 7) / 8;
  if (smem <= 160 * 1024) {
    static PerDeviceFlag attr_set;
    if (!attr_set.get()) {
      JPS_CHECK_CUDA(cudaFuncSetAttribute(xi_bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_set.set();
    }
    int per_sm = 1;
    JPS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, xi_bin_kernel<true>, 256, smem));
    const int blocks = (int)std::min<long long>(want, (long long)kNumSMs * std::max(per_sm, 1));
    ScopedLaunch L(K_XI_BIN, s);
    xi_bin_kernel<true><<<blocks, 256, smem, s>>>(P);
  } else {
    const int blocks = (int)std::min<long long>(want, (long long)kNumSMs * 8);
    ScopedLaunch L(K_XI_BIN, s);
    xi_bin_kernel<false><<<blocks, 256, 0, s>>>(P);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

int xi_finalize(jps_plan* plan, const BinTable& T, float box_size, int nb, int first_bin,
                int fundamental, float* r3d, float* xi3d, float* nmodes, double* sums,
                int64_t* counts, cudaStream_t s) {
  if (nb < 1) return JPS_OK;
  XiFinalizeParams F;
  F.nb = nb; F.first_bin = first_bin; F.bin_to_compact = T.bin_to_compact; F.edges = T.edges;
  F.acc = plan->acc; F.cnt = T.cnt; F.ksum = T.ksum;
  const double n3 = (double)plan->n * plan->n * plan->n;
  F.inv_n6 = 1.0 / n3 / n3;
  F.cell = box_size * 1.0f / (float)plan->n;
  F.fundamental = fundamental;
  F.r3d = r3d; F.xi3d = xi3d; F.nmodes = nmodes; F.sums = sums; F.counts = counts;
  {
    ScopedLaunch L(K_PK_FINALIZE, s);
    xi_finalize_kernel<<<(nb + 127) / 128, 128, 0, s>>>(F);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

// kedges for xi: k_edges = kF*s_edges*dims/box_size ; kedges = k_edges/kF (float32), :126,170
void xi_grid_edges(const float* s_edges, int nb, float box_size, int n, std::vector<float>& out) {
  const float kF = ref_kF(box_size);
  out.resize((size_t)nb + 1);
  for (int i = 0; i <= nb; ++i) {
    const float ke = ((kF * s_edges[i]) * (float)n) / box_size;
    out[(size_t)i] = ke / kF;
  }
}

}  // namespace jps

using namespace jps;

extern "C" int jps_xi(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                      const float* s_edges, int nb, int mas_order, int guard_mu, float* r3d,
                      float* xi3d, float* nmodes, double* sums, int64_t* counts, void* stream) {
  JPS_REQUIRE(plan && mesh && s_edges && r3d && xi3d && nmodes, "jps_xi: NULL argument");
  JPS_REQUIRE(nb >= 1 && nb <= kMaxUserBins, "jps_xi: nb=%d out of range", nb);
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4, "jps_xi: mas_order must be 2, 3 or 4");
  JPS_REQUIRE(box_size > 0.0f, "jps_xi: box_size must be > 0");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<float> kg;
  xi_grid_edges(s_edges, nb, box_size, plan->n, kg);
  BinTable* T = nullptr;
  int rc = ensure_bin_table(plan, kg.data(), nb, TABLE_XI_EDGES, s, &T);
  if (rc) return rc;
  rc = forward_fft(plan, mesh, s);
  if (rc) return rc;
  rc = xi_from_dk(plan, *T, normalise, mas_order, guard_mu, s);
  if (rc) return rc;
  return xi_finalize(plan, *T, box_size, nb, 0, 0, r3d, xi3d, nmodes, sums, counts, s);
}

extern "C" int jps_xi_fundamental(jps_plan_t* plan, const float* mesh, int normalise,
                                  float box_size, int mas_order, float* r3d, float* xi3d,
                                  float* nmodes, double* sums, int64_t* counts, void* stream) {
  JPS_REQUIRE(plan && mesh && r3d && xi3d && nmodes, "jps_xi_fundamental: NULL argument");
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4, "jps_xi_fundamental: mas_order must be 2, 3 or 4");
  JPS_REQUIRE(box_size > 0.0f, "jps_xi_fundamental: box_size must be > 0");
  cudaStream_t s = (cudaStream_t)stream;
  BinTable* T = nullptr;
  int rc = ensure_bin_table(plan, nullptr, 0, TABLE_XI_FUNDAMENTAL, s, &T);
  if (rc) return rc;
  rc = forward_fft(plan, mesh, s);
  if (rc) return rc;
  rc = xi_from_dk(plan, *T, normalise, mas_order, 0, s);
  if (rc) return rc;
  return xi_finalize(plan, *T, box_size, jps_fundamental_nbins(plan->n), 1, 1, r3d, xi3d, nmodes,
                     sums, counts, s);
}
