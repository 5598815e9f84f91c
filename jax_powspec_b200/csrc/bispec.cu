// FFT ("Scoccimarro") bispectrum: jps_bispec and the shared "from delta_k" stage.
//
// Replaces /root/reference/src/correlations.py:334-462 (bispec) and the bispectrum block of
// compute_all_correlations (:557-632):
//   shells j = (k1, k2, k3(theta_b)...), half-width kF, mask lo_j <= |k| < hi_j in grid units;
//   d_j = irfftn(mask_j * delta_k), I_j = irfftn(mask_j);
//   P_j = sum d_j^2 / sum I_j^2 * (box/N^2)^3
//   B_b = sum d_0 d_1 d_{b+2} / sum I_0 I_1 I_{b+2} * (box^2/N^3)^3 ; Q_b = B_b/(P0P1+P0P3+P1P3)
//
// Here the [N,N,N/2+1,bins+2] boolean mask of the reference is never materialised: shell
// membership is an integer test on k^2 against thresholds derived from the float32 shell bounds
// (same construction as the P(k) bin edges), one kernel writes the masked delta_k AND the
// indicator in a single read of delta_k, cuFFT C2R runs in place, and one fused kernel reads
// the six real fields once to produce the four sums a triangle bin needs.  cuFFT's C2R is
// unnormalised; the N^3 factors cancel in every ratio, exactly as the 1/N^3 of irfftn does in
// the reference (Q20).
#include "common.cuh"
#include "fold.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>

namespace jps {

constexpr int kMaxShells = 250;      // 2 doubles per shell in the 1024-double scratch
constexpr int kMaxShellsDev = 256;

__global__ void __launch_bounds__(256) shell_filter_kernel(const float2* __restrict__ dk, int n,
                                                           int nz, int pitch,
                                                           const float* __restrict__ wl,
                                                           int normalise, int tlo, int thi,
                                                           float2* __restrict__ out_delta,
                                                           float2* __restrict__ out_ind) {
  float scale = 1.0f;
  if (normalise) scale = (float)((double)n * (double)n * (double)n / (double)dk[0].x);
  const int mid = n / 2;
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const int kx = ix > mid ? ix - n : ix, ky = iy > mid ? iy - n : iy;
    const int k2xy = kx * kx + ky * ky;
    const float wxy = wl[ix] * wl[iy];
    const float2* src = dk + (size_t)row * pitch;
    float2* dd = out_delta + (size_t)row * pitch;
    float2* di = out_ind ? out_ind + (size_t)row * pitch : nullptr;   // indicator only when its sums are not cached
    for (int kz = threadIdx.x; kz < pitch; kz += blockDim.x) {
      float2 od = make_float2(0.0f, 0.0f), oi = make_float2(0.0f, 0.0f);
      if (kz < nz) {
        const int k2 = k2xy + kz * kz;
        if (k2 >= tlo && k2 < thi) {
          const float c = (wxy * wl[kz]) * scale;
          const float2 d = src[kz];
          od = make_float2(d.x * c, d.y * c);
          if (normalise && k2 == 0) od = make_float2(0.0f, 0.0f);
          oi.x = 1.0f;
        }
      }
      dd[kz] = od;
      if (di) di[kz] = oi;
    }
  }
}

// host launcher for other translation units (grad_est.cu): masked, deconvolved delta_k of one shell
int launch_shell_filter(jps_plan* plan, int mas_order, int tlo, int thi, float* out, cudaStream_t s) {
  const int n = plan->n;
  ScopedLaunch L(K_SHELL_FILTER, s);
  const int blocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 16);
  shell_filter_kernel<<<blocks, 256, 0, s>>>(plan->dk, n, plan->nz, plan->pitch,
                                             plan->wlut + (size_t)(mas_order - 2) * n, 0, tlo, thi, (float2*)out, nullptr);
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

// out[0] += sum d3^2, out[1] += sum i3^2, and when `triple`: out[2] += sum d0 d1 d3,
// out[3] += sum i0 i1 i3.  Fields are [n][n][rowpitch] reals (in-place C2R layout).
__global__ void __launch_bounds__(256) triple_reduce_kernel(const float* __restrict__ d0,
                                                            const float* __restrict__ d1,
                                                            const float* __restrict__ d3,
                                                            const float* __restrict__ i0,
                                                            const float* __restrict__ i1,
                                                            const float* __restrict__ i3, int n,
                                                            int rowpitch, int triple, int with_ind,
                                                            double* __restrict__ out,
                                                            double* __restrict__ out_i2,
                                                            double* __restrict__ out_i3) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const long long rows = (long long)n * n;
  const int half = n / 2;                         // float2 loads (rows are 8-byte aligned, n even)
  const bool vec = (n % 2 == 0);
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const size_t base = (size_t)row * rowpitch;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;   // float32 partials per row segment
    if (vec) {
      for (int j = threadIdx.x; j < half; j += blockDim.x) {
        const float2 x3 = *reinterpret_cast<const float2*>(d3 + base + 2 * j);
        s0 += x3.x * x3.x + x3.y * x3.y;
        float2 y3 = make_float2(0.0f, 0.0f);
        if (with_ind) {
          y3 = *reinterpret_cast<const float2*>(i3 + base + 2 * j);
          s1 += y3.x * y3.x + y3.y * y3.y;
        }
        if (triple) {
          const float2 x0 = *reinterpret_cast<const float2*>(d0 + base + 2 * j);
          const float2 x1 = *reinterpret_cast<const float2*>(d1 + base + 2 * j);
          s2 += x0.x * x1.x * x3.x + x0.y * x1.y * x3.y;
          if (with_ind) {
            const float2 y0 = *reinterpret_cast<const float2*>(i0 + base + 2 * j);
            const float2 y1 = *reinterpret_cast<const float2*>(i1 + base + 2 * j);
            s3 += y0.x * y1.x * y3.x + y0.y * y1.y * y3.y;
          }
        }
      }
    } else {
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float x3 = d3[base + j];
        const float y3 = with_ind ? i3[base + j] : 0.0f;
        s0 += x3 * x3;
        s1 += y3 * y3;
        if (triple) {
          s2 += d0[base + j] * d1[base + j] * x3;
          if (with_ind) s3 += i0[base + j] * i1[base + j] * y3;
        }
      }
    }
    a0 += (double)s0; a1 += (double)s1; a2 += (double)s2; a3 += (double)s3;
  }
  // block reduction in float64
  __shared__ double red[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a0 += __shfl_down_sync(0xffffffffu, a0, off);
    a1 += __shfl_down_sync(0xffffffffu, a1, off);
    a2 += __shfl_down_sync(0xffffffffu, a2, off);
    a3 += __shfl_down_sync(0xffffffffu, a3, off);
  }
  if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = a2; red[3][warp] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
    // data sums go to the per-call scratch, indicator sums to their cached device slots
    if (threadIdx.x == 0) atomicAdd(out + 0, t);
    if (threadIdx.x == 2 && triple) atomicAdd(out + 1, t);
    if (threadIdx.x == 1 && with_ind) atomicAdd(out_i2, t);
    if (threadIdx.x == 3 && triple && with_ind) atomicAdd(out_i3, t);
  }
}

// scal layout: [2*j + 0..1] for shell j: sum d_j^2, sum d_0 d_1 d_j.  The indicator sums
// sum I_j^2 / sum I_0 I_1 I_j live in isum[slot2[j]] / isum[slot3[j]] (slots passed in `slots`:
// [nshell] pair slots then [nshell] triple slots).
__global__ void bispec_finalize_kernel(const double* __restrict__ scal, const double* __restrict__ isum,
                                       const int* __restrict__ slots, int nshell, double vol_p,
                                       double vol_b, float* __restrict__ pk, float* __restrict__ B,
                                       float* __restrict__ Q) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nshell) return;
  const double pj = scal[2 * j] / isum[slots[j]] * vol_p;
  pk[j] = (float)pj;
  if (j >= 2) {
    const double p0 = scal[0] / isum[slots[0]] * vol_p, p1 = scal[2] / isum[slots[1]] * vol_p;
    const double b = scal[2 * j + 1] / isum[slots[nshell + j]] * vol_b;
    B[j - 2] = (float)b;
    Q[j - 2] = (float)(b / (p0 * p1 + p0 * pj + p1 * pj));
  }
}


// Round-1 formulation, kept for A/B runs and as an independent cross-check (JPS_BISPEC=realspace): one
// filter + C2R + real-space reduce PER SHELL (bins + 2 inverse FFTs per call).
// Needs 6 shell fields: d0, d1, i0, i1 and a (d3, i3) pair.
static int bispec_from_dk_realspace(jps_plan* plan, int normalise, float box_size, float k1, float k2,
                   const float* theta, int nbins, int mas_order, float* k_all_out, float* pk_out,
                   float* B_out, float* Q_out, cudaStream_t s) {
  if (plan->n_shell_fields < 6) {
    set_error("bispec needs a plan created with n_shell_fields >= 6");
    return JPS_ERR_WORKSPACE;
  }
  const int nshell = nbins + 2;
  JPS_REQUIRE(nbins >= 1 && nshell <= kMaxShells, "bispec: number of theta bins %d out of range [1,%d]", nbins, kMaxShells - 2);
  const int n = plan->n;
  const float kF = ref_kF(box_size);
  // k_all = [k1, k2, k3(theta)...] in float32, :347-357 (Q19)
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
    tlo[(size_t)j] = (int)edge_threshold(lo, false, plan->k2max);     // |k| >= lo
    thi[(size_t)j] = (int)edge_threshold(hi, false, plan->k2max);     // |k| <  hi
  }
  JPS_CHECK_CUDA(cudaMemcpyAsync(k_all_out, k_all.data(), (size_t)nshell * 4, cudaMemcpyHostToDevice, s));
  // cached indicator sums: find / assign device slots
  if ((int)plan->isum_slot.size() + 2 * nshell > kIsumSlots) plan->isum_slot.clear();   // simple eviction
  std::vector<int> slots((size_t)2 * nshell, 0);
  std::vector<char> fresh2((size_t)nshell, 0), fresh3((size_t)nshell, 0);
  auto slot_of = [&](const std::vector<int>& key, char& fresh) {
    auto it = plan->isum_slot.find(key);
    if (it != plan->isum_slot.end()) { fresh = 0; return it->second; }
    const int sidx = (int)plan->isum_slot.size();
    plan->isum_slot[key] = sidx;
    fresh = 1;
    return sidx;
  };
  for (int j = 0; j < nshell; ++j) {
    slots[(size_t)j] = slot_of({tlo[(size_t)j], thi[(size_t)j]}, fresh2[(size_t)j]);
    if (j >= 2)
      slots[(size_t)nshell + j] = slot_of({tlo[0], thi[0], tlo[1], thi[1], tlo[(size_t)j], thi[(size_t)j]}, fresh3[(size_t)j]);
  }
  // a fresh triple needs I_0 and I_1 as fields: recompute them even if their own sums are cached
  bool any_fresh3 = false;
  for (int j = 2; j < nshell; ++j) any_fresh3 = any_fresh3 || fresh3[(size_t)j];
  for (int j = 0; j < nshell; ++j) {
    if (fresh2[(size_t)j]) JPS_CHECK_CUDA(cudaMemsetAsync(plan->isum + slots[(size_t)j], 0, 8, s));
    if (j >= 2 && fresh3[(size_t)j]) JPS_CHECK_CUDA(cudaMemsetAsync(plan->isum + slots[(size_t)nshell + j], 0, 8, s));
  }
  int* slots_dev = reinterpret_cast<int*>(plan->scal + 768);        // tail of the 1024-double scratch
  JPS_CHECK_CUDA(cudaMemcpyAsync(slots_dev, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  {
    ScopedLaunch L(K_MEMSET, s);
    JPS_CHECK_CUDA(cudaMemsetAsync(plan->scal, 0, 762 * sizeof(double), s));   // sums + dump; [768..) holds the slot table
  }
  const size_t field_floats = (size_t)n * n * 2 * plan->pitch;
  float* F[6];
  for (int i = 0; i < 6; ++i) F[i] = plan->shell + (size_t)i * field_floats;
  // F[0]=d0 F[1]=d1 F[2]=i0 F[3]=i1 F[4]=d3 F[5]=i3
  const float* wl = plan->wlut + (size_t)(mas_order - 2) * n;
  const int fblocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 16);
  const int rblocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 8);
  JPS_CHECK_CUFFT(cufftSetStream(plan->c2r, s));
  for (int j = 0; j < nshell; ++j) {
    float* dj = (j == 0) ? F[0] : (j == 1) ? F[1] : F[4];
    float* ij = (j == 0) ? F[2] : (j == 1) ? F[3] : F[5];
    // the indicator field of this shell is needed if one of its sums is not cached yet
    // (shells 0 and 1: also when any triple with them is new)
    const bool need_ind = fresh2[(size_t)j] || (j >= 2 && fresh3[(size_t)j]) || (j < 2 && any_fresh3);
    {
      ScopedLaunch L(K_SHELL_FILTER, s);
      shell_filter_kernel<<<fblocks, 256, 0, s>>>(plan->dk, n, plan->nz, plan->pitch, wl, normalise,
                                                  tlo[(size_t)j], thi[(size_t)j], (float2*)dj,
                                                  need_ind ? (float2*)ij : nullptr);
    }
    JPS_CHECK_LAUNCH();
    {
      ScopedLaunch L(K_FFT_C2R, s);
      JPS_CHECK_CUFFT(cufftExecC2R(plan->c2r, (cufftComplex*)dj, (cufftReal*)dj));
    }
    if (need_ind) {
      ScopedLaunch L(K_FFT_C2R, s);
      JPS_CHECK_CUFFT(cufftExecC2R(plan->c2r, (cufftComplex*)ij, (cufftReal*)ij));
    }
    {
      // with_ind: accumulate the indicator sums of this shell.  A cached pair sum must not be added
      // to again: route it to a dummy slot when only the triple is fresh.
      const int with_ind = need_ind && (fresh2[(size_t)j] || (j >= 2 && fresh3[(size_t)j])) ? 1 : 0;
      double* dump = plan->scal + 760;
      double* o2 = fresh2[(size_t)j] ? plan->isum + slots[(size_t)j] : dump;
      double* o3 = (j >= 2 && fresh3[(size_t)j]) ? plan->isum + slots[(size_t)nshell + j] : dump + 1;
      ScopedLaunch L(K_TRIPLE_REDUCE, s);
      triple_reduce_kernel<<<rblocks, 256, 0, s>>>(F[0], F[1], dj, F[2], F[3], ij, n, 2 * plan->pitch,
                                                   j >= 2 ? 1 : 0, with_ind, plan->scal + 2 * j, o2, o3);
    }
    JPS_CHECK_LAUNCH();
  }
  const float tp = box_size / (float)(n * n);                 // (box_size / dims**2)**3, :398
  const float tb = (box_size * box_size) / (float)((long long)n * n * n);   // (box_size**2 / dims**3)**3, :451
  {
    ScopedLaunch L(K_PK_FINALIZE, s);
    bispec_finalize_kernel<<<(nshell + 127) / 128, 128, 0, s>>>(plan->scal, plan->isum, slots_dev, nshell,
                                                               (double)(tp * tp * tp), (double)(tb * tb * tb),
                                                               pk_out, B_out, Q_out);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}


// ------------------------------------------------------------------ Fourier-space formulation (default)
// The triangle sums are evaluated where the shell masks live.  With unnormalised transforms, A = d_0 d_1
// (pointwise product of the two shell fields) and A^ = R2C(A):
//     sum_x d_0 d_1 d_j  = sum_x A(x) C2R(M_j D)(x) = sum_k M_j(k) Re[D(k) conj(A^(k))]      (Parseval)
//     sum_x d_j^2        = N^3 sum_k M_j(k) |D(k)|^2
// over the FULL k-space, i.e. over the stored half with weight 2 for 0 < kz < N/2 (D = deconvolved delta_k).
// The same holds for the indicator fields with D = 1.  So a call needs TWO inverse transforms (shells k1, k2),
// one product, ONE forward transform and one streaming pass over delta_k and A^ that serves every theta bin at
// once -- instead of bins + 2 inverse transforms and bins + 2 full real-space reductions
// (/root/reference/src/correlations.py:424-457 does two irfftn per theta bin).  Sums are float64; every
// quantity stays in the units of the real-space formulation (the indicator-sum cache is shared with it).
__global__ void __launch_bounds__(256) field_product_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            float* __restrict__ out, int n, int rowpitch) {
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const size_t base = (size_t)row * rowpitch;
    for (int j = threadIdx.x; j < n; j += blockDim.x) out[base + j] = a[base + j] * b[base + j];
  }
}

struct CrossParams {
  const float2* dk;        // delta_k (not deconvolved); unused when IND
  const float2* ahat;      // R2C of the product field (may be null: pair sums only)
  int n, nz, pitch;
  const float* wl;
  int normalise;
  int nshell;
  const int* tlo;          // device [nshell]
  const int* thi;
  double* out_pair;        // [j * pair_stride]: N^3 sum mult M_j |D|^2   (IND: N^3 sum mult M_j)
  double* out_triple;      // [j * triple_stride]: sum mult M_j Re[D conj(A^)]   (j >= 2)
  int pair_stride, triple_stride;
  const int* pair_slot;    // IND: device slot tables (index into out_pair / out_triple), -1 = do not store
  const int* triple_slot;
};

template <bool IND>
__global__ void __launch_bounds__(256) shell_cross_reduce_kernel(CrossParams P) {
  __shared__ int s_lo[kMaxShellsDev], s_hi[kMaxShellsDev];
  __shared__ double s_pair[kMaxShellsDev], s_trip[kMaxShellsDev];
  __shared__ int s_min, s_max;
  for (int j = threadIdx.x; j < P.nshell; j += blockDim.x) {
    s_lo[j] = P.tlo[j]; s_hi[j] = P.thi[j]; s_pair[j] = 0.0; s_trip[j] = 0.0;
  }
  if (threadIdx.x == 0) {
    int mn = 0x7fffffff, mx = 0;
    for (int j = 0; j < P.nshell; ++j) { mn = min(mn, P.tlo[j]); mx = max(mx, P.thi[j]); }
    s_min = mn; s_max = mx;
  }
  __syncthreads();
  const int n = P.n, nz = P.nz, mid = n / 2;
  float scale = 1.0f;
  if (!IND && P.normalise) scale = (float)((double)n * (double)n * (double)n / (double)P.dk[0].x);
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const int kx = ix > mid ? ix - n : ix, ky = iy > mid ? iy - n : iy;
    const int k2xy = kx * kx + ky * ky;
    if (k2xy >= s_max) continue;
    const float wxy = IND ? 1.0f : P.wl[ix] * P.wl[iy];
    const size_t base = (size_t)row * P.pitch;
    for (int kz = threadIdx.x; kz < nz; kz += blockDim.x) {
      const int k2 = k2xy + kz * kz;
      if (k2 < s_min || k2 >= s_max) continue;
      float2 d = make_float2(1.0f, 0.0f);
      if (!IND) {
        const float c = (wxy * P.wl[kz]) * scale;
        const float2 r = P.dk[base + kz];
        d = make_float2(r.x * c, r.y * c);
        if (P.normalise && k2 == 0) d = make_float2(0.0f, 0.0f);
      }
      const float mult = (kz > 0 && 2 * kz != n) ? 2.0f : 1.0f;
      const float pw = mult * (d.x * d.x + d.y * d.y);
      float cr = 0.0f;
      if (P.ahat) {
        const float2 a = P.ahat[base + kz];
        cr = mult * (d.x * a.x + d.y * a.y);               // Re[d conj(a)]
      }
      for (int j = 0; j < P.nshell; ++j) {
        if (k2 >= s_lo[j] && k2 < s_hi[j]) {
          atomicAdd(&s_pair[j], (double)pw);
          if (j >= 2 && P.ahat) atomicAdd(&s_trip[j], (double)cr);
        }
      }
    }
  }
  __syncthreads();
  const double n3 = (double)n * (double)n * (double)n;
  for (int j = threadIdx.x; j < P.nshell; j += blockDim.x) {
    if (IND) {
      if (P.pair_slot[j] >= 0 && s_pair[j] != 0.0) atomicAdd(P.out_pair + P.pair_slot[j], s_pair[j] * n3);
      if (j >= 2 && P.triple_slot[j] >= 0 && s_trip[j] != 0.0) atomicAdd(P.out_triple + P.triple_slot[j], s_trip[j]);
    } else {
      if (s_pair[j] != 0.0) atomicAdd(P.out_pair + (size_t)j * P.pair_stride, s_pair[j] * n3);
      if (j >= 2 && s_trip[j] != 0.0) atomicAdd(P.out_triple + (size_t)j * P.triple_stride, s_trip[j]);
    }
  }
}

// Which shells the fields F[0], F[1] (and their indicators F[2], F[3]) currently hold: lets a sweep over
// (k1, k2) pairs (jps_bispec_pairs) skip the filter + inverse transform of a shell the previous pair left
// behind (the row-major upper-triangle order keeps k1 for a whole run of pairs).
struct ShellCache {
  int lo[2] = {-1, -1}, hi[2] = {-1, -1};
  bool ind[2] = {false, false};
};

static int bispec_from_dk_fourier(jps_plan* plan, int normalise, float box_size, float k1, float k2,
                                  const float* theta, int nbins, int mas_order, float* k_all_out, float* pk_out,
                                  float* B_out, float* Q_out, ShellCache* cache, cudaStream_t s) {
  if (plan->n_shell_fields < 6 || !plan->r2c_ip_ok) {
    set_error("bispec needs a plan created with n_shell_fields >= 6");
    return JPS_ERR_WORKSPACE;
  }
  const int nshell = nbins + 2;
  JPS_REQUIRE(nbins >= 1 && nshell <= kMaxShells, "bispec: number of theta bins %d out of range [1,%d]", nbins, kMaxShells - 2);
  const int n = plan->n;
  const float kF = ref_kF(box_size);
  // k_all = [k1, k2, k3(theta)...] in float32, :347-357 (Q19)
  std::vector<float> k_all((size_t)nshell);
  k_all[0] = k1; k_all[1] = k2;
  for (int b = 0; b < nbins; ++b) {
    const float sn = k2 * sinf(theta[b]);
    const float cs = k2 * cosf(theta[b]) + k1;
    k_all[(size_t)b + 2] = sqrtf(sn * sn + cs * cs);
  }
  std::vector<int> tl((size_t)2 * nshell);        // [0, nshell): lower thresholds, [nshell, 2 nshell): upper
  int* tlo = tl.data();
  int* thi = tl.data() + nshell;
  for (int j = 0; j < nshell; ++j) {
    const float lo = (k_all[(size_t)j] - kF) / kF, hi = (k_all[(size_t)j] + kF) / kF;
    tlo[j] = (int)edge_threshold(lo, false, plan->k2max);     // |k| >= lo
    thi[j] = (int)edge_threshold(hi, false, plan->k2max);     // |k| <  hi
  }
  JPS_CHECK_CUDA(cudaMemcpyAsync(k_all_out, k_all.data(), (size_t)nshell * 4, cudaMemcpyHostToDevice, s));
  // Cached indicator sums: find the device slots of the known ones, RESERVE slots for the new ones.  The
  // reservations become visible in plan->isum_slot only after every launch of this call has been enqueued
  // without error: a failed call never leaves a key that points at a half-computed sum.
  if ((int)plan->isum_slot.size() + 2 * nshell > kIsumSlots) plan->isum_slot.clear();   // simple eviction
  std::map<std::vector<int>, int> reserved;
  std::vector<int> slots((size_t)2 * nshell, 0), fresh_pair((size_t)nshell, -1), fresh_triple((size_t)nshell, -1);
  auto slot_of = [&](const std::vector<int>& key, bool& fresh) {
    auto it = plan->isum_slot.find(key);
    if (it != plan->isum_slot.end()) { fresh = false; return it->second; }
    auto ir = reserved.find(key);
    if (ir != reserved.end()) { fresh = false; return ir->second; }        // same shell twice in one call
    const int sidx = (int)(plan->isum_slot.size() + reserved.size());
    reserved[key] = sidx;
    fresh = true;
    return sidx;
  };
  bool any_fresh3 = false, any_fresh2 = false;
  for (int j = 0; j < nshell; ++j) {
    bool f2 = false, f3 = false;
    slots[(size_t)j] = slot_of({tlo[j], thi[j]}, f2);
    if (f2) { fresh_pair[(size_t)j] = slots[(size_t)j]; any_fresh2 = true; }
    if (j >= 2) {
      slots[(size_t)nshell + j] = slot_of({tlo[0], thi[0], tlo[1], thi[1], tlo[j], thi[j]}, f3);
      if (f3) { fresh_triple[(size_t)j] = slots[(size_t)nshell + j]; any_fresh3 = true; }
    }
  }
  for (int j = 0; j < nshell; ++j) {
    if (fresh_pair[(size_t)j] >= 0) JPS_CHECK_CUDA(cudaMemsetAsync(plan->isum + fresh_pair[(size_t)j], 0, 8, s));
    if (fresh_triple[(size_t)j] >= 0) JPS_CHECK_CUDA(cudaMemsetAsync(plan->isum + fresh_triple[(size_t)j], 0, 8, s));
  }
  // device tables in the tail of the 1024-double scratch: [768..) ints: slots[2 nshell], then tlo/thi[2 nshell],
  // then fresh_pair[nshell], fresh_triple[nshell]   (6 * kMaxShells ints = 750 doubles' worth does NOT fit there,
  // so the tables live in the dump area of the accumulator array instead: plan->acc is unused by the bispectrum)
  int* tab = reinterpret_cast<int*>(plan->acc);
  std::vector<int> host_tab;
  host_tab.insert(host_tab.end(), slots.begin(), slots.end());
  host_tab.insert(host_tab.end(), tl.begin(), tl.end());
  host_tab.insert(host_tab.end(), fresh_pair.begin(), fresh_pair.end());
  host_tab.insert(host_tab.end(), fresh_triple.begin(), fresh_triple.end());
  JPS_CHECK_CUDA(cudaMemcpyAsync(tab, host_tab.data(), host_tab.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  const int* slots_dev = tab;
  const int* tlo_dev = tab + 2 * nshell;
  const int* thi_dev = tab + 3 * nshell;
  const int* fpair_dev = tab + 4 * nshell;
  const int* ftrip_dev = tab + 5 * nshell;
  {
    ScopedLaunch L(K_MEMSET, s);
    JPS_CHECK_CUDA(cudaMemsetAsync(plan->scal, 0, 762 * sizeof(double), s));
  }
  const size_t field_floats = (size_t)n * n * 2 * plan->pitch;
  float* F[6];
  for (int i = 0; i < 6; ++i) F[i] = plan->shell + (size_t)i * field_floats;
  // F[0]=d0 F[1]=d1 F[2]=i0 F[3]=i1 F[4]=d0*d1 -> A^ F[5]=i0*i1 -> I^
  const float* wl = plan->wlut + (size_t)(mas_order - 2) * n;
  const int fblocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 16);
  JPS_CHECK_CUFFT(cufftSetStream(plan->c2r, s));
  JPS_CHECK_CUFFT(cufftSetStream(plan->r2c_ip, s));
  for (int j = 0; j < 2; ++j) {
    const bool need_ind = any_fresh3;            // a new triple needs I_0 and I_1 as real-space fields
    if (cache && cache->lo[j] == tlo[j] && cache->hi[j] == thi[j] && (!need_ind || cache->ind[j])) continue;
    {
      ScopedLaunch L(K_SHELL_FILTER, s);
      shell_filter_kernel<<<fblocks, 256, 0, s>>>(plan->dk, n, plan->nz, plan->pitch, wl, normalise, tlo[j], thi[j],
                                                  (float2*)F[j], need_ind ? (float2*)F[2 + j] : nullptr);
    }
    JPS_CHECK_LAUNCH();
    {
      ScopedLaunch L(K_FFT_C2R, s);
      JPS_CHECK_CUFFT(cufftExecC2R(plan->c2r, (cufftComplex*)F[j], (cufftReal*)F[j]));
    }
    if (need_ind) {
      ScopedLaunch L(K_FFT_C2R, s);
      JPS_CHECK_CUFFT(cufftExecC2R(plan->c2r, (cufftComplex*)F[2 + j], (cufftReal*)F[2 + j]));
    }
    if (cache) { cache->lo[j] = tlo[j]; cache->hi[j] = thi[j]; cache->ind[j] = need_ind; }
  }
  const int rblocks = (int)std::min<long long>((long long)n * n, (long long)kNumSMs * 8);
  for (int pass = 0; pass < 2; ++pass) {         // 0: data fields, 1: indicator fields (only when something is new)
    if (pass == 1 && !(any_fresh2 || any_fresh3)) break;
    const bool want_triple = (pass == 0) || any_fresh3;
    float* prod = F[4 + pass];
    if (want_triple) {
      {
        ScopedLaunch L(K_TRIPLE_REDUCE, s);
        field_product_kernel<<<rblocks, 256, 0, s>>>(F[2 * pass], F[2 * pass + 1], prod, n, 2 * plan->pitch);
      }
      JPS_CHECK_LAUNCH();
      ScopedLaunch L(K_FFT_R2C, s);
      JPS_CHECK_CUFFT(cufftExecR2C(plan->r2c_ip, (cufftReal*)prod, (cufftComplex*)prod));
    }
    CrossParams C;
    C.dk = plan->dk; C.ahat = want_triple ? (const float2*)prod : nullptr;
    C.n = n; C.nz = plan->nz; C.pitch = plan->pitch; C.wl = wl; C.normalise = normalise; C.nshell = nshell;
    C.tlo = tlo_dev; C.thi = thi_dev;
    if (pass == 0) {
      C.out_pair = plan->scal; C.out_triple = plan->scal + 1; C.pair_stride = 2; C.triple_stride = 2;
      C.pair_slot = nullptr; C.triple_slot = nullptr;
      ScopedLaunch L(K_TRIPLE_REDUCE, s);
      shell_cross_reduce_kernel<false><<<rblocks, 256, 0, s>>>(C);
    } else {
      C.out_pair = plan->isum; C.out_triple = plan->isum; C.pair_stride = 0; C.triple_stride = 0;
      C.pair_slot = fpair_dev; C.triple_slot = ftrip_dev;
      ScopedLaunch L(K_TRIPLE_REDUCE, s);
      shell_cross_reduce_kernel<true><<<rblocks, 256, 0, s>>>(C);
    }
    JPS_CHECK_LAUNCH();
  }
  const float tp = box_size / (float)(n * n);                 // (box_size / dims**2)**3, :398
  const float tb = (box_size * box_size) / (float)((long long)n * n * n);   // (box_size**2 / dims**3)**3, :451
  {
    ScopedLaunch L(K_PK_FINALIZE, s);
    bispec_finalize_kernel<<<(nshell + 127) / 128, 128, 0, s>>>(plan->scal, plan->isum, slots_dev, nshell,
                                                               (double)(tp * tp * tp), (double)(tb * tb * tb),
                                                               pk_out, B_out, Q_out);
  }
  JPS_CHECK_LAUNCH();
  for (const auto& kv : reserved) plan->isum_slot[kv.first] = kv.second;      // everything enqueued: publish the new sums
  return JPS_OK;
}

int bispec_from_dk_cached(jps_plan* plan, int normalise, float box_size, float k1, float k2, const float* theta,
                          int nbins, int mas_order, float* k_all, float* pk, float* B, float* Q, ShellCache* cache,
                          cudaStream_t s) {
  static const bool realspace = [] { const char* e = getenv("JPS_BISPEC"); return e && !strcmp(e, "realspace"); }();
  if (realspace) return bispec_from_dk_realspace(plan, normalise, box_size, k1, k2, theta, nbins, mas_order, k_all, pk, B, Q, s);
  return bispec_from_dk_fourier(plan, normalise, box_size, k1, k2, theta, nbins, mas_order, k_all, pk, B, Q, cache, s);
}

// Bispectrum stage given plan->dk (entry point of the other translation units).
int bispec_from_dk(jps_plan* plan, int normalise, float box_size, float k1, float k2,
                   const float* theta, int nbins, int mas_order, float* k_all_out, float* pk_out,
                   float* B_out, float* Q_out, cudaStream_t s) {
  return bispec_from_dk_cached(plan, normalise, box_size, k1, k2, theta, nbins, mas_order, k_all_out, pk_out, B_out,
                               Q_out, nullptr, s);
}

// xi.cu
int xi_from_dk(jps_plan* plan, const BinTable& T, int normalise, int mas_order, int guard_mu, cudaStream_t s);
int xi_finalize(jps_plan* plan, const BinTable& T, float box_size, int nb, int first_bin, int fundamental,
                float* r3d, float* xi3d, float* nmodes, double* sums, int64_t* counts, cudaStream_t s);
void xi_grid_edges(const float* s_edges, int nb, float box_size, int n, std::vector<float>& out);
// powspec.cu
int pk_from_dk(jps_plan* plan, float box_size, const float* k_edges, int nb, int normalise, int mas_order,
               float shot_noise, float* k3d, float* pk3d, float* nmodes, double* sums, int64_t* counts,
               cudaStream_t s);

}  // namespace jps

using namespace jps;

extern "C" int jps_bispec(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                          float k1, float k2, const float* theta, int nbins, int mas_order,
                          float* k_all, float* pk, float* B, float* Q, void* stream) {
  JPS_REQUIRE(plan && mesh && theta && k_all && pk && B && Q, "jps_bispec: NULL argument");
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4, "jps_bispec: mas_order must be 2, 3 or 4");
  JPS_REQUIRE(box_size > 0.0f, "jps_bispec: box_size must be > 0");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = forward_fft(plan, mesh, s);
  if (rc) return rc;
  return bispec_from_dk(plan, normalise, box_size, k1, k2, theta, nbins, mas_order, k_all, pk, B, Q, s);
}

extern "C" int jps_bispec_pairs(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                                const float* k1, const float* k2, int npairs, const float* theta,
                                int nbins, int mas_order, float* k_all, float* pk, float* B, float* Q,
                                void* stream) {
  JPS_REQUIRE(plan && mesh && k1 && k2 && theta && k_all && pk && B && Q, "jps_bispec_pairs: NULL argument");
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4, "jps_bispec_pairs: mas_order must be 2, 3 or 4");
  JPS_REQUIRE(box_size > 0.0f && npairs >= 1, "jps_bispec_pairs: box_size must be > 0 and npairs >= 1");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = forward_fft(plan, mesh, s);                     // ONE rfftn for every pair
  if (rc) return rc;
  const size_t nshell = (size_t)nbins + 2;
  ShellCache cache;                                        // shells left in the fields by the previous pair
  for (int p = 0; p < npairs; ++p) {
    rc = bispec_from_dk_cached(plan, normalise, box_size, k1[p], k2[p], theta, nbins, mas_order,
                               k_all + (size_t)p * nshell, pk + (size_t)p * nshell, B + (size_t)p * (size_t)nbins,
                               Q + (size_t)p * (size_t)nbins, &cache, s);
    if (rc) return rc;
  }
  return JPS_OK;
}

extern "C" int jps_compute_2pt_correlations(jps_plan_t* plan, const float* mesh, int normalise,
                                            float box_size, const float* s_edges, int ns,
                                            const float* k_edges, int nk, int mas_order,
                                            float* k3d, float* pk3d, float* nmodes_pk, float* r3d,
                                            float* xi3d, float* nmodes_xi, void* stream) {
  JPS_REQUIRE(plan && mesh && s_edges && k_edges && k3d && pk3d && nmodes_pk && r3d && xi3d && nmodes_xi,
              "jps_compute_2pt_correlations: NULL argument");
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4, "jps_compute_2pt_correlations: mas_order must be 2, 3 or 4");
  JPS_REQUIRE(box_size > 0.0f && ns >= 1 && nk >= 1, "jps_compute_2pt_correlations: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = forward_fft(plan, mesh, s);                     // ONE rfftn shared by both estimators (:656)
  if (rc) return rc;
  rc = pk_from_dk(plan, box_size, k_edges, nk, normalise, mas_order, 0.0f, k3d, pk3d, nmodes_pk, nullptr, nullptr, s);
  if (rc) return rc;
  std::vector<float> kg;
  xi_grid_edges(s_edges, ns, box_size, plan->n, kg);
  BinTable* T = nullptr;
  rc = ensure_bin_table(plan, kg.data(), ns, TABLE_XI_EDGES, s, &T);
  if (rc) return rc;
  rc = xi_from_dk(plan, *T, normalise, mas_order, /*guard_mu=*/1, s);        // composites guard mu (:527)
  if (rc) return rc;
  return xi_finalize(plan, *T, box_size, ns, 0, 0, r3d, xi3d, nmodes_xi, nullptr, nullptr, s);
}

extern "C" int jps_compute_all_correlations(jps_plan_t* plan, const float* mesh, int normalise,
                                            float box_size, const float* s_edges, int ns,
                                            const float* k_edges, int nk, float k1, float k2,
                                            const float* theta, int nbins, int mas_order,
                                            float* k3d, float* pk3d, float* nmodes_pk, float* r3d,
                                            float* xi3d, float* nmodes_xi, float* k_all,
                                            float* pk_shell, float* B, float* Q, void* stream) {
  JPS_REQUIRE(theta && k_all && pk_shell && B && Q, "jps_compute_all_correlations: NULL argument");
  int rc = jps_compute_2pt_correlations(plan, mesh, normalise, box_size, s_edges, ns, k_edges, nk,
                                        mas_order, k3d, pk3d, nmodes_pk, r3d, xi3d, nmodes_xi, stream);
  if (rc) return rc;
  // delta_k from the shared forward FFT is still in the plan
  return bispec_from_dk(plan, normalise, box_size, k1, k2, theta, nbins, mas_order, k_all, pk_shell,
                        B, Q, (cudaStream_t)stream);
}
