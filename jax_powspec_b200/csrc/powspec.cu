// P(k) multipoles: bin table, the fused streaming binning kernel (K4) and the API calls.
//
// Replaces /root/reference/src/correlations.py:25-54 (window multiply, |delta_k|^2, mu,
// four jnp.histogram passes, normalisation) with ONE pass over delta_k.
//
// Design (B200): delta_k is read exactly once (8 B per stored mode, HBM bound).  All stored
// modes (+-kx, +-ky) and (+-ky, +-kx) at the same kz share |k|, mu, the window factor, the
// Legendre weights and the bin, so one warp owns a canonical pair (a <= b) = (|kx|,|ky|),
// streams its <= 8 rows along kz with coalesced 8-byte loads, and folds them in registers
// before any binning work.  Lanes that fall in the same bin are contiguous in kz (|k| is
// monotone in kz), so a shuffle-based segmented reduction leaves one partial sum per
// (warp, bin); those go to warp-private shared-memory accumulators with plain stores (no
// atomics: shared-memory float atomics are CAS loops on sm_100a), are merged per block and
// leave the SM as one float64 global red per (block, bin).
//
// Bin membership is decided with integers only: the float32 edges of jnp.histogram are
// converted on the host into thresholds on k^2 = kx^2+ky^2+kz^2 (the smallest integer m
// with sqrt_f32(m) >= edge), which reproduces searchsorted(kedges, sqrt_f32(k^2), 'right')
// bit for bit -> mode counts are exact.
#include "common.cuh"
#include "fold.cuh"

#include <algorithm>
#include <cmath>

namespace jps {

// ------------------------------------------------------------------ host: bin table
// smallest integer m in [0, k2max+1] with sqrtf(m) >= e   (strict: > e)
int64_t edge_threshold(float e, bool strict, int64_t k2max) {
  auto pass = [&](int64_t m) {
    const float k = sqrtf((float)m);          // exact conversion: m < 2^24 for n <= 4096
    return strict ? (k > e) : (k >= e);
  };
  if (std::isnan(e)) return k2max + 1;
  if (pass(0)) return 0;
  if (!pass(k2max)) return k2max + 1;
  int64_t lo = 0, hi = k2max;                 // pass(lo) false, pass(hi) true; sqrtf is monotone
  while (hi - lo > 1) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (pass(mid)) hi = mid; else lo = mid;
  }
  return hi;
}

struct PkParams {
  const float2* dk;
  int n, nz, pitch;
  const int32_t* lut;
  const float* wl;         // [n] window factor per array index for the chosen order
  int nbc;
  double* acc;             // [nbc][4]
  int normalise;
  int npairs;
  int herm;                // 1: weight stored modes with 0 < kz < N/2 twice (TABLE_PK_EDGES_HERM)
};

// K4: fold + bin.  Accumulation MODE (chosen by the number of reachable bins):
//   ACC_WARP   warp-private shared accumulators, plain += (no atomics)        nbc <= kMaxSmemBins
//   ACC_BLOCK  one shared accumulator set per CTA, shared atomics by the few segment heads
//   ACC_GLOBAL float64 global reds by the segment heads                      (huge bin counts)
template <int MODE>
__global__ void __launch_bounds__(256) pk_fold_bin_kernel(PkParams P) {
  extern __shared__ float sacc[];
  constexpr bool SMEM = (MODE != ACC_GLOBAL);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int nacc = P.nbc * 3;
  float* my = sacc + (MODE == ACC_WARP ? (size_t)warp * nacc : 0);
  if (MODE == ACC_WARP) {
    for (int i = lane; i < nacc; i += 32) my[i] = 0.0f;
    __syncwarp();
  } else if (MODE == ACC_BLOCK) {
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) sacc[i] = 0.0f;
    __syncthreads();
  }
  float scale2 = 1.0f;
  if (P.normalise) {
    // delta_k = rho_k * N^3 / rho_0 for k != 0 (tests/correlations.py:49-50 in Fourier space)
    const double dc = (double)P.dk[0].x;
    const double s = (double)P.n * (double)P.n * (double)P.n / dc;
    scale2 = (float)(s * s);
  }
  const int n = P.n, nz = P.nz;
  const size_t pitch = (size_t)P.pitch;
  for (int p = blockIdx.x * nwarps + warp; p < P.npairs; p += gridDim.x * nwarps) {
    const PairDecode ab = decode_pair(p);
    const RowSet rows = make_rows(ab.a, ab.b, n);
    const float wab = P.wl[ab.a] * P.wl[ab.b];      // (c(kx)*c(ky)), symmetric in a<->b
    const int k2ab = ab.a * ab.a + ab.b * ab.b;
    const float2* rp[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      rp[r] = P.dk + ((size_t)rows.ix[r] * n + rows.iy[r]) * pitch;
    constexpr int UNR = 1;                         // kz chunks in flight per lane (8 rows each); 2 measured no faster
    for (int kzb = 0; kzb < nz; kzb += 32 * UNR) {
      float2 d[UNR][8];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int kz = kzb + 32 * u + lane;
#pragma unroll
        for (int r = 0; r < 8; ++r)
          d[u][r] = (r < rows.nrows && kz < nz) ? __ldg(rp[r] + kz) : make_float2(0.0f, 0.0f);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int kz0 = kzb + 32 * u;
        if (kz0 >= nz) break;                      // warp-uniform
        const int kz = kz0 + lane;
        float v[3] = {0.0f, 0.0f, 0.0f};
        int cb = -2;
        if (kz < nz) {
          const int k2 = k2ab + kz * kz;
          cb = __ldg(P.lut + k2);
          const float c = wab * P.wl[kz];             // (c(kx)*c(ky))*c(kz), correlations.py:32
          float sum = 0.0f;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float re = d[u][r].x * c, im = d[u][r].y * c;   // delta_k *= correction (:39)
            sum += re * re + im * im;                             // (delta_k * conj(delta_k)).real (:40)
          }
          sum *= scale2;
          if (P.herm && kz > 0 && 2 * kz != n) sum *= 2.0f;   // its mirror image -k is not stored
          float mu2 = 0.0f;
          if (k2 > 0) mu2 = (float)(kz * kz) / (float)k2;   // mu = kz/|k| (:37-38), LOS = z (Q14)
          else if (P.normalise) sum = 0.0f;                 // delta_0 = 0 after rho/mean - 1
          v[0] = sum;
          v[1] = sum * (3.0f * mu2 - 1.0f) * 0.5f;
          v[2] = sum * (35.0f * mu2 * mu2 - 30.0f * mu2 + 3.0f) * 0.125f;
        }
        const int prev = __shfl_up_sync(0xffffffffu, cb, 1);
        const bool head = (lane == 0) || (cb != prev);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        segmented_reduce<3>(v, heads, lane);
        if (head && cb >= 0) {
          if (MODE == ACC_WARP) {
            float* a = my + cb * 3;
            a[0] += v[0]; a[1] += v[1]; a[2] += v[2];
          } else if (MODE == ACC_BLOCK) {
            float* a = my + cb * 3;
            atomicAdd(a + 0, v[0]); atomicAdd(a + 1, v[1]); atomicAdd(a + 2, v[2]);
          } else {
            double* a = P.acc + (size_t)cb * 4;
            atomicAdd(a + 0, (double)v[0]);
            atomicAdd(a + 1, (double)v[1]);
            atomicAdd(a + 2, (double)v[2]);
          }
        }
        if (MODE == ACC_WARP) __syncwarp();
      }
    }
  }
  if (SMEM) {
    __syncthreads();
    const int nsets = (MODE == ACC_WARP) ? nwarps : 1;
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
      double s = 0.0;
      for (int w = 0; w < nsets; ++w) s += (double)sacc[(size_t)w * nacc + i];
      if (s != 0.0) atomicAdd(P.acc + (size_t)(i / 3) * 4 + (i % 3), s);
    }
  }
}

// Geometry only (no delta_k): exact mode counts, sum of |k| and largest C-order flat index per
// bin (for the reference's k3D[.].set(...) in powspec_vec_fundamental, Q18).  Runs once per
// bin table, results cached in the plan.
struct CountParams {
  int full_grid;            // 0: half-space (P(k)); 1: full n^3 grid (xi): z index folded like x,y
  int herm;                 // 1: half-space array with Hermitian weights (same z multiplicity as the full grid)
  int n, nz;
  const int32_t* lut;
  unsigned long long* cnt;
  double* ksum;
  unsigned long long* lastidx;
  int npairs;
};

__global__ void __launch_bounds__(256) pk_count_kernel(CountParams P) {
  const int n = P.n, nz = P.nz;
  const long long total = (long long)P.npairs * nz;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t / nz);
    const int kz = (int)(t - (long long)p * nz);
    const PairDecode ab = decode_pair(p);
    const int k2 = ab.a * ab.a + ab.b * ab.b + kz * kz;
    const int cb = P.lut[k2];
    if (cb < 0) continue;
    const RowSet rows = make_rows(ab.a, ab.b, n);
    unsigned long long last = 0;
    for (int r = 0; r < rows.nrows; ++r) {
      const unsigned long long flat = ((unsigned long long)rows.ix[r] * n + rows.iy[r]) * nz + kz;
      last = flat > last ? flat : last;
    }
    // full grid: +-kz are both stored (except 0 and the Nyquist index)
    const int zmult = ((P.full_grid || P.herm) && kz > 0 && 2 * kz != n) ? 2 : 1;
    const unsigned long long mult = (unsigned long long)rows.nrows * zmult;
    atomicAdd(P.cnt + cb, mult);
    atomicAdd(P.ksum + cb, (double)mult * (double)sqrtf((float)k2));
    atomicMax(P.lastidx + cb, last);
  }
}

struct FinalizeParams {
  int nb;                       // rows written
  int first_bin;                // user bin of output row 0 (1 for the fundamental variant)
  const int32_t* bin_to_compact;
  const float* edges;           // grid units, may be null (fundamental)
  const double* acc;
  const unsigned long long* cnt;
  const double* ksum;
  const unsigned long long* lastidx;
  int n, nz;
  float kF;
  double vol;                   // (box/N^2)^3 evaluated in float32 on the host
  float shot_noise;
  int kmode;                    // 0: bin centres (powspec_vec, Q10); 1: reference .set quirk (Q18); 2: mean k
  float* k3d; float* pk3d; float* nmodes;
  double* sums; int64_t* counts;
};

__global__ void pk_finalize_kernel(FinalizeParams F) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= F.nb) return;
  const int bin = j + F.first_bin;
  const int c = F.bin_to_compact[bin];
  double s0 = 0, s2 = 0, s4 = 0, ks = 0;
  unsigned long long cnt = 0, last = 0;
  if (c >= 0) {
    s0 = F.acc[(size_t)c * 4 + 0]; s2 = F.acc[(size_t)c * 4 + 1]; s4 = F.acc[(size_t)c * 4 + 2];
    cnt = F.cnt[c]; ks = F.ksum[c]; last = F.lastidx[c];
  }
  const double nm = (double)(float)cnt;          // the reference's counts are float32 (Q8)
  // empty bins: 0/0 -> NaN, as the reference (Q11)
  F.pk3d[j * 3 + 0] = (float)(s0 / nm * F.vol - (double)F.shot_noise);
  F.pk3d[j * 3 + 1] = (float)(s2 / nm * 5.0 * F.vol);
  F.pk3d[j * 3 + 2] = (float)(s4 / nm * 9.0 * F.vol);
  F.nmodes[j] = (float)cnt;
  if (F.kmode == 0) {
    F.k3d[j] = (0.5f * (F.edges[bin + 1] + F.edges[bin])) * F.kF;      // correlations.py:54
  } else if (F.kmode == 1) {
    // k3D.at[k_index].set(k): the last stored mode of the bin in C order wins, then /Nmodes*kF
    float kl = 0.0f;
    if (cnt > 0) {
      const int kz = (int)(last % F.nz);
      const unsigned long long xy = last / F.nz;
      int iy = (int)(xy % F.n), ix = (int)(xy / F.n);
      const int mid = F.n / 2;
      const int kx = ix > mid ? ix - F.n : ix, ky = iy > mid ? iy - F.n : iy;
      kl = sqrtf((float)(kx * kx + ky * ky + kz * kz));
    }
    F.k3d[j] = kl / (float)cnt * F.kF;
  } else {
    F.k3d[j] = (float)(ks / (double)cnt) * F.kF;
  }
  if (F.sums) { F.sums[j * 3 + 0] = s0; F.sums[j * 3 + 1] = s2; F.sums[j * 3 + 2] = s4; }
  if (F.counts) F.counts[j] = (int64_t)cnt;
}

int npairs_for(int n) {
  const int m = n / 2 + 1;                     // |k| values 0..n/2
  return m * (m + 1) / 2;
}

// Find or build the k^2 -> bin table for (mode, edges).  Tables live in kNumTables LRU slots so
// that P(k), xi(s) and the fundamental variants can alternate without rebuilding.
int ensure_bin_table(jps_plan* plan, const float* kedges_grid, int nb, int mode, cudaStream_t s,
                     BinTable** out) {
  const bool user_edges = (mode == TABLE_PK_EDGES || mode == TABLE_XI_EDGES || mode == TABLE_PK_EDGES_HERM);
  std::vector<float> key;
  key.push_back((float)mode);
  if (user_edges) key.insert(key.end(), kedges_grid, kedges_grid + nb + 1);
  BinTable* slot = nullptr;
  for (int i = 0; i < kNumTables; ++i) {
    BinTable& T = plan->tables[i];
    if (T.valid && T.key.size() == key.size() &&
        std::memcmp(T.key.data(), key.data(), key.size() * sizeof(float)) == 0) {
      T.stamp = ++plan->stamp;
      *out = &T;
      return JPS_OK;
    }
  }
  for (int i = 0; i < kNumTables; ++i) {          // least recently used (or empty) slot
    BinTable& T = plan->tables[i];
    if (!slot || !T.valid || (slot->valid && T.stamp < slot->stamp)) slot = &T;
    if (!T.valid) break;
  }
  BinTable& T = *slot;
  T.valid = false;
  const int64_t k2max = plan->k2max;
  std::vector<int32_t> lut((size_t)k2max + 1, -1);
  if (user_edges) {
    JPS_REQUIRE(nb >= 1 && nb <= kMaxUserBins, "number of bins %d out of range [1,%d]", nb, kMaxUserBins);
    for (int i = 0; i < nb; ++i)
      JPS_REQUIRE(!(kedges_grid[i + 1] < kedges_grid[i]), "bin edges must be ascending");
    // thresholds: th[i] = min m with sqrtf(m) >= e_i; last edge inclusive -> strict threshold
    std::vector<int64_t> th((size_t)nb + 1);
    for (int i = 0; i < nb; ++i) th[(size_t)i] = edge_threshold(kedges_grid[i], false, k2max);
    th[(size_t)nb] = edge_threshold(kedges_grid[nb], true, k2max);
    for (int i = 0; i < nb; ++i) {
      // searchsorted(...,'right') puts m in the LAST bin whose lower edge it reaches
      const int64_t lo = th[(size_t)i];
      const int64_t hi = th[(size_t)i + 1];
      for (int64_t m = lo; m < hi && m <= k2max; ++m) lut[(size_t)m] = i;
    }
  } else {
    nb = jps_fundamental_nbins(plan->n) + 1;
    for (int64_t m = 0; m <= k2max; ++m) {
      const int b = (int)sqrtf((float)m);
      lut[(size_t)m] = (b < nb) ? b : -1;
    }
  }
  // compact the reachable bins
  std::vector<int32_t> b2c((size_t)nb, -1), c2b;
  {
    std::vector<char> seen((size_t)nb, 0);
    for (int64_t m = 0; m <= k2max; ++m) if (lut[(size_t)m] >= 0) seen[(size_t)lut[(size_t)m]] = 1;
    for (int i = 0; i < nb; ++i) if (seen[(size_t)i]) { b2c[(size_t)i] = (int32_t)c2b.size(); c2b.push_back(i); }
    for (int64_t m = 0; m <= k2max; ++m) if (lut[(size_t)m] >= 0) lut[(size_t)m] = b2c[(size_t)lut[(size_t)m]];
  }
  const int nbc = (int)c2b.size();
  if (nbc > plan->acc_cap) {
    set_error("%d reachable bins exceed the plan capacity %d", nbc, plan->acc_cap);
    return JPS_ERR_UNSUPPORTED;
  }
  JPS_CHECK_CUDA(cudaMemcpyAsync(T.lut, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, s));
  JPS_CHECK_CUDA(cudaMemcpyAsync(T.bin_to_compact, b2c.data(), b2c.size() * 4, cudaMemcpyHostToDevice, s));
  if (nbc) JPS_CHECK_CUDA(cudaMemcpyAsync(T.compact_to_bin, c2b.data(), c2b.size() * 4, cudaMemcpyHostToDevice, s));
  if (user_edges)
    JPS_CHECK_CUDA(cudaMemcpyAsync(T.edges, kedges_grid, (size_t)(nb + 1) * 4, cudaMemcpyHostToDevice, s));
  // pageable sources are staged before cudaMemcpyAsync returns, so the vectors may die here
  JPS_CHECK_CUDA(cudaMemsetAsync(T.cnt, 0, (size_t)plan->acc_cap * 8, s));
  JPS_CHECK_CUDA(cudaMemsetAsync(T.ksum, 0, (size_t)plan->acc_cap * 8, s));
  JPS_CHECK_CUDA(cudaMemsetAsync(T.lastidx, 0, (size_t)plan->acc_cap * 8, s));
  CountParams C;
  C.full_grid = (mode == TABLE_XI_EDGES || mode == TABLE_XI_FUNDAMENTAL) ? 1 : 0;
  C.herm = (mode == TABLE_PK_EDGES_HERM) ? 1 : 0;
  C.n = plan->n; C.nz = plan->nz; C.lut = T.lut; C.cnt = T.cnt; C.ksum = T.ksum;
  C.lastidx = T.lastidx; C.npairs = npairs_for(plan->n);
  {
    ScopedLaunch L(K_PK_COUNT, s);
    pk_count_kernel<<<kNumSMs * 8, 256, 0, s>>>(C);
  }
  JPS_CHECK_LAUNCH();
  // segment form of the lut (see BinTable): breakpoints where the value changes + entry point per integer |k|
  {
    std::vector<int32_t> bp, val;
    for (int64_t m = 0; m <= k2max; ++m)
      if (m == 0 || lut[(size_t)m] != lut[(size_t)m - 1]) { bp.push_back((int32_t)m); val.push_back(lut[(size_t)m]); }
    const int ncoarse = (int)sqrt((double)k2max) + 2;
    std::vector<int32_t> coarse((size_t)ncoarse, 0);
    int seg = 0, max_scan = 0;
    for (int r = 0; r < ncoarse; ++r) {
      const int64_t lo = (int64_t)r * r, hi = (int64_t)(r + 1) * (r + 1);
      while (seg + 1 < (int)bp.size() && bp[(size_t)seg + 1] <= lo) ++seg;
      coarse[(size_t)r] = seg;
      int scan = 0;
      for (int t = seg; t + 1 < (int)bp.size() && bp[(size_t)t + 1] < hi; ++t) ++scan;
      max_scan = std::max(max_scan, scan);
    }
    T.nseg = 0; T.ncoarse = ncoarse;
    if ((int)bp.size() <= kMaxSegments && max_scan <= 8) {
      T.nseg = (int)bp.size();
      JPS_CHECK_CUDA(cudaMemcpyAsync(T.seg_bp, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice, s));
      JPS_CHECK_CUDA(cudaMemcpyAsync(T.seg_val, val.data(), val.size() * 4, cudaMemcpyHostToDevice, s));
      JPS_CHECK_CUDA(cudaMemcpyAsync(T.coarse, coarse.data(), coarse.size() * 4, cudaMemcpyHostToDevice, s));
    }
  }
  T.key = key; T.mode = mode; T.nb = nb; T.nbc = nbc; T.valid = true; T.stamp = ++plan->stamp;
  *out = &T;
  return JPS_OK;
}

// ------------------------------------------------------------------ pencil decomposition of the R2C transform
// Batched 2-D transpose of complex64: in[b][r][c] -> out[b][c][r].  64 x 64 tiles through shared memory,
// 256 threads, 16 elements per thread: loads and stores are both 256-byte runs per warp row.
__global__ void __launch_bounds__(256) transpose_c64_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                            long long rows, long long cols, long long in_pitch) {
  __shared__ float2 tile[64][65];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;              // 8 rows of 32 lanes
  const long long c0 = (long long)blockIdx.x * 64, r0 = (long long)blockIdx.y * 64;
  const float2* src = in + (size_t)blockIdx.z * rows * in_pitch;
  float2* dst = out + (size_t)blockIdx.z * rows * cols;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r0 + ty + 8 * i;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long c = c0 + tx + 32 * h;
      if (r < rows && c < cols) tile[ty + 8 * i][tx + 32 * h] = __ldg(src + r * in_pitch + c);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long c = c0 + ty + 8 * i;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long r = r0 + tx + 32 * h;
      if (c < cols && r < rows) dst[c * rows + r] = tile[tx + 32 * h][ty + 8 * i];
    }
  }
}

// in[b][r][0..cols) with row pitch in_pitch (>= cols) -> out[b][c][r], dense
static int transpose_c64(const float2* in, float2* out, long long batch, long long rows, long long cols, long long in_pitch,
                         cudaStream_t s) {
  const long long gx = (cols + 63) / 64, gy = (rows + 63) / 64;
  JPS_REQUIRE(gy <= 65535 && batch <= 65535 && gx <= 2147483647LL, "transpose: grid too large");
  ScopedLaunch L(K_TRANSPOSE, s);
  transpose_c64_kernel<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)batch), 256, 0, s>>>(in, out, rows, cols, in_pitch);
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

// Real-to-complex "untangle" step fused with the first transpose.  The z-pass reads the n reals of a line as
// M = n/2 complex numbers z[m] = x[2m] + i x[2m+1] and transforms them with a C2C of length M (Z).  The spectrum of
// the real line follows from pairs (Z[k], Z[M-k]):
//     E = (Z[k] + conj Z[M-k]) / 2,  O = -i (Z[k] - conj Z[M-k]) / 2,  T = e^{-2 pi i k / n} O
//     X[k] = E + T,   X[M-k] = conj(E - T)            (k = 0 .. M/2;  Z[M] = Z[0];  X[M] comes from k = 0)
// cuFFT does this in a separate pass over the array (its R2C of 2048 reals = C2C-1024 + `postprocess_kernel`,
// 17 ms of a 26.7 ms z-pass at 2048^3); here it rides on the transpose that is needed anyway: a CTA loads a
// 64 (y) x 32 (k) tile and the mirrored tile (coalesced along k, forwards and backwards), and writes rows k and
// M-k of the transposed array out[x][kz][y] (coalesced along y).  One read and one write of the array instead of three.
__global__ void __launch_bounds__(256) r2c_untangle_transpose_kernel(const float2* __restrict__ Z, float2* __restrict__ out,
                                                                     const float2* __restrict__ tw, int n) {
  __shared__ float2 ta[64][33], tb[64][33];
  const int M = n / 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;              // 8 rows of 32 lanes
  const int k0 = blockIdx.x * 32, y0 = blockIdx.y * 64;
  const float2* src = Z + (size_t)blockIdx.z * n * M;                  // plane x: [y][M]
  float2* dst = out + (size_t)blockIdx.z * (size_t)(M + 1) * n;        // plane x: [kz][y]
  const int k = k0 + tx;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int y = y0 + ty + 8 * i;
    if (y < n && k <= M / 2) {
      ta[ty + 8 * i][tx] = __ldg(src + (size_t)y * M + k);
      tb[ty + 8 * i][tx] = __ldg(src + (size_t)y * M + (k == 0 ? 0 : M - k));
    }
  }
  __syncthreads();
  // thread (t = k index within the tile, lanes along y)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = ty + 8 * i;
    const int kk = k0 + t;
    if (kk > M / 2) continue;
    const float2 w = __ldg(tw + kk);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int yl = tx + 32 * h, y = y0 + yl;
      if (y >= n) continue;
      const float2 a = ta[yl][t], b = tb[yl][t];
      const float ex = 0.5f * (a.x + b.x), ey = 0.5f * (a.y - b.y);          // E = (a + conj b) / 2
      const float ox = 0.5f * (a.y + b.y), oy = -0.5f * (a.x - b.x);         // O = -i (a - conj b) / 2
      const float tr = w.x * ox - w.y * oy, ti = w.x * oy + w.y * ox;        // T = w O
      dst[(size_t)kk * n + y] = make_float2(ex + tr, ey + ti);
      if (M - kk != kk) dst[(size_t)(M - kk) * n + y] = make_float2(ex - tr, -(ey - ti));
    }
  }
}

int launch_r2c_untangle_transpose(const float2* Z, float2* out, const float2* tw, int n, int batch, cudaStream_t s) {
  const int M = n / 2;
  JPS_REQUIRE(n % 2 == 0 && batch >= 1 && batch <= 65535, "r2c untangle: bad sizes");
  ScopedLaunch L(K_TRANSPOSE, s);
  r2c_untangle_transpose_kernel<<<dim3((unsigned)((M / 2 + 1 + 31) / 32), (unsigned)((n + 63) / 64), (unsigned)batch), 256, 0, s>>>(
      Z, out, tw, n);
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

void host_r2c_twiddles(int n, std::vector<float2>& tw) {
  tw.resize((size_t)n / 4 + 1);
  for (int k = 0; k <= n / 4; ++k) {
    const double ang = -2.0 * M_PI * (double)k / (double)n;
    tw[(size_t)k] = make_float2((float)cos(ang), (float)sin(ang));
  }
}

// mesh[x][y][z] -> dk[kz][ky][kx]: C2C of length n/2 along z on the real lines read as complex pairs, untangle +
// transpose (y fastest), 1-D C2C along y, transpose (x fastest), 1-D C2C along x.  Every cuFFT pass is a contiguous
// batched C2C (one read + one write of the array at ~6.5 TB/s, measured), where the monolithic 3-D plan and the
// strided 1-D plans need the equivalent of 7-8 passes at 2048^3.
static int forward_fft_pencil(jps_plan* plan, const float* mesh, cudaStream_t s) {
  const long long n = plan->n, nz = plan->nz;
  JPS_REQUIRE(((uintptr_t)mesh & 7) == 0, "jps_powspec: the mesh must be 8-byte aligned for a JPS_PLAN_FFT_PENCIL plan");
  JPS_CHECK_CUFFT(cufftSetStream(plan->fz, s));
  JPS_CHECK_CUFFT(cufftSetStream(plan->fy, s));
  JPS_CHECK_CUFFT(cufftSetStream(plan->fx, s));
  {
    ScopedLaunch L(K_FFT_R2C, s);                                                   // Z[x][y][n/2]
    JPS_CHECK_CUFFT(cufftExecC2C(plan->fz, (cufftComplex*)const_cast<float*>(mesh), (cufftComplex*)plan->dk, CUFFT_FORWARD));
  }
  int rc0 = launch_r2c_untangle_transpose(plan->dk, plan->dk2, plan->ztw, plan->n, plan->n, s);         // [x][kz][y]
  if (rc0) return rc0;
  {
    ScopedLaunch L(K_FFT_C2C_Y, s);
    JPS_CHECK_CUFFT(cufftExecC2C(plan->fy, (cufftComplex*)plan->dk2, (cufftComplex*)plan->dk2, CUFFT_FORWARD));
  }
  int rc = transpose_c64(plan->dk2, plan->dk, 1, n, nz * n, nz * n, s);             // [kz][y][x]
  if (rc) return rc;
  {
    ScopedLaunch L(K_FFT_C2C_X, s);
    JPS_CHECK_CUFFT(cufftExecC2C(plan->fx, (cufftComplex*)plan->dk, (cufftComplex*)plan->dk, CUFFT_FORWARD));
  }
  return JPS_OK;
}

int forward_fft(jps_plan* plan, const float* mesh, cudaStream_t s) {
  if (plan->pencil) return forward_fft_pencil(plan, mesh, s);
  if (!plan->r2c_ok) {
    set_error("this plan was created with JPS_PLAN_TABLES_ONLY: it has no 3-D FFT");
    return JPS_ERR_UNSUPPORTED;
  }
  JPS_CHECK_CUFFT(cufftSetStream(plan->r2c, s));
  ScopedLaunch L(K_FFT_R2C, s);
  JPS_CHECK_CUFFT(cufftExecR2C(plan->r2c, (cufftReal*)mesh, (cufftComplex*)plan->dk));
  return JPS_OK;
}

// fold + bin plan->dk into plan->acc with table T
static int bin_from_dk(jps_plan* plan, const BinTable& T, int normalise, int mas_order, cudaStream_t s) {
  const int nbc = T.nbc;
  {
    ScopedLaunch L(K_MEMSET, s);
    JPS_CHECK_CUDA(cudaMemsetAsync(plan->acc, 0, (size_t)std::max(nbc, 1) * 4 * 8, s));
  }
  if (nbc == 0) return JPS_OK;
  if (plan->pencil) {                            // spectrum stored [kz][ky][kx]: the x-fastest binning kernel
    JPS_REQUIRE(T.mode == TABLE_PK_EDGES, "a JPS_PLAN_FFT_PENCIL plan serves jps_powspec / jps_paint_powspec only");
    return bin_xfast_layout(plan, plan->dk, plan->n, 0, 1, T, reinterpret_cast<const float*>(plan->dk), normalise, mas_order, s);
  }
  PkParams P;
  P.dk = plan->dk; P.n = plan->n; P.nz = plan->nz; P.pitch = plan->pitch; P.lut = T.lut;
  P.wl = plan->wlut + (size_t)(mas_order - 2) * plan->n;
  P.nbc = nbc; P.acc = plan->acc; P.normalise = normalise; P.npairs = npairs_for(plan->n);
  P.herm = (T.mode == TABLE_PK_EDGES_HERM) ? 1 : 0;
  const int threads = 256, warps = threads / 32;
  const int want = (P.npairs + warps - 1) / warps;
  static PerDeviceFlag attr_set;
  if (!attr_set.get()) {
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_fold_bin_kernel<ACC_WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((size_t)warps * kMaxSmemBins * 3 * sizeof(float))));
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_fold_bin_kernel<ACC_BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((size_t)kMaxBlockBins * 3 * sizeof(float))));
    attr_set.set();
  }
  int per_sm = 1;
  if (nbc <= kMaxSmemBins) {
    const size_t smem = (size_t)warps * nbc * 3 * sizeof(float);
    JPS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk_fold_bin_kernel<ACC_WARP>, threads, smem));
    const int blocks = std::min(want, kNumSMs * std::max(per_sm, 1));
    ScopedLaunch L(K_PK_FOLD_BIN, s);
    pk_fold_bin_kernel<ACC_WARP><<<blocks, threads, smem, s>>>(P);
  } else if (nbc <= kMaxBlockBins) {
    const size_t smem = (size_t)nbc * 3 * sizeof(float);
    JPS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk_fold_bin_kernel<ACC_BLOCK>, threads, smem));
    const int blocks = std::min(want, kNumSMs * std::max(per_sm, 1));
    ScopedLaunch L(K_PK_FOLD_BIN, s);
    pk_fold_bin_kernel<ACC_BLOCK><<<blocks, threads, smem, s>>>(P);
  } else {
    const int blocks = std::min(want, kNumSMs * 8);
    ScopedLaunch L(K_PK_FOLD_BIN, s);
    pk_fold_bin_kernel<ACC_GLOBAL><<<blocks, threads, 0, s>>>(P);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

int bin_from_dk_public(jps_plan* plan, const BinTable& T, int normalise, int mas_order, cudaStream_t s) {
  return bin_from_dk(plan, T, normalise, mas_order, s);
}

double ref_volume(float box_size, int n) {
  const float t = box_size / (float)(n * n);    // (box_size/dims**2)**3 in float32, :50
  return (double)(t * t * t);
}

float ref_kF(float box_size) { return (float)(2.0 * M_PI) / box_size; }   // :12

int pk_from_dk(jps_plan* plan, float box_size, const float* k_edges, int nb, int normalise,
               int mas_order, float shot_noise, float* k3d, float* pk3d, float* nmodes,
               double* sums, int64_t* counts, cudaStream_t s);

static void fill_finalize(FinalizeParams& F, jps_plan* plan, const BinTable& T) {
  F.bin_to_compact = T.bin_to_compact; F.edges = T.edges; F.acc = plan->acc; F.cnt = T.cnt;
  F.ksum = T.ksum; F.lastidx = T.lastidx; F.n = plan->n; F.nz = plan->nz;
}

}  // namespace jps

using namespace jps;

extern "C" int jps_fundamental_nbins(int n_mesh) {
  const int mid = n_mesh / 2;
  return (int)sqrtf((float)(3 * mid * mid));   // jnp.int32(jnp.sqrt(3*middle**2)), :68
}

extern "C" int jps_powspec(jps_plan_t* plan, const float* mesh, int normalise, float box_size,
                           const float* k_edges, int nb, int mas_order, float shot_noise,
                           float* k3d, float* pk3d, float* nmodes, double* sums, int64_t* counts,
                           void* stream) {
  JPS_REQUIRE(plan && mesh && k_edges && k3d && pk3d && nmodes, "jps_powspec: NULL argument");
  JPS_REQUIRE(nb >= 1 && nb <= kMaxUserBins, "jps_powspec: nb=%d out of range", nb);
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4, "jps_powspec: mas_order must be 2, 3 or 4");
  JPS_REQUIRE(box_size > 0.0f, "jps_powspec: box_size must be > 0");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = forward_fft(plan, mesh, s);
  if (rc) return rc;
  return pk_from_dk(plan, box_size, k_edges, nb, normalise, mas_order, shot_noise, k3d, pk3d, nmodes,
                    sums, counts, s);
}

namespace jps {
// P(k) stage given plan->dk (forward FFT already done).
int pk_from_dk_mode(jps_plan* plan, float box_size, const float* k_edges, int nb, int normalise,
                    int mas_order, float shot_noise, int table_mode, float* k3d, float* pk3d, float* nmodes,
                    double* sums, int64_t* counts, cudaStream_t s);

int pk_from_dk(jps_plan* plan, float box_size, const float* k_edges, int nb, int normalise,
               int mas_order, float shot_noise, float* k3d, float* pk3d, float* nmodes,
               double* sums, int64_t* counts, cudaStream_t s) {
  return pk_from_dk_mode(plan, box_size, k_edges, nb, normalise, mas_order, shot_noise, TABLE_PK_EDGES, k3d, pk3d,
                         nmodes, sums, counts, s);
}

// table_mode: TABLE_PK_EDGES (the reference's half-space counting) or TABLE_PK_EDGES_HERM
int pk_from_dk_mode(jps_plan* plan, float box_size, const float* k_edges, int nb, int normalise,
                    int mas_order, float shot_noise, int table_mode, float* k3d, float* pk3d, float* nmodes,
                    double* sums, int64_t* counts, cudaStream_t s) {
  const float kF = ref_kF(box_size);
  std::vector<float> kg((size_t)nb + 1);
  for (int i = 0; i <= nb; ++i) kg[(size_t)i] = k_edges[i] / kF;       // kedges = k_edges / kF, :42 (Q9)
  BinTable* T = nullptr;
  int rc = ensure_bin_table(plan, kg.data(), nb, table_mode, s, &T);
  if (rc) return rc;
  rc = bin_from_dk(plan, *T, normalise, mas_order, s);
  if (rc) return rc;
  FinalizeParams F;
  fill_finalize(F, plan, *T);
  F.nb = nb; F.first_bin = 0; F.kF = kF; F.vol = ref_volume(box_size, plan->n);
  F.shot_noise = shot_noise; F.kmode = 0;
  F.k3d = k3d; F.pk3d = pk3d; F.nmodes = nmodes; F.sums = sums; F.counts = counts;
  {
    ScopedLaunch L(K_PK_FINALIZE, s);
    pk_finalize_kernel<<<(nb + 127) / 128, 128, 0, s>>>(F);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}
}  // namespace jps

namespace jps {
// ------------------------------------------------------------------ interlacing (no reference counterpart)
// dk <- (dk + dk2 * exp(-i pi (kx + ky + kz) / N)) / 2, where dk2 is the transform of the SAME particles
// painted on the grid displaced by +half a cell (xmin + cell/2): the images k + 2 k_N m with odd
// m_x + m_y + m_z cancel (Sefusatti et al. 2016).  One streaming pass, 16 B read + 8 B written per mode.
// Frequencies follow the reference's map (index N/2 -> +N/2, Q15); the phase is evaluated with
// sincospif on the integer sum reduced mod 2N, i.e. to float32 rounding.
__global__ void __launch_bounds__(256) interlace_combine_kernel(float2* __restrict__ dk,
                                                                const float2* __restrict__ dk2, int n, int nz,
                                                                int pitch) {
  const int mid = n / 2;
  const long long rows = (long long)n * n;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int iy = (int)(row % n), ix = (int)(row / n);
    const int kxy = (ix > mid ? ix - n : ix) + (iy > mid ? iy - n : iy);
    float2* a = dk + (size_t)row * pitch;
    const float2* b = dk2 + (size_t)row * pitch;
    for (int kz = threadIdx.x; kz < nz; kz += blockDim.x) {
      int sidx = (kxy + kz) % (2 * n);
      if (sidx < 0) sidx += 2 * n;
      float sn, cs;
      sincospif((float)sidx / (float)n, &sn, &cs);           // exp(-i pi s/N) = cs - i sn
      const float2 u = a[kz], v = b[kz];
      float2 o;
      o.x = 0.5f * (u.x + (v.x * cs + v.y * sn));
      o.y = 0.5f * (u.y + (v.y * cs - v.x * sn));
      a[kz] = o;
    }
  }
}

}  // namespace jps

extern "C" int jps_powspec_ex(jps_plan_t* plan, const float* mesh, const float* mesh2, int normalise,
                              float box_size, const float* k_edges, int nb, int mas_order, float shot_noise,
                              int flags, float* k3d, float* pk3d, float* nmodes, double* sums,
                              int64_t* counts, void* stream) {
  JPS_REQUIRE(plan && mesh && k_edges && k3d && pk3d && nmodes, "jps_powspec_ex: NULL argument");
  JPS_REQUIRE(nb >= 1 && nb <= kMaxUserBins, "jps_powspec_ex: nb=%d out of range", nb);
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4, "jps_powspec_ex: mas_order must be 2, 3 or 4");
  JPS_REQUIRE(box_size > 0.0f, "jps_powspec_ex: box_size must be > 0");
  JPS_REQUIRE((flags & ~JPS_PK_HERMITIAN) == 0, "jps_powspec_ex: unknown flags 0x%x", flags);
  cudaStream_t s = (cudaStream_t)stream;
  if (mesh2 != nullptr) {
    if (plan->n_shell_fields < 1) {
      set_error("jps_powspec_ex: interlacing needs a plan created with n_shell_fields >= 1");
      return JPS_ERR_WORKSPACE;
    }
    // second spectrum into shell field 0 (same bytes as delta_k), with the plan's out-of-place R2C
    JPS_CHECK_CUFFT(cufftSetStream(plan->r2c, s));
    {
      ScopedLaunch L(K_FFT_R2C, s);
      JPS_CHECK_CUFFT(cufftExecR2C(plan->r2c, (cufftReal*)mesh2, (cufftComplex*)plan->shell));
    }
  }
  int rc = forward_fft(plan, mesh, s);
  if (rc) return rc;
  if (mesh2 != nullptr) {
    const int blocks = (int)std::min<long long>((long long)plan->n * plan->n, (long long)kNumSMs * 16);
    ScopedLaunch L(K_INTERLACE, s);
    interlace_combine_kernel<<<blocks, 256, 0, s>>>(plan->dk, (const float2*)plan->shell, plan->n, plan->nz,
                                                   plan->pitch);
  }
  JPS_CHECK_LAUNCH();
  return pk_from_dk_mode(plan, box_size, k_edges, nb, normalise, mas_order, shot_noise,
                         (flags & JPS_PK_HERMITIAN) ? TABLE_PK_EDGES_HERM : TABLE_PK_EDGES, k3d, pk3d, nmodes,
                         sums, counts, s);
}

extern "C" int jps_powspec_fundamental(jps_plan_t* plan, const float* mesh, int normalise,
                                       float box_size, int mas_order, int compat, float* k3d,
                                       float* pk3d, float* nmodes, double* sums, int64_t* counts,
                                       void* stream) {
  JPS_REQUIRE(plan && mesh && k3d && pk3d && nmodes, "jps_powspec_fundamental: NULL argument");
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4, "jps_powspec_fundamental: mas_order must be 2, 3 or 4");
  JPS_REQUIRE(box_size > 0.0f, "jps_powspec_fundamental: box_size must be > 0");
  cudaStream_t s = (cudaStream_t)stream;
  BinTable* T = nullptr;
  int rc = ensure_bin_table(plan, nullptr, 0, TABLE_PK_FUNDAMENTAL, s, &T);
  if (rc) return rc;
  rc = forward_fft(plan, mesh, s);
  if (rc) return rc;
  rc = bin_from_dk(plan, *T, normalise, mas_order, s);
  if (rc) return rc;
  const int kmax = jps_fundamental_nbins(plan->n);
  if (kmax < 1) return JPS_OK;
  FinalizeParams F;
  fill_finalize(F, plan, *T);
  F.edges = nullptr;
  F.nb = kmax; F.first_bin = 1; F.kF = ref_kF(box_size); F.vol = ref_volume(box_size, plan->n);
  F.shot_noise = 0.0f; F.kmode = (compat == JPS_COMPAT_REFERENCE) ? 1 : 2;
  F.k3d = k3d; F.pk3d = pk3d; F.nmodes = nmodes; F.sums = sums; F.counts = counts;
  {
    ScopedLaunch L(K_PK_FINALIZE, s);
    pk_finalize_kernel<<<(kmax + 127) / 128, 128, 0, s>>>(F);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

extern "C" int jps_paint_powspec(jps_plan_t* plan, const float* x, const float* y, const float* z,
                                 const float* w, int64_t stride, int64_t n_part, float xmin,
                                 float ymin, float zmin, float box_size, int order, int wrap,
                                 int compat, int method, const float* k_edges, int nb,
                                 float shot_noise, float* mesh, void* paint_workspace,
                                 size_t paint_workspace_bytes, float* k3d, float* pk3d,
                                 float* nmodes, double* sums, int64_t* counts, void* stream) {
  JPS_REQUIRE(plan && mesh, "jps_paint_powspec: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n3 = (size_t)plan->n * plan->n * plan->n;
  {
    ScopedLaunch L(K_MEMSET, s);
    JPS_CHECK_CUDA(cudaMemsetAsync(mesh, 0, n3 * sizeof(float), s));
  }
  int rc = jps_paint(plan->n, x, y, z, w, stride, n_part, xmin, ymin, zmin, box_size, order, wrap,
                     compat, JPS_VARIANT_VEC, method, mesh, paint_workspace, paint_workspace_bytes,
                     stream);
  if (rc) return rc;
  return jps_powspec(plan, mesh, 1, box_size, k_edges, nb, order, shot_noise, k3d, pk3d, nmodes,
                     sums, counts, stream);
}
