// Slab-sharded mesh: the per-rank compute stages of the distributed paint -> FFT -> P(k) path
// (the reference has no multi-device code).
//
// Decomposition: rank r of P owns x-planes [r*N/P, (r+1)*N/P).  The forward transform is
//   (1) jps_slab_fft_yz : batched 2-D R2C over (y,z) of the owned planes   [nxl][N][N] -> [nxl][N][nz]
//   (2) jps_slab_pack   : regroup by destination rank                       -> [P][nxl][nyl][nz]
//       + ONE all-to-all of (P-1)/P^2 * 8 N^2 nz bytes per rank (host side: torch.distributed / NCCL)
//   (3) jps_slab_fft_x  : batched 1-D C2C along x, in place: strided on [N][nyl][nz] (x-slow layout) or
//                         contiguous on [nyl][nz][N] (x-fast layout, jps_slab_set_layout)
// and the spectrum stays y-sharded: the binning kernel only needs each element's (kx,ky,kz).
//   (4) jps_slab_powspec_partial : fold +-kx, window, Legendre weights, k-bin sums of the local shard
//       + allreduce of nb*3 float64 (host side), then jps_slab_powspec_finalize.
// Exact mode counts are geometry only and are computed identically on every rank.
#include "common.cuh"
#include "fold.cuh"

#include <algorithm>
#include <cstdlib>
#include <cmath>

constexpr int kMaxRanks = 64;
// Pieces the owned planes are cut into for the FFT / transfer overlap: the 2-D transform of piece c+1 runs
// while the peer stores of piece c drain, so only the FIRST piece's transform is exposed.  8 pieces of 32
// planes at 2048^3 on 8 GPUs (JPS_SLAB_CHUNKS overrides; round 1 used 4).
static int slab_chunks() {
  static const int k = [] { const char* e = getenv("JPS_SLAB_CHUNKS"); const int v = e ? atoi(e) : 0; return v >= 1 ? v : 8; }();
  return k;
}

static int chunk_planes_for(int nxl) {
  for (int k = slab_chunks(); k >= 2; k >>= 1)
    if (nxl % k == 0 && nxl / k >= 2) return nxl / k;
  return 0;
}

struct jps_slab_plan {
  int n = 0, nz = 0, nranks = 1, rank = 0, nxl = 0, nyl = 0;
  void** peer_dev = nullptr;        // device copy of the peers' receive-buffer pointers (pack_p2p)
  void* peer_host[kMaxRanks] = {};  // last pointers uploaded
  cufftHandle fft_yz = 0, fft_x = 0;
  bool yz_ok = false, x_ok = false;
  // A second 2-D plan over a quarter of the owned planes: the transform of chunk c+1 overlaps the
  // NVLink transfer of chunk c (jps_slab_fft_yz_planes / jps_slab_pack_p2p_planes on two streams).
  cufftHandle fft_yz_chunk = 0;
  bool yzc_ok = false;
  int chunk_planes = 0;
  // Layout of the transposed shard: 0 = [x][yl][kz] (x slowest; strided 1-D FFT, two passes over the
  // shard inside cuFFT), 1 = [yl][kz][x] (x fastest; the peer-store kernel transposes 32x32 tiles on
  // the way, the 1-D FFT is contiguous and the binning kernel runs its lanes along kx).
  int xfast = 0;
  cufftHandle fft_x_contig = 0;
  bool xc_ok = false;
  void* work = nullptr;
  size_t work_bytes = 0;
  // "Pencil" form of the 2-D transform of the owned planes (x-fast layout, even n): C2C of length n/2 along z on the
  // real lines read as complex pairs -> untangle + transpose (y fastest) -> contiguous C2C along y.  Three streaming
  // passes at 5-6.5 TB/s instead of cuFFT's 2-D R2C plan (24.3 ms for 1024 planes of 2048^2 on 2 GPUs = 2.3x the 2-pass
  // model).  The transformed planes are then [xl][kz][y] and the peer-store kernel transposes (xl <-> y) on the way.
  bool pencil_ok = false;         // the pencil plans exist
  bool use_pencil = false;        // ... and are used (default rule, JPS_SLAB_FFT, or jps_slab_set_layout)
  cufftHandle pz = 0, py = 0, pz_chunk = 0, py_chunk = 0;
  bool pz_ok = false, py_ok = false, pzc_ok = false, pyc_ok = false;
  float2* zbuf = nullptr;         // [nxl][n][n/2] complex: output of the z-pass
  float2* ztw = nullptr;          // untangle twiddles
  jps_plan* tables = nullptr;     // bin tables + accumulators (no 3-D FFT inside)
};

namespace jps {

static int make_fft_yz(int n, int nxl, cufftHandle* h, size_t* work) {
  JPS_CHECK_CUFFT(cufftCreate(h));
  JPS_CHECK_CUFFT(cufftSetAutoAllocation(*h, 0));
  const int nz = n / 2 + 1;
  long long dims[2] = {n, n};
  long long inembed[2] = {n, n};
  long long onembed[2] = {n, nz};
  JPS_CHECK_CUFFT(cufftMakePlanMany64(*h, 2, dims, inembed, 1, (long long)n * n, onembed, 1,
                                      (long long)n * nz, CUFFT_R2C, nxl, work));
  return JPS_OK;
}

static int make_fft_x(int n, int nyl, cufftHandle* h, size_t* work) {
  JPS_CHECK_CUFFT(cufftCreate(h));
  JPS_CHECK_CUFFT(cufftSetAutoAllocation(*h, 0));
  const int nz = n / 2 + 1;
  long long dims[1] = {n};
  long long embed[1] = {n};
  const long long stride = (long long)nyl * nz;      // distance between consecutive x
  JPS_CHECK_CUFFT(cufftMakePlanMany64(*h, 1, dims, embed, stride, 1, embed, stride, 1, CUFFT_C2C,
                                      stride, work));
  return JPS_OK;
}

static int make_fft_x_contig(int n, int nyl, cufftHandle* h, size_t* work) {
  JPS_CHECK_CUFFT(cufftCreate(h));
  JPS_CHECK_CUFFT(cufftSetAutoAllocation(*h, 0));
  const int nz = n / 2 + 1;
  long long dims[1] = {n};
  long long embed[1] = {n};
  JPS_CHECK_CUFFT(cufftMakePlanMany64(*h, 1, dims, embed, 1, n, embed, 1, n, CUFFT_C2C, (long long)nyl * nz, work));
  return JPS_OK;
}

static int make_c2c_contig(int len, long long batch, cufftHandle* h, size_t* work) {
  JPS_CHECK_CUFFT(cufftCreate(h));
  JPS_CHECK_CUFFT(cufftSetAutoAllocation(*h, 0));
  long long dims[1] = {len};
  long long embed[1] = {len};
  JPS_CHECK_CUFFT(cufftMakePlanMany64(*h, 1, dims, embed, 1, len, embed, 1, len, CUFFT_C2C, batch, work));
  return JPS_OK;
}

// Pencil form or cuFFT's batched 2-D R2C plan for the owned planes?  Measured at 2048^3: 1024 planes per rank (2 GPUs)
// 16.8 ms against 24.3 ms; 256 planes per rank (8 GPUs) 4.22 ms against 4.11 ms, and the peer-store kernel that reads
// the pencil layout is 4 % slower (its remote 256-byte runs land 16 MB apart instead of 16 KB): step 20.76 against
// 20.38 ms.  So by default the pencil form is taken when a rank's planes hold 6 GB or more; jps_slab_set_layout (or
// JPS_SLAB_FFT=pencil|cufft2d) forces one.  The pencil plans exist whenever the mesh size is even.
static bool slab_pencil_possible(int n, int nranks) { return nranks > 1 && n % 2 == 0; }

static bool slab_pencil_default(int n, int nranks) {
  static const int forced = [] {
    const char* e = getenv("JPS_SLAB_FFT");
    return !e ? 0 : !strcmp(e, "cufft2d") ? 1 : !strcmp(e, "pencil") ? 2 : 0;
  }();
  if (!slab_pencil_possible(n, nranks) || forced == 1) return false;
  if (forced == 2) return true;
  return (double)(n / nranks) * n * n * 4.0 >= 6.0e9;
}

struct SlabPkParams {
  const float2* dk;      // [n][nyl][nz]
  int n, nz, nyl, y0;
  const int32_t* lut;
  const float* wl;
  int nbc;
  double* acc;           // [nbc][4]
  const float* dc;       // device: Re rho_hat(k=0) (used when normalise)
  int normalise;
  int kz_major;          // x-fast kernel only: 0 = dk[yl][kz][x], 1 = dk[kz][yl][x]
  // x-fast kernel only: the lut in segment form, staged in shared memory (nseg == 0: read the lut from L2)
  const int32_t* seg_bp; const int32_t* seg_val; const int32_t* coarse;
  int nseg, ncoarse;
};

// k^2 -> compact bin through the shared-memory segment list: entry point from floor(sqrt(k^2)) (exact in float32
// for k^2 < 2^22), then a short forward scan (bounded by the table builder, <= 8 steps).
__device__ __forceinline__ int seg_lookup(const int* __restrict__ bp, const int* __restrict__ val,
                                          const int* __restrict__ coarse, int nseg, int k2) {
  int sidx = coarse[(int)sqrtf((float)k2)];
  while (sidx + 1 < nseg && k2 >= bp[sidx + 1]) ++sidx;
  return val[sidx];
}

// One warp per (a = |kx|, local y): folds the rows ix = a and ix = n-a, lanes along kz.
template <int MODE>
__global__ void __launch_bounds__(256) pk_bin_ysharded_kernel(SlabPkParams P) {
  extern __shared__ float sacc[];
  constexpr bool SMEM = (MODE != ACC_GLOBAL);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
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
    const double s = (double)P.n * (double)P.n * (double)P.n / (double)P.dc[0];
    scale2 = (float)(s * s);
  }
  const int n = P.n, nz = P.nz, mid = n / 2;
  const long long items = (long long)(mid + 1) * P.nyl;
  for (long long it = (long long)blockIdx.x * nwarps + warp; it < items; it += (long long)gridDim.x * nwarps) {
    const int a = (int)(it / P.nyl), yl = (int)(it % P.nyl);
    const int iy = P.y0 + yl;
    const int ky = iy > mid ? iy - n : iy;
    const bool two = (a > 0 && 2 * a != n);
    const float2* r0 = P.dk + ((size_t)a * P.nyl + yl) * nz;
    const float2* r1 = P.dk + ((size_t)(two ? n - a : a) * P.nyl + yl) * nz;
    const float wab = P.wl[a] * P.wl[iy];
    const int k2ab = a * a + ky * ky;
    constexpr int UNR = 4;                         // kz chunks in flight: 2 rows x 4 chunks = 8 loads per lane
    for (int kzb = 0; kzb < nz; kzb += 32 * UNR) {
      float2 d0[UNR], d1[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int kz = kzb + 32 * u + lane;
        d0[u] = (kz < nz) ? __ldg(r0 + kz) : make_float2(0.0f, 0.0f);
        d1[u] = (two && kz < nz) ? __ldg(r1 + kz) : make_float2(0.0f, 0.0f);
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
          const float c = wab * P.wl[kz];
          float re = d0[u].x * c, im = d0[u].y * c;
          float sum = re * re + im * im;
          re = d1[u].x * c; im = d1[u].y * c;
          sum += re * re + im * im;
          sum *= scale2;
          float mu2 = 0.0f;
          if (k2 > 0) mu2 = (float)(kz * kz) / (float)k2;
          else if (P.normalise) sum = 0.0f;
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
            float* q = my + cb * 3;
            q[0] += v[0]; q[1] += v[1]; q[2] += v[2];
          } else if (MODE == ACC_BLOCK) {
            float* q = my + cb * 3;
            atomicAdd(q + 0, v[0]); atomicAdd(q + 1, v[1]); atomicAdd(q + 2, v[2]);
          } else {
            double* q = P.acc + (size_t)cb * 4;
            atomicAdd(q + 0, (double)v[0]); atomicAdd(q + 1, (double)v[1]); atomicAdd(q + 2, (double)v[2]);
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

// Same estimator on the x-fast layout dk[yl][kz][x]: one warp per (local y, kz) row, lanes along
// a = |kx| (|k| is monotone in a, so the warp-segmented reduction applies unchanged); the partner
// row element n - a is the same row read backwards.  The window product keeps the reference's
// order (c(kx) c(ky)) c(kz).
template <int MODE, bool FOLD_Y>
__global__ void __launch_bounds__(256) pk_bin_xfast_kernel(SlabPkParams P) {
  extern __shared__ float sacc[];
  constexpr bool SMEM = (MODE != ACC_GLOBAL);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
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
    const double s = (double)P.n * (double)P.n * (double)P.n / (double)P.dc[0];
    scale2 = (float)(s * s);
  }
  const int n = P.n, nz = P.nz, mid = n / 2;
  // segment tables behind the accumulators (dynamic shared memory sized by the launcher)
  const int nacc_all = (MODE == ACC_WARP ? nwarps : (MODE == ACC_BLOCK ? 1 : 0)) * nacc;
  int* s_bp = reinterpret_cast<int*>(sacc + nacc_all);
  int* s_val = s_bp + P.nseg;
  int* s_coarse = s_val + P.nseg;
  if (P.nseg) {
    for (int i = threadIdx.x; i < P.nseg; i += blockDim.x) { s_bp[i] = P.seg_bp[i]; s_val[i] = P.seg_val[i]; }
    for (int i = threadIdx.x; i < P.ncoarse; i += blockDim.x) s_coarse[i] = P.coarse[i];
    __syncthreads();
  }
  // kz_major (the pencil plan's [kz][y][x]): the whole ky axis is local, so the rows y = b and y = n - b are folded
  // together (4 modes per lane and step share |k|, mu, window, weights and bin); a slab shard folds +-kx only.
  constexpr bool fold_y = FOLD_Y;               // compile time: a slab shard must not carry the second row's registers
  const long long items = fold_y ? (long long)nz * (mid + 1) : (long long)P.nyl * nz;
  for (long long it = (long long)blockIdx.x * nwarps + warp; it < items; it += (long long)gridDim.x * nwarps) {
    int yl, kz;
    if (fold_y) { kz = (int)(it / (mid + 1)); yl = (int)(it % (mid + 1)); }
    else { yl = (int)(it / nz); kz = (int)(it % nz); }
    const int iy = P.y0 + yl;
    const int ky = iy > mid ? iy - n : iy;
    const bool two_y = fold_y && yl > 0 && 2 * yl != n;
    const float2* r = P.dk + (fold_y ? ((size_t)kz * n + yl) * n : (size_t)it * n);
    const float2* r2 = two_y ? P.dk + ((size_t)kz * n + (n - yl)) * n : r;
    const float wy = P.wl[iy], wz = P.wl[kz];
    const int k2yz = ky * ky + kz * kz;
    const float kz2 = (float)(kz * kz);
    constexpr int UNR = 4;
    for (int ab = 0; ab <= mid; ab += 32 * UNR) {
      float2 d0[UNR], d1[UNR], d2[UNR], d3[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int a = ab + 32 * u + lane;
        const bool in = a <= mid;
        const bool two = in && a > 0 && 2 * a != n;
        d0[u] = in ? __ldg(r + a) : make_float2(0.0f, 0.0f);
        d1[u] = two ? __ldg(r + (n - a)) : make_float2(0.0f, 0.0f);
        if (FOLD_Y) {
          d2[u] = (in && two_y) ? __ldg(r2 + a) : make_float2(0.0f, 0.0f);
          d3[u] = (two && two_y) ? __ldg(r2 + (n - a)) : make_float2(0.0f, 0.0f);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int a0 = ab + 32 * u;
        if (a0 > mid) break;                       // warp-uniform
        const int a = a0 + lane;
        float v[3] = {0.0f, 0.0f, 0.0f};
        int cb = -2;
        if (a <= mid) {
          const int k2 = k2yz + a * a;
          cb = P.nseg ? seg_lookup(s_bp, s_val, s_coarse, P.nseg, k2) : __ldg(P.lut + k2);
          const float c = (P.wl[a] * wy) * wz;
          float re = d0[u].x * c, im = d0[u].y * c;
          float sum = re * re + im * im;
          re = d1[u].x * c; im = d1[u].y * c;
          sum += re * re + im * im;
          if (FOLD_Y) {
            re = d2[u].x * c; im = d2[u].y * c;
            sum += re * re + im * im;
            re = d3[u].x * c; im = d3[u].y * c;
            sum += re * re + im * im;
          }
          sum *= scale2;
          float mu2 = 0.0f;
          if (k2 > 0) mu2 = kz2 / (float)k2;
          else if (P.normalise) sum = 0.0f;
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
            float* q = my + cb * 3;
            q[0] += v[0]; q[1] += v[1]; q[2] += v[2];
          } else if (MODE == ACC_BLOCK) {
            float* q = my + cb * 3;
            atomicAdd(q + 0, v[0]); atomicAdd(q + 1, v[1]); atomicAdd(q + 2, v[2]);
          } else {
            double* q = P.acc + (size_t)cb * 4;
            atomicAdd(q + 0, (double)v[0]); atomicAdd(q + 1, (double)v[1]); atomicAdd(q + 2, (double)v[2]);
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

// Peer-store kernel for the x-fast layout: 32 (xl) x 32 (kz) tiles go through shared memory so that
// the loads are coalesced along kz and the (remote) stores along x:
//   dst_q[((yl * nz) + kz) * n + rank * nxl + xl] = yz[(xl * n + q * nyl + yl) * nz + kz]
__global__ void __launch_bounds__(256) slab_pack_p2p_xfast_kernel(const float2* __restrict__ yz,
                                                                  void* const* __restrict__ peers, int n, int nz,
                                                                  int nxl, int nyl, int nranks, int rank,
                                                                  int x_begin, int x_count) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 8 rows of 32 lanes
  const int ntx = (x_count + 31) / 32, ntz = (nz + 31) / 32;
  const long long ntiles = (long long)n * ntx * ntz;                 // over all y = q * nyl + yl
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // consecutive CTAs take consecutive y -> different yl of the same peer, then the next peer
    const int y = (int)(t % n);
    const long long rest = t / n;
    const int bz = (int)(rest % ntz), bx = (int)(rest / ntz);
    const int q = y / nyl, yl = y % nyl;
    const int xl0 = x_begin + bx * 32, kz0 = bz * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xl = xl0 + ty + 8 * i, kz = kz0 + tx;
      if (xl < x_begin + x_count && kz < nz) tile[ty + 8 * i][tx] = __ldg(yz + ((size_t)xl * n + y) * nz + kz);
    }
    __syncthreads();
    float2* dst = reinterpret_cast<float2*>(peers[q]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kz = kz0 + ty + 8 * i, xl = xl0 + tx;
      if (kz < nz && xl < x_begin + x_count)
        dst[((size_t)yl * nz + kz) * n + (size_t)rank * nxl + xl] = tile[tx][ty + 8 * i];
    }
    __syncthreads();
  }
}

// Same transposing peer-store with the TMA doing the remote writes (cp.async.bulk, shared -> global):
// a CTA stages a TX (x) x 32 (kz) tile in shared memory in OUTPUT order [kz][x] (loads coalesced along
// kz), then 32 lanes each hand ONE row -- TX * 8 bytes, contiguous along x in the peer's shard -- to the
// bulk-copy engine.  The stores are asynchronous: with two tile buffers the loads of tile i+1 run while
// the NVLink writes of tile i drain, every remote write is a 256/512-byte burst, and no thread waits on
// a store.  (The plain kernel above issues 8-byte stores per thread and two barriers per 8 KB tile.)
//   dst_q[((yl * nz) + kz) * n + rank * nxl + xl] = yz[(xl * n + q * nyl + yl) * nz + kz]
__device__ __forceinline__ void bulk_store_row(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned src = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src), "r"(bytes) : "memory");
}

template <int TX>
__global__ void __launch_bounds__(256) slab_pack_p2p_xfast_tma_kernel(const float2* __restrict__ yz,
                                                                      void* const* __restrict__ peers, int n, int nz,
                                                                      int nxl, int nyl, int nranks, int rank,
                                                                      int x_begin, int x_count) {
  constexpr int TZ = 32, PITCH = TX + 2;               // pitch: rows stay 16-byte aligned, 2-way bank conflicts at most
  __shared__ __align__(128) float2 tile[2][TZ][PITCH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ntx = x_count / TX, ntz = (nz + TZ - 1) / TZ;
  const long long ntiles = (long long)n * ntx * ntz;                 // over all y = q * nyl + yl
  int it = 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    // consecutive CTAs take consecutive y -> different yl of the same peer, then the next peer
    const int y = (int)(t % n);
    const long long rest = t / n;
    const int bz = (int)(rest % ntz), bx = (int)(rest / ntz);
    const int q = y / nyl, yl = y % nyl;
    const int xl0 = x_begin + bx * TX, kz0 = bz * TZ;
    const int b = it & 1;
    // buffer b was handed to the bulk engine two tiles ago: wait until it has been READ (one younger group may be in flight)
    if (warp == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    const int kz = kz0 + lane;
    constexpr int ROWS = TX / 8;                                     // x rows per warp
    float2 v[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const int xl = xl0 + warp + 8 * i;
      v[i] = (kz < nz) ? __ldg(yz + ((size_t)xl * n + y) * nz + kz) : make_float2(0.0f, 0.0f);
    }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) tile[b][lane][warp + 8 * i] = v[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // these writes -> visible to the bulk-copy engine
    __syncthreads();
    if (warp == 0) {
      if (kz < nz) {
        float2* dst = reinterpret_cast<float2*>(peers[q]) + ((size_t)yl * nz + kz) * n + (size_t)rank * nxl + xl0;
        bulk_store_row(dst, &tile[b][lane][0], TX * (unsigned)sizeof(float2));
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");      // every lane commits (possibly empty) groups in step
    }
  }
  if (warp == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all writes performed before the grid retires
}

// Transposing peer-store for planes transformed in pencil form, yz[xl][kz][y] (y fastest): 32 (xl) x 32 (y) tiles at
// fixed kz go through shared memory, loads coalesced along y, (remote) stores coalesced along x:
//   dst_q[((yl * nz) + kz) * n + rank * nxl + xl] = yz[(xl * nz + kz) * n + q * nyl + yl]
__global__ void __launch_bounds__(256) slab_pack_p2p_xfast_ykz_kernel(const float2* __restrict__ yz,
                                                                      void* const* __restrict__ peers, int n, int nz,
                                                                      int nxl, int nyl, int nranks, int rank,
                                                                      int x_begin, int x_count) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 8 rows of 32 lanes
  const int ntx = (x_count + 31) / 32, nty = (n + 31) / 32;
  const long long ntiles = (long long)nz * ntx * nty;
  // consecutive CTAs take y tiles a whole peer apart, so that all links are busy at once
  const int per_peer = (nty % nranks == 0) ? nty / nranks : 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    int by = (int)(t % nty);
    const long long rest = t / nty;
    const int bx = (int)(rest % ntx), kz = (int)(rest / ntx);
    if (per_peer) by = (by % nranks) * per_peer + by / nranks;
    const int xl0 = x_begin + bx * 32, y0 = by * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xl = xl0 + ty + 8 * i, y = y0 + tx;
      if (xl < x_begin + x_count && y < n) tile[ty + 8 * i][tx] = __ldg(yz + ((size_t)xl * nz + kz) * n + y);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int y = y0 + ty + 8 * i, xl = xl0 + tx;
      if (y < n && xl < x_begin + x_count) {
        const int q = y / nyl, yl = y - q * nyl;
        reinterpret_cast<float2*>(peers[q])[((size_t)yl * nz + kz) * n + (size_t)rank * nxl + xl] = tile[tx][ty + 8 * i];
      }
    }
    __syncthreads();
  }
}

// Fused pack + all-to-all over NVLink peer memory: block q of this rank's yz-transformed planes is
// written DIRECTLY into rank q's receive buffer (peer pointers obtained through CUDA IPC on the host
// side), at the slot of this rank -- no packed send buffer, no NCCL copy kernels.  One launch moves
// (P-1)/P of the shard over NVLink and 1/P locally; CTAs interleave destinations so all links are
// busy at once.
//   dst_q[(rank*nxl + xl)*nyl*nz + e] = yz[xl*n*nz + q*nyl*nz + e],  e in [0, nyl*nz)
__global__ void __launch_bounds__(256) slab_pack_p2p_kernel(const float2* __restrict__ yz,
                                                            void* const* __restrict__ peers, int n,
                                                            int nz, int nxl, int nyl, int nranks, int rank,
                                                            int x_begin, int x_count) {
  const long long run = (long long)nyl * nz;                 // contiguous complex elements per (xl, q)
  const long long nrun = (long long)x_count * nranks;
  const bool vec = (run % 2 == 0);                           // 16-byte moves when every run is 16-byte aligned
  for (long long r = blockIdx.x; r < nrun; r += gridDim.x) {
    const int q = (int)(r % nranks);                         // consecutive CTAs -> different peers
    const int xl = x_begin + (int)(r / nranks);
    const float2* src = yz + (size_t)xl * n * nz + (size_t)q * run;
    float2* dst = reinterpret_cast<float2*>(peers[q]) + ((size_t)rank * nxl + xl) * run;
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (long long e = threadIdx.x; e < run / 2; e += blockDim.x) d4[e] = s4[e];
    } else {
      for (long long e = threadIdx.x; e < run; e += blockDim.x) dst[e] = src[e];
    }
  }
}

// compact accumulators -> user-bin arrays
__global__ void slab_expand_kernel(int nb, const int32_t* __restrict__ bin_to_compact,
                                   const double* __restrict__ acc,
                                   const unsigned long long* __restrict__ cnt,
                                   double* __restrict__ sums, int64_t* __restrict__ counts) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb) return;
  const int c = bin_to_compact[j];
  sums[j * 3 + 0] = c >= 0 ? acc[(size_t)c * 4 + 0] : 0.0;
  sums[j * 3 + 1] = c >= 0 ? acc[(size_t)c * 4 + 1] : 0.0;
  sums[j * 3 + 2] = c >= 0 ? acc[(size_t)c * 4 + 2] : 0.0;
  if (counts) counts[j] = c >= 0 ? (int64_t)cnt[c] : 0;
}

__global__ void slab_finalize_kernel(int nb, const float* __restrict__ edges, const double* __restrict__ sums,
                                     const int64_t* __restrict__ counts, float kF, double vol, float shot,
                                     float* __restrict__ k3d, float* __restrict__ pk3d, float* __restrict__ nmodes) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb) return;
  const double nm = (double)(float)counts[j];
  pk3d[j * 3 + 0] = (float)(sums[j * 3 + 0] / nm * vol - (double)shot);
  pk3d[j * 3 + 1] = (float)(sums[j * 3 + 1] / nm * 5.0 * vol);
  pk3d[j * 3 + 2] = (float)(sums[j * 3 + 2] / nm * 9.0 * vol);
  nmodes[j] = (float)counts[j];
  k3d[j] = (0.5f * (edges[j + 1] + edges[j])) * kF;
}

// Launch of the y-sharded / x-fast binning kernels on one spectrum shard (accumulators zeroed first).
static int bin_layout(jps_plan* tp, const float2* dk, int nyl, int y0, int xfast, int kz_major, const BinTable& T,
                      const float* dc, int normalise, int mas_order, cudaStream_t s) {
  {
    ScopedLaunch L(K_MEMSET, s);
    JPS_CHECK_CUDA(cudaMemsetAsync(tp->acc, 0, (size_t)std::max(T.nbc, 1) * 4 * 8, s));
  }
  if (T.nbc == 0) return JPS_OK;
  SlabPkParams P;
  P.dk = dk; P.n = tp->n; P.nz = tp->nz; P.nyl = nyl; P.y0 = y0;
  P.lut = T.lut; P.wl = tp->wlut + (size_t)(mas_order - 2) * tp->n; P.nbc = T.nbc; P.acc = tp->acc;
  P.dc = dc; P.normalise = normalise; P.kz_major = kz_major;
  // JPS_BIN_LUT=segments stages the k^2 -> bin table in shared memory in segment form.  Measured on the 2048^3 shards
  // of 2 B200s: 7.6 ms against 6.1 ms with the plain L2-resident lut -- the sqrt + scan costs more instructions than the
  // sector traffic it saves (the kernel is bound by instructions per byte), so the lut stays the default.
  static const bool no_seg = [] { const char* e = getenv("JPS_BIN_LUT"); return !(e && !strcmp(e, "segments")); }();
  const bool seg = xfast && T.nseg > 0 && !no_seg;
  P.seg_bp = T.seg_bp; P.seg_val = T.seg_val; P.coarse = T.coarse;
  P.nseg = seg ? T.nseg : 0; P.ncoarse = seg ? T.ncoarse : 0;
  const size_t seg_bytes = seg ? (size_t)(2 * T.nseg + T.ncoarse) * sizeof(int) : 0;
  // (Measured and dropped: warp-private accumulators for ~1000 bins with 4 warps per CTA -- 24.6 ms against 13.4 ms
  // on a 2048^3 spectrum.  The kernel is bound by instructions per byte (segment lookup, segmented reduction and
  // Legendre weights serve only the modes folded before them), not by its shared atomics: fold more instead.)
  const int warps = 8;
  const bool warp_private = T.nbc <= kMaxSmemBins;
  const int threads = warps * 32;
  const long long items = (xfast && kz_major) ? (long long)tp->nz * (tp->n / 2 + 1)
                          : xfast ? (long long)nyl * tp->nz : (long long)(tp->n / 2 + 1) * nyl;
  const long long want = (items + warps - 1) / warps;
  static PerDeviceFlag attr_set;
  if (!attr_set.get()) {
    const int sw = (int)((size_t)8 * kMaxSmemBins * 3 * sizeof(float));
    const int sb = (int)((size_t)kMaxBlockBins * 3 * sizeof(float));
    const int sseg = (int)((size_t)(2 * kMaxSegments + 8192) * sizeof(int));       // segment tables (x-fast kernels)
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_ysharded_kernel<ACC_WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, sw));
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_ysharded_kernel<ACC_BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_xfast_kernel<ACC_WARP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sw + sseg));
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_xfast_kernel<ACC_BLOCK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb + sseg));
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_xfast_kernel<ACC_GLOBAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sseg));
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_xfast_kernel<ACC_WARP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sw + sseg));
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_xfast_kernel<ACC_BLOCK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb + sseg));
    JPS_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_xfast_kernel<ACC_GLOBAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sseg));
    attr_set.set();
  }
  using KernelFn = void (*)(SlabPkParams);
  KernelFn fn;
  size_t smem = 0;
  long long cap_per_sm = 0;
  if (warp_private) {
    fn = xfast ? (kz_major ? pk_bin_xfast_kernel<ACC_WARP, true> : pk_bin_xfast_kernel<ACC_WARP, false>) : pk_bin_ysharded_kernel<ACC_WARP>;
    smem = (size_t)warps * T.nbc * 3 * sizeof(float) + seg_bytes;
  } else if (T.nbc <= kMaxBlockBins) {
    fn = xfast ? (kz_major ? pk_bin_xfast_kernel<ACC_BLOCK, true> : pk_bin_xfast_kernel<ACC_BLOCK, false>) : pk_bin_ysharded_kernel<ACC_BLOCK>;
    smem = (size_t)T.nbc * 3 * sizeof(float) + seg_bytes;
  } else {
    fn = xfast ? (kz_major ? pk_bin_xfast_kernel<ACC_GLOBAL, true> : pk_bin_xfast_kernel<ACC_GLOBAL, false>) : pk_bin_ysharded_kernel<ACC_GLOBAL>;
    smem = seg_bytes;
    cap_per_sm = 8;
  }
  if (!cap_per_sm) {
    int per_sm = 1;
    JPS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    cap_per_sm = std::max(per_sm, 1);
  }
  const int blocks = (int)std::min<long long>(want, (long long)kNumSMs * cap_per_sm);
  {
    ScopedLaunch L(K_PK_FOLD_BIN, s);
    fn<<<blocks, threads, smem, s>>>(P);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

int bin_xfast_layout(jps_plan* tables, const float2* dk, int nyl, int y0, int kz_major, const BinTable& T,
                     const float* dc, int normalise, int mas_order, cudaStream_t s) {
  return bin_layout(tables, dk, nyl, y0, 1, kz_major, T, dc, normalise, mas_order, s);
}

static size_t tables_bytes(int n) {
  size_t b = 0;
  return jps_plan_workspace_bytes(n, 0, JPS_PLAN_TABLES_ONLY, &b) == JPS_OK ? b : 0;
}

}  // namespace jps

using namespace jps;

extern "C" int jps_slab_plan_workspace_bytes(int n_mesh, int nranks, size_t* bytes) {
  JPS_REQUIRE(bytes != nullptr, "jps_slab_plan_workspace_bytes: bytes is NULL");
  JPS_REQUIRE(n_mesh >= 2 && n_mesh <= 4096, "jps_slab_plan_workspace_bytes: n_mesh out of range");
  JPS_REQUIRE(nranks >= 1 && n_mesh % nranks == 0, "jps_slab_plan_workspace_bytes: n_mesh=%d must be divisible by nranks=%d", n_mesh, nranks);
  cufftHandle h;
  size_t w1 = 0, w2 = 0;
  int rc = make_fft_yz(n_mesh, n_mesh / nranks, &h, &w1);
  cufftDestroy(h);
  if (rc) return rc;
  rc = make_fft_x(n_mesh, n_mesh / nranks, &h, &w2);
  cufftDestroy(h);
  if (rc) return rc;
  size_t w3 = 0, w4 = 0;
  if (chunk_planes_for(n_mesh / nranks)) {
    rc = make_fft_yz(n_mesh, chunk_planes_for(n_mesh / nranks), &h, &w3);
    cufftDestroy(h);
    if (rc) return rc;
  }
  rc = make_fft_x_contig(n_mesh, n_mesh / nranks, &h, &w4);
  cufftDestroy(h);
  if (rc) return rc;
  const size_t tb = tables_bytes(n_mesh);
  JPS_REQUIRE(tb > 0, "jps_slab_plan_workspace_bytes: table sizing failed");
  size_t wp = 0, pencil_bytes = 0;
  if (slab_pencil_possible(n_mesh, nranks)) {
    const int nxl = n_mesh / nranks, nzz = n_mesh / 2 + 1;
    size_t a = 0, b = 0;
    rc = make_c2c_contig(n_mesh / 2, (long long)nxl * n_mesh, &h, &a);
    cufftDestroy(h);
    if (rc) return rc;
    rc = make_c2c_contig(n_mesh, (long long)nxl * nzz, &h, &b);
    cufftDestroy(h);
    if (rc) return rc;
    wp = std::max(a, b);
    pencil_bytes = align_up((size_t)nxl * n_mesh * (n_mesh / 2) * sizeof(float2), 256) + align_up((size_t)(n_mesh / 4 + 2) * sizeof(float2), 256);
  }
  *bytes = align_up(std::max(std::max(std::max(w1, w2), std::max(w3, w4)), wp), 256) + align_up(tb, 256) + 1024 + 512 + pencil_bytes;
  return JPS_OK;
}

extern "C" int jps_slab_plan_destroy(jps_slab_plan_t* p) {
  if (!p) return JPS_OK;
  if (p->yz_ok) cufftDestroy(p->fft_yz);
  if (p->x_ok) cufftDestroy(p->fft_x);
  if (p->yzc_ok) cufftDestroy(p->fft_yz_chunk);
  if (p->xc_ok) cufftDestroy(p->fft_x_contig);
  if (p->pz_ok) cufftDestroy(p->pz);
  if (p->py_ok) cufftDestroy(p->py);
  if (p->pzc_ok) cufftDestroy(p->pz_chunk);
  if (p->pyc_ok) cufftDestroy(p->py_chunk);
  if (p->tables) jps_plan_destroy(p->tables);
  delete p;
  return JPS_OK;
}

extern "C" int jps_slab_plan_create(int n_mesh, int nranks, int rank, void* workspace,
                                    size_t workspace_bytes, jps_slab_plan_t** out) {
  JPS_REQUIRE(out != nullptr, "jps_slab_plan_create: plan is NULL");
  *out = nullptr;
  JPS_REQUIRE(n_mesh >= 2 && n_mesh <= 4096, "jps_slab_plan_create: n_mesh out of range");
  JPS_REQUIRE(nranks >= 1 && n_mesh % nranks == 0, "jps_slab_plan_create: n_mesh=%d must be divisible by nranks=%d", n_mesh, nranks);
  JPS_REQUIRE(rank >= 0 && rank < nranks, "jps_slab_plan_create: bad rank");
  JPS_REQUIRE(workspace && ((uintptr_t)workspace & 255) == 0, "jps_slab_plan_create: workspace must be 256-byte aligned");
  jps_slab_plan* p = new jps_slab_plan();
  p->n = n_mesh; p->nz = n_mesh / 2 + 1; p->nranks = nranks; p->rank = rank;
  p->nxl = n_mesh / nranks; p->nyl = n_mesh / nranks;
  size_t w1 = 0, w2 = 0;
  int rc = make_fft_yz(n_mesh, p->nxl, &p->fft_yz, &w1);
  if (rc) { delete p; return rc; }
  p->yz_ok = true;
  rc = make_fft_x(n_mesh, p->nyl, &p->fft_x, &w2);
  if (rc) { jps_slab_plan_destroy(p); return rc; }
  p->x_ok = true;
  size_t w3 = 0;
  p->chunk_planes = chunk_planes_for(p->nxl);
  if (p->chunk_planes) {
    rc = make_fft_yz(n_mesh, p->chunk_planes, &p->fft_yz_chunk, &w3);
    if (rc) { jps_slab_plan_destroy(p); return rc; }
    p->yzc_ok = true;
  }
  size_t w4 = 0;
  rc = make_fft_x_contig(n_mesh, p->nyl, &p->fft_x_contig, &w4);
  if (rc) { jps_slab_plan_destroy(p); return rc; }
  p->xc_ok = true;
  size_t wp = 0, pencil_bytes = 0;
  const bool want_pencil = slab_pencil_possible(n_mesh, nranks);
  if (want_pencil) {
    size_t a = 0;
    rc = make_c2c_contig(n_mesh / 2, (long long)p->nxl * n_mesh, &p->pz, &a);
    if (rc) { jps_slab_plan_destroy(p); return rc; }
    p->pz_ok = true; wp = std::max(wp, a);
    rc = make_c2c_contig(n_mesh, (long long)p->nxl * p->nz, &p->py, &a);
    if (rc) { jps_slab_plan_destroy(p); return rc; }
    p->py_ok = true; wp = std::max(wp, a);
    if (p->chunk_planes) {
      rc = make_c2c_contig(n_mesh / 2, (long long)p->chunk_planes * n_mesh, &p->pz_chunk, &a);
      if (rc) { jps_slab_plan_destroy(p); return rc; }
      p->pzc_ok = true; wp = std::max(wp, a);
      rc = make_c2c_contig(n_mesh, (long long)p->chunk_planes * p->nz, &p->py_chunk, &a);
      if (rc) { jps_slab_plan_destroy(p); return rc; }
      p->pyc_ok = true; wp = std::max(wp, a);
    }
    pencil_bytes = align_up((size_t)p->nxl * n_mesh * (n_mesh / 2) * sizeof(float2), 256) + align_up((size_t)(n_mesh / 4 + 2) * sizeof(float2), 256);
  }
  const size_t wb = align_up(std::max(std::max(std::max(w1, w2), std::max(w3, w4)), wp), 256) + 1024;       // + the peer pointer table
  const size_t tb = tables_bytes(n_mesh);
  if (workspace_bytes < wb + align_up(tb, 256) + pencil_bytes) {
    set_error("jps_slab_plan_create: workspace has %zu bytes, %zu needed", workspace_bytes, wb + align_up(tb, 256) + pencil_bytes);
    jps_slab_plan_destroy(p);
    return JPS_ERR_WORKSPACE;
  }
  p->work = workspace; p->work_bytes = wb - 1024;
  p->peer_dev = (void**)((char*)workspace + wb - 1024);
  if (cufftSetWorkArea(p->fft_yz, p->work) != CUFFT_SUCCESS || cufftSetWorkArea(p->fft_x, p->work) != CUFFT_SUCCESS ||
      (p->yzc_ok && cufftSetWorkArea(p->fft_yz_chunk, p->work) != CUFFT_SUCCESS) ||
      cufftSetWorkArea(p->fft_x_contig, p->work) != CUFFT_SUCCESS) {
    set_error("jps_slab_plan_create: cufftSetWorkArea failed");
    jps_slab_plan_destroy(p);
    return JPS_ERR_CUFFT;
  }
  rc = jps_plan_create(n_mesh, 0, JPS_PLAN_TABLES_ONLY, (char*)workspace + wb, align_up(tb, 256), &p->tables);
  if (rc) { jps_slab_plan_destroy(p); return rc; }
  if (want_pencil) {
    char* base = (char*)workspace + wb + align_up(tb, 256);
    p->zbuf = (float2*)base;
    p->ztw = (float2*)(base + align_up((size_t)p->nxl * n_mesh * (n_mesh / 2) * sizeof(float2), 256));
    std::vector<float2> tw;
    host_r2c_twiddles(n_mesh, tw);
    if (cudaMemcpy(p->ztw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cufftSetWorkArea(p->pz, p->work) != CUFFT_SUCCESS || cufftSetWorkArea(p->py, p->work) != CUFFT_SUCCESS ||
        (p->pzc_ok && cufftSetWorkArea(p->pz_chunk, p->work) != CUFFT_SUCCESS) ||
        (p->pyc_ok && cufftSetWorkArea(p->py_chunk, p->work) != CUFFT_SUCCESS)) {
      set_error("jps_slab_plan_create: pencil transform setup failed");
      jps_slab_plan_destroy(p);
      return JPS_ERR_CUFFT;
    }
    p->pencil_ok = true;
    p->use_pencil = slab_pencil_default(n_mesh, nranks);
  }
  *out = p;
  return JPS_OK;
}

extern "C" int jps_slab_fft_yz(jps_slab_plan_t* p, const float* slab, void* out, void* stream) {
  JPS_REQUIRE(p && slab && out, "jps_slab_fft_yz: NULL argument");
  return jps_slab_fft_yz_planes(p, slab, out, 0, p->nxl, stream);
}

extern "C" int jps_slab_pack(jps_slab_plan_t* p, const void* in, void* out, void* stream) {
  JPS_REQUIRE(p && in && out, "jps_slab_pack: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  // out[q][xl][yl][kz] = in[xl][q*nyl + yl][kz]: for each q a 2-D copy (rows = xl) of nyl*nz complex
  const size_t row = (size_t)p->nyl * p->nz * sizeof(float2);
  const size_t spitch = (size_t)p->n * p->nz * sizeof(float2);
  for (int q = 0; q < p->nranks; ++q) {
    ScopedLaunch L(K_MISC, s);
    JPS_CHECK_CUDA(cudaMemcpy2DAsync((char*)out + (size_t)q * p->nxl * row, row,
                                     (const char*)in + (size_t)q * row, spitch, row, (size_t)p->nxl,
                                     cudaMemcpyDeviceToDevice, s));
  }
  return JPS_OK;
}

// Open a CUDA IPC memory handle exported by another process ON THE CURRENT DEVICE, so that kernels
// of this device can dereference the peer's memory over NVLink (peer access is enabled lazily by
// the driver).  `handle` is the 64-byte cudaIpcMemHandle_t.
extern "C" int jps_ipc_open(const void* handle, void** out_ptr) {
  JPS_REQUIRE(handle && out_ptr, "jps_ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(&h, handle, sizeof(h));
  JPS_CHECK_CUDA(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return JPS_OK;
}

extern "C" int jps_ipc_close(void* ptr) {
  if (ptr) JPS_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return JPS_OK;
}

extern "C" int jps_enable_peer_access(int peer_device) {
  int dev = 0;
  JPS_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev == peer_device) return JPS_OK;
  int can = 0;
  JPS_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) {
    set_error("jps_enable_peer_access: device %d cannot access device %d", dev, peer_device);
    return JPS_ERR_UNSUPPORTED;
  }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return JPS_OK; }
  JPS_CHECK_CUDA(e);
  return JPS_OK;
}

extern "C" int jps_slab_pack_p2p_planes(jps_slab_plan_t* p, const void* yz, void* const* peer_recv, int x_begin,
                                        int x_count, void* stream) {
  JPS_REQUIRE(p && yz && peer_recv, "jps_slab_pack_p2p: NULL argument");
  JPS_REQUIRE(p->nranks <= kMaxRanks, "jps_slab_pack_p2p: too many ranks");
  JPS_REQUIRE(x_begin >= 0 && x_count >= 1 && x_begin + x_count <= p->nxl, "jps_slab_pack_p2p: plane range [%d, %d) outside [0, %d)",
              x_begin, x_begin + x_count, p->nxl);
  cudaStream_t s = (cudaStream_t)stream;
  bool changed = false;
  for (int q = 0; q < p->nranks; ++q) {
    JPS_REQUIRE(peer_recv[q] != nullptr && ((uintptr_t)peer_recv[q] & 15) == 0, "jps_slab_pack_p2p: peer pointer %d is NULL or misaligned", q);
    if (p->peer_host[q] != peer_recv[q]) { p->peer_host[q] = peer_recv[q]; changed = true; }
  }
  if (changed)
    JPS_CHECK_CUDA(cudaMemcpyAsync(p->peer_dev, p->peer_host, (size_t)p->nranks * sizeof(void*), cudaMemcpyHostToDevice, s));
  JPS_REQUIRE(((uintptr_t)yz & 15) == 0, "jps_slab_pack_p2p: yz must be 16-byte aligned");
  {
    ScopedLaunch L(K_MISC, s);
    const long long nrun = (long long)x_count * p->nranks;
    // A partial range runs NEXT TO the transform of the following piece (on a high-priority stream): it is NVLink
    // bound, so it must leave most of every SM to cuFFT.  Round 1 launched up to 8 CTAs of 256 threads per SM for the
    // transposing kernels -- every thread slot of the GPU -- and the "overlap" hid nothing (measured on 8 GPUs: fused
    // stage 9.45 ms = transform 4.11 + transfer 5.46).  Now 4 CTAs per SM for a partial range (JPS_PACK_CTAS_PER_SM;
    // fused stage on 2 GPUs, 2048^3: 1 / 2 / 3 / 4 / 6 per SM -> 37.2 / 28.8 / 26.9 / 25.8 / 26.0 ms), 8 when the whole
    // slab is sent in one launch (nothing to share the SMs with).
    static const int per_sm_env = [] { const char* e = getenv("JPS_PACK_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
    const long long per_sm = (x_count == p->nxl) ? 8 : (per_sm_env > 0 ? per_sm_env : 4);
    const long long cap = (long long)kNumSMs * per_sm;
    // JPS_PACK_KERNEL=tma selects the TMA bulk-store variant.  Measured on 2 B200s (2048^3, 8.6 GB leaving each
    // rank, kernel alone): straight contiguous peer copy 12.4 ms = 692 GB/s; the plain load/store transposing kernel
    // 12.6 ms = 0.99 of it; the TMA variant 14.0 ms = 0.89 (one warp serialises 32 UBLKCP issues per 16 KB tile while
    // the other seven wait at the barrier).  The link is already saturated by the plain kernel, so it stays the default.
    static const bool no_tma = [] { const char* e = getenv("JPS_PACK_KERNEL"); return !(e && !strcmp(e, "tma")); }();
    static const int tma_ctas = [] { const char* e = getenv("JPS_PACK_TMA_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
    if (p->xfast && p->use_pencil) {
      const long long ntiles = (long long)p->nz * ((x_count + 31) / 32) * ((p->n + 31) / 32);
      slab_pack_p2p_xfast_ykz_kernel<<<(int)std::min<long long>(ntiles, cap), 256, 0, s>>>(
          (const float2*)yz, p->peer_dev, p->n, p->nz, p->nxl, p->nyl, p->nranks, p->rank, x_begin, x_count);
    } else if (p->xfast && !no_tma && x_count % 32 == 0 && x_begin % 2 == 0 && p->nxl % 2 == 0 && p->n % 2 == 0) {
      const int tx = (x_count % 64 == 0) ? 64 : 32;
      const long long ntiles = (long long)p->n * (x_count / tx) * ((p->nz + 31) / 32);
      const int blocks = (int)std::min<long long>(ntiles, (long long)kNumSMs * (tma_ctas > 0 ? tma_ctas : 2));
      if (tx == 64)
        slab_pack_p2p_xfast_tma_kernel<64><<<blocks, 256, 0, s>>>((const float2*)yz, p->peer_dev, p->n, p->nz, p->nxl, p->nyl,
                                                                 p->nranks, p->rank, x_begin, x_count);
      else
        slab_pack_p2p_xfast_tma_kernel<32><<<blocks, 256, 0, s>>>((const float2*)yz, p->peer_dev, p->n, p->nz, p->nxl, p->nyl,
                                                                 p->nranks, p->rank, x_begin, x_count);
    } else if (p->xfast) {
      const long long ntiles = (long long)p->n * ((x_count + 31) / 32) * ((p->nz + 31) / 32);
      slab_pack_p2p_xfast_kernel<<<(int)std::min<long long>(ntiles, cap), 256, 0, s>>>(
          (const float2*)yz, p->peer_dev, p->n, p->nz, p->nxl, p->nyl, p->nranks, p->rank, x_begin, x_count);
    } else {
      slab_pack_p2p_kernel<<<(int)std::min<long long>(nrun, cap), 256, 0, s>>>(
          (const float2*)yz, p->peer_dev, p->n, p->nz, p->nxl, p->nyl, p->nranks, p->rank, x_begin, x_count);
    }
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

extern "C" int jps_slab_pack_p2p(jps_slab_plan_t* p, const void* yz, void* const* peer_recv, void* stream) {
  JPS_REQUIRE(p != nullptr, "jps_slab_pack_p2p: NULL argument");
  return jps_slab_pack_p2p_planes(p, yz, peer_recv, 0, p->nxl, stream);
}

extern "C" int jps_slab_chunk_planes(jps_slab_plan_t* p) { return p ? p->chunk_planes : 0; }

extern "C" int jps_slab_set_layout(jps_slab_plan_t* p, int layout) {
  JPS_REQUIRE(p != nullptr, "jps_slab_set_layout: NULL plan");
  const int xfast = layout & JPS_SLAB_LAYOUT_XFAST;
  JPS_REQUIRE(!xfast || p->nranks > 1, "jps_slab_set_layout: the x-fast layout is produced by the peer-store kernel (nranks > 1)");
  JPS_REQUIRE(!((layout & JPS_SLAB_FFT_PENCIL) && (layout & JPS_SLAB_FFT_CUFFT2D)), "jps_slab_set_layout: contradictory transform flags");
  JPS_REQUIRE(!(layout & JPS_SLAB_FFT_PENCIL) || (xfast && p->pencil_ok), "jps_slab_set_layout: the pencil form needs the x-fast layout and an even mesh size");
  p->xfast = xfast ? 1 : 0;
  if (layout & JPS_SLAB_FFT_PENCIL) p->use_pencil = true;
  else if (layout & JPS_SLAB_FFT_CUFFT2D) p->use_pencil = false;
  else p->use_pencil = p->pencil_ok && slab_pencil_default(p->n, p->nranks);
  return JPS_OK;
}

// 2-D transform of the owned planes [x_begin, x_begin + x_count): x_count is the whole slab or
// jps_slab_chunk_planes() (the two batch sizes a cuFFT plan exists for).  `slab` / `yz` are the
// BASE pointers of the owned planes / of the output.
extern "C" int jps_slab_fft_yz_planes(jps_slab_plan_t* p, const float* slab, void* out, int x_begin, int x_count,
                                      void* stream) {
  JPS_REQUIRE(p && slab && out, "jps_slab_fft_yz_planes: NULL argument");
  JPS_REQUIRE(x_begin >= 0 && x_begin + x_count <= p->nxl, "jps_slab_fft_yz_planes: plane range outside the slab");
  cudaStream_t s = (cudaStream_t)stream;
  const bool whole = (x_count == p->nxl), chunk = (p->chunk_planes && x_count == p->chunk_planes);
  if (!whole && !chunk) {
    set_error("jps_slab_fft_yz_planes: x_count=%d is neither the slab (%d) nor the chunk size (%d)", x_count, p->nxl, p->chunk_planes);
    return JPS_ERR_INVALID;
  }
  const size_t n = (size_t)p->n;
  if (p->use_pencil && p->xfast) {
    // out[xl][kz][y]: C2C of length n/2 along z (real lines read as complex pairs), untangle + transpose, C2C along y
    const cufftHandle hz = whole ? p->pz : p->pz_chunk, hy = whole ? p->py : p->py_chunk;
    const size_t M = n / 2;
    float2* zb = p->zbuf + (size_t)x_begin * n * M;
    float2* ob = (float2*)out + (size_t)x_begin * n * p->nz;
    JPS_REQUIRE((((uintptr_t)slab) & 7) == 0, "jps_slab_fft_yz_planes: the planes must be 8-byte aligned");
    JPS_CHECK_CUFFT(cufftSetStream(hz, s));
    JPS_CHECK_CUFFT(cufftSetStream(hy, s));
    {
      ScopedLaunch L(K_FFT_R2C, s);
      JPS_CHECK_CUFFT(cufftExecC2C(hz, (cufftComplex*)const_cast<float*>(slab + (size_t)x_begin * n * n), (cufftComplex*)zb, CUFFT_FORWARD));
    }
    int rc = launch_r2c_untangle_transpose(zb, ob, p->ztw, p->n, x_count, s);
    if (rc) return rc;
    ScopedLaunch L(K_FFT_C2C_Y, s);
    JPS_CHECK_CUFFT(cufftExecC2C(hy, (cufftComplex*)ob, (cufftComplex*)ob, CUFFT_FORWARD));
    return JPS_OK;
  }
  const cufftHandle h = whole ? p->fft_yz : p->fft_yz_chunk;
  JPS_REQUIRE(whole || p->yzc_ok, "jps_slab_fft_yz_planes: no chunk plan");
  JPS_CHECK_CUFFT(cufftSetStream(h, s));
  ScopedLaunch L(K_FFT_R2C, s);
  JPS_CHECK_CUFFT(cufftExecR2C(h, (cufftReal*)(slab + (size_t)x_begin * n * n),
                               (cufftComplex*)((float2*)out + (size_t)x_begin * n * p->nz)));
  return JPS_OK;
}

extern "C" int jps_slab_fft_x(jps_slab_plan_t* p, void* data, void* stream) {
  JPS_REQUIRE(p && data, "jps_slab_fft_x: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  const cufftHandle h = p->xfast ? p->fft_x_contig : p->fft_x;
  JPS_CHECK_CUFFT(cufftSetStream(h, s));
  ScopedLaunch L(K_FFT_R2C, s);
  JPS_CHECK_CUFFT(cufftExecC2C(h, (cufftComplex*)data, (cufftComplex*)data, CUFFT_FORWARD));
  return JPS_OK;
}

extern "C" int jps_slab_powspec_partial(jps_slab_plan_t* p, const void* dk, const float* dc,
                                        int normalise, float box_size, const float* k_edges, int nb,
                                        int mas_order, double* sums, int64_t* counts, void* stream) {
  JPS_REQUIRE(p && dk && k_edges && sums, "jps_slab_powspec_partial: NULL argument");
  JPS_REQUIRE(!normalise || dc, "jps_slab_powspec_partial: normalise needs the DC mode (dc)");
  JPS_REQUIRE(nb >= 1 && nb <= kMaxUserBins, "jps_slab_powspec_partial: nb out of range");
  JPS_REQUIRE(mas_order >= 2 && mas_order <= 4 && box_size > 0.0f, "jps_slab_powspec_partial: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  jps_plan* tp = p->tables;
  const float kF = ref_kF(box_size);
  std::vector<float> kg((size_t)nb + 1);
  for (int i = 0; i <= nb; ++i) kg[(size_t)i] = k_edges[i] / kF;
  BinTable* T = nullptr;
  int rc = ensure_bin_table(tp, kg.data(), nb, TABLE_PK_EDGES, s, &T);
  if (rc) return rc;
  rc = bin_layout(tp, (const float2*)dk, p->nyl, p->rank * p->nyl, p->xfast ? 1 : 0, 0, *T, dc, normalise, mas_order, s);
  if (rc) return rc;
  {
    ScopedLaunch L(K_PK_FINALIZE, s);
    slab_expand_kernel<<<(nb + 127) / 128, 128, 0, s>>>(nb, T->bin_to_compact, tp->acc, T->cnt, sums, counts);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

extern "C" int jps_slab_powspec_finalize(jps_slab_plan_t* p, float box_size, const float* k_edges,
                                         int nb, const double* sums, const int64_t* counts,
                                         float shot_noise, float* k3d, float* pk3d, float* nmodes,
                                         void* stream) {
  JPS_REQUIRE(p && k_edges && sums && counts && k3d && pk3d && nmodes, "jps_slab_powspec_finalize: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  const float kF = ref_kF(box_size);
  std::vector<float> kg((size_t)nb + 1);
  for (int i = 0; i <= nb; ++i) kg[(size_t)i] = k_edges[i] / kF;
  BinTable* T = nullptr;
  int rc = ensure_bin_table(p->tables, kg.data(), nb, TABLE_PK_EDGES, s, &T);   // cached: gives the device edges
  if (rc) return rc;
  {
    ScopedLaunch L(K_PK_FINALIZE, s);
    slab_finalize_kernel<<<(nb + 127) / 128, 128, 0, s>>>(nb, T->edges, sums, counts, kF,
                                                         ref_volume(box_size, p->n), shot_noise, k3d, pk3d, nmodes);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}
