// Shared declarations for libjps.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/jps.h"

namespace jps {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);

#define JPS_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      jps::set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)e__,      \
                     cudaGetErrorString(e__), #expr);                                     \
      return JPS_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define JPS_CHECK_CUFFT(expr)                                                             \
  do {                                                                                    \
    cufftResult r__ = (expr);                                                             \
    if (r__ != CUFFT_SUCCESS) {                                                           \
      jps::set_error("%s:%d cuFFT error %d in %s", __FILE__, __LINE__, (int)r__, #expr);  \
      return JPS_ERR_CUFFT;                                                               \
    }                                                                                     \
  } while (0)

#define JPS_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      jps::set_error(__VA_ARGS__);                                                        \
      return JPS_ERR_INVALID;                                                             \
    }                                                                                     \
  } while (0)

#define JPS_CHECK_LAUNCH() JPS_CHECK_CUDA(cudaGetLastError())

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

// cudaFuncSetAttribute applies to the CURRENT device only.  The product runs one process per GPU, but
// nothing in the C ABI forbids one process driving several: keep one "already set" flag per device.
struct PerDeviceFlag {
  bool done[64] = {};
  static int dev() {
    int d = 0;
    return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) ? d : 0;
  }
  bool get() const { return done[dev()]; }
  void set() { done[dev()] = true; }
};

// ---------------------------------------------------------------- launch accounting
// Every kernel launch of the library goes through a ScopedLaunch: it counts launches (always)
// and, when profiling is enabled with jps_profile_enable(1), brackets the launch with CUDA
// events on the launching stream so that bench.py can report per-kernel device time.
enum KernelId {
  K_PAINT_ATOMIC = 0,
  K_BUCKET_COUNT,
  K_BUCKET_SCAN,
  K_BUCKET_SCATTER,
  K_BUCKET_FINE,
  K_PAINT_TILE,
  K_PK_FOLD_BIN,
  K_PK_COUNT,
  K_PK_FINALIZE,
  K_FFT_R2C,        // cuFFT (library) -- timed, not counted as one of ours
  K_FFT_C2R,
  K_MEMSET,
  K_SHELL_FILTER,
  K_TRIPLE_REDUCE,
  K_XI_BIN,
  K_MISC,
  K_TEXT_INDEX,
  K_TEXT_PARSE,
  K_TEXT_COMPACT,
  K_MOCK_FIELD,
  K_MOCK_POPULATE,
  K_INTERLACE,
  K_TRANSPOSE,
  K_FFT_C2C_Y,      // pencil plan: contiguous 1-D passes along y and x (cuFFT, library)
  K_FFT_C2C_X,
  K_NUM
};

struct ScopedLaunch {
  ScopedLaunch(int id, cudaStream_t s);
  ~ScopedLaunch();
  int id;
  cudaStream_t s;
  cudaEvent_t e0, e1;
  bool timed;
};

// ---------------------------------------------------------------- plan
// Maximum number of distinct (reachable) k-bins the warp-private shared-memory
// accumulators can hold; above it the binning kernel accumulates with global atomics.
constexpr int kMaxSmemBins = 576;
// Up to this many bins one shared accumulator set per CTA is used (shared atomics by segment
// heads); above it the kernels fall back to global float64 reds.
constexpr int kMaxBlockBins = 8192;
constexpr int kIsumSlots = 32768;   // 256 KB; an all-triangles run at 256^3 (276 pairs x 20 angles) uses ~12k
enum AccMode { ACC_WARP = 0, ACC_BLOCK = 1, ACC_GLOBAL = 2 };
constexpr int kMaxUserBins = 1 << 18;

// bin-table kinds
enum TableMode {
  TABLE_PK_EDGES = 0,        // powspec_vec: user edges, half-space mode counts
  TABLE_PK_FUNDAMENTAL = 1,  // powspec_vec_fundamental: bin = int(|k|)
  TABLE_XI_EDGES = 2,        // xi_vec: user edges, full-grid pair counts
  TABLE_XI_FUNDAMENTAL = 3,  // xi_vec_fundamental
  TABLE_PK_EDGES_HERM = 4    // user edges, stored modes with 0 < kz < N/2 counted twice (Q7 corrected)
};

constexpr int kNumTables = 4;
constexpr int kMaxSegments = 4096;    // 32 KB of shared memory for (seg_bp, seg_val)

struct BinTable {               // cached k^2 -> bin lookup for one set of edges (one slot)
  std::vector<float> key;       // mode tag + edges in grid units (float32) that produced it
  int mode = 0;
  int nb = 0;                   // user bins
  int nbc = 0;                  // compact (reachable) bins
  bool valid = false;
  unsigned long long stamp = 0; // LRU
  // device storage (inside the plan workspace)
  int32_t* lut = nullptr;              // [k2max+1]  k^2 -> compact bin (or -1)
  int32_t* compact_to_bin = nullptr;   // [acc_cap]
  int32_t* bin_to_compact = nullptr;   // [kMaxUserBins] user bin -> compact bin or -1
  float* edges = nullptr;              // [nb+1] bin edges in grid units
  unsigned long long* cnt = nullptr;   // [acc_cap] exact mode / cell counts   } geometry only,
  double* ksum = nullptr;              // [acc_cap] sum of |k| (grid units)     } computed once
  unsigned long long* lastidx = nullptr;  // [acc_cap] largest C-order flat index (Q18) } per table
  // The lut as a short list of SEGMENTS of constant value (k^2 in [seg_bp[i], seg_bp[i+1]) -> seg_val[i]) with a
  // per-integer-|k| entry point coarse[floor(sqrt(k^2))]: small enough for shared memory (kernels whose lanes
  // run along one axis of a 2048^3 spectrum otherwise pull one 32-byte L2 sector of the 12 MB lut per mode).
  int32_t* seg_bp = nullptr;           // [kMaxSegments]
  int32_t* seg_val = nullptr;          // [kMaxSegments]
  int32_t* coarse = nullptr;           // [isqrt(k2max) + 2]
  int nseg = 0;                        // 0: too many segments / too long scans -> use the lut
  int ncoarse = 0;
};

}  // namespace jps

struct jps_plan {
  int n = 0;                    // mesh cells per side
  int nz = 0;                   // n/2 + 1
  int pitch = 0;                // complex elements per (kx,ky) row of delta_k
  int device = 0;
  int n_shell_fields = 0;
  int64_t k2max = 0;            // 3 * (n/2)^2

  cufftHandle r2c = 0;
  bool r2c_ok = false;
  cufftHandle c2r = 0;          // single inverse transform (xi, bispectrum shells)
  bool c2r_ok = false;
  cufftHandle r2c_ip = 0;       // forward transform IN PLACE on a shell field (estimator gradients);
  bool r2c_ip_ok = false;       // only when the plan has shell fields
  // JPS_PLAN_FFT_PENCIL: three contiguous batched 1-D plans + a second delta_k-sized buffer; the
  // spectrum ends in `dk` as [kz][ky][kx] (x fastest)
  bool pencil = false;
  cufftHandle fz = 0, fy = 0, fx = 0;
  bool fz_ok = false, fy_ok = false, fx_ok = false;
  float2* dk2 = nullptr;
  float2* ztw = nullptr;        // [n/4 + 1] twiddles e^{-2 pi i k / n} of the real-to-complex untangle step

  // workspace partition (all device pointers inside the caller's workspace)
  char* ws = nullptr;
  size_t ws_bytes = 0;
  float2* dk = nullptr;         // [n][n][pitch] complex64
  void* fft_work = nullptr;
  size_t fft_work_bytes = 0;
  float* wlut = nullptr;        // [3][n] per-axis window factors for p = 2,3,4
  double* acc = nullptr;        // [acc_cap][4] per-call sums of the l=0,2,4 weights (4th slot spare)
  double* scal = nullptr;       // [1024] small float64 scratch (bispectrum sums)
  // Sums over the shell INDICATOR fields (sum I_j^2, sum I_0 I_1 I_j) depend only on (N, shell
  // bounds), not on the data: they are computed once, kept ON THE DEVICE in `isum` (so no host
  // read-back / synchronisation is ever needed) and found again through this host-side key -> slot map.
  double* isum = nullptr;       // [kIsumSlots]
  std::map<std::vector<int>, int> isum_slot;
  float* shell = nullptr;       // n_shell_fields real fields [n][n][2*pitch] (in-place C2R layout)
  int acc_cap = 0;

  jps::BinTable tables[jps::kNumTables];
  unsigned long long stamp = 0;
};

namespace jps {

// host side: window factors as the reference computes them in float32
void host_window_axis(int n, int p, float* out);

// powspec.cu
int ensure_bin_table(jps_plan* plan, const float* kedges_grid, int nb, int mode, cudaStream_t s,
                     BinTable** out);
int64_t edge_threshold(float e, bool strict, int64_t k2max);
int forward_fft(jps_plan* plan, const float* mesh, cudaStream_t s);
// powspec.cu: the real-to-complex untangle step fused with a transpose (see r2c_untangle_transpose_kernel):
// Z[b][y][n/2] (C2C of the real lines read as complex pairs) -> out[b][kz][y], b = 0 .. batch-1
int launch_r2c_untangle_transpose(const float2* Z, float2* out, const float2* tw, int n, int batch, cudaStream_t s);
void host_r2c_twiddles(int n, std::vector<float2>& tw);      // e^{-2 pi i k / n}, k = 0 .. n/4, evaluated in double

// slab.cu: fold +-kx / window / Legendre weights / k-bin sums of a spectrum stored x-fastest, into
// tables->acc (zeroed first).  kz_major = 0: dk[yl][kz][x] (y-shard [y0, y0 + nyl) of a slab rank);
// kz_major = 1: dk[kz][yl][x] (the pencil plan's layout, nyl = n, y0 = 0: the +-ky partner row is local and is
// folded too).  dc: device Re rho_hat(0).
int bin_xfast_layout(jps_plan* tables, const float2* dk, int nyl, int y0, int kz_major, const BinTable& T,
                     const float* dc, int normalise, int mas_order, cudaStream_t s);
double ref_volume(float box_size, int n);
float ref_kF(float box_size);

}  // namespace jps
