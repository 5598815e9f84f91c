// Mock generator on the device: Gaussian field in k-space, and
// Poisson sampling of a density mesh into particles.
//
// Replaces, behind the same function names in jax_powspec_b200/mocks.py:
//   gaussian_field(grid, kf, Pkf, Rayleigh_sampling, seed, BoxSize)   /root/reference/src/gauss_field.py:5-80
//   populate_field(rho, n_bins, box_size, density, key)               /root/reference/src/populate_field.py:11-29
// and the pipeline that strings them together, /root/reference/tests/create_lognormal.py:44-55
// (Gaussian field -> irfftn -> exp(b1 * g) -> populate).  The per-mode / per-cell arithmetic lives in
// mockgen.cuh (shared with a host build for the CPU tests); this file is the launch geometry:
//
//   M1 gaussian_field_kernel : one thread per stored mode, 8-byte coalesced stores, P(k) table in L1/L2
//   M2 density_partial/_final: mean of rho (or of exp(bias g)) in float64, fixed reduction tree
//                              (deterministic: the Poisson rates must not depend on atomic order)
//   M3 poisson_count_kernel  : one Philox counter per cell -> count; per-CTA totals
//   M4 scan_totals_kernel    : exclusive scan of the per-CTA totals (one CTA, 64-bit)
//   M5 populate_fill_kernel  : CTA-local scan of the counts in cell order, then every thread writes its
//                              cells' particles at their final rows of the [Np][3] float32 output
//
// All of it is HBM-bound streaming: M1 writes 8 B per mode, M3 reads 4 B and writes 4 B per cell,
// M5 reads 4 B per cell and writes 12 B per particle.  Particles come out grouped by cell in C order
// (the reference orders cells by descending density first, populate_field.py:18-20 -- an artefact of
// its argsort that no caller relies on).
#include "common.cuh"
#include "mockgen.cuh"

#include <algorithm>

namespace jps {

using namespace mock;

constexpr int MOCK_THREADS = 256;
constexpr int MOCK_ITEMS = 8;                               // cells per thread in M3 / M5
constexpr int MOCK_CHUNK = MOCK_THREADS * MOCK_ITEMS;       // cells per CTA
constexpr int MOCK_PARTIALS = 148 * 8;                      // CTAs of the density reduction

__global__ void __launch_bounds__(MOCK_THREADS) gaussian_field_kernel(float2* __restrict__ dk, int n,
                                                                      const double* __restrict__ kf,
                                                                      const double* __restrict__ pkf, int nk,
                                                                      int rayleigh, unsigned long long seed,
                                                                      double box_size) {
  const int nzc = n / 2 + 1;
  const size_t total = (size_t)n * n * nzc;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int iz = (int)(i % nzc);
    const size_t row = i / nzc;
    const int iy = (int)(row % n), ix = (int)(row / n);
    float re, im;
    gaussian_mode(ix, iy, iz, n, kf, pkf, nk, rayleigh, seed, box_size, re, im);
    dk[i] = make_float2(re, im);
  }
}

// ---- M2: deterministic float64 mean
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;                                                 // valid in thread 0
}

__global__ void __launch_bounds__(MOCK_THREADS) density_partial_kernel(const float* __restrict__ rho, size_t ncell,
                                                                       int lognormal, double bias,
                                                                       double* __restrict__ partial) {
  __shared__ double red[MOCK_THREADS / 32];
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += stride)
    acc += cell_density(rho[i], lognormal, bias);
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(MOCK_THREADS) density_final_kernel(const double* __restrict__ partial, int np,
                                                                     double* __restrict__ sum_out) {
  __shared__ double red[MOCK_THREADS / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) acc += partial[i];
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) *sum_out = t;
}

// ---- M3
__global__ void __launch_bounds__(MOCK_THREADS) poisson_count_kernel(const float* __restrict__ rho, size_t ncell,
                                                                     int lognormal, double bias,
                                                                     const double* __restrict__ sum, double mean_obj,
                                                                     unsigned long long seed,
                                                                     unsigned* __restrict__ counts,
                                                                     unsigned long long* __restrict__ block_tot) {
  __shared__ unsigned long long red[MOCK_THREADS / 32];
  const double scale = mean_obj / (*sum / (double)ncell);   // mean_obj_per_cell / rho.mean(), populate_field.py:16
  const size_t base = (size_t)blockIdx.x * MOCK_CHUNK;
  unsigned long long mine = 0;
#pragma unroll 1
  for (int it = 0; it < MOCK_ITEMS; ++it) {
    const size_t c = base + (size_t)it * MOCK_THREADS + threadIdx.x;
    if (c < ncell) {
      const unsigned k = poisson_draw(cell_density(rho[c], lognormal, bias) * scale, seed, (uint64_t)c);
      counts[c] = k;
      mine += k;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < MOCK_THREADS / 32; ++w) t += red[w];
    block_tot[blockIdx.x] = t;
  }
}

// CTA-wide exclusive scan of one value per thread (64-bit); returns the exclusive prefix and leaves the
// CTA total in *total_out (same value in every thread).  `warp_tot` holds blockDim/32 + 1 entries.
__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* warp_tot,
                                                                   unsigned long long* total_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += o;
  }
  __syncthreads();                                          // warp_tot may still be read from the previous call
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  unsigned long long before = 0, total = 0;
  for (int w = 0; w < nwarp; ++w) {
    const unsigned long long t = warp_tot[w];
    if (w < warp) before += t;
    total += t;
  }
  *total_out = total;
  return before + inc - v;
}

// ---- M4: exclusive scan of block_tot[nb] in place; grand total to total_out[0] (int64) .
__global__ void __launch_bounds__(1024) scan_totals_kernel(unsigned long long* __restrict__ block_tot, int nb,
                                                           long long* __restrict__ total_out) {
  __shared__ unsigned long long warp_tot[33];
  unsigned long long carry = 0;
  for (int base = 0; base < nb; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const unsigned long long v = i < nb ? block_tot[i] : 0ull;
    unsigned long long chunk_total;
    const unsigned long long ex = block_exclusive_scan(v, warp_tot, &chunk_total);
    if (i < nb) block_tot[i] = carry + ex;
    carry += chunk_total;
  }
  if (threadIdx.x == 0) *total_out = (long long)carry;
}

// ---- M5
__global__ void __launch_bounds__(MOCK_THREADS) populate_fill_kernel(const unsigned* __restrict__ counts, size_t ncell,
                                                                     const unsigned long long* __restrict__ block_off,
                                                                     int n, float bin_size, float box,
                                                                     unsigned long long seed, long long n_out,
                                                                     float* __restrict__ pos) {
  __shared__ unsigned long long warp_tot[MOCK_THREADS / 32 + 1];
  const size_t base = (size_t)blockIdx.x * MOCK_CHUNK;
  unsigned long long run = block_off[blockIdx.x];           // first output row of this CTA's cells
  const size_t n2 = (size_t)n * n;
#pragma unroll 1
  for (int it = 0; it < MOCK_ITEMS; ++it) {                 // cell order inside the chunk: it-major, thread-minor
    const size_t c = base + (size_t)it * MOCK_THREADS + threadIdx.x;
    const unsigned m = c < ncell ? counts[c] : 0u;
    unsigned long long round_total;
    const unsigned long long first = run + block_exclusive_scan((unsigned long long)m, warp_tot, &round_total);
    run += round_total;
    if (m) {
      const int ix = (int)(c / n2), iy = (int)((c / n) % n), iz = (int)(c % n);
      for (unsigned j = 0; j < m; ++j) {
        const unsigned long long p = first + j;
        if ((long long)p >= n_out) break;
        float x, y, z;
        particle_position(ix, iy, iz, (uint64_t)p, seed, bin_size, box, x, y, z);
        float* o = pos + 3 * (size_t)p;
        o[0] = x; o[1] = y; o[2] = z;
      }
    }
  }
}

struct PopulateLayout { size_t sum, partial, block_tot, counts, total; int nblocks; };

static PopulateLayout populate_layout(int n) {
  PopulateLayout L;
  const size_t ncell = (size_t)n * n * n;
  L.nblocks = (int)((ncell + MOCK_CHUNK - 1) / MOCK_CHUNK);
  size_t off = 0;
  auto take = [&](size_t b) { const size_t o = off; off = align_up(off + b, 256); return o; };
  L.sum = take(16);                                          // [0] sum of the density, [1] grand total (as int64)
  L.partial = take((size_t)MOCK_PARTIALS * 8);
  L.block_tot = take((size_t)L.nblocks * 8);
  L.counts = take(ncell * 4);
  L.total = off;
  return L;
}

}  // namespace jps

using namespace jps;

extern "C" size_t jps_mock_field_workspace_bytes(int nk) { return nk > 0 ? align_up((size_t)nk * 16, 256) : 0; }

extern "C" int jps_mock_gaussian_field(int n_mesh, const double* kf, const double* pkf, int nk, int rayleigh,
                                       unsigned long long seed, float box_size, void* delta_k, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  JPS_REQUIRE(n_mesh >= 2 && n_mesh <= 4096, "jps_mock_gaussian_field: n_mesh %d out of range [2,4096]", n_mesh);
  JPS_REQUIRE(kf && pkf && nk >= 2, "jps_mock_gaussian_field: the P(k) table needs at least two points");
  JPS_REQUIRE(box_size > 0.0f, "jps_mock_gaussian_field: box_size must be > 0");
  JPS_REQUIRE(delta_k != nullptr, "jps_mock_gaussian_field: NULL output");
  if (workspace == nullptr || workspace_bytes < jps_mock_field_workspace_bytes(nk)) {
    set_error("jps_mock_gaussian_field: workspace has %zu bytes, %zu needed (jps_mock_field_workspace_bytes)",
              workspace_bytes, jps_mock_field_workspace_bytes(nk));
    return JPS_ERR_WORKSPACE;
  }
  for (int i = 1; i < nk; ++i)
    JPS_REQUIRE(kf[i] > kf[i - 1], "jps_mock_gaussian_field: kf must be strictly increasing (entry %d)", i);
  cudaStream_t s = (cudaStream_t)stream;
  double* tab = (double*)workspace;
  JPS_CHECK_CUDA(cudaMemcpyAsync(tab, kf, (size_t)nk * 8, cudaMemcpyHostToDevice, s));
  JPS_CHECK_CUDA(cudaMemcpyAsync(tab + nk, pkf, (size_t)nk * 8, cudaMemcpyHostToDevice, s));
  const size_t total = (size_t)n_mesh * n_mesh * (n_mesh / 2 + 1);
  const int blocks = (int)std::min<size_t>((total + MOCK_THREADS - 1) / MOCK_THREADS, (size_t)kNumSMs * 32);
  {
    ScopedLaunch L(K_MOCK_FIELD, s);
    gaussian_field_kernel<<<blocks, MOCK_THREADS, 0, s>>>((float2*)delta_k, n_mesh, tab, tab + nk, nk, rayleigh ? 1 : 0,
                                                          seed, (double)box_size);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

extern "C" size_t jps_mock_populate_workspace_bytes(int n_mesh) {
  return n_mesh >= 1 ? populate_layout(n_mesh).total : 0;
}

extern "C" size_t jps_mock_populate_counts_offset(int n_mesh) {
  return n_mesh >= 1 ? populate_layout(n_mesh).counts : 0;
}

extern "C" int jps_mock_populate_count(const float* rho, int n_mesh, float box_size, float density, int lognormal,
                                       float bias, unsigned long long seed, void* workspace, size_t workspace_bytes,
                                       int64_t* total, void* stream) {
  JPS_REQUIRE(rho && total, "jps_mock_populate_count: NULL argument");
  JPS_REQUIRE(n_mesh >= 1 && n_mesh <= 4096, "jps_mock_populate_count: n_mesh %d out of range [1,4096]", n_mesh);
  JPS_REQUIRE(box_size > 0.0f && density >= 0.0f, "jps_mock_populate_count: box_size must be > 0 and density >= 0");
  const PopulateLayout L = populate_layout(n_mesh);
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_error("jps_mock_populate_count: workspace has %zu bytes, %zu needed (jps_mock_populate_workspace_bytes)",
              workspace_bytes, L.total);
    return JPS_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const size_t ncell = (size_t)n_mesh * n_mesh * n_mesh;
  const double bin = (double)box_size / (double)n_mesh;
  const double mean_obj = bin * bin * bin * (double)density;            // populate_field.py:12-14
  {
    ScopedLaunch T(K_MOCK_POPULATE, s);
    density_partial_kernel<<<MOCK_PARTIALS, MOCK_THREADS, 0, s>>>(rho, ncell, lognormal ? 1 : 0, (double)bias,
                                                                   (double*)(ws + L.partial));
  }
  JPS_CHECK_LAUNCH();
  {
    ScopedLaunch T(K_MOCK_POPULATE, s);
    density_final_kernel<<<1, MOCK_THREADS, 0, s>>>((const double*)(ws + L.partial), MOCK_PARTIALS, (double*)(ws + L.sum));
  }
  JPS_CHECK_LAUNCH();
  {
    ScopedLaunch T(K_MOCK_POPULATE, s);
    poisson_count_kernel<<<L.nblocks, MOCK_THREADS, 0, s>>>(rho, ncell, lognormal ? 1 : 0, (double)bias,
                                                             (const double*)(ws + L.sum), mean_obj, seed,
                                                             (unsigned*)(ws + L.counts),
                                                             (unsigned long long*)(ws + L.block_tot));
  }
  JPS_CHECK_LAUNCH();
  {
    ScopedLaunch T(K_MOCK_POPULATE, s);
    scan_totals_kernel<<<1, 1024, 0, s>>>((unsigned long long*)(ws + L.block_tot), L.nblocks, (long long*)total);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

extern "C" int jps_mock_populate_fill(int n_mesh, float box_size, unsigned long long seed, const void* workspace,
                                      size_t workspace_bytes, int64_t n_out, float* pos, void* stream) {
  JPS_REQUIRE(n_mesh >= 1 && n_mesh <= 4096, "jps_mock_populate_fill: n_mesh %d out of range [1,4096]", n_mesh);
  JPS_REQUIRE(box_size > 0.0f && n_out >= 0, "jps_mock_populate_fill: box_size must be > 0 and n_out >= 0");
  const PopulateLayout L = populate_layout(n_mesh);
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_error("jps_mock_populate_fill: workspace has %zu bytes, %zu needed", workspace_bytes, L.total);
    return JPS_ERR_WORKSPACE;
  }
  if (n_out == 0) return JPS_OK;
  JPS_REQUIRE(pos != nullptr, "jps_mock_populate_fill: NULL output");
  cudaStream_t s = (cudaStream_t)stream;
  const char* ws = (const char*)workspace;
  const size_t ncell = (size_t)n_mesh * n_mesh * n_mesh;
  const float bin_size = box_size / (float)n_mesh;
  {
    ScopedLaunch T(K_MOCK_POPULATE, s);
    populate_fill_kernel<<<L.nblocks, MOCK_THREADS, 0, s>>>((const unsigned*)(ws + L.counts), ncell,
                                                             (const unsigned long long*)(ws + L.block_tot), n_mesh,
                                                             bin_size, box_size, seed, (long long)n_out, pos);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}
