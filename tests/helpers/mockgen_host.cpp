// Host build of jax_powspec_b200/csrc/mockgen.cuh (the per-mode / per-cell arithmetic the device mock
// generator runs) so that the CPU test-suite can check the generator's statistics and known answers
// without a GPU, and the GPU tests can compare the kernels with it element by element.
// Test infrastructure only: nothing in the package loads this.
#include "../../jax_powspec_b200/csrc/mockgen.cuh"

#include <vector>

using namespace jps::mock;

extern "C" {

void mock_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  U4 c{ctr[0], ctr[1], ctr[2], ctr[3]};
  const U4 r = philox4x32_10(c, key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

double mock_uniform_open(uint32_t a, uint32_t b) { return uniform_open(a, b); }
float mock_uniform_f32(uint32_t a) { return uniform_f32(a); }
double mock_interp_power(const double* kf, const double* pkf, int nk, double k) { return interp_power(kf, pkf, nk, k); }

// out: float pairs [n][n][n/2+1][2]
void mock_gaussian_field(int n, const double* kf, const double* pkf, int nk, int rayleigh, unsigned long long seed,
                         double box, float* out) {
  const int nzc = n / 2 + 1;
  for (int ix = 0; ix < n; ++ix)
    for (int iy = 0; iy < n; ++iy)
      for (int iz = 0; iz < nzc; ++iz) {
        float re, im;
        gaussian_mode(ix, iy, iz, n, kf, pkf, nk, rayleigh, seed, box, re, im);
        float* o = out + 2 * (((size_t)ix * n + iy) * nzc + iz);
        o[0] = re; o[1] = im;
      }
}

// The two uniforms (phase, amplitude) mode (ix, iy, iz) draws from ITS OWN counter, in C order: fed to the
// unmodified reference loop in place of np.random.random() they must reproduce mock_gaussian_field
// (oracle/make_mock_golden.py).  out: double [n][n][n/2+1][2]
void mock_field_uniforms(int n, unsigned long long seed, double* out) {
  const int nzc = n / 2 + 1;
  for (unsigned long long flat = 0; flat < (unsigned long long)n * n * nzc; ++flat) {
    const U4 r = draw(seed, STREAM_FIELD, flat, 0u);
    out[2 * flat] = uniform_open(r.x, r.y);
    out[2 * flat + 1] = uniform_open(r.z, r.w);
  }
}

// `count` draws at the same rate, cells cell0 .. cell0 + count - 1
void mock_poisson_many(double lam, unsigned long long seed, unsigned long long cell0, long long count, uint32_t* out) {
  for (long long i = 0; i < count; ++i) out[i] = poisson_draw(lam, seed, cell0 + (unsigned long long)i);
}

// The device's fixed float64 reduction tree (mockgen.cu: density_partial_kernel + density_final_kernel,
// `blocks` CTAs of `threads` threads), replayed serially so that the Poisson rates are bit-identical.
static double tree_sum(const std::vector<double>& per_thread, int threads) {
  // per_thread: one accumulator per thread of ONE block; shuffle-down tree per warp, then warps in order
  double total = 0.0;
  for (int w = 0; w < threads / 32; ++w) {
    double v[32];
    for (int l = 0; l < 32; ++l) v[l] = per_thread[(size_t)w * 32 + l];
    for (int off = 16; off > 0; off >>= 1)
      for (int l = 0; l < 32; ++l) v[l] = v[l] + (l + off < 32 ? v[l + off] : v[l]);   // shfl_down: out of range keeps own value
    total += v[0];
  }
  return total;
}

double mock_density_sum(const float* rho, long long ncell, int lognormal, double bias, int blocks, int threads) {
  std::vector<double> partial((size_t)blocks);
  const long long stride = (long long)blocks * threads;
  std::vector<double> acc((size_t)threads);
  for (int b = 0; b < blocks; ++b) {
    for (int t = 0; t < threads; ++t) {
      double a = 0.0;
      for (long long i = (long long)b * threads + t; i < ncell; i += stride) a += cell_density(rho[i], lognormal, bias);
      acc[(size_t)t] = a;
    }
    partial[(size_t)b] = tree_sum(acc, threads);
  }
  for (int t = 0; t < threads; ++t) {
    double a = 0.0;
    for (int i = t; i < blocks; i += threads) a += partial[(size_t)i];
    acc[(size_t)t] = a;
  }
  return tree_sum(acc, threads);
}

// counts[ncell]; returns the total
long long mock_populate_count(const float* rho, int n, double box, double density, int lognormal, double bias,
                              unsigned long long seed, double sum, uint32_t* counts) {
  const long long ncell = (long long)n * n * n;
  const double bin = box / (double)n;
  const double mean_obj = bin * bin * bin * density;
  const double scale = mean_obj / (sum / (double)ncell);
  long long total = 0;
  for (long long c = 0; c < ncell; ++c) {
    counts[c] = poisson_draw(cell_density(rho[c], lognormal, bias) * scale, seed, (unsigned long long)c);
    total += counts[c];
  }
  return total;
}

void mock_populate_fill(const uint32_t* counts, int n, float box, unsigned long long seed, float* pos) {
  const float bin_size = box / (float)n;
  unsigned long long p = 0;
  for (int ix = 0; ix < n; ++ix)
    for (int iy = 0; iy < n; ++iy)
      for (int iz = 0; iz < n; ++iz) {
        const uint32_t m = counts[((size_t)ix * n + iy) * n + iz];
        for (uint32_t j = 0; j < m; ++j, ++p) {
          float x, y, z;
          particle_position(ix, iy, iz, p, seed, bin_size, box, x, y, z);
          pos[3 * p] = x; pos[3 * p + 1] = y; pos[3 * p + 2] = z;
        }
      }
}

void mock_tri_offsets(const uint32_t* bits, long long count, float bin_size, float* out) {
  for (long long i = 0; i < count; ++i) out[i] = tri_offset(bits[i], bin_size);
}
}
