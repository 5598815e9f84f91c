// Shared-memory integer atomics on B200: throughput of ATOMS.ADD (with and without a returned value) against the bank
// pattern of the 32 lanes of an instruction.  Decides whether the tile deposit (csrc/paint_sorted.cu) can gain from a
// bank-aware particle order.      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_atoms tools/microbench_atoms.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;} }while(0)
__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352du; x^=x>>15; x*=0x846ca68bu; x^=x>>16; return x; }
constexpr int WORDS = 8192;        // 32 KB tile
constexpr int NPAT = 64;           // address patterns per thread, cycled

// pattern 0: 32 distinct banks (a row of 32 consecutive words, row varies)   1: random words   2: two lanes per bank
// 3: four lanes per bank   4: all lanes one bank (32 rows)   5: all lanes ONE word   6: groups of 4 consecutive words at
// 8 random 4-aligned positions (a PCS z-row per particle, 8 particles per instruction)
template <bool RET>
__global__ void __launch_bounds__(512) k_atoms(unsigned* out, int pattern, int iters) {
  __shared__ unsigned sm[WORDS];
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sm[i] = 0u;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned idx[NPAT];
#pragma unroll
  for (int j = 0; j < NPAT; ++j) {
    const unsigned row = hash32(warp * 131u + j * 7u + blockIdx.x) % (WORDS / 32);
    unsigned a;
    switch (pattern) {
      case 0: a = row * 32u + lane; break;
      case 1: a = hash32(threadIdx.x * 977u + j * 13u + blockIdx.x * 7919u) % WORDS; break;
      case 2: a = ((row + (lane & 1u) * 3u) % (WORDS / 32)) * 32u + (lane >> 1); break;
      case 3: a = ((row + (lane & 3u) * 5u) % (WORDS / 32)) * 32u + (lane >> 2); break;
      case 4: a = ((row + lane) % (WORDS / 32)) * 32u + (j & 31u); break;
      case 5: a = row * 32u + (j & 31u); break;
      default: a = (hash32((warp * 8u + (lane >> 2)) * 31u + j * 17u + blockIdx.x) % (WORDS / 4)) * 4u + (lane & 3u); break;
    }
    idx[j] = a;
  }
  __syncthreads();
  unsigned acc = 0u;
  const unsigned val = threadIdx.x * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NPAT; ++j) {
      if (RET) acc += atomicAdd(sm + idx[j], val);
      else atomicAdd(sm + idx[j], val);
    }
  }
  __syncthreads();
  unsigned s = acc;
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) s += sm[i];
  if (s == 0xdeadbeefu) out[0] = s;
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  unsigned* out; CK(cudaMalloc(&out, 4));
  const int blocks = pr.multiProcessorCount * 3, threads = 512, iters = 64;
  const char* names[7] = {"32 distinct banks", "random words", "2 lanes per bank", "4 lanes per bank", "32 lanes one bank",
                          "32 lanes one word", "8 x (4 consecutive words)"};
  printf("device %s, %d SMs, max clock %.0f MHz; %d CTAs x %d threads, %d atomics per thread\n", pr.name, pr.multiProcessorCount,
         clk_khz / 1e3, blocks, threads, iters * NPAT);
  for (int ret = 0; ret < 2; ++ret)
    for (int p = 0; p < 7; ++p) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        if (ret) k_atoms<true><<<blocks, threads>>>(out, p, iters); else k_atoms<false><<<blocks, threads>>>(out, p, iters);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
      }
      const double lane_ops = (double)blocks * threads * iters * NPAT;
      const double per_clk_sm = lane_ops / (best * 1e-3) / pr.multiProcessorCount / (clk_khz * 1e3);
      printf("%-10s %-28s %8.3f ms  %7.1f G lane-atomics/s  %5.2f lanes/clk/SM  %5.2f clk per warp instruction\n",
             ret ? "ATOMS ret" : "ATOMS", names[p], best, lane_ops / best / 1e6, per_clk_sm, 32.0 / per_clk_sm);
    }
  return 0;
}
