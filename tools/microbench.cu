// Micro-benchmarks that decide the painter design on B200 (run under gpurun, not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu -lcufft
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s){ s = s*1664525u + 1013904223u; return s; }
__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352du; x^=x>>15; x*=0x846ca68bu; x^=x>>16; return x; }

// random scalar reds into region of `cells` floats
__global__ void k_red_scalar(float* g, uint32_t cells_mask, int iters){
  uint32_t s = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ uint32_t a = hash32(lcg(s)) & cells_mask; atomicAdd(g+a, 1.0f); }
}
// random v4 reds (16B aligned)
__global__ void k_red_v4(float* g, uint32_t cells_mask, int iters){
  uint32_t s = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ uint32_t a = (hash32(lcg(s)) & cells_mask) & ~3u; float* p=g+a;
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"::"l"(p),"f"(1.f),"f"(1.f),"f"(1.f),"f"(1.f):"memory"); }
}
// v2
__global__ void k_red_v2(float* g, uint32_t cells_mask, int iters){
  uint32_t s = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ uint32_t a = (hash32(lcg(s)) & cells_mask) & ~1u; float* p=g+a;
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};"::"l"(p),"f"(1.f),"f"(1.f):"memory"); }
}
// CIC-like: per "particle" 4 rows x (2 contiguous cells) scalar reds, rows offset by n and n*n (locality like a real stencil)
__global__ void k_red_cic(float* g, uint32_t cells_mask, int n, int iters){
  uint32_t s = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ uint32_t a = hash32(lcg(s)) & cells_mask;
    #pragma unroll
    for(int r=0;r<4;r++){ uint32_t b = (a + (r&1)*n + (r>>1)*n*n) & cells_mask; atomicAdd(g+b,1.f); atomicAdd(g+((b+1)&cells_mask),1.f);} }
}
// smem CAS float atomics into a tile of `tile` floats
__global__ void k_smem_atomic(float* out, int tile, int iters){
  extern __shared__ float sm[];
  for(int i=threadIdx.x;i<tile;i+=blockDim.x) sm[i]=0.f;
  __syncthreads();
  uint32_t s = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ uint32_t a = hash32(lcg(s)) % tile; atomicAdd(sm+a, 1.0f); }
  __syncthreads();
  float acc=0; for(int i=threadIdx.x;i<tile;i+=blockDim.x) acc+=sm[i];
  if(acc==-1.f) out[0]=acc;
}
// smem int atomics (native)
__global__ void k_smem_atomic_int(int* out, int tile, int iters){
  extern __shared__ int smi[];
  for(int i=threadIdx.x;i<tile;i+=blockDim.x) smi[i]=0;
  __syncthreads();
  uint32_t s = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ uint32_t a = hash32(lcg(s)) % tile; atomicAdd(smi+a, 1); }
  __syncthreads();
  int acc=0; for(int i=threadIdx.x;i<tile;i+=blockDim.x) acc+=smi[i];
  if(acc==-1) out[0]=acc;
}
// smem plain RMW (racy; measures the non-atomic LDS+FADD+STS rate the coloured scheme would get)
__global__ void k_smem_rmw(float* out, int tile, int iters){
  extern __shared__ float sm[];
  for(int i=threadIdx.x;i<tile;i+=blockDim.x) sm[i]=0.f;
  __syncthreads();
  uint32_t s = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ uint32_t a = hash32(lcg(s)) % tile; volatile float* p = sm+a; *p = *p + 1.0f; }
  __syncthreads();
  float acc=0; for(int i=threadIdx.x;i<tile;i+=blockDim.x) acc+=sm[i];
  if(acc==-1.f) out[0]=acc;
}
// scattered 16B stores (bucket scatter emulation) into region of n4 float4
__global__ void k_scatter16(float4* g, uint32_t mask, int iters){
  uint32_t s = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ uint32_t a = hash32(lcg(s)) & mask; g[a] = make_float4(1,2,3,4); }
}
__global__ void k_copy(const float4* a, float4* b, size_t n){ for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x;i<n;i+=(size_t)gridDim.x*blockDim.x) b[i]=a[i]; }

template<class F> float timeit(F f, int reps=3){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; }
  return best;
}

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  printf("device %s SMs %d L2 %d MB\n", pr.name, pr.multiProcessorCount, pr.l2CacheSize>>20);
  const int blocks = 148*8, threads=256, iters=256;
  const double nthreads = (double)blocks*threads;
  float* g; CK(cudaMalloc(&g, (size_t)1<<31));  // 2 GB
  CK(cudaMemset(g,0,(size_t)1<<31));
  for(int lg : {20, 23, 24, 25, 27, 29}){   // region floats: 4MB, 32MB, 64MB, 128MB, 512MB, 2GB
    uint32_t mask=(1u<<lg)-1;
    float t1=timeit([&]{k_red_scalar<<<blocks,threads>>>(g,mask,iters);});
    float t2=timeit([&]{k_red_v2<<<blocks,threads>>>(g,mask,iters);});
    float t4=timeit([&]{k_red_v4<<<blocks,threads>>>(g,mask,iters);});
    float tc=timeit([&]{k_red_cic<<<blocks,threads>>>(g,mask,1<<(lg/3),iters/8);});
    printf("region %5d MB: red.f32 %.1f G/s | red.v2 %.1f Gop/s | red.v4 %.1f Gop/s | cic-like(8 reds) %.2f Gpart/s\n",
      (int)((4ull<<lg)>>20), nthreads*iters/t1/1e6, nthreads*iters/t2/1e6, nthreads*iters/t4/1e6, nthreads*(iters/8)/tc/1e6);
  }
  for(int tile : {2048, 8192, 32768}){
    size_t sm = tile*4; 
    cudaFuncSetAttribute(k_smem_atomic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
    cudaFuncSetAttribute(k_smem_atomic_int, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
    cudaFuncSetAttribute(k_smem_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
    int b2 = 148*4; double nt=(double)b2*threads; int it=2048;
    float ta=timeit([&]{k_smem_atomic<<<b2,threads,sm>>>(g,tile,it);});
    float ti=timeit([&]{k_smem_atomic_int<<<b2,threads,sm>>>((int*)g,tile,it);});
    float tr=timeit([&]{k_smem_rmw<<<b2,threads,sm>>>(g,tile,it);});
    printf("smem tile %6d floats: float atomicAdd(CAS) %.1f G/s | int atomicAdd %.1f G/s | plain RMW %.1f G/s\n", tile, nt*it/ta/1e6, nt*it/ti/1e6, nt*it/tr/1e6);
  }
  { uint32_t mask=(1u<<27)-1; float t=timeit([&]{k_scatter16<<<blocks,threads>>>((float4*)g,mask,iters);});
    printf("scattered 16B stores into 2GB: %.1f G stores/s = %.1f GB/s useful\n", nthreads*iters/t/1e6, nthreads*iters*16/t/1e6); }
  { size_t n=(size_t)1<<26; float t=timeit([&]{k_copy<<<148*16,256>>>((float4*)g,(float4*)g+n,n);});
    printf("copy 1GB->1GB: %.1f GB/s (r+w)\n", 2.0*n*16/t/1e6); }
  // cuFFT R2C
  for(int n : {256, 512, 1024}){
    cufftHandle h; size_t ws; cufftCreate(&h); long long d[3]={n,n,n};
    if(cufftMakePlanMany64(h,3,d,NULL,1,0,NULL,1,0,CUFFT_R2C,1,&ws)!=CUFFT_SUCCESS){printf("plan fail %d\n",n);continue;}
    float* in=g; cufftComplex* out=(cufftComplex*)(g + ((size_t)1<<30)/4*1);  // out at +1GB
    if((size_t)n*n*n*4 > ((size_t)1<<30) ){ printf("skip fft %d (buffer)\n",n); cufftDestroy(h); continue; }
    float t=timeit([&]{cufftExecR2C(h,in,out);});
    printf("cuFFT R2C %d^3: %.3f ms  (24 N^3 model: %.1f GB/s) work area %.1f MB\n", n, t, 24.0*n*n*n/t/1e6, ws/1048576.0);
    cufftDestroy(h);
  }
  return 0;
}
