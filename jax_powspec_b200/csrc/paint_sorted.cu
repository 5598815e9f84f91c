// Bucketed painter: K1 (bucket_count / bucket_scan / bucket_scatter) + K2 (paint_tile).
//
// Why (measured on B200, tools/microbench.cu, profiles/r1_microbench.txt): a 512^3 mesh
// (537 MB) does not fit the 126 MB L2, and random global reds then run at ~25 G/s (DRAM sector
// read-modify-write) against ~190 G/s when the target is L2 resident and ~750 G/s for
// shared-memory float atomics.  So particles are first bucketed by the 16^3-cell tile that
// holds the lowest node of their stencil; one CTA per tile then deposits its bucket into a
// shared-memory copy of the tile (+ (order-1) halo cells on the high side of each axis) and
// flushes it once with 16-byte vector reds (red.global.add.v4.f32) -- every mesh cell is
// touched by O(1) global operations instead of O(particles).
//
// Bucketed record = (px, py, pz, w): the float32 grid coordinate (x-xmin)*inv the reference
// computes (src/mas.py:103-105) and the weight, so K2 needs no further per-catalogue scalars
// and cell choice stays bit-identical to the reference's float32 arithmetic.
#include "paint_common.cuh"

#include <cuda.h>      // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cstdlib>

namespace jps {

constexpr int TILE = 16;                 // cells per tile side

struct TileGeom {
  int n;        // global mesh side
  int x0, nx;   // local slab: global plane of local plane 0, allocated planes (0, n for a full mesh)
  int nt;       // tiles per y / z axis = ceil(n / TILE)
  int ntx;      // tiles along x = ceil(nx / TILE)
  int ntiles;   // ntx * nt * nt
  int two_level;  // 1: K1 v2 (two-level partition, rep == 1); 0: single-level atomic-cursor scatter
  int rep;      // counter replicas per tile (power of two): bucket = tile*rep + (block & (rep-1)).
                // Same-address global atomics serialise in L2; with ~3000 particles per tile and
                // ~3e5 threads in flight the single-counter version ran at half the red rate of
                // spread addresses.  Replicas of one tile are adjacent, so a tile's particles stay
                // contiguous: [offsets[tile*rep], offsets[(tile+1)*rep]).
                // Bucket ntiles*rep collects the particles the tile kernel cannot take.
};

// Lowest stencil node of a particle along one axis, wrapped into [0, n); -1 if the particle
// cannot be handled by the tile kernel (reference-compat CIC outside the box).
template <int ORDER>
__device__ __forceinline__ int anchor_base(float pos) {      // unwrapped lowest node of the B-spline stencil
  if (ORDER == 2) return (int)floorf(pos);
  if (ORDER == 3) return (int)floorf(pos + 0.5f) - 1;
  return (int)floorf(pos) - 1;
}

template <int ORDER, bool REFCIC>
__device__ __forceinline__ int anchor_axis(float pos, int n) {
  if (REFCIC) {
    const int i = (int)pos;              // truncation (Q3); in-box particles have 0 <= i < n
    return (pos >= 0.0f && i < n) ? i : -1;
  }
  return pymod(anchor_base<ORDER>(pos), n);
}

// bucket id of a particle; `block` selects the counter replica (count and scatter passes use
// the same particle -> block mapping, so a particle sees the same replica in both).
// The three anchors are wrapped with branch-free selects and share ONE rarely taken branch to the
// out-of-line division (particles more than a box length outside the box).
template <int ORDER, bool REFCIC>
__device__ __forceinline__ int tile_of(float px, float py, float pz, const TileGeom& g, unsigned block) {
  int ax, ay, az;
  if (REFCIC) {
    ax = anchor_axis<ORDER, REFCIC>(px, g.n);
    ay = anchor_axis<ORDER, REFCIC>(py, g.n);
    az = anchor_axis<ORDER, REFCIC>(pz, g.n);
  } else {
    const int n = g.n;
    ax = wrap_once(anchor_base<ORDER>(px), n);
    ay = wrap_once(anchor_base<ORDER>(py), n);
    az = wrap_once(anchor_base<ORDER>(pz), n);
    if (((unsigned)ax >= (unsigned)n) | ((unsigned)ay >= (unsigned)n) | ((unsigned)az >= (unsigned)n)) {
      ax = pymod_slow(ax, n);
      ay = pymod_slow(ay, n);
      az = pymod_slow(az, n);
    }
  }
  if (g.nx != g.n || g.x0 != 0) ax = local_plane(ax, g.x0, g.nx, g.n);   // slab: not held here -> dropped
  if ((ax | ay | az) < 0) return g.ntiles * g.rep;
  const unsigned t = (((unsigned)ax / TILE) * g.nt + ((unsigned)ay / TILE)) * g.nt + ((unsigned)az / TILE);
  return (int)(t * g.rep + (block & (unsigned)(g.rep - 1)));
}

// ---------------------------------------------------------------- catalogue access
// The passes that read the caller's catalogue give every thread FOUR consecutive particles.  For the
// usual (N, 3) float32 array (x, y, z interleaved, 16-byte aligned) those are 48 contiguous bytes =
// three 16-byte loads instead of twelve 4-byte loads with 64-bit index arithmetic each; any other
// layout (separate arrays, other strides) takes the scalar path.
struct Quad {
  float x[4], y[4], z[4];
};

__device__ __forceinline__ bool catalogue_is_aos(const PaintParams& p) {
  return p.stride == 3 && p.y == p.x + 1 && p.z == p.x + 2 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
}

// i is a multiple of 4; cnt = number of valid particles in the quad (1..4)
__device__ __forceinline__ void load_quad(const PaintParams& p, bool aos, int64_t i, int cnt, Quad& q) {
  if (aos && cnt == 4) {
    const float4* b = reinterpret_cast<const float4*>(p.x + 3 * i);
    const float4 a = __ldg(b), m = __ldg(b + 1), c = __ldg(b + 2);
    q.x[0] = a.x; q.y[0] = a.y; q.z[0] = a.z;
    q.x[1] = a.w; q.y[1] = m.x; q.z[1] = m.y;
    q.x[2] = m.z; q.y[2] = m.w; q.z[2] = c.x;
    q.x[3] = c.y; q.y[3] = c.z; q.z[3] = c.w;
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = (i + (u < cnt ? u : 0)) * p.stride;
      q.x[u] = p.x[j]; q.y[u] = p.y[j]; q.z[u] = p.z[j];
    }
  }
}

__device__ __forceinline__ void load_quad_weights(const PaintParams& p, int64_t i, int cnt, float (&w)[4]) {
  if (!p.w) {
    w[0] = w[1] = w[2] = w[3] = 1.0f;
  } else if (cnt == 4 && (reinterpret_cast<uintptr_t>(p.w) & 15) == 0) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + i));
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) w[u] = p.w[i + (u < cnt ? u : 0)];
  }
}

__device__ __forceinline__ float quad_wmax(const float (&w)[4], int cnt, float wmax) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float a = fabsf(w[u]);
    if (u < cnt && a < 3.0e38f) wmax = fmaxf(wmax, a);     // ignores inf / NaN
  }
  return wmax;
}

// ---------------------------------------------------------------- K1a: histogram of tile ids
// One global red per particle.  BUCKET_UNROLL particles per thread keep that many independent
// loads / atomics in flight (the pass is latency bound, not bandwidth bound).
constexpr int BUCKET_UNROLL = 4;

template <int ORDER, bool REFCIC>
__global__ void __launch_bounds__(256) bucket_count_kernel(PaintParams p, TileGeom g,
                                                           unsigned* __restrict__ counts,
                                                           unsigned* __restrict__ wmax_bits) {
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  float wmax = p.w ? 0.0f : 1.0f;                  // max |w| (sets the fixed-point scale of K2)
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < p.n_part;
       i0 += BUCKET_UNROLL * T) {
    int tile[BUCKET_UNROLL];
#pragma unroll
    for (int u = 0; u < BUCKET_UNROLL; ++u) {
      const int64_t i = i0 + u * T;
      tile[u] = -1;
      if (i < p.n_part) {
        const float px = grid_pos(p.x[i * p.stride], p.xmin, p.inv);
        const float py = grid_pos(p.y[i * p.stride], p.ymin, p.inv);
        const float pz = grid_pos(p.z[i * p.stride], p.zmin, p.inv);
        tile[u] = tile_of<ORDER, REFCIC>(px, py, pz, g, blockIdx.x);
        if (p.w) {
          const float a = fabsf(p.w[i]);
          if (a < 3.0e38f) wmax = fmaxf(wmax, a);     // ignores inf / NaN
        }
      }
    }
#pragma unroll
    for (int u = 0; u < BUCKET_UNROLL; ++u)
      if (tile[u] >= 0) atomicAdd(counts + tile[u], 1u);
  }
  if (!p.w) {                                       // unit weights: nothing to reduce
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(wmax_bits, __float_as_uint(1.0f));
    return;
  }
  __shared__ float wmax_warp[8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
  if ((threadIdx.x & 31) == 0) wmax_warp[threadIdx.x >> 5] = wmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < (int)(blockDim.x >> 5); ++q) wmax = fmaxf(wmax, wmax_warp[q]);
    atomicMax(wmax_bits, __float_as_uint(wmax));    // non-negative floats order as uints; one per CTA
  }
}

// K1a, small-mesh variant: when all tile counters fit in one SM's shared memory (<= 48 K tiles,
// i.e. N <= 576) every CTA histograms its share of the catalogue with native shared-memory integer
// atomics and adds its 128 KB table to the global counters once -- 4.8 M global reds instead of
// 1e8, which turns the pass from L2-atomic-bound (0.69 ms on C2) into a plain read of x, y, z.
constexpr int kSmemCountMaxTiles = 49152;

template <int ORDER, bool REFCIC>
__global__ void __launch_bounds__(1024) bucket_count_smem_kernel(PaintParams p, TileGeom g,
                                                                 unsigned* __restrict__ counts,
                                                                 unsigned* __restrict__ wmax_bits) {
  extern __shared__ unsigned hist[];               // [ntiles + 1]; requires g.rep == 1
  const int nb = g.ntiles + 1;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  const int64_t nquads = (p.n_part + 3) >> 2;
  const bool aos = catalogue_is_aos(p);
  float wmax = p.w ? 0.0f : 1.0f;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += T) {
    const int64_t i = q << 2;
    const int cnt = (int)min((int64_t)4, p.n_part - i);
    Quad c;
    load_quad(p, aos, i, cnt, c);
    if (p.w) {
      float w[4];
      load_quad_weights(p, i, cnt, w);
      wmax = quad_wmax(w, cnt, wmax);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int tile = tile_of<ORDER, REFCIC>(grid_pos(c.x[u], p.xmin, p.inv), grid_pos(c.y[u], p.ymin, p.inv),
                                              grid_pos(c.z[u], p.zmin, p.inv), g, 0u);
      if (u < cnt) atomicAdd(hist + tile, 1u);
    }
  }
  __syncthreads();
  // every CTA adds its table to the same global counters: rotate the start so that at any moment
  // the CTAs touch different addresses (same-address reds serialise in L2)
  const int rot = (int)(((long long)blockIdx.x * nb) / gridDim.x);
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    int i = k + rot;
    if (i >= nb) i -= nb;
    const unsigned c = hist[i];
    if (c) atomicAdd(counts + i, c);
  }
  if (!p.w) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(wmax_bits, __float_as_uint(1.0f));
    return;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
  if ((threadIdx.x & 31) == 0) atomicMax(wmax_bits, __float_as_uint(wmax));    // 32 per CTA, 148 CTAs
}

// K1a, big-mesh variant (N > 576: the tile table no longer fits one SM's shared memory): the same
// 16-byte quad loads, one global red per particle into the per-tile counters.  The table is small
// against the L2 (2048^3: 2.1 M tiles = 8.4 MB), so the reds stay on chip.  With the tile offsets known
// before the partition, the fine pass reads its records ONCE instead of twice (2048^3 on one GPU:
// 30.4 -> ~15 ms), which is worth far more than the reds cost over the shared-memory group histogram.
template <int ORDER, bool REFCIC>
__global__ void __launch_bounds__(256) bucket_count_global_kernel(PaintParams p, TileGeom g,
                                                                  unsigned* __restrict__ counts,
                                                                  unsigned* __restrict__ wmax_bits) {
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  const int64_t nquads = (p.n_part + 3) >> 2;
  const bool aos = catalogue_is_aos(p);
  float wmax = p.w ? 0.0f : 1.0f;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += T) {
    const int64_t i = q << 2;
    const int cnt = (int)min((int64_t)4, p.n_part - i);
    Quad c;
    load_quad(p, aos, i, cnt, c);
    if (p.w) {
      float w[4];
      load_quad_weights(p, i, cnt, w);
      wmax = quad_wmax(w, cnt, wmax);
    }
    int tile[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      tile[u] = tile_of<ORDER, REFCIC>(grid_pos(c.x[u], p.xmin, p.inv), grid_pos(c.y[u], p.ymin, p.inv),
                                       grid_pos(c.z[u], p.zmin, p.inv), g, 0u);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < cnt) atomicAdd(counts + tile[u], 1u);
  }
  if (!p.w) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(wmax_bits, __float_as_uint(1.0f));
    return;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
  if ((threadIdx.x & 31) == 0) atomicMax(wmax_bits, __float_as_uint(wmax));
}

// ---------------------------------------------------------------- K1b: exclusive scan
// Three small launches: per-block totals (SCAN_BLOCK counters per CTA, coalesced), a one-CTA scan of
// the block totals, and the per-block rescan that writes offsets[] and the scatter cursors.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ROUNDS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ROUNDS;      // counters per CTA

__device__ __forceinline__ unsigned block_exclusive_scan_256(unsigned v, unsigned* warp_tot, unsigned& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  unsigned before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    const unsigned t = warp_tot[w];
    if (w < warp) before += t;
    tot += t;
  }
  __syncthreads();                                // warp_tot may be reused by the caller's next round
  total = tot;
  return before + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_totals_kernel(const unsigned* __restrict__ counts,
                                                                         unsigned* __restrict__ block_tot, int m) {
  __shared__ unsigned warp_tot[SCAN_THREADS / 32];
  const int base = blockIdx.x * SCAN_BLOCK;
  unsigned v = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ROUNDS; ++j) {
    const int i = base + j * SCAN_THREADS + threadIdx.x;
    if (i < m) v += counts[i];
  }
  unsigned total;
  block_exclusive_scan_256(v, warp_tot, total);
  if (threadIdx.x == 0) block_tot[blockIdx.x] = total;
}

// in place: block_tot[b] <- sum of block_tot[0..b-1]; block_tot[nblocks] <- grand total
__global__ void __launch_bounds__(SCAN_THREADS) scan_of_totals_kernel(unsigned* __restrict__ block_tot, int nblocks) {
  __shared__ unsigned warp_tot[SCAN_THREADS / 32];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const unsigned v = i < nblocks ? block_tot[i] : 0u;
    unsigned total;
    const unsigned ex = block_exclusive_scan_256(v, warp_tot, total);
    const unsigned c = carry;
    if (i < nblocks) block_tot[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_tot[nblocks] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_write_kernel(const unsigned* __restrict__ counts,
                                                                  const unsigned* __restrict__ block_off,
                                                                  unsigned* __restrict__ offsets,
                                                                  unsigned* __restrict__ cursor, int m, int nblocks) {
  __shared__ unsigned warp_tot[SCAN_THREADS / 32];
  const int base = blockIdx.x * SCAN_BLOCK;
  unsigned run = block_off[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ROUNDS; ++j) {
    const int i = base + j * SCAN_THREADS + threadIdx.x;
    const unsigned v = i < m ? counts[i] : 0u;
    unsigned total;
    const unsigned ex = block_exclusive_scan_256(v, warp_tot, total);
    if (i < m) { offsets[i] = run + ex; cursor[i] = run + ex; }
    run += total;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) offsets[m] = block_off[nblocks];
}

// ---------------------------------------------------------------- K1c: scatter into buckets
// slot = atomicAdd(cursor[tile], 1).  Slots of one tile are handed out in time order, so each
// bucket is filled at a moving frontier and the 16-byte records merge into full lines in L2
// before they reach DRAM (ncu: dram bytes written == 16 B/particle).  (Taking the slot from the
// count pass instead -- no second atomic -- was measured 1.8x SLOWER: it breaks that locality.)
template <int ORDER, bool REFCIC, int UNR>
__global__ void __launch_bounds__(256) bucket_scatter_kernel(PaintParams p, TileGeom g,
                                                             unsigned* __restrict__ cursor,
                                                             float4* __restrict__ sorted) {
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < p.n_part;
       i0 += UNR * T) {
    float4 rec[UNR];
    int tile[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t i = i0 + u * T;
      tile[u] = -1;
      if (i < p.n_part) {
        rec[u].x = grid_pos(p.x[i * p.stride], p.xmin, p.inv);
        rec[u].y = grid_pos(p.y[i * p.stride], p.ymin, p.inv);
        rec[u].z = grid_pos(p.z[i * p.stride], p.zmin, p.inv);
        rec[u].w = p.w ? p.w[i] : 1.0f;
        tile[u] = tile_of<ORDER, REFCIC>(rec[u].x, rec[u].y, rec[u].z, g, blockIdx.x);
      }
    }
    unsigned slot[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      slot[u] = (tile[u] >= 0) ? atomicAdd(cursor + tile[u], 1u) : 0u;
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (tile[u] >= 0) sorted[slot[u]] = rec[u];
  }
}

// ---------------------------------------------------------------- K1 v2: two-level partition
// The single-level scatter above pays one global atomic (with return) and one scattered 16-byte
// store per particle and sits at ~41 G particles/s whatever the unroll / occupancy / replica count
// (tools/sweep_bucket.sh).  Two-level partition, no per-particle global atomics:
//   A0 coarse_count    : per-CTA shared histogram over GROUPS of 2^gshift consecutive tiles
//   A1 coarse_scatter  : a CTA takes 8192 particles, ranks them per group with shared-memory integer
//                        atomics, reserves its run in every group with ONE global atomic per
//                        (CTA, group), stages the records in shared memory in group order and
//                        writes them out as coalesced runs
//   B  fine_scatter    : one CTA per group (its records are contiguous and L2-sized): shared
//                        histogram over the group's tiles -> tile offsets -> second pass places every
//                        record with a warp-aggregated SHARED cursor; writes go to <= 2^gshift
//                        frontiers per resident group, which L2 merges into full lines.
#ifndef JPS_COARSE_CHUNK
#define JPS_COARSE_CHUNK 4096
#define JPS_COARSE_THREADS 512
#define JPS_COARSE_MINB 2
#endif
#ifndef JPS_FINE_MINB
#define JPS_FINE_MINB 1
#endif
#ifndef JPS_FINE_UNR
#define JPS_FINE_UNR 4
#endif
#ifdef JPS_FINE_LDCS
#define JPS_FINE_LOAD(ptr) __ldcs(ptr)          // evict-first: the coarse records are read once
#else
#define JPS_FINE_LOAD(ptr) (*(ptr))
#endif
constexpr int COARSE_CHUNK = JPS_COARSE_CHUNK;     // 64 KB of staged records per CTA
constexpr int COARSE_THREADS = JPS_COARSE_THREADS;
constexpr int COARSE_QPT = COARSE_CHUNK / (4 * COARSE_THREADS);   // quads (4 particles) per thread
constexpr int kGroupSplit = 1 << 22;     // records: a group above this is placed by several CTAs (fine_heavy_kernel)
constexpr int kGroupSlice = 1 << 16;     // records per CTA of a heavy group
constexpr int kMaxGroups = 4096;         // capacity (12-bit group ids in the coarse pass); the partition uses max_groups() of them

// Groups of the coarse partition.  More groups = fewer tiles per group for the fine pass (which is latency bound on
// its scattered write frontiers) but shorter runs per (CTA, group) in the coarse pass.  JPS_MAX_GROUPS overrides (A/B).
static int max_groups() {
  static const int g = [] { const char* e = getenv("JPS_MAX_GROUPS"); const int v = e ? atoi(e) : 0; return (v >= 64 && v <= kMaxGroups) ? v : 2048; }();
  return g;
}

// exclusive scan of a[0..n) in shared memory, in place, by all threads of the CTA (n <= 4 * blockDim);
// returns the total.  `scratch` holds >= 33 words.  n <= 8 * blockDim.
__device__ __forceinline__ unsigned block_scan_inplace(unsigned* a, int n, unsigned* scratch) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
  const int ipt = (n + blockDim.x - 1) / blockDim.x;
  unsigned v[8];
  unsigned tsum = 0;
  for (int j = 0; j < ipt; ++j) {
    const int i = tid * ipt + j;
    v[j] = i < n ? a[i] : 0u;
    tsum += v[j];
  }
  unsigned incl = tsum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < nwarps ? scratch[lane] : 0u;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += t;
    }
    scratch[lane] = w;                             // inclusive scan of the warp totals
  }
  __syncthreads();
  unsigned run = (warp ? scratch[warp - 1] : 0u) + (incl - tsum);
  const unsigned total = scratch[nwarps - 1];
  for (int j = 0; j < ipt; ++j) {
    const int i = tid * ipt + j;
    if (i < n) a[i] = run;
    run += v[j];
  }
  __syncthreads();
  return total;
}

template <int ORDER, bool REFCIC>
__global__ void __launch_bounds__(256) coarse_count_kernel(PaintParams p, TileGeom g, int gshift, int ngroups,
                                                           unsigned* __restrict__ gcounts,
                                                           unsigned* __restrict__ wmax_bits) {
  extern __shared__ unsigned csm[];                // [ngroups]
  for (int i = threadIdx.x; i < ngroups; i += blockDim.x) csm[i] = 0u;
  __syncthreads();
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  const int64_t nquads = (p.n_part + 3) >> 2;
  const bool aos = catalogue_is_aos(p);
  float wmax = p.w ? 0.0f : 1.0f;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += T) {
    const int64_t i = q << 2;
    const int cnt = (int)min((int64_t)4, p.n_part - i);
    Quad c;
    load_quad(p, aos, i, cnt, c);
    if (p.w) {
      float w[4];
      load_quad_weights(p, i, cnt, w);
      wmax = quad_wmax(w, cnt, wmax);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int tile = tile_of<ORDER, REFCIC>(grid_pos(c.x[u], p.xmin, p.inv), grid_pos(c.y[u], p.ymin, p.inv),
                                              grid_pos(c.z[u], p.zmin, p.inv), g, 0u);
      if (u < cnt) atomicAdd(csm + (tile >> gshift), 1u);
    }
  }
  __syncthreads();
  const int rot = (int)(((long long)blockIdx.x * ngroups) / gridDim.x);
  for (int k = threadIdx.x; k < ngroups; k += blockDim.x) {
    int i = k + rot;
    if (i >= ngroups) i -= ngroups;
    const unsigned c = csm[i];
    if (c) atomicAdd(gcounts + i, c);
  }
  if (!p.w) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(wmax_bits, __float_as_uint(1.0f));
    return;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
  if ((threadIdx.x & 31) == 0) atomicMax(wmax_bits, __float_as_uint(wmax));
}

// gbase[0..ngroups] = exclusive scan of gcounts; gcursor = gbase (one CTA; ngroups <= kMaxGroups)
__global__ void __launch_bounds__(1024) group_scan_kernel(const unsigned* __restrict__ gcounts,
                                                          unsigned* __restrict__ gbase,
                                                          unsigned* __restrict__ gcursor, int ngroups) {
  __shared__ unsigned a[kMaxGroups];
  __shared__ unsigned scratch[33];
  for (int i = threadIdx.x; i < ngroups; i += blockDim.x) a[i] = gcounts[i];
  __syncthreads();
  const unsigned total = block_scan_inplace(a, ngroups, scratch);
  for (int i = threadIdx.x; i < ngroups; i += blockDim.x) { gbase[i] = a[i]; gcursor[i] = a[i]; }
  if (threadIdx.x == 0) gbase[ngroups] = total;
}

// group bases from per-tile offsets: gbase[g] = offsets[g << gshift], gbase[ngroups] = total
__global__ void group_bases_from_offsets_kernel(const unsigned* __restrict__ offsets, int nbuckets, int gshift,
                                                int ngroups, unsigned* __restrict__ gbase,
                                                unsigned* __restrict__ gcursor) {
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx > ngroups) return;
  const unsigned v = offsets[min(gidx << gshift, nbuckets)];
  gbase[gidx] = v;
  if (gidx < ngroups) gcursor[gidx] = v;
}

// TILE_COUNTS: also build the per-tile histogram (one global red per particle into the L2-resident table) -- the
// big-mesh path, where the pass that precedes this one then only needs the cheap shared-memory GROUP histogram.  The
// reds ride on a kernel that is bound by its shared-memory staging, not by the L2.
template <int ORDER, bool REFCIC, bool TILE_COUNTS>
__global__ void __launch_bounds__(COARSE_THREADS, JPS_COARSE_MINB) coarse_scatter_kernel(PaintParams p, TileGeom g, int gshift,
                                                                        int ngroups,
                                                                        unsigned* __restrict__ gcursor,
                                                                        float4* __restrict__ tmp,
                                                                        unsigned* __restrict__ tile_counts) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float4* stage = reinterpret_cast<float4*>(smraw);                                   // [COARSE_CHUNK]
  unsigned short* skey = reinterpret_cast<unsigned short*>(stage + COARSE_CHUNK);     // [COARSE_CHUNK]
  unsigned* h = reinterpret_cast<unsigned*>(skey + COARSE_CHUNK);                     // [ngroups] counts -> local offsets
  unsigned* gb = h + ngroups;                                                         // [ngroups] reserved global base
  unsigned* scratch = gb + ngroups;                                                   // [33]
  const int tid = threadIdx.x;
  const bool aos = catalogue_is_aos(p);
  const int64_t nchunks = (p.n_part + COARSE_CHUNK - 1) / COARSE_CHUNK;
  // Thread `tid` owns quads tid + k * COARSE_THREADS of the chunk (k < COARSE_QPT), held in registers
  // from the prefetch (issued before the PREVIOUS chunk's write-out, so the HBM latency hides behind
  // it) until they are staged.
  Quad cq[COARSE_QPT];
  float cw[COARSE_QPT][4];
  auto fetch = [&](int64_t ch) {
    const int64_t c0 = ch * COARSE_CHUNK;
    const int m = (int)min((int64_t)COARSE_CHUNK, p.n_part - c0);
#pragma unroll
    for (int k = 0; k < COARSE_QPT; ++k) {
      const int q4 = 4 * (tid + k * COARSE_THREADS);
      const int cnt = min(4, m - q4);
      if (cnt > 0) {
        load_quad(p, aos, c0 + q4, cnt, cq[k]);
        load_quad_weights(p, c0 + q4, cnt, cw[k]);
      }
    }
  };
  int64_t ch = blockIdx.x;
  if (ch < nchunks) fetch(ch);
  for (; ch < nchunks; ch += gridDim.x) {
    const int64_t c0 = ch * COARSE_CHUNK;
    const int m = (int)min((int64_t)COARSE_CHUNK, p.n_part - c0);
    for (int i = tid; i < ngroups; i += COARSE_THREADS) h[i] = 0u;
    __syncthreads();
    unsigned key[COARSE_QPT][4];                   // group | rank << 12
#pragma unroll
    for (int k = 0; k < COARSE_QPT; ++k) {
      const int cnt = m - 4 * (tid + k * COARSE_THREADS);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cq[k].x[u] = grid_pos(cq[k].x[u], p.xmin, p.inv);
        cq[k].y[u] = grid_pos(cq[k].y[u], p.ymin, p.inv);
        cq[k].z[u] = grid_pos(cq[k].z[u], p.zmin, p.inv);
        const int tile = tile_of<ORDER, REFCIC>(cq[k].x[u], cq[k].y[u], cq[k].z[u], g, 0u);
        const unsigned grp = (unsigned)(tile >> gshift);
        key[k][u] = 0xffffffffu;
        if (u < cnt) {
          key[k][u] = grp | (atomicAdd(h + grp, 1u) << 12);
          if (TILE_COUNTS) atomicAdd(tile_counts + tile, 1u);
        }
      }
    }
    __syncthreads();
    for (int k = tid; k < ngroups; k += COARSE_THREADS) {
      const unsigned c = h[k];
      gb[k] = c ? atomicAdd(gcursor + k, c) : 0u;  // this CTA's run inside group k
    }
    block_scan_inplace(h, ngroups, scratch);       // h -> local exclusive offsets (syncs inside)
#pragma unroll
    for (int k = 0; k < COARSE_QPT; ++k) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (key[k][u] != 0xffffffffu) {
          const unsigned grp = key[k][u] & 0xfffu, pos = h[grp] + (key[k][u] >> 12);
          stage[pos] = make_float4(cq[k].x[u], cq[k].y[u], cq[k].z[u], cw[k][u]);
          skey[pos] = (unsigned short)grp;
        }
      }
    }
    __syncthreads();
    if (ch + gridDim.x < nchunks) fetch(ch + gridDim.x);
    for (int j = tid; j < m; j += COARSE_THREADS) {
      const unsigned grp = skey[j];
      tmp[gb[grp] + ((unsigned)j - h[grp])] = stage[j];        // consecutive j of a group -> consecutive slots
    }
    __syncthreads();
  }
}

// HAVE_OFFSETS: the per-tile offsets already exist (small meshes: the shared-memory tile histogram
// of bucket_count_smem_kernel), so the group's records are read ONCE; otherwise pass 1 builds them.
template <int ORDER, bool REFCIC, bool HAVE_OFFSETS>
__global__ void __launch_bounds__(512, JPS_FINE_MINB) fine_scatter_kernel(const float4* __restrict__ tmp,
                                                           const unsigned* __restrict__ gbase, TileGeom g,
                                                           int gshift, int ngroups, int nbuckets,
                                                           unsigned* offsets,
                                                           float4* __restrict__ sorted, int skip_heavy) {
  extern __shared__ unsigned fsm[];                // [G] counts -> offsets, [G] cursors, [33] scratch
  const int G = 1 << gshift;
  unsigned* fh = fsm;
  unsigned* cur = fsm + G;
  unsigned* scratch = cur + G;
  const int grp = blockIdx.x;
  const unsigned beg = gbase[grp], end = gbase[grp + 1];
  if (skip_heavy && end - beg > (unsigned)kGroupSplit) return;          // left to fine_heavy_kernel
  const int t0 = grp << gshift;
  const int lane = threadIdx.x & 31;
  constexpr int FINE_UNR = JPS_FINE_UNR;           // records in flight per thread
  if (HAVE_OFFSETS) {
    for (int i = threadIdx.x; i < G; i += blockDim.x)
      cur[i] = (t0 + i < nbuckets) ? offsets[t0 + i] - beg : 0u;
    __syncthreads();
  } else {
  for (int i = threadIdx.x; i < G; i += blockDim.x) fh[i] = 0u;
  __syncthreads();
  // pass 1: histogram over the group's tiles (warp-aggregated: lanes with the same tile add once)
  for (unsigned i0 = beg; i0 < end; i0 += FINE_UNR * blockDim.x) {
    float4 r[FINE_UNR];
    bool ok[FINE_UNR];
#pragma unroll
    for (int u = 0; u < FINE_UNR; ++u) {
      const unsigned i = i0 + u * blockDim.x + threadIdx.x;
      ok[u] = i < end;
      if (ok[u]) r[u] = tmp[i];
    }
    // no early exit inside the unrolled body: the FINE_UNR match -> atomic chains stay independent
    // and overlap (a CTA-uniform `break` serialised them)
#pragma unroll
    for (int u = 0; u < FINE_UNR; ++u) {
      const int f = ok[u] ? tile_of<ORDER, REFCIC>(r[u].x, r[u].y, r[u].z, g, 0u) - t0 : -1;
      const unsigned same = __match_any_sync(0xffffffffu, f);
      if (f >= 0 && lane == __ffs(same) - 1) atomicAdd(fh + f, (unsigned)__popc(same));
    }
  }
  __syncthreads();
  block_scan_inplace(fh, G, scratch);              // fh -> exclusive offsets inside the group
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    cur[i] = fh[i];
    if (t0 + i < nbuckets) offsets[t0 + i] = beg + fh[i];
  }
  if (grp == ngroups - 1 && threadIdx.x == 0) offsets[nbuckets] = end;
  __syncthreads();
  }
  // pass 2: place
  for (unsigned i0 = beg; i0 < end; i0 += FINE_UNR * blockDim.x) {
    float4 r[FINE_UNR];
    bool ok[FINE_UNR];
#pragma unroll
    for (int u = 0; u < FINE_UNR; ++u) {
      const unsigned i = i0 + u * blockDim.x + threadIdx.x;
      ok[u] = i < end;
      if (ok[u]) r[u] = JPS_FINE_LOAD(tmp + i);
    }
    int f[FINE_UNR];
    unsigned same[FINE_UNR], base[FINE_UNR];
#pragma unroll
    for (int u = 0; u < FINE_UNR; ++u) {
      f[u] = ok[u] ? tile_of<ORDER, REFCIC>(r[u].x, r[u].y, r[u].z, g, 0u) - t0 : -1;
      same[u] = __match_any_sync(0xffffffffu, f[u]);
    }
#pragma unroll
    for (int u = 0; u < FINE_UNR; ++u) {
      base[u] = 0;
      if (f[u] >= 0 && lane == __ffs(same[u]) - 1) base[u] = atomicAdd(cur + f[u], (unsigned)__popc(same[u]));
    }
#pragma unroll
    for (int u = 0; u < FINE_UNR; ++u) {
      base[u] = __shfl_sync(0xffffffffu, base[u], __ffs(same[u]) - 1);
      if (f[u] >= 0) sorted[beg + base[u] + __popc(same[u] & ((1u << lane) - 1u))] = r[u];
    }
  }
}

// ---- fine pass, staged form (big groups).  The direct form above hands every record to its tile with one lone
// 16-byte store; with hundreds of tiles per group those stores have no neighbours in their warp and every 32-byte
// sector is written by two separate requests.  Here one CTA per group works like the coarse pass: a chunk of
// FS_CHUNK records is ranked by tile with shared integer atomics, laid out in tile order in shared memory and
// written out as coalesced runs (FS_CHUNK / tiles-per-group records each).  The CTA owns its group, so the tile
// cursors live in shared memory and no global atomic is needed.  The next chunk's loads are issued before the
// write-out of the current one.  Needs the per-tile offsets (HAVE_OFFSETS path).
// Two shapes (measured on a B200, 1e9 particles on 2048^3 = 2048 tiles per group: direct 26.5 ms, 4096-record chunks
// 20.4 ms, 8192-record chunks 12.0 ms; one rank of the 8-GPU decomposition = 256 tiles per group: direct 1.57 ms,
// 4096-record chunks 1.03 ms; C2 = 32 tiles per group: direct 0.755 ms, staged 0.778 ms): what matters is the run
// length chunk / tiles-per-group, so big groups take the big chunk (one 1024-thread CTA per SM) and medium ones the
// small chunk (two 512-thread CTAs per SM, which overlap each other's barriers).
constexpr int kFineStagedMinShift = 7;             // auto rule: staged form from 128 tiles per group
constexpr int kFineStagedBigShift = 10;            //            big chunk from 1024 tiles per group
constexpr bool kCountInCoarseDefault = false;      // auto rule for JPS_COUNT (until measured)

template <int CHUNK>
static size_t fine_staged_smem(int gshift) {
  return (size_t)CHUNK * (sizeof(float4) + 2) + (size_t)(3 * (1 << gshift) + 1) * 4 + 33 * 4;
}

template <int ORDER, bool REFCIC, int CHUNK, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fine_staged_kernel(const float4* __restrict__ tmp,
                                                                    const unsigned* __restrict__ gbase, TileGeom g,
                                                                    int gshift, int nbuckets,
                                                                    const unsigned* __restrict__ offsets,
                                                                    float4* __restrict__ sorted, int skip_heavy) {
  static_assert(CHUNK % THREADS == 0 && CHUNK <= 65536, "chunk ranks are kept in 16 bits");
  constexpr int RPT = CHUNK / THREADS;             // records per thread and chunk
  extern __shared__ __align__(16) unsigned char fs_raw[];
  const int G = 1 << gshift;
  float4* stage = reinterpret_cast<float4*>(fs_raw);                                  // [CHUNK] records in tile order
  unsigned short* skey = reinterpret_cast<unsigned short*>(stage + CHUNK);            // [CHUNK] their local tile
  unsigned* h = reinterpret_cast<unsigned*>(skey + CHUNK);                            // [G + 1] chunk histogram -> offsets
  unsigned* cur = h + G + 1;                                                          // [G] next free slot of every tile
  unsigned* wb = cur + G;                                                             // [G] cur - h: slot of staged record 0
  unsigned* scratch = wb + G;                                                         // [33]
  const int tid = threadIdx.x;
  const int grp = blockIdx.x;
  const unsigned beg = gbase[grp], end = gbase[grp + 1];
  if (beg == end) return;
  if (skip_heavy && end - beg > (unsigned)kGroupSplit) return;          // left to fine_heavy_kernel
  const int t0 = grp << gshift;
  for (int i = tid; i < G; i += THREADS) cur[i] = offsets[min(t0 + i, nbuckets)];
  float4 r[RPT];
  auto fetch = [&](unsigned c0) {
    const unsigned m = min((unsigned)CHUNK, end - c0);
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const unsigned j = (unsigned)tid + (unsigned)k * THREADS;
      if (j < m) r[k] = JPS_FINE_LOAD(tmp + c0 + j);
    }
  };
  fetch(beg);
  for (unsigned c0 = beg; c0 < end; c0 += CHUNK) {
    const unsigned m = min((unsigned)CHUNK, end - c0);
    for (int i = tid; i < G; i += THREADS) h[i] = 0u;
    if (tid == 0) h[G] = m;
    __syncthreads();
    unsigned key[RPT];                             // local tile | rank inside the chunk << 16
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const unsigned j = (unsigned)tid + (unsigned)k * THREADS;
      key[k] = 0xffffffffu;
      if (j < m) {
        const unsigned f = (unsigned)(tile_of<ORDER, REFCIC>(r[k].x, r[k].y, r[k].z, g, 0u) - t0);
        key[k] = f | (atomicAdd(h + f, 1u) << 16);
      }
    }
    __syncthreads();
    block_scan_inplace(h, G, scratch);             // h -> exclusive offsets inside the chunk (syncs inside)
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      if (key[k] != 0xffffffffu) {
        const unsigned f = key[k] & 0xffffu, pos = h[f] + (key[k] >> 16);
        stage[pos] = r[k];
        skey[pos] = (unsigned short)f;
      }
    }
    for (int i = tid; i < G; i += THREADS) {
      const unsigned c = cur[i], o = h[i];
      wb[i] = c - o;                               // mod 2^32: wb + (index in the chunk) is the record's slot
      cur[i] = c + (h[i + 1] - o);
    }
    __syncthreads();
    if (c0 + CHUNK < end) fetch(c0 + CHUNK);
    for (unsigned j = tid; j < m; j += THREADS) sorted[wb[skey[j]] + j] = stage[j];   // a tile's records: consecutive slots
    // no barrier here: the next round writes h only after its first barrier, stage / skey / wb after three more
  }
}

template <int ORDER, bool REFCIC, int CHUNK, int THREADS, int MINB>
static int launch_fine_staged(const float4* tmp, const unsigned* gbase, const TileGeom& g, int gshift, int ngroups, int nbuckets,
                              const unsigned* offsets, float4* sorted, int skip_heavy, cudaStream_t s) {
  JPS_REQUIRE(gshift <= 11, "fine_staged: 2^%d tiles per group do not fit the shared-memory tables", gshift);
  static PerDeviceFlag attr_set;
  if (!attr_set.get()) {
    JPS_CHECK_CUDA(cudaFuncSetAttribute(fine_staged_kernel<ORDER, REFCIC, CHUNK, THREADS, MINB>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fine_staged_smem<CHUNK>(11)));
    attr_set.set();
  }
  fine_staged_kernel<ORDER, REFCIC, CHUNK, THREADS, MINB><<<ngroups, THREADS, fine_staged_smem<CHUNK>(gshift), s>>>(
      tmp, gbase, g, gshift, nbuckets, offsets, sorted, skip_heavy);
  return JPS_OK;
}

// ---------------------------------------------------------------- K2: per-tile deposit
__device__ __forceinline__ void red_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// Per-axis anchor (lowest node, wrapped global index) and weights.
template <int ORDER>
__device__ __forceinline__ void tile_axis(float pos, int n, int wrap, int& anchor, float (&w)[ORDER]) {
  int idx[ORDER];
  bspline_axis<ORDER>(pos, n, 1, idx, w);        // wrapped indices; idx[0] is the anchor
  anchor = idx[0];
  if (!wrap) {                                   // non-periodic: drop nodes outside the mesh
    int base;
    if (ORDER == 2) base = (int)floorf(pos);
    else if (ORDER == 3) base = (int)floorf(pos + 0.5f) - 1;
    else base = (int)floorf(pos) - 1;
#pragma unroll
    for (int s = 0; s < ORDER; ++s)
      if (base + s < 0 || base + s >= n) w[s] = 0.0f;
  }
}

template <int ORDER, bool REFCIC>
__global__ void __launch_bounds__(256) paint_tile_kernel(const float4* __restrict__ sorted,
                                                         const unsigned* __restrict__ offsets,
                                                         TileGeom g, int wrap, int variant,
                                                         int mesh_vec_ok, float* __restrict__ mesh) {
  constexpr int L = TILE + ORDER - 1;            // local extent per axis (halo on the + side)
  constexpr int LP = (L + 3) & ~3;               // z pitch: rows stay 16-byte aligned for the flush
  __shared__ __align__(16) float tile[L * L * LP];
  const int t = blockIdx.x;
  const unsigned beg = offsets[t * g.rep], end = offsets[(t + 1) * g.rep];
  if (beg == end) return;                        // empty tile: nothing to flush
  const int tz = t % g.nt, ty = (t / g.nt) % g.nt, tx = t / (g.nt * g.nt);
  const int ox = tx * TILE, oy = ty * TILE, oz = tz * TILE;     // ox is a LOCAL plane index
  const int n = g.n;
  for (int i = threadIdx.x; i < L * L * LP; i += blockDim.x) tile[i] = 0.0f;
  __syncthreads();

  for (unsigned i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const float4 r = sorted[i];
    if (REFCIC) {
      int x0, x1, y0, y1, z0, z1;
      float mdx, ddx, mdy, ddy, mdz, ddz;
      cic_reference_axis(r.x, n, wrap, variant, x0, x1, mdx, ddx);
      cic_reference_axis(r.y, n, wrap, variant, y0, y1, mdy, ddy);
      cic_reference_axis(r.z, n, wrap, variant, z0, z1, mdz, ddz);
      // in-box particles: x1 == (x0+1) mod n for every variant, i.e. local index +1
      float* c = tile + ((local_plane(x0, g.x0, g.nx, n) - ox) * L + (y0 - oy)) * LP + (z0 - oz);
      const float wgt = r.w;
      constexpr int SX = L * LP, SY = LP;
      atomicAdd(c, ((mdx * mdy) * mdz) * wgt);
      atomicAdd(c + SX, ((ddx * mdy) * mdz) * wgt);
      atomicAdd(c + SY, ((mdx * ddy) * mdz) * wgt);
      atomicAdd(c + 1, ((mdx * mdy) * ddz) * wgt);
      atomicAdd(c + SX + SY, ((ddx * ddy) * mdz) * wgt);
      atomicAdd(c + SX + 1, ((ddx * mdy) * ddz) * wgt);
      atomicAdd(c + SY + 1, ((mdx * mdy) * ddz) * wgt);       // Q1 (reference weight)
      atomicAdd(c + SX + SY + 1, ((ddx * ddy) * ddz) * wgt);
    } else {
      int ax, ay, az;
      float wx[ORDER], wy[ORDER], wz[ORDER];
      tile_axis<ORDER>(r.x, n, wrap, ax, wx);
      tile_axis<ORDER>(r.y, n, wrap, ay, wy);
      tile_axis<ORDER>(r.z, n, wrap, az, wz);
      float* c = tile + ((local_plane(ax, g.x0, g.nx, n) - ox) * L + (ay - oy)) * LP + (az - oz);
#pragma unroll
      for (int a = 0; a < ORDER; ++a) {
#pragma unroll
        for (int b = 0; b < ORDER; ++b) {
          const float wxy = wx[a] * wy[b];
#pragma unroll
          for (int cc = 0; cc < ORDER; ++cc)
            atomicAdd(c + (a * L + b) * LP + cc, (wxy * wz[cc]) * r.w);
        }
      }
    }
  }
  __syncthreads();

  // flush: one (i,j) row of L floats at a time; 16-byte vector reds where the row segment is
  // aligned and does not wrap, scalar reds for the halo tail.
  const size_t n2 = (size_t)n * n;
  const bool vec_ok = (n % 4 == 0) && mesh_vec_ok;
  constexpr int NV = TILE / 4;                   // aligned float4 groups per row
  constexpr int ROW_ITEMS = NV + (ORDER - 1);    // + scalar halo cells
  for (int item = threadIdx.x; item < L * L * ROW_ITEMS; item += blockDim.x) {
    const int q = item % ROW_ITEMS;
    const int row = item / ROW_ITEMS;
    const int j = row % L, i = row / L;
    int gx = ox + i;                               // local plane
    if (g.nx == n) gx %= n;                        // full mesh: periodic in x
    else if (gx >= g.nx) continue;                 // slab: ghost planes are part of the allocation
    const int gy = (oy + j) % n;
    float* grow = mesh + (size_t)gx * n2 + (size_t)gy * n;
    const float* trow = tile + (i * L + j) * LP;
    if (q < NV) {
      const int k = q * 4;
      const float4 v = *reinterpret_cast<const float4*>(trow + k) ;
      if (v.x == 0.0f && v.y == 0.0f && v.z == 0.0f && v.w == 0.0f) continue;
      const int gz = oz + k;
      if (vec_ok && gz + 3 < n) {
        red_v4(grow + gz, v.x, v.y, v.z, v.w);
      } else {
        if (v.x != 0.0f) atomicAdd(grow + (gz % n), v.x);
        if (v.y != 0.0f) atomicAdd(grow + ((gz + 1) % n), v.y);
        if (v.z != 0.0f) atomicAdd(grow + ((gz + 2) % n), v.z);
        if (v.w != 0.0f) atomicAdd(grow + ((gz + 3) % n), v.w);
      }
    } else {
      const int k = TILE + (q - NV);
      const float v = trow[k];
      if (v != 0.0f) atomicAdd(grow + ((oz + k) % n), v);
    }
  }
}

// ---------------------------------------------------------------- K2 (v3): fixed-point deposit
// Shared-memory float atomicAdd is a compare-and-swap loop on sm_100a (ATOMS.CAST.SPIN); the only
// native shared atomic add is the 32-bit integer one (ATOMS.ADD), measured 2x faster and immune to
// same-address retries.  So the tile is accumulated in 64-bit fixed point held as two 32-bit words
// per cell: lo += q (native atomic, returns the old value -> carry), hi += carry (native, only on
// overflow of the low word: ~4% of the updates).  Each contribution is the SAME float32 product the
// reference forms, scaled by a power of two 2^(fx_bits-e) with 2^e >= max|w| (exact in float32; fx_bits = 31 or 28,
// below) and rounded to an integer, i.e. quantised at 2^-fx_bits of the largest weight -- far below float32
// resolution -- and integer addition is associative, so the tile sum is independent of the order
// in which particles arrive (the float32 path is only reproducible to ~1e-7).
// A cleverer ordering of the particles (in-tile counting sort by cell, tried in round 1) was
// SLOWER with float CAS: adjacent lanes then hit the same cell and every collision costs a full
// CAS retry; with native integer atomics ordering no longer matters.
// Heavy tiles.  One CTA deposits at most kTileSplit particles of a tile; a tile that holds more is finished by further
// CTAs (a second launch over the list of (tile, part) pairs built here), each with its own shared-memory copy of the
// tile -- the flush is additive, so nothing else changes.  Measured on a B200 (1e8 particles, TSC, 512^3,
// tools/clustered_paint.py, before / after): half of the particles in 40 blobs (2.3e5 in the densest tile, 74x the
// mean) 2.15 ms against 1.35 ms for a uniform catalogue; half in 4 blobs of 1 % of the box (7.6e6 in one tile)
// 31.4 ms -- one CTA working through millions of particles while the GPU idles.
constexpr int kTileSplit = 16384;

__global__ void __launch_bounds__(256) tile_parts_kernel(const unsigned* __restrict__ offsets, int ntiles, int rep,
                                                         uint2* __restrict__ parts, unsigned* __restrict__ nparts, int cap) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const unsigned c = offsets[(t + 1) * rep] - offsets[t * rep];
  if (c <= (unsigned)kTileSplit) return;
  const unsigned k = (c - 1u) / (unsigned)kTileSplit;         // parts beyond the first
  const unsigned base = atomicAdd(nparts, k);
  for (unsigned i = 0; i < k && base + i < (unsigned)cap; ++i) parts[base + i] = make_uint2((unsigned)t, i + 1u);
}

// Heavy groups.  The fine pass gives a group of tiles to ONE CTA (its tile cursors live in shared memory); a group that
// holds more than kGroupSplit records -- a catalogue with millions of particles in one tile -- is left out by that CTA
// and cut into slices of kGroupSlice records, each placed by a CTA of its own through the GLOBAL tile cursors the scan
// left behind (one warp-aggregated global atomic per tile and warp).  (Same catalogue as above: fine pass 7.25 ms with
// 7.6e6 records in one group against 0.76 ms for a uniform catalogue.)
__global__ void __launch_bounds__(256) group_slices_kernel(const unsigned* __restrict__ gbase, int ngroups,
                                                           uint2* __restrict__ slices, unsigned* __restrict__ nslices, int cap) {
  const int grp = blockIdx.x * blockDim.x + threadIdx.x;
  if (grp >= ngroups) return;
  const unsigned c = gbase[grp + 1] - gbase[grp];
  if (c <= (unsigned)kGroupSplit) return;
  const unsigned k = (c + (unsigned)kGroupSlice - 1u) / (unsigned)kGroupSlice;
  const unsigned base = atomicAdd(nslices, k);
  for (unsigned i = 0; i < k && base + i < (unsigned)cap; ++i) slices[base + i] = make_uint2((unsigned)grp, i);
}

template <int ORDER, bool REFCIC>
__global__ void __launch_bounds__(512) fine_heavy_kernel(const float4* __restrict__ tmp, const unsigned* __restrict__ gbase,
                                                         TileGeom g, const uint2* __restrict__ slices,
                                                         const unsigned* __restrict__ nslices,
                                                         unsigned* __restrict__ cursor, float4* __restrict__ sorted) {
  if (blockIdx.x >= *nslices) return;
  const uint2 e = slices[blockIdx.x];
  const unsigned gend = gbase[e.x + 1];
  const unsigned beg = gbase[e.x] + e.y * (unsigned)kGroupSlice;
  if (beg >= gend) return;
  const unsigned end = (gend - beg > (unsigned)kGroupSlice) ? beg + (unsigned)kGroupSlice : gend;
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned i0 = beg; i0 < end; i0 += blockDim.x) {         // CTA-uniform trip count: every warp stays whole
    const unsigned i = i0 + threadIdx.x;
    const bool ok = i < end;
    float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    int tile = -1 - (int)lane;                                   // idle lanes match nobody
    if (ok) {
      r = tmp[i];
      tile = tile_of<ORDER, REFCIC>(r.x, r.y, r.z, g, 0u);
    }
    const unsigned same = __match_any_sync(0xffffffffu, tile);
    const int leader = __ffs(same) - 1;
    unsigned base = 0u;
    if (ok && (int)lane == leader) base = atomicAdd(cursor + tile, (unsigned)__popc(same));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ok) sorted[base + (unsigned)__popc(same & ((1u << lane) - 1u))] = r;
  }
}

// Fixed-point position of the tile accumulators (kernel argument fx_bits = 31; JPS_FX_BITS=24..31 overrides): a
// contribution is quantised at 2^-fx_bits of the tile's largest weight.  At 31 |contribution| fills the low word and,
// with PCS, 0.8 % of the updates carry into the high word (two thirds of the warp-level rows of four take the slow
// path); 28 makes carries 8x rarer at a quantum of 3.7e-9 -- measured: no change in the kernel's time (5.25 ms per C4
// rank either way: it is bound by the shared-memory atomic pipe, not by instruction issue), so the finest position stays.
template <int ORDER>
struct TileDims {
  static constexpr int L = TILE + ORDER - 1;
  static constexpr int LP = (L + 3) & ~3;
  static constexpr int CELLS = L * L * LP;
};

// v >= 0 is the magnitude of one contribution in units of 2^-fx_bits of the weight scale (< 2^32, one
// F2I); NEG selects subtraction (negative particle weight).  64-bit two's-complement arithmetic on
// the (hi, lo) word pair: the low-word atomic returns the old value, from which carry / borrow follow.
// K contributions at once: all K low-word atomics are issued back to back (independent, K in
// flight), and the K carry tests share ONE rarely taken branch.  (One atomic + test + branch per
// contribution serialised every update behind the previous atomic's return and cost 9 instructions
// per update instead of 6.)
template <bool NEG, int K>
__device__ __forceinline__ void fx_add_batch(unsigned* __restrict__ lo, unsigned* __restrict__ hi,
                                             const int (&idx)[K], const float (&v)[K]) {
  unsigned q[K], old[K];
#pragma unroll
  for (int j = 0; j < K; ++j) q[j] = __float2uint_rn(v[j]);     // q == 0: no carry, no borrow
#pragma unroll
  for (int j = 0; j < K; ++j) old[j] = atomicAdd(lo + idx[j], NEG ? 0u - q[j] : q[j]);
  // carry of word j <=> q > ~old; borrow <=> old < q.  The slow path re-derives the flags from the
  // wrapped sum so that the compiler keeps the fast path to K compares (it otherwise materialises
  // both senses of every predicate before the branch).
  bool any = false;
  if (NEG) {
#pragma unroll
    for (int j = 0; j < K; ++j) any |= (old[j] < q[j]);
  } else {
    // number of carries as the high words of 64-bit sums: one add with carry-out + one add-with-carry per update
    // (IADD3 / IADD3.X) instead of a complement and two compares
    // (inline PTX: from the C expression the compiler falls back to compares and selects)
    unsigned nc = 0u;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      unsigned t;
      asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %1, 0;" : "=r"(t), "+r"(nc) : "r"(old[j]), "r"(q[j]));
    }
    any = nc != 0u;
  }
  if (any) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const unsigned sum = NEG ? old[j] - q[j] : old[j] + q[j];
      if (NEG ? (sum > old[j]) : (sum < old[j]))
        atomicAdd(hi + idx[j], NEG ? 0xffffffffu : 1u);          // borrow / carry into the high word
    }
  }
}

// One axis of a particle that is KNOWN to belong to the tile whose local cell 0 is the global node
// `origin` (the bucketing put it there): the B-spline weights exactly as bspline_axis() forms them, and
// the anchor as a TILE-LOCAL index in [0, TILE).  Because the answer is known to lie in the tile, the
// periodic wrap is one compare (taken only by the few particles whose stencil straddles the box edge)
// instead of ORDER general wraps per axis: the deposit's preamble shrinks from ~170 to ~60 instructions.
template <int ORDER>
__device__ __forceinline__ int tile_local_axis(float pos, int n, int origin, int wrap, float (&w)[ORDER]) {
  int base;
  if (ORDER == 2) {
    const float f = floorf(pos);
    const float d = pos - f;
    w[0] = 1.0f - d;
    w[1] = d;
    base = (int)f;
  } else if (ORDER == 3) {
    const float f = floorf(pos + 0.5f);
    const float d = pos - f;
    const float a = 0.5f - d, b = 0.5f + d;
    w[0] = 0.5f * a * a;
    w[1] = 0.75f - d * d;
    w[2] = 0.5f * b * b;
    base = (int)f - 1;
  } else {
    const float f = floorf(pos);
    const float d = pos - f;
    const float e = 1.0f - d;
    const float sixth = 1.0f / 6.0f;
    w[0] = e * e * e * sixth;
    w[1] = (4.0f - 6.0f * d * d + 3.0f * d * d * d) * sixth;
    w[2] = (4.0f - 6.0f * e * e + 3.0f * e * e * e) * sixth;
    w[ORDER - 1] = d * d * d * sixth;
    base = (int)f - 1;
  }
  if (!wrap) {                                   // non-periodic: drop nodes outside the mesh
#pragma unroll
    for (int s = 0; s < ORDER; ++s)
      if (base + s < 0 || base + s >= n) w[s] = 0.0f;
  }
  int l = base - origin;
  if ((unsigned)l >= (unsigned)TILE) l = pymod(l, n);
  return l;
}

// The anchor alone (same arithmetic as tile_local_axis): the cell a particle's stencil starts in decides which
// shared-memory bank its atomics hit.
template <int ORDER>
__device__ __forceinline__ int tile_local_anchor(float pos, int n, int origin) {
  int l = anchor_base<ORDER>(pos) - origin;
  if ((unsigned)l >= (unsigned)TILE) l = pymod(l, n);
  return l;
}

template <int ORDER, bool REFCIC, bool NEG>
__device__ __forceinline__ void fx_deposit(const float4& r, float ws, unsigned* lo, unsigned* hi,
                                           const TileGeom& g, int wrap, int variant, int ox, int oy, int oz) {
  constexpr int L = TileDims<ORDER>::L, LP = TileDims<ORDER>::LP;
  const int n = g.n;
  // ws = |w| * 2^(fx_bits-e): the power-of-two scale is folded into the weight, which commutes exactly
  // with the float32 products below, so each contribution is the reference's float32 product
  // (src/mas.py:142-151, left to right) times 2^(fx_bits-e).
  if (REFCIC) {
    int x0, x1, y0, y1, z0, z1;
    float mdx, ddx, mdy, ddy, mdz, ddz;
    cic_reference_axis(r.x, n, wrap, variant, x0, x1, mdx, ddx);
    cic_reference_axis(r.y, n, wrap, variant, y0, y1, mdy, ddy);
    cic_reference_axis(r.z, n, wrap, variant, z0, z1, mdz, ddz);
    const int c = ((local_plane(x0, g.x0, g.nx, n) - ox) * L + (y0 - oy)) * LP + (z0 - oz);
    constexpr int SX = L * LP, SY = LP;
    const int ia[4] = {c, c + SX, c + SY, c + 1};
    const float va[4] = {((mdx * mdy) * mdz) * ws, ((ddx * mdy) * mdz) * ws, ((mdx * ddy) * mdz) * ws,
                         ((mdx * mdy) * ddz) * ws};
    fx_add_batch<NEG, 4>(lo, hi, ia, va);
    const int ib[4] = {c + SX + SY, c + SX + 1, c + SY + 1, c + SX + SY + 1};
    const float vb[4] = {((ddx * ddy) * mdz) * ws, ((ddx * mdy) * ddz) * ws,
                         ((mdx * mdy) * ddz) * ws,                       // Q1 (reference weight)
                         ((ddx * ddy) * ddz) * ws};
    fx_add_batch<NEG, 4>(lo, hi, ib, vb);
  } else {
    float wx[ORDER], wy[ORDER], wz[ORDER];
    const int lx = tile_local_axis<ORDER>(r.x, n, g.x0 + ox, wrap, wx);   // ox is a local plane: global node g.x0 + ox
    const int ly = tile_local_axis<ORDER>(r.y, n, oy, wrap, wy);
    const int lz = tile_local_axis<ORDER>(r.z, n, oz, wrap, wz);
    const int c = (lx * L + ly) * LP + lz;
#pragma unroll
    for (int cc = 0; cc < ORDER; ++cc) wz[cc] *= ws;   // one multiply per update instead of two
#pragma unroll
    for (int a = 0; a < ORDER; ++a) {
#pragma unroll
      for (int b = 0; b < ORDER; ++b) {
        const float wxy = wx[a] * wy[b];
        int idx[ORDER];
        float v[ORDER];
#pragma unroll
        for (int cc = 0; cc < ORDER; ++cc) {
          idx[cc] = c + (a * L + b) * LP + cc;
          v[cc] = wxy * wz[cc];
        }
        fx_add_batch<NEG, ORDER>(lo, hi, idx, v);     // one z-row of the stencil
      }
    }
  }
}

// ---- tile flush through the TMA: cp.reduce.async.bulk.tensor (.add.f32), shared -> global
// After the deposit the fixed-point tile is converted IN PLACE to float32 (the low-word array becomes a
// dense [L][L][LP] float box).  One elected thread then hands the whole box to the TMA unit, which adds
// it to the mesh in L2 -- one instruction per tile instead of L*L*5 per-thread red.global.add.v4 (1 805
// for PCS), and no per-thread address arithmetic.  Box elements outside the tensor (the ghost planes a
// slab does not hold) are clipped by the hardware.  Tiles whose box would have to WRAP around the periodic
// box (last tile of an axis), meshes with N % 4 != 0 and unaligned meshes take the per-thread red path.
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1, int c2) {
  const unsigned src = (unsigned)__cvta_generic_to_shared(smem_src);
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tmap), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the box has been read: smem may be released
}

// One particle straight into the global mesh (float atomics, the general per-particle arithmetic of
// paint_atomic_kernel): used by the tile kernel for the rare particles it must not put into the fixed-point
// tile (non-finite weights).
template <int ORDER, bool REFCIC>
__device__ __noinline__ void global_deposit(float rx, float ry, float rz, float rw, int n, int gx0, int gnx, int wrap,
                                            int variant, float* __restrict__ mesh) {
  const float4 r = make_float4(rx, ry, rz, rw);
  struct { int x0, nx; } g = {gx0, gnx};
  const size_t n2 = (size_t)n * n;
  if (REFCIC) {
    int x0, x1, y0, y1, z0, z1;
    float mdx, ddx, mdy, ddy, mdz, ddz;
    cic_reference_axis(r.x, n, wrap, variant, x0, x1, mdx, ddx);
    cic_reference_axis(r.y, n, wrap, variant, y0, y1, mdy, ddy);
    cic_reference_axis(r.z, n, wrap, variant, z0, z1, mdz, ddz);
    x0 = local_plane(x0, g.x0, g.nx, n);
    x1 = local_plane(x1, g.x0, g.nx, n);
#define JPS_CORNER(ix, iy, iz, wx, wy, wz)                                           \
  if (((ix) | (iy) | (iz)) >= 0)                                                      \
    atomicAdd(mesh + (size_t)(ix) * n2 + (size_t)(iy) * n + (iz), (((wx) * (wy)) * (wz)) * r.w);
    JPS_CORNER(x0, y0, z0, mdx, mdy, mdz)
    JPS_CORNER(x1, y0, z0, ddx, mdy, mdz)
    JPS_CORNER(x0, y1, z0, mdx, ddy, mdz)
    JPS_CORNER(x0, y0, z1, mdx, mdy, ddz)
    JPS_CORNER(x1, y1, z0, ddx, ddy, mdz)
    JPS_CORNER(x1, y0, z1, ddx, mdy, ddz)
    JPS_CORNER(x0, y1, z1, mdx, mdy, ddz)
    JPS_CORNER(x1, y1, z1, ddx, ddy, ddz)
#undef JPS_CORNER
  } else {
    int ix[ORDER], iy[ORDER], iz[ORDER];
    float wx[ORDER], wy[ORDER], wz[ORDER];
    bspline_axis<ORDER>(r.x, n, wrap, ix, wx);
    bspline_axis<ORDER>(r.y, n, wrap, iy, wy);
    bspline_axis<ORDER>(r.z, n, wrap, iz, wz);
#pragma unroll
    for (int a = 0; a < ORDER; ++a) {
      const int lx = local_plane(ix[a], g.x0, g.nx, n);
#pragma unroll
      for (int b = 0; b < ORDER; ++b) {
        if ((lx | iy[b]) < 0) continue;
        float* row = mesh + (size_t)lx * n2 + (size_t)iy[b] * n;
        const float wxy = wx[a] * wy[b];
#pragma unroll
        for (int c = 0; c < ORDER; ++c)
          if (iz[c] >= 0) atomicAdd(row + iz[c], (wxy * wz[c]) * r.w);
      }
    }
  }
}

template <int ORDER, bool REFCIC>
__global__ void __launch_bounds__(512, 3) paint_tile_fx_kernel(const float4* __restrict__ sorted,
                                                            const unsigned* __restrict__ offsets,
                                                            TileGeom g, int wrap, int variant,
                                                            int mesh_vec_ok, int has_w,
                                                            float* __restrict__ mesh,
                                                            const __grid_constant__ CUtensorMap tmap, int use_tma,
                                                            int tile_offset, int bank_order, int fx_bits,
                                                            const uint2* __restrict__ parts,
                                                            const unsigned* __restrict__ nparts, int tile_count) {
  constexpr int L = TileDims<ORDER>::L, LP = TileDims<ORDER>::LP, NC = TileDims<ORDER>::CELLS;
  extern __shared__ __align__(128) unsigned fx_smem[];
  unsigned* lo = fx_smem;
  unsigned* hi = fx_smem + NC;
  // first launch: CTA b takes the first kTileSplit particles of tile b; second launch (parts != nullptr): CTA b takes
  // entry b of the list of further parts of heavy tiles
  int t;
  unsigned beg, end;
  if (parts == nullptr) {
    t = blockIdx.x + tile_offset;
    beg = offsets[t * g.rep];
    end = offsets[(t + 1) * g.rep];
  } else {
    if (blockIdx.x >= *nparts) return;
    const uint2 e = parts[blockIdx.x];
    t = (int)e.x;
    if (t < tile_offset || t >= tile_offset + tile_count) return;      // deposit of a range of tile rows
    beg = offsets[t * g.rep] + e.y * (unsigned)kTileSplit;        // < offsets[(t + 1) * rep]: the list holds existing parts only
    end = offsets[(t + 1) * g.rep];
  }
  if (beg >= end) return;
  if (end - beg > (unsigned)kTileSplit) end = beg + (unsigned)kTileSplit;
  const int tz = t % g.nt, ty = (t / g.nt) % g.nt, tx = t / (g.nt * g.nt);
  const int ox = tx * TILE, oy = ty * TILE, oz = tz * TILE;
  const int n = g.n;
  static_assert((2 * NC) % 4 == 0, "tile words are zeroed 16 bytes at a time");
  for (int i = threadIdx.x; i < (2 * NC) / 4; i += blockDim.x) reinterpret_cast<uint4*>(fx_smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  // Power-of-two scale of THIS tile: 2^e >= max|w| over the tile's own particles, so that
  // |contribution| * 2^(fx_bits-e) <= 2^31 fits one 32-bit word and the quantum is 2^-fx_bits of the tile's largest
  // weight -- one outlier weight costs precision in its own 16^3 cells only, not in the whole mesh.  With unit
  // weights (no w array) the scale is 2^31 and the extra pass over the tile's records is skipped.  Particles
  // with a non-finite weight never enter the fixed-point tile: they are deposited straight into the mesh with
  // float atomics, so inf / NaN propagate as they do in the reference's float32 scatter.
  __shared__ float s_wmax[32];
  float wmax = 1.0f;
  if (has_w) {
    float m = 0.0f;
    for (unsigned i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const float a = fabsf(sorted[i].w);
      if (a < 3.0e38f) m = fmaxf(m, a);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) s_wmax[threadIdx.x >> 5] = m;
    __syncthreads();
    m = 0.0f;
    for (int q = 0; q < (int)((blockDim.x + 31) >> 5); ++q) m = fmaxf(m, s_wmax[q]);
    wmax = m > 0.0f ? m : 1.0f;
  }
  int e;
  frexpf(wmax, &e);
  e = max(-90, min(90, e));
  const float scale = ldexpf(1.0f, fx_bits - e);
  __syncthreads();

  if (!REFCIC && bank_order) {
    // Bank-class order.  Every atomic of a warp instruction adds the SAME stencil offset to each lane's anchor cell, so
    // the bank conflicts of all order^3 instructions are decided by the 32 anchors' word index mod 32 -- with particles
    // in arrival order that is 32 balls in 32 bins (3.5 wavefronts per instruction: 54 % of the kernel's shared-memory
    // wavefronts were conflict replays).  So each batch of blockDim particles is counting-sorted by that residue, and
    // warp w takes the sorted positions w, w + nw, w + 2 nw, ... (nw = warps in the batch): its lanes walk through the
    // classes and land on distinct banks but for the Poisson drift of the class sizes (2.0 wavefronts per instruction on
    // C4's tiles, 1.6 on C2's), with every lane busy.  MEASURED AND NOT THE DEFAULT (JPS_TILE_ORDER=bank enables it):
    // on one C4 rank (ncu) the conflict replays fall from 6.7e8 to 3.2e8 and the shared-memory wavefronts from 12.4e8
    // to 9.1e8 (l1tex 85 % -> 64 % busy) exactly as the model says -- tools/microbench_atoms.cu: an ATOMS costs its
    // conflict degree + 0.3 clocks, 1.33 with 32 distinct banks, 3.84 with random words -- but the sort adds 10 % to the
    // instruction count and four barriers per batch, and the kernel, already issuing 69 % of the time, ends up bound by
    // issue and barriers instead: PCS 5.25 -> 5.49 ms per C4 rank, TSC on C2 1.40 -> 1.89 ms, CIC 0.38 -> 0.44 ms.
    __shared__ unsigned s_cls[32], s_start[32];
    __shared__ unsigned short s_perm[512];                                 // [warp][lane] -> particle of the batch
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (unsigned b0 = beg; b0 < end; b0 += blockDim.x) {
      const unsigned pb = min((unsigned)blockDim.x, end - b0);
      const unsigned nwb = (pb + 31u) >> 5;
      if (threadIdx.x < 32) s_cls[threadIdx.x] = 0u;
      __syncthreads();
      const bool valid = threadIdx.x < pb;
      float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      unsigned cls = 0xffffu, rank = 0u;
      if (valid) {
        r = sorted[b0 + threadIdx.x];
        const int lx = tile_local_anchor<ORDER>(r.x, n, g.x0 + ox), ly = tile_local_anchor<ORDER>(r.y, n, oy),
                  lz = tile_local_anchor<ORDER>(r.z, n, oz);
        cls = (unsigned)((lx * L + ly) * LP + lz) & 31u;
      }
      {
        const unsigned same = __match_any_sync(0xffffffffu, cls);
        const int leader = __ffs(same) - 1;
        unsigned base = 0u;
        if (valid && (int)lane == leader) base = atomicAdd(&s_cls[cls], (unsigned)__popc(same));
        base = __shfl_sync(0xffffffffu, base, leader);
        rank = base + (unsigned)__popc(same & ((1u << lane) - 1u));
      }
      __syncthreads();
      if (warp == 0) {                             // exclusive scan of the 32 class sizes
        const unsigned c = s_cls[lane];
        unsigned incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
          if ((int)lane >= off) incl += t;
        }
        s_start[lane] = incl - c;
      }
      __syncthreads();
      if (valid) {
        const unsigned p = s_start[cls] + rank;    // sorted position -> (warp p % nwb, lane p / nwb)
        const unsigned l = (p * (65536u / nwb + 1u)) >> 16;                // p / nwb, exact for p < 512, nwb <= 16
        s_perm[(p - l * nwb) * 32u + l] = (unsigned short)threadIdx.x;
      }
      __syncthreads();
      if (warp < nwb && warp + nwb * lane < pb) {
        r = sorted[b0 + s_perm[threadIdx.x]];      // second read of the batch's records: L1 / L2 hits
        const float ws = fabsf(r.w) * scale;
        if (has_w && !(fabsf(r.w) < 3.0e38f)) global_deposit<ORDER, REFCIC>(r.x, r.y, r.z, r.w, n, g.x0, g.nx, wrap, variant, mesh);
        else if (r.w >= 0.0f) fx_deposit<ORDER, REFCIC, false>(r, ws, lo, hi, g, wrap, variant, ox, oy, oz);
        else fx_deposit<ORDER, REFCIC, true>(r, ws, lo, hi, g, wrap, variant, ox, oy, oz);
      }
      // no barrier: the next batch touches s_cls after its own first barrier and s_perm after three
    }
  } else {
    for (unsigned i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const float4 r = sorted[i];
      const float ws = fabsf(r.w) * scale;
      if (has_w && !(fabsf(r.w) < 3.0e38f)) global_deposit<ORDER, REFCIC>(r.x, r.y, r.z, r.w, n, g.x0, g.nx, wrap, variant, mesh);
      else if (r.w >= 0.0f) fx_deposit<ORDER, REFCIC, false>(r, ws, lo, hi, g, wrap, variant, ox, oy, oz);
      else fx_deposit<ORDER, REFCIC, true>(r, ws, lo, hi, g, wrap, variant, ox, oy, oz);
    }
  }
  __syncthreads();

  // fixed point -> float32 (one rounding), in place over the low words: lo[] becomes a dense float box
  // A cell whose sum fits one word (high word 0: every cell of a sparse tile) takes one I2F and an exact multiply by the
  // power of two -- the same value as the general 64-bit -> double -> float path, at a quarter of its instructions.
  const float inv_scale_f = ldexpf(1.0f, e - fx_bits);
  const double inv_scale = (double)inv_scale_f;
  float* ftile = reinterpret_cast<float*>(lo);
  for (int i = threadIdx.x; i < NC; i += blockDim.x) {
    const unsigned l = lo[i], h = hi[i];
    float v;
    if (h == 0u) v = __uint2float_rn(l) * inv_scale_f;
    else v = (float)((double)(long long)(((unsigned long long)h << 32) | l) * inv_scale);
    ftile[i] = v;
  }
  // whole box inside the mesh along y, z (and x for a full mesh; a slab's missing planes are clipped)?
  const bool inside = (oz + LP <= n) && (oy + L <= n) && (g.nx != n || ox + L <= n);
  if (use_tma && inside) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the TMA
    __syncthreads();
    if (threadIdx.x == 0) tma_reduce_add_3d(&tmap, ftile, oz, oy, ox);
    return;
  }
  __syncthreads();

  // per-thread flush: 16-byte vector reds where aligned (boundary tiles and meshes the TMA cannot take)
  const size_t n2 = (size_t)n * n;
  const bool vec_ok = (n % 4 == 0) && mesh_vec_ok;
  constexpr int NV = TILE / 4;
  // per row: NV aligned quads + ONE more quad for the halo tail [TILE, L) -- its padding cells
  // [L, LP) are zero in the tile, and adding +0.0 to the mesh is free, so the tail costs one vector
  // red instead of order-1 scalar ones
  constexpr int ROW_ITEMS = NV + 1;
  static_assert(LP == TILE + 4, "tail quad = cells [TILE, TILE + 4)");
  for (int item = threadIdx.x; item < L * L * ROW_ITEMS; item += blockDim.x) {
    const int q = item % ROW_ITEMS;
    const int row = item / ROW_ITEMS;
    const int j = row % L, i = row / L;
    int gx = ox + i;                               // local plane
    if (g.nx == n) gx %= n;                        // full mesh: periodic in x
    else if (gx >= g.nx) continue;                 // slab: ghost planes are part of the allocation
    const int gy = (oy + j) % n;
    float* grow = mesh + (size_t)gx * n2 + (size_t)gy * n;
    const int k = q * 4;
    const float4 v = *reinterpret_cast<const float4*>(ftile + (i * L + j) * LP + k);
    if (v.x == 0.0f && v.y == 0.0f && v.z == 0.0f && v.w == 0.0f) continue;
    const int gz = (oz + k) % n;
    if (vec_ok && gz + 3 < n) {
      red_v4(grow + gz, v.x, v.y, v.z, v.w);
    } else {
      if (v.x != 0.0f) atomicAdd(grow + gz, v.x);
      if (v.y != 0.0f) atomicAdd(grow + ((gz + 1) % n), v.y);
      if (v.z != 0.0f) atomicAdd(grow + ((gz + 2) % n), v.z);
      if (v.w != 0.0f) atomicAdd(grow + ((gz + 3) % n), v.w);
    }
  }
}

// Particles the tile kernel cannot take (reference-compat CIC outside the box): the last bucket,
// painted with the general per-particle path straight from the bucketed records.
__global__ void __launch_bounds__(256) paint_outliers_kernel(const float4* __restrict__ sorted,
                                                             const unsigned* __restrict__ offsets,
                                                             int bucket, int n, int gx0, int nx, int wrap,
                                                             int variant, float* __restrict__ mesh) {
  const unsigned beg = offsets[bucket], end = offsets[bucket + 1];
  const size_t n2 = (size_t)n * n;
  for (unsigned i = beg + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
    const float4 r = sorted[i];
    int x0, x1, y0, y1, z0, z1;
    float mdx, ddx, mdy, ddy, mdz, ddz;
    cic_reference_axis(r.x, n, wrap, variant, x0, x1, mdx, ddx);
    cic_reference_axis(r.y, n, wrap, variant, y0, y1, mdy, ddy);
    cic_reference_axis(r.z, n, wrap, variant, z0, z1, mdz, ddz);
    x0 = local_plane(x0, gx0, nx, n);
    x1 = local_plane(x1, gx0, nx, n);
#define JPS_CORNER(ix, iy, iz, wx, wy, wz)                                           \
  if (((ix) | (iy) | (iz)) >= 0)                                                      \
    atomicAdd(mesh + (size_t)(ix) * n2 + (size_t)(iy) * n + (iz), (((wx) * (wy)) * (wz)) * r.w);
    JPS_CORNER(x0, y0, z0, mdx, mdy, mdz)
    JPS_CORNER(x1, y0, z0, ddx, mdy, mdz)
    JPS_CORNER(x0, y1, z0, mdx, ddy, mdz)
    JPS_CORNER(x0, y0, z1, mdx, mdy, ddz)
    JPS_CORNER(x1, y1, z0, ddx, ddy, mdz)
    JPS_CORNER(x1, y0, z1, ddx, mdy, ddz)
    JPS_CORNER(x0, y1, z1, mdx, mdy, ddz)
    JPS_CORNER(x1, y1, z1, ddx, ddy, ddz)
#undef JPS_CORNER
  }
}

// ---------------------------------------------------------------- host side
struct SortedLayout {
  size_t sorted, tmp, counts, offsets, cursor, block_tot, wmax, gcounts, gbase, gcursor, parts, nparts, total;
  int parts_cap, slices_cap;
  size_t slices;
  int nbuckets;
};

static int replicas_for(int ntiles) {
  static const int rep_env = [] { const char* e = getenv("JPS_BUCKET_REP"); return e ? atoi(e) : 0; }();
  if (!rep_env && ntiles <= kSmemCountMaxTiles) return 1;   // shared-memory histogram: no same-address pressure
  int rep = rep_env ? rep_env : 8;
  while (rep > 1 && (long long)ntiles * rep > (1 << 19)) rep >>= 1;   // keep the scan + hot lines small
  return rep;
}

static SortedLayout sorted_layout(int n, int nx, int64_t n_part) {
  SortedLayout L;
  const int nt = (n + TILE - 1) / TILE, ntx = (nx + TILE - 1) / TILE;
  L.nbuckets = ntx * nt * nt * replicas_for(ntx * nt * nt) + 1;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
  L.sorted = take((size_t)(n_part > 0 ? n_part : 1) * sizeof(float4));
  L.counts = take((size_t)(L.nbuckets + 1) * 4);
  L.offsets = take((size_t)(L.nbuckets + 1) * 4);
  L.cursor = take((size_t)(L.nbuckets + 1) * 4);
  L.block_tot = take((size_t)(L.nbuckets / 2048 + 4) * 4);
  L.wmax = take(256);
  L.tmp = take((size_t)(n_part > 0 ? n_part : 1) * sizeof(float4));       // coarse-partitioned records (K1 v2)
  L.gcounts = take((size_t)(kMaxGroups + 1) * 4);
  L.gbase = take((size_t)(kMaxGroups + 1) * 4);
  L.gcursor = take((size_t)(kMaxGroups + 1) * 4);
  L.parts_cap = (int)((n_part > 0 ? n_part : 0) / kTileSplit) + 1;        // sum over tiles of floor((count - 1) / split) <= n_part / split
  L.parts = take((size_t)L.parts_cap * sizeof(uint2));
  L.nparts = take(256);                            // [0]: parts of heavy tiles, [4]: slices of heavy groups
  L.slices_cap = (int)((n_part > 0 ? n_part : 0) / kGroupSlice + (n_part > 0 ? n_part : 0) / kGroupSplit) + 2;
  L.slices = take((size_t)L.slices_cap * sizeof(uint2));
  L.total = off;
  return L;
}

size_t paint_sorted_workspace(int n, int64_t n_part, int order) {
  (void)order;
  return sorted_layout(n, n, n_part).total + 4096;   // a full mesh bounds every slab of it
}

// bucket one piece of the catalogue (count -> scan -> scatter) on stream `s`
template <int ORDER, bool REFCIC>
static int run_bucket(const PaintParams& p, const TileGeom& g, const SortedLayout& L, char* ws, cudaStream_t s) {
  unsigned* counts = (unsigned*)(ws + L.counts);
  unsigned* offsets = (unsigned*)(ws + L.offsets);
  unsigned* cursor = (unsigned*)(ws + L.cursor);
  float4* sorted = (float4*)(ws + L.sorted);
  unsigned* wmax_bits = (unsigned*)(ws + L.wmax);
  const int threads = 256;
  const int64_t want = (p.n_part + (int64_t)threads * BUCKET_UNROLL - 1) / ((int64_t)threads * BUCKET_UNROLL);
  const int blocks = (int)std::min<int64_t>(want, (int64_t)kNumSMs * 8 * 2);
  {
    ScopedLaunch T(K_MEMSET, s);
    JPS_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)(L.nbuckets + 1) * 4, s));
    JPS_CHECK_CUDA(cudaMemsetAsync(wmax_bits, 0, 4, s));
  }
  if (g.rep == 1 && g.ntiles <= kSmemCountMaxTiles) {
    const int smem = (g.ntiles + 1) * (int)sizeof(unsigned);
    static PerDeviceFlag attr_set;
    if (!attr_set.get()) {
      JPS_CHECK_CUDA(cudaFuncSetAttribute(bucket_count_smem_kernel<ORDER, REFCIC>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (kSmemCountMaxTiles + 1) * (int)sizeof(unsigned)));
      attr_set.set();
    }
    const int64_t w1 = (p.n_part + 1024 * BUCKET_UNROLL - 1) / (1024 * BUCKET_UNROLL);
    ScopedLaunch T(K_BUCKET_COUNT, s);
    bucket_count_smem_kernel<ORDER, REFCIC><<<(int)std::min<int64_t>(w1, kNumSMs), 1024, smem, s>>>(p, g, counts, wmax_bits);
  } else {
    ScopedLaunch T(K_BUCKET_COUNT, s);
    bucket_count_kernel<ORDER, REFCIC><<<blocks, threads, 0, s>>>(p, g, counts, wmax_bits);
  }
  JPS_CHECK_LAUNCH();
  {
    const int nsb = (L.nbuckets + SCAN_BLOCK - 1) / SCAN_BLOCK;
    unsigned* block_tot = (unsigned*)(ws + L.block_tot);
    { ScopedLaunch T(K_BUCKET_SCAN, s); scan_block_totals_kernel<<<nsb, SCAN_THREADS, 0, s>>>(counts, block_tot, L.nbuckets); }
    { ScopedLaunch T(K_BUCKET_SCAN, s); scan_of_totals_kernel<<<1, SCAN_THREADS, 0, s>>>(block_tot, nsb); }
    { ScopedLaunch T(K_BUCKET_SCAN, s); scan_write_kernel<<<nsb, SCAN_THREADS, 0, s>>>(counts, block_tot, offsets, cursor, L.nbuckets, nsb); }
  }
  JPS_CHECK_LAUNCH();
  {
    static const int unr = [] { const char* e = getenv("JPS_SCATTER_UNROLL"); return e ? atoi(e) : 4; }();
    static const int bpsm = [] { const char* e = getenv("JPS_SCATTER_BLOCKS_PER_SM"); return e ? atoi(e) : 16; }();
    const int64_t w2 = (p.n_part + (int64_t)threads * unr - 1) / ((int64_t)threads * unr);
    const int b2 = (int)std::min<int64_t>(w2, (int64_t)kNumSMs * bpsm);
    ScopedLaunch T(K_BUCKET_SCATTER, s);
    if (unr == 1) bucket_scatter_kernel<ORDER, REFCIC, 1><<<b2, threads, 0, s>>>(p, g, cursor, sorted);
    else if (unr == 2) bucket_scatter_kernel<ORDER, REFCIC, 2><<<b2, threads, 0, s>>>(p, g, cursor, sorted);
    else bucket_scatter_kernel<ORDER, REFCIC, 4><<<b2, threads, 0, s>>>(p, g, cursor, sorted);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

// K1 v2 host side.  `g.rep` must be 1 (offsets are per tile).
template <int ORDER, bool REFCIC>
static int run_bucket_two_level(const PaintParams& p, const TileGeom& g, const SortedLayout& L, char* ws,
                                cudaStream_t s) {
  unsigned* offsets = (unsigned*)(ws + L.offsets);
  float4* sorted = (float4*)(ws + L.sorted);
  float4* tmp = (float4*)(ws + L.tmp);
  unsigned* wmax_bits = (unsigned*)(ws + L.wmax);
  unsigned* gcounts = (unsigned*)(ws + L.gcounts);
  unsigned* gbase = (unsigned*)(ws + L.gbase);
  unsigned* gcursor = (unsigned*)(ws + L.gcursor);
  const int nbuckets = g.ntiles + 1;               // + the outlier bucket
  int gshift = 0;
  while (((nbuckets + (1 << gshift) - 1) >> gshift) > max_groups()) ++gshift;
  const int ngroups = (nbuckets + (1 << gshift) - 1) >> gshift;
  // Tile offsets BEFORE the partition (the fine pass then reads its records once): from the per-SM
  // shared-memory histogram when the tile table fits it, from global reds into the (L2-resident) table
  // otherwise.  JPS_FINE_TWOPASS=1 restores round 1's big-mesh path (group histogram, fine pass reads twice).
  static const bool two_pass = [] { const char* e = getenv("JPS_FINE_TWOPASS"); return e && atoi(e) != 0; }();
  // JPS_COUNT=fused|separate: big meshes only -- tile histogram from reds issued by the coarse pass (the count pass is
  // then the shared-memory group histogram) or by a pass of its own
  static const int count_forced = [] {
    const char* e = getenv("JPS_COUNT");
    return !e ? 0 : !strcmp(e, "fused") ? 1 : !strcmp(e, "separate") ? 2 : 0;
  }();
  const bool smem_hist = g.ntiles <= kSmemCountMaxTiles;        // tile histogram fits one SM's shared memory
  const bool have_offsets = smem_hist || !two_pass;
  const bool count_in_coarse = !smem_hist && !two_pass && (count_forced == 1 || (count_forced == 0 && kCountInCoarseDefault));
  const int nsb = (nbuckets + SCAN_BLOCK - 1) / SCAN_BLOCK;
  unsigned* block_tot = (unsigned*)(ws + L.block_tot);
  if (have_offsets && !count_in_coarse) {
    // tile-level histogram (shared-memory privatised) + scan -> offsets[]; group bases are a sample of it
    unsigned* counts = (unsigned*)(ws + L.counts);
    unsigned* cursor = (unsigned*)(ws + L.cursor);
    {
      ScopedLaunch T(K_MEMSET, s);
      JPS_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)(nbuckets + 1) * 4, s));
      JPS_CHECK_CUDA(cudaMemsetAsync(wmax_bits, 0, 4, s));
    }
    if (smem_hist) {
      const int smem = (g.ntiles + 1) * (int)sizeof(unsigned);
      static PerDeviceFlag attr_set;
      if (!attr_set.get()) {
        JPS_CHECK_CUDA(cudaFuncSetAttribute(bucket_count_smem_kernel<ORDER, REFCIC>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (kSmemCountMaxTiles + 1) * (int)sizeof(unsigned)));
        attr_set.set();
      }
      const int64_t w1 = (p.n_part + 1024 * BUCKET_UNROLL - 1) / (1024 * BUCKET_UNROLL);
      ScopedLaunch T(K_BUCKET_COUNT, s);
      bucket_count_smem_kernel<ORDER, REFCIC><<<(int)std::min<int64_t>(w1, kNumSMs), 1024, smem, s>>>(p, g, counts, wmax_bits);
    } else {
      const int64_t want = (p.n_part + 256 * 4 - 1) / (256 * 4);
      ScopedLaunch T(K_BUCKET_COUNT, s);
      bucket_count_global_kernel<ORDER, REFCIC><<<(int)std::min<int64_t>(want, (int64_t)kNumSMs * 8), 256, 0, s>>>(p, g, counts, wmax_bits);
    }
    JPS_CHECK_LAUNCH();
    {
      { ScopedLaunch T(K_BUCKET_SCAN, s); scan_block_totals_kernel<<<nsb, SCAN_THREADS, 0, s>>>(counts, block_tot, nbuckets); }
      { ScopedLaunch T(K_BUCKET_SCAN, s); scan_of_totals_kernel<<<1, SCAN_THREADS, 0, s>>>(block_tot, nsb); }
      { ScopedLaunch T(K_BUCKET_SCAN, s); scan_write_kernel<<<nsb, SCAN_THREADS, 0, s>>>(counts, block_tot, offsets, cursor, nbuckets, nsb); }
      { ScopedLaunch T(K_BUCKET_SCAN, s); group_bases_from_offsets_kernel<<<(ngroups + 256) / 256, 256, 0, s>>>(offsets, nbuckets, gshift, ngroups, gbase, gcursor); }
    }
    JPS_CHECK_LAUNCH();
  } else {
  {
    ScopedLaunch T(K_MEMSET, s);
    JPS_CHECK_CUDA(cudaMemsetAsync(gcounts, 0, (size_t)(ngroups + 1) * 4, s));
    JPS_CHECK_CUDA(cudaMemsetAsync(wmax_bits, 0, 4, s));
    if (count_in_coarse) JPS_CHECK_CUDA(cudaMemsetAsync(ws + L.counts, 0, (size_t)(nbuckets + 1) * 4, s));
  }
  {
    const int64_t want = (p.n_part + 256 * BUCKET_UNROLL - 1) / (256 * BUCKET_UNROLL);
    ScopedLaunch T(K_BUCKET_COUNT, s);
    coarse_count_kernel<ORDER, REFCIC><<<(int)std::min<int64_t>(want, (int64_t)kNumSMs * 8), 256, ngroups * 4, s>>>(
        p, g, gshift, ngroups, gcounts, wmax_bits);
  }
  JPS_CHECK_LAUNCH();
  {
    ScopedLaunch T(K_BUCKET_SCAN, s);
    group_scan_kernel<<<1, 1024, 0, s>>>(gcounts, gbase, gcursor, ngroups);
  }
  JPS_CHECK_LAUNCH();
  }
  {
    const size_t smem = (size_t)COARSE_CHUNK * (sizeof(float4) + 2) + (size_t)ngroups * 8 + 33 * 4;
    const int smem_max = (int)((size_t)COARSE_CHUNK * (sizeof(float4) + 2) + (size_t)kMaxGroups * 8 + 33 * 4);
    static PerDeviceFlag attr_set;
    if (!attr_set.get()) {
      JPS_CHECK_CUDA(cudaFuncSetAttribute(coarse_scatter_kernel<ORDER, REFCIC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      JPS_CHECK_CUDA(cudaFuncSetAttribute(coarse_scatter_kernel<ORDER, REFCIC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      attr_set.set();
    }
    const int64_t nchunks = (p.n_part + COARSE_CHUNK - 1) / COARSE_CHUNK;
    const int cblocks = (int)std::min<int64_t>(nchunks, JPS_COARSE_MINB * kNumSMs);
    ScopedLaunch T(K_BUCKET_SCATTER, s);
    if (count_in_coarse)
      coarse_scatter_kernel<ORDER, REFCIC, true><<<cblocks, COARSE_THREADS, smem, s>>>(p, g, gshift, ngroups, gcursor, tmp,
                                                                                      (unsigned*)(ws + L.counts));
    else
      coarse_scatter_kernel<ORDER, REFCIC, false><<<cblocks, COARSE_THREADS, smem, s>>>(p, g, gshift, ngroups, gcursor, tmp, nullptr);
  }
  JPS_CHECK_LAUNCH();
  if (count_in_coarse) {                           // tile offsets from the counts the coarse pass left behind
    unsigned* counts = (unsigned*)(ws + L.counts);
    unsigned* cursor = (unsigned*)(ws + L.cursor);
    { ScopedLaunch T(K_BUCKET_SCAN, s); scan_block_totals_kernel<<<nsb, SCAN_THREADS, 0, s>>>(counts, block_tot, nbuckets); }
    { ScopedLaunch T(K_BUCKET_SCAN, s); scan_of_totals_kernel<<<1, SCAN_THREADS, 0, s>>>(block_tot, nsb); }
    { ScopedLaunch T(K_BUCKET_SCAN, s); scan_write_kernel<<<nsb, SCAN_THREADS, 0, s>>>(counts, block_tot, offsets, cursor, nbuckets, nsb); }
    JPS_CHECK_LAUNCH();
  }
  {
    const size_t smem = (size_t)(2 << gshift) * 4 + 33 * 4;
    // staged form: JPS_FINE=staged|direct forces one (A/B runs, tests)
    static const int fine_forced = [] {
      const char* e = getenv("JPS_FINE");
      return !e ? 0 : !strcmp(e, "staged") ? 1 : !strcmp(e, "direct") ? 2 : 0;
    }();
    // JPS_FINE_CHUNK=big|small forces the chunk shape of the staged form
    static const int chunk_forced = [] {
      const char* e = getenv("JPS_FINE_CHUNK");
      return !e ? 0 : !strcmp(e, "big") ? 1 : !strcmp(e, "small") ? 2 : 0;
    }();
    const bool staged = have_offsets && gshift <= 11 &&
                        (fine_forced == 1 || (fine_forced == 0 && gshift >= kFineStagedMinShift));
    const bool big = chunk_forced == 1 || (chunk_forced == 0 && gshift >= kFineStagedBigShift);
    // heavy groups (> kGroupSplit records) go to fine_heavy_kernel; it needs the global tile cursors of the scan
    const int skip_heavy = have_offsets ? 1 : 0;
    uint2* slices = (uint2*)(ws + L.slices);
    unsigned* nslices = (unsigned*)(ws + L.nparts + 16);
    ScopedLaunch T(K_BUCKET_FINE, s);
    if (skip_heavy) {
      JPS_CHECK_CUDA(cudaMemsetAsync(nslices, 0, 4, s));
      group_slices_kernel<<<(ngroups + 255) / 256, 256, 0, s>>>(gbase, ngroups, slices, nslices, L.slices_cap);
    }
    if (staged) {
      const int rc = big ? launch_fine_staged<ORDER, REFCIC, 8192, 1024, 1>(tmp, gbase, g, gshift, ngroups, nbuckets, offsets, sorted, skip_heavy, s)
                         : launch_fine_staged<ORDER, REFCIC, 4096, 512, 2>(tmp, gbase, g, gshift, ngroups, nbuckets, offsets, sorted, skip_heavy, s);
      if (rc) return rc;
    } else if (have_offsets)
      fine_scatter_kernel<ORDER, REFCIC, true><<<ngroups, 512, smem, s>>>(tmp, gbase, g, gshift, ngroups, nbuckets, offsets, sorted, skip_heavy);
    else
      fine_scatter_kernel<ORDER, REFCIC, false><<<ngroups, 512, smem, s>>>(tmp, gbase, g, gshift, ngroups, nbuckets, offsets, sorted, 0);
    if (skip_heavy)
      fine_heavy_kernel<ORDER, REFCIC><<<L.slices_cap, 512, 0, s>>>(tmp, gbase, g, slices, nslices, (unsigned*)(ws + L.cursor), sorted);
  }
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// float32 mesh [nx][n][n] as a 3-D tensor (z fastest) with a [L][L][LP] box; false if the TMA cannot take it
static bool make_mesh_tensor_map(CUtensorMap* tm, float* mesh, int n, int nx, int L, int LP) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc || n % 4 != 0 || ((uintptr_t)mesh & 15) != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)nx};
  const cuuint64_t strides[2] = {(cuuint64_t)n * 4, (cuuint64_t)n * n * 4};     // bytes, dims 1 and 2
  const cuuint32_t box[3] = {(cuuint32_t)LP, (cuuint32_t)L, (cuuint32_t)L};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, mesh, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// deposit a bucketed piece into the mesh on stream `s`
template <int ORDER, bool REFCIC>
static int run_deposit(const PaintParams& p, const TileGeom& g, const SortedLayout& L, char* ws, cudaStream_t s,
                       int tx_begin, int tx_end) {
  const int tile_offset = tx_begin * g.nt * g.nt;
  const int tile_count = (tx_end - tx_begin) * g.nt * g.nt;
  if (tile_count <= 0) return JPS_OK;
  const unsigned* offsets = (const unsigned*)(ws + L.offsets);
  const float4* sorted = (const float4*)(ws + L.sorted);
  const unsigned* wmax_bits = (const unsigned*)(ws + L.wmax);
  {
    ScopedLaunch T(K_PAINT_TILE, s);
    const int mesh_vec_ok = (((uintptr_t)p.mesh) & 15) == 0 ? 1 : 0;
    static const bool plain = [] { const char* e = getenv("JPS_TILE_KERNEL"); return e && !strcmp(e, "plain"); }();
    if (plain) {
      JPS_REQUIRE(tile_count == g.ntiles, "JPS_TILE_KERNEL=plain cannot deposit a range of tile rows");
      paint_tile_kernel<ORDER, REFCIC><<<g.ntiles, 256, 0, s>>>(sorted, offsets, g, p.wrap, p.variant,
                                                              mesh_vec_ok, p.mesh);
    } else {
      constexpr int smem = 2 * TileDims<ORDER>::CELLS * (int)sizeof(unsigned);
      static PerDeviceFlag attr_set;
      if (!attr_set.get()) {
        JPS_CHECK_CUDA(cudaFuncSetAttribute(paint_tile_fx_kernel<ORDER, REFCIC>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set.set();
      }
      // CTA size (measured, C2 / a C4 rank): 256 threads 1.94 ms, 384 1.56, 512 1.61, 768 1.89 for TSC;
      // PCS on sparse tiles 9.2 (256), 6.67 (384), 6.41 (512) ms; CIC flat between 320 and 512.
      static const int tpb_env = [] { const char* e = getenv("JPS_TILE_THREADS"); return e ? atoi(e) : 0; }();
      const int tpb = tpb_env > 0 ? std::max(32, std::min(tpb_env, 512) & ~31) : (ORDER == 3 ? 384 : 512);   // whole warps
      // JPS_TILE_FLUSH=red forces the per-thread red flush everywhere (A/B runs, tests)
      static const bool no_tma = [] { const char* e = getenv("JPS_TILE_FLUSH"); return e && !strcmp(e, "red"); }();
      CUtensorMap tmap;
      memset(&tmap, 0, sizeof(tmap));
      const int use_tma = (!no_tma && make_mesh_tensor_map(&tmap, p.mesh, g.n, g.nx, TileDims<ORDER>::L, TileDims<ORDER>::LP)) ? 1 : 0;
      // JPS_TILE_ORDER=bank: particles of a tile in bank-class order instead of arrival order (measured slower, see
      // the kernel); JPS_FX_BITS=24..31: fixed-point position (default 31)
      static const int order_forced = [] {
        const char* e = getenv("JPS_TILE_ORDER");
        return !e ? 0 : !strcmp(e, "bank") ? 1 : !strcmp(e, "arrival") ? 2 : 0;
      }();
      static const int fx_env = [] { const char* e = getenv("JPS_FX_BITS"); const int v = e ? atoi(e) : 0; return (v >= 24 && v <= 31) ? v : 0; }();
      const int bank_order = (!REFCIC && order_forced == 1) ? 1 : 0;
      const int fx_bits = fx_env ? fx_env : 31;
      paint_tile_fx_kernel<ORDER, REFCIC><<<tile_count, tpb, smem, s>>>(sorted, offsets, g, p.wrap, p.variant,
                                                                      mesh_vec_ok, p.w ? 1 : 0, p.mesh, tmap, use_tma,
                                                                      tile_offset, bank_order, fx_bits, nullptr, nullptr, tile_count);
      // further parts of heavy tiles: at most parts_cap list entries, CTAs beyond the list's length return at once
      paint_tile_fx_kernel<ORDER, REFCIC><<<L.parts_cap, tpb, smem, s>>>(sorted, offsets, g, p.wrap, p.variant,
                                                                       mesh_vec_ok, p.w ? 1 : 0, p.mesh, tmap, use_tma,
                                                                       tile_offset, bank_order, fx_bits,
                                                                       (const uint2*)(ws + L.parts), (const unsigned*)(ws + L.nparts),
                                                                       tile_count);
    }
  }
  JPS_CHECK_LAUNCH();
  if (REFCIC && tx_begin == 0) {                  // the outlier bucket goes with the first range
    ScopedLaunch T(K_PAINT_ATOMIC, s);
    paint_outliers_kernel<<<kNumSMs, 256, 0, s>>>(sorted, offsets, g.ntiles * g.rep, g.n, g.x0, g.nx, p.wrap,
                                                  p.variant, p.mesh);
    JPS_CHECK_LAUNCH();
  }
  return JPS_OK;
}

// (Measured and dropped in round 1: cutting the catalogue into 4 pieces and bucketing piece c+1 on
// an auxiliary stream while piece c is deposited.  The step got SLOWER, 6.9 -> 8.5 ms on C2: each
// piece re-zeroes and re-flushes every tile, so the deposit grows from 2.7 to 4 x 1.1 ms, and the
// two kernels do not overlap well enough to pay that back.)
// phase 0: bucket + deposit everything; 1: bucket only; 2: deposit tile rows [tx_begin, tx_end) of a workspace
// bucketed by an earlier phase-1 call with the same arguments
template <int ORDER, bool REFCIC>
static int run_sorted(const PaintParams& p, const TileGeom& g, char* ws, size_t ws_bytes, cudaStream_t s, int phase,
                      int tx_begin, int tx_end) {
  (void)ws_bytes;
  const SortedLayout L = sorted_layout(p.n, p.nx, p.n_part);
  if (phase != 2) {
    int rc = g.two_level ? run_bucket_two_level<ORDER, REFCIC>(p, g, L, ws, s) : run_bucket<ORDER, REFCIC>(p, g, L, ws, s);
    if (rc) return rc;
    {                                              // list of the further parts of heavy tiles (usually empty)
      ScopedLaunch T(K_BUCKET_SCAN, s);
      JPS_CHECK_CUDA(cudaMemsetAsync(ws + L.nparts, 0, 4, s));
      tile_parts_kernel<<<(g.ntiles + 255) / 256, 256, 0, s>>>((const unsigned*)(ws + L.offsets), g.ntiles, g.rep,
                                                                (uint2*)(ws + L.parts), (unsigned*)(ws + L.nparts), L.parts_cap);
    }
    JPS_CHECK_LAUNCH();
  }
  if (phase == 1) return JPS_OK;
  if (phase == 0) { tx_begin = 0; tx_end = g.ntx; }
  JPS_REQUIRE(tx_begin >= 0 && tx_end <= g.ntx && tx_begin <= tx_end, "jps_paint: tile rows [%d, %d) outside [0, %d)", tx_begin, tx_end, g.ntx);
  return run_deposit<ORDER, REFCIC>(p, g, L, ws, s, tx_begin, tx_end);
}

int paint_sorted(const PaintParams& p, int order, int compat, void* ws, size_t ws_bytes,
                 cudaStream_t s, int phase, int tx_begin, int tx_end) {
  if (p.n_part == 0) return JPS_OK;
  JPS_REQUIRE(p.n_part < ((int64_t)1 << 32) - 1, "jps_paint: the sorted painter takes < 2^32 particles per call");
  const SortedLayout L = sorted_layout(p.n, p.nx, p.n_part);
  if (ws == nullptr || ws_bytes < L.total) {
    set_error("jps_paint: workspace has %zu bytes, %zu needed (jps_paint_workspace_bytes)", ws_bytes, L.total);
    return JPS_ERR_WORKSPACE;
  }
  JPS_REQUIRE(((uintptr_t)ws & 15) == 0, "jps_paint: workspace must be 16-byte aligned");
  TileGeom g;
  g.n = p.n;
  g.x0 = p.x0;
  g.nx = p.nx;
  g.nt = (p.n + TILE - 1) / TILE;
  g.ntx = (p.nx + TILE - 1) / TILE;
  g.ntiles = g.ntx * g.nt * g.nt;
  // Bucketing flavour.  Measured on C2 (N=512, 1e8 particles): single-level 0.60 + 2.36 = 2.96 ms;
  // two-level 0.60 (tile histogram) + 1.10 (coarse) + 0.78 (fine, one read) = 2.51 ms, or
  // 0.38 + 1.10 + 1.32 = 2.80 ms when the tile histogram does not fit shared memory and the fine pass
  // reads its records twice.  The two-level partition is the default whenever the tile count fits
  // its group table; JPS_BUCKET=atomic|two forces one (A/B runs, tests).
  static const int forced = [] {
    const char* e = getenv("JPS_BUCKET");
    return !e ? 0 : !strcmp(e, "atomic") ? 1 : !strcmp(e, "two") ? 2 : 0;
  }();
  const bool fits_two = g.ntiles + 1 <= kMaxGroups * 2048;
  g.two_level = (forced != 1 && fits_two) ? 1 : 0;
  g.rep = g.two_level ? 1 : replicas_for(g.ntiles);
  char* w = (char*)ws;
  if (order == 2 && compat == JPS_COMPAT_REFERENCE) return run_sorted<2, true>(p, g, w, ws_bytes, s, phase, tx_begin, tx_end);
  if (order == 2) return run_sorted<2, false>(p, g, w, ws_bytes, s, phase, tx_begin, tx_end);
  if (order == 3) return run_sorted<3, false>(p, g, w, ws_bytes, s, phase, tx_begin, tx_end);
  return run_sorted<4, false>(p, g, w, ws_bytes, s, phase, tx_begin, tx_end);
}

}  // namespace jps
