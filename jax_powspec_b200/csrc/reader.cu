// Catalogue text reader on the device.
//
// Replaces the host-side  np.loadtxt(path, usecols=(0,1,2), dtype=np.float32)  /
// pd.read_csv(path, usecols=(0,1,2), delim_whitespace=True).values.astype(np.float32)  followed by
// the box mask  ((p < box_size) & (p > 0)).all(axis=1)  of the reference's scripts
// (/root/reference/tests/correlations.py:29-31, tests/positions.py:25-27).  The raw bytes of the file
// are copied to the GPU once; everything else is byte work at HBM speed:
//   T1 text_count_newlines : 32 bytes per thread, byte-compare SIMD (__vcmpeq4), one count per 8 KB chunk
//   T2 scan_u32_to_u64     : exclusive scan of the chunk counts (one CTA)
//   T3 text_index_newlines : positions of all '\n' (int64), in file order
//   T4 text_parse_lines    : one thread per line; requested columns -> float32 with the exact
//                            decimal -> double -> float conversion of textparse.cuh; row keep flag
//                            (blank / comment lines dropped, optional box mask); kept rows per block
//   T5 scan (same kernel)  : block offsets of the kept rows
//   T6 text_compact_rows   : rows in file order into the dense [n_rows][ncols] output; the rare rows
//                            the device does not convert exactly are listed for the host
#include "common.cuh"
#include "textparse.cuh"

namespace jps {

constexpr int TEXT_THREADS = 256;
constexpr int TEXT_BYTES_PER_THREAD = 32;
constexpr int TEXT_CHUNK = TEXT_THREADS * TEXT_BYTES_PER_THREAD;   // 8 KB per CTA step
constexpr int TEXT_MAXC = 8;                                        // requested columns per row

struct TextCols {
  int col[TEXT_MAXC];
  int n;
};

// row status bits (one byte per line)
constexpr unsigned char ROW_KEEP = 1, ROW_SLOW = 2, ROW_BAD = 4;

// newlines among the 32 bytes at `base + 32 * t` (bytes at or beyond nbytes do not count)
__device__ __forceinline__ unsigned newline_mask32(const char* __restrict__ text, int64_t pos, int64_t nbytes) {
  unsigned mask = 0;
  if (pos + TEXT_BYTES_PER_THREAD <= nbytes) {
    const uint4* v = reinterpret_cast<const uint4*>(text + pos);
    const uint4 a = __ldg(v), b = __ldg(v + 1);
    const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned eq = __vcmpeq4(w[k], 0x0a0a0a0au);            // 0xff in every byte equal to '\n'
      // one bit per byte: bits 0, 8, 16, 24 of eq -> bits 4k .. 4k+3
      const unsigned bits = ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u));
      mask |= bits << (4 * k);
    }
  } else {
    for (int k = 0; k < TEXT_BYTES_PER_THREAD; ++k)
      if (pos + k < nbytes && text[pos + k] == '\n') mask |= 1u << k;
  }
  return mask;
}

__device__ __forceinline__ unsigned block_exclusive_scan_u32(unsigned v, unsigned* warp_tot, unsigned& total) {
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
  for (int w = 0; w < TEXT_THREADS / 32; ++w) {
    const unsigned t = warp_tot[w];
    if (w < warp) before += t;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return before + incl - v;
}

__global__ void __launch_bounds__(TEXT_THREADS) text_count_newlines_kernel(const char* __restrict__ text, int64_t nbytes,
                                                                           int64_t nchunks,
                                                                           unsigned* __restrict__ chunk_count) {
  __shared__ unsigned warp_tot[TEXT_THREADS / 32];
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int64_t pos = ch * TEXT_CHUNK + (int64_t)threadIdx.x * TEXT_BYTES_PER_THREAD;
    unsigned c = pos < nbytes ? __popc(newline_mask32(text, pos, nbytes)) : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned t = 0;
      for (int w = 0; w < TEXT_THREADS / 32; ++w) t += warp_tot[w];
      chunk_count[ch] = t;
    }
    __syncthreads();
  }
}

// off[0..m] = exclusive scan of c[0..m) in 64 bits (one CTA of 1024 threads)
__global__ void __launch_bounds__(1024) scan_u32_to_u64_kernel(const unsigned* __restrict__ c,
                                                               unsigned long long* __restrict__ off, int64_t m) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < m; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const unsigned long long v = i < m ? c[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;                            // inclusive scan of the warp totals
    }
    __syncthreads();
    const unsigned long long before = carry + (warp ? warp_tot[warp - 1] : 0ull);
    if (i < m) off[i] = before + incl - v;
    const unsigned long long total = warp_tot[31];
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[m] = carry;
}

__global__ void text_line_count_kernel(const char* __restrict__ text, int64_t nbytes,
                                       const unsigned long long* __restrict__ total_newlines,
                                       int64_t* __restrict__ n_lines) {
  // a last line without a trailing '\n' still counts
  const int64_t nl = (int64_t)*total_newlines;
  *n_lines = nl + ((nbytes > 0 && text[nbytes - 1] != '\n') ? 1 : 0);
}

__global__ void __launch_bounds__(TEXT_THREADS) text_index_newlines_kernel(const char* __restrict__ text, int64_t nbytes,
                                                                           int64_t nchunks,
                                                                           const unsigned long long* __restrict__ chunk_off,
                                                                           int64_t capacity, int64_t* __restrict__ nl_pos) {
  __shared__ unsigned warp_tot[TEXT_THREADS / 32];
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int64_t pos = ch * TEXT_CHUNK + (int64_t)threadIdx.x * TEXT_BYTES_PER_THREAD;
    unsigned mask = pos < nbytes ? newline_mask32(text, pos, nbytes) : 0u;
    unsigned total;
    const unsigned ex = block_exclusive_scan_u32(__popc(mask), warp_tot, total);
    int64_t o = (int64_t)chunk_off[ch] + ex;
    while (mask) {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      if (o < capacity) nl_pos[o] = pos + k;
      ++o;
    }
  }
}

struct TextParseParams {
  const char* text;
  int64_t nbytes;
  int64_t n_lines, n_newlines_capacity;
  const int64_t* nl_pos;
  int skiprows;
  char comment;
  int filter;
  float lo, hi;
  TextCols cols;
};

__device__ __forceinline__ void line_extent(const TextParseParams& p, int64_t i, int64_t& beg, int64_t& end) {
  beg = i == 0 ? 0 : p.nl_pos[i - 1] + 1;
  end = i < p.n_newlines_capacity ? p.nl_pos[i] : p.nbytes;
  if (end > p.nbytes) end = p.nbytes;
}

// The 256 lines of a CTA are one contiguous byte range (~12 KB for a 3-column catalogue): it is
// staged in shared memory with coalesced 16-byte loads and parsed from there, instead of ~50
// dependent one-byte global loads per thread.  Ranges longer than the staging buffer (very long
// lines) are parsed straight from global memory.
constexpr int TEXT_STAGE = 40960;

__global__ void __launch_bounds__(TEXT_THREADS) text_parse_lines_kernel(TextParseParams p, float* __restrict__ vals,
                                                                        unsigned char* __restrict__ status,
                                                                        unsigned* __restrict__ block_keep,
                                                                        unsigned long long* __restrict__ counters) {
  __shared__ unsigned warp_tot[TEXT_THREADS / 32];
  __shared__ __align__(16) char stext[TEXT_STAGE];
  __shared__ int64_t s_range[2];
  const int64_t i0 = (int64_t)blockIdx.x * TEXT_THREADS;
  const int64_t i = i0 + threadIdx.x;
  if (threadIdx.x == 0) {
    int64_t b0, e0, b1, e1;
    line_extent(p, i0, b0, e0);
    line_extent(p, min(i0 + TEXT_THREADS, p.n_lines) - 1, b1, e1);
    s_range[0] = b0 & ~(int64_t)15;                  // text is 16-byte aligned
    s_range[1] = e1;
  }
  __syncthreads();
  const int64_t abeg = s_range[0], span = s_range[1] - s_range[0];
  const bool staged = span <= TEXT_STAGE;
  if (staged) {
    for (int64_t off = (int64_t)threadIdx.x * 16; off < span; off += TEXT_THREADS * 16) {
      if (abeg + off + 16 <= p.nbytes) {
        *reinterpret_cast<uint4*>(stext + off) = __ldg(reinterpret_cast<const uint4*>(p.text + abeg + off));
      } else {
        for (int k = 0; k < 16 && abeg + off + k < p.nbytes; ++k) stext[off + k] = p.text[abeg + off + k];
      }
    }
    __syncthreads();
  }
  unsigned char st = 0;
  if (i < p.n_lines && i >= p.skiprows) {
    int64_t beg, end;
    line_extent(p, i, beg, end);
    const char* base = staged ? stext - abeg : p.text;
    float v[TEXT_MAXC];
    int fs;
    const int row = text::parse_line<TEXT_MAXC>(base + beg, base + end, p.comment, p.cols.col, p.cols.n, v, fs);
    if (row) {
      if (fs >= text::FIELD_BAD) {
        st = ROW_BAD;
        atomicAdd(counters + 2, 1ull);
        atomicMin(counters + 3, (unsigned long long)i);
      } else if (fs == text::FIELD_SLOW) {
        st = ROW_KEEP | ROW_SLOW;                      // the host converts (and, if asked, masks) this row
      } else {
        bool keep = true;
        if (p.filter)
          for (int c = 0; c < p.cols.n; ++c) keep = keep && (v[c] < p.hi) && (v[c] > p.lo);   // NaN fails both
        st = keep ? ROW_KEEP : 0;
      }
      if (st & ROW_KEEP)
        for (int c = 0; c < p.cols.n; ++c) vals[i * p.cols.n + c] = v[c];
    }
  }
  if (i < p.n_lines) status[i] = st;
  unsigned c = (st & ROW_KEEP) ? 1u : 0u;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < TEXT_THREADS / 32; ++w) t += warp_tot[w];
    block_keep[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(TEXT_THREADS) text_compact_rows_kernel(TextParseParams p, const float* __restrict__ vals,
                                                                         const unsigned char* __restrict__ status,
                                                                         const unsigned long long* __restrict__ block_off,
                                                                         int64_t nblocks, float* __restrict__ out,
                                                                         unsigned long long* __restrict__ counters,
                                                                         int64_t* __restrict__ slow_rows, int64_t slow_cap) {
  __shared__ unsigned warp_tot[TEXT_THREADS / 32];
  const int64_t i = (int64_t)blockIdx.x * TEXT_THREADS + threadIdx.x;
  const unsigned char st = i < p.n_lines ? status[i] : 0;
  unsigned total;
  const unsigned ex = block_exclusive_scan_u32((st & ROW_KEEP) ? 1u : 0u, warp_tot, total);
  if (st & ROW_KEEP) {
    const int64_t row = (int64_t)block_off[blockIdx.x] + ex;
    for (int c = 0; c < p.cols.n; ++c) out[row * p.cols.n + c] = vals[i * p.cols.n + c];
    if (st & ROW_SLOW) {
      const unsigned long long k = atomicAdd(counters + 1, 1ull);
      if ((int64_t)k < slow_cap) {
        int64_t beg, end;
        line_extent(p, i, beg, end);
        slow_rows[2 * k] = row;
        slow_rows[2 * k + 1] = beg;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] = block_off[nblocks];
}

struct TextLayout {
  size_t chunk_count, chunk_off, nl_pos, vals, status, block_keep, block_off, total;
  int64_t nchunks, nblocks;
};

static TextLayout text_layout(int64_t nbytes, int64_t n_lines, int ncols) {
  TextLayout L;
  L.nchunks = (nbytes + TEXT_CHUNK - 1) / TEXT_CHUNK;
  L.nblocks = (n_lines + TEXT_THREADS - 1) / TEXT_THREADS;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
  L.chunk_count = take((size_t)(L.nchunks + 1) * 4);
  L.chunk_off = take((size_t)(L.nchunks + 2) * 8);
  L.nl_pos = take((size_t)(n_lines + 1) * 8);
  L.vals = take((size_t)(n_lines + 1) * (size_t)(ncols > 0 ? ncols : 1) * 4);
  L.status = take((size_t)(n_lines + 1));
  L.block_keep = take((size_t)(L.nblocks + 1) * 4);
  L.block_off = take((size_t)(L.nblocks + 2) * 8);
  L.total = off;
  return L;
}

static int index_newlines(const char* text, int64_t nbytes, const TextLayout& L, char* ws, bool write_pos,
                          int64_t capacity, cudaStream_t s) {
  unsigned* chunk_count = reinterpret_cast<unsigned*>(ws + L.chunk_count);
  unsigned long long* chunk_off = reinterpret_cast<unsigned long long*>(ws + L.chunk_off);
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(L.nchunks, 1), 8 * kNumSMs);
  {
    ScopedLaunch Lc(K_TEXT_INDEX, s);
    text_count_newlines_kernel<<<grid, TEXT_THREADS, 0, s>>>(text, nbytes, L.nchunks, chunk_count);
    JPS_CHECK_LAUNCH();
  }
  {
    ScopedLaunch Lc(K_TEXT_INDEX, s);
    scan_u32_to_u64_kernel<<<1, 1024, 0, s>>>(chunk_count, chunk_off, L.nchunks);
    JPS_CHECK_LAUNCH();
  }
  if (write_pos) {
    ScopedLaunch Lc(K_TEXT_INDEX, s);
    text_index_newlines_kernel<<<grid, TEXT_THREADS, 0, s>>>(text, nbytes, L.nchunks, chunk_off, capacity,
                                                             reinterpret_cast<int64_t*>(ws + L.nl_pos));
    JPS_CHECK_LAUNCH();
  }
  return JPS_OK;
}

}  // namespace jps

using namespace jps;

extern "C" {

JPS_API size_t jps_text_workspace_bytes(int64_t nbytes, int64_t n_lines, int ncols) {
  if (nbytes < 0 || n_lines < 0 || ncols < 0 || ncols > TEXT_MAXC) return 0;
  return text_layout(nbytes, n_lines, ncols).total + 256;
}

JPS_API int jps_text_count_lines(const char* text, int64_t nbytes, int64_t* n_lines, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  JPS_REQUIRE(nbytes >= 0 && n_lines && (text || nbytes == 0), "jps_text_count_lines: bad arguments");
  JPS_REQUIRE((reinterpret_cast<uintptr_t>(text) & 15) == 0, "jps_text_count_lines: text must be 16-byte aligned");
  const TextLayout L = text_layout(nbytes, 0, 0);
  JPS_REQUIRE(workspace && workspace_bytes >= L.total, "jps_text_count_lines: workspace too small (%zu < %zu)",
              workspace_bytes, L.total);
  char* ws = static_cast<char*>(workspace);
  if (nbytes == 0) {
    JPS_CHECK_CUDA(cudaMemsetAsync(n_lines, 0, sizeof(int64_t), s));
    return JPS_OK;
  }
  const int rc = index_newlines(text, nbytes, L, ws, false, 0, s);
  if (rc) return rc;
  ScopedLaunch Lc(K_TEXT_INDEX, s);
  text_line_count_kernel<<<1, 1, 0, s>>>(text, nbytes,
                                         reinterpret_cast<unsigned long long*>(ws + L.chunk_off) + L.nchunks, n_lines);
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

JPS_API int jps_text_parse(const char* text, int64_t nbytes, int64_t n_lines, int skiprows, int comment,
                           const int* usecols, int ncols, int filter, float lo, float hi, float* out,
                           int64_t* counters, int64_t* slow_rows, int64_t slow_capacity, void* workspace,
                           size_t workspace_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  JPS_REQUIRE(nbytes >= 0 && n_lines >= 0 && counters && (text || nbytes == 0), "jps_text_parse: bad arguments");
  JPS_REQUIRE(ncols >= 1 && ncols <= TEXT_MAXC && usecols, "jps_text_parse: 1..%d columns", TEXT_MAXC);
  JPS_REQUIRE(skiprows >= 0 && slow_capacity >= 0 && (slow_rows || slow_capacity == 0), "jps_text_parse: bad arguments");
  JPS_REQUIRE((reinterpret_cast<uintptr_t>(text) & 15) == 0, "jps_text_parse: text must be 16-byte aligned");
  for (int c = 0; c < ncols; ++c) JPS_REQUIRE(usecols[c] >= 0, "jps_text_parse: negative column index");
  JPS_CHECK_CUDA(cudaMemsetAsync(counters, 0, 3 * sizeof(int64_t), s));
  JPS_CHECK_CUDA(cudaMemsetAsync(counters + 3, 0xff, sizeof(int64_t), s));      // first bad line: "none"
  if (n_lines == 0 || nbytes == 0) return JPS_OK;
  JPS_REQUIRE(out, "jps_text_parse: out is null");
  const TextLayout L = text_layout(nbytes, n_lines, ncols);
  JPS_REQUIRE(workspace && workspace_bytes >= L.total, "jps_text_parse: workspace too small (%zu < %zu)",
              workspace_bytes, L.total);
  char* ws = static_cast<char*>(workspace);
  // A last line without '\n' has no entry in nl_pos: pre-fill its slot with a huge offset, which
  // line_extent() clamps to the end of the file (the index kernel overwrites it when the '\n' exists).
  JPS_CHECK_CUDA(cudaMemsetAsync(ws + L.nl_pos + (size_t)(n_lines - 1) * 8, 0x7f, 8, s));
  const int rc = index_newlines(text, nbytes, L, ws, true, n_lines, s);
  if (rc) return rc;

  TextParseParams p;
  p.text = text;
  p.nbytes = nbytes;
  p.n_lines = n_lines;
  p.n_newlines_capacity = n_lines;
  p.nl_pos = reinterpret_cast<const int64_t*>(ws + L.nl_pos);
  p.skiprows = skiprows;
  p.comment = (char)comment;
  p.filter = filter;
  p.lo = lo;
  p.hi = hi;
  p.cols.n = ncols;
  for (int c = 0; c < TEXT_MAXC; ++c) p.cols.col[c] = c < ncols ? usecols[c] : 0;
  float* vals = reinterpret_cast<float*>(ws + L.vals);
  unsigned char* status = reinterpret_cast<unsigned char*>(ws + L.status);
  unsigned* block_keep = reinterpret_cast<unsigned*>(ws + L.block_keep);
  unsigned long long* block_off = reinterpret_cast<unsigned long long*>(ws + L.block_off);
  unsigned long long* ctr = reinterpret_cast<unsigned long long*>(counters);
  {
    ScopedLaunch Lc(K_TEXT_PARSE, s);
    text_parse_lines_kernel<<<(unsigned)L.nblocks, TEXT_THREADS, 0, s>>>(p, vals, status, block_keep, ctr);
    JPS_CHECK_LAUNCH();
  }
  {
    ScopedLaunch Lc(K_TEXT_COMPACT, s);
    scan_u32_to_u64_kernel<<<1, 1024, 0, s>>>(block_keep, block_off, L.nblocks);
    JPS_CHECK_LAUNCH();
  }
  {
    ScopedLaunch Lc(K_TEXT_COMPACT, s);
    text_compact_rows_kernel<<<(unsigned)L.nblocks, TEXT_THREADS, 0, s>>>(p, vals, status, block_off, L.nblocks, out,
                                                                          ctr, slow_rows, slow_capacity);
    JPS_CHECK_LAUNCH();
  }
  return JPS_OK;
}

}  // extern "C"
