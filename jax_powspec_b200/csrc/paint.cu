// Mesh painting: jps_paint and its kernels.
//
// K2' paint_atomic: one thread per particle, red.global.add.f32 per stencil cell.  Correct for
// any particle order; the small-mesh (L2-resident) fast path and the fallback for particles
// outside the box in compat=reference.
#include "paint_common.cuh"

namespace jps {

__device__ __forceinline__ void red_add(float* addr, float v) {
  atomicAdd(addr, v);   // result unused -> REDG.E.ADD.F32
}

template <int ORDER, bool REFCIC>
__global__ void __launch_bounds__(256) paint_atomic_kernel(PaintParams p) {
  const int n = p.n;
  const size_t n2 = (size_t)n * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_part;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float wgt = p.w ? p.w[i] : 1.0f;
    const float px = grid_pos(p.x[i * p.stride], p.xmin, p.inv);
    const float py = grid_pos(p.y[i * p.stride], p.ymin, p.inv);
    const float pz = grid_pos(p.z[i * p.stride], p.zmin, p.inv);
    if (REFCIC) {
      int x0, x1, y0, y1, z0, z1;
      float mdx, ddx, mdy, ddy, mdz, ddz;
      cic_reference_axis(px, n, p.wrap, p.variant, x0, x1, mdx, ddx);
      cic_reference_axis(py, n, p.wrap, p.variant, y0, y1, mdy, ddy);
      cic_reference_axis(pz, n, p.wrap, p.variant, z0, z1, mdz, ddz);
      // the 8 scatters of src/mas.py:142-151, weights multiplied left to right
      x0 = local_plane(x0, p.x0, p.nx, n);
      x1 = local_plane(x1, p.x0, p.nx, n);
#define JPS_CORNER(ix, iy, iz, wx, wy, wz)                                          \
  if (((ix) | (iy) | (iz)) >= 0)                                                     \
    red_add(p.mesh + (size_t)(ix) * n2 + (size_t)(iy) * n + (iz), (((wx) * (wy)) * (wz)) * wgt);
      JPS_CORNER(x0, y0, z0, mdx, mdy, mdz)
      JPS_CORNER(x1, y0, z0, ddx, mdy, mdz)
      JPS_CORNER(x0, y1, z0, mdx, ddy, mdz)
      JPS_CORNER(x0, y0, z1, mdx, mdy, ddz)
      JPS_CORNER(x1, y1, z0, ddx, ddy, mdz)
      JPS_CORNER(x1, y0, z1, ddx, mdy, ddz)
      JPS_CORNER(x0, y1, z1, mdx, mdy, ddz)   // Q1: reference weight (textbook: mdx*ddy*ddz)
      JPS_CORNER(x1, y1, z1, ddx, ddy, ddz)
#undef JPS_CORNER
    } else {
      int ix[ORDER], iy[ORDER], iz[ORDER];
      float wx[ORDER], wy[ORDER], wz[ORDER];
      bspline_axis<ORDER>(px, n, p.wrap, ix, wx);
      bspline_axis<ORDER>(py, n, p.wrap, iy, wy);
      bspline_axis<ORDER>(pz, n, p.wrap, iz, wz);
#pragma unroll
      for (int a = 0; a < ORDER; ++a) {
        const int lx = local_plane(ix[a], p.x0, p.nx, n);
#pragma unroll
        for (int b = 0; b < ORDER; ++b) {
          if ((lx | iy[b]) < 0) continue;
          float* row = p.mesh + (size_t)lx * n2 + (size_t)iy[b] * n;
          const float wxy = wx[a] * wy[b];
#pragma unroll
          for (int c = 0; c < ORDER; ++c) {
            if (iz[c] >= 0) red_add(row + iz[c], (wxy * wz[c]) * wgt);
          }
        }
      }
    }
  }
}

static int launch_atomic(const PaintParams& p, int order, int compat, cudaStream_t s) {
  if (p.n_part == 0) return JPS_OK;
  const int threads = 256;
  int64_t blocks64 = (p.n_part + threads - 1) / threads;
  const int64_t cap = (int64_t)kNumSMs * 8 * 16;
  int blocks = (int)(blocks64 < cap ? blocks64 : cap);
  ScopedLaunch L(K_PAINT_ATOMIC, s);
  if (order == 2 && compat == JPS_COMPAT_REFERENCE)
    paint_atomic_kernel<2, true><<<blocks, threads, 0, s>>>(p);
  else if (order == 2)
    paint_atomic_kernel<2, false><<<blocks, threads, 0, s>>>(p);
  else if (order == 3)
    paint_atomic_kernel<3, false><<<blocks, threads, 0, s>>>(p);
  else
    paint_atomic_kernel<4, false><<<blocks, threads, 0, s>>>(p);
  JPS_CHECK_LAUNCH();
  return JPS_OK;
}

int paint_sorted(const PaintParams& p, int order, int compat, void* ws, size_t ws_bytes,
                 cudaStream_t s, int phase, int tx_begin, int tx_end);   // paint_sorted.cu
size_t paint_sorted_workspace(int n, int64_t n_part, int order);

}  // namespace jps

using namespace jps;

extern "C" int jps_paint_workspace_bytes(int n_mesh, int64_t n_part, int order, int method,
                                         size_t* bytes) {
  JPS_REQUIRE(bytes != nullptr, "jps_paint_workspace_bytes: bytes is NULL");
  JPS_REQUIRE(n_mesh >= 2 && n_part >= 0, "jps_paint_workspace_bytes: bad sizes");
  JPS_REQUIRE(order >= 2 && order <= 4, "jps_paint_workspace_bytes: order must be 2, 3 or 4");
  if (method == JPS_PAINT_ATOMIC) { *bytes = 0; return JPS_OK; }
  *bytes = paint_sorted_workspace(n_mesh, n_part, order);
  return JPS_OK;
}

extern "C" int jps_paint(int n_mesh, const float* x, const float* y, const float* z,
                         const float* w, int64_t stride, int64_t n_part, float xmin, float ymin,
                         float zmin, float box_size, int order, int wrap, int compat, int variant,
                         int method, float* mesh, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return jps_paint_slab(n_mesh, 0, n_mesh, x, y, z, w, stride, n_part, xmin, ymin, zmin, box_size, order,
                        wrap, compat, variant, method, mesh, workspace, workspace_bytes, stream);
}

extern "C" int jps_paint_slab(int n_mesh, int x0, int nx_alloc, const float* x, const float* y,
                              const float* z, const float* w, int64_t stride, int64_t n_part,
                              float xmin, float ymin, float zmin, float box_size, int order, int wrap,
                              int compat, int variant, int method, float* mesh, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return jps_paint_slab_phase(n_mesh, x0, nx_alloc, x, y, z, w, stride, n_part, xmin, ymin, zmin, box_size, order, wrap,
                              compat, variant, method, mesh, workspace, workspace_bytes, JPS_PAINT_PHASE_ALL, 0, 0, stream);
}

extern "C" int jps_paint_tile_rows(int nx_alloc) { return (nx_alloc + 15) / 16; }

extern "C" int jps_paint_slab_phase(int n_mesh, int x0, int nx_alloc, const float* x, const float* y,
                                    const float* z, const float* w, int64_t stride, int64_t n_part,
                                    float xmin, float ymin, float zmin, float box_size, int order, int wrap,
                                    int compat, int variant, int method, float* mesh, void* workspace,
                                    size_t workspace_bytes, int phase, int tx_begin, int tx_end, void* stream) {
  JPS_REQUIRE(phase == JPS_PAINT_PHASE_ALL || phase == JPS_PAINT_PHASE_BUCKET || phase == JPS_PAINT_PHASE_DEPOSIT,
              "jps_paint_slab_phase: unknown phase %d", phase);
  JPS_REQUIRE(phase == JPS_PAINT_PHASE_ALL || method == JPS_PAINT_SORTED,
              "jps_paint_slab_phase: the bucket / deposit phases exist for method = JPS_PAINT_SORTED only");
  JPS_REQUIRE(nx_alloc >= 1 && nx_alloc <= n_mesh, "jps_paint_slab: nx_alloc=%d out of range [1,%d]", nx_alloc, n_mesh);
  JPS_REQUIRE(n_mesh >= 2 && n_mesh <= 4096, "jps_paint: n_mesh=%d out of range [2,4096]", n_mesh);
  JPS_REQUIRE(n_part >= 0, "jps_paint: n_part < 0");
  JPS_REQUIRE(order >= 2 && order <= 4, "jps_paint: order must be 2 (CIC), 3 (TSC) or 4 (PCS)");
  JPS_REQUIRE(compat == JPS_COMPAT_REFERENCE || compat == JPS_COMPAT_FIXED, "jps_paint: bad compat");
  JPS_REQUIRE(variant == JPS_VARIANT_VEC || variant == JPS_VARIANT_SCAN, "jps_paint: bad variant");
  JPS_REQUIRE(mesh != nullptr, "jps_paint: mesh is NULL");
  JPS_REQUIRE(n_part == 0 || (x && y && z), "jps_paint: x/y/z is NULL");
  JPS_REQUIRE(stride >= 1, "jps_paint: stride must be >= 1");
  JPS_REQUIRE(box_size > 0.0f, "jps_paint: box_size must be > 0");
  PaintParams p;
  p.n = n_mesh;
  p.x0 = x0;
  p.nx = nx_alloc;
  p.wrap = wrap ? 1 : 0;
  p.variant = variant;
  p.xmin = xmin; p.ymin = ymin; p.zmin = zmin;
  const float bin_size = box_size / (float)n_mesh;   // src/mas.py:100-101, float32
  p.inv = 1.0f / bin_size;
  p.stride = stride;
  p.n_part = n_part;
  p.x = x; p.y = y; p.z = z; p.w = w;
  p.mesh = mesh;
  cudaStream_t s = (cudaStream_t)stream;
  if (method == JPS_PAINT_AUTO) {
    // small meshes stay L2-resident: global reds are already local; few particles: not worth sorting
    const size_t mesh_bytes = (size_t)nx_alloc * n_mesh * n_mesh * 4;
    method = (mesh_bytes <= (size_t)48 << 20 || n_part < (int64_t)1 << 18) ? JPS_PAINT_ATOMIC
                                                                          : JPS_PAINT_SORTED;
    if (method == JPS_PAINT_SORTED && workspace_bytes < paint_sorted_workspace(n_mesh, n_part, order))  // sized for a full mesh: always enough
      method = JPS_PAINT_ATOMIC;                       // caller gave no room: still correct
  }
  if (method == JPS_PAINT_ATOMIC) return launch_atomic(p, order, compat, s);
  JPS_REQUIRE(method == JPS_PAINT_SORTED, "jps_paint: unknown method %d", method);
  return paint_sorted(p, order, compat, workspace, workspace_bytes, s, phase, tx_begin, tx_end);
}
