// jax.ffi handlers that forward to libjps.so -- GATED: not part of the default build.
//
// north_star asks for jax.ffi custom calls; neither `jax` nor the XLA FFI headers
// (xla/ffi/api/ffi.h, shipped inside jaxlib: jax.ffi.include_dir()) exist in this image or on the
// GPU box, so this file cannot be linked or exercised here.  It is, however, COMPILED in the CPU test
// suite (tests/test_ffi_shim.py) against ffi/stub/xla/ffi/api/ffi.h, a stub of the binding API that
// rejects any handler whose Bind() chain and implementation disagree in arity or type.
// The C ABI of include/jps.h is shaped so each handler is a 1:1 forward: XLA owns every buffer
// (inputs, outputs, scratch), hands us its cudaStream_t, and we never allocate or synchronise.
//
// Build where JAX is installed:
//   g++ -O2 -fPIC -shared -std=c++17 -DJPS_WITH_JAX_FFI -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include ffi/jax_ffi_shim.cc -L jax_powspec_b200 -ljps -lcudart -o jps_jax_ffi.so
//
// Handlers (one per @jax.jit entry point of the reference's hot path):
//   JpsPaint                cic_mas_vec / cic_mas (+ TSC, PCS)   src/mas.py:5,88
//   JpsPowspec              powspec_vec                          src/correlations.py:7
//   JpsPowspecFundamental   powspec_vec_fundamental              src/correlations.py:60
//   JpsBispec               bispec                               src/correlations.py:334
//   JpsPaintPowspec         paint -> rho/mean-1 -> powspec_vec fused (tests/correlations.py:41-78)
//   JpsPaintGrad, JpsPowspecGrad   backward passes for jax.custom_vjp (tests/lognormal.py:99-107)
#if defined(JPS_WITH_JAX_FFI)

#include <cstdint>

#include <cuda_runtime_api.h>

#include "jps.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status(int rc) {
  if (rc != JPS_OK) return ffi::Error(ffi::ErrorCode::kInternal, jps_last_error());
  return ffi::Error::Success();
}

// cic_mas_vec(delta, x, y, z, w, ...) -> delta'.  Operand 0 is the mesh that is accumulated into
// (src/mas.py:89-153, Q5); the Python side passes input_output_aliases={0: 0}, so XLA normally hands
// the SAME buffer as operand 0 and as the result.  If it did not alias (the operand is still live
// elsewhere), the input mesh is copied into the result first -- the call stays functional either way.
static ffi::Error PaintImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> delta, ffi::Buffer<ffi::F32> x,
                            ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::F32> z, ffi::Buffer<ffi::F32> w,
                            ffi::Buffer<ffi::U8> scratch, float xmin, float ymin, float zmin,
                            float box_size, int32_t order, int32_t wrap, int32_t compat,
                            int32_t variant, ffi::ResultBuffer<ffi::F32> mesh) {
  const int n = static_cast<int>(mesh->dimensions()[0]);
  const int64_t np = static_cast<int64_t>(x.element_count());
  if (mesh->typed_data() != delta.typed_data()) {
    if (cudaMemcpyAsync(mesh->typed_data(), delta.typed_data(), delta.size_bytes(), cudaMemcpyDeviceToDevice,
                        stream) != cudaSuccess)
      return ffi::Error(ffi::ErrorCode::kInternal, "jps_paint: copy of the input mesh failed");
  }
  return status(jps_paint(n, x.typed_data(), y.typed_data(), z.typed_data(), w.typed_data(), /*stride=*/1, np,
                          xmin, ymin, zmin, box_size, order, wrap, compat, variant, JPS_PAINT_AUTO,
                          mesh->typed_data(), scratch.typed_data(), scratch.size_bytes(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsPaint, PaintImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // delta (operand 0, aliased to the result)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y
                                  .Arg<ffi::Buffer<ffi::F32>>()   // z
                                  .Arg<ffi::Buffer<ffi::F32>>()   // w
                                  .Arg<ffi::Buffer<ffi::U8>>()    // scratch (jps_paint_workspace_bytes)
                                  .Attr<float>("xmin").Attr<float>("ymin").Attr<float>("zmin")
                                  .Attr<float>("box_size").Attr<int32_t>("order").Attr<int32_t>("wrap")
                                  .Attr<int32_t>("compat").Attr<int32_t>("variant")
                                  .Ret<ffi::Buffer<ffi::F32>>());  // mesh

// powspec_vec(delta, box_size, k_edges) -> (k3D, Pk3D, Nmodes3D)   (src/correlations.py:8-56)
// The plan (cuFFT handles + partition of its workspace) is created once per (N, device, stream) by the
// Python wrapper and passed as an int64 attribute; k_edges is a static (host) attribute because bin
// membership is resolved on the host into integer thresholds.
static ffi::Error PowspecImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> delta, int64_t plan,
                              float box_size, ffi::Span<const float> k_edges, int32_t mas_order,
                              ffi::ResultBuffer<ffi::F32> k3d, ffi::ResultBuffer<ffi::F32> pk3d,
                              ffi::ResultBuffer<ffi::F32> nmodes) {
  return status(jps_powspec(reinterpret_cast<jps_plan_t*>(plan), delta.typed_data(), /*normalise=*/0, box_size,
                            k_edges.begin(), static_cast<int>(k_edges.size()) - 1, mas_order, /*shot_noise=*/0.f,
                            k3d->typed_data(), pk3d->typed_data(), nmodes->typed_data(), nullptr, nullptr, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsPowspec, PowspecImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan").Attr<float>("box_size")
                                  .Attr<ffi::Span<const float>>("k_edges").Attr<int32_t>("mas_order")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// powspec_vec_fundamental(delta, box_size) -> (k3D, Pk3D, Nmodes3D), kmax = jps_fundamental_nbins(N) rows
// (src/correlations.py:60-117; compat = 0 reproduces the k3D `.set` quirk Q18)
static ffi::Error PowspecFundamentalImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> delta, int64_t plan,
                                         float box_size, int32_t mas_order, int32_t compat,
                                         ffi::ResultBuffer<ffi::F32> k3d, ffi::ResultBuffer<ffi::F32> pk3d,
                                         ffi::ResultBuffer<ffi::F32> nmodes) {
  return status(jps_powspec_fundamental(reinterpret_cast<jps_plan_t*>(plan), delta.typed_data(), /*normalise=*/0,
                                        box_size, mas_order, compat, k3d->typed_data(), pk3d->typed_data(),
                                        nmodes->typed_data(), nullptr, nullptr, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsPowspecFundamental, PowspecFundamentalImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan").Attr<float>("box_size").Attr<int32_t>("mas_order")
                                  .Attr<int32_t>("compat")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// bispec(delta, box_size, k1, k2, theta) -> (k_all, Pk, B, Q)   (src/correlations.py:334-462; theta is
// returned unchanged by the Python wrapper).  theta is a static host attribute: the shell bounds are
// turned into integer k^2 thresholds on the host.  The plan needs n_shell_fields >= 6.
static ffi::Error BispecImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> delta, int64_t plan, float box_size,
                             float k1, float k2, ffi::Span<const float> theta, int32_t mas_order,
                             ffi::ResultBuffer<ffi::F32> k_all, ffi::ResultBuffer<ffi::F32> pk,
                             ffi::ResultBuffer<ffi::F32> B, ffi::ResultBuffer<ffi::F32> Q) {
  return status(jps_bispec(reinterpret_cast<jps_plan_t*>(plan), delta.typed_data(), /*normalise=*/0, box_size, k1,
                           k2, theta.begin(), static_cast<int>(theta.size()), mas_order, k_all->typed_data(),
                           pk->typed_data(), B->typed_data(), Q->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsBispec, BispecImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan").Attr<float>("box_size").Attr<float>("k1").Attr<float>("k2")
                                  .Attr<ffi::Span<const float>>("theta").Attr<int32_t>("mas_order")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());

// Fused particles -> mesh -> delta_k -> multipoles (what bench.py times): `mesh_scratch` is an XLA-owned
// float32 [N,N,N] operand used as the mesh, `scratch` the bucketing workspace.
static ffi::Error PaintPowspecImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y,
                                   ffi::Buffer<ffi::F32> z, ffi::Buffer<ffi::F32> w,
                                   ffi::Buffer<ffi::F32> mesh_scratch, ffi::Buffer<ffi::U8> scratch, int64_t plan,
                                   float xmin, float ymin, float zmin, float box_size, int32_t order,
                                   int32_t wrap, int32_t compat, ffi::Span<const float> k_edges, float shot_noise,
                                   ffi::ResultBuffer<ffi::F32> k3d, ffi::ResultBuffer<ffi::F32> pk3d,
                                   ffi::ResultBuffer<ffi::F32> nmodes) {
  return status(jps_paint_powspec(reinterpret_cast<jps_plan_t*>(plan), x.typed_data(), y.typed_data(), z.typed_data(),
                                  w.typed_data(), /*stride=*/1, static_cast<int64_t>(x.element_count()), xmin, ymin,
                                  zmin, box_size, order, wrap, compat, JPS_PAINT_AUTO, k_edges.begin(),
                                  static_cast<int>(k_edges.size()) - 1, shot_noise, mesh_scratch.typed_data(),
                                  scratch.typed_data(), scratch.size_bytes(), k3d->typed_data(), pk3d->typed_data(),
                                  nmodes->typed_data(), nullptr, nullptr, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsPaintPowspec, PaintPowspecImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y
                                  .Arg<ffi::Buffer<ffi::F32>>()   // z
                                  .Arg<ffi::Buffer<ffi::F32>>()   // w
                                  .Arg<ffi::Buffer<ffi::F32>>()   // mesh scratch [N,N,N]
                                  .Arg<ffi::Buffer<ffi::U8>>()    // bucketing scratch
                                  .Attr<int64_t>("plan")
                                  .Attr<float>("xmin").Attr<float>("ymin").Attr<float>("zmin")
                                  .Attr<float>("box_size").Attr<int32_t>("order").Attr<int32_t>("wrap")
                                  .Attr<int32_t>("compat").Attr<ffi::Span<const float>>("k_edges")
                                  .Attr<float>("shot_noise")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// ---- backward passes for jax.custom_vjp (the reference is differentiated with jax.value_and_grad through
// paint -> powspec_vec, tests/lognormal.py:99-107; INTEGRATION.md shows the Python wrappers)
// d loss / d (x, y, z, w) from d loss / d mesh
static ffi::Error PaintGradImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y,
                                ffi::Buffer<ffi::F32> z, ffi::Buffer<ffi::F32> w, ffi::Buffer<ffi::F32> grad_mesh,
                                float xmin, float ymin, float zmin, float box_size, int32_t order, int32_t wrap,
                                int32_t compat, int32_t variant, ffi::ResultBuffer<ffi::F32> gx,
                                ffi::ResultBuffer<ffi::F32> gy, ffi::ResultBuffer<ffi::F32> gz,
                                ffi::ResultBuffer<ffi::F32> gw) {
  const int n = static_cast<int>(grad_mesh.dimensions()[0]);
  return status(jps_paint_grad(n, x.typed_data(), y.typed_data(), z.typed_data(), w.typed_data(), /*stride=*/1,
                               static_cast<int64_t>(x.element_count()), xmin, ymin, zmin, box_size, order, wrap, compat,
                               variant, grad_mesh.typed_data(), gx->typed_data(), gy->typed_data(), gz->typed_data(),
                               gw->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsPaintGrad, PaintGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // grad_mesh
                                  .Attr<float>("xmin").Attr<float>("ymin").Attr<float>("zmin")
                                  .Attr<float>("box_size").Attr<int32_t>("order").Attr<int32_t>("wrap")
                                  .Attr<int32_t>("compat").Attr<int32_t>("variant")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());

// d loss / d delta from d loss / d Pk3D
static ffi::Error PowspecGradImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> delta, ffi::Buffer<ffi::F32> grad_pk,
                                  int64_t plan, float box_size, ffi::Span<const float> k_edges, int32_t mas_order,
                                  ffi::ResultBuffer<ffi::F32> grad_mesh) {
  return status(jps_powspec_grad(reinterpret_cast<jps_plan_t*>(plan), delta.typed_data(), /*normalise=*/0, box_size,
                                 k_edges.begin(), static_cast<int>(k_edges.size()) - 1, mas_order, grad_pk.typed_data(),
                                 grad_mesh->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsPowspecGrad, PowspecGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan").Attr<float>("box_size")
                                  .Attr<ffi::Span<const float>>("k_edges").Attr<int32_t>("mas_order")
                                  .Ret<ffi::Buffer<ffi::F32>>());

#endif  // JPS_WITH_JAX_FFI
