// jax.ffi handlers that forward to libjps.so -- GATED: not part of the default build.
//
// north_star asks for jax.ffi custom calls; neither `jax` nor the XLA FFI headers
// (xla/ffi/api/ffi.h, shipped inside jaxlib: jax.ffi.include_dir()) exist in this image or on the
// GPU box, so this file cannot be compiled or exercised here.  It documents, in code, that the C
// ABI of include/jps.h is shaped so each handler is a 1:1 forward: XLA owns every buffer
// (inputs, outputs, scratch), hands us its cudaStream_t, and we never allocate or synchronise.
//
// Build where JAX is installed:
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include ffi/jax_ffi_shim.cc -L jax_powspec_b200 -ljps -o jps_jax_ffi.so
#if defined(JPS_WITH_JAX_FFI)

#include <cstdint>

#include "jps.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// cic_mas_vec(delta, x, y, z, w, ...) -> delta'   (input_output_aliases={0: 0} on the Python side,
// so `mesh` below is the accumulated-into buffer, src/mas.py:89-153)
static ffi::Error PaintImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y,
                            ffi::Buffer<ffi::F32> z, ffi::Buffer<ffi::F32> w,
                            ffi::Buffer<ffi::U8> scratch, float xmin, float ymin, float zmin,
                            float box_size, int32_t order, int32_t wrap, int32_t compat,
                            int32_t variant, ffi::ResultBuffer<ffi::F32> mesh) {
  const int n = static_cast<int>(mesh->dimensions()[0]);
  const int64_t np = static_cast<int64_t>(x.element_count());
  int rc = jps_paint(n, x.typed_data(), y.typed_data(), z.typed_data(), w.typed_data(), /*stride=*/1, np,
                     xmin, ymin, zmin, box_size, order, wrap, compat, variant, JPS_PAINT_AUTO,
                     mesh->typed_data(), scratch.typed_data(), scratch.size_bytes(), stream);
  if (rc != JPS_OK) return ffi::Error(ffi::ErrorCode::kInternal, jps_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsPaint, PaintImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x
                                  .Arg<ffi::Buffer<ffi::F32>>()   // y
                                  .Arg<ffi::Buffer<ffi::F32>>()   // z
                                  .Arg<ffi::Buffer<ffi::F32>>()   // w
                                  .Arg<ffi::Buffer<ffi::U8>>()    // scratch (jps_paint_workspace_bytes)
                                  .Attr<float>("xmin").Attr<float>("ymin").Attr<float>("zmin")
                                  .Attr<float>("box_size").Attr<int32_t>("order").Attr<int32_t>("wrap")
                                  .Attr<int32_t>("compat").Attr<int32_t>("variant")
                                  .Ret<ffi::Buffer<ffi::F32>>());  // mesh (aliased to operand 0)

// powspec_vec(delta, box_size, k_edges) -> (k3D, Pk3D, Nmodes3D)   (src/correlations.py:8-56)
// The plan (cuFFT handles + partition of `plan_ws`) is created once per (N, device) by the Python
// wrapper and passed as an int64 attribute; k_edges is a static (host) attribute because bin
// membership is resolved on the host into integer thresholds.
static ffi::Error PowspecImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> delta, int64_t plan,
                              float box_size, ffi::Span<const float> k_edges, int32_t mas_order,
                              ffi::ResultBuffer<ffi::F32> k3d, ffi::ResultBuffer<ffi::F32> pk3d,
                              ffi::ResultBuffer<ffi::F32> nmodes) {
  int rc = jps_powspec(reinterpret_cast<jps_plan_t*>(plan), delta.typed_data(), /*normalise=*/0, box_size,
                       k_edges.begin(), static_cast<int>(k_edges.size()) - 1, mas_order, /*shot_noise=*/0.f,
                       k3d->typed_data(), pk3d->typed_data(), nmodes->typed_data(), nullptr, nullptr, stream);
  if (rc != JPS_OK) return ffi::Error(ffi::ErrorCode::kInternal, jps_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JpsPowspec, PowspecImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan").Attr<float>("box_size")
                                  .Attr<ffi::Span<const float>>("k_edges").Attr<int32_t>("mas_order")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

#endif  // JPS_WITH_JAX_FFI
