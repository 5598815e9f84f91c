// Bucketed painter (K1 + K2) -- placeholder until the tile kernels land.
#include "paint_common.cuh"

namespace jps {

size_t paint_sorted_workspace(int n, int64_t n_part, int order) {
  (void)n; (void)order;
  return align_up((size_t)n_part * 16, 256) + ((size_t)1 << 20);
}

int paint_sorted(const PaintParams& p, int order, int compat, void* ws, size_t ws_bytes,
                 cudaStream_t s) {
  (void)p; (void)order; (void)compat; (void)ws; (void)ws_bytes; (void)s;
  set_error("jps_paint: JPS_PAINT_SORTED is not built yet");
  return JPS_ERR_UNSUPPORTED;
}

}  // namespace jps
