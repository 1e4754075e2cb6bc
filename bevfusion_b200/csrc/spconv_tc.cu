// tcgen05 tensor-core sparse convolution (placeholder until the UMMA path lands).
#include "common.cuh"
namespace bevb200 {
int spconv_forward_tc(const float *, const float *, const int32_t *, int, int, int, int, int,
                      const float *, const float *, const float *, int, int, float *, cudaStream_t) {
  snprintf(g_last_error, sizeof(g_last_error), "spconv_forward: tensor-core path not built");
  return BEVB200_EUNSUPPORTED;
}
}  // namespace bevb200
