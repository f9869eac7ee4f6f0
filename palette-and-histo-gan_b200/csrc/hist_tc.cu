// tcgen05 engine placeholder (filled in by the tensor-core milestone).
#include "common.cuh"
#include "hist_internal.cuh"

namespace ph {
bool tc_supported(int64_t, int, int) { return false; }
size_t tc_workspace_bytes(int64_t, int64_t, int) { return 0; }
int tc_hist_forward(const float*, int64_t, int64_t, int, const float*, int, int, float, float, float*, float*,
                    void*, cudaStream_t) {
  set_error("tensor-core engine not built");
  return PH_ERR_UNSUPPORTED;
}
int tc_hist_backward(const float*, int64_t, int64_t, int, const float*, int, int, float, float, const float*,
                     const float*, const float*, const float*, const double*, int64_t, const float*, float*, void*,
                     cudaStream_t) {
  set_error("tensor-core engine not built");
  return PH_ERR_UNSUPPORTED;
}
}  // namespace ph
