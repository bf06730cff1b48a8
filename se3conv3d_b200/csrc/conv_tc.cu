// bf16 tensor-core path (precision 1) -- placeholder until the mma.sync / tcgen05 kernels land.
#include "common.cuh"
namespace se3 {
size_t conv_tc_fwd_workspace_bytes(const se3_conv_desc*) { return 0; }
size_t conv_tc_bwd_workspace_bytes(const se3_conv_desc*) { return 0; }
size_t conv_tc_saved_bytes(const se3_conv_desc*) { return 0; }
int conv_tc_fwd(const se3_conv_desc*, const float*, float*, void*, void*, size_t, cudaStream_t) {
  set_error("precision 1 (tensor-core path) is not built in this library");
  return SE3_EINVAL;
}
int conv_tc_bwd(const se3_conv_desc*, const float*, const float*, const void*, float*, float*, float*, float*, void*,
                size_t, cudaStream_t) {
  set_error("precision 1 (tensor-core path) is not built in this library");
  return SE3_EINVAL;
}
}  // namespace se3
