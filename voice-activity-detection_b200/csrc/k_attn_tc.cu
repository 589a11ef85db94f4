// placeholder: replaced by the tcgen05 kernel
#include "vadb_common.cuh"
namespace vadb {
cudaError_t launch_attn_tc(const bf16*, const bf16*, const bf16*, bf16*, const int32_t*, int, int,
                           cudaStream_t, std::string* err) {
  if (err) *err = "tcgen05 attention kernel not built";
  return cudaErrorNotSupported;
}
}  // namespace vadb
