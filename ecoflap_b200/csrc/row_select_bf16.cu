// Instantiations of the fast per-row Wanda select for ECF_BF16 weights (one translation unit per dtype: they compile in parallel).
#include "row_select_fast.cuh"

namespace ecf {
int row_select_fast_bf16(void* W, int64_t R, int64_t C, int64_t ld, const float* s, int64_t k, int nv_max, bool keep, uint8_t* mask,
                        int64_t mask_ld, unsigned long long* nz, cudaStream_t stream) {
  return run_row_select_fast<ECF_BF16>(W, R, C, ld, s, k, nv_max, keep, mask, mask_ld, nz, stream);
}
}  // namespace ecf
