// Instantiations of the bulk-copy per-row Wanda select for ECF_BF16 weights (one translation unit per dtype: they compile in parallel).
#include "row_select_tma.cuh"

namespace ecf {
int row_select_tma_bf16(RfBatch& tb, int stages, int share, size_t pad_smem, cudaStream_t stream) {
  return run_row_select_tma<ECF_BF16>(tb, stages, share, pad_smem, stream);
}
}  // namespace ecf
