// Instantiations of the fast per-row Wanda select for ECF_F16 weights (one translation unit per dtype: they compile in parallel).
#include "row_select_fast.cuh"

namespace ecf {
int row_select_fast_f16(RfBatch& tb, int nv_max, bool keep, cudaStream_t stream) {
  return run_row_select_fast<ECF_F16>(tb, nv_max, keep, stream);
}
}  // namespace ecf
