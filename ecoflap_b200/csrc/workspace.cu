// Workspace-size queries (one place so the Python side and the kernels cannot disagree).
#include "common.cuh"

namespace ecf {
size_t sqnorm_workspace_bytes(int64_t T, int64_t C);
size_t layer_thresh_workspace_bytes(int64_t R, int64_t C);
size_t group_reduce_workspace_bytes(int64_t total_chunks);
size_t hessian_workspace_bytes(int64_t T, int64_t C);
size_t obs_workspace_bytes(int64_t R, int64_t C);
size_t global_select_workspace_bytes(int64_t nseg);
}  // namespace ecf

extern "C" size_t ecf_workspace_bytes(int op, int64_t R, int64_t C) {
  using namespace ecf;
  if (R < 0 || C < 0) return 0;
  switch (op) {
    case ECF_OP_SQNORM: return sqnorm_workspace_bytes(R, C);
    case ECF_OP_ROW_SELECT: return 256;  // none needed; a non-zero size keeps callers uniform
    case ECF_OP_LAYER_THRESH: return layer_thresh_workspace_bytes(R, C);
    case ECF_OP_GROUP_REDUCE: return group_reduce_workspace_bytes(C);
    case ECF_OP_HESSIAN: return hessian_workspace_bytes(R, C);
    case ECF_OP_OBS: return obs_workspace_bytes(R, C);
    case ECF_OP_GLOBAL_SELECT: return global_select_workspace_bytes(R);
  }
  return 0;
}
