// placeholder until the OBS kernels land
#include "common.cuh"
namespace ecf { size_t obs_workspace_bytes(int64_t, int64_t) { return 256; } }
extern "C" int ecf_obs_prune(float*, int64_t, int64_t, int64_t, const float*, int64_t, const int64_t*, int, void*, size_t, ecf_stream_t) {
  ecf::set_error("obs_prune: not implemented yet");
  return ECF_ERR_INVALID;
}
