// C-ABI plumbing: version, thread-local error string, device query.
#include <cstring>

#include "common.cuh"

namespace ecf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_dev_state = 0;  // 0 unknown, 1 ok, -1 none
static int g_sms = 0;

int check_device() {
  if (g_dev_state == 1) return ECF_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device visible (ecoflap_b200 has no CPU fallback)");
    return ECF_ERR_NO_DEVICE;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    set_error("cudaGetDeviceProperties failed");
    return ECF_ERR_CUDA;
  }
  if (p.major != 10) {
    set_error("device '%s' is sm_%d%d; this library is built for sm_100a only", p.name, p.major, p.minor);
    return ECF_ERR_NO_DEVICE;
  }
  g_sms = p.multiProcessorCount;
  g_dev_state = 1;
  return ECF_OK;
}

int sm_count() { return g_sms > 0 ? g_sms : 148; }

}  // namespace ecf

extern "C" {

int ecf_version(void) { return ECF_ABI_VERSION; }

const char* ecf_last_error(void) { return ecf::g_err; }

int ecf_device_sm_count(void) {
  int st = ecf::check_device();
  if (st != ECF_OK) return st;
  return ecf::sm_count();
}

}  // extern "C"
