// Shared device/host helpers for the ecoflap_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/ecoflap_b200.h"

namespace ecf {

// ------------------------------------------------------------------ error plumbing (host)
void set_error(const char* fmt, ...);
int check_device();  // ECF_OK or ECF_ERR_NO_DEVICE (cached)
int sm_count();

#define ECF_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::ecf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                      \
      return ECF_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define ECF_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::ecf::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------ dtype traits (device)
template <int DT>
struct DType;
template <>
struct DType<ECF_F32> {
  using T = float;
  static constexpr int kBytes = 4;
  static constexpr int kVec = 4;  // elements per 16-byte vector
};
template <>
struct DType<ECF_F16> {
  using T = __half;
  static constexpr int kBytes = 2;
  static constexpr int kVec = 8;
};
template <>
struct DType<ECF_BF16> {
  using T = __nv_bfloat16;
  static constexpr int kBytes = 2;
  static constexpr int kVec = 8;
};

__host__ __device__ inline int dtype_bytes(int dt) { return dt == ECF_F32 ? 4 : 2; }

// 16-byte streaming load (read-once data: do not pollute L1)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// 16-byte coherent load that bypasses L1 allocation: streaming data the same kernel later overwrites in place, read next
// to small tables (q = sqrt(scaler_row)) that must stay L1 resident
__device__ __forceinline__ uint4 ldg_noalloc(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
// 16-byte load that may be re-read by the same kernel / a follow-up pass (keep in L2)
__device__ __forceinline__ uint4 ldg_v4(const void* p) {
  return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// Unpack one 32-bit word holding two 16-bit floats into two fp32 values (lo = element 0).
template <int DT>
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi) {
  if constexpr (DT == ECF_BF16) {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
  } else {
    __half2 h = *reinterpret_cast<__half2*>(&w);
    float2 f = __half22float2(h);
    lo = f.x;
    hi = f.y;
  }
}

// Scalar element load as fp32 (slow/generic paths only).
template <int DT>
__device__ __forceinline__ float load_elem(const void* base, int64_t i) {
  if constexpr (DT == ECF_F32) return reinterpret_cast<const float*>(base)[i];
  if constexpr (DT == ECF_F16) return __half2float(reinterpret_cast<const __half*>(base)[i]);
  if constexpr (DT == ECF_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
  return 0.f;
}
template <int DT>
__device__ __forceinline__ void store_zero(void* base, int64_t i) {
  if constexpr (DT == ECF_F32)
    reinterpret_cast<float*>(base)[i] = 0.f;
  else
    reinterpret_cast<uint16_t*>(base)[i] = 0;
}

// Wanda score, bit-exact with torch: fp32(|w|) * sqrtf(s).  The multiply must stay a lone
// IEEE fp32 multiply (no FMA contraction) -- __fmul_rn guarantees that.
__device__ __forceinline__ float wanda_score(float w, float sqrt_s) { return __fmul_rn(fabsf(w), sqrt_s); }

// Order-preserving key of a score.  Scores are >= +0 or NaN; NaN sorts last like torch.sort.
__device__ __forceinline__ uint32_t score_key(float s) {
  uint32_t u = __float_as_uint(s) & 0x7fffffffu;  // -0 -> +0 (equal under torch's comparison)
  return u > 0x7f800000u ? 0x7fffffffu : u;       // every NaN -> one maximal key
}

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace ecf
