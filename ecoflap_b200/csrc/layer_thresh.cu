// A3+A5+A7 -- fused Wanda score / per-LAYER threshold select / in-place apply.
//
// Replaces  thres = torch.sort(W_metric.flatten())[0][int(numel * s)];  W[W_metric <= thres] = 0
// (LAVIS/lavis/compression/pruners/wanda_pruner.py:541,553-558; UPop wanda_pruner.py:502,512-517;
//  LLaMA/image_classifiers/prune_utils.py:28-31).
//
// The exact kth_index-th smallest fp32 score of the whole matrix is found with a 3-level radix
// select (11 + 10 + 10 key bits): each level is one pass that recomputes the score on the fly and
// histograms the digit of the keys that still match the prefix; the first pass reads W from HBM,
// the later ones and the apply pass are served from the 126 MB L2 (the largest matrix on this
// variant, ViT-g fc1/fc2, is 17 MB).  The score matrix is never materialised.
// Bound: HBM.  Algorithmic bytes per call: 2*R*C*sizeof(w) + 4*C (re-reads are L2 hits).
#include "common.cuh"
#include "radix_select.cuh"

namespace ecf {

constexpr int kLtBins = 2048;
constexpr int kLtThreads = 256;


template <int DT, bool ALIGNED>
__device__ __forceinline__ void lt_load_chunk(const char* wrow, int64_t c0, int64_t C, uint32_t (&raw)[8]) {
  // raw[j] = bit pattern of element j widened to 32 bits (fp32 bits, or 16-bit pattern in the low half)
  if (ALIGNED) {
    if constexpr (DT == ECF_F32) {
      const uint4 a = ldg_v4(wrow + c0 * 4), b = ldg_v4(wrow + c0 * 4 + 16);
      raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w; raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
    } else {
      const uint4 a = ldg_v4(wrow + c0 * 2);
      raw[0] = a.x & 0xffffu; raw[1] = a.x >> 16; raw[2] = a.y & 0xffffu; raw[3] = a.y >> 16;
      raw[4] = a.z & 0xffffu; raw[5] = a.z >> 16; raw[6] = a.w & 0xffffu; raw[7] = a.w >> 16;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c0 + j < C) {
        if constexpr (DT == ECF_F32)
          raw[j] = reinterpret_cast<const uint32_t*>(wrow)[c0 + j];
        else
          raw[j] = reinterpret_cast<const uint16_t*>(wrow)[c0 + j];
      } else {
        raw[j] = 0;
      }
    }
  }
}

template <int DT>
__device__ __forceinline__ float lt_to_float(uint32_t raw) {
  if constexpr (DT == ECF_F32) return __uint_as_float(raw);
  if constexpr (DT == ECF_BF16) return __uint_as_float(raw << 16);
  return __half2float(__ushort_as_half((unsigned short)raw));
}

// PASS 0: digit = key >> 20 (11 bits);  PASS 1: (key >> 10) & 1023 given prefix (11 bits);
// PASS 2: key & 1023 given prefix (21 bits)
template <int DT, bool ALIGNED, int PASS>
__global__ void __launch_bounds__(kLtThreads)
    lt_hist_kernel(const void* __restrict__ W, int64_t R, int64_t C, int64_t ld, const float* __restrict__ scaler_row,
                   const LtState* __restrict__ state, unsigned* __restrict__ hist) {
  __shared__ unsigned sh[kLtBins];
  for (int i = threadIdx.x; i < kLtBins; i += kLtThreads) sh[i] = 0;
  __syncthreads();
  const uint32_t prefix = PASS == 0 ? 0u : state->prefix;
  const int64_t nvec = (C + 7) / 8;
  const int64_t total = R * nvec;
  for (int64_t v = (int64_t)blockIdx.x * kLtThreads + threadIdx.x; v < total; v += (int64_t)gridDim.x * kLtThreads) {
    const int64_t row = v / nvec;
    const int64_t c0 = (v - row * nvec) * 8;
    const char* wrow = reinterpret_cast<const char*>(W) + row * ld * DType<DT>::kBytes;
    uint32_t raw[8];
    lt_load_chunk<DT, ALIGNED>(wrow, c0, C, raw);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c0 + j < C) {
        const uint32_t key = score_key(wanda_score(lt_to_float<DT>(raw[j]), sqrtf(scaler_row[c0 + j])));
        if (PASS == 0) {
          atomicAdd(&sh[key >> 20], 1u);
        } else if (PASS == 1) {
          if ((key >> 20) == prefix) atomicAdd(&sh[(key >> 10) & 1023u], 1u);
        } else {
          if ((key >> 10) == prefix) atomicAdd(&sh[key & 1023u], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kLtBins; i += kLtThreads) {
    const unsigned c = sh[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

template <int DT, bool ALIGNED>
__global__ void __launch_bounds__(kLtThreads)
    lt_apply_kernel(void* __restrict__ W, int64_t R, int64_t C, int64_t ld, const float* __restrict__ scaler_row,
                    const LtState* __restrict__ state, float* __restrict__ thres_out, uint8_t* __restrict__ mask_bits,
                    int64_t mask_ld, unsigned long long* __restrict__ n_zero) {
  const uint32_t tkey = state->prefix;
  if (blockIdx.x == 0 && threadIdx.x == 0 && thres_out != nullptr) *thres_out = __uint_as_float(tkey);
  const int64_t nvec = (C + 7) / 8;
  const int64_t total = R * nvec;
  int zeros = 0;
  for (int64_t v = (int64_t)blockIdx.x * kLtThreads + threadIdx.x; v < total; v += (int64_t)gridDim.x * kLtThreads) {
    const int64_t row = v / nvec;
    const int64_t c0 = (v - row * nvec) * 8;
    char* wrow = reinterpret_cast<char*>(W) + row * ld * DType<DT>::kBytes;
    uint32_t raw[8];
    lt_load_chunk<DT, ALIGNED>(wrow, c0, C, raw);
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c0 + j < C) {
        const uint32_t key = score_key(wanda_score(lt_to_float<DT>(raw[j]), sqrtf(scaler_row[c0 + j])));
        const bool p = key <= tkey;
        m |= (p ? 1u : 0u) << j;
        if (p) raw[j] = 0;
        const uint32_t absmask = DT == ECF_F32 ? 0x7fffffffu : 0x7fffu;
        zeros += ((raw[j] & absmask) == 0) ? 1 : 0;
      }
    }
    if (m) {
      if (ALIGNED) {
        if constexpr (DT == ECF_F32) {
          stg_v4(wrow + c0 * 4, make_uint4(raw[0], raw[1], raw[2], raw[3]));
          stg_v4(wrow + c0 * 4 + 16, make_uint4(raw[4], raw[5], raw[6], raw[7]));
        } else {
          stg_v4(wrow + c0 * 2, make_uint4(raw[0] | raw[1] << 16, raw[2] | raw[3] << 16, raw[4] | raw[5] << 16,
                                           raw[6] | raw[7] << 16));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < C && (m >> j & 1)) store_zero<DT>(wrow, c0 + j);
      }
    }
    if (mask_bits != nullptr) mask_bits[row * mask_ld + (c0 >> 3)] = (uint8_t)m;
  }
  if (n_zero != nullptr) {
    const int z = warp_sum(zeros);
    if ((threadIdx.x & 31) == 0 && z) atomicAdd(n_zero, (unsigned long long)z);
  }
}

size_t layer_thresh_workspace_bytes() { return 256 + kLtBins * sizeof(unsigned); }

template <int DT, bool ALIGNED>
static int run_layer_thresh(void* W, int64_t R, int64_t C, int64_t ld, const float* s, int64_t kth, float* thres_out,
                            uint8_t* mask, int64_t mask_ld, unsigned long long* nz, void* ws, cudaStream_t stream) {
  LtState* state = reinterpret_cast<LtState*>(ws);
  unsigned* hist = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + 256);
  ECF_CUDA_OK(cudaMemsetAsync(ws, 0, layer_thresh_workspace_bytes(), stream));
  const int64_t nvec = (C + 7) / 8;
  const int64_t total = R * nvec;
  int64_t want = (total + kLtThreads - 1) / kLtThreads;
  const int64_t cap = (int64_t)sm_count() * 8;
  const unsigned grid = (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
  lt_hist_kernel<DT, ALIGNED, 0><<<grid, kLtThreads, 0, stream>>>(W, R, C, ld, s, state, hist);
  lt_scan_kernel<11, true><<<1, 1024, 0, stream>>>(state, hist, (unsigned long long)kth);
  lt_hist_kernel<DT, ALIGNED, 1><<<grid, kLtThreads, 0, stream>>>(W, R, C, ld, s, state, hist);
  lt_scan_kernel<10, false><<<1, 1024, 0, stream>>>(state, hist, 0ull);
  lt_hist_kernel<DT, ALIGNED, 2><<<grid, kLtThreads, 0, stream>>>(W, R, C, ld, s, state, hist);
  lt_scan_kernel<10, false><<<1, 1024, 0, stream>>>(state, hist, 0ull);
  lt_apply_kernel<DT, ALIGNED><<<grid, kLtThreads, 0, stream>>>(W, R, C, ld, s, state, thres_out, mask, mask_ld, nz);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

template <int DT>
static int dispatch_lt(void* W, int64_t R, int64_t C, int64_t ld, const float* s, int64_t kth, float* thres_out,
                       uint8_t* mask, int64_t mask_ld, unsigned long long* nz, void* ws, cudaStream_t stream) {
  const int V = DType<DT>::kVec;
  const bool aligned = (C % 8 == 0) && (ld % V == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (aligned) return run_layer_thresh<DT, true>(W, R, C, ld, s, kth, thres_out, mask, mask_ld, nz, ws, stream);
  return run_layer_thresh<DT, false>(W, R, C, ld, s, kth, thres_out, mask, mask_ld, nz, ws, stream);
}

}  // namespace ecf

extern "C" int ecf_wanda_layer_thresh_apply(void* W, int w_dtype, int64_t R, int64_t C, int64_t ld,
                                            const float* scaler_row, int64_t kth_index, float* thres_out,
                                            uint8_t* mask_bits, int64_t mask_ld, unsigned long long* n_zero, void* ws,
                                            size_t ws_bytes, ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(W != nullptr && scaler_row != nullptr, ECF_ERR_INVALID, "layer_thresh: null pointer");
  ECF_REQUIRE(R > 0 && C > 0 && ld >= C, ECF_ERR_INVALID, "layer_thresh: bad shape R=%lld C=%lld ld=%lld",
              (long long)R, (long long)C, (long long)ld);
  // python indexing: sort(...)[idx] raises IndexError for idx >= numel; negative idx is never produced
  ECF_REQUIRE(kth_index >= 0 && kth_index < R * C, ECF_ERR_RANGE,
              "layer_thresh: kth_index %lld out of range for %lld elements (the reference raises IndexError)",
              (long long)kth_index, (long long)(R * C));
  ECF_REQUIRE(mask_bits == nullptr || mask_ld >= (C + 7) / 8, ECF_ERR_INVALID, "layer_thresh: mask_ld too small");
  ECF_REQUIRE(ws != nullptr && ws_bytes >= layer_thresh_workspace_bytes(), ECF_ERR_WORKSPACE,
              "layer_thresh: workspace %zu < %zu bytes", ws_bytes, layer_thresh_workspace_bytes());
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (w_dtype) {
    case ECF_F32:
      return dispatch_lt<ECF_F32>(W, R, C, ld, scaler_row, kth_index, thres_out, mask_bits, mask_ld, n_zero, ws, s);
    case ECF_F16:
      return dispatch_lt<ECF_F16>(W, R, C, ld, scaler_row, kth_index, thres_out, mask_bits, mask_ld, n_zero, ws, s);
    case ECF_BF16:
      return dispatch_lt<ECF_BF16>(W, R, C, ld, scaler_row, kth_index, thres_out, mask_bits, mask_ld, n_zero, ws, s);
  }
  set_error("layer_thresh: unknown dtype %d", w_dtype);
  return ECF_ERR_INVALID;
}
