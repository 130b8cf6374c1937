// A11 zeroth-order perturbation, A14 segmented |W| / W^2 reduction, A17 zero count.
// All three are single-pass HBM-bound streams with 128-bit accesses.
#include "common.cuh"

namespace ecf {

// ------------------------------------------------------------------------------------------------
// A11  w = rn(w + rn(rn(scaling * z) * eps))     layer_single_base_pruner.py:473-486
// torch evaluates `param.data + scaling_factor * z * zo_eps` as three separate elementwise kernels,
// each rounding to the parameter dtype, so the three roundings are reproduced explicitly
// (__fmul_rn/__fadd_rn forbid FMA contraction in the fp32 case).
// Algorithmic bytes: numel * (2*sizeof(w) + sizeof(z)).
// ------------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ float round_dt(float x) {
  if constexpr (DT == ECF_F32) return x;
  if constexpr (DT == ECF_F16) return __half2float(__float2half_rn(x));
  return __bfloat162float(__float2bfloat16_rn(x));
}

template <int DT>
__device__ __forceinline__ float zo_one(float w, float z, float scaling, float eps) {
  const float t1 = round_dt<DT>(__fmul_rn(z, scaling));
  const float t2 = round_dt<DT>(__fmul_rn(t1, eps));
  return round_dt<DT>(__fadd_rn(w, t2));
}

template <int DT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (DT == ECF_F16) {
    return (uint32_t)__half_as_ushort(__float2half_rn(lo)) | ((uint32_t)__half_as_ushort(__float2half_rn(hi)) << 16);
  } else {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) |
           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
  }
}

template <int DT>
__global__ void __launch_bounds__(256) zo_perturb_kernel(void* __restrict__ W, const void* __restrict__ Z, int64_t n,
                                                         float scaling, float eps, bool vec) {
  constexpr int V = DType<DT>::kVec;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const int64_t nv = n / V;
    for (int64_t i = tid; i < nv; i += stride) {
      uint4 w = ldg_v4(reinterpret_cast<const char*>(W) + i * 16);
      const uint4 z = ldg_stream(reinterpret_cast<const char*>(Z) + i * 16);
      uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      const uint32_t zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if constexpr (DT == ECF_F32) {
          ww[j] = __float_as_uint(zo_one<DT>(__uint_as_float(ww[j]), __uint_as_float(zz[j]), scaling, eps));
        } else {
          float wl, wh, zl, zh;
          unpack2<DT>(ww[j], wl, wh);
          unpack2<DT>(zz[j], zl, zh);
          ww[j] = pack2<DT>(zo_one<DT>(wl, zl, scaling, eps), zo_one<DT>(wh, zh, scaling, eps));
        }
      }
      stg_v4(reinterpret_cast<char*>(W) + i * 16, make_uint4(ww[0], ww[1], ww[2], ww[3]));
    }
    for (int64_t i = nv * V + tid; i < n; i += stride) {
      const float r = zo_one<DT>(load_elem<DT>(W, i), load_elem<DT>(Z, i), scaling, eps);
      if constexpr (DT == ECF_F32) reinterpret_cast<float*>(W)[i] = r;
      if constexpr (DT == ECF_F16) reinterpret_cast<__half*>(W)[i] = __float2half_rn(r);
      if constexpr (DT == ECF_BF16) reinterpret_cast<__nv_bfloat16*>(W)[i] = __float2bfloat16_rn(r);
    }
  } else {
    for (int64_t i = tid; i < n; i += stride) {
      const float r = zo_one<DT>(load_elem<DT>(W, i), load_elem<DT>(Z, i), scaling, eps);
      if constexpr (DT == ECF_F32) reinterpret_cast<float*>(W)[i] = r;
      if constexpr (DT == ECF_F16) reinterpret_cast<__half*>(W)[i] = __float2half_rn(r);
      if constexpr (DT == ECF_BF16) reinterpret_cast<__nv_bfloat16*>(W)[i] = __float2bfloat16_rn(r);
    }
  }
}

static unsigned stream_grid(int64_t work_items, int threads) {
  int64_t want = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (unsigned)want;
}

// ------------------------------------------------------------------------------------------------
// A17  count of zero-valued elements (either sign of zero, like torch's W == 0)
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) count_zero_kernel(const void* __restrict__ W, int64_t n, bool vec,
                                                         unsigned long long* __restrict__ out) {
  constexpr int V = DType<DT>::kVec;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int z = 0;
  int64_t done = 0;
  if (vec) {
    const int64_t nv = n / V;
    for (int64_t i = tid; i < nv; i += stride) {
      const uint4 w = ldg_stream(reinterpret_cast<const char*>(W) + i * 16);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if constexpr (DT == ECF_F32) {
          z += ((ww[j] & 0x7fffffffu) == 0);
        } else {
          z += ((ww[j] & 0x00007fffu) == 0) + ((ww[j] & 0x7fff0000u) == 0);
        }
      }
    }
    done = nv * V;
  }
  for (int64_t i = done + tid; i < n; i += stride) z += (load_elem<DT>(W, i) == 0.f);
  z = warp_sum(z);
  __shared__ int ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = z;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < 8; ++i) t += ws[i];
    if (t) atomicAdd(out, t);
  }
}

// ------------------------------------------------------------------------------------------------
// A14  segmented reduction over a device table of tensors: sum|w| and sum w^2 per tensor.
// One CTA per 32768-element chunk; fp32 per-thread partials (128 elements), fp32 block tree, and a
// last-CTA ticket that combines the chunk partials of every tensor in chunk order in fp64, so the
// result is deterministic.  Algorithmic bytes: sum numel * sizeof(w).
// ------------------------------------------------------------------------------------------------
constexpr int64_t kGrChunk = 32768;
constexpr int kGrThreads = 256;

__device__ __forceinline__ void gr_acc(float f, float& sa, float& sq) {
  sa += fabsf(f);
  sq = fmaf(f, f, sq);
}

template <int DT>
__device__ __forceinline__ void gr_chunk(const void* base, int64_t begin, int64_t end, float& sa, float& sq) {
  constexpr int V = DType<DT>::kVec;
  const char* p = reinterpret_cast<const char*>(base);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) + begin * DType<DT>::kBytes) & 15) == 0;
  int64_t i = begin;
  if (vec) {
    const int64_t nv = (end - begin) / V;
    for (int64_t v = threadIdx.x; v < nv; v += kGrThreads) {
      const uint4 w = ldg_stream(p + (begin + v * V) * DType<DT>::kBytes);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if constexpr (DT == ECF_F32) {
          gr_acc(__uint_as_float(ww[j]), sa, sq);
        } else {
          float lo, hi;
          unpack2<DT>(ww[j], lo, hi);
          gr_acc(lo, sa, sq);
          gr_acc(hi, sa, sq);
        }
      }
    }
    i = begin + nv * V;
  }
  for (int64_t e = i + threadIdx.x; e < end; e += kGrThreads) gr_acc(load_elem<DT>(base, e), sa, sq);
}

__global__ void __launch_bounds__(kGrThreads)
    group_reduce_kernel(const ecf_tensor_desc* __restrict__ table, int n, int64_t total_chunks,
                        float2* __restrict__ partial, unsigned* __restrict__ ticket, double* __restrict__ sum_abs,
                        double* __restrict__ sum_sq) {
  __shared__ int s_tensor;
  __shared__ float s_a[8], s_q[8];
  __shared__ bool s_last;
  const int64_t chunk = blockIdx.x;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n - 1;  // last tensor whose chunk_begin <= chunk
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid].chunk_begin <= chunk) lo = mid; else hi = mid - 1;
    }
    s_tensor = lo;
  }
  __syncthreads();
  const ecf_tensor_desc d = table[s_tensor];
  const int64_t begin = (chunk - d.chunk_begin) * kGrChunk;
  const int64_t end = min(d.numel, begin + kGrChunk);
  float sa = 0.f, sq = 0.f;
  if (begin < end) {
    if (d.dtype == ECF_F32) gr_chunk<ECF_F32>(d.ptr, begin, end, sa, sq);
    else if (d.dtype == ECF_F16) gr_chunk<ECF_F16>(d.ptr, begin, end, sa, sq);
    else gr_chunk<ECF_BF16>(d.ptr, begin, end, sa, sq);
  }
  sa = warp_sum(sa);
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) {
    s_a[threadIdx.x >> 5] = sa;
    s_q[threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, q = 0.f;
    for (int i = 0; i < 8; ++i) { a += s_a[i]; q += s_q[i]; }
    partial[chunk] = make_float2(a, q);
    __threadfence();
    const unsigned prev = atomicAdd(ticket, 1u);
    s_last = (prev == (unsigned)(total_chunks - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // final combine: one warp per tensor, chunk order, fp64
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t = wid; t < n; t += kGrThreads / 32) {
    const int64_t cb = table[t].chunk_begin;
    const int64_t ce = (t + 1 < n) ? table[t + 1].chunk_begin : total_chunks;
    double a = 0.0, q = 0.0;
    for (int64_t c = cb + lane; c < ce; c += 32) {
      const float2 p = __ldcg(&partial[c]);
      a += (double)p.x;
      q += (double)p.y;
    }
    a = warp_sum(a);
    q = warp_sum(q);
    if (lane == 0) {
      sum_abs[t] = a;
      sum_sq[t] = q;
    }
  }
  if (threadIdx.x == 0) *ticket = 0;
}

size_t group_reduce_workspace_bytes(int64_t total_chunks) { return 256 + (size_t)total_chunks * sizeof(float2); }

}  // namespace ecf

extern "C" {

int ecf_zo_perturb(void* W, int w_dtype, int64_t numel, const void* z, double scaling, double eps,
                   ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(W != nullptr && z != nullptr && numel >= 0, ECF_ERR_INVALID, "zo_perturb: bad arguments");
  if (numel == 0) return ECF_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(z)) & 15) == 0;
  const int V = w_dtype == ECF_F32 ? 4 : 8;
  const unsigned grid = stream_grid((numel + V - 1) / V, 256);
  const float fs = (float)scaling, fe = (float)eps;
  switch (w_dtype) {
    case ECF_F32: zo_perturb_kernel<ECF_F32><<<grid, 256, 0, s>>>(W, z, numel, fs, fe, vec); break;
    case ECF_F16: zo_perturb_kernel<ECF_F16><<<grid, 256, 0, s>>>(W, z, numel, fs, fe, vec); break;
    case ECF_BF16: zo_perturb_kernel<ECF_BF16><<<grid, 256, 0, s>>>(W, z, numel, fs, fe, vec); break;
    default: set_error("zo_perturb: unknown dtype %d", w_dtype); return ECF_ERR_INVALID;
  }
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

int ecf_count_zero(const void* W, int w_dtype, int64_t numel, unsigned long long* n_zero, ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(W != nullptr && n_zero != nullptr && numel >= 0, ECF_ERR_INVALID, "count_zero: bad arguments");
  if (numel == 0) return ECF_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = (reinterpret_cast<uintptr_t>(W) & 15) == 0;
  const int V = w_dtype == ECF_F32 ? 4 : 8;
  const unsigned grid = stream_grid((numel + V - 1) / V, 256);
  switch (w_dtype) {
    case ECF_F32: count_zero_kernel<ECF_F32><<<grid, 256, 0, s>>>(W, numel, vec, n_zero); break;
    case ECF_F16: count_zero_kernel<ECF_F16><<<grid, 256, 0, s>>>(W, numel, vec, n_zero); break;
    case ECF_BF16: count_zero_kernel<ECF_BF16><<<grid, 256, 0, s>>>(W, numel, vec, n_zero); break;
    default: set_error("count_zero: unknown dtype %d", w_dtype); return ECF_ERR_INVALID;
  }
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

int64_t ecf_group_reduce_chunk_elems(void) { return ecf::kGrChunk; }

int ecf_group_abs_reduce(const ecf_tensor_desc* d_table, int n_tensors, int64_t total_chunks, double* sum_abs,
                         double* sum_sq, void* ws, size_t ws_bytes, ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(d_table != nullptr && sum_abs != nullptr && sum_sq != nullptr, ECF_ERR_INVALID,
              "group_reduce: null pointer");
  ECF_REQUIRE(n_tensors > 0 && total_chunks > 0 && total_chunks < (1ll << 31), ECF_ERR_INVALID,
              "group_reduce: bad table size n=%d chunks=%lld", n_tensors, (long long)total_chunks);
  const size_t need = group_reduce_workspace_bytes(total_chunks);
  ECF_REQUIRE(ws != nullptr && ws_bytes >= need, ECF_ERR_WORKSPACE, "group_reduce: workspace %zu < %zu bytes",
              ws_bytes, need);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  unsigned* ticket = reinterpret_cast<unsigned*>(ws);
  float2* partial = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) + 256);
  ECF_CUDA_OK(cudaMemsetAsync(ticket, 0, sizeof(unsigned), s));
  group_reduce_kernel<<<(unsigned)total_chunks, kGrThreads, 0, s>>>(d_table, n_tensors, total_chunks, partial, ticket,
                                                                    sum_abs, sum_sq);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

}  // extern "C"
