// N3 -- global-pruner baselines: one threshold over the scores of MANY tensors (or one per tensor), then  w *= (score > thr).
//
// Replaces  BLIPT5GlobalPruner.get_mask / get_layerwise_mask  (LAVIS/lavis/compression/pruners/global_pruner.py:116-157)
// and LayerSparsity.get_mask (layer_single_base_pruner.py:156-181, the 'Real*' ratio oracle):
//     all_scores = cat(flatten(score_t));  thr = topk(all_scores, int(p * numel), largest=False)[-1]
//     mask_t = score_t > thr;  w_t *= mask_t
// with the optional per-tensor protection of the  int(numel_t * (1 - max_sparsity))  largest scores (set to finfo.max
// before the global top-k).  The reference materialises every score tensor on the CPU in fp32 (3.7 G elements for BLIP-2)
// and runs torch.topk over their concatenation; here the scores are recomputed on the fly from W (and the accumulated
// |grad| sums) inside a three-digit radix select (11 + 11 + 10 bits) over a device table of tensors -- nothing is
// materialised, every pass streams the weights once.
//
// Score modes (what the reference's compute_importance_scores variants produce per element):
//   MAG           float(w)                       -- SIGNED, as the reference has it (global_pruner.py:249-251: no abs)
//   GRAD_MAG_ABS  |float(w)| * |G / nb|          -- global_pruner.py:256-300, layer_single_base_pruner.py:466
//   GRAD_MAG_SQ   float(w)^2 * (G / nb)          -- layer_single_base_pruner.py:464  (G accumulates g^2)
//   GRAD_ONLY     |G / nb|                       -- layer_single_base_pruner.py:468
// G: fp32 per-element sum over the batches of |grad| (or grad^2), nb: number of batches; G / nb is one IEEE fp32 division
// and every product one fp32 multiply, like the torch expressions.
// Keys: order-preserving uint32 image of the fp32 score (sign-magnitude -> biased), -0 == +0, NaN last (torch.topk order).
// Segmented mode: one select per tensor (get_layerwise_mask, the protection thresholds) with one histogram per tensor.
// Bound: HBM for the apply and the two refinement passes; the first histogram pass is bound by shared-memory atomics
// (this is a baseline the reference runs in minutes, not the hot path).
#include "common.cuh"
#include "radix_select.cuh"

namespace ecf {

constexpr int64_t kGsChunk = 32768;
constexpr int kGsThreads = 256;
constexpr uint32_t kGsProtected = 0xff7fffffu;  // key of finfo(float32).max
constexpr uint32_t kGsNaN = 0xffffffffu;

__device__ __forceinline__ uint32_t gs_key(float s) {
  uint32_t u = __float_as_uint(s);
  if ((u << 1) == 0u) u = 0u;                          // -0 == +0
  if ((u & 0x7fffffffu) > 0x7f800000u) return kGsNaN;  // NaN sorts last
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int MODE>
__device__ __forceinline__ float gs_score(float w, float g, float nb) {
  if constexpr (MODE == ECF_GLOBAL_MAG) return w;
  else if constexpr (MODE == ECF_GLOBAL_GRAD_MAG_ABS) return __fmul_rn(fabsf(w), fabsf(__fdiv_rn(g, nb)));
  else if constexpr (MODE == ECF_GLOBAL_GRAD_MAG_SQ) return __fmul_rn(__fmul_rn(w, w), __fdiv_rn(g, nb));
  else return fabsf(__fdiv_rn(g, nb));
}

__device__ __forceinline__ float gs_load(const void* W, int dtype, int64_t i) {
  if (dtype == ECF_F32) return reinterpret_cast<const float*>(W)[i];
  if (dtype == ECF_F16) return __half2float(reinterpret_cast<const __half*>(W)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(W)[i]);
}

__device__ __forceinline__ int gs_find_tensor(const ecf_global_desc* table, int n, int64_t chunk) {
  int lo = 0, hi = n - 1;  // last tensor whose chunk_begin <= chunk
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].chunk_begin <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <int MODE>
__device__ __forceinline__ uint32_t gs_elem_key(const ecf_global_desc& d, int64_t i, float nb, uint32_t prot) {
  const float w = gs_load(d.W, d.dtype, i);
  const float g = MODE == ECF_GLOBAL_MAG ? 0.f : d.G[i];
  const uint32_t key = gs_key(gs_score<MODE>(w, g, nb));
  return (key >= prot && key != kGsNaN) ? kGsProtected : key;
}

// PASS 0: digit key >> 21; PASS 1: (key >> 10) & 2047 inside prefix; PASS 2: key & 1023 inside prefix
template <int MODE, int PASS>
__global__ void __launch_bounds__(kGsThreads)
    gs_hist_kernel(const ecf_global_desc* __restrict__ table, int n, float nb, int segmented, const uint32_t* __restrict__ prot,
                   const LtState* __restrict__ state, unsigned* __restrict__ hist) {
  __shared__ unsigned sh[2048];
  __shared__ int s_tensor;
  for (int i = threadIdx.x; i < 2048; i += kGsThreads) sh[i] = 0u;
  if (threadIdx.x == 0) s_tensor = gs_find_tensor(table, n, blockIdx.x);
  __syncthreads();
  const int t = s_tensor;
  const ecf_global_desc d = table[t];
  const int seg = segmented ? t : 0;
  const uint32_t prefix = PASS == 0 ? 0u : state[seg].prefix;
  const uint32_t pk = prot != nullptr ? prot[t] : 0xffffffffu;
  const int64_t begin = ((int64_t)blockIdx.x - d.chunk_begin) * kGsChunk;
  const int64_t end = min(d.numel, begin + kGsChunk);
  for (int64_t i = begin + threadIdx.x; i < end; i += kGsThreads) {
    const uint32_t key = gs_elem_key<MODE>(d, i, nb, pk);
    if (PASS == 0) atomicAdd(&sh[key >> 21], 1u);
    else if (PASS == 1) { if ((key >> 21) == prefix) atomicAdd(&sh[(key >> 10) & 2047u], 1u); }
    else { if ((key >> 10) == prefix) atomicAdd(&sh[key & 1023u], 1u); }
  }
  __syncthreads();
  unsigned* gh = hist + (size_t)seg * 2048;
  for (int i = threadIdx.x; i < 2048; i += kGsThreads) {
    const unsigned c = sh[i];
    if (c) atomicAdd(&gh[i], c);
  }
}

// one CTA per segment: the bin holding rank rem, extend the prefix, clear the segment's histogram (64-bit running sums:
// a global select ranks billions of elements, one bin may hold more than 2^32 only if numel does -- rejected on the host)
template <int BITS, bool FIRST>
__global__ void __launch_bounds__(1024) gs_scan_kernel(LtState* state, unsigned* hist, const long long* __restrict__ ranks) {
  constexpr int NB = 1 << BITS;
  __shared__ unsigned long long warp_tot[32];
  LtState* st = state + blockIdx.x;
  unsigned* h = hist + (size_t)blockIdx.x * 2048;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned long long rem = FIRST ? (unsigned long long)ranks[blockIdx.x] : st->rem;
  const uint32_t old_prefix = FIRST ? 0u : st->prefix;
  constexpr int PER = NB / 1024 > 0 ? NB / 1024 : 1;
  unsigned long long mine[PER], sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int b = tid * PER + j;
    mine[j] = b < NB ? h[b] : 0ull;
    sum += mine[j];
  }
  unsigned long long inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const unsigned long long w = warp_tot[lane];
    unsigned long long winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;
  }
  __syncthreads();
  unsigned long long run = warp_tot[wid] + inc - sum;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int b = tid * PER + j;
    if (b < NB && rem >= run && rem < run + mine[j]) {
      st->prefix = (old_prefix << BITS) | (uint32_t)b;
      st->rem = rem - run;
    }
    run += mine[j];
  }
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int b = tid * PER + j;
    if (b < NB) h[b] = 0u;
  }
}

__global__ void gs_export_kernel(const LtState* __restrict__ state, int nseg, uint32_t* __restrict__ tkeys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nseg) tkeys[i] = state[i].prefix;
}

// w *= (score > thr): pruned weights keep their sign bit (w * 0.0 is -0.0 for negative w, exactly like `v.data *= mask`)
template <int MODE>
__global__ void __launch_bounds__(kGsThreads)
    gs_apply_kernel(const ecf_global_desc* __restrict__ table, int n, float nb, int segmented, const uint32_t* __restrict__ prot,
                    const uint32_t* __restrict__ tkeys, unsigned long long* __restrict__ n_pruned) {
  __shared__ int s_tensor;
  __shared__ unsigned s_cnt;
  if (threadIdx.x == 0) {
    s_tensor = gs_find_tensor(table, n, blockIdx.x);
    s_cnt = 0u;
  }
  __syncthreads();
  const int t = s_tensor;
  const ecf_global_desc d = table[t];
  const uint32_t tkey = tkeys[segmented ? t : 0];
  const uint32_t pk = prot != nullptr ? prot[t] : 0xffffffffu;
  const int64_t begin = ((int64_t)blockIdx.x - d.chunk_begin) * kGsChunk;
  const int64_t end = min(d.numel, begin + kGsChunk);
  unsigned cnt = 0;
  for (int64_t i = begin + threadIdx.x; i < end; i += kGsThreads) {
    const uint32_t key = gs_elem_key<MODE>(d, i, nb, pk);
    // mask = score > thr; a NaN score compares false and is "pruned" (w * 0 = NaN stays NaN in torch; finite weights never get here)
    if (key <= tkey || key == kGsNaN) {
      ++cnt;
      if (d.dtype == ECF_F32) reinterpret_cast<uint32_t*>(d.W)[i] &= 0x80000000u;
      else reinterpret_cast<uint16_t*>(d.W)[i] &= 0x8000u;
    }
  }
  cnt = (unsigned)__reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0 && n_pruned != nullptr && s_cnt) atomicAdd(n_pruned + t, (unsigned long long)s_cnt);
}

// A13 / A14 -- per-tensor SUM of the per-element score (what LayerSparsity.return_sparsity reads of a first-order score
// tensor, layer_single_base_pruner.py:361-370): fp32 partials per thread, fp64 per-CTA result, one fp64 atomic per chunk.
template <int MODE>
__global__ void __launch_bounds__(kGsThreads)
    gs_sum_kernel(const ecf_global_desc* __restrict__ table, int n, float nb, double* __restrict__ sums) {
  __shared__ int s_tensor;
  __shared__ double s_part[kGsThreads / 32];
  if (threadIdx.x == 0) s_tensor = gs_find_tensor(table, n, blockIdx.x);
  __syncthreads();
  const int t = s_tensor;
  const ecf_global_desc d = table[t];
  const int64_t begin = ((int64_t)blockIdx.x - d.chunk_begin) * kGsChunk;
  const int64_t end = min(d.numel, begin + kGsChunk);
  float acc = 0.f;
  for (int64_t i = begin + threadIdx.x; i < end; i += kGsThreads) {
    const float w = gs_load(d.W, d.dtype, i);
    const float g = MODE == ECF_GLOBAL_MAG ? 0.f : d.G[i];
    acc += gs_score<MODE>(w, g, nb);
  }
  double v = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < kGsThreads / 32; ++i) tot += s_part[i];
    atomicAdd(sums + t, tot);
  }
}

// G += |g| (or g^2): the device-side accumulation of the first-order loops (the reference adds the gradients up on the CPU
// in fp32, layer_single_base_pruner.py:447-450, global_pruner.py:288)
template <int DT>
__global__ void __launch_bounds__(256) gs_grad_accum_kernel(float* __restrict__ G, const void* __restrict__ g, int64_t n, int square) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float v = load_elem<DT>(g, i);
    G[i] = __fadd_rn(G[i], square ? __fmul_rn(v, v) : fabsf(v));
  }
}

size_t global_select_workspace_bytes(int64_t nseg) {
  return 256 + (size_t)nseg * (sizeof(LtState) + 2048 * sizeof(unsigned));
}

template <int MODE>
static int gs_select(const ecf_global_desc* table, int n, unsigned chunks, float nb, int segmented, const uint32_t* prot,
                     const long long* ranks, uint32_t* tkeys, void* ws, cudaStream_t s) {
  const int nseg = segmented ? n : 1;
  LtState* state = reinterpret_cast<LtState*>(ws);
  unsigned* hist = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + align_up((size_t)nseg * sizeof(LtState), 256));
  ECF_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)nseg * 2048 * sizeof(unsigned), s));
  gs_hist_kernel<MODE, 0><<<chunks, kGsThreads, 0, s>>>(table, n, nb, segmented, prot, state, hist);
  gs_scan_kernel<11, true><<<nseg, 1024, 0, s>>>(state, hist, ranks);
  gs_hist_kernel<MODE, 1><<<chunks, kGsThreads, 0, s>>>(table, n, nb, segmented, prot, state, hist);
  gs_scan_kernel<11, false><<<nseg, 1024, 0, s>>>(state, hist, ranks);
  gs_hist_kernel<MODE, 2><<<chunks, kGsThreads, 0, s>>>(table, n, nb, segmented, prot, state, hist);
  gs_scan_kernel<10, false><<<nseg, 1024, 0, s>>>(state, hist, ranks);
  gs_export_kernel<<<(nseg + 255) / 256, 256, 0, s>>>(state, nseg, tkeys);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

}  // namespace ecf

extern "C" int64_t ecf_global_chunk_elems(void) { return ecf::kGsChunk; }

extern "C" int ecf_global_select(const ecf_global_desc* d_table, int n_tensors, int64_t total_chunks, int mode, double n_batches,
                                 int segmented, const uint32_t* d_protect, const long long* d_ranks, uint32_t* d_tkeys, void* ws,
                                 size_t ws_bytes, ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(d_table != nullptr && d_ranks != nullptr && d_tkeys != nullptr, ECF_ERR_INVALID, "global_select: null pointer");
  ECF_REQUIRE(n_tensors >= 1 && total_chunks >= n_tensors && total_chunks < (1ll << 31), ECF_ERR_INVALID,
              "global_select: bad table size n=%d chunks=%lld", n_tensors, (long long)total_chunks);
  ECF_REQUIRE(mode >= ECF_GLOBAL_MAG && mode <= ECF_GLOBAL_GRAD_ONLY, ECF_ERR_INVALID, "global_select: unknown score mode %d", mode);
  const size_t need = global_select_workspace_bytes(segmented ? n_tensors : 1);
  ECF_REQUIRE(ws != nullptr && ws_bytes >= need, ECF_ERR_WORKSPACE, "global_select: workspace %zu < %zu bytes", ws_bytes, need);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float nb = (float)n_batches;
  const unsigned chunks = (unsigned)total_chunks;
  switch (mode) {
    case ECF_GLOBAL_MAG: return gs_select<ECF_GLOBAL_MAG>(d_table, n_tensors, chunks, nb, segmented, d_protect, d_ranks, d_tkeys, ws, s);
    case ECF_GLOBAL_GRAD_MAG_ABS: return gs_select<ECF_GLOBAL_GRAD_MAG_ABS>(d_table, n_tensors, chunks, nb, segmented, d_protect, d_ranks, d_tkeys, ws, s);
    case ECF_GLOBAL_GRAD_MAG_SQ: return gs_select<ECF_GLOBAL_GRAD_MAG_SQ>(d_table, n_tensors, chunks, nb, segmented, d_protect, d_ranks, d_tkeys, ws, s);
    default: return gs_select<ECF_GLOBAL_GRAD_ONLY>(d_table, n_tensors, chunks, nb, segmented, d_protect, d_ranks, d_tkeys, ws, s);
  }
}

extern "C" int ecf_global_apply(const ecf_global_desc* d_table, int n_tensors, int64_t total_chunks, int mode, double n_batches,
                                int segmented, const uint32_t* d_protect, const uint32_t* d_tkeys, unsigned long long* d_n_pruned,
                                ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(d_table != nullptr && d_tkeys != nullptr, ECF_ERR_INVALID, "global_apply: null pointer");
  ECF_REQUIRE(n_tensors >= 1 && total_chunks >= n_tensors && total_chunks < (1ll << 31), ECF_ERR_INVALID,
              "global_apply: bad table size n=%d chunks=%lld", n_tensors, (long long)total_chunks);
  ECF_REQUIRE(mode >= ECF_GLOBAL_MAG && mode <= ECF_GLOBAL_GRAD_ONLY, ECF_ERR_INVALID, "global_apply: unknown score mode %d", mode);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float nb = (float)n_batches;
  const unsigned chunks = (unsigned)total_chunks;
  switch (mode) {
    case ECF_GLOBAL_MAG: gs_apply_kernel<ECF_GLOBAL_MAG><<<chunks, kGsThreads, 0, s>>>(d_table, n_tensors, nb, segmented, d_protect, d_tkeys, d_n_pruned); break;
    case ECF_GLOBAL_GRAD_MAG_ABS: gs_apply_kernel<ECF_GLOBAL_GRAD_MAG_ABS><<<chunks, kGsThreads, 0, s>>>(d_table, n_tensors, nb, segmented, d_protect, d_tkeys, d_n_pruned); break;
    case ECF_GLOBAL_GRAD_MAG_SQ: gs_apply_kernel<ECF_GLOBAL_GRAD_MAG_SQ><<<chunks, kGsThreads, 0, s>>>(d_table, n_tensors, nb, segmented, d_protect, d_tkeys, d_n_pruned); break;
    default: gs_apply_kernel<ECF_GLOBAL_GRAD_ONLY><<<chunks, kGsThreads, 0, s>>>(d_table, n_tensors, nb, segmented, d_protect, d_tkeys, d_n_pruned); break;
  }
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

extern "C" int ecf_global_score_sum(const ecf_global_desc* d_table, int n_tensors, int64_t total_chunks, int mode, double n_batches,
                                    double* d_sums, ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(d_table != nullptr && d_sums != nullptr, ECF_ERR_INVALID, "global_score_sum: null pointer");
  ECF_REQUIRE(n_tensors >= 1 && total_chunks >= n_tensors && total_chunks < (1ll << 31), ECF_ERR_INVALID,
              "global_score_sum: bad table size n=%d chunks=%lld", n_tensors, (long long)total_chunks);
  ECF_REQUIRE(mode >= ECF_GLOBAL_MAG && mode <= ECF_GLOBAL_GRAD_ONLY, ECF_ERR_INVALID, "global_score_sum: unknown score mode %d", mode);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float nb = (float)n_batches;
  const unsigned chunks = (unsigned)total_chunks;
  switch (mode) {
    case ECF_GLOBAL_MAG: gs_sum_kernel<ECF_GLOBAL_MAG><<<chunks, kGsThreads, 0, s>>>(d_table, n_tensors, nb, d_sums); break;
    case ECF_GLOBAL_GRAD_MAG_ABS: gs_sum_kernel<ECF_GLOBAL_GRAD_MAG_ABS><<<chunks, kGsThreads, 0, s>>>(d_table, n_tensors, nb, d_sums); break;
    case ECF_GLOBAL_GRAD_MAG_SQ: gs_sum_kernel<ECF_GLOBAL_GRAD_MAG_SQ><<<chunks, kGsThreads, 0, s>>>(d_table, n_tensors, nb, d_sums); break;
    default: gs_sum_kernel<ECF_GLOBAL_GRAD_ONLY><<<chunks, kGsThreads, 0, s>>>(d_table, n_tensors, nb, d_sums); break;
  }
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

extern "C" int ecf_grad_accum(float* G, const void* g, int g_dtype, int64_t numel, int square, ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(G != nullptr && g != nullptr && numel >= 0, ECF_ERR_INVALID, "grad_accum: null pointer");
  ECF_REQUIRE(g_dtype >= ECF_F32 && g_dtype <= ECF_BF16, ECF_ERR_INVALID, "grad_accum: unknown dtype %d", g_dtype);
  if (numel == 0) return ECF_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int64_t grid = (numel + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  switch (g_dtype) {
    case ECF_F32: gs_grad_accum_kernel<ECF_F32><<<(unsigned)grid, 256, 0, s>>>(G, g, numel, square); break;
    case ECF_F16: gs_grad_accum_kernel<ECF_F16><<<(unsigned)grid, 256, 0, s>>>(G, g, numel, square); break;
    default: gs_grad_accum_kernel<ECF_BF16><<<(unsigned)grid, 256, 0, s>>>(G, g, numel, square); break;
  }
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}
