// A3+A5+A7 -- fused Wanda score / per-LAYER threshold select / in-place apply, batched over the Linears of a block.
//
// Replaces  thres = torch.sort(W_metric.flatten())[0][int(numel * s)];  W[W_metric <= thres] = 0
// (LAVIS/lavis/compression/pruners/wanda_pruner.py:541,553-558; UPop wanda_pruner.py:502,512-517;
//  LLaMA/image_classifiers/prune_utils.py:28-31).
//
// The exact kth_index-th smallest fp32 score of a whole matrix needs grid-wide agreement several times, and at
// these sizes (2-17 MB per matrix, a few microseconds of HBM time) every kernel boundary costs as much as the data.
// So ONE persistent cooperative kernel serves all the matrices of a block (ViT-g: qkv, proj, fc1, fc2 = 50 MB) and
// moves through its phases with grid barriers:
//   P1  sample   1 vector in S of every matrix -> 32 768-bin histogram of the upper 16 key bits (global REDs on
//                <= 131 072 samples per matrix); the same phase writes q = sqrt(scaler_row) once per column
//   P2  bracket  one CTA per matrix scans its histogram: coarse bracket [lo, hi) = sample ranks k_s -+ 2.5 sqrt(n_s)
//   P3  count    the only HBM read of W: #(key < lo) in registers + 2 048-bin histogram of the top digit of
//                (key - lo) for the ~1-2 % of elements inside the bracket (shared-memory atomics are affordable
//                there: they retire ~0.5 elements/clk/SM, a full-matrix histogram would need ~12)
//   P4  refine   (1-2 times, W now L2 resident) next 11-bit digit of the elements matching the prefix
//   P5  apply    score <= thres -> zero, in place (L2 read, HBM write); optional packed mask / zero count
// (grid barriers: cooperative_groups grid.sync(), the launch is cooperative so all CTAs are co-resident)
// If the k-th score falls outside the sampled bracket (probability ~1e-6, or adversarial ties) the bracket is
// replaced by the side that holds it and P3 is repeated -- the result is always exact.
// Bound: HBM.  Algorithmic bytes per matrix: 2*R*C*sizeof(w) + 4*C (re-reads are L2 hits).
#include <cooperative_groups.h>

#include "common.cuh"

namespace ecf {

constexpr int kLtThreads = 512;
constexpr int kLtMaxMat = ECF_LAYER_MAX_BATCH;
constexpr int kLtCoarseBins = 32768;
constexpr int kLtBins = 2048;
constexpr int kLtSampleVecs = 16384;  // sampled 8-element vectors per matrix (131 072 scores)

struct LtMat {
  void* W;
  const float* s;
  float* q;                 // workspace: sqrt(scaler_row) + 0
  unsigned* coarse;         // workspace: [kLtCoarseBins] sample histogram (self-cleaning)
  unsigned* hist;           // workspace: [3][kLtBins] digit histograms (self-cleaning)
  unsigned long long* cnt;  // workspace: [0] #(key < lo), [1] #(lo <= key < hi)
  uint32_t* bracket;        // workspace: [0] lo, [1] hi (coarse, hi exclusive, <= 0x8000), [2] retry flag
  float* thres_out;
  uint8_t* mask;
  unsigned long long* n_zero;
  int64_t R, C, ld, mask_ld;
  int64_t kth;
  int64_t nvpr;             // vectors per row = ceil(C / 8)
  int64_t nvec;             // R * nvpr
  int64_t vec_begin;        // prefix over the matrices of the launch
  int64_t sample_stride;    // S
  int64_t col_begin;        // prefix of C over the matrices (q-table work split)
  int dtype, aligned;
};

struct LtBatch {
  LtMat m[kLtMaxMat];
  int n;
  int64_t total_vec, total_cols;
  unsigned long long* stamps;  // workspace: %globaltimer of CTA 0 at the phase boundaries (profiling aid)
};

__device__ __forceinline__ void lt_stamp(const LtBatch& b, int i) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    b.stamps[i] = t;
  }
}

namespace cg = cooperative_groups;

template <int DT, bool ALIGNED>
__device__ __forceinline__ void lt_load_chunk(const char* wrow, int64_t c0, int64_t C, uint32_t (&raw)[8]) {
  // raw[j] = bit pattern of element j widened to 32 bits (fp32 bits, or the 16-bit pattern in the low half)
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

// keys of the 8 elements of vector `lv` of matrix M (out-of-range columns get key 0xffffffff: never counted)
template <int DT, bool ALIGNED, bool SQRT_INLINE>
__device__ __forceinline__ void lt_keys(const LtMat& M, int64_t lv, uint32_t (&key)[8], uint32_t (&raw)[8], int64_t& row, int64_t& c0) {
  if (M.nvec < (1ll << 31)) {  // 32-bit division: every BASELINE.json matrix
    const uint32_t r32 = (uint32_t)lv / (uint32_t)M.nvpr;
    row = r32;
    c0 = (int64_t)((uint32_t)lv - r32 * (uint32_t)M.nvpr) * 8;
  } else {
    row = lv / M.nvpr;
    c0 = (lv - row * M.nvpr) * 8;
  }
  const char* wrow = reinterpret_cast<const char*>(M.W) + row * M.ld * DType<DT>::kBytes;
  lt_load_chunk<DT, ALIGNED>(wrow, c0, M.C, raw);
  float q[8];
  if (!SQRT_INLINE && ALIGNED) {
    const float4 qa = *reinterpret_cast<const float4*>(M.q + c0), qb = *reinterpret_cast<const float4*>(M.q + c0 + 4);
    q[0] = qa.x; q[1] = qa.y; q[2] = qa.z; q[3] = qa.w; q[4] = qb.x; q[5] = qb.y; q[6] = qb.z; q[7] = qb.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c0 + j < M.C) q[j] = SQRT_INLINE ? __fadd_rn(sqrtf(M.s[c0 + j]), 0.f) : M.q[c0 + j];
      else q[j] = 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    key[j] = (c0 + j < M.C) ? score_key(wanda_score(lt_to_float<DT>(raw[j]), q[j])) : 0xffffffffu;
}

// dispatch a generic lambda on (dtype, aligned) of a matrix
#define LT_DISPATCH(M, CALL)                                                  \
  do {                                                                        \
    switch ((M).dtype * 2 + (M).aligned) {                                    \
      case ECF_F32 * 2 + 1: { constexpr int DT = ECF_F32; constexpr bool AL = true; CALL; } break;   \
      case ECF_F32 * 2 + 0: { constexpr int DT = ECF_F32; constexpr bool AL = false; CALL; } break;  \
      case ECF_F16 * 2 + 1: { constexpr int DT = ECF_F16; constexpr bool AL = true; CALL; } break;   \
      case ECF_F16 * 2 + 0: { constexpr int DT = ECF_F16; constexpr bool AL = false; CALL; } break;  \
      case ECF_BF16 * 2 + 1: { constexpr int DT = ECF_BF16; constexpr bool AL = true; CALL; } break; \
      default: { constexpr int DT = ECF_BF16; constexpr bool AL = false; CALL; } break;              \
    }                                                                         \
  } while (0)

// per-matrix select state every CTA keeps (identical in all CTAs: derived from global memory after a barrier)
struct LtSel {
  uint32_t lo32, range_hi;  // bracket as fp32 keys: [lo32, lo32 + range)  (range_hi: range - 1, fits 32 bits)
  int nd;                   // 11-bit digits needed for (key - lo32)
  uint32_t prefix;          // digits fixed so far (right aligned)
  unsigned long long rem;   // rank still to resolve inside the prefix bucket
  int done;                 // digits resolved
  int active;               // participates in the current P3 round
};

__device__ __forceinline__ int lt_shift(int nd, int level) { return (nd - 1 - level) * 11; }

// block-wide: find the bin of a 2 048-bin global histogram holding rank `rem`; returns bin, updates rem
__device__ __forceinline__ uint32_t lt_find_bin(const unsigned* hist, unsigned long long& rem, unsigned long long* warp_tot /*[32] smem*/,
                                                uint32_t* out /*[2] smem*/) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int PER = kLtBins / kLtThreads;  // 4
  unsigned long long mine[PER], sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    mine[j] = __ldcg(hist + tid * PER + j);
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
    unsigned long long w = lane < kLtThreads / 32 ? warp_tot[lane] : 0ull;
    unsigned long long winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kLtThreads / 32) warp_tot[lane] = winc - w;  // exclusive
  }
  __syncthreads();
  unsigned long long run = warp_tot[wid] + inc - sum;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    if (rem >= run && rem < run + mine[j]) {
      out[0] = (uint32_t)(tid * PER + j);
      reinterpret_cast<unsigned long long*>(warp_tot)[33] = rem - run;
    }
    run += mine[j];
  }
  __syncthreads();
  rem = reinterpret_cast<unsigned long long*>(warp_tot)[33];
  const uint32_t bin = out[0];
  __syncthreads();
  return bin;
}

__global__ void __launch_bounds__(kLtThreads, 2) layer_thresh_batched_kernel(const __grid_constant__ LtBatch b) {
  extern __shared__ unsigned sh_hist[];  // [n][kLtBins]
  __shared__ unsigned sh_cnt[kLtMaxMat][2];  // per-CTA counts fit 32 bits (a CTA sees < 2^32 elements)
  __shared__ unsigned long long sh_scan[36];
  __shared__ uint32_t sh_out[2];
  __shared__ LtSel sel[kLtMaxMat];
  const int tid = threadIdx.x;
  const int64_t gthreads = (int64_t)gridDim.x * kLtThreads;
  const int64_t gtid = (int64_t)blockIdx.x * kLtThreads + tid;
  cg::grid_group grid = cg::this_grid();
  lt_stamp(b, 0);

  // ================= P1: q tables + sample histogram of the upper 16 key bits ======================================
  for (int64_t i = gtid; i < (int64_t)b.n * 3 * kLtBins; i += gthreads) {  // digit histograms start from zero
    const int mi = (int)(i / (3 * kLtBins));
    b.m[mi].hist[i - (int64_t)mi * 3 * kLtBins] = 0u;
  }
  for (int64_t c = gtid; c < b.total_cols; c += gthreads) {
    int mi = 0;
    while (mi + 1 < b.n && c >= b.m[mi + 1].col_begin) ++mi;
    const LtMat& M = b.m[mi];
    const int64_t cc = c - M.col_begin;
    M.q[cc] = __fadd_rn(sqrtf(M.s[cc]), 0.f);
  }
  for (int mi = 0; mi < b.n; ++mi) {
    const LtMat& M = b.m[mi];
    const int64_t nsv = (M.nvec + M.sample_stride - 1) / M.sample_stride;
    for (int64_t j = gtid; j < nsv; j += gthreads) {
      int64_t lv = j * M.sample_stride;
      if (M.sample_stride > 1) lv += (int64_t)(((uint32_t)j * 2654435761u) >> 8) % M.sample_stride;
      if (lv >= M.nvec) lv = M.nvec - 1;
      uint32_t key[8], raw[8];
      int64_t row, c0;
      LT_DISPATCH(M, (lt_keys<DT, AL, true>(M, lv, key, raw, row, c0)));
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (key[e] != 0xffffffffu) atomicAdd(M.coarse + (key[e] >> 16), 1u);
    }
  }
  grid.sync();
  lt_stamp(b, 1);

  // ================= P2: one CTA per matrix turns the sample histogram into a coarse bracket ======================
  if ((int)blockIdx.x < b.n) {
    const LtMat& M = b.m[blockIdx.x];
    constexpr int PER = kLtCoarseBins / kLtThreads;  // 64 consecutive bins per thread, read as 16 independent 128-bit loads
    const uint4* my_bins = reinterpret_cast<const uint4*>(M.coarse + tid * PER);
    unsigned long long sum = 0;
    {
      uint4 v[PER / 4];
#pragma unroll
      for (int j = 0; j < PER / 4; ++j) v[j] = __ldcg(my_bins + j);
#pragma unroll
      for (int j = 0; j < PER / 4; ++j) sum += (unsigned long long)v[j].x + v[j].y + v[j].z + v[j].w;
    }
    const int lane = tid & 31, wid = tid >> 5;
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sh_scan[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      unsigned long long w = lane < kLtThreads / 32 ? sh_scan[lane] : 0ull;
      unsigned long long winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      if (lane < kLtThreads / 32) sh_scan[lane] = winc - w;
      if (lane == kLtThreads / 32 - 1) sh_scan[34] = winc;  // total number of samples
    }
    __syncthreads();
    const unsigned long long ns = sh_scan[34];
    const double numel = (double)M.R * (double)M.C;
    const long long rs = (long long)((double)M.kth * (double)ns / numel);
    const long long delta = M.sample_stride > 1 ? (long long)(2.5 * sqrt((double)ns)) + 4 : 0;
    const long long r_lo = rs - delta, r_hi = rs + delta;
    if (tid == 0) {
      M.bracket[0] = 0u;
      M.bracket[1] = 0x8000u;
      M.bracket[2] = 0u;
      M.cnt[0] = 0ull;
      M.cnt[1] = 0ull;
    }
    __syncthreads();
    unsigned long long run = sh_scan[wid] + inc - sum;  // samples in the bins before this thread's first bin
    if (sum) {  // only threads whose bins hold samples can contain the two ranks
#pragma unroll 4
      for (int j = 0; j < PER / 4; ++j) {
        const uint4 v = __ldcg(my_bins + j);
        const unsigned cs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const unsigned long long c = cs[e];
          const int bin = tid * PER + 4 * j + e;
          if (c) {
            if (r_lo >= 0 && (unsigned long long)r_lo >= run && (unsigned long long)r_lo < run + c) M.bracket[0] = (uint32_t)bin;
            if (r_hi >= 0 && (unsigned long long)r_hi >= run && (unsigned long long)r_hi < run + c) M.bracket[1] = (uint32_t)bin + 1u;
          }
          run += c;
        }
      }
#pragma unroll
      for (int j = 0; j < PER / 4; ++j)  // self-cleaning for the next launch
        reinterpret_cast<uint4*>(M.coarse + tid * PER)[j] = make_uint4(0, 0, 0, 0);
    }
  }
  grid.sync();
  lt_stamp(b, 2);

  // ================= P3 (+ retry): #(key < lo) and the top-digit histogram of the bracket =========================
  for (int mi = tid; mi < b.n; mi += kLtThreads) sel[mi].active = 1;
  __syncthreads();
  for (int round = 0;; ++round) {
    for (int mi = tid; mi < b.n; mi += kLtThreads) {
      const LtMat& M = b.m[mi];
      if (sel[mi].active) {
        const uint32_t lo = __ldcg(M.bracket), hi = __ldcg(M.bracket + 1);
        const uint32_t lo32 = lo << 16;
        const uint32_t range_hi = (hi >= 0x8000u ? 0x80000000u : (hi << 16)) - lo32 - 1u;
        const int bits = 32 - __clz(range_hi | 1u);
        sel[mi].lo32 = lo32;
        sel[mi].range_hi = range_hi;
        sel[mi].nd = (bits + 10) / 11;
        sel[mi].prefix = 0;
        sel[mi].done = 0;
      }
      sh_cnt[mi][0] = 0u;
      sh_cnt[mi][1] = 0u;
    }
    for (int i = tid; i < b.n * kLtBins; i += kLtThreads) sh_hist[i] = 0;
    __syncthreads();
    {
      int mi = 0;
      unsigned c_lt = 0, c_band = 0;
      for (int64_t v = gtid; v < b.total_vec; v += gthreads) {
        if (v >= b.m[mi].vec_begin + b.m[mi].nvec) {
          if (c_lt) atomicAdd(&sh_cnt[mi][0], c_lt);
          if (c_band) atomicAdd(&sh_cnt[mi][1], c_band);
          c_lt = c_band = 0;
          while (v >= b.m[mi].vec_begin + b.m[mi].nvec) ++mi;
        }
        if (!sel[mi].active) continue;
        const LtMat& M = b.m[mi];
        uint32_t key[8], raw[8];
        int64_t row, c0;
        LT_DISPATCH(M, (lt_keys<DT, AL, false>(M, v - M.vec_begin, key, raw, row, c0)));
        const uint32_t lo32 = sel[mi].lo32, rh = sel[mi].range_hi;
        const int sh = lt_shift(sel[mi].nd, 0);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t d = key[e] - lo32;
          c_lt += key[e] < lo32 ? 1u : 0u;  // padding keys (0xffffffff) are never below lo32
          if (key[e] >= lo32 && d <= rh) {
            ++c_band;
            atomicAdd(&sh_hist[mi * kLtBins + (d >> sh)], 1u);
          }
        }
      }
      // one shared atomic per warp when all its lanes ended in the same matrix (the common case)
      if (__all_sync(0xffffffffu, mi == __shfl_sync(0xffffffffu, mi, 0))) {
        c_lt = (unsigned)warp_sum((int)c_lt);
        c_band = (unsigned)warp_sum((int)c_band);
        if ((tid & 31) != 0) c_lt = c_band = 0;
      }
      if (c_lt) atomicAdd(&sh_cnt[mi][0], c_lt);
      if (c_band) atomicAdd(&sh_cnt[mi][1], c_band);
    }
    __syncthreads();
    for (int i = tid; i < b.n * kLtBins; i += kLtThreads) {
      const unsigned c = sh_hist[i];
      const int mi = i / kLtBins;
      if (c) atomicAdd(b.m[mi].hist + (i - mi * kLtBins), c);
    }
    if (tid < b.n * 2) {
      const unsigned long long c = sh_cnt[tid >> 1][tid & 1];
      if (c) atomicAdd(b.m[tid >> 1].cnt + (tid & 1), c);  // 64-bit global atomic: native
    }
    grid.sync();
    // every CTA checks the brackets (same global values everywhere)
    int any_retry = 0;
    for (int mi = 0; mi < b.n; ++mi) {
      const LtMat& M = b.m[mi];
      const unsigned long long c_lo = __ldcg(M.cnt), c_band = __ldcg(M.cnt + 1);
      const unsigned long long kth = (unsigned long long)M.kth;
      const bool ok = kth >= c_lo && kth < c_lo + c_band;
      if (!ok) any_retry = 1;
      __syncthreads();
      if (tid == 0) {
        sel[mi].active = ok ? 0 : 1;
        if (ok) sel[mi].rem = kth - c_lo;
      }
    }
    __syncthreads();
    if (!any_retry) break;
    // rare: the k-th score is outside the sampled bracket.  Move to the side that holds it and count again.
    grid.sync();  // everybody has read cnt / bracket
    if (blockIdx.x == 0) {
      for (int mi = 0; mi < b.n; ++mi) {
        if (!sel[mi].active) continue;
        const LtMat& M = b.m[mi];
        if (tid == 0) {
          const uint32_t lo = M.bracket[0], hi = M.bracket[1];
          if ((unsigned long long)M.kth < M.cnt[0]) { M.bracket[0] = 0u; M.bracket[1] = lo; }
          else { M.bracket[0] = hi; M.bracket[1] = 0x8000u; }
          M.cnt[0] = 0ull;
          M.cnt[1] = 0ull;
        }
        for (int i = tid; i < kLtBins; i += kLtThreads) M.hist[i] = 0u;
      }
    }
    grid.sync();
  }

  lt_stamp(b, 3);
  // ================= P4: resolve the digits (level 0 histogram is already in global memory) =======================
  int max_nd = 1;
  for (int mi = 0; mi < b.n; ++mi) max_nd = max(max_nd, sel[mi].nd);
  for (int level = 0; level < max_nd; ++level) {
    // scan level `level` (every CTA computes the same bins)
    for (int mi = 0; mi < b.n; ++mi) {
      if (sel[mi].nd <= level) continue;
      unsigned long long rem = sel[mi].rem;
      const uint32_t bin = lt_find_bin(b.m[mi].hist + level * kLtBins, rem, sh_scan, sh_out);
      if (tid == 0) {
        sel[mi].prefix = (sel[mi].prefix << 11) | bin;
        sel[mi].rem = rem;
        sel[mi].done = level + 1;
      }
      __syncthreads();
    }
    if (level + 1 >= max_nd) break;
    // histogram of the next digit for the elements matching the prefix (W is L2 resident by now)
    for (int i = tid; i < b.n * kLtBins; i += kLtThreads) sh_hist[i] = 0;
    __syncthreads();
    {
      int mi = 0;
      for (int64_t v = gtid; v < b.total_vec; v += gthreads) {
        while (v >= b.m[mi].vec_begin + b.m[mi].nvec) ++mi;
        if (sel[mi].nd <= level + 1) continue;
        const LtMat& M = b.m[mi];
        uint32_t key[8], raw[8];
        int64_t row, c0;
        LT_DISPATCH(M, (lt_keys<DT, AL, false>(M, v - M.vec_begin, key, raw, row, c0)));
        const uint32_t lo32 = sel[mi].lo32, rh = sel[mi].range_hi, prefix = sel[mi].prefix;
        const int sh_prev = lt_shift(sel[mi].nd, level), sh = lt_shift(sel[mi].nd, level + 1);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t d = key[e] - lo32;
          if (key[e] >= lo32 && d <= rh && (d >> sh_prev) == prefix) atomicAdd(&sh_hist[mi * kLtBins + ((d >> sh) & (kLtBins - 1))], 1u);
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < b.n * kLtBins; i += kLtThreads) {
      const unsigned c = sh_hist[i];
      const int mi = i / kLtBins;
      if (c) atomicAdd(b.m[mi].hist + (level + 1) * kLtBins + (i - mi * kLtBins), c);
    }
    grid.sync();
  }

  lt_stamp(b, 4);
  // ================= P5: apply  score <= thres  in place ===========================================================
  {
    int mi = 0;
    int zeros = 0;
    for (int64_t v = gtid; v < b.total_vec; v += gthreads) {
      if (v >= b.m[mi].vec_begin + b.m[mi].nvec) {
        if (b.m[mi].n_zero != nullptr && zeros) atomicAdd(b.m[mi].n_zero, (unsigned long long)zeros);
        zeros = 0;
        while (v >= b.m[mi].vec_begin + b.m[mi].nvec) ++mi;
      }
      const LtMat& M = b.m[mi];
      const uint32_t tkey = sel[mi].lo32 + sel[mi].prefix;
      // float semantics of `W_metric <= thres`: NaN scores are never pruned, a NaN threshold prunes nothing
      const uint32_t tcmp = tkey > 0x7f800000u ? 0u : tkey + 1u;  // prune iff key < tcmp
      uint32_t key[8], raw[8];
      int64_t row, c0;
      LT_DISPATCH(M, (lt_keys<DT, AL, false>(M, v - M.vec_begin, key, raw, row, c0)));
      uint32_t m = 0;
      const uint32_t absmask = M.dtype == ECF_F32 ? 0x7fffffffu : 0x7fffu;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool p = key[e] < tcmp;
        m |= (p ? 1u : 0u) << e;
        if (p) raw[e] = 0;
        zeros += (c0 + e < M.C && (raw[e] & absmask) == 0) ? 1 : 0;
      }
      char* wrow = reinterpret_cast<char*>(M.W) + row * M.ld * dtype_bytes(M.dtype);
      if (m) {
        if (M.aligned) {
          if (M.dtype == ECF_F32) {
            stg_v4(wrow + c0 * 4, make_uint4(raw[0], raw[1], raw[2], raw[3]));
            stg_v4(wrow + c0 * 4 + 16, make_uint4(raw[4], raw[5], raw[6], raw[7]));
          } else {
            stg_v4(wrow + c0 * 2, make_uint4(raw[0] | raw[1] << 16, raw[2] | raw[3] << 16, raw[4] | raw[5] << 16, raw[6] | raw[7] << 16));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (c0 + e < M.C && (m >> e & 1)) {
              if (M.dtype == ECF_F32) reinterpret_cast<float*>(wrow)[c0 + e] = 0.f;
              else reinterpret_cast<uint16_t*>(wrow)[c0 + e] = 0;
            }
          }
        }
      }
      if (M.mask != nullptr) M.mask[row * M.mask_ld + (c0 >> 3)] = (uint8_t)m;
    }
    // final flush: one atomic per warp when all its lanes ended in the same matrix (the common case)
    const bool uni = __all_sync(0xffffffffu, mi == __shfl_sync(0xffffffffu, mi, 0));
    if (uni) {
      const int z = warp_sum(zeros);
      if ((tid & 31) == 0 && z && b.m[mi].n_zero != nullptr) atomicAdd(b.m[mi].n_zero, (unsigned long long)z);
    } else if (b.m[mi].n_zero != nullptr && zeros) {
      atomicAdd(b.m[mi].n_zero, (unsigned long long)zeros);
    }
  }
  lt_stamp(b, 5);
  if (blockIdx.x == 0) {
    for (int mi = tid; mi < b.n; mi += kLtThreads)
      if (b.m[mi].thres_out != nullptr) *b.m[mi].thres_out = __uint_as_float(sel[mi].lo32 + sel[mi].prefix);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static size_t lt_mat_fixed_bytes() {
  // coarse hist + 3 digit hists + counters + bracket, each 256-byte aligned
  return align_up((size_t)kLtCoarseBins * 4, 256) + align_up((size_t)3 * kLtBins * 4, 256) + 256 + 256;
}

size_t layer_thresh_batched_workspace_bytes(const ecf_layer_desc* descs, int n) {
  if (descs == nullptr || n < 1 || n > kLtMaxMat) return 0;
  size_t total = 256;
  for (int i = 0; i < n; ++i) {
    if (descs[i].C <= 0) return 0;
    total += lt_mat_fixed_bytes() + align_up((size_t)descs[i].C * sizeof(float), 256);
  }
  return total;
}

size_t layer_thresh_workspace_bytes(int64_t R, int64_t C) {
  (void)R;
  return 256 + lt_mat_fixed_bytes() + align_up((size_t)(C > 0 ? C : 1) * sizeof(float), 256);
}

}  // namespace ecf

extern "C" size_t ecf_layer_thresh_batched_workspace_bytes(const ecf_layer_desc* descs, int n) {
  return ecf::layer_thresh_batched_workspace_bytes(descs, n);
}

extern "C" int ecf_wanda_layer_thresh_apply_batched(const ecf_layer_desc* descs, int n, void* ws, size_t ws_bytes,
                                                    ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(descs != nullptr && n >= 1 && n <= kLtMaxMat, ECF_ERR_INVALID, "layer_thresh: batch size %d outside [1, %d]", n, kLtMaxMat);
  for (int i = 0; i < n; ++i) {
    const ecf_layer_desc& d = descs[i];
    ECF_REQUIRE(d.W != nullptr && d.scaler_row != nullptr, ECF_ERR_INVALID, "layer_thresh: null pointer (matrix %d)", i);
    ECF_REQUIRE(d.R > 0 && d.C > 0 && d.ld >= d.C, ECF_ERR_INVALID, "layer_thresh: bad shape R=%lld C=%lld ld=%lld (matrix %d)",
                (long long)d.R, (long long)d.C, (long long)d.ld, i);
    ECF_REQUIRE(d.dtype >= 0 && d.dtype <= 2, ECF_ERR_INVALID, "layer_thresh: unknown dtype %d (matrix %d)", d.dtype, i);
    // python indexing: sort(...)[idx] raises IndexError for idx >= numel; negative idx is never produced
    ECF_REQUIRE(d.kth_index >= 0 && d.kth_index < d.R * d.C, ECF_ERR_RANGE,
                "layer_thresh: kth_index %lld out of range for %lld elements (the reference raises IndexError)",
                (long long)d.kth_index, (long long)(d.R * d.C));
    ECF_REQUIRE(d.mask_bits == nullptr || d.mask_ld >= (d.C + 7) / 8, ECF_ERR_INVALID, "layer_thresh: mask_ld too small (matrix %d)", i);
    for (int j = 0; j < i; ++j)
      ECF_REQUIRE(descs[j].W != d.W, ECF_ERR_INVALID, "layer_thresh: matrices %d and %d are the same tensor", j, i);
  }
  const size_t need = layer_thresh_batched_workspace_bytes(descs, n);
  ECF_REQUIRE(ws != nullptr && ws_bytes >= need, ECF_ERR_WORKSPACE, "layer_thresh: workspace %zu < %zu bytes", ws_bytes, need);

  LtBatch b;
  b.n = n;
  char* p = reinterpret_cast<char*>(ws);
  b.stamps = reinterpret_cast<unsigned long long*>(p);
  p += 256;
  int64_t vec = 0, cols = 0;
  for (int i = 0; i < n; ++i) {
    const ecf_layer_desc& d = descs[i];
    LtMat& M = b.m[i];
    M.W = d.W; M.s = d.scaler_row; M.R = d.R; M.C = d.C; M.ld = d.ld; M.dtype = d.dtype; M.kth = d.kth_index;
    M.thres_out = d.thres_out; M.mask = d.mask_bits; M.mask_ld = d.mask_ld; M.n_zero = d.n_zero;
    const int V = d.dtype == ECF_F32 ? 4 : 8;
    M.aligned = ((d.C % 8 == 0) && (d.ld % V == 0) && ((reinterpret_cast<uintptr_t>(d.W) & 15) == 0)) ? 1 : 0;
    M.nvpr = (d.C + 7) / 8;
    M.nvec = d.R * M.nvpr;
    M.vec_begin = vec;
    vec += M.nvec;
    M.col_begin = cols;
    cols += d.C;
    M.sample_stride = M.nvec / kLtSampleVecs;
    if (M.sample_stride < 1) M.sample_stride = 1;
    M.coarse = reinterpret_cast<unsigned*>(p); p += align_up((size_t)kLtCoarseBins * 4, 256);
    M.hist = reinterpret_cast<unsigned*>(p); p += align_up((size_t)3 * kLtBins * 4, 256);
    M.cnt = reinterpret_cast<unsigned long long*>(p); p += 256;
    M.bracket = reinterpret_cast<uint32_t*>(p); p += 256;
    M.q = reinterpret_cast<float*>(p); p += align_up((size_t)d.C * sizeof(float), 256);
  }
  b.total_vec = vec;
  b.total_cols = cols;

  const size_t smem = (size_t)n * kLtBins * sizeof(unsigned);
  static size_t smem_opted = 0;
  if (smem > smem_opted) {
    ECF_CUDA_OK(cudaFuncSetAttribute(layer_thresh_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_opted = smem;
  }
  int occ = 0;
  ECF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layer_thresh_batched_kernel, kLtThreads, smem));
  ECF_REQUIRE(occ >= 1, ECF_ERR_CUDA, "layer_thresh: kernel does not fit on an SM");
  int64_t want = (vec + kLtThreads - 1) / kLtThreads;
  const int64_t cap = (int64_t)sm_count() * occ;
  if (want > cap) want = cap;
  if (want < n) want = n;  // P2 needs one CTA per matrix (n <= 8 <= SM count)
  void* args[] = {(void*)&b};
  ECF_CUDA_OK(cudaLaunchCooperativeKernel((const void*)layer_thresh_batched_kernel, dim3((unsigned)want), dim3(kLtThreads), args, smem,
                                          reinterpret_cast<cudaStream_t>(stream)));
  return ECF_OK;
}

extern "C" int ecf_wanda_layer_thresh_apply(void* W, int w_dtype, int64_t R, int64_t C, int64_t ld,
                                            const float* scaler_row, int64_t kth_index, float* thres_out,
                                            uint8_t* mask_bits, int64_t mask_ld, unsigned long long* n_zero, void* ws,
                                            size_t ws_bytes, ecf_stream_t stream) {
  ecf_layer_desc d;
  d.W = W; d.scaler_row = scaler_row; d.R = R; d.C = C; d.ld = ld; d.dtype = w_dtype; d.kth_index = kth_index;
  d.thres_out = thres_out; d.mask_bits = mask_bits; d.mask_ld = mask_ld; d.n_zero = n_zero;
  return ecf_wanda_layer_thresh_apply_batched(&d, 1, ws, ws_bytes, stream);
}
