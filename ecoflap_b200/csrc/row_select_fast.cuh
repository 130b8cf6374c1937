// A3+A4+A7 -- fused Wanda score / per-row k-smallest select / in-place apply: the HBM-roofline path.
//
// Replaces  W_metric = |W| * sqrt(scaler_row);  sort(W_metric, dim=-1, stable=True);  indices[:, :k];
//           scatter_;  W[mask] = 0
// (LAVIS/lavis/compression/pruners/wanda_pruner.py:260,272-279; CoOp wanda_pruner.py:357,379-383;
//  UPop wanda_pruner.py:243,253-260; LLaMA/image_classifiers/prune_utils.py:35-38).
//
// Budget.  At 6.5 TB/s a 16-bit row element (read + write) may cost ~22 issue slots per SM-clock
// across both ALU pipes, so a 31-round bisection on 32-bit keys (2 ops per element per round) cannot keep
// up.  The select is therefore coarse-to-fine:
//   * every element's exact fp32 score is computed once; its upper 16 bits (clamped to 0x7bff so that
//     the pattern is a finite fp16 number) are kept PACKED two per register.  Comparing such patterns
//     as fp16 values equals comparing them as integers, so one HSET2 + one HADD2 counts two elements
//     against a pivot: 1 op per element per pass.
//   * a 32-sample in-warp bitonic sort seeds a bracket [lo, hi) around the k-th coarse key; offset
//     interpolation passes (aiming M/3 elements either side of rank k, bisection when a pass fails to
//     halve the bracket) shrink it until it holds <= M elements or a single coarse value (4-7 passes).
//   * only the elements inside the bracket are ranked exactly by (fp32 key, column), which reproduces
//     torch.sort(stable=True)[:, :k] bit for bit.  Heavy ties (thousands of equal coarse keys, e.g.
//     already-pruned weights) take a slower in-register bisection on the low key bits and the column.
// sqrt(scaler_row) is computed once per CTA into shared memory (CTAs are persistent over rows).
// One GROUP of G lanes (32 ... 512) owns a row; a lane holds NV 8-element vectors of it.
// Bound: HBM.  Algorithmic bytes per call: 2*R*C*sizeof(w) + 4*C.
#pragma once

#include "common.cuh"

namespace ecf {

constexpr int kRfMaxWarps = 16;   // warps per row group (G <= 512)
constexpr int kRfMaxGroups = 8;   // row groups per CTA (G >= 32, CTA of 256 threads)
constexpr int kRfCap = 256;       // exact-ranking capacity per row
constexpr int kRfBand = 32;       // stop narrowing once the bracket holds this many elements
constexpr uint32_t kRfInf = 0x7c00u;  // fp16 +inf pattern: sorts after every real coarse key

constexpr int kRfMaxMat = ECF_ROW_MAX_BATCH;  // matrices (same C, dtype) served by one launch

// one matrix of a batched launch
struct RfMat {
  void* W;
  const float* s;
  uint8_t* mask;
  unsigned long long* n_zero;
  int64_t R, ld, mask_ld;
  int k;
  int batch_begin;  // first row batch (BLOCK / G rows) of this matrix in the launch
};
struct RfBatch {
  RfMat m[kRfMaxMat];
  int n, C, G, total_batches;
  int prefetch;  // issue L2 prefetches for a group's next row (A/B switch ECF_RS_PREFETCH)
};

struct RfShared {
  int slots[2][kRfMaxGroups][kRfMaxWarps];
  uint32_t seed[kRfMaxGroups][2];
  unsigned long long thr[kRfMaxGroups];
  int cand_n[kRfMaxGroups];
  unsigned long long cand[kRfMaxGroups][kRfCap];  // key << 32 | column
};

__device__ __forceinline__ __half2 rf_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t rf_dup(uint32_t p) { return p | (p << 16); }

template <bool MULTI>
__device__ __forceinline__ void rf_gsync(int group, int G) {
  if constexpr (MULTI) named_bar_sync(1 + group, G);
  else __syncwarp();
}

// sum of a (packed) int over the G lanes of a row group
template <bool MULTI>
__device__ __forceinline__ int rf_gsum(int v, int G, int group, int wig, int lane, RfShared& sh, int& parity) {
  const int w = warp_sum(v);
  if constexpr (!MULTI) return w;
  int* s = sh.slots[parity][group];
  if (lane == 0) s[wig] = w;
  named_bar_sync(1 + group, G);
  parity ^= 1;
  int tot = 0;
  const int nw = G >> 5;
  for (int j = 0; j < nw; ++j) tot += s[j];
  return tot;
}

// number of packed coarse keys below the pivot (both halves): one HSET2 + one HADD2 per TWO elements
template <int NP>
__device__ __forceinline__ int rf_count1(const uint32_t (&co)[NP], uint32_t pivot) {
  const __half2 pv = rf_h2(rf_dup(pivot));
  __half2 a[4] = {rf_h2(0u), rf_h2(0u), rf_h2(0u), rf_h2(0u)};  // NP is a multiple of 4: four independent chains
#pragma unroll
  for (int p = 0; p < NP; p += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = __hadd2(a[u], __hlt2(rf_h2(co[p + u]), pv));
  }
  const float2 f = __half22float2(__hadd2(__hadd2(a[0], a[1]), __hadd2(a[2], a[3])));
  return (int)(f.x + f.y);
}

template <int NP>
__device__ __forceinline__ int rf_count2(const uint32_t (&co)[NP], uint32_t lo, uint32_t hi) {
  const __half2 pl = rf_h2(rf_dup(lo)), ph = rf_h2(rf_dup(hi));
  __half2 a0 = rf_h2(0u), a1 = rf_h2(0u), b0 = rf_h2(0u), b1 = rf_h2(0u);
#pragma unroll
  for (int p = 0; p < NP; p += 2) {
    a0 = __hadd2(a0, __hlt2(rf_h2(co[p]), pl));
    b0 = __hadd2(b0, __hlt2(rf_h2(co[p]), ph));
    a1 = __hadd2(a1, __hlt2(rf_h2(co[p + 1]), pl));
    b1 = __hadd2(b1, __hlt2(rf_h2(co[p + 1]), ph));
  }
  const float2 fa = __half22float2(__hadd2(a0, a1)), fb = __half22float2(__hadd2(b0, b1));
  return (int)(fa.x + fa.y) | ((int)(fb.x + fb.y) << 16);
}

__device__ __forceinline__ uint32_t rf_coarse_pair(float w0, float w1, float q0, float q1) {
  const uint32_t b0 = __float_as_uint(__fmul_rn(fabsf(w0), q0)), b1 = __float_as_uint(__fmul_rn(fabsf(w1), q1));
  // upper halves of both scores, each clamped to the largest finite fp16 pattern (one VIMNMX.U16x2)
  return __vminu2(__byte_perm(b0, b1, 0x7632), 0x7bff7bffu);
}

// Lane-local walk over the bracket elements recorded in `bm` (bit e = element e of this lane: vector e >> 3,
// slot e & 7; `ebase` = 0 for the low word, 32 for the high word).  The weight is re-read from global memory (an
// L1/L2 hit: this lane has just streamed the row), so no register array is indexed dynamically and the code stays
// small.
//   MODE 0: append (key, column) to the shared candidate list
//   MODE 1: count keys < a                    MODE 2: count keys == a with column < b
//   MODE 3: zero the element when (key, column) <= thr; patch the packed mask; count new zeros
template <int DT, int MODE>
__device__ __forceinline__ int rf_walk(uint32_t bm, int ebase, char* wrow, const float* qtab, int G, int gl, uint32_t a, uint32_t b,
                                       unsigned long long thr, unsigned long long* cand, int* cand_n, uint8_t* mask_row) {
  int c = 0;
  while (bm) {
    const int e = ebase + __ffs((int)bm) - 1;
    bm &= bm - 1;
    const uint32_t col = (uint32_t)(((e >> 3) * G + gl) * 8 + (e & 7));
    const float w = load_elem<DT>(wrow, col);
    const uint32_t key = score_key(wanda_score(w, qtab[col]));
    if constexpr (MODE == 0) {
      cand[atomicAdd(cand_n, 1)] = ((unsigned long long)key << 32) | col;
    } else if constexpr (MODE == 1) {
      c += key < a ? 1 : 0;
    } else if constexpr (MODE == 2) {
      c += (key == a && col < b) ? 1 : 0;
    } else {
      if ((((unsigned long long)key << 32) | col) <= thr) {
        store_zero<DT>(wrow, col);
        if (mask_row != nullptr) mask_row[col >> 3] |= (uint8_t)(1u << (col & 7));
        c += w != 0.f ? 1 : 0;
      }
    }
  }
  return c;
}

template <int DT, int MODE>
__device__ __forceinline__ int rf_walk2(uint32_t bm0, uint32_t bm1, char* wrow, const float* qtab, int G, int gl, uint32_t a,
                                        uint32_t b, unsigned long long thr, unsigned long long* cand, int* cand_n,
                                        uint8_t* mask_row) {
  return rf_walk<DT, MODE>(bm0, 0, wrow, qtab, G, gl, a, b, thr, cand, cand_n, mask_row) +
         rf_walk<DT, MODE>(bm1, 32, wrow, qtab, G, gl, a, b, thr, cand, cand_n, mask_row);
}

// byte-granular OR into the packed mask from a lane that does not own the byte
__device__ __forceinline__ void rf_mask_or(uint8_t* mask_row, uint32_t col) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(mask_row + (col >> 3));
  atomicOr(reinterpret_cast<unsigned*>(addr & ~uintptr_t(3)), 1u << ((addr & 3) * 8 + (col & 7)));
}

template <int DT, int NV, bool MULTI, int BLOCK, bool KEEP>
__global__ void __launch_bounds__(BLOCK, (BLOCK == 256 ? ((NV <= 4 || !KEEP) ? 3 : 2) : 1))
    row_select_fast_kernel(const __grid_constant__ RfBatch tb) {
  constexpr int NP = 4 * NV;  // packed pairs per lane
  constexpr bool F32 = (DT == ECF_F32);
  constexpr bool REREAD = F32 || !KEEP;  // weights are not kept in registers: the apply pass re-reads the row (L2 hit)
  extern __shared__ __align__(16) float qtab[];
  __shared__ RfShared sh;

  const int tid = threadIdx.x;
  const bool PREFETCH = tb.prefetch != 0;
  const int C = tb.C, G = tb.G;
  const int cpad = NV * G * 8;
  const int group = tid / G;
  const int gl = tid - group * G;
  const int lane = tid & 31;
  const int wig = gl >> 5;  // warp inside the group
  const int rows_per_cta = BLOCK / G;
  int parity = 0;
  int zeros = 0;

  // a CTA owns a contiguous range of row batches: it crosses at most a couple of matrix boundaries, and at each
  // one the whole CTA rebuilds sqrt(scaler_row) in shared memory
  const int per = (tb.total_batches + (int)gridDim.x - 1) / (int)gridDim.x;
  const int bt0 = (int)blockIdx.x * per, bt1 = min(tb.total_batches, bt0 + per);
  int mi = -1;
  void* W = nullptr;
  int64_t R = 0, ld = 0, mask_ld = 0;
  uint8_t* mask_bits = nullptr;
  unsigned long long* n_zero = nullptr;
  int k = 0;

  // Outer loop: the matrices this CTA's range of row batches crosses (at most a couple); everything that depends on the
  // matrix -- descriptor fields, the sqrt(scaler_row) table, the end of its batch range -- is set up here, so that the
  // per-row loop carries a row index and a row pointer and nothing else (round 1 re-derived the matrix per row, with a
  // 64-bit division: ~12 % of the kernel's instructions).
  for (int bt = bt0; bt < bt1;) {
    if (mi < 0) mi = 0;
    while (mi + 1 < tb.n && bt >= tb.m[mi + 1].batch_begin) ++mi;
    const RfMat& M = tb.m[mi];
    const int seg_end = min(bt1, mi + 1 < tb.n ? tb.m[mi + 1].batch_begin : tb.total_batches);
    W = M.W; R = M.R; ld = M.ld; mask_bits = M.mask; mask_ld = M.mask_ld; n_zero = M.n_zero; k = M.k;
    __syncthreads();  // every group is done with the previous matrix' table
    for (int c = tid; c < cpad; c += BLOCK) qtab[c] = c < C ? __fadd_rn(sqrtf(M.s[c]), 0.f) : 0.f;
    __syncthreads();
    int64_t row = (int64_t)(bt - M.batch_begin) * rows_per_cta + group;
    const int64_t row_step_bytes = (int64_t)rows_per_cta * ld * DType<DT>::kBytes;
    char* wrow = reinterpret_cast<char*>(W) + row * ld * DType<DT>::kBytes;
  for (; bt < seg_end; ++bt, row += rows_per_cta, wrow += row_step_bytes) {
    if (row >= R) continue;
    if (PREFETCH && bt + 1 < seg_end && row + rows_per_cta < R) {
      // the group's next row starts its trip from HBM to L2 now; its loads (a whole selection later) then miss L1 only
      const char* nrow = wrow + row_step_bytes;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c0 = (i * G + gl) * 8;
        if (c0 < C) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + (int64_t)c0 * DType<DT>::kBytes));
          if constexpr (F32) asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + (int64_t)c0 * 4 + 16));
        }
      }
    }
    uint8_t* mask_row = mask_bits != nullptr ? mask_bits + row * mask_ld : nullptr;
    uint32_t co[NP];
    uint32_t raw[REREAD ? 1 : NP];

    // ---- load the row once, exact scores -> packed coarse keys ----------------------------------
    if constexpr (!F32) {
      uint4 v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c0 = (i * G + gl) * 8;
        v[i] = c0 < C ? ldg_v4(wrow + (int64_t)c0 * 2) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c0 = (i * G + gl) * 8;
        const uint32_t rv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        if constexpr (!REREAD) { raw[4 * i + 0] = rv[0]; raw[4 * i + 1] = rv[1]; raw[4 * i + 2] = rv[2]; raw[4 * i + 3] = rv[3]; }
        const float4 qa = *reinterpret_cast<const float4*>(qtab + c0), qb = *reinterpret_cast<const float4*>(qtab + c0 + 4);
        const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float w0, w1;
          unpack2<DT>(rv[j], w0, w1);
          co[4 * i + j] = c0 < C ? rf_coarse_pair(w0, w1, q[2 * j], q[2 * j + 1]) : rf_dup(kRfInf);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c0 = (i * G + gl) * 8;
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (c0 < C) {
          a = ldg_v4(wrow + (int64_t)c0 * 4);
          b = ldg_v4(wrow + (int64_t)c0 * 4 + 16);
        }
        const float4 qa = *reinterpret_cast<const float4*>(qtab + c0), qb = *reinterpret_cast<const float4*>(qtab + c0 + 4);
        const uint32_t inf2 = rf_dup(kRfInf);
        co[4 * i + 0] = c0 < C ? rf_coarse_pair(__uint_as_float(a.x), __uint_as_float(a.y), qa.x, qa.y) : inf2;
        co[4 * i + 1] = c0 < C ? rf_coarse_pair(__uint_as_float(a.z), __uint_as_float(a.w), qa.z, qa.w) : inf2;
        co[4 * i + 2] = c0 < C ? rf_coarse_pair(__uint_as_float(b.x), __uint_as_float(b.y), qb.x, qb.y) : inf2;
        co[4 * i + 3] = c0 < C ? rf_coarse_pair(__uint_as_float(b.z), __uint_as_float(b.w), qb.z, qb.w) : inf2;
      }
    }
    if constexpr (REREAD) raw[0] = 0;

    // ---- coarse bracket [lo, hi):  #(coarse < lo) < k <= #(coarse < hi) ----------------------------
    uint32_t lo = 0, hi = 0;
    int c_lo = 0, c_hi = 0;
    if (k >= C) {
      lo = hi = kRfInf;  // every real element is below the bracket
    } else if (k > 0) {
      if (gl == 0) sh.cand_n[group] = 0;
      if constexpr (!MULTI) __syncwarp();
      // seed: 32 samples (first element of vector 0 of the group's first warp), bitonic sort
      const int ns = min(32, C >> 3);
      if (wig == 0) {
        uint32_t v = (gl * 8 < C) ? (co[0] & 0xffffu) : kRfInf;
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
          for (int j = kk >> 1; j > 0; j >>= 1) {
            const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = (lane & kk) == 0, lower = (lane & j) == 0;
            v = (up == lower) ? min(v, o) : max(v, o);
          }
        }
        const int r = (int)(((int64_t)k * ns) / C);
        const int il = r - 3, ih = r + 3;
        const uint32_t vlo = __shfl_sync(0xffffffffu, v, il < 0 ? 0 : il);
        const uint32_t vhi = __shfl_sync(0xffffffffu, v, ih > 31 ? 31 : ih);
        lo = il >= 0 ? vlo : 0u;
        hi = ih < ns ? vhi + 1u : kRfInf;
        if constexpr (MULTI) {
          if (lane == 0) { sh.seed[group][0] = lo; sh.seed[group][1] = hi; }
        }
      }
      if constexpr (MULTI) {
        named_bar_sync(1 + group, G);
        lo = sh.seed[group][0];
        hi = sh.seed[group][1];
      }
      {
        const int cc = rf_gsum<MULTI>(rf_count2<NP>(co, lo, hi), G, group, wig, lane, sh, parity);
        c_lo = cc & 0xffff;
        c_hi = cc >> 16;
      }
      if (c_lo >= k) { hi = lo; c_hi = c_lo; lo = 0; c_lo = 0; }
      else if (c_hi < k) { lo = hi; c_lo = c_hi; hi = kRfInf; c_hi = C; }
      int stall = 0;
      while (c_hi - c_lo > kRfBand && hi - lo > 1) {
        const int n = c_hi - c_lo;
        uint32_t p;
        if (stall >= 2) {
          p = (lo + hi) >> 1;
          stall = 0;
        } else {
          const int dl = k - c_lo, dh = c_hi - k;
          const float target = (float)dl + (dl >= dh ? -(float)(kRfBand / 3) : (float)(kRfBand / 3));
          const float f = __saturatef(__fdividef(target, (float)n));
          p = lo + (uint32_t)__float2int_rn((float)(hi - lo) * f);
        }
        p = min(max(p, lo + 1), hi - 1);
        const int c = rf_gsum<MULTI>(rf_count1<NP>(co, p), G, group, wig, lane, sh, parity);
        if (c < k) { lo = p; c_lo = c; } else { hi = p; c_hi = c; }
        stall = (c_hi - c_lo) * 2 > n ? stall + 1 : 0;
      }
    }

    // ---- coarse apply (everything below the bracket goes) + store; record this lane's bracket elements ----
    uint32_t bm0 = 0, bm1 = 0;  // bit e: element e (vectors 0-3) / 32 + e (vectors 4-7) of this lane lies in the bracket
    {
      const __half2 pl = rf_h2(rf_dup(lo)), ph = rf_h2(rf_dup(hi));
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c0 = (i * G + gl) * 8;
        uint32_t ml[4];
        uint32_t bacc = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = 4 * i + j;
          ml[j] = __hlt2_mask(rf_h2(co[p]), pl);
          const uint32_t mh = __hlt2_mask(rf_h2(co[p]), ph);
          bacc |= (mh ^ ml[j]) & ((1u << (2 * j)) | (0x10000u << (2 * j + 1)));
        }
        if (i < 4) bm0 |= ((bacc | (bacc >> 16)) & 0xffu) << (8 * (i & 3));
        else bm1 |= ((bacc | (bacc >> 16)) & 0xffu) << (8 * (i & 3));
        if (c0 >= C) continue;
        if constexpr (!F32) {
          uint32_t v[4];
          if constexpr (REREAD) {
            const uint4 a = ldg_v4(wrow + (int64_t)c0 * 2);
            v[0] = a.x & ~ml[0]; v[1] = a.y & ~ml[1]; v[2] = a.z & ~ml[2]; v[3] = a.w & ~ml[3];
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = raw[4 * i + j] & ~ml[j];
          }
          stg_v4(wrow + (int64_t)c0 * 2, make_uint4(v[0], v[1], v[2], v[3]));
          if (n_zero != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) zeros += ((v[j] & 0x00007fffu) == 0 ? 1 : 0) + ((v[j] & 0x7fff0000u) == 0 ? 1 : 0);
          }
        } else {
          const uint32_t any = ml[0] | ml[1] | ml[2] | ml[3];
          if (any || n_zero != nullptr) {
            uint4 a = ldg_v4(wrow + (int64_t)c0 * 4), b = ldg_v4(wrow + (int64_t)c0 * 4 + 16);
            uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (ml[j] & 0x0000ffffu) v[2 * j] = 0;
              if (ml[j] & 0xffff0000u) v[2 * j + 1] = 0;
            }
            if (any) {
              stg_v4(wrow + (int64_t)c0 * 4, make_uint4(v[0], v[1], v[2], v[3]));
              stg_v4(wrow + (int64_t)c0 * 4 + 16, make_uint4(v[4], v[5], v[6], v[7]));
            }
            if (n_zero != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) zeros += ((v[j] & 0x7fffffffu) == 0) ? 1 : 0;
            }
          }
        }
        if (mask_row != nullptr) {
          uint32_t mb = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) mb |= ((ml[j] & 1u) | ((ml[j] >> 15) & 2u)) << (2 * j);
          mask_row[c0 >> 3] = (uint8_t)mb;
        }
      }
    }

    // ---- exact ranking of the bracket by (fp32 key, column) ----------------------------------------------
    const int m = c_hi - c_lo, need = k - c_lo;  // bracket size, how many of it must go (1 <= need <= m when m > 0)
    if (m > 0) {
      if (m <= kRfCap) {
        // gather the bracket into shared memory (lane-local walk), rank, then one lane per candidate patches it.
        // The barrier after the gather also orders every lane's vector stores before the scalar patches.
        rf_walk2<DT, 0>(bm0, bm1, wrow, qtab, G, gl, 0, 0, 0ull, sh.cand[group], &sh.cand_n[group], nullptr);
        rf_gsync<MULTI>(group, G);
        unsigned long long thr = ~0ull;  // need == m: the whole bracket goes
        if (need < m) {
          for (int t = gl; t < m; t += G) {
            const unsigned long long me = sh.cand[group][t];
            int rank = 0;
#pragma unroll 4
            for (int j = 0; j < m; ++j) rank += sh.cand[group][j] < me ? 1 : 0;
            if (rank == need - 1) sh.thr[group] = me;
          }
          rf_gsync<MULTI>(group, G);
          thr = sh.thr[group];
        }
        for (int t = gl; t < m; t += G) {
          const unsigned long long me = sh.cand[group][t];
          if (me <= thr) {
            const uint32_t col = (uint32_t)me;
            if (n_zero != nullptr) zeros += load_elem<DT>(wrow, col) != 0.f ? 1 : 0;
            store_zero<DT>(wrow, col);
            if (mask_row != nullptr) rf_mask_or(mask_row, col);
          }
        }
      } else {
        // heavy ties: the bracket is a single coarse value.  Bisect the fp32 key, then the column.
        unsigned long long thr = ~0ull;
        if (need < m) {
          uint32_t L = lo << 16, H = (lo >= 0x7bffu) ? 0x80000000u : ((lo + 1u) << 16);
          int cL = 0;
          while (H - L > 1) {
            const uint32_t pv = L + ((H - L) >> 1);
            const int c = rf_gsum<MULTI>(rf_walk2<DT, 1>(bm0, bm1, wrow, qtab, G, gl, pv, 0, 0ull, nullptr, nullptr, nullptr), G,
                                         group, wig, lane, sh, parity);
            if (c < need) { L = pv; cL = c; } else { H = pv; }
          }
          const int need2 = need - cL;        // ties at key L that must go, lowest column first
          uint32_t CL = 0, CH = (uint32_t)C;  // #(ties with col < CL) < need2 <= #(ties with col < CH)
          while (CH - CL > 1) {
            const uint32_t pc = (CL + CH) >> 1;
            const int c = rf_gsum<MULTI>(rf_walk2<DT, 2>(bm0, bm1, wrow, qtab, G, gl, L, pc, 0ull, nullptr, nullptr, nullptr), G,
                                         group, wig, lane, sh, parity);
            if (c < need2) CL = pc; else CH = pc;
          }
          thr = ((unsigned long long)L << 32) | CL;
        }
        // scalar patch of this lane's own bracket elements (after its own vector stores: same-thread order)
        zeros += rf_walk2<DT, 3>(bm0, bm1, wrow, qtab, G, gl, 0, 0, thr, nullptr, nullptr, mask_row);
      }
      // the next row reuses sh.cand / sh.thr / sh.seed: every lane of the group must be done with them
      rf_gsync<MULTI>(group, G);
    }
  }
    if (n_zero != nullptr) {  // this matrix' zero count
      const int z = warp_sum(zeros);
      if (lane == 0 && z) atomicAdd(n_zero, (unsigned long long)z);
    }
    zeros = 0;
  }
}

// ---------------------------------------------------------------------------------------------- host
template <int DT, int NV, bool MULTI, int BLOCK, bool KEEP>
static int rf_launch(RfBatch& tb, cudaStream_t stream) {
  auto kern = row_select_fast_kernel<DT, NV, MULTI, BLOCK, KEEP>;
  const int G = tb.G;
  const size_t smem = (size_t)NV * G * 8 * sizeof(float);
  static size_t smem_opted = 0;  // largest dynamic size this instantiation has been opted in for
  if (smem + sizeof(RfShared) > 48 * 1024 && smem > smem_opted) {
    ECF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_opted = smem;
  }
  int occ = 0;
  ECF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem));
  if (occ < 1) occ = 1;
  const int rows_per_cta = BLOCK / G;
  int64_t batches = 0;
  for (int i = 0; i < tb.n; ++i) {
    tb.m[i].batch_begin = (int)batches;
    batches += (tb.m[i].R + rows_per_cta - 1) / rows_per_cta;
  }
  ECF_REQUIRE(batches < (1ll << 31), ECF_ERR_INVALID, "row_select: too many rows in one launch");
  tb.total_batches = (int)batches;
  const int64_t cap = (int64_t)sm_count() * occ;
  const unsigned grid = (unsigned)(batches < cap ? batches : cap);
  kern<<<grid, BLOCK, smem, stream>>>(tb);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

template <int DT, int NV>
static int rf_dispatch_g(RfBatch& tb, bool keep, cudaStream_t stream) {
  const int G = tb.G;
  // KEEP only changes code for 16-bit weights with more than 4 vectors per lane (fp32 always re-reads)
  if constexpr (DT != ECF_F32 && NV > 4) {
    if (keep) {
      if (G == 32) return rf_launch<DT, NV, false, 256, true>(tb, stream);
      if (G <= 256) return rf_launch<DT, NV, true, 256, true>(tb, stream);
      return rf_launch<DT, NV, true, 512, true>(tb, stream);
    }
    if (G == 32) return rf_launch<DT, NV, false, 256, false>(tb, stream);
    if (G <= 256) return rf_launch<DT, NV, true, 256, false>(tb, stream);
    return rf_launch<DT, NV, true, 512, false>(tb, stream);
  } else {
    if (G == 32) return rf_launch<DT, NV, false, 256, true>(tb, stream);
    if (G <= 256) return rf_launch<DT, NV, true, 256, true>(tb, stream);
    return rf_launch<DT, NV, true, 512, true>(tb, stream);
  }
}

// Requires (every matrix of the batch): the same C with C % 8 == 0, 16-byte aligned rows, C <= 32768.
template <int DT>
static int run_row_select_fast(RfBatch& tb, int nv_max, bool keep, cudaStream_t stream) {
  const int64_t C = tb.C;
  const int64_t nvec = C / 8;
  int G = 32;
  while (G < 512 && (nvec + G - 1) / G > nv_max) G <<= 1;
  int nv = (int)((nvec + G - 1) / G);
  ECF_REQUIRE(nv <= 8, ECF_ERR_INVALID, "row_select: C=%lld exceeds the supported row length 32768", (long long)C);
  if (nv == 7) nv = 8;
  tb.G = G;
  switch (nv) {
#define ECF_CASE(N) \
  case N: return rf_dispatch_g<DT, N>(tb, keep, stream);
    ECF_CASE(1) ECF_CASE(2) ECF_CASE(3) ECF_CASE(4) ECF_CASE(5) ECF_CASE(6) ECF_CASE(8)
#undef ECF_CASE
  }
  set_error("row_select: unsupported vectors-per-lane %d", nv);
  return ECF_ERR_INVALID;
}

}  // namespace ecf
