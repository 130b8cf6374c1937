// A3+A4+A7 -- per-row Wanda select, round-3 kernel: one WARP owns a row, rows travel HBM -> shared memory as 1-D bulk
// TMA copies (cp.async.bulk + mbarrier), the bracket around the k-th score is carried from row to row.
//
// Same result as row_select_fast.cuh (bit for bit torch.sort(stable=True)[:, :k] + scatter_ + W[mask] = 0,
// LAVIS/lavis/compression/pruners/wanda_pruner.py:260,272-279); what changed is where the time went:
//   * round 2 measured 1 709 warp instructions per 2 048-element row at 56 % issue utilisation (16 warps per SM, the
//     loads of a row held in registers).  Here the next row is in flight as ONE bulk copy per warp that occupies no
//     registers, the weights are re-read from shared memory in the apply pass instead of being kept (raw[] gone), and
//     the kernel runs at 64 registers: 4 CTAs = 32 warps per SM for C <= 2048.  The buffer is free after the apply pass
//     (bracket elements are re-read from L2), so the next row's copy has the gather / ranking / patch to land.
//   * the rows of a weight matrix have nearly the same score distribution.  A row therefore starts from the previous
//     row's result: one counting pass at the previous threshold, then one at the key the local density (coarse keys per
//     element, carried as well) predicts for rank k +- 16 on the other side of it.  Most rows are bracketed to <= 32
//     elements after two or three pivot evaluations (4.2 on average, the seeded first rows included); round 2 needed a
//     32-sample sort plus 6-7.  Whatever the guess, the bracket invariant  #(key < lo) < k <= #(key < hi)  comes from
//     exact counts, so the carry only affects speed, never the result; a bad guess falls through to the interpolation /
//     bisection loop of round 2.
//   * the bracket is finished without atomics: a prefix sum over the lanes' candidate counts gives every candidate a
//     slot, the lanes list the COLUMNS of their candidates, then candidate t gets its exact key from lane t; brackets
//     that span <= 8 coarse keys rank one 32-bit word per candidate (19 key bits + 13 column bits).
//   * a CTA owns a contiguous range of ROWS (not of 8-row batches) of the concatenated matrices of the launch, so that
//     the warps of the grid differ by at most one row.
//   * next to a short-row launch (row_select.cu forks a second stream) the long-row variant runs as four-warp CTAs, one per
//     SM: both kernels then share every SM and the long rows' latency hides behind the short rows' work.
// Requirements (host side checks them, everything else takes row_select_fast.cuh): 16-bit weights, C = NV * 256 exactly,
// 16-byte aligned rows, no packed-mask / zero-count outputs.
// Bound: HBM.  Algorithmic bytes per call: 2*R*C*sizeof(w) + 4*C.
#pragma once

#include "row_select_fast.cuh"
#include "umma.cuh"

namespace ecf {

using umma::fence_barrier_init;
using umma::mbar_expect_tx;
using umma::mbar_init;
using umma::mbar_wait;
using umma::smem_u32;

constexpr int kRtWarps = 8;    // rows in flight per CTA
constexpr int kRtCap = 64;     // exact-ranking capacity per row (brackets end at <= 32 elements unless they are one coarse key)
constexpr int kRtBand = 32;    // stop narrowing once the bracket holds this many elements

struct RtShared {
  uint64_t mbar[kRtWarps][2];
  unsigned long long thr[kRtWarps];
  uint16_t cols[kRtWarps][kRtCap];
  alignas(16) unsigned long long cand[kRtWarps][kRtCap + 2];  // key << 32 | column, or packed 32-bit candidates (+ padding)
};

__device__ __forceinline__ void rt_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  mbar_expect_tx(bar, bytes);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// q = sqrt(scaler_row) of column `col` in the split layout: columns 8v..8v+3 of vector v at qs[4v..], columns 8v+4..8v+7
// at qs[C/2 + 4v..] (two conflict-free 128-bit reads per vector)
template <int C>
__device__ __forceinline__ float rt_q(const float* qs, uint32_t col) {
  return qs[((col & 4u) ? C / 2 : 0) + ((col >> 3) << 2) + (col & 3u)];
}

// Lane-local walk over this lane's bracket elements (bit e of `bm`: element ebase + e of the lane: vector e >> 3, slot e & 7).
//   MODE 5: write the column to slots b, b + 1, ... of `cols` (returns the next free slot)
//   MODE 1: count keys < a          MODE 2: count keys == a with column < b
//   MODE 3: zero the element in global memory when (key, column) <= thr
template <int DT, int C, int MODE>
__device__ __forceinline__ int rt_walk(uint32_t bm, int ebase, const unsigned char* buf, char* wrow, const float* qs, int lane, uint32_t a,
                                       uint32_t b, unsigned long long thr, unsigned long long* cand, uint16_t* cols) {
  int c = 0;
  if constexpr (MODE == 5) {
    if (bm == 0) return (int)b;
  }
  while (bm) {
    const int e = ebase + __ffs((int)bm) - 1;
    bm &= bm - 1;
    const uint32_t col = (uint32_t)(((e >> 3) * 32 + lane) * 8 + (e & 7));
    const float w = load_elem<DT>(buf, col);
    const uint32_t key = score_key(wanda_score(w, rt_q<C>(qs, col)));
    if constexpr (MODE == 5) {
      cols[b + c] = (uint16_t)col;
      ++c;
    } else if constexpr (MODE == 1) {
      c += key < a ? 1 : 0;
    } else if constexpr (MODE == 2) {
      c += (key == a && col < b) ? 1 : 0;
    } else {
      if ((((unsigned long long)key << 32) | col) <= thr) store_zero<DT>(wrow, col);
    }
  }
  if constexpr (MODE == 5) return (int)b + c;
  return c;
}

template <int DT, int NV, int STAGES, int MINB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, MINB) row_select_tma_kernel(const __grid_constant__ RfBatch tb) {
  constexpr int NP = 4 * NV;     // packed pairs per lane
  constexpr int C = NV * 256;    // row length
  constexpr uint32_t ROWB = C * 2;
  constexpr int NBM = (NV + 3) / 4;  // 32-element bitmask words per lane
  extern __shared__ __align__(128) unsigned char rt_dsm[];
  float* qs = reinterpret_cast<float*>(rt_dsm);
  __shared__ RtShared sh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* mybuf = rt_dsm + (size_t)C * 4 + (size_t)warp * STAGES * ROWB;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&sh.mbar[warp][s], 1);
  }
  fence_barrier_init();
  __syncthreads();

  // this CTA's rows [g0, g1) of the concatenated matrices
  const int total = tb.total_batches;
  const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int g0 = min(total, (int)blockIdx.x * per), g1 = min(total, g0 + per);
  uint32_t phase = 0;
  int stage = 0;

  for (int mi = 0; mi < tb.n; ++mi) {
    const RfMat& M = tb.m[mi];
    const int mb = M.batch_begin;
    const int ra = max(g0, mb) - mb, rb = min(g1, mb + (int)M.R) - mb;
    if (ra >= rb) continue;  // uniform over the CTA
    const int k = M.k;
    const int64_t ld = M.ld;
    // the first row of every warp starts its trip now (its buffer is free: the previous matrix' last row ended with a
    // __syncwarp); the table of sqrt(scaler_row) is built underneath that latency
    int r = ra + warp;
    char* wrow = reinterpret_cast<char*>(M.W) + (int64_t)r * ld * 2;
    const int64_t row_step = (int64_t)WARPS * ld * 2;
    if (r < rb && lane == 0) rt_bulk_load(mybuf + stage * ROWB, wrow, ROWB, &sh.mbar[warp][stage]);
    {
      constexpr int QV = C / 4, QIT = (QV + WARPS * 32 - 1) / (WARPS * 32);
      float4 s4[QIT];
#pragma unroll
      for (int it = 0; it < QIT; ++it) {  // all loads in flight before the first use
        const int idx = tid + it * WARPS * 32;
        s4[it] = idx < QV ? __ldg(reinterpret_cast<const float4*>(M.s) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();  // every warp is done with the previous matrix' table
#pragma unroll
      for (int it = 0; it < QIT; ++it) {
        const int idx = tid + it * WARPS * 32;  // columns 4 idx .. 4 idx + 3: half (idx & 1) of vector idx >> 1
        if (idx < QV)
          *reinterpret_cast<float4*>(qs + ((idx & 1) ? C / 2 : 0) + ((idx >> 1) << 2)) =
              make_float4(__fadd_rn(sqrtf(s4[it].x), 0.f), __fadd_rn(sqrtf(s4[it].y), 0.f), __fadd_rn(sqrtf(s4[it].z), 0.f),
                          __fadd_rn(sqrtf(s4[it].w), 0.f));
      }
      __syncthreads();
    }
    bool carry = false;
    uint32_t t_prev = 0;
    float inv_rho = 1.f;  // coarse-key units per element near the threshold

    for (; r < rb; r += WARPS, wrow += row_step) {
      const bool has_next = r + WARPS < rb;
      if constexpr (STAGES == 2) {
        // the other buffer was last read before the __syncwarp that ended the previous row
        if (has_next && lane == 0) rt_bulk_load(mybuf + (stage ^ 1) * ROWB, wrow + row_step, ROWB, &sh.mbar[warp][stage ^ 1]);
      }
      mbar_wait(&sh.mbar[warp][stage], (phase >> stage) & 1u);
      phase ^= 1u << stage;
      const unsigned char* buf = mybuf + stage * ROWB;

      // ---- exact scores -> packed coarse keys (upper 16 bits of the fp32 score, clamped to the largest finite fp16 pattern) ----
      uint32_t co[NP];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = i * 32 + lane;
        const uint4 a = *reinterpret_cast<const uint4*>(buf + v * 16);
        const float4 qa = *reinterpret_cast<const float4*>(qs + 4 * v), qb = *reinterpret_cast<const float4*>(qs + C / 2 + 4 * v);
        const uint32_t rv[4] = {a.x, a.y, a.z, a.w};
        const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float w0, w1;
          unpack2<DT>(rv[j], w0, w1);
          co[4 * i + j] = rf_coarse_pair(w0, w1, q[2 * j], q[2 * j + 1]);
        }
        if ((i & 1) == 1) asm volatile("" ::: "memory");  // keep at most two vectors' loads live (register budget)
      }

      // ---- coarse bracket [lo, hi):  #(coarse < lo) < k <= #(coarse < hi) ----
      uint32_t lo = 0, hi = 0;
      int c_lo = 0, c_hi = 0;
      if (k >= C) {
        lo = hi = kRfInf;  // every element is below the bracket
      } else if (k > 0) {
        lo = 0; c_lo = 0; hi = kRfInf; c_hi = C;
        if (carry) {
          // pass 1: where does rank k sit relative to the previous row's threshold?
          const uint32_t p = min(max(t_prev, 1u), kRfInf - 1u);
          const int c = warp_sum(rf_count1<NP>(co, p));
          if (c < k) { lo = p; c_lo = c; } else { hi = p; c_hi = c; }
          // pass 2: one pivot at the key the carried density predicts for rank k + band/2 (k-th above p) or k - band/2
          // (k-th below p): the bracket between the two pivots then holds |k - c| + band/2 elements
          const float dist = ((float)abs(k - c) + (float)(kRtBand / 2)) * inv_rho;
          const int step = max(__float2int_ru(dist), 1);
          if (c < k) {
            const uint32_t p2 = min(p + (uint32_t)step, kRfInf - 1u);
            if (p2 > lo) {
              const int c2 = warp_sum(rf_count1<NP>(co, p2));
              if (c2 < k) { lo = p2; c_lo = c2; } else { hi = p2; c_hi = c2; }
            }
          } else if ((int)p - step >= 1) {
            const uint32_t p2 = p - (uint32_t)step;
            const int c2 = warp_sum(rf_count1<NP>(co, p2));
            if (c2 < k) { lo = p2; c_lo = c2; } else { hi = p2; c_hi = c2; }
          }
        } else {
          // first row of this warp in this matrix: 32 samples, bitonic sort, two pivots around the sample quantile
          uint32_t v = (lane & 1) ? (co[NP - 1] >> 16) : (co[0] & 0xffffu);
#pragma unroll
          for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
              const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
              const bool up = (lane & kk) == 0, lower = (lane & j) == 0;
              v = (up == lower) ? min(v, o) : max(v, o);
            }
          }
          const int rk = (int)(((int64_t)k * 32) / C);
          const int il = rk - 3, ih = rk + 3;
          const uint32_t vlo = __shfl_sync(0xffffffffu, v, il < 0 ? 0 : il);
          const uint32_t vhi = __shfl_sync(0xffffffffu, v, ih > 31 ? 31 : ih);
          const uint32_t pl = il >= 0 ? vlo : 0u, ph = ih < 32 ? vhi + 1u : kRfInf;
          const int cc = warp_sum(rf_count2<NP>(co, pl, ph));
          const int cl = cc & 0xffff, ch = cc >> 16;
          if (cl >= k) { hi = pl; c_hi = cl; }
          else if (ch < k) { lo = ph; c_lo = ch; }
          else { lo = pl; c_lo = cl; hi = ph; c_hi = ch; }
        }
        int stall = 0;
        while (c_hi - c_lo > kRtBand && hi - lo > 1) {
          const int n = c_hi - c_lo;
          uint32_t p;
          if (stall >= 2) {
            p = (lo + hi) >> 1;
            stall = 0;
          } else {
            const int dl = k - c_lo, dh = c_hi - k;
            const float target = (float)dl + (dl >= dh ? -(float)(kRtBand / 3) : (float)(kRtBand / 3));
            const float f = __saturatef(__fdividef(target, (float)n));
            p = lo + (uint32_t)__float2int_rn((float)(hi - lo) * f);
          }
          p = min(max(p, lo + 1), hi - 1);
          const int c = warp_sum(rf_count1<NP>(co, p));
          if (c < k) { lo = p; c_lo = c; } else { hi = p; c_hi = c; }
          stall = (c_hi - c_lo) * 2 > n ? stall + 1 : 0;
        }
        // carry for the next row: the middle of the bracket and the density inside it (smoothed)
        if (hi - lo <= 256u && c_hi > c_lo) {
          const float d = __fdividef((float)(hi - lo), (float)(c_hi - c_lo));
          inv_rho = carry ? 0.75f * inv_rho + 0.25f * d : d;
          inv_rho = fminf(fmaxf(inv_rho, 1.f / 4096.f), 64.f);
          t_prev = (lo + hi) >> 1;
          carry = true;
        } else {
          carry = false;
        }
      }

      // ---- coarse apply (everything below the bracket goes) + store; record this lane's bracket elements ----
      uint32_t bm[NBM];
#pragma unroll
      for (int i = 0; i < NBM; ++i) bm[i] = 0;
      {
        const __half2 pl = rf_h2(rf_dup(lo)), ph = rf_h2(rf_dup(hi));
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = i * 32 + lane;
          const uint4 a = *reinterpret_cast<const uint4*>(buf + v * 16);
          const uint32_t rv[4] = {a.x, a.y, a.z, a.w};
          uint32_t o[4];
          uint32_t bacc = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int p = 4 * i + j;
            const uint32_t ml = __hlt2_mask(rf_h2(co[p]), pl);
            const uint32_t mh = __hlt2_mask(rf_h2(co[p]), ph);
            bacc |= (mh ^ ml) & ((1u << (2 * j)) | (0x10000u << (2 * j + 1)));
            o[j] = rv[j] & ~ml;
          }
          bm[i >> 2] |= ((bacc | (bacc >> 16)) & 0xffu) << (8 * (i & 3));
          stg_v4(wrow + v * 16, make_uint4(o[0], o[1], o[2], o[3]));
          if ((i & 1) == 1) asm volatile("" ::: "memory");
        }
      }

      // The row buffer is free from here on (bracket elements are re-read from global memory: the vector stores above left
      // them untouched and this warp's lines are L2 hits), so with one stage the next row's copy starts now and has the
      // gather / ranking / patch of this row to land.
      __syncwarp();  // every lane has read the buffer; also orders the vector stores before the scalar reads / patches below
      if constexpr (STAGES == 1) {
        if (has_next && lane == 0) rt_bulk_load(mybuf, wrow + row_step, ROWB, &sh.mbar[warp][0]);
      }
      const unsigned char* src = reinterpret_cast<const unsigned char*>(wrow);

      // ---- exact ranking of the bracket by (fp32 key, column) ----
      const int m = c_hi - c_lo, need = k - c_lo;  // bracket size, how many of it must go (1 <= need <= m when m > 0)
      if (m > 0) {
        if (m <= kRtCap) {
          // slots of this lane's candidates: exclusive prefix sum of the per-lane counts (no atomics)
          int mine = 0;
#pragma unroll
          for (int i = 0; i < NBM; ++i) mine += __popc(bm[i]);
          int incl = mine;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          int slot = incl - mine;
          // When the bracket spans at most 8 coarse keys below the clamp, (exact key - lo * 2^16) fits 19 bits and the column
          // 13: a candidate is ONE 32-bit word and the ranking loop costs half.
          const bool packed = hi - lo <= 8u && hi <= 0x7bffu;
          unsigned long long* cand = sh.cand[warp];
          uint32_t* cand32 = reinterpret_cast<uint32_t*>(cand);
          // gather in two steps: the lanes first list the COLUMNS of their bracket elements (a short lane-local loop over
          // set bits, as many trips as the fullest lane has candidates), then candidate t gets its exact key from lane
          // t mod 32 -- one parallel step instead of one per trip
          uint16_t* cols = sh.cols[warp];
#pragma unroll
          for (int i = 0; i < NBM; ++i) slot = rt_walk<DT, C, 5>(bm[i], 32 * i, src, wrow, qs, lane, 0, (uint32_t)slot, 0ull, cand, cols);
          __syncwarp();
          for (int t = lane; t < m; t += 32) {
            const uint32_t col = cols[t];
            const uint32_t key = score_key(wanda_score(load_elem<DT>(src, col), rt_q<C>(qs, col)));
            if (packed) cand32[t] = ((key - (lo << 16)) << 13) | col;
            else cand[t] = ((unsigned long long)key << 32) | col;
          }
          if (packed && lane < 4) cand32[m + lane] = 0xffffffffu;  // padding up to a multiple of 4
          __syncwarp();  // also orders every lane's vector stores before the scalar patches below
          if (packed) {
            uint32_t thr = 0xffffffffu;  // need == m: the whole bracket goes
            if (need < m) {
              for (int t = lane; t < m; t += 32) {
                const uint32_t me = cand32[t];
                int rank = 0;
#pragma unroll 2
                for (int j = 0; j < m; j += 4) {
                  const uint4 o = *reinterpret_cast<const uint4*>(cand32 + j);
                  rank += (o.x < me ? 1 : 0) + (o.y < me ? 1 : 0) + (o.z < me ? 1 : 0) + (o.w < me ? 1 : 0);
                }
                if (rank == need - 1) sh.thr[warp] = me;
              }
              __syncwarp();
              thr = (uint32_t)sh.thr[warp];
            }
            for (int t = lane; t < m; t += 32) {
              const uint32_t me = cand32[t];
              if (me <= thr) store_zero<DT>(wrow, me & 0x1fffu);
            }
          } else {
            unsigned long long thr = ~0ull;
            if (need < m) {
              for (int t = lane; t < m; t += 32) {
                const unsigned long long me = cand[t];
                int rank = 0;
#pragma unroll 4
                for (int j = 0; j < m; ++j) rank += cand[j] < me ? 1 : 0;
                if (rank == need - 1) sh.thr[warp] = me;
              }
              __syncwarp();
              thr = sh.thr[warp];
            }
            for (int t = lane; t < m; t += 32) {
              const unsigned long long me = cand[t];
              if (me <= thr) store_zero<DT>(wrow, (uint32_t)me);
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
              int c = 0;
#pragma unroll
              for (int i = 0; i < NBM; ++i) c += rt_walk<DT, C, 1>(bm[i], 32 * i, src, wrow, qs, lane, pv, 0, 0ull, nullptr, nullptr);
              c = warp_sum(c);
              if (c < need) { L = pv; cL = c; } else { H = pv; }
            }
            const int need2 = need - cL;        // ties at key L that must go, lowest column first
            uint32_t CL = 0, CH = (uint32_t)C;  // #(ties with col < CL) < need2 <= #(ties with col < CH)
            while (CH - CL > 1) {
              const uint32_t pc = (CL + CH) >> 1;
              int c = 0;
#pragma unroll
              for (int i = 0; i < NBM; ++i) c += rt_walk<DT, C, 2>(bm[i], 32 * i, src, wrow, qs, lane, L, pc, 0ull, nullptr, nullptr);
              c = warp_sum(c);
              if (c < need2) CL = pc; else CH = pc;
            }
            thr = ((unsigned long long)L << 32) | CL;
          }
          // scalar patch of this lane's own bracket elements (after its own vector stores: same-thread order)
#pragma unroll
          for (int i = 0; i < NBM; ++i) rt_walk<DT, C, 3>(bm[i], 32 * i, src, wrow, qs, lane, 0, 0, thr, nullptr, nullptr);
        }
      }
      __syncwarp();  // every lane is done with this row's candidate list and threshold slot
      if constexpr (STAGES == 2) stage ^= 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------- host
// share != 0: this launch runs NEXT TO another row-select launch on a second stream (row_select.cu).  The grid is then capped
// at `share` CTAs per SM and the dynamic shared memory padded to `pad_smem`, so that the two kernels split every SM the same
// way whichever is dispatched first.
template <int DT, int NV, int STAGES, int MINB, int WARPS = kRtWarps>
static int rt_launch(RfBatch& tb, int share, size_t pad_smem, cudaStream_t stream) {
  static_assert(WARPS <= kRtWarps, "RtShared is sized for kRtWarps rows per CTA");
  auto kern = row_select_tma_kernel<DT, NV, STAGES, MINB, WARPS>;
  size_t smem = (size_t)NV * 256 * 4 + (size_t)WARPS * STAGES * NV * 512;
  if (share != 0 && smem < pad_smem) smem = pad_smem;
  static size_t opted = 0;
  if (smem > opted) {
    ECF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the largest shared-memory carve-out whatever the launch needs: two kernels only share an SM when they agree on it
    ECF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    opted = smem;
  }
  int occ = 0;
  ECF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  if (occ < 1) occ = 1;
  if (share != 0 && occ > share) occ = share;
  int64_t rows = 0;
  for (int i = 0; i < tb.n; ++i) {
    tb.m[i].batch_begin = (int)rows;
    rows += tb.m[i].R;
  }
  ECF_REQUIRE(rows < (1ll << 31), ECF_ERR_INVALID, "row_select: too many rows in one launch");
  tb.total_batches = (int)rows;
  const int64_t want = (rows + WARPS - 1) / WARPS, cap = (int64_t)sm_count() * occ;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  kern<<<grid, WARPS * 32, smem, stream>>>(tb);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

template <int DT>
static int run_row_select_tma(RfBatch& tb, int stages, int share, size_t pad_smem, cudaStream_t stream) {
  switch (tb.C) {
    case 768: return stages == 2 ? rt_launch<DT, 3, 2, 4>(tb, share, pad_smem, stream) : rt_launch<DT, 3, 1, 4>(tb, share, pad_smem, stream);
    case 1024: return stages == 2 ? rt_launch<DT, 4, 2, 4>(tb, share, pad_smem, stream) : rt_launch<DT, 4, 1, 4>(tb, share, pad_smem, stream);
    case 2048: return stages == 2 ? rt_launch<DT, 8, 2, 3>(tb, share, pad_smem, stream) : rt_launch<DT, 8, 1, 4>(tb, share, pad_smem, stream);
    // long rows next to a short-row launch (share == 1): CTAs of four warps, one per SM, so that the short-row kernel keeps three
    case 3072: return share == 1 ? rt_launch<DT, 12, 1, 4, 4>(tb, share, pad_smem, stream) : rt_launch<DT, 12, 1, 2>(tb, share, pad_smem, stream);
    case 4096: return share == 1 ? rt_launch<DT, 16, 1, 4, 4>(tb, share, pad_smem, stream) : rt_launch<DT, 16, 1, 2>(tb, share, pad_smem, stream);
    case 5120: return share == 1 ? rt_launch<DT, 20, 1, 4, 4>(tb, share, pad_smem, stream) : rt_launch<DT, 20, 1, 2>(tb, share, pad_smem, stream);
  }
  set_error("row_select: row length %d is not served by the bulk-copy kernel", tb.C);
  return ECF_ERR_INVALID;
}

}  // namespace ecf
