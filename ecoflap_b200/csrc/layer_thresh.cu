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
//   P1  sample   1 vector in S of every matrix (all matrices in one flat index space) -> 31 744-bin histogram of the
//                upper 16 key bits (global REDs on <= 131 072 samples per matrix); also q = sqrt(scaler_row)
//   P2  bracket  one CTA per matrix scans its histogram: coarse bracket [lo, hi) = sample ranks k_s -+ 2.5 sqrt(n_s)
//   P3  count    the only HBM read of W.  Per 8-element vector: exact fp32 scores -> upper 16 bits packed two per
//                register (clamped to the finite fp16 range, so HSET2 compares them like integers) -> #(key < lo)
//                by mask popcounts, and "has an element in [lo, hi)" as one bit.  Vectors with that bit (~10 %) are
//                only APPENDED to a per-CTA shared-memory list; the CTA then walks its list with all lanes busy and
//                histograms the top 11-bit digit of the ~1-2 % of elements inside the bracket (shared atomics are
//                affordable there, a full-matrix histogram is not: ~2 clk per element per SM)
//   P4  refine   the next digit(s) come from the SAME per-CTA lists (no second pass over W)
//   P5  apply    score <= thres -> zero, in place (L2 read, HBM write); optional packed mask / zero count
// A CTA whose list overflows (heavy ties inside the bracket) histograms those vectors on the spot and re-scans its
// share of W for the later digits -- slower, still exact.  If the k-th score falls outside the sampled bracket
// (probability ~1e-6, or adversarial ties) the bracket is replaced by the side that holds it and P3 is repeated.
// Bound: HBM.  Algorithmic bytes per matrix: 2*R*C*sizeof(w) + 4*C (re-reads are L2 hits).
#include <cstdlib>

#include "common.cuh"
#include "layer_cut.cuh"

namespace ecf {

constexpr int kLtThreads = 512;
constexpr int kLtMaxMat = ECF_LAYER_MAX_BATCH;
constexpr int kLtCoarseBins = 32768;
constexpr int kLtBins = 2048;
constexpr int kLtSampleVecs = 16384;  // sampled 8-element vectors per matrix (131 072 scores)
constexpr int kLtListCap = 6144;      // bracket-vector list per CTA (24 KB of shared memory)
constexpr uint32_t kLtTop = 0x7c00u;  // end of the coarse key domain: keys are clamped to the finite fp16 patterns
constexpr int kLtSampWords = kLtTop / 2;        // per-CTA sample histogram: 16-bit counters packed two per word (62 KB)
constexpr int kLtSampChunk = 4096;              // sampled vectors per CTA between flushes (8 * 4096 < 65 536 per bin)
constexpr int kLtHeaderBytes = 8192;            // workspace header: phase stamps, grid barrier state
constexpr int kLtBarMaxGroups = 32;             // grid barrier: two-level arrival tree, one counter per 128 bytes

struct LtMat {
  void* W;
  const float* s;
  float* q;                 // workspace: sqrt(scaler_row) + 0
  unsigned* coarse;         // workspace: [kLtCoarseBins] sample histogram (self-cleaning)
  unsigned* hist;           // workspace: [3][kLtBins] digit histograms (self-cleaning)
  unsigned long long* cnt;  // workspace: [0] #(key < lo), [1] #(lo <= key < hi)
  uint32_t* bracket;        // workspace: [0] lo, [1] hi (coarse, hi exclusive, <= kLtTop), [3] P1 ticket, [4] ~min / [5] max+1 occupied sample bin
  float* thres_out;
  uint8_t* mask;
  unsigned long long* n_zero;
  int64_t R, C, ld, mask_ld;
  int64_t kth;
  int64_t nvpr;             // vectors per row = ceil(C / 8)
  int64_t nvec;             // R * nvpr
  int64_t vec_begin;        // prefix over the matrices of the launch
  int64_t sample_stride;    // S
  int64_t sample_begin;     // prefix of the sampled vectors over the matrices
  int64_t col_begin;        // prefix of C over the matrices (q-table work split)
  int64_t step_rows, step_cols;  // grid stride (threads) = step_rows * nvpr + step_cols
  int dtype, aligned;
};

struct LtBatch {
  LtMat m[kLtMaxMat];
  int n;
  int64_t total_vec, total_cols, total_samples;
  unsigned long long* stamps;  // workspace: %globaltimer of CTA 0 at the phase boundaries (profiling aid)
  unsigned* bar;               // workspace: [0] generation, [1] top counter, [32 * (1 + g)] counter of arrival group g
  unsigned* fallback;          // workspace: set by the split path when it gives a block up
  int after_split;             // launched behind the split path: do nothing unless *fallback != 0
};

__device__ __forceinline__ void lt_stamp(const LtBatch& b, int i) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    b.stamps[i] = t;
  }
}

__device__ __forceinline__ __half2 lt_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t lt_dup(uint32_t p) { return p | (p << 16); }

// position of a vector inside the launch: matrix, row, column (in 8-element vectors)
struct LtCur {
  int mi;
  uint32_t row, col;  // R and ceil(C / 8) are checked to fit 31 bits on the host
};

__device__ __forceinline__ void lt_seek(const LtBatch& b, int64_t v, LtCur& c) {
  int mi = c.mi;
  while (v >= b.m[mi].vec_begin + b.m[mi].nvec) ++mi;
  c.mi = mi;
  const LtMat& M = b.m[mi];
  const int64_t lv = v - M.vec_begin;
  if (M.nvec < (1ll << 31)) {  // 32-bit division: every BASELINE.json matrix
    const uint32_t r32 = (uint32_t)lv / (uint32_t)M.nvpr;
    c.row = r32;
    c.col = (uint32_t)lv - r32 * (uint32_t)M.nvpr;
  } else {
    const int64_t r64 = lv / M.nvpr;
    c.row = (uint32_t)r64;
    c.col = (uint32_t)(lv - r64 * M.nvpr);
  }
}
// advance by one grid stride; v is the NEW flat index (seek again when it leaves the matrix)
__device__ __forceinline__ void lt_next(const LtBatch& b, int64_t v, LtCur& c) {
  const LtMat& M = b.m[c.mi];
  if (v >= M.vec_begin + M.nvec) {
    if (v < b.total_vec) lt_seek(b, v, c);
    return;
  }
  c.col += (uint32_t)M.step_cols;
  c.row += (uint32_t)M.step_rows;
  if (c.col >= (uint32_t)M.nvpr) {
    c.col -= (uint32_t)M.nvpr;
    ++c.row;
  }
}

// the 16 / 32 bytes of an aligned vector, loaded ahead of their use (software pipelining of the streaming passes)
struct LtRaw {
  uint4 a;  // first 16 bytes of the vector (all of it for 16-bit types; fp32 vectors fetch their second half at use)
};
__device__ __forceinline__ void lt_issue(const LtMat& M, const LtCur& c, LtRaw& r) {
  if (!M.aligned) return;
  const int eb = M.dtype == ECF_F32 ? 4 : 2;
  const char* p = reinterpret_cast<const char*>(M.W) + ((int64_t)c.row * M.ld + (int64_t)c.col * 8) * eb;
  r.a = ldg_v4(p);
}

// Bit patterns u[e] of the exact scores fp32(|w|) * q of vector (row, col); out-of-range columns (ragged C) get
// 0xffffffff.  raw[] = the weight words (aligned: 4 packed words for 16-bit types, 8 words for fp32), taken from the
// pre-issued load `r` when the matrix is aligned.
template <int DT, bool ALIGNED, bool SQRT_INLINE>
__device__ __forceinline__ void lt_score_bits(const LtMat& M, uint32_t row, uint32_t col, const LtRaw& r, uint32_t (&u)[8],
                                              uint32_t (&raw)[8]) {
  const int64_t c0 = (int64_t)col * 8;
  if constexpr (ALIGNED) {
    float q[8];
    if constexpr (SQRT_INLINE) {
      const float4 sa = *reinterpret_cast<const float4*>(M.s + c0), sb = *reinterpret_cast<const float4*>(M.s + c0 + 4);
      const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) q[j] = __fadd_rn(sqrtf(sv[j]), 0.f);
    } else {
      const float4 qa = *reinterpret_cast<const float4*>(M.q + c0), qb = *reinterpret_cast<const float4*>(M.q + c0 + 4);
      q[0] = qa.x; q[1] = qa.y; q[2] = qa.z; q[3] = qa.w; q[4] = qb.x; q[5] = qb.y; q[6] = qb.z; q[7] = qb.w;
    }
    if constexpr (DT == ECF_F32) {
      const uint4 rb = ldg_v4(reinterpret_cast<const char*>(M.W) + ((int64_t)row * M.ld + c0) * 4 + 16);
      raw[0] = r.a.x; raw[1] = r.a.y; raw[2] = r.a.z; raw[3] = r.a.w; raw[4] = rb.x; raw[5] = rb.y; raw[6] = rb.z; raw[7] = rb.w;
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = __float_as_uint(wanda_score(__uint_as_float(raw[j]), q[j]));
    } else {
      raw[0] = r.a.x; raw[1] = r.a.y; raw[2] = r.a.z; raw[3] = r.a.w;
      raw[4] = raw[5] = raw[6] = raw[7] = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float w0, w1;
        unpack2<DT>(raw[j], w0, w1);
        u[2 * j] = __float_as_uint(wanda_score(w0, q[2 * j]));
        u[2 * j + 1] = __float_as_uint(wanda_score(w1, q[2 * j + 1]));
      }
    }
  } else {
    const char* wrow = reinterpret_cast<const char*>(M.W) + (int64_t)row * M.ld * DType<DT>::kBytes;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      raw[j] = 0;
      if (c0 + j < M.C) {
        const float w = load_elem<DT>(wrow, c0 + j);
        const float q = SQRT_INLINE ? __fadd_rn(sqrtf(M.s[c0 + j]), 0.f) : M.q[c0 + j];
        u[j] = __float_as_uint(wanda_score(w, q)) & 0x7fffffffu;  // keep real scores apart from the padding pattern
      } else {
        u[j] = 0xffffffffu;
      }
    }
  }
}

// upper 16 bits of the 8 scores, packed two per word, clamped to the finite fp16 patterns (padding -> +inf pattern)
template <bool ALIGNED>
__device__ __forceinline__ void lt_coarse(const uint32_t (&u)[8], uint32_t (&co)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t p = __vminu2(__byte_perm(u[2 * j], u[2 * j + 1], 0x7632) & 0x7fff7fffu, 0x7bff7bffu);
    if constexpr (!ALIGNED) {
      if (u[2 * j] == 0xffffffffu) p = (p & 0xffff0000u) | kLtTop;
      if (u[2 * j + 1] == 0xffffffffu) p = (p & 0x0000ffffu) | (kLtTop << 16);
    }
    co[j] = p;
  }
}

// dispatch a statement on (dtype, aligned) of a matrix
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
  uint32_t lo, hi;          // coarse bracket
  uint32_t lo32, range_hi;  // bracket as fp32 keys: [lo32, lo32 + range)  (range_hi: range - 1, fits 32 bits)
  int nd;                   // 11-bit digits needed for (key - lo32)
  uint32_t prefix;          // digits fixed so far (right aligned)
  unsigned long long rem;   // rank still to resolve inside the prefix bucket
  int done;                 // digits resolved
  int active;               // participates in the current P3 round
};

__device__ __forceinline__ int lt_shift(int nd, int level) { return (nd - 1 - level) * 11; }

// P3 main path for one vector: 16 x #(key < lo) via mask popcounts; returns non-zero when an element is in [lo, hi)
template <int DT, bool AL>
__device__ __forceinline__ uint32_t lt_count_vec(const LtMat& M, uint32_t row, uint32_t col, const LtRaw& r, __half2 pl, __half2 ph,
                                                 int& pc16, uint32_t (&u)[8]) {
  uint32_t raw[8], co[4];
  lt_score_bits<DT, AL, false>(M, row, col, r, u, raw);
  lt_coarse<AL>(u, co);
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t ml = __hlt2_mask(lt_h2(co[j]), pl), mh = __hlt2_mask(lt_h2(co[j]), ph);
    x |= ml ^ mh;
    pc16 += __popc(ml);
  }
  return x;
}

// bracket elements of one vector -> digit histogram (level 0: every bracket element, counted in nb; level > 0: the
// elements whose higher digits equal `prefix`)
__device__ __forceinline__ void lt_hist_vec(const uint32_t (&u)[8], const LtSel& S, int level, unsigned* hist, unsigned& nb) {
  const int sh = lt_shift(S.nd, level);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const uint32_t key = score_key(__uint_as_float(u[e]));
    const uint32_t d = key - S.lo32;
    if (u[e] != 0xffffffffu && key >= S.lo32 && d <= S.range_hi) {
      if (level == 0) {
        ++nb;
        atomicAdd(hist + (d >> sh), 1u);
      } else if ((d >> (sh + 11)) == S.prefix) {
        atomicAdd(hist + ((d >> sh) & (kLtBins - 1)), 1u);
      }
    }
  }
}

// P5 for one vector: zero the elements with score bits < tcmp
template <int DT, bool AL>
__device__ __forceinline__ void lt_apply_vec(const LtMat& M, uint32_t row, uint32_t col, const LtRaw& r, uint32_t tcmp, int& zeros) {
  uint32_t u[8], raw[8];
  lt_score_bits<DT, AL, false>(M, row, col, r, u, raw);
  const int64_t c0 = (int64_t)col * 8;
  char* wrow = reinterpret_cast<char*>(M.W) + (int64_t)row * M.ld * DType<DT>::kBytes;
  // padding / NaN patterns are never below tcmp.  The packed mask byte is only assembled when somebody wants it.
  if constexpr (AL) {
    if constexpr (DT == ECF_F32) {
      uint32_t lo4 = 0, hi4 = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool p = u[e] < tcmp;
        if (p) raw[e] = 0;
        if (e < 4) lo4 |= p ? 1u : 0u; else hi4 |= p ? 1u : 0u;
      }
      if (lo4) stg_v4(wrow + c0 * 4, make_uint4(raw[0], raw[1], raw[2], raw[3]));
      if (hi4) stg_v4(wrow + c0 * 4 + 16, make_uint4(raw[4], raw[5], raw[6], raw[7]));
      if (M.n_zero != nullptr) {
#pragma unroll
        for (int e = 0; e < 8; ++e) zeros += (raw[e] & 0x7fffffffu) == 0 ? 1 : 0;
      }
    } else {
      uint32_t any = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t z = (u[2 * j] < tcmp ? 0x0000ffffu : 0u) | (u[2 * j + 1] < tcmp ? 0xffff0000u : 0u);
        any |= z;
        raw[j] &= ~z;
      }
      if (any) stg_v4(wrow + c0 * 2, make_uint4(raw[0], raw[1], raw[2], raw[3]));
      if (M.n_zero != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) zeros += ((raw[j] & 0x00007fffu) == 0 ? 1 : 0) + ((raw[j] & 0x7fff0000u) == 0 ? 1 : 0);
      }
    }
    if (M.mask != nullptr) {
      uint32_t m = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) m |= (u[e] < tcmp ? 1u : 0u) << e;
      M.mask[(int64_t)row * M.mask_ld + col] = (uint8_t)m;
    }
  } else {
    uint32_t m = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) m |= (u[e] < tcmp ? 1u : 0u) << e;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (c0 + e < M.C) {
        const bool p = m >> e & 1;
        if (p) store_zero<DT>(wrow, c0 + e);
        if (M.n_zero != nullptr) zeros += (p || load_elem<DT>(wrow, c0 + e) == 0.f) ? 1 : 0;
      }
    }
    if (M.mask != nullptr) M.mask[(int64_t)row * M.mask_ld + col] = (uint8_t)m;
  }
}

// Grid-wide barrier for the co-resident (cooperative) grid.  cooperative_groups' grid.sync() funnels every CTA
// through ONE counter -- ~300 same-address atomics serialise in the L2 for ~6 us on B200.  Here CTAs arrive on one of
// ~sqrt(grid) group counters (distinct 128-byte lines), the last arrival of a group moves on to the top counter and
// the last one there publishes the new generation, which everybody polls.  Counters reset themselves.
__device__ __forceinline__ void lt_grid_barrier(unsigned* bar, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++gen;
    const unsigned nb = gridDim.x;
    unsigned gs = 1;
    while (gs * gs < nb) ++gs;
    if ((nb + gs - 1) / gs > (unsigned)kLtBarMaxGroups) gs = (nb + kLtBarMaxGroups - 1) / kLtBarMaxGroups;
    const unsigned ng = (nb + gs - 1) / gs;
    const unsigned g = blockIdx.x / gs;
    const unsigned members = min(gs, nb - g * gs);
    __threadfence();
    if (atomicAdd(bar + 32 * (1 + g), 1u) == members - 1) {
      bar[32 * (1 + g)] = 0u;
      __threadfence();
      if (atomicAdd(bar + 1, 1u) == ng - 1) {
        bar[1] = 0u;
        __threadfence();
        atomicExch(bar, gen);
      }
    }
    while (*reinterpret_cast<volatile unsigned*>(bar) != gen) {}
    __threadfence();
  }
  __syncthreads();
}

// Four warps per matrix: find the bin of its 2 048-bin digit histogram (level `level`) that holds rank sel.rem.
// Every lane reads its 16 bins with four coalesced 128-bit loads that are all in flight together and keeps them in
// registers for the search (one L2 round trip instead of a chain of them).
__device__ __forceinline__ void lt_find_bins(const LtBatch& b, LtSel* sel, int level, unsigned long long (*sh_q)[4]) {
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kPerRound = max(1, (int)(blockDim.x >> 7));  // matrices handled per round (four warps each)
  for (int m0 = 0; m0 < b.n; m0 += kPerRound) {
    const int mi = m0 + (wid >> 2), qt = wid & 3;
    const bool on = mi < b.n && sel[mi < b.n ? mi : 0].nd > level;
    uint4 v[4];
    unsigned long long inc[4], tot[4], quarter = 0;
    if (on) {
      const uint4* src = reinterpret_cast<const uint4*>(b.m[mi].hist + level * kLtBins + qt * (kLtBins / 4));
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldcg(src + j * 32 + lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unsigned long long x = (unsigned long long)v[j].x + v[j].y + v[j].z + v[j].w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long t = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += t;
        }
        inc[j] = x;
        tot[j] = __shfl_sync(0xffffffffu, x, 31);
        quarter += tot[j];
      }
      if (lane == 0) sh_q[wid >> 2][qt] = quarter;
    }
    __syncthreads();
    if (on) {
      unsigned long long run = 0;
      for (int q = 0; q < qt; ++q) run += sh_q[wid >> 2][q];
      const unsigned long long rem = sel[mi].rem;
      if (rem >= run && rem < run + quarter) {  // exactly one warp
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (rem >= run && rem < run + tot[j]) {  // exactly one chunk (warp-uniform)
            const unsigned cs[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            const unsigned long long mine = (unsigned long long)cs[0] + cs[1] + cs[2] + cs[3];
            unsigned long long r2 = run + inc[j] - mine;  // elements before this lane's first bin
            if (rem >= r2 && rem < r2 + mine) {  // exactly one lane
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (rem >= r2 && rem < r2 + cs[e]) {
                  sel[mi].prefix = (sel[mi].prefix << 11) | (uint32_t)(qt * (kLtBins / 4) + (j * 32 + lane) * 4 + e);
                  sel[mi].rem = rem - r2;
                  sel[mi].done = level + 1;
                }
                r2 += cs[e];
              }
            }
          }
          run += tot[j];
        }
      }
    }
    __syncthreads();
  }
}

// P2: the calling CTA turns the sample histogram of matrix M into a coarse bracket (and cleans the histogram).
// Only the occupied bin range [bmin, bmax) recorded by the sampling CTAs is scanned: typically a few hundred bins,
// i.e. one or two per thread, instead of 64.
__device__ __forceinline__ void lt_bracket(const LtMat& M, unsigned long long* sh_scan /*[36]*/) {
  const int tid = threadIdx.x;
  const uint32_t enc_min = __ldcg(M.bracket + 4), bmax = min(__ldcg(M.bracket + 5), (uint32_t)kLtTop);
  const uint32_t bmin = min(0xffffffffu - enc_min, bmax);
  const uint32_t per = (bmax - bmin + kLtThreads - 1) / kLtThreads;
  const uint32_t b0 = min(bmax, bmin + (uint32_t)tid * per), b1 = min(bmax, b0 + per);
  unsigned long long sum = 0;
  for (uint32_t bin = b0; bin < b1; bin += 8) {
    unsigned c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] = bin + u < b1 ? __ldcg(M.coarse + bin + u) : 0u;
#pragma unroll
    for (int u = 0; u < 8; ++u) sum += c[u];
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
    if (lane == kLtThreads / 32 - 1) {
      const unsigned long long ns = winc;  // total number of samples
      const double numel = (double)M.R * (double)M.C;
      const long long rs = (long long)((double)M.kth * (double)ns / numel);
      const long long delta = M.sample_stride > 1 ? (long long)(2.5f * sqrtf((float)ns)) + 4 : 0;
      reinterpret_cast<long long*>(sh_scan)[32] = rs - delta;
      reinterpret_cast<long long*>(sh_scan)[33] = rs + delta;
      M.bracket[0] = 0u;
      M.bracket[1] = kLtTop;
      M.bracket[4] = 0u;  // occupied range: empty again for the next launch
      M.bracket[5] = 0u;
      M.cnt[0] = 0ull;
      M.cnt[1] = 0ull;
    }
  }
  __syncthreads();
  const long long r_lo = reinterpret_cast<long long*>(sh_scan)[32], r_hi = reinterpret_cast<long long*>(sh_scan)[33];
  unsigned long long run = sh_scan[wid] + inc - sum;  // samples in the bins before this thread's first bin
  if (sum) {  // only threads whose bins hold samples can contain the two ranks
    for (uint32_t bin = b0; bin < b1; bin += 8) {
      unsigned c[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) c[u] = bin + u < b1 ? __ldcg(M.coarse + bin + u) : 0u;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (c[u]) {
          if (r_lo >= 0 && (unsigned long long)r_lo >= run && (unsigned long long)r_lo < run + c[u]) M.bracket[0] = bin + u;
          if (r_hi >= 0 && (unsigned long long)r_hi >= run && (unsigned long long)r_hi < run + c[u]) M.bracket[1] = bin + u + 1u;
          M.coarse[bin + u] = 0u;  // self-cleaning for the next launch
        }
        run += c[u];
      }
    }
  }
  __syncthreads();
}

// P1 + P2 for the calling grid of kLtThreads-wide CTAs: q tables, zeroed digit histograms, per-CTA shared-memory sample
// histograms -> global, and the coarse bracket by the last CTA of every matrix.  No grid barrier inside.
__device__ __forceinline__ void lt_phase_sample(const LtBatch& b, unsigned* sh_dyn, unsigned long long* sh_scan, unsigned* sh_range,
                                                int* sh_last_p) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t gthreads = (int64_t)gridDim.x * kLtThreads;
  const int64_t gtid = (int64_t)blockIdx.x * kLtThreads + tid;
  int& sh_last = *sh_last_p;
  // ================= P1: q tables + per-CTA sample histograms of the upper 16 key bits -> global ===================
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
  {
    // CTA c samples slice c / n of matrix c % n into a shared-memory histogram (16-bit counters, two per word) and
    // adds its non-empty bins to the global one: ~300 global REDs per CTA instead of one per sample on a few hundred
    // hot addresses.  The last CTA of a matrix (ticket) goes straight on to P2 for it: no grid barrier in between.
    if (tid == 0) sh_range[0] = sh_range[1] = 0u;
    const int nslices = max(1, (int)gridDim.x / b.n);
    const int mi = (int)blockIdx.x % b.n, slice = (int)blockIdx.x / b.n;
    if (slice < nslices) {
      const LtMat& M = b.m[mi];
      const int64_t nsv = (M.nvec + M.sample_stride - 1) / M.sample_stride;
      const int64_t per = (nsv + nslices - 1) / nslices;
      const int64_t j0 = min(nsv, (int64_t)slice * per), j1 = min(nsv, j0 + per);
      for (int64_t cb = j0; cb < j1; cb += kLtSampChunk) {
        for (int i = tid; i < kLtSampWords; i += kLtThreads) sh_dyn[i] = 0u;
        __syncthreads();
        const int64_t ce = min(j1, cb + (int64_t)kLtSampChunk);
        for (int64_t j = cb + tid; j < ce; j += kLtThreads) {
          int64_t lv = j * M.sample_stride;
          if (M.sample_stride > 1) lv += (int64_t)(((uint32_t)j * 2654435761u) >> 8) % M.sample_stride;
          if (lv >= M.nvec) lv = M.nvec - 1;
          LtCur cur;
          cur.mi = mi;
          lt_seek(b, M.vec_begin + lv, cur);
          LtRaw r;
          lt_issue(M, cur, r);
          uint32_t u[8], raw[8];
          LT_DISPATCH(M, (lt_score_bits<DT, AL, true>(M, cur.row, cur.col, r, u, raw)));
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (u[e] != 0xffffffffu) {
              const uint32_t bin = min(score_key(__uint_as_float(u[e])) >> 16, 0x7bffu);
              atomicAdd(&sh_dyn[bin >> 1], 1u << ((bin & 1u) * 16));
            }
          }
        }
        __syncthreads();
        unsigned wmin = 0xffffffffu, wmax = 0u;  // occupied words of this CTA's histogram
        for (int i = tid; i < kLtSampWords; i += kLtThreads) {
          const unsigned w = sh_dyn[i];
          if (w) {
            wmin = min(wmin, (unsigned)i);
            wmax = max(wmax, (unsigned)i + 1u);
          }
          if (w & 0xffffu) atomicAdd(M.coarse + 2 * i, w & 0xffffu);
          if (w >> 16) atomicAdd(M.coarse + 2 * i + 1, w >> 16);
        }
        wmin = __reduce_min_sync(0xffffffffu, wmin);
        wmax = __reduce_max_sync(0xffffffffu, wmax);
        if (lane == 0 && wmax) {  // both encoded so that a zeroed workspace means "empty": max of (~min) and of (max)
          atomicMax(&sh_range[0], 0xffffffffu - 2u * wmin);
          atomicMax(&sh_range[1], 2u * wmax);
        }
        __syncthreads();
        if (tid == 0) {
          if (sh_range[1]) {
            atomicMax(M.bracket + 4, sh_range[0]);
            atomicMax(M.bracket + 5, sh_range[1]);
          }
          sh_range[0] = sh_range[1] = 0u;
        }
        __syncthreads();
      }
      lt_stamp(b, 13);
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const unsigned prev = atomicAdd(M.bracket + 3, 1u);
        sh_last = (prev == (unsigned)nslices - 1);
        if (sh_last) M.bracket[3] = 0u;
      }
      __syncthreads();
      if (sh_last) {
        __threadfence();
        lt_bracket(M, sh_scan);  // ================= P2 =================
      }
    }
  }
}

__global__ void __launch_bounds__(kLtThreads, 2) layer_thresh_batched_kernel(const __grid_constant__ LtBatch b) {
  // dynamic shared memory: P1 uses it as the packed sample histogram; afterwards [n][kLtBins] digit histograms followed
  // by the bracket-vector list
  extern __shared__ unsigned sh_dyn[];
  unsigned* sh_hist = sh_dyn;
  uint32_t* sh_list = sh_dyn + b.n * kLtBins;
  __shared__ unsigned sh_cnt[kLtMaxMat][2];  // per-CTA counts fit 32 bits (a CTA sees < 2^32 elements)
  __shared__ unsigned long long sh_scan[36];
  __shared__ unsigned long long sh_q[kLtThreads / 32 / 4][4];
  __shared__ LtSel sel[kLtMaxMat];
  __shared__ unsigned sh_list_n;
  __shared__ unsigned sh_range[2];
  __shared__ int sh_overflow, sh_last;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t gthreads = (int64_t)gridDim.x * kLtThreads;
  const int64_t gtid = (int64_t)blockIdx.x * kLtThreads + tid;
  const bool list_ok = b.total_vec < (1ll << 32);  // the list stores 32-bit flat vector indices
  if (b.after_split && __ldcg(b.fallback) == 0u) return;  // the split path handled the block
  unsigned bar_gen = *reinterpret_cast<volatile unsigned*>(b.bar);  // read before this CTA's first arrival: nobody can have advanced it
  lt_stamp(b, 0);

  lt_phase_sample(b, sh_dyn, sh_scan, sh_range, &sh_last);
  lt_stamp(b, 14);
  lt_grid_barrier(b.bar, bar_gen);
  lt_stamp(b, 1);
  lt_stamp(b, 2);

  // ================= P3 (+ retry): #(key < lo), bracket vectors -> list -> top-digit histogram ====================
  for (int mi = tid; mi < b.n; mi += kLtThreads) sel[mi].active = 1;
  __syncthreads();
  int rounds = 0;
  for (;; ++rounds) {
    for (int mi = tid; mi < b.n; mi += kLtThreads) {
      const LtMat& M = b.m[mi];
      if (sel[mi].active) {
        const uint32_t lo = __ldcg(M.bracket), hi = __ldcg(M.bracket + 1);
        const uint32_t lo32 = lo << 16;
        const uint32_t range_hi = (hi >= kLtTop ? 0x80000000u : (hi << 16)) - lo32 - 1u;
        const int bits = 32 - __clz(range_hi | 1u);
        sel[mi].lo = lo;
        sel[mi].hi = hi;
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
    if (tid == 0) {
      sh_list_n = 0u;
      sh_overflow = 0;
    }
    __syncthreads();
    {
      LtCur cur, nxt;
      LtRaw ra, rb;
      cur.mi = 0;
      int cmi = -1;  // matrix the running counter and the pivots belong to
      int pc16 = 0;
      __half2 pl = lt_h2(0u), ph = lt_h2(0u);
      int64_t v = gtid;
      if (v < b.total_vec) {
        lt_seek(b, v, cur);
        if (sel[cur.mi].active) lt_issue(b.m[cur.mi], cur, ra);
      }
      for (int64_t base = gtid - lane; base < b.total_vec; base += gthreads, v += gthreads) {  // warp-uniform trip count
        bool cand = false;
        uint32_t u[8];
        if (v < b.total_vec) {
          // the next vector's load goes out before this one is processed: two 16-byte requests in flight per thread
          nxt = cur;
          lt_next(b, v + gthreads, nxt);
          if (v + gthreads < b.total_vec && sel[nxt.mi].active) lt_issue(b.m[nxt.mi], nxt, rb);
          if (cur.mi != cmi) {
            if (pc16) atomicAdd(&sh_cnt[cmi][0], (unsigned)pc16 >> 4);
            pc16 = 0;
            cmi = cur.mi;
            pl = lt_h2(lt_dup(sel[cmi].lo));
            ph = lt_h2(lt_dup(sel[cmi].hi));
          }
          if (sel[cmi].active) {
            const LtMat& M = b.m[cmi];
            uint32_t x = 0;
            LT_DISPATCH(M, (x = lt_count_vec<DT, AL>(M, cur.row, cur.col, ra, pl, ph, pc16, u)));
            cand = x != 0;
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (bal) {
          unsigned pos = 0;
          if (lane == 0) pos = atomicAdd(&sh_list_n, (unsigned)__popc(bal));
          pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(bal & ((1u << lane) - 1u));
          if (cand) {
            if (list_ok && pos < (unsigned)kLtListCap) {
              sh_list[pos] = (uint32_t)v;
            } else {  // list full: histogram this vector on the spot; later digits re-scan this CTA's share of W
              sh_overflow = 1;
              unsigned nb = 0;
              lt_hist_vec(u, sel[cmi], 0, sh_hist + cmi * kLtBins, nb);
              if (nb) atomicAdd(&sh_cnt[cmi][1], nb);
            }
          }
        }
        cur = nxt;
        ra = rb;
      }
      if (pc16) atomicAdd(&sh_cnt[cmi][0], (unsigned)pc16 >> 4);
    }
    __syncthreads();
    lt_stamp(b, 6);
    {
      // walk the list: every lane has a bracket vector (re-read: L1/L2 hit), exact keys, top digit
      const int n_list = list_ok ? (int)min(sh_list_n, (unsigned)kLtListCap) : 0;
      LtCur cur;
      for (int i = tid; i < n_list; i += kLtThreads) {
        cur.mi = 0;
        lt_seek(b, (int64_t)sh_list[i], cur);
        const LtMat& M = b.m[cur.mi];
        LtRaw r;
        lt_issue(M, cur, r);
        uint32_t u[8], raw[8];
        LT_DISPATCH(M, (lt_score_bits<DT, AL, false>(M, cur.row, cur.col, r, u, raw)));
        unsigned nb = 0;
        lt_hist_vec(u, sel[cur.mi], 0, sh_hist + cur.mi * kLtBins, nb);
        if (nb) atomicAdd(&sh_cnt[cur.mi][1], nb);
      }
    }
    __syncthreads();
    lt_stamp(b, 7);
    for (int i = tid; i < b.n * kLtBins; i += kLtThreads) {
      const unsigned c = sh_hist[i];
      const int mi = i / kLtBins;
      if (c) atomicAdd(b.m[mi].hist + (i - mi * kLtBins), c);
    }
    if (tid < b.n * 2) {
      const unsigned long long c = sh_cnt[tid >> 1][tid & 1];
      if (c) atomicAdd(b.m[tid >> 1].cnt + (tid & 1), c);  // 64-bit global atomic: native
    }
    __syncthreads();
    lt_stamp(b, 8);
    lt_grid_barrier(b.bar, bar_gen);
    // every CTA checks the brackets (same global values everywhere)
    int any_retry = 0;
    for (int mi = 0; mi < b.n; ++mi) {
      const LtMat& M = b.m[mi];
      const unsigned long long c_lo = __ldcg(M.cnt), c_band = __ldcg(M.cnt + 1);
      const unsigned long long kth = (unsigned long long)M.kth;
      const bool was_active = sel[mi].active != 0;
      const bool ok = !was_active || (kth >= c_lo && kth < c_lo + c_band);
      if (!ok) any_retry = 1;
      __syncthreads();
      if (tid == 0 && was_active) {
        sel[mi].active = ok ? 0 : 1;
        if (ok) sel[mi].rem = kth - c_lo;
      }
    }
    __syncthreads();
    if (!any_retry) break;
    // rare: the k-th score is outside the sampled bracket.  Move to the side that holds it and count again.
    lt_grid_barrier(b.bar, bar_gen);  // everybody has read cnt / bracket
    if (blockIdx.x == 0) {
      for (int mi = 0; mi < b.n; ++mi) {
        if (!sel[mi].active) continue;
        const LtMat& M = b.m[mi];
        if (tid == 0) {
          const uint32_t lo = M.bracket[0], hi = M.bracket[1];
          if ((unsigned long long)M.kth < M.cnt[0]) { M.bracket[0] = 0u; M.bracket[1] = lo; }
          else { M.bracket[0] = hi; M.bracket[1] = kLtTop; }
          M.cnt[0] = 0ull;
          M.cnt[1] = 0ull;
        }
        for (int i = tid; i < kLtBins; i += kLtThreads) M.hist[i] = 0u;
      }
    }
    lt_grid_barrier(b.bar, bar_gen);
  }

  lt_stamp(b, 3);
  // ================= P4: resolve the digits (level 0 histogram is already in global memory) =======================
  // NB a retry round only re-lists the matrices that were still active, so after a retry the list no longer covers
  // the others: their later digits come from the re-scan path as well.
  int max_nd = 1;
  for (int mi = 0; mi < b.n; ++mi) max_nd = max(max_nd, sel[mi].nd);
  for (int level = 0; level < max_nd; ++level) {
    lt_find_bins(b, sel, level, sh_q);  // every CTA computes the same bins
    lt_stamp(b, 9 + 3 * (level > 0 ? 1 : 0));
    if (level + 1 >= max_nd) break;
    // histogram of the next digit for the elements matching the prefix
    for (int i = tid; i < b.n * kLtBins; i += kLtThreads) sh_hist[i] = 0;
    __syncthreads();
    if (sh_overflow || !list_ok || rounds > 0) {
      LtCur cur;
      cur.mi = 0;
      int64_t v = gtid;
      if (v < b.total_vec) lt_seek(b, v, cur);
      for (; v < b.total_vec; v += gthreads) {
        if (sel[cur.mi].nd > level + 1) {
          const LtMat& M = b.m[cur.mi];
          LtRaw r;
          lt_issue(M, cur, r);
          uint32_t u[8], raw[8];
          LT_DISPATCH(M, (lt_score_bits<DT, AL, false>(M, cur.row, cur.col, r, u, raw)));
          unsigned nb = 0;
          lt_hist_vec(u, sel[cur.mi], level + 1, sh_hist + cur.mi * kLtBins, nb);
        }
        lt_next(b, v + gthreads, cur);
      }
    } else {
      const int n_list = (int)min(sh_list_n, (unsigned)kLtListCap);
      LtCur cur;
      for (int i = tid; i < n_list; i += kLtThreads) {
        cur.mi = 0;
        lt_seek(b, (int64_t)sh_list[i], cur);
        if (sel[cur.mi].nd <= level + 1) continue;
        const LtMat& M = b.m[cur.mi];
        LtRaw r;
        lt_issue(M, cur, r);
        uint32_t u[8], raw[8];
        LT_DISPATCH(M, (lt_score_bits<DT, AL, false>(M, cur.row, cur.col, r, u, raw)));
        unsigned nb = 0;
        lt_hist_vec(u, sel[cur.mi], level + 1, sh_hist + cur.mi * kLtBins, nb);
      }
    }
    __syncthreads();
    lt_stamp(b, 10);
    for (int i = tid; i < b.n * kLtBins; i += kLtThreads) {
      const unsigned c = sh_hist[i];
      const int mi = i / kLtBins;
      if (c) atomicAdd(b.m[mi].hist + (level + 1) * kLtBins + (i - mi * kLtBins), c);
    }
    __syncthreads();
    lt_stamp(b, 11);
    lt_grid_barrier(b.bar, bar_gen);
  }

  lt_stamp(b, 4);
  // ================= P5: apply  score <= thres  in place ===========================================================
  for (int mi = tid; mi < b.n; mi += kLtThreads) sh_cnt[mi][0] = 0u;
  __syncthreads();
  {
    LtCur cur, nxt;
    LtRaw ra, rb;
    cur.mi = 0;
    int cmi = -1;
    int zeros = 0;
    uint32_t tcmp = 0;
    int64_t v = gtid;
    if (v < b.total_vec) {
      lt_seek(b, v, cur);
      lt_issue(b.m[cur.mi], cur, ra);
    }
    for (; v < b.total_vec; v += gthreads) {
      nxt = cur;
      lt_next(b, v + gthreads, nxt);
      if (v + gthreads < b.total_vec) lt_issue(b.m[nxt.mi], nxt, rb);
      if (cur.mi != cmi) {
        if (zeros) atomicAdd(&sh_cnt[cmi][0], (unsigned)zeros);
        zeros = 0;
        cmi = cur.mi;
        const uint32_t tkey = sel[cmi].lo32 + sel[cmi].prefix;
        // float semantics of `W_metric <= thres`: NaN scores are never pruned, a NaN threshold prunes nothing
        tcmp = tkey > 0x7f800000u ? 0u : tkey + 1u;  // prune iff key < tcmp
      }
      const LtMat& M = b.m[cmi];
      LT_DISPATCH(M, (lt_apply_vec<DT, AL>(M, cur.row, cur.col, ra, tcmp, zeros)));
      cur = nxt;
      ra = rb;
    }
    if (zeros) atomicAdd(&sh_cnt[cmi][0], (unsigned)zeros);
    __syncthreads();
    if (tid < b.n && sh_cnt[tid][0] && b.m[tid].n_zero != nullptr) atomicAdd(b.m[tid].n_zero, (unsigned long long)sh_cnt[tid][0]);
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

static size_t lc_mat_bytes() {
  // cutoff path, per matrix: bracket histogram, one line of counters / select state / results, the bin list
  return align_up((size_t)kLcCoarse * kLcCoarseStride * sizeof(unsigned), 256) + 256 + align_up((size_t)kLcListCap * sizeof(uint2), 256);
}

static size_t lt_split_bytes() { return (size_t)kLtMaxMat * lc_mat_bytes(); }

size_t layer_thresh_batched_workspace_bytes(const ecf_layer_desc* descs, int n) {
  if (descs == nullptr || n < 1 || n > kLtMaxMat) return 0;
  size_t total = kLtHeaderBytes + lt_split_bytes();
  for (int i = 0; i < n; ++i) {
    if (descs[i].C <= 0) return 0;
    total += lt_mat_fixed_bytes() + align_up((size_t)descs[i].C * sizeof(float), 256);
  }
  return total;
}

size_t layer_thresh_workspace_bytes(int64_t R, int64_t C) {
  (void)R;
  return kLtHeaderBytes + lt_split_bytes() + lt_mat_fixed_bytes() + align_up((size_t)(C > 0 ? C : 1) * sizeof(float), 256);
}

}  // namespace ecf

extern "C" size_t ecf_layer_thresh_flag_offset(void) { return 128 + (1 + ecf::kLtBarMaxGroups) * 128 + 4; }

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
    ECF_REQUIRE(d.R < (1ll << 31) && d.C < (1ll << 31), ECF_ERR_INVALID, "layer_thresh: R or C beyond 2^31 (matrix %d)", i);
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
  b.bar = reinterpret_cast<unsigned*>(p + 128);  // 128 bytes of stamps, then (1 + kLtBarMaxGroups) 128-byte barrier lines
  static_assert(128 + (1 + kLtBarMaxGroups) * 128 + 256 + kLtMaxMat * 16 <= kLtHeaderBytes, "workspace header too small");
  // header words of the cutoff path: [0] fallback flag (live during a launch, cleared by its last kernel),
  // [1] what it was in the last launch (test / profiling aid), [2] raised inside K1, [16] K1 ticket, [17] K4 ticket
  unsigned* lc_hdr = reinterpret_cast<unsigned*>(p + 128 + (1 + kLtBarMaxGroups) * 128);
  b.fallback = lc_hdr;
  b.after_split = 0;
  p += kLtHeaderBytes;
  char* lc_ws = p;
  p += (size_t)kLtMaxMat * lc_mat_bytes();
  LcBatch cb;
  cb.n = 0;
  cb.total_ctas = 0;
  cb.fallback = lc_hdr;
  cb.ticket = lc_hdr + 16;
  cb.nsigma = 4.0f;
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
  int64_t samples = 0;
  for (int i = 0; i < n; ++i) {
    LtMat& M = b.m[i];
    M.sample_begin = samples;
    samples += (M.nvec + M.sample_stride - 1) / M.sample_stride;
  }
  b.total_samples = samples;

  size_t smem = (size_t)n * kLtBins * sizeof(unsigned) + (size_t)kLtListCap * sizeof(uint32_t);
  if (smem < (size_t)kLtSampWords * sizeof(unsigned)) smem = (size_t)kLtSampWords * sizeof(unsigned);
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
  const int64_t gthreads = want * kLtThreads;
  for (int i = 0; i < n; ++i) {
    b.m[i].step_rows = gthreads / b.m[i].nvpr;
    b.m[i].step_cols = gthreads % b.m[i].nvpr;
  }
  // ---- cutoff path when every matrix qualifies (16-bit, aligned); everything else goes to the cooperative kernel
  bool fast = vec < (1ll << 31);
  for (int i = 0; i < n; ++i) fast = fast && b.m[i].aligned && b.m[i].dtype != ECF_F32 && b.m[i].nvec < (1ll << 29) && b.m[i].nvec >= 1;
  // ECF_LT_CUT=0 switches the cutoff path off (A/B against the cooperative kernel alone)
  static const bool cut_off = [] { const char* v = getenv("ECF_LT_CUT"); return v != nullptr && v[0] == '0'; }();
  // half-width of the sampled bracket in standard deviations of the sample rank (read per call: tests force the exact
  // fallback with ECF_LT_NSIGMA=0)
  const char* nsig_env = getenv("ECF_LT_NSIGMA");
  const float cut_nsigma = nsig_env != nullptr ? (float)atof(nsig_env) : 4.0f;
  bool cut = fast && !cut_off;
  for (int i = 0; i < n; ++i)  // 32-bit element indices in the bin list; one dtype per launch (the kernels are specialised)
    cut = cut && (uint64_t)descs[i].R * (uint64_t)descs[i].C < (1ull << 32) && descs[i].dtype == descs[0].dtype;
  if (cut) {
    // ---- cutoff path (layer_cut.cuh): K0 sample, K1 count, K3 apply, K4 fix-up / exact cluster select (fallback)
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cb.n = n;
    cb.nsigma = cut_nsigma;
    uint64_t units = 0;
    for (int i = 0; i < n; ++i) {
      const ecf_layer_desc& d = descs[i];
      LcMat& M = cb.m[i];
      M.W = d.W; M.s = d.scaler_row; M.thres_out = d.thres_out; M.mask = d.mask_bits; M.n_zero = d.n_zero;
      M.ld = d.ld; M.mask_ld = d.mask_ld; M.kth = d.kth_index;
      M.frac = (double)d.kth_index / ((double)d.R * (double)d.C);
      M.R = (uint32_t)d.R; M.C = (uint32_t)d.C; M.nvpr = (uint32_t)(d.C / 8); M.dtype = d.dtype;
      M.slabs = (M.nvpr + kLcSlabVecs - 1) / kLcSlabVecs;
      units += (uint64_t)M.slabs * M.R;
      char* q = lc_ws + (size_t)i * lc_mat_bytes();
      M.hist = reinterpret_cast<unsigned*>(q); q += align_up((size_t)kLcCoarse * kLcCoarseStride * sizeof(unsigned), 256);
      M.cnt = reinterpret_cast<unsigned long long*>(q);
      M.sel = reinterpret_cast<int32_t*>(q + 32);
      M.res = reinterpret_cast<int32_t*>(q + 64);
      M.list_n = reinterpret_cast<unsigned*>(q + 128);
      q += 256;
      M.list = reinterpret_cast<uint2*>(q);
    }
    const uint64_t g_target = (uint64_t)sm_count() * kLcCtasPerSm;
    uint64_t rpi = (units + g_target - 1) / g_target;
    if (rpi < (uint64_t)kLcRowsPerIter) rpi = kLcRowsPerIter;
    for (;;) {  // the smallest row range per CTA that keeps the grid to one resident wave
      uint64_t ctas = 0;
      for (int i = 0; i < n; ++i) ctas += (uint64_t)cb.m[i].slabs * ((cb.m[i].R + rpi - 1) / rpi);
      if (ctas <= g_target) break;
      ++rpi;
    }
    unsigned ctas = 0;
    for (int i = 0; i < n; ++i) {
      LcMat& M = cb.m[i];
      M.rows_per_item = (uint32_t)(rpi < M.R ? rpi : M.R);
      M.cta_begin = ctas;
      M.cta_count = M.slabs * (unsigned)((M.R + rpi - 1) / rpi);
      ctas += M.cta_count;
    }
    cb.total_ctas = ctas;
    static bool lc_attr_done = false;
    if (!lc_attr_done) {
      ECF_CUDA_OK(cudaFuncSetAttribute(lc_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLcSampleSmem));
      ECF_CUDA_OK(cudaFuncSetAttribute(lc_count_kernel<ECF_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLcRingCountBytes));
      ECF_CUDA_OK(cudaFuncSetAttribute(lc_count_kernel<ECF_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLcRingCountBytes));
      ECF_CUDA_OK(cudaFuncSetAttribute(lc_apply_kernel<ECF_F16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLcRingApplyBytes));
      ECF_CUDA_OK(cudaFuncSetAttribute(lc_apply_kernel<ECF_F16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLcRingApplyBytes));
      ECF_CUDA_OK(cudaFuncSetAttribute(lc_apply_kernel<ECF_BF16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLcRingApplyBytes));
      ECF_CUDA_OK(cudaFuncSetAttribute(lc_apply_kernel<ECF_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLcRingApplyBytes));
      lc_attr_done = true;
    }
    // ECF_LT_STOP=k (profiling aid): issue only the first k kernels of the chain -- the result is then incomplete
    const char* stop_env = getenv("ECF_LT_STOP");
    const int stop = stop_env != nullptr ? atoi(stop_env) : 4;
    lc_sample_kernel<<<(unsigned)(n * kLcSampleCluster), kLcSampleCtaThreads, kLcSampleSmem, st>>>(cb);
    if (stop < 2) return ECF_OK;
    bool extras = false;
    for (int i = 0; i < n; ++i) extras = extras || descs[i].mask_bits != nullptr || descs[i].n_zero != nullptr;
    // ECF_LT_PDL=1 launches K1, K3 and K4 with programmatic stream serialisation (PDL, see layer_cut.cuh).  Measured under
    // CUDA-graph replay (profiles/r2/lt_cut_*_pdl{0,1}.log): no gain -- the graph's kernel-to-kernel gaps are ~1 us already
    // and the chain is bound by the kernels themselves -- so plain stream order is the default (the griddepcontrol
    // instructions are no-ops then).
    static const bool pdl = [] { const char* v = getenv("ECF_LT_PDL"); return v != nullptr && v[0] == '1'; }();
    auto launch = [&](auto kern, unsigned grid, unsigned block, size_t smem_bytes) -> cudaError_t {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(block);
      cfg.dynamicSmemBytes = smem_bytes;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl ? 1 : 0;
      return cudaLaunchKernelEx(&cfg, kern, cb);
    };
    const bool f16 = descs[0].dtype == ECF_F16;
    ECF_CUDA_OK(f16 ? launch(lc_count_kernel<ECF_F16>, ctas, kLcThreads, kLcRingCountBytes)
                    : launch(lc_count_kernel<ECF_BF16>, ctas, kLcThreads, kLcRingCountBytes));
    if (stop < 3) return ECF_OK;
    if (f16) ECF_CUDA_OK(extras ? launch(lc_apply_kernel<ECF_F16, true>, ctas, kLcThreads, kLcRingApplyBytes)
                                : launch(lc_apply_kernel<ECF_F16, false>, ctas, kLcThreads, kLcRingApplyBytes));
    else ECF_CUDA_OK(extras ? launch(lc_apply_kernel<ECF_BF16, true>, ctas, kLcThreads, kLcRingApplyBytes)
                            : launch(lc_apply_kernel<ECF_BF16, false>, ctas, kLcThreads, kLcRingApplyBytes));
    if (stop < 4) return ECF_OK;
    ECF_CUDA_OK(cudaGetLastError());
    // K4: fix-up, or the exact cluster select when the flag is up (one cluster per matrix)
    ECF_CUDA_OK(launch(lc_finish_kernel, (unsigned)(n * kLcCluster), kLcFinishThreads, 0));
    ECF_CUDA_OK(cudaGetLastError());
    return ECF_OK;
  }
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
