// A3+A5+A7 -- per-LAYER threshold select, CUTOFF path (16-bit, aligned matrices: every BLIP-2 / LLaMA Linear).
//
// Replaces  thres = torch.sort(W_metric.flatten())[0][int(numel * s)];  W[W_metric <= thres] = 0
// (LAVIS/lavis/compression/pruners/wanda_pruner.py:541,553-558; UPop wanda_pruner.py:502,512-517;
//  LLaMA/image_classifiers/prune_utils.py:28-31).
//
// Idea.  All the weights of column c share q_c = sqrtf(scaler_row[c]) and an IEEE fp32 multiply by a non-negative
// constant is monotone, so for any fp32 score value t
//        fl32(fp32(|w|) * q_c) <= t      <=>      |w| <= cut_c(t)
// where cut_c(t) is the largest 16-bit magnitude pattern whose exact score is <= t (found per column with the very
// same fl32 multiply, so the equivalence is exact, not approximate).  Counting / bracketing / applying a threshold
// over the whole matrix then needs no score at all: one AND (|w|) and one packed 16-bit compare per PAIR of weights
// instead of two conversions, two multiplies, a pack and the compares -- the streaming passes drop from ~100 to ~25
// instructions per 8-element vector and become what they should be: memory bound.
//
// Four launches per block (all the Linears of the block together), in stream order:
//   K0 sample   a cluster of 8 CTAs per matrix: 16 384 stratified samples -> exact scores -> a score bracket (lo, hi]
//               that holds the k-th score with ~4 sigma on either side (a sorted 256-key sub-sample seeds it, 2 048-bin
//               shared histograms of the samples inside the seed bracket, merged through DSMEM, refine it)
//   K1 count    the only HBM read of W: #(score <= lo) by per-column cutoffs; vectors holding a bracket element (~20 %)
//               are parked in shared memory, their bracket elements (~3 % of the matrix) get their exact key and go into a
//               256-bin histogram of (key - lo) >> shift (shared, then one RED per bin and CTA, every global bin on its own
//               L2 slice).  The last CTA (ticket) checks the bracket and finds the bin B that holds the k-th score.
//   K3 apply    (L2 read, HBM write) zero everything below bin B by cutoffs; the elements inside B (~1 500 per matrix) are
//               appended to a list with their exact keys
//   K4 finish   a cluster per matrix: CTA 0 finds the exact k-th key among the listed elements (11 bits per level),
//               writes the threshold and zeroes the listed elements at or below it.  Anything unusual (k-th score outside
//               the bracket, non-finite sqrt(scaler_row), a bin beyond the list capacity) is known before K3 writes
//               anything and raises a flag: K3 then does nothing and K4's clusters run an exact three-digit radix select
//               over the whole matrix instead (shared histograms merged through DSMEM) -- exact either way.
// Work split of K1 / K3: a CTA owns (matrix, slab of 16 vectors = 128 columns, row range); a thread keeps its 8
// columns for the whole range, so its cutoffs live in registers and the address stream is one add per row.  CTAs are
// numbered slab-fastest so that the CTAs running together read whole rows.  The grid is one resident wave.
// Bound: HBM.  Algorithmic bytes per matrix: 2*R*C*sizeof(w) + 4*C.
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"

namespace ecf {

constexpr int kLcThreads = 256;
constexpr int kLcCtasPerSm = 3;
constexpr int kLcSlabVecs = 16;                  // vectors per slab row (128 columns, 256 bytes)
constexpr int kLcSlabCols = kLcSlabVecs * 8;
constexpr int kLcRowsPerIter = kLcThreads / kLcSlabVecs;  // 16 rows per CTA iteration
constexpr int kLcBins = 2048;                    // bins of the shared-memory histograms (sample refinement, fix-up, exact select)
constexpr int kLcCoarse = 256;                   // bins of the global bracket histogram
constexpr int kLcCoarseStride = 64;              // words between two of them: every bin on its own 256-byte line, i.e. its own
                                                 // L2 slice -- atomics on neighbouring words serialise in ONE slice (measured:
                                                 // 780 k REDs into 6 KB took 35 us)
constexpr int kLcListCap = 8192;                 // elements of the k-th score's bin (expected ~1 500)
constexpr uint32_t kLcMaxFinite = 0x7f7fffffu;   // brackets never reach the inf / NaN keys (those blocks fall back)

struct LcMat {
  void* W;
  const float* s;
  float* thres_out;
  uint8_t* mask;
  unsigned long long* n_zero;
  int64_t ld, mask_ld;
  int64_t kth;
  double frac;              // kth / numel
  uint32_t R, C, nvpr;
  uint32_t slabs, rows_per_item, cta_begin, cta_count;
  int dtype;
  // workspace
  unsigned* hist;           // [kLcCoarse] bracket histogram, bin i at word i * kLcCoarseStride
  unsigned long long* cnt;  // [0] #(score <= lo), [1] #(lo < score <= hi)
  int32_t* sel;             // [0] lo (exclusive, -1 = nothing below), [1] hi (inclusive), [2] shift
  int32_t* res;             // [0] tA: zero everything with key <= tA (-1: nothing), [1] tB: keys in (tA, tB] go to the list,
                            // [2] rank left inside the bin, [3] elements in the bin, [4] first key of the bin, [5] log2 of its width
  unsigned* list_n;
  uint2* list;              // [kLcListCap] (element index, key)
};

struct LcBatch {
  LcMat m[ECF_LAYER_MAX_BATCH];
  int n;
  unsigned total_ctas;
  unsigned* fallback;       // [0] != 0: K4 redoes the block exactly (written by K0 and by the LAST CTA of K1 only, so that it is
                            // stable while a kernel reads it); [1] the flag of the previous launch (test aid); [2] raised by any
                            // CTA of K1 (non-finite norm), folded into [0] by K1's last CTA
  unsigned* ticket;         // [0] K1 arrivals, [1] K4 arrivals
  float nsigma;
};

// ------------------------------------------------------------------------------------------------ small helpers
// explicit shared-space accesses: the hot loops must not fall back to generic addressing
__device__ __forceinline__ uint32_t lc_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lc_sts_v4(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void lc_sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lc_lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lc_lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lc_lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lc_lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void lc_red_shared(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lc_atom_shared(uint32_t a, uint32_t v) {
  uint32_t r;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(v) : "memory");
  return r;
}

// programmatic dependent launch (PDL): the four kernels of a block are launched with programmatic stream serialisation, so
// the next kernel's CTAs become resident -- and run their prologue: ring loads of W, the q table -- while the previous
// kernel is still finishing; lc_pdl_wait() blocks until the previous kernel has completed and its writes are visible.
__device__ __forceinline__ void lc_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void lc_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// fp32 value of a 16-bit magnitude pattern
template <int DT>
__device__ __forceinline__ float lc_mag(uint32_t m) {
  if constexpr (DT == ECF_BF16) return __uint_as_float(m << 16);
  else return __half2float(__ushort_as_half((unsigned short)m));
}
template <int DT>
__device__ __forceinline__ uint32_t lc_max_finite() { return DT == ECF_BF16 ? 0x7f7fu : 0x7bffu; }
template <int DT>
__device__ __forceinline__ uint32_t lc_none() { return DT == ECF_BF16 ? 0xbf80u : 0xbc00u; }  // -1.0: |w| <= -1 never holds

// packed compare |w| <= cut on two 16-bit floats at once -> 0xffff per true half (NaN weights compare false, like
// `NaN <= thres` in the reference)
template <int DT>
__device__ __forceinline__ uint32_t lc_le2(uint32_t a, uint32_t cut) {
  if constexpr (DT == ECF_BF16) {
    return __hle2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&cut));
  } else {
    return __hle2_mask(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&cut));
  }
}
template <int DT>
__device__ __forceinline__ uint32_t lc_eq0(uint32_t a) {  // 0xffff per half holding +-0
  const uint32_t z = 0u;
  if constexpr (DT == ECF_BF16) {
    return __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&z));
  } else {
    return __heq2_mask(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&z));
  }
}
// one bit per element of a vector from the four packed compare masks
__device__ __forceinline__ uint32_t lc_bits8(const uint32_t (&x)[4]) {
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) bits |= ((x[j] & 1u) | ((x[j] >> 15) & 2u)) << (2 * j);
  return bits;
}

// cut_c(T): the largest finite magnitude pattern m with score_key(fl32(mag(m) * q)) <= T, or "none".  T is a key
// (int32, -1 = nothing qualifies, <= kLcMaxFinite), q is finite and >= 0.
template <int DT>
__device__ __forceinline__ uint32_t lc_cutoff(int32_t T, float q) {
  if (T < 0) return lc_none<DT>();
  const uint32_t top = lc_max_finite<DT>();
  if (q == 0.f) return top;  // every finite weight scores 0 <= T
  const float g = __uint_as_float((uint32_t)T) / q;  // IEEE division: the guess is within a pattern of the answer
  uint32_t m;
  if (!(g < lc_mag<DT>(top))) {
    m = top;
  } else if constexpr (DT == ECF_BF16) {
    m = __float_as_uint(g) >> 16;  // truncation = round toward zero
  } else {
    m = (uint32_t)__half_as_ushort(__float2half_rz(g));
  }
  // exact fix-up with the kernel's own multiply (the guess is off by a pattern or two at most; monotone => terminates)
  while (m < top && score_key(__fmul_rn(lc_mag<DT>(m + 1u), q)) <= (uint32_t)T) ++m;
  while (m > 0u && score_key(__fmul_rn(lc_mag<DT>(m), q)) > (uint32_t)T) --m;
  return m;  // m == 0: the score of a zero weight is 0 <= T
}

__device__ __forceinline__ int lc_find_mat(const LcBatch& b, unsigned cta) {
  int mi = 0;
  while (mi + 1 < b.n && cta >= b.m[mi + 1].cta_begin) ++mi;
  return mi;
}

// ------------------------------------------------------------------------------------------------ K0: sample
// A cluster of kLcSampleCluster CTAs per matrix.  Scattered loads cost the LSU ~2 cycles per lane whatever their width and
// shared atomics as much, and one SM gathering and binning 16 384 samples took 42 k cycles (measured), so (1) a sample
// location is 4 consecutive weights of a row (8 bytes + a 16-byte norm load: 4 096 locations per matrix), (2) the locations
// and the histogram work are spread over 8 SMs, (3) every CTA sorts the same 256-key sub-sample (no broadcast step) and
// (4) the per-CTA histograms are pushed into CTA 0 through DSMEM: two cluster barriers in all.
constexpr int kLcSub = 256;            // sorted sub-sample that seeds the bracket (every 64th sample)
constexpr int kLcSampleCluster = 8;
constexpr int kLcSampleCtaThreads = 256;
constexpr int kLcSampleKeys = 16384;   // per matrix: 8 CTAs x 256 threads x 2 locations x 4 weights
constexpr int kLcLocsPerThread = 2;
constexpr size_t kLcSampleSmem = (size_t)kLcSampleCluster * kLcBins * sizeof(unsigned);  // gather area (used in CTA 0)

__global__ void __cluster_dims__(kLcSampleCluster, 1, 1) __launch_bounds__(kLcSampleCtaThreads, 1)
    lc_sample_kernel(const __grid_constant__ LcBatch b) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned sh_gather[];  // [kLcSampleCluster][kLcBins]: every CTA's histogram, in CTA 0
  __shared__ __align__(16) uint32_t raw_sub[kLcSub];
  __shared__ uint32_t sub[kLcSub];
  __shared__ __align__(16) unsigned h2[kLcBins];
  __shared__ unsigned wsum[kLcSampleCtaThreads / 32];
  __shared__ unsigned s_below[kLcSampleCluster];
  __shared__ unsigned s_my_below;
  __shared__ int s_jlo, s_jhi;
  __shared__ uint32_t s_a, s_b;
  __shared__ int s_sh, s_fail, s_lo_none, s_hi_max;
  __shared__ long long s_rlo, s_rhi;
  cg::cluster_group cluster = cg::this_cluster();
  lc_pdl_trigger();  // K1's CTAs may take the free SMs now: their W loads do not depend on the bracket
  const unsigned rank = cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int mi = (int)blockIdx.x / kLcSampleCluster;
  const LcMat& M = b.m[mi];
  const uint64_t numel = (uint64_t)M.R * M.C;
  const bool exact = numel <= (uint64_t)kLcSampleKeys;  // everything is sampled
  const uint32_t ns = exact ? (uint32_t)numel : (uint32_t)kLcSampleKeys;
  const uint32_t nloc = ns / 4;  // C % 8 == 0, so numel % 4 == 0
  // ---- this thread's locations: row i*R/4096 (evenly stratified), column group frac(i * phi) * C/4 (low discrepancy)
  uint32_t key[kLcLocsPerThread * 4];
  {
    int64_t off[kLcLocsPerThread];
    uint32_t cols[kLcLocsPerThread];
    uint32_t loc[kLcLocsPerThread];
#pragma unroll
    for (int l = 0; l < kLcLocsPerThread; ++l) {
      loc[l] = (rank * kLcSampleCtaThreads + (uint32_t)tid) * kLcLocsPerThread + (uint32_t)l;
      const uint32_t i = min(loc[l], nloc - 1u);
      uint32_t row;
      if (exact) {
        row = (4u * i) / M.C;
        cols[l] = 4u * i - row * M.C;
      } else {
        row = __umulhi(i << 20, M.R);                       // i * R / 2^12
        cols[l] = 4u * __umulhi(i * 2654435769u, M.C >> 2);
      }
      off[l] = (int64_t)row * M.ld + cols[l];
    }
    uint2 wv[kLcLocsPerThread];
    float4 sv[kLcLocsPerThread];
    const uint16_t* W16 = reinterpret_cast<const uint16_t*>(M.W);
#pragma unroll
    for (int l = 0; l < kLcLocsPerThread; ++l) wv[l] = *reinterpret_cast<const uint2*>(W16 + off[l]);
#pragma unroll
    for (int l = 0; l < kLcLocsPerThread; ++l) sv[l] = *reinterpret_cast<const float4*>(M.s + cols[l]);
    for (int i = tid; i < kLcBins; i += kLcSampleCtaThreads) h2[i] = 0u;
    if (tid == 0) {
      s_my_below = 0u;
      s_jlo = -1;
      s_jhi = -1;
    }
#pragma unroll
    for (int l = 0; l < kLcLocsPerThread; ++l) {
      const uint32_t wb[4] = {wv[l].x & 0x7fffu, (wv[l].x >> 16) & 0x7fffu, wv[l].y & 0x7fffu, (wv[l].y >> 16) & 0x7fffu};
      const float sq[4] = {sv[l].x, sv[l].y, sv[l].z, sv[l].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float q = __fadd_rn(sqrtf(sq[e]), 0.f);
        const float w = M.dtype == ECF_BF16 ? lc_mag<ECF_BF16>(wb[e]) : lc_mag<ECF_F16>(wb[e]);
        key[4 * l + e] = loc[l] < nloc ? score_key(__fmul_rn(w, q)) : 0xffffffffu;
      }
    }
    // every 64th key of the sample (the first weight of every 16th location) goes to every CTA's sub-sample
    if ((tid & 7) == 0) {
      const uint32_t si = (rank * kLcSampleCtaThreads + (uint32_t)tid) >> 3;
#pragma unroll
      for (int r = 0; r < kLcSampleCluster; ++r) cluster.map_shared_rank(raw_sub, r)[si] = key[0];
    }
  }
  cluster.sync();  // #1: the sub-sample is complete everywhere
  {  // sort it by counting: rank = #(smaller) + #(equal, earlier); 64 broadcast LDS.128 per thread, no barriers
    const uint32_t me = raw_sub[tid];
    int r = 0;
#pragma unroll 4
    for (int j = 0; j < kLcSub; j += 4) {
      const uint4 o = *reinterpret_cast<const uint4*>(raw_sub + j);
      r += (o.x < me || (o.x == me && j + 0 < tid)) ? 1 : 0;
      r += (o.y < me || (o.y == me && j + 1 < tid)) ? 1 : 0;
      r += (o.z < me || (o.z == me && j + 2 < tid)) ? 1 : 0;
      r += (o.w < me || (o.w == me && j + 3 < tid)) ? 1 : 0;
    }
    sub[r] = me;
  }
  __syncthreads();
  if (tid == 0) {  // identical in every CTA of the cluster
    const int nsub = (int)((ns + 63u) / 64u);
    const double p = M.frac;
    const double sig = sqrt(max(p * (1.0 - p), 0.0));
    // ranks (in the full sample) that bracket the k-th score
    const double c_s = p * (double)ns, d_s = exact ? 0.0 : (double)b.nsigma * sig * sqrt((double)ns) + 4.0;
    const long long r_lo = exact ? (long long)M.kth : (long long)floor(c_s - d_s);
    const long long r_hi = exact ? (long long)M.kth : (long long)ceil(c_s + d_s);
    // the seed bracket must contain both with 4.5 sigma of its own; towards the ends of the distribution it is the end
    const double c_b = p * (double)nsub, d_b = 4.5 * sig * sqrt((double)nsub) + 3.0 + d_s * (double)nsub / (double)ns;
    const int ia = (int)floor(c_b - d_b), ib = (int)ceil(c_b + d_b);
    uint32_t a = (ia < 0 || r_lo < 128) ? 0u : sub[min(ia, nsub - 1)];
    uint32_t bb = (ib >= nsub || r_hi + 128 >= (long long)ns) ? kLcMaxFinite : sub[ib];
    if (bb > kLcMaxFinite) bb = kLcMaxFinite;
    const bool fail = a > bb;
    const uint32_t rng = fail ? 0u : bb - a;
    s_a = a;
    s_b = bb;
    s_sh = max(0, (32 - __clz((int)(rng | 1u))) - 11);  // (key - a) >> sh < 2048
    s_fail = fail ? 1 : 0;
    s_rlo = r_lo;
    s_rhi = r_hi;
    s_lo_none = r_lo < 0 ? 1 : 0;
    s_hi_max = r_hi >= (long long)ns ? 1 : 0;
  }
  __syncthreads();
  const uint32_t a = s_a, bb = s_b;
  const int sh = s_sh;
  {
    unsigned below = 0;
    const uint32_t h2a = lc_saddr(h2);
#pragma unroll
    for (int j = 0; j < kLcLocsPerThread * 4; ++j) {
      if (key[j] == 0xffffffffu) continue;
      if (key[j] < a) ++below;
      else if (key[j] <= bb) lc_red_shared(h2a + 4u * ((key[j] - a) >> sh), 1u);
    }
    below = __reduce_add_sync(0xffffffffu, below);
    if (lane == 0 && below) atomicAdd(&s_my_below, below);
  }
  __syncthreads();
  {  // push this CTA's histogram and count into CTA 0
    unsigned* dst = cluster.map_shared_rank(sh_gather, 0) + rank * kLcBins;
    for (int i = tid; i < kLcBins / 4; i += kLcSampleCtaThreads) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(h2)[i];
    if (tid == 0) cluster.map_shared_rank(s_below, 0)[rank] = s_my_below;
  }
  cluster.sync();  // #2: everything has landed in CTA 0
  if (rank != 0) return;
  // ---- scan the 2 048 merged bins (eight per thread): bins that hold ranks r_lo and r_hi of the sample
  {
    const long long r_lo = s_rlo, r_hi = s_rhi;
    constexpr int PER = kLcBins / kLcSampleCtaThreads;
    static_assert(PER == 8, "eight bins per thread");
    unsigned cs[PER], tot = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) cs[e] = 0u;
#pragma unroll
    for (int r = 0; r < kLcSampleCluster; ++r) {
      const uint4 x = *reinterpret_cast<const uint4*>(sh_gather + r * kLcBins + PER * tid);
      const uint4 y = *reinterpret_cast<const uint4*>(sh_gather + r * kLcBins + PER * tid + 4);
      cs[0] += x.x; cs[1] += x.y; cs[2] += x.z; cs[3] += x.w; cs[4] += y.x; cs[5] += y.y; cs[6] += y.z; cs[7] += y.w;
    }
#pragma unroll
    for (int e = 0; e < PER; ++e) tot += cs[e];
    unsigned inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    long long run = (long long)inc - tot;
    for (int r = 0; r < kLcSampleCluster; ++r) run += s_below[r];
    for (int w = 0; w < wid; ++w) run += wsum[w];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      if (cs[e]) {
        if (r_lo >= run && r_lo < run + (long long)cs[e]) s_jlo = PER * tid + e;
        if (r_hi >= run && r_hi < run + (long long)cs[e]) s_jhi = PER * tid + e;
      }
      run += cs[e];
    }
  }
  __syncthreads();
  // per-launch state of K1 / K3 / K4 (after the cluster barriers: their release fences would wait for these stores)
  M.hist[tid * kLcCoarseStride] = 0u;
  if (tid < 2) M.cnt[tid] = 0ull;
  if (tid == 0) {
    *M.list_n = 0u;
    if (mi == 0) *b.ticket = 0u;
    bool fail = s_fail != 0;
    const bool lo_none = s_lo_none != 0, hi_max = s_hi_max != 0;
    if (!lo_none && s_jlo < 0) fail = true;
    if (!hi_max && s_jhi < 0) fail = true;
    long long lo = lo_none ? -1ll : (long long)a + ((long long)max(s_jlo, 0) << sh) - 1ll;
    long long hi = hi_max ? (long long)kLcMaxFinite : (long long)a + (((long long)max(s_jhi, 0) + 1ll) << sh) - 1ll;
    if (hi > (long long)kLcMaxFinite) hi = (long long)kLcMaxFinite;
    if (hi <= lo) {
      fail = true;
      hi = lo + 1;
    }
    const long long range = hi - lo;  // keys in (lo, hi]
    const int bits = 64 - __clzll((long long)((unsigned long long)(range - 1) | 1ull));
    M.sel[0] = (int32_t)lo;
    M.sel[1] = (int32_t)hi;
    M.sel[2] = max(0, bits - 8);  // (key - lo - 1) >> shift < kLcCoarse
    M.sel[3] = 0;
    // the flag starts at 0: the last CTA of the previous launch's K4 cleared it (a reset here would race with the other clusters)
    if (fail) atomicExch(b.fallback, 1u);
  }
}

// ------------------------------------------------------------------------------------------------ K1 / K3 common
struct LcItem {
  int mi;
  uint32_t slab, row0, rows;   // rows [row0, row0 + rows)
  uint32_t colvec;             // this thread's vector column
  uint32_t rsub;               // this thread's row lane (0..15)
  int nit;                     // iterations of this thread
  bool active;
};

__device__ __forceinline__ void lc_item(const LcBatch& b, LcItem& it) {
  it.mi = lc_find_mat(b, blockIdx.x);
  const LcMat& M = b.m[it.mi];
  const uint32_t local = blockIdx.x - M.cta_begin;
  const uint32_t chunk = local / M.slabs;
  it.slab = local - chunk * M.slabs;
  it.row0 = chunk * M.rows_per_item;
  it.rows = min(M.rows_per_item, M.R - it.row0);
  it.colvec = it.slab * kLcSlabVecs + (threadIdx.x & (kLcSlabVecs - 1));
  it.rsub = threadIdx.x / kLcSlabVecs;
  it.active = it.colvec < M.nvpr;
  it.nit = (it.active && it.rows > it.rsub) ? (int)((it.rows - it.rsub + kLcRowsPerIter - 1) / kLcRowsPerIter) : 0;
}

constexpr uint32_t kLcNaN2 = 0x7fff7fffu;  // a pair of NaN patterns: compares false against every cutoff

__device__ __forceinline__ void lc_load4(uint4 (&v)[4], const char* p, int64_t step, int it, int nit) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (it + u < nit) v[u] = ldg_noalloc(p + (int64_t)u * step);
    else v[u] = make_uint4(kLcNaN2, kLcNaN2, kLcNaN2, kLcNaN2);
  }
}
// the same without the range checks (steady state of the streaming loops: four rows, three adds)
__device__ __forceinline__ void lc_load4_full(uint4 (&v)[4], const char* p, int64_t step, int64_t step2, int64_t step3) {
  v[0] = ldg_noalloc(p);
  v[1] = ldg_noalloc(p + step);
  v[2] = ldg_noalloc(p + step2);
  v[3] = ldg_noalloc(p + step3);
}

// ---- asynchronous streaming ring.  A thread's vectors go global -> shared with cp.async (no registers hold data in
// flight), kRing groups of four rows deep; the thread reads back only its own slots, so the only synchronisation is
// cp.async.wait_group.  Slot (group, u) of thread t lives at ring + ((group * 4 + u) * kLcThreads + t) * 16: consecutive
// threads, consecutive 16 bytes -- conflict-free both ways.  With 3-4 CTAs of 256 threads per SM this keeps 100-190 KB per
// SM in flight (round 2: the register double buffer held 48 KB and the loops were latency bound, ncu long_scoreboard).
__device__ __forceinline__ void lc_cp16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void lc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lc_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr uint32_t kLcSlotStride = kLcThreads * 16u;  // bytes between the slots u and u + 1 of a group

// rows it0 .. it0 + 3 of this thread's column into the group at `group_sa` (missing rows read as NaN pairs), one commit
__device__ __forceinline__ void lc_issue_group(uint32_t group_sa, const char* p, int64_t step, int it0, int nit) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (it0 + u < nit) lc_cp16(group_sa + (uint32_t)u * kLcSlotStride, p + (int64_t)u * step);
    else lc_sts_v4(group_sa + (uint32_t)u * kLcSlotStride, make_uint4(kLcNaN2, kLcNaN2, kLcNaN2, kLcNaN2));
  }
  lc_cp_commit();
}

// column tables of a CTA (shared memory): q = sqrt(scaler_row) and the cutoffs of two keys, for its 128 columns
struct LcTables {
  __align__(16) float q[kLcSlabCols];
  __align__(16) uint16_t ca[kLcSlabCols];
  __align__(16) uint16_t cb[kLcSlabCols];
};

// `sval` = scaler_row of column slab * 128 + tid (loaded by the caller, early), threads tid < 128
template <int DT>
__device__ __forceinline__ void lc_tables(const LcMat& M, uint32_t slab, float sval, int32_t TA, int32_t TB, LcTables& T, unsigned* fallback) {
  const int tid = threadIdx.x;
  if (tid < kLcSlabCols) {
    const uint32_t c = slab * kLcSlabCols + tid;
    float q = 0.f;
    uint32_t ca = lc_none<DT>(), cb = lc_none<DT>();
    if (c < M.C) {
      q = __fadd_rn(sqrtf(sval), 0.f);
      if (!(q <= 3.0e38f)) {  // NaN or inf norm: the exact-score path handles it
        atomicExch(fallback, 1u);
        q = 0.f;
      }
      ca = lc_cutoff<DT>(TA, q);
      cb = TB == TA ? ca : lc_cutoff<DT>(TB, q);
    }
    T.q[tid] = q;
    T.ca[tid] = (uint16_t)ca;
    T.cb[tid] = (uint16_t)cb;
  }
}

// ------------------------------------------------------------------------------------------------ K1: count
constexpr int kLcStash = 1792;  // bracket vectors a CTA parks in shared memory (expected ~20 % of its ~7 000 vectors)

constexpr int kLcRingCount = 2;  // groups in flight per thread in K1 (32 KB of ring next to the 30 KB stash: 3 CTAs per SM)
constexpr int kLcRingApply = 3;  // ... in K3 (48 KB, 4 CTAs per SM)

constexpr size_t kLcRingCountBytes = (size_t)kLcRingCount * 4 * kLcThreads * 16;
constexpr size_t kLcRingApplyBytes = (size_t)kLcRingApply * 4 * kLcThreads * 16;

struct LcCountShared {
  LcTables T;
  __align__(16) uint4 stash[kLcStash];
  uint8_t stash_l16[kLcStash];
  unsigned hist[kLcCoarse];
  unsigned stash_n;
  unsigned cnt[2];
  int last;
};

// exact keys of the bracket elements of one parked vector -> the CTA's shared histogram; returns how many
template <int DT>
__device__ __forceinline__ unsigned lc_hist_vec(const uint4& v, uint32_t l16, uint32_t vec_saddr, const LcTables& T, uint32_t hist_saddr,
                                                int32_t lo, int shift) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  const uint4 cl4 = *reinterpret_cast<const uint4*>(T.ca + l16 * 8), ch4 = *reinterpret_cast<const uint4*>(T.cb + l16 * 8);
  const uint32_t cl[4] = {cl4.x, cl4.y, cl4.z, cl4.w}, ch[4] = {ch4.x, ch4.y, ch4.z, ch4.w};
  uint32_t x[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = w[j] & 0x7fff7fffu;
    x[j] = lc_le2<DT>(a, cl[j]) ^ lc_le2<DT>(a, ch[j]);
  }
  uint32_t bits = lc_bits8(x);
  unsigned n = 0;
  const uint32_t q_saddr = lc_saddr(T.q + l16 * 8);
  while (bits) {  // the cutoffs are exact: every flagged element has lo < key <= hi
    const int e = __ffs((int)bits) - 1;
    bits &= bits - 1;
    const uint32_t m = lc_lds_u16(vec_saddr + 2u * e) & 0x7fffu;
    const uint32_t key = score_key(__fmul_rn(lc_mag<DT>(m), lc_lds_f32(q_saddr + 4u * e)));
    lc_red_shared(hist_saddr + 4u * ((key - (uint32_t)(lo + 1)) >> shift), 1u);
    ++n;
  }
  return n;
}

template <int DT>
__global__ void __launch_bounds__(kLcThreads, kLcCtasPerSm) lc_count_kernel(const __grid_constant__ LcBatch b) {
  extern __shared__ __align__(16) uint4 lc_ring[];  // [kLcRingCount][4][kLcThreads]
  __shared__ LcCountShared S;
  static_assert(kLcCoarse == kLcThreads, "one coarse bin per thread");
  constexpr int G = kLcRingCount;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int l16 = tid & (kLcSlabVecs - 1);
  LcItem I;
  lc_item(b, I);
  const LcMat& M = b.m[I.mi];
  const int64_t step = (int64_t)kLcRowsPerIter * M.ld * 2, step4 = 4 * step;
  const char* p0 = reinterpret_cast<const char*>(M.W) + ((int64_t)(I.row0 + I.rsub) * M.ld + (int64_t)I.colvec * 8) * 2;
  // everything this CTA needs from memory goes out at once: its first groups of vectors, the flag, the bracket, its norms
  const uint32_t ring_sa = lc_saddr(lc_ring) + (uint32_t)tid * 16u;
  const int ngroups = (I.nit + 3) >> 2;
  const char* pi = p0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    lc_issue_group(ring_sa + (uint32_t)g * 4u * kLcSlotStride, pi, step, 4 * g, I.nit);
    pi += step4;
  }
  lc_pdl_trigger();
  float sval = 0.f;
  if (tid < kLcSlabCols && I.slab * kLcSlabCols + tid < M.C) sval = M.s[I.slab * kLcSlabCols + tid];
  lc_pdl_wait();  // K0 is complete: flag, bracket and the cleared counters are visible
  const unsigned flag = *b.fallback;
  const int4 sel = *reinterpret_cast<const int4*>(M.sel);
  if (tid < 2) S.cnt[tid] = 0u;
  if (tid == 0) S.stash_n = 0u;
  S.hist[tid] = 0u;
  if (flag != 0u) {  // K0 gave the block up (its bracket words are not meaningful then)
    lc_cp_wait<0>();
    return;
  }
  const int32_t lo = sel.x, hi = sel.y;
  const int shift = sel.z;
  lc_tables<DT>(M, I.slab, sval, lo, hi, S.T, b.fallback + 2);
  __syncthreads();
  const uint4 cl4 = *reinterpret_cast<const uint4*>(S.T.ca + l16 * 8), ch4 = *reinterpret_cast<const uint4*>(S.T.cb + l16 * 8);
  const uint32_t cl[4] = {cl4.x, cl4.y, cl4.z, cl4.w}, ch[4] = {ch4.x, ch4.y, ch4.z, ch4.w};
  const uint32_t stash_sa = lc_saddr(S.stash), l16_sa = lc_saddr(S.stash_l16), n_sa = lc_saddr(&S.stash_n), hist_sa = lc_saddr(S.hist);
  unsigned acc = 0, n_in = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  // four vectors: count against both cutoffs; vectors holding a bracket element are parked in shared memory (raw bits +
  // column lane) and finished by the whole CTA after the stream -- no second trip to L2, no dependency chain in the stream.
  // Both half-warps take the same number of trips (the append is a warp collective; missing rows read as NaN).
  const int ngroups_w = __reduce_max_sync(0xffffffffu, ngroups);
  int slot = 0;
  for (int g = 0; g < ngroups_w; ++g) {
    lc_cp_wait<G - 1>();
    const uint32_t gsa = ring_sa + (uint32_t)slot * 4u * kLcSlotStride;
    uint4 cu[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) cu[u] = lc_lds_v4(gsa + (uint32_t)u * kLcSlotStride);
    uint32_t hit4 = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t w[4] = {cu[u].x, cu[u].y, cu[u].z, cu[u].w};
      uint32_t x = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t a = w[j] & 0x7fff7fffu;
        const uint32_t ml = lc_le2<DT>(a, cl[j]), mh = lc_le2<DT>(a, ch[j]);
        acc = __dp4a(ml, 0x01010101u, acc);  // += 510 per weight at or below lo
        x |= ml ^ mh;
      }
      hit4 |= (x != 0u ? 1u : 0u) << u;
    }
    // one warp-aggregated append per four vectors: the per-lane counts (0..4) are scanned with three ballots, one shared
    // atomic per warp (per-lane atomics on the CTA's single cursor were measured slower: they serialise across 8 warps)
    const int c = __popc(hit4);
    const unsigned b0 = __ballot_sync(0xffffffffu, c & 1), b1 = __ballot_sync(0xffffffffu, c & 2), b2 = __ballot_sync(0xffffffffu, c & 4);
    if (b0 | b1 | b2) {
      const unsigned total = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
      unsigned pos = 0;
      if (lane == 0) pos = lc_atom_shared(n_sa, total);
      pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(b0 & lt_mask) + 2u * __popc(b1 & lt_mask) + 4u * __popc(b2 & lt_mask);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (hit4 & (1u << u)) {
          if (pos < (unsigned)kLcStash) {
            lc_sts_v4(stash_sa + 16u * pos, cu[u]);
            lc_sts_u8(l16_sa + pos, (uint32_t)l16);
          } else {  // stash full (heavy ties inside the bracket): finish this vector on the spot, from its ring slot
            n_in += lc_hist_vec<DT>(cu[u], (uint32_t)l16, gsa + (uint32_t)u * kLcSlotStride, S.T, hist_sa, lo, shift);
          }
          ++pos;
        }
      }
    }
    // refill the slot just consumed (its contents are in registers / the stash by now) with group g + G
    if (g + G < ngroups_w) lc_issue_group(gsa, pi, step, 4 * (g + G), I.nit);
    else lc_cp_commit();  // an empty group keeps wait_group's count in step
    pi += step4;
    slot = slot + 1 == G ? 0 : slot + 1;
  }
  __syncthreads();
  const unsigned n_stash = min(S.stash_n, (unsigned)kLcStash);
  for (unsigned i = tid; i < n_stash; i += kLcThreads)
    n_in += lc_hist_vec<DT>(lc_lds_v4(stash_sa + 16u * i), lc_lds_u8(l16_sa + i), stash_sa + 16u * i, S.T, hist_sa, lo, shift);
  const unsigned c_lo = __reduce_add_sync(0xffffffffu, acc / 510u), c_in = __reduce_add_sync(0xffffffffu, n_in);
  if (lane == 0) {
    if (c_lo) atomicAdd(&S.cnt[0], c_lo);
    if (c_in) atomicAdd(&S.cnt[1], c_in);
  }
  __syncthreads();
  {
    // the CTA's ~1 500 bracket elements leave as <= 256 REDs, one per thread, each to its own L2 slice
    const unsigned cc = S.hist[tid];
    if (cc) atomicAdd(M.hist + tid * kLcCoarseStride, cc);
  }
  if (tid < 2 && S.cnt[tid]) atomicAdd(M.cnt + tid, (unsigned long long)S.cnt[tid]);
  // ---- the last CTA of the grid: bracket check, the bin of the k-th score, what K3 / K4 need.  One warp per matrix.
  __threadfence();
  __syncthreads();
  if (tid == 0) S.last = atomicAdd(b.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!S.last) return;
  __threadfence();
  if (tid == 0 && __ldcg(b.fallback + 2) != 0u) atomicExch(b.fallback, 1u);
  static_assert(ECF_LAYER_MAX_BATCH <= kLcThreads / 32, "one warp per matrix");
  if (wid < b.n) {
    const LcMat& Mw = b.m[wid];
    constexpr int PER = kLcCoarse / 32;  // 8 consecutive bins per lane
    unsigned loc[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) loc[j] = __ldcg(Mw.hist + (lane * PER + j) * kLcCoarseStride);
    const unsigned long long c_lo2 = __ldcg(Mw.cnt), c_in2 = __ldcg(Mw.cnt + 1), kth = (unsigned long long)Mw.kth;
    const int32_t lo2 = Mw.sel[0], hi2 = Mw.sel[1];
    const int shift2 = Mw.sel[2];
    const bool ok = kth >= c_lo2 && kth < c_lo2 + c_in2;  // the k-th score lies inside the sampled bracket
    const unsigned long long rem0 = kth - c_lo2;
    unsigned long long sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) sum += loc[j];
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    unsigned long long run = inc - sum;
    if (ok) {
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (rem0 >= run && rem0 < run + loc[j]) {
          const uint32_t bin = (uint32_t)(lane * PER + j);
          const uint32_t first = (uint32_t)(lo2 + 1) + (bin << shift2);
          uint32_t last = first + ((1u << shift2) - 1u);
          if (last > (uint32_t)hi2) last = (uint32_t)hi2;
          Mw.res[0] = shift2 == 0 ? (int32_t)first : (int32_t)first - 1;
          Mw.res[1] = (int32_t)last;
          Mw.res[2] = (int32_t)(rem0 - run);
          Mw.res[3] = (int32_t)loc[j];
          Mw.res[4] = (int32_t)first;
          Mw.res[5] = shift2;
          if (shift2 != 0 && loc[j] > (unsigned)kLcListCap) atomicExch(b.fallback, 1u);  // (heavy ties inside the bin)
        }
        run += loc[j];
      }
    } else if (lane == 0) {
      atomicExch(b.fallback, 1u);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K3: apply
template <int DT, bool EXTRAS>
__global__ void __launch_bounds__(kLcThreads, EXTRAS ? 3 : 4) lc_apply_kernel(const __grid_constant__ LcBatch b) {
  extern __shared__ __align__(16) uint4 lc_ring[];  // [kLcRingApply][4][kLcThreads]
  __shared__ LcTables T;
  __shared__ unsigned sh_zero;
  constexpr int G = kLcRingApply;
  const int tid = threadIdx.x, lane = tid & 31;
  const int l16 = tid & (kLcSlabVecs - 1);
  LcItem I;
  lc_item(b, I);
  const LcMat& M = b.m[I.mi];
  const int64_t step = (int64_t)kLcRowsPerIter * M.ld * 2, step4 = 4 * step;
  char* p0 = reinterpret_cast<char*>(M.W) + ((int64_t)(I.row0 + I.rsub) * M.ld + (int64_t)I.colvec * 8) * 2;
  const uint32_t ring_sa = lc_saddr(lc_ring) + (uint32_t)tid * 16u;
  const int ngroups = (I.nit + 3) >> 2;
  const char* pi = p0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    lc_issue_group(ring_sa + (uint32_t)g * 4u * kLcSlotStride, pi, step, 4 * g, I.nit);
    pi += step4;
  }
  lc_pdl_trigger();
  float sval = 0.f;
  if (tid < kLcSlabCols && I.slab * kLcSlabCols + tid < M.C) sval = M.s[I.slab * kLcSlabCols + tid];
  if (tid == 0) sh_zero = 0u;
  lc_pdl_wait();  // K1 is complete (it only read W, so the ring loads above were safe to issue early)
  const unsigned flag = *b.fallback;  // final
  const int2 tt = *reinterpret_cast<const int2*>(M.res);
  if (flag != 0u) {
    lc_cp_wait<0>();
    return;
  }
  const int32_t TA = tt.x, TB = tt.y;
  lc_tables<DT>(M, I.slab, sval, TA, TB, T, b.fallback + 2);  // (cannot fire: K1 saw the same norms)
  __syncthreads();
  const uint4 ca4 = *reinterpret_cast<const uint4*>(T.ca + l16 * 8), cb4 = *reinterpret_cast<const uint4*>(T.cb + l16 * 8);
  const uint32_t ca[4] = {ca4.x, ca4.y, ca4.z, ca4.w}, cb[4] = {cb4.x, cb4.y, cb4.z, cb4.w};
  // EXTRAS: somebody asked for the zero count or the packed mask (tests, check_sparsity); the plain variant is leaner
  const bool want_zero = EXTRAS && M.n_zero != nullptr, want_mask = EXTRAS && M.mask != nullptr;
  uint8_t* mrow = want_mask ? M.mask + (int64_t)(I.row0 + I.rsub) * M.mask_ld + I.colvec : nullptr;
  const int64_t mstep = (int64_t)kLcRowsPerIter * M.mask_ld;
  const uint32_t q_sa = lc_saddr(T.q + l16 * 8);
  unsigned zacc = 0;
  // four vectors of this thread's column: zero what lies at or below cutoff A, store; elements between the cutoffs (the k-th
  // score's bin, ~1 500 per matrix -- one lane in a few thousand vectors) go to the list with their exact keys, one global
  // reservation per lane that has any.  TA == TB makes the cutoffs equal, so the bin test is empty without a branch.
  auto process = [&](const uint4 (&cu)[4], char* pc, int it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t w[4] = {cu[u].x, cu[u].y, cu[u].z, cu[u].w};
      uint32_t xs[4];
      uint32_t x = 0, any = 0, mb = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t a = w[j] & 0x7fff7fffu;
        const uint32_t ma = lc_le2<DT>(a, ca[j]);
        xs[j] = ma ^ lc_le2<DT>(a, cb[j]);
        x |= xs[j];
        any |= ma;
        w[j] &= ~ma;
        if (want_zero) zacc = __dp4a(lc_eq0<DT>(w[j]), 0x01010101u, zacc);
        if (want_mask) mb |= ((ma & 1u) | ((ma >> 15) & 2u)) << (2 * j);
      }
      if (any) stg_v4(pc + (int64_t)u * step, make_uint4(w[0], w[1], w[2], w[3]));
      if (want_mask && it + u < I.nit) mrow[(int64_t)(it + u) * mstep] = (uint8_t)mb;
      if (x != 0u) {
        uint32_t bits = lc_bits8(xs);
        unsigned pos = atomicAdd(M.list_n, (unsigned)__popc(bits));
        const uint32_t row = I.row0 + I.rsub + (uint32_t)(it + u) * kLcRowsPerIter;
        const uint32_t w0[4] = {cu[u].x, cu[u].y, cu[u].z, cu[u].w};
        while (bits) {
          const int e = __ffs((int)bits) - 1;
          bits &= bits - 1;
          uint32_t pr = w0[0];  // element e of the vector, without indexing the register array dynamically
          if ((e >> 1) == 1) pr = w0[1];
          if ((e >> 1) == 2) pr = w0[2];
          if ((e >> 1) == 3) pr = w0[3];
          const uint32_t m = (pr >> ((e & 1) * 16)) & 0x7fffu;
          const uint32_t key = score_key(__fmul_rn(lc_mag<DT>(m), lc_lds_f32(q_sa + 4u * e)));
          if (pos < (unsigned)kLcListCap) M.list[pos] = make_uint2(row * M.C + I.colvec * 8 + (uint32_t)e, key);
          ++pos;
        }
      }
    }
  };
  char* pc = p0;
  int slot = 0;
  for (int g = 0; g < ngroups; ++g) {
    lc_cp_wait<G - 1>();
    const uint32_t gsa = ring_sa + (uint32_t)slot * 4u * kLcSlotStride;
    uint4 cu[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) cu[u] = lc_lds_v4(gsa + (uint32_t)u * kLcSlotStride);
    process(cu, pc, 4 * g);  // (missing rows read as NaN: never below a cutoff, never stored)
    if (g + G < ngroups) lc_issue_group(gsa, pi, step, 4 * (g + G), I.nit);
    else lc_cp_commit();
    pi += step4;
    pc += step4;
    slot = slot + 1 == G ? 0 : slot + 1;
  }
  if (want_zero) {
    const unsigned z = __reduce_add_sync(0xffffffffu, zacc / 510u);
    if (lane == 0 && z) atomicAdd(&sh_zero, z);
    __syncthreads();
    if (tid == 0 && sh_zero) atomicAdd(M.n_zero, (unsigned long long)sh_zero);
  }
}

// ------------------------------------------------------------------------------------------------ K4: fix-up
// One CTA per matrix (block size a multiple of 32, >= 256): exact k-th key among the listed elements of the bin, 11 bits of
// (key - first) per level (a bin is 2^shift keys wide), threshold out, zero the listed elements at or below it.
// `sh_h`: kLcBins words, `sh_keys`: kLcListCap words of shared memory.
__device__ __forceinline__ uint32_t lc_scan_bins(const unsigned* sh_h, unsigned rem, unsigned* sh_w /*[34]*/) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) sh_w[32] = 0xffffffffu;
  unsigned loc[8], sum = 0;
  const bool scan = tid < kLcBins / 8;
  if (scan) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      loc[j] = sh_h[tid * 8 + j];
      sum += loc[j];
    }
  }
  unsigned inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (scan && lane == 31) sh_w[wid] = inc;
  __syncthreads();
  if (scan) {
    unsigned run = inc - sum;
    for (int w = 0; w < wid; ++w) run += sh_w[w];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (rem >= run && rem < run + loc[j]) {
        sh_w[32] = (unsigned)(tid * 8 + j);
        sh_w[33] = rem - run;
      }
      run += loc[j];
    }
  }
  __syncthreads();
  return sh_w[32];  // 0xffffffff cannot happen: K1 counted exactly these elements
}

__device__ __forceinline__ void lc_fixup(const LcMat& M, const int4& r0, const int2& r1, unsigned n_list, unsigned* sh_h, uint32_t* sh_keys,
                                         unsigned* sh_w /*[34]*/) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
  const int32_t TA = r0.x, TB = r0.y;
  if (TA == TB) {  // the histogram resolved the key exactly
    if (tid == 0 && M.thres_out != nullptr) *M.thres_out = __uint_as_float((uint32_t)TA);
    return;
  }
  unsigned rem = (unsigned)r0.z;
  uint32_t base = (uint32_t)r1.x;
  int width = r1.y;  // the k-th key lies in [base, base + 2^width)
  const unsigned n = min(n_list, (unsigned)kLcListCap);
  for (unsigned i = tid; i < n; i += nthr) sh_keys[i] = M.list[i].y;  // one trip to L2; the levels work on shared memory
  while (width > 0) {
    const int s1 = max(0, width - 11);
    for (int i = tid; i < kLcBins; i += nthr) sh_h[i] = 0u;
    __syncthreads();
    for (unsigned i = tid; i < n; i += nthr) {
      const uint32_t key = sh_keys[i];
      if (key >= base && ((key - base) >> width) == 0u) atomicAdd(&sh_h[(key - base) >> s1], 1u);
    }
    __syncthreads();
    const uint32_t d = lc_scan_bins(sh_h, rem, sh_w);
    rem = sh_w[33];
    __syncthreads();
    base += d << s1;
    width = s1;
  }
  const uint32_t tkey = base;
  if (tid == 0 && M.thres_out != nullptr) *M.thres_out = __uint_as_float(tkey);
  unsigned zeros = 0;
  for (unsigned i = tid; i < n; i += nthr) {
    if (sh_keys[i] <= tkey) {
      const uint32_t idx = M.list[i].x;
      const uint32_t row = idx / M.C, col = idx - row * M.C;
      uint16_t* w = reinterpret_cast<uint16_t*>(M.W) + (int64_t)row * M.ld + col;
      if (M.n_zero != nullptr && (*w & 0x7fffu) != 0u) ++zeros;
      *w = 0;
      if (M.mask != nullptr) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(M.mask + (int64_t)row * M.mask_ld + (col >> 3));
        atomicOr(reinterpret_cast<unsigned*>(addr & ~uintptr_t(3)), 1u << ((addr & 3) * 8 + (col & 7)));
      }
    }
  }
  if (M.n_zero != nullptr) {
    zeros = __reduce_add_sync(0xffffffffu, zeros);
    if (lane == 0 && zeros) atomicAdd(M.n_zero, (unsigned long long)zeros);
  }
}

// ------------------------------------------------------------------------------------------------ K4: finish
// One thread-block CLUSTER of kLcCluster CTAs per matrix.  Normal case: CTA 0 of the cluster runs the fix-up above, the
// others leave at once.  Fallback (flag set before K3 wrote anything): the cluster runs an exact three-digit radix select
// over the matrix' fp32 score keys (11 + 11 + 9 bits) with per-CTA shared histograms merged through distributed shared
// memory and the hardware cluster barrier -- no cooperative launch, no capacity limits (heavy ties, non-finite norms,
// NaN / inf scores all take this path) -- and applies score <= thres.  Slow (four L2 passes by 8 SMs) but exact.
constexpr int kLcCluster = 8;
constexpr int kLcFinishThreads = 512;

template <int DT>
__device__ __forceinline__ void lc_exact_keys(const LcMat& M, uint32_t row, uint32_t colvec, uint32_t (&key)[8], uint4& raw) {
  raw = ldg_v4(reinterpret_cast<const char*>(M.W) + ((int64_t)row * M.ld + (int64_t)colvec * 8) * 2);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  const float4 sa = *reinterpret_cast<const float4*>(M.s + colvec * 8), sb = *reinterpret_cast<const float4*>(M.s + colvec * 8 + 4);
  const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float w0, w1;
    unpack2<DT>(w[e >> 1], w0, w1);
    key[e] = score_key(wanda_score((e & 1) ? w1 : w0, __fadd_rn(sqrtf(sv[e]), 0.f)));
  }
}

template <int DT>
__device__ __noinline__ void lc_exact_select(const LcMat& M, unsigned* sh_h, unsigned* sh_tot, unsigned long long* sh_w, uint32_t* sh_sel) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t r0 = (uint32_t)(((uint64_t)M.R * rank) / kLcCluster), r1 = (uint32_t)(((uint64_t)M.R * (rank + 1)) / kLcCluster);
  const uint64_t nv = (uint64_t)(r1 - r0) * M.nvpr;
  unsigned* tot0 = cluster.map_shared_rank(sh_tot, 0);
  uint32_t prefix = 0;
  unsigned long long rem = (unsigned long long)M.kth;
  const int shifts[3] = {20, 9, 0}, widths[3] = {11, 11, 9};
  for (int pass = 0; pass < 3; ++pass) {
    const int sh = shifts[pass], wd = widths[pass];
    for (int i = tid; i < kLcBins; i += kLcFinishThreads) {
      sh_h[i] = 0u;
      sh_tot[i] = 0u;
    }
    cluster.sync();  // every CTA's total histogram is clear before anybody adds to rank 0's
    for (uint64_t v = tid; v < nv; v += kLcFinishThreads) {
      const uint32_t row = r0 + (uint32_t)(v / M.nvpr), colvec = (uint32_t)(v % M.nvpr);
      uint32_t key[8];
      uint4 raw;
      lc_exact_keys<DT>(M, row, colvec, key, raw);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (pass == 0 || (key[e] >> (sh + wd)) == prefix) atomicAdd(&sh_h[(key[e] >> sh) & ((1u << wd) - 1u)], 1u);
      }
    }
    __syncthreads();
    for (int i = tid; i < kLcBins; i += kLcFinishThreads) {
      const unsigned c = sh_h[i];
      if (c) atomicAdd(tot0 + i, c);
    }
    cluster.sync();
    // every CTA scans rank 0's totals on its own (4 bins per thread): the bin holding rank `rem`
    unsigned loc[kLcBins / kLcFinishThreads];
    unsigned long long sum = 0;
#pragma unroll
    for (int j = 0; j < kLcBins / kLcFinishThreads; ++j) {
      loc[j] = tot0[tid * (kLcBins / kLcFinishThreads) + j];
      sum += loc[j];
    }
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sh_w[wid] = inc;
    __syncthreads();
    unsigned long long run = inc - sum;
    for (int w = 0; w < wid; ++w) run += sh_w[w];
#pragma unroll
    for (int j = 0; j < kLcBins / kLcFinishThreads; ++j) {
      if (rem >= run && rem < run + loc[j]) {
        sh_sel[0] = (uint32_t)(tid * (kLcBins / kLcFinishThreads) + j);
        sh_sel[1] = (uint32_t)(rem - run);
      }
      run += loc[j];
    }
    cluster.sync();  // all CTAs have read rank 0's totals (and this CTA's sh_sel is written)
    prefix = (prefix << wd) | sh_sel[0];
    rem = sh_sel[1];
  }
  const uint32_t tkey = prefix;
  // float semantics of `W_metric <= thres`: NaN scores are never pruned, a NaN threshold prunes nothing
  const uint32_t tcmp = tkey > 0x7f800000u ? 0u : tkey + 1u;
  if (rank == 0 && tid == 0 && M.thres_out != nullptr) *M.thres_out = __uint_as_float(tkey);
  unsigned zeros = 0;
  for (uint64_t v = tid; v < nv; v += kLcFinishThreads) {
    const uint32_t row = r0 + (uint32_t)(v / M.nvpr), colvec = (uint32_t)(v % M.nvpr);
    uint32_t key[8];
    uint4 raw;
    lc_exact_keys<DT>(M, row, colvec, key, raw);
    uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t mb = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (key[e] < tcmp) {
        w[e >> 1] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
        mb |= 1u << e;
      }
    }
    if (mb) stg_v4(reinterpret_cast<char*>(M.W) + ((int64_t)row * M.ld + (int64_t)colvec * 8) * 2, make_uint4(w[0], w[1], w[2], w[3]));
    if (M.mask != nullptr) M.mask[(int64_t)row * M.mask_ld + colvec] = (uint8_t)mb;
    if (M.n_zero != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) zeros += ((w[j] & 0x00007fffu) == 0 ? 1 : 0) + ((w[j] & 0x7fff0000u) == 0 ? 1 : 0);
    }
  }
  if (M.n_zero != nullptr) {
    zeros = __reduce_add_sync(0xffffffffu, zeros);
    if (lane == 0 && zeros) atomicAdd(M.n_zero, (unsigned long long)zeros);
  }
  cluster.sync();  // nobody leaves while a peer may still touch its shared memory
}

__global__ void __cluster_dims__(kLcCluster, 1, 1) __launch_bounds__(kLcFinishThreads, 1)
    lc_finish_kernel(const __grid_constant__ LcBatch b) {
  __shared__ unsigned sh_h[kLcBins];
  __shared__ uint32_t sh_big[kLcListCap];  // fix-up: the listed keys; exact select: the cluster's total histogram
  __shared__ unsigned long long sh_w[kLcFinishThreads / 32];
  __shared__ unsigned sh_fix[34];
  __shared__ uint32_t sh_sel[2];
  const int mi = (int)blockIdx.x / kLcCluster;
  const LcMat& M = b.m[mi];
  lc_pdl_wait();  // K3 is complete
  // all the words the fix-up depends on in one round trip
  const unsigned flag = *b.fallback;
  const int4 r0 = *reinterpret_cast<const int4*>(M.res);
  const int2 r1 = *reinterpret_cast<const int2*>(M.res + 4);
  const unsigned n_list = *M.list_n;
  if (flag == 0u) {
    if (blockIdx.x % kLcCluster == 0) lc_fixup(M, r0, r1, n_list, sh_h, sh_big, sh_fix);
  } else if (M.dtype == ECF_F16) {
    lc_exact_select<ECF_F16>(M, sh_h, sh_big, sh_w, sh_sel);
  } else {
    lc_exact_select<ECF_BF16>(M, sh_h, sh_big, sh_w, sh_sel);
  }
  // the last CTA of the launch leaves the header ready for the next one: flag -> "last launch" word, flag and ticket clear
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(b.ticket + 1, 1u) == gridDim.x - 1) {
      b.fallback[1] = flag;
      b.fallback[2] = 0u;
      b.ticket[1] = 0u;
      __threadfence();
      atomicExch(b.fallback, 0u);
    }
  }
}

}  // namespace ecf
