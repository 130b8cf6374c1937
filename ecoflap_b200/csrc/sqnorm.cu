// A1 -- per-input-channel squared-norm accumulation of [T, C] activation blocks, batched over the Linears of a
// transformer block.
//
// Replaces WrappedGPT.add_batch (LAVIS/lavis/compression/pruners/wanda_pruner.py:71-84; CoOp wanda_pruner.py:
// 159-172; UPop wanda_pruner.py:65-78): the reference materialises an fp32 [C, T] copy per hooked Linear, runs
// torch.norm (sqrt) and squares it again -- three small kernels per hook call, 9 408 hook calls for BLIP-2.  Here
// every X is streamed exactly once with 128-bit loads and accumulated in fp32 registers, and ONE launch serves
// all the hook calls of a block forward (up to kSqMaxDesc descriptors: the 4-11 Linears of a block see 2-50 MB
// together, where a launch per Linear would be pure launch latency).  q/k/v (and wi_0/wi_1) share their input:
// their descriptors read the same X inside one launch, so the repeats are L2 hits, not HBM traffic.
//
// Descriptors that update the SAME accumulator (the 16 calibration batches of a block) are merged into one group:
// the reference's sequential  s = s*n/(n+B) + sum/(n+B)  has the closed form  s*prod(r_j) + sum_j w_j*colsum_j  with
// w_j = inv_n_j * prod_{l>j} r_l, so a whole block's calibration sweep can be ONE launch of hundreds of MB instead of
// 16 launches whose 10-50 MB each are dominated by launch ramp and reduction tail.
//
// Layout: a CTA is 32 column-vectors (16 bytes each: 512 contiguous bytes per warp request) by 8 row lanes and
// owns `rows_per_cta` tokens of one 32-vector column tile of one hook call.  CTAs are ordered group-major,
// tile-major, (call, split)-minor; each writes one weighted partial row into the workspace slot of its own block
// index, and the last CTA to finish a tile (atomic ticket) adds that tile's partials in order -- deterministic -- and
// applies   scaler_row = scaler_row * rescale + sum.  Tickets reset themselves: the first kSqCounterBytes of the
// workspace must be zero before the FIRST call only (no memset node per call).
// Bound: HBM.  Algorithmic bytes per hook call: T*C*sizeof(x) + 8*C.
#include "common.cuh"

namespace ecf {

constexpr int kSqTX = 32;      // column vectors per CTA
constexpr int kSqTY = 8;       // row lanes per CTA
constexpr int kSqUnroll = 8;   // independent 16-byte loads in flight per thread
constexpr int kSqMaxSeg = ECF_SQNORM_MAX_BATCH;  // hook calls per launch
constexpr int kSqMaxGroup = 32;                  // distinct accumulators per launch
constexpr int kSqCounterBytes = 64 * 1024;       // 16 384 column tiles per launch
constexpr int kSqMinRows = kSqTY * kSqUnroll;
constexpr int kSqPartialCols = kSqTX * 8;

// one hook call: a [T, C] activation block (C, dtype come from its group)
struct SqSeg {
  const void* x;
  int64_t T, ld;
  float weight;     // inv_n of this call times the rescale factors of the later calls on the same accumulator
  int split_begin;  // first split of this segment inside its group
};
// one accumulator: all the hook calls of the launch that update the same scaler_row, applied as
//   scaler_row = scaler_row * rescale + sum_j weight_j * colsum(x_j^2)
struct SqGroup {
  float* scaler_row;
  int64_t C;
  int dtype, vec;     // vec: 128-bit path usable by every segment (C, ld multiples of the vector width, aligned bases)
  float rescale;      // product of the per-call rescale factors
  int nx, splits;     // column tiles, token splits over all segments
  int seg_begin, seg_end;
  int cta_begin;      // first block index of this group
  int tile_begin;     // first ticket of this group
};

struct SqBatch {
  SqSeg seg[kSqMaxSeg];
  SqGroup grp[kSqMaxGroup];
  int n_groups;
  int64_t rows_per_cta;
};

template <int DT, bool VEC>
__device__ __forceinline__ void sq_accumulate(const SqSeg& sg, int64_t C, int tile, int split, int64_t rows_per_cta,
                                              float (&red)[kSqTY][kSqPartialCols], int& cols_out) {
  constexpr int V = VEC ? DType<DT>::kVec : 1;
  constexpr int kCols = kSqTX * V;
  cols_out = kCols;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t col0 = (int64_t)tile * kCols + (int64_t)tx * V;
  const bool col_ok = col0 < C;  // VEC: C % V == 0, so the whole vector is in range
  const int64_t t0 = (int64_t)split * rows_per_cta;
  const int64_t t1 = min(sg.T, t0 + rows_per_cta);

  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;

  if (col_ok) {
    const char* base = reinterpret_cast<const char*>(sg.x) + col0 * DType<DT>::kBytes;
    const int64_t row_bytes = sg.ld * DType<DT>::kBytes;
    for (int64_t t = t0 + ty; t < t1; t += kSqTY * kSqUnroll) {
      if constexpr (VEC) {
        uint4 buf[kSqUnroll];
#pragma unroll
        for (int u = 0; u < kSqUnroll; ++u) {
          const int64_t tt = t + (int64_t)u * kSqTY;
          buf[u] = tt < t1 ? ldg_stream(base + tt * row_bytes) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kSqUnroll; ++u) {
          const uint32_t w[4] = {buf[u].x, buf[u].y, buf[u].z, buf[u].w};
          if constexpr (DT == ECF_F32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float f = __uint_as_float(w[j]);
              acc[j] = fmaf(f, f, acc[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float lo, hi;
              unpack2<DT>(w[j], lo, hi);
              acc[2 * j] = fmaf(lo, lo, acc[2 * j]);
              acc[2 * j + 1] = fmaf(hi, hi, acc[2 * j + 1]);
            }
          }
        }
      } else {
        float buf[kSqUnroll];
#pragma unroll
        for (int u = 0; u < kSqUnroll; ++u) {
          const int64_t tt = t + (int64_t)u * kSqTY;
          buf[u] = tt < t1 ? load_elem<DT>(sg.x, tt * sg.ld + col0) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kSqUnroll; ++u) acc[0] = fmaf(buf[u], buf[u], acc[0]);
      }
    }
  }
  // cross-row-lane reduction through shared memory (vector stores: conflict free)
#pragma unroll
  for (int v = 0; v < V; ++v) red[ty][tx * V + v] = acc[v];
}

__global__ void __launch_bounds__(kSqTX* kSqTY)
    sqnorm_batched_kernel(const __grid_constant__ SqBatch b, float* __restrict__ partial, unsigned* __restrict__ counters) {
  __shared__ __align__(16) float red[kSqTY][kSqPartialCols];
  __shared__ bool is_last;
  // which group / tile / segment / split is this CTA?  (CTAs: group-major, tile-major, split-minor)
  int gi = 0;
  while (gi + 1 < b.n_groups && (int)blockIdx.x >= b.grp[gi + 1].cta_begin) ++gi;
  const SqGroup& g = b.grp[gi];
  const int local = (int)blockIdx.x - g.cta_begin;
  const int tile = local / g.splits, gsplit = local - tile * g.splits;
  int lo = g.seg_begin, hi = g.seg_end - 1;  // last segment whose split_begin <= gsplit
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (b.seg[mid].split_begin <= gsplit) lo = mid; else hi = mid - 1;
  }
  const SqSeg& sg = b.seg[lo];
  const int split = gsplit - sg.split_begin;
  const int tid = threadIdx.y * kSqTX + threadIdx.x;

  int kCols = 0;
  switch (g.dtype * 2 + g.vec) {
    case ECF_F32 * 2 + 1: sq_accumulate<ECF_F32, true>(sg, g.C, tile, split, b.rows_per_cta, red, kCols); break;
    case ECF_F32 * 2 + 0: sq_accumulate<ECF_F32, false>(sg, g.C, tile, split, b.rows_per_cta, red, kCols); break;
    case ECF_F16 * 2 + 1: sq_accumulate<ECF_F16, true>(sg, g.C, tile, split, b.rows_per_cta, red, kCols); break;
    case ECF_F16 * 2 + 0: sq_accumulate<ECF_F16, false>(sg, g.C, tile, split, b.rows_per_cta, red, kCols); break;
    case ECF_BF16 * 2 + 1: sq_accumulate<ECF_BF16, true>(sg, g.C, tile, split, b.rows_per_cta, red, kCols); break;
    default: sq_accumulate<ECF_BF16, false>(sg, g.C, tile, split, b.rows_per_cta, red, kCols); break;
  }
  __syncthreads();
  float* my_partial = partial + (int64_t)blockIdx.x * kSqPartialCols;
  if (tid < kCols) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kSqTY; ++r) s += red[r][tid];
    my_partial[tid] = s * sg.weight;
  }
  // ticket: the last CTA of this column tile folds the partials in (segment, split) order (deterministic)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(&counters[g.tile_begin + tile], 1u);
    is_last = (prev == (unsigned)g.splits - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int64_t out_col = (int64_t)tile * kCols + tid;
  if (tid < kCols && out_col < g.C) {
    const float* p0 = partial + (int64_t)(g.cta_begin + tile * g.splits) * kSqPartialCols + tid;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int sp = 0;
    for (; sp + 4 <= g.splits; sp += 4) {  // four independent L2 loads in flight
      s0 += __ldcg(p0 + (int64_t)(sp + 0) * kSqPartialCols);
      s1 += __ldcg(p0 + (int64_t)(sp + 1) * kSqPartialCols);
      s2 += __ldcg(p0 + (int64_t)(sp + 2) * kSqPartialCols);
      s3 += __ldcg(p0 + (int64_t)(sp + 3) * kSqPartialCols);
    }
    for (; sp < g.splits; ++sp) s0 += __ldcg(p0 + (int64_t)sp * kSqPartialCols);
    g.scaler_row[out_col] = g.scaler_row[out_col] * g.rescale + ((s0 + s1) + (s2 + s3));
  }
  if (tid == 0) counters[g.tile_begin + tile] = 0;  // self-cleaning for the next call on this workspace
}

// Fills the launch plan; returns the number of CTAs (negative status on error with the message set).
static int64_t sq_plan(const ecf_sqnorm_desc* descs, int n, SqBatch& b, int64_t& tiles_total) {
  const int sms = sm_count();
  // group the descriptors by accumulator, keeping call order inside a group
  int order[kSqMaxSeg];
  int ng = 0, ns = 0;
  bool used[kSqMaxSeg];
  for (int i = 0; i < n; ++i) used[i] = false;
  int64_t max_T = 1;
  for (int i = 0; i < n; ++i) {
    if (used[i]) continue;
    if (ng == kSqMaxGroup) {
      set_error("sqnorm: more than %d distinct accumulators in one launch", kSqMaxGroup);
      return ECF_ERR_INVALID;
    }
    SqGroup& g = b.grp[ng];
    const ecf_sqnorm_desc& s0 = descs[i];
    g.scaler_row = s0.scaler_row; g.C = s0.C; g.dtype = s0.dtype; g.seg_begin = ns;
    const int V = s0.dtype == ECF_F32 ? 4 : 8;
    bool vec = (s0.C % V == 0);
    double rescale = 1.0;
    for (int j = i; j < n; ++j) {
      if (descs[j].scaler_row != s0.scaler_row) continue;
      const ecf_sqnorm_desc& s = descs[j];
      if (s.C != s0.C || s.dtype != s0.dtype) {
        set_error("sqnorm: descriptors %d and %d share an accumulator but differ in C or dtype", i, j);
        return ECF_ERR_INVALID;
      }
      used[j] = true;
      order[ns++] = j;
      vec = vec && (s.ld % V == 0) && ((reinterpret_cast<uintptr_t>(s.x) & 15) == 0);
      rescale *= (double)s.rescale;
      if (s.T > max_T) max_T = s.T;
    }
    g.seg_end = ns;
    g.vec = vec ? 1 : 0;
    g.rescale = (float)rescale;
    const int64_t cols = (int64_t)kSqTX * (g.vec ? V : 1);
    g.nx = (int)((s0.C + cols - 1) / cols);
    // weight_j = inv_n_j * prod_{l > j} rescale_l   (sequential application, closed form)
    double tail = 1.0;
    for (int p = g.seg_end - 1; p >= g.seg_begin; --p) {
      const ecf_sqnorm_desc& s = descs[order[p]];
      SqSeg& sg = b.seg[p];
      sg.x = s.x; sg.T = s.T; sg.ld = s.ld;
      sg.weight = (float)((double)s.inv_n * tail);
      tail *= (double)s.rescale;
    }
    ++ng;
  }
  b.n_groups = ng;
  // rows per CTA: the smallest multiple of kSqMinRows that keeps the grid at <= 8 CTAs per SM
  int64_t rows = kSqMinRows;
  for (;;) {
    int64_t ctas = 0;
    for (int gi = 0; gi < ng; ++gi) {
      int64_t sp = 0;
      for (int p = b.grp[gi].seg_begin; p < b.grp[gi].seg_end; ++p) sp += (b.seg[p].T + rows - 1) / rows;
      ctas += sp * b.grp[gi].nx;
    }
    if (ctas <= (int64_t)sms * 8 || rows >= max_T) break;
    rows += kSqMinRows * ((ctas / ((int64_t)sms * 8) > 2) ? (ctas / ((int64_t)sms * 16)) : 1);
  }
  b.rows_per_cta = rows;
  int64_t cta = 0, tile = 0;
  for (int gi = 0; gi < ng; ++gi) {
    SqGroup& g = b.grp[gi];
    int64_t sp = 0;
    for (int p = g.seg_begin; p < g.seg_end; ++p) {
      b.seg[p].split_begin = (int)sp;
      sp += (b.seg[p].T + rows - 1) / rows;
    }
    g.splits = (int)sp;
    g.cta_begin = (int)cta;
    g.tile_begin = (int)tile;
    cta += (int64_t)g.nx * sp;
    tile += g.nx;
    if (cta >= (1ll << 31)) {
      set_error("sqnorm: grid too large");
      return ECF_ERR_INVALID;
    }
  }
  tiles_total = tile;
  return cta;
}

static int sq_check(const ecf_sqnorm_desc* descs, int n) {
  ECF_REQUIRE(descs != nullptr && n >= 1 && n <= kSqMaxSeg, ECF_ERR_INVALID, "sqnorm: batch size %d outside [1, %d]", n, kSqMaxSeg);
  for (int i = 0; i < n; ++i) {
    const ecf_sqnorm_desc& s = descs[i];
    ECF_REQUIRE(s.x != nullptr && s.scaler_row != nullptr, ECF_ERR_INVALID, "sqnorm: null pointer (descriptor %d)", i);
    ECF_REQUIRE(s.T >= 0 && s.C > 0 && s.ld >= s.C, ECF_ERR_INVALID, "sqnorm: bad shape T=%lld C=%lld ld=%lld (descriptor %d)",
                (long long)s.T, (long long)s.C, (long long)s.ld, i);
    ECF_REQUIRE(s.T > 0, ECF_ERR_INVALID, "sqnorm: empty activation block (T == 0, descriptor %d)", i);
    ECF_REQUIRE(s.dtype >= 0 && s.dtype <= 2, ECF_ERR_INVALID, "sqnorm: unknown dtype %d (descriptor %d)", s.dtype, i);
  }
  return ECF_OK;
}

size_t sqnorm_batched_workspace_bytes(const ecf_sqnorm_desc* descs, int n) {
  if (descs == nullptr || n < 1 || n > kSqMaxSeg) return 0;
  for (int i = 0; i < n; ++i)
    if (descs[i].T <= 0 || descs[i].C <= 0 || descs[i].dtype < 0 || descs[i].dtype > 2) return 0;
  SqBatch b;
  int64_t tiles = 0;
  const int64_t ctas = sq_plan(descs, n, b, tiles);
  if (ctas < 0) return 0;
  return (size_t)kSqCounterBytes + (size_t)ctas * kSqPartialCols * sizeof(float);
}

size_t sqnorm_workspace_bytes(int64_t T, int64_t C) {
  // worst case over dtypes / alignment for a single descriptor
  size_t worst = 0;
  for (int dt = 0; dt < 3; ++dt)
    for (int al = 0; al < 2; ++al) {
      ecf_sqnorm_desc s;
      s.x = reinterpret_cast<const void*>(al ? uintptr_t(16) : uintptr_t(2));
      s.scaler_row = reinterpret_cast<float*>(uintptr_t(16)); s.T = T > 0 ? T : 1; s.C = C > 0 ? C : 1; s.ld = s.C; s.dtype = dt;
      s.rescale = 0.f; s.inv_n = 0.f;
      const size_t b = sqnorm_batched_workspace_bytes(&s, 1);
      if (b > worst) worst = b;
    }
  return worst;
}

}  // namespace ecf

extern "C" size_t ecf_sqnorm_batched_workspace_bytes(const ecf_sqnorm_desc* descs, int n) {
  return ecf::sqnorm_batched_workspace_bytes(descs, n);
}

extern "C" int ecf_sqnorm_accum_batched(const ecf_sqnorm_desc* descs, int n, void* ws, size_t ws_bytes, ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  if ((st = sq_check(descs, n)) != ECF_OK) return st;
  SqBatch b;
  int64_t tiles = 0;
  const int64_t ctas = sq_plan(descs, n, b, tiles);
  if (ctas < 0) return (int)ctas;
  ECF_REQUIRE(tiles * (int64_t)sizeof(unsigned) <= kSqCounterBytes, ECF_ERR_INVALID, "sqnorm: %lld column tiles in one launch (max %d)",
              (long long)tiles, kSqCounterBytes / (int)sizeof(unsigned));
  const size_t need = (size_t)kSqCounterBytes + (size_t)ctas * kSqPartialCols * sizeof(float);
  ECF_REQUIRE(ws != nullptr && ws_bytes >= need, ECF_ERR_WORKSPACE, "sqnorm: workspace %zu < %zu bytes", ws_bytes, need);
  unsigned* counters = reinterpret_cast<unsigned*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kSqCounterBytes);
  sqnorm_batched_kernel<<<(unsigned)ctas, dim3(kSqTX, kSqTY), 0, reinterpret_cast<cudaStream_t>(stream)>>>(b, partial, counters);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

extern "C" int ecf_sqnorm_accum(const void* x, int x_dtype, int64_t T, int64_t C, int64_t ld, float* scaler_row,
                                float rescale, float inv_n, void* ws, size_t ws_bytes, ecf_stream_t stream) {
  ecf_sqnorm_desc s;
  s.x = x; s.scaler_row = scaler_row; s.T = T; s.C = C; s.ld = ld; s.dtype = x_dtype; s.rescale = rescale; s.inv_n = inv_n;
  return ecf_sqnorm_accum_batched(&s, 1, ws, ws_bytes, stream);
}
