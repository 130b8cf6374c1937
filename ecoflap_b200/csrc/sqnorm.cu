// A1 -- per-input-channel squared-norm accumulation of [T, C] activation blocks, batched over the Linears of a
// transformer block.
//
// Replaces WrappedGPT.add_batch (LAVIS/lavis/compression/pruners/wanda_pruner.py:71-84; CoOp wanda_pruner.py:
// 159-172; UPop wanda_pruner.py:65-78): the reference materialises an fp32 [C, T] copy per hooked Linear, runs
// torch.norm (sqrt) and squares it again -- three small kernels per hook call, 9 408 hook calls for BLIP-2.  Here
// every X is streamed exactly once with 128-bit loads and accumulated in fp32 registers, and ONE launch serves
// all the hook calls of a block sweep (up to ECF_SQNORM_MAX_BATCH descriptors).
//
// Descriptors that update the SAME accumulator (the 16 calibration batches of a Linear) form one group: the
// reference's sequential  s = s*n/(n+B) + sum/(n+B)  has the closed form  s*prod(r_j) + sum_j w_j*colsum_j  with
// w_j = inv_n_j * prod_{l>j} r_l.  Groups whose call lists are identical (q/k/v, wi_0/wi_1, cross-attention k/v
// hooks see the very same input tensors) are computed ONCE and written to every accumulator of the set.
//
// Work split.  The rows of a group's calls are concatenated; a CTA owns one column tile (TX = 32 16-byte vectors by
// 8 row lanes; wider tiles are an A/B switch) and a row range that may cross call boundaries (each call's partial is scaled by its weight before it is added).  Row ranges are
// sized in BYTES so that every CTA of the launch streams the same amount and the whole grid is one resident wave
// (sms x 4 CTAs): with equal shares of a saturated HBM all CTAs finish together -- no tail, no second wave.
// Each CTA writes one partial row; tickets fold them in two levels (16 rows per chunk, then the chunk sums) in a
// fixed order -- deterministic, and the serial fold after the last CTA is two short steps instead of `splits` loads.
// Tickets reset themselves: the first kSqCounterBytes of the workspace must be zero before the FIRST call only.
// Bound: HBM.  Algorithmic bytes per hook call: T*C*sizeof(x) + 8*C.
#include <cstdlib>

#include "common.cuh"

namespace ecf {

constexpr int kSqThreads = 256;
constexpr int kSqUnroll = 8;   // independent 16-byte loads in flight per thread
constexpr int kSqOcc = 4;      // resident CTAs per SM (__launch_bounds__ below)
constexpr int kSqMaxSeg = ECF_SQNORM_MAX_BATCH;  // hook calls per launch
constexpr int kSqMaxGroup = 32;                  // distinct accumulators per launch
constexpr int kSqMaxOut = 4;                     // accumulators fed by one (shared-input) group
constexpr int kSqChunk = 16;                     // partial rows folded per first-level ticket
constexpr int kSqCounterBytes = 64 * 1024;
constexpr int kSqMaxCols = kSqThreads * 8;       // widest CTA tile in columns
constexpr int64_t kSqMinCtaBytes = 48 * 1024;    // do not split below this (launch of a single small hook call)

// one hook call: a [T, C] activation block (C, dtype come from its group)
struct SqSeg {
  const void* x;
  int64_t T, ld;
  int64_t row_begin;  // first row of this call in the group's concatenated row space
  float weight;       // inv_n of this call times the rescale factors of the later calls on the same accumulator
};
// one set of accumulators fed by the same call list:  out = out * rescale + sum_j weight_j * colsum(x_j^2)
struct SqGroup {
  float* out[kSqMaxOut];
  int n_out;
  int64_t C;
  int dtype, vec;     // vec: 128-bit path usable by every call (C, ld multiples of the vector width, aligned bases)
  int txl;            // log2 of the 16-byte column vectors per CTA row request (5..8)
  float rescale;      // product of the per-call rescale factors
  int nx, splits;     // column tiles, row ranges per tile
  int64_t rows_total, rows_per_cta;
  int seg_begin, seg_end;
  int cta_begin;      // first block index of this group
  int ticket_begin;   // first ticket of this group: per tile 1 + ceil(splits / kSqChunk) counters
  int64_t partial_begin;  // float offset of this group's partial rows (tile-major, split-minor, kCols floats each)
};

struct SqBatch {
  SqSeg seg[kSqMaxSeg];
  SqGroup grp[kSqMaxGroup];
  int n_groups;
};

template <int DT, bool VEC>
__device__ __forceinline__ int sq_accumulate(const SqBatch& b, const SqGroup& g, int tile, int64_t r0, int64_t r1,
                                             float* __restrict__ red) {
  constexpr int V = VEC ? DType<DT>::kVec : 1;
  const int txl = VEC ? g.txl : 5;
  const int TX = 1 << txl, TY = kSqThreads >> txl;
  const int kCols = TX * V;
  const int tid = threadIdx.x;
  const int tx = tid & (TX - 1), ty = tid >> txl;
  const int64_t col0 = (int64_t)tile * kCols + (int64_t)tx * V;
  const bool col_ok = col0 < g.C;  // VEC: C % V == 0, so the whole vector is in range

  float tot[V];
#pragma unroll
  for (int v = 0; v < V; ++v) tot[v] = 0.f;

  // first call touched by [r0, r1): last segment whose row_begin <= r0
  int lo = g.seg_begin, hi = g.seg_end - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (b.seg[mid].row_begin <= r0) lo = mid; else hi = mid - 1;
  }
  int64_t r = r0;
  for (int si = lo; r < r1; ++si) {
    const SqSeg& sg = b.seg[si];
    const int64_t t0 = r - sg.row_begin;
    const int64_t t1 = min(sg.T, r1 - sg.row_begin);
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    if (col_ok) {
      const char* base = reinterpret_cast<const char*>(sg.x) + col0 * DType<DT>::kBytes;
      const int64_t row_bytes = sg.ld * DType<DT>::kBytes;
      const int64_t step = (int64_t)TY * kSqUnroll;
      for (int64_t t = t0 + ty; t < t1; t += step) {
        if constexpr (VEC) {
          uint4 buf[kSqUnroll];
#pragma unroll
          for (int u = 0; u < kSqUnroll; ++u) {
            const int64_t tt = t + (int64_t)u * TY;
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
                float lo2, hi2;
                unpack2<DT>(w[j], lo2, hi2);
                acc[2 * j] = fmaf(lo2, lo2, acc[2 * j]);
                acc[2 * j + 1] = fmaf(hi2, hi2, acc[2 * j + 1]);
              }
            }
          }
        } else {
          float buf[kSqUnroll];
#pragma unroll
          for (int u = 0; u < kSqUnroll; ++u) {
            const int64_t tt = t + (int64_t)u * TY;
            buf[u] = tt < t1 ? load_elem<DT>(sg.x, tt * sg.ld + col0) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < kSqUnroll; ++u) acc[0] = fmaf(buf[u], buf[u], acc[0]);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) tot[v] = fmaf(acc[v], sg.weight, tot[v]);
    r = sg.row_begin + t1;
  }
  // red[(ty * TX + tx) * V + v] = red[ty * kCols + tx * V + v]: row lane ty, column tx * V + v of the tile
  if constexpr (V == 8) {
    reinterpret_cast<float4*>(red)[tid * 2] = make_float4(tot[0], tot[1], tot[2], tot[3]);
    reinterpret_cast<float4*>(red)[tid * 2 + 1] = make_float4(tot[4], tot[5], tot[6], tot[7]);
  } else if constexpr (V == 4) {
    reinterpret_cast<float4*>(red)[tid] = make_float4(tot[0], tot[1], tot[2], tot[3]);
  } else {
    red[tid] = tot[0];
  }
  return txl;
}

// sum of `nrows` partial rows (row i at base + i * stride floats), four columns per thread, fixed order
__device__ __forceinline__ float4 sq_fold4(const float* base, int nrows, int64_t stride, int q) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* p = reinterpret_cast<const float4*>(base) + q;
  const int64_t s4 = stride >> 2;
  for (int r = 0; r < nrows; r += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (r + u < nrows) ? __ldcg(p + (int64_t)(r + u) * s4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
  }
  return a;
}

__global__ void __launch_bounds__(kSqThreads, kSqOcc)
    sqnorm_batched_kernel(const __grid_constant__ SqBatch b, float* __restrict__ partial, unsigned* __restrict__ counters) {
  __shared__ __align__(16) float red[kSqMaxCols];
  __shared__ int s_last;
  // which group / tile / row range is this CTA?  (CTAs: group-major, tile-major, split-minor)
  int gi = 0;
  while (gi + 1 < b.n_groups && (int)blockIdx.x >= b.grp[gi + 1].cta_begin) ++gi;
  const SqGroup& g = b.grp[gi];
  const int local = (int)blockIdx.x - g.cta_begin;
  const int tile = local / g.splits, split = local - tile * g.splits;
  const int64_t r0 = (int64_t)split * g.rows_per_cta;
  const int64_t r1 = min(g.rows_total, r0 + g.rows_per_cta);
  const int tid = threadIdx.x;

  int txl;
  switch (g.dtype * 2 + g.vec) {
    case ECF_F32 * 2 + 1: txl = sq_accumulate<ECF_F32, true>(b, g, tile, r0, r1, red); break;
    case ECF_F32 * 2 + 0: txl = sq_accumulate<ECF_F32, false>(b, g, tile, r0, r1, red); break;
    case ECF_F16 * 2 + 1: txl = sq_accumulate<ECF_F16, true>(b, g, tile, r0, r1, red); break;
    case ECF_F16 * 2 + 0: txl = sq_accumulate<ECF_F16, false>(b, g, tile, r0, r1, red); break;
    case ECF_BF16 * 2 + 1: txl = sq_accumulate<ECF_BF16, true>(b, g, tile, r0, r1, red); break;
    default: txl = sq_accumulate<ECF_BF16, false>(b, g, tile, r0, r1, red); break;
  }
  const int V = g.vec ? (g.dtype == ECF_F32 ? 4 : 8) : 1;
  const int kCols = V << txl, TY = kSqThreads >> txl;
  __syncthreads();
  // cross-row-lane fold, one partial row per CTA
  float* tile_partial = partial + g.partial_begin + (int64_t)tile * g.splits * kCols;
  float* my_partial = tile_partial + (int64_t)split * kCols;
  for (int c = tid; c < kCols; c += kSqThreads) {
    float s = 0.f;
    for (int r = 0; r < TY; ++r) s += red[r * kCols + c];
    my_partial[c] = s;
  }
  // ---- level 1: the last CTA of a chunk of kSqChunk row ranges folds the chunk into its first row --------------
  const int nch = (g.splits + kSqChunk - 1) / kSqChunk;
  unsigned* tile_ticket = counters + g.ticket_begin + tile * (1 + nch);
  const int ch = split / kSqChunk;
  const int ch_rows = min(kSqChunk, g.splits - ch * kSqChunk);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(tile_ticket + 1 + ch, 1u);
    s_last = (prev == (unsigned)ch_rows - 1);
    if (s_last) tile_ticket[1 + ch] = 0;  // self-cleaning for the next call on this workspace
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float* chunk_row = tile_partial + (int64_t)ch * kSqChunk * kCols;
  const int nq = kCols >> 2;
  // nq <= 512: at most two float4 columns per thread
  const int q0 = tid, q1 = tid + kSqThreads;
  float4 sum0 = make_float4(0.f, 0.f, 0.f, 0.f), sum1 = sum0;
  if (q0 < nq) sum0 = sq_fold4(chunk_row, ch_rows, kCols, q0);
  if (q1 < nq) sum1 = sq_fold4(chunk_row, ch_rows, kCols, q1);
  if (nch > 1) {
    // every thread has finished READING the chunk before its first row is overwritten with the chunk sum
    __syncthreads();
    if (q0 < nq) reinterpret_cast<float4*>(chunk_row)[q0] = sum0;
    if (q1 < nq) reinterpret_cast<float4*>(chunk_row)[q1] = sum1;
    // ---- level 2: the last chunk folder of the tile adds the chunk sums in chunk order ----------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(tile_ticket, 1u);
      s_last = (prev == (unsigned)nch - 1);
      if (s_last) tile_ticket[0] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (q0 < nq) sum0 = sq_fold4(tile_partial, nch, (int64_t)kSqChunk * kCols, q0);
    if (q1 < nq) sum1 = sq_fold4(tile_partial, nch, (int64_t)kSqChunk * kCols, q1);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int q = h ? q1 : q0;
    if (q >= nq) continue;
    const float4 sv = h ? sum1 : sum0;
    const float s4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t col = (int64_t)tile * kCols + q * 4 + e;
      if (col < g.C) {
        for (int o = 0; o < g.n_out; ++o) g.out[o][col] = g.out[o][col] * g.rescale + s4[e];
      }
    }
  }
}

static int sq_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v != nullptr && *v ? atoi(v) : dflt;
}

// Fills the launch plan; returns the number of CTAs (negative status on error with the message set).
static int64_t sq_plan(const ecf_sqnorm_desc* descs, int n, SqBatch& b, int64_t& tickets_total, int64_t& partial_floats) {
  // tuning / A-B switches (read once): ECF_SQ_TXL forces log2(vectors per CTA row request), ECF_SQ_OCC the CTAs per SM
  // the grid is sized for, ECF_SQ_DEDUP=0 keeps shared-input groups separate.
  static const int env_txl = sq_env("ECF_SQ_TXL", 0);
  static const int env_occ = sq_env("ECF_SQ_OCC", kSqOcc);
  static const bool dedup = sq_env("ECF_SQ_DEDUP", 1) != 0;
  const int64_t cap = (int64_t)sm_count() * (env_occ < 1 ? 1 : env_occ);

  // ---- group the descriptors by accumulator, keeping call order inside a group ---------------------------------
  int order[kSqMaxSeg], gb[kSqMaxGroup + 1];
  bool used[kSqMaxSeg];
  for (int i = 0; i < n; ++i) used[i] = false;
  int ng = 0, ns = 0;
  for (int i = 0; i < n; ++i) {
    if (used[i]) continue;
    if (ng == kSqMaxGroup) {
      set_error("sqnorm: more than %d distinct accumulators in one launch", kSqMaxGroup);
      return ECF_ERR_INVALID;
    }
    gb[ng] = ns;
    for (int j = i; j < n; ++j) {
      if (descs[j].scaler_row != descs[i].scaler_row) continue;
      if (descs[j].C != descs[i].C || descs[j].dtype != descs[i].dtype) {
        set_error("sqnorm: descriptors %d and %d share an accumulator but differ in C or dtype", i, j);
        return ECF_ERR_INVALID;
      }
      used[j] = true;
      order[ns++] = j;
    }
    ++ng;
  }
  gb[ng] = ns;
  // ---- merge groups fed by the very same call list (shared hook inputs) ------------------------------------------
  int alias_of[kSqMaxGroup];
  for (int a = 0; a < ng; ++a) {
    alias_of[a] = -1;
    if (!dedup) continue;
    for (int c = 0; c < a && alias_of[a] < 0; ++c) {
      if (alias_of[c] >= 0 || gb[c + 1] - gb[c] != gb[a + 1] - gb[a]) continue;
      bool same = true;
      for (int p = 0; p < gb[a + 1] - gb[a] && same; ++p) {
        const ecf_sqnorm_desc& u = descs[order[gb[a] + p]];
        const ecf_sqnorm_desc& w = descs[order[gb[c] + p]];
        same = u.x == w.x && u.T == w.T && u.C == w.C && u.ld == w.ld && u.dtype == w.dtype && u.rescale == w.rescale &&
               u.inv_n == w.inv_n;
      }
      if (same) alias_of[a] = c;
    }
  }
  int slot[kSqMaxGroup];
  int out_g = 0, out_s = 0;
  double total_bytes = 0;
  for (int a = 0; a < ng; ++a) {
    if (alias_of[a] >= 0) {
      SqGroup& g = b.grp[slot[alias_of[a]]];
      if (g.n_out < kSqMaxOut) {
        g.out[g.n_out++] = descs[order[gb[a]]].scaler_row;
        slot[a] = slot[alias_of[a]];
        continue;
      }
    }
    slot[a] = out_g;
    SqGroup& g = b.grp[out_g++];
    const ecf_sqnorm_desc& s0 = descs[order[gb[a]]];
    g.out[0] = s0.scaler_row; g.n_out = 1; g.C = s0.C; g.dtype = s0.dtype;
    g.seg_begin = out_s;
    const int V = s0.dtype == ECF_F32 ? 4 : 8;
    bool vec = (s0.C % V == 0);
    double rescale = 1.0;
    int64_t rows = 0;
    for (int p = gb[a]; p < gb[a + 1]; ++p) {
      const ecf_sqnorm_desc& s = descs[order[p]];
      vec = vec && (s.ld % V == 0) && ((reinterpret_cast<uintptr_t>(s.x) & 15) == 0);
      rescale *= (double)s.rescale;
      SqSeg& sg = b.seg[out_s++];
      sg.x = s.x; sg.T = s.T; sg.ld = s.ld; sg.row_begin = rows;
      rows += s.T;
    }
    g.seg_end = out_s;
    g.rows_total = rows;
    g.vec = vec ? 1 : 0;
    g.rescale = (float)rescale;
    // weight_j = inv_n_j * prod_{l > j} rescale_l   (sequential application, closed form)
    double tail = 1.0;
    for (int p = gb[a + 1] - 1, q = g.seg_end - 1; p >= gb[a]; --p, --q) {
      const ecf_sqnorm_desc& s = descs[order[p]];
      b.seg[q].weight = (float)((double)s.inv_n * tail);
      tail *= (double)s.rescale;
    }
    // column vectors per CTA row request.  Measured on B200 (tools/sq_probe.py, ECF_SQ_TXL sweep): 32 vectors x 8 row
    // lanes (512 contiguous bytes per warp request) streams as fast as whole-row requests on the large launches and
    // is 15-30 % faster on the 40 us T5 block launches (8x fewer partial bytes per CTA), so it is the default.
    int txl = 5;
    if (g.vec && env_txl >= 5 && env_txl <= 8) txl = env_txl;
    g.txl = txl;
    const int64_t cols = (g.vec ? (int64_t)V : 1) << txl;
    g.nx = (int)((s0.C + cols - 1) / cols);
    total_bytes += (double)rows * (double)s0.C * dtype_bytes(s0.dtype);
  }
  b.n_groups = out_g;
  // ---- equal BYTES per CTA, one resident wave ----------------------------------------------------------------------
  double target = total_bytes / (double)cap;
  if (target < (double)kSqMinCtaBytes) target = (double)kSqMinCtaBytes;
  int64_t ctas = 0;
  for (int it = 0; it < 64; ++it) {
    ctas = 0;
    for (int gi = 0; gi < out_g; ++gi) {
      SqGroup& g = b.grp[gi];
      const int64_t gran = (int64_t)(kSqThreads >> g.txl) * kSqUnroll;  // whole unrolled trips of every row lane
      const double tile_row_bytes = (double)g.C * dtype_bytes(g.dtype) / (double)g.nx;
      int64_t rows = (int64_t)(target / tile_row_bytes + 0.5);
      rows = (rows + gran - 1) / gran * gran;
      if (rows < gran) rows = gran;
      g.rows_per_cta = rows;
      g.splits = (int)((g.rows_total + rows - 1) / rows);
      ctas += (int64_t)g.splits * g.nx;
    }
    if (ctas <= cap || target > total_bytes) break;
    target *= 1.03;
  }
  int64_t cta = 0, ticket = 0, pf = 0;
  for (int gi = 0; gi < out_g; ++gi) {
    SqGroup& g = b.grp[gi];
    g.cta_begin = (int)cta;
    g.ticket_begin = (int)ticket;
    g.partial_begin = pf;
    const int64_t cols = (g.vec ? (g.dtype == ECF_F32 ? 4ll : 8ll) : 1ll) << g.txl;
    cta += (int64_t)g.nx * g.splits;
    ticket += (int64_t)g.nx * (1 + (g.splits + kSqChunk - 1) / kSqChunk);
    pf += (int64_t)g.nx * g.splits * cols;
    if (cta >= (1ll << 31)) {
      set_error("sqnorm: grid too large");
      return ECF_ERR_INVALID;
    }
  }
  tickets_total = ticket;
  partial_floats = pf;
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
  int64_t tickets = 0, pf = 0;
  const int64_t ctas = sq_plan(descs, n, b, tickets, pf);
  if (ctas < 0) return 0;
  return (size_t)kSqCounterBytes + (size_t)pf * sizeof(float);
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
  int64_t tickets = 0, pf = 0;
  const int64_t ctas = sq_plan(descs, n, b, tickets, pf);
  if (ctas < 0) return (int)ctas;
  ECF_REQUIRE(tickets * (int64_t)sizeof(unsigned) <= kSqCounterBytes, ECF_ERR_INVALID, "sqnorm: %lld tickets in one launch (max %d)",
              (long long)tickets, kSqCounterBytes / (int)sizeof(unsigned));
  const size_t need = (size_t)kSqCounterBytes + (size_t)pf * sizeof(float);
  ECF_REQUIRE(ws != nullptr && ws_bytes >= need, ECF_ERR_WORKSPACE, "sqnorm: workspace %zu < %zu bytes", ws_bytes, need);
  unsigned* counters = reinterpret_cast<unsigned*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kSqCounterBytes);
  sqnorm_batched_kernel<<<(unsigned)ctas, kSqThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(b, partial, counters);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

extern "C" int ecf_sqnorm_accum(const void* x, int x_dtype, int64_t T, int64_t C, int64_t ld, float* scaler_row,
                                float rescale, float inv_n, void* ws, size_t ws_bytes, ecf_stream_t stream) {
  ecf_sqnorm_desc s;
  s.x = x; s.scaler_row = scaler_row; s.T = T; s.C = C; s.ld = ld; s.dtype = x_dtype; s.rescale = rescale; s.inv_n = inv_n;
  return ecf_sqnorm_accum_batched(&s, 1, ws, ws_bytes, stream);
}
