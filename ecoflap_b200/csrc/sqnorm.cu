// A1 -- per-input-channel squared-norm accumulation of a [T, C] activation block.
//
// Replaces WrappedGPT.add_batch (LAVIS/lavis/compression/pruners/wanda_pruner.py:71-84): the
// reference materialises an fp32 [C, T] copy, runs torch.norm (sqrt) and squares it again; here X is
// streamed exactly once with 128-bit loads, squared and accumulated in fp32 registers.
//
// Layout: a CTA is 32 column-vectors (16 bytes each: 512 contiguous bytes per warp request) by 8
// row lanes; grid.x tiles the channels, grid.y splits the tokens so that ~4 CTAs per SM are in
// flight.  Each CTA writes one partial row into the workspace; the last CTA to finish a channel
// tile (atomic ticket) adds the partials in a fixed order, so the result is deterministic, and
// applies   scaler_row = scaler_row * rescale + sum * inv_n.  The tickets reset themselves, so the first
// 4 KB of the workspace must be zero before the FIRST call only (no per-call memset node).
// Bound: HBM.  Algorithmic bytes per call: T*C*sizeof(x) + 8*C.
#include "common.cuh"

namespace ecf {

constexpr int kSqTX = 32;      // column vectors per CTA
constexpr int kSqTY = 8;       // row lanes per CTA
constexpr int kSqUnroll = 8;   // independent 16-byte loads in flight per thread
constexpr int kSqCounterBytes = 4096;

template <int DT, bool VEC>
__global__ void __launch_bounds__(kSqTX* kSqTY)
    sqnorm_kernel(const void* __restrict__ x, int64_t T, int64_t C, int64_t ld, int64_t rows_per_cta,
                  float* __restrict__ partial, int64_t cpad, unsigned* __restrict__ counters,
                  float* __restrict__ scaler_row, float rescale, float inv_n) {
  constexpr int V = VEC ? DType<DT>::kVec : 1;
  constexpr int kCols = kSqTX * V;  // columns per CTA
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * kSqTX + tx;
  const int64_t col0 = (int64_t)blockIdx.x * kCols + (int64_t)tx * V;
  const bool col_ok = col0 < C;  // VEC: C % V == 0, so the whole vector is in range
  const int64_t t0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t t1 = min(T, t0 + rows_per_cta);

  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;

  if (col_ok) {
    const char* base = reinterpret_cast<const char*>(x) + col0 * DType<DT>::kBytes;
    const int64_t row_bytes = ld * DType<DT>::kBytes;
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
          buf[u] = tt < t1 ? load_elem<DT>(x, tt * ld + col0) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kSqUnroll; ++u) acc[0] = fmaf(buf[u], buf[u], acc[0]);
      }
    }
  }

  // cross-row-lane reduction through shared memory (vector stores: conflict free)
  __shared__ __align__(16) float red[kSqTY][kCols];
#pragma unroll
  for (int v = 0; v < V; ++v) red[ty][tx * V + v] = acc[v];
  __syncthreads();
  const int64_t out_col = (int64_t)blockIdx.x * kCols + tid;
  if (tid < kCols) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kSqTY; ++r) s += red[r][tid];
    partial[(int64_t)blockIdx.y * cpad + out_col] = s;
  }

  // ticket: the last CTA of this channel tile folds the partials in split order (deterministic)
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(&counters[blockIdx.x], 1u);
    is_last = (prev == gridDim.y - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid < kCols && out_col < C) {
    float s = 0.f;
    for (unsigned sp = 0; sp < gridDim.y; ++sp) s += __ldcg(&partial[(int64_t)sp * cpad + out_col]);
    scaler_row[out_col] = scaler_row[out_col] * rescale + s * inv_n;
  }
  if (tid == 0) counters[blockIdx.x] = 0;  // self-cleaning for the next call on this workspace
}

struct SqPlan {
  bool vec;
  int V;
  int64_t nx, splits, rows_per_cta, cpad;
};

static SqPlan sq_plan(int dt, int64_t T, int64_t C, bool vec, int sms) {
  SqPlan p;
  p.vec = vec;
  p.V = vec ? (dt == ECF_F32 ? 4 : 8) : 1;
  const int64_t cols = (int64_t)kSqTX * p.V;
  p.nx = (C + cols - 1) / cols;
  p.cpad = p.nx * cols;
  const int64_t min_rows = kSqTY * kSqUnroll;
  int64_t want = ((int64_t)sms * 4 + p.nx - 1) / p.nx;
  int64_t maxs = (T + min_rows - 1) / min_rows;
  int64_t s = want < 1 ? 1 : want;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  int64_t rows = (T + s - 1) / s;
  rows = (rows + kSqTY - 1) / kSqTY * kSqTY;
  if (rows < 1) rows = kSqTY;
  p.rows_per_cta = rows;
  p.splits = (T + rows - 1) / rows;
  if (p.splits < 1) p.splits = 1;
  return p;
}

size_t sqnorm_workspace_bytes(int64_t T, int64_t C) {
  const int sms = sm_count();
  size_t worst = 0;
  for (int dt = 0; dt < 3; ++dt)
    for (int vec = 0; vec < 2; ++vec) {
      SqPlan p = sq_plan(dt, T, C, vec != 0, sms);
      size_t b = (size_t)p.splits * (size_t)p.cpad * sizeof(float);
      if (b > worst) worst = b;
    }
  return kSqCounterBytes + worst;
}

template <int DT>
static int launch_sqnorm(const void* x, int64_t T, int64_t C, int64_t ld, float* scaler_row, float rescale,
                         float inv_n, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const int V = DType<DT>::kVec;
  const bool vec = (C % V == 0) && (ld % V == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  SqPlan p = sq_plan(DT, T, C, vec, sm_count());
  ECF_REQUIRE(p.nx * sizeof(unsigned) <= (size_t)kSqCounterBytes, ECF_ERR_INVALID,
              "sqnorm: C=%lld too large", (long long)C);
  const size_t need = kSqCounterBytes + (size_t)p.splits * p.cpad * sizeof(float);
  ECF_REQUIRE(ws != nullptr && ws_bytes >= need, ECF_ERR_WORKSPACE,
              "sqnorm: workspace %zu < %zu bytes", ws_bytes, need);
  unsigned* counters = reinterpret_cast<unsigned*>(ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kSqCounterBytes);
  dim3 grid((unsigned)p.nx, (unsigned)p.splits), block(kSqTX, kSqTY);
  if (vec)
    sqnorm_kernel<DT, true><<<grid, block, 0, stream>>>(x, T, C, ld, p.rows_per_cta, partial, p.cpad, counters,
                                                        scaler_row, rescale, inv_n);
  else
    sqnorm_kernel<DT, false><<<grid, block, 0, stream>>>(x, T, C, ld, p.rows_per_cta, partial, p.cpad,
                                                         counters, scaler_row, rescale, inv_n);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

}  // namespace ecf

extern "C" int ecf_sqnorm_accum(const void* x, int x_dtype, int64_t T, int64_t C, int64_t ld, float* scaler_row,
                                float rescale, float inv_n, void* ws, size_t ws_bytes, ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(x != nullptr && scaler_row != nullptr, ECF_ERR_INVALID, "sqnorm: null pointer");
  ECF_REQUIRE(T >= 0 && C > 0 && ld >= C, ECF_ERR_INVALID, "sqnorm: bad shape T=%lld C=%lld ld=%lld",
              (long long)T, (long long)C, (long long)ld);
  ECF_REQUIRE(T > 0, ECF_ERR_INVALID, "sqnorm: empty activation block (T == 0)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (x_dtype) {
    case ECF_F32: return launch_sqnorm<ECF_F32>(x, T, C, ld, scaler_row, rescale, inv_n, ws, ws_bytes, s);
    case ECF_F16: return launch_sqnorm<ECF_F16>(x, T, C, ld, scaler_row, rescale, inv_n, ws, ws_bytes, s);
    case ECF_BF16: return launch_sqnorm<ECF_BF16>(x, T, C, ld, scaler_row, rescale, inv_n, ws, ws_bytes, s);
  }
  set_error("sqnorm: unknown dtype %d", x_dtype);
  return ECF_ERR_INVALID;
}
