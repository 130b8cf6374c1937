// A3+A4+A7 -- fused Wanda score / per-row k-smallest select / in-place apply: C-ABI entry + GENERIC kernel.
//
// Aligned rows (C % 8 == 0, 16-byte aligned) -- every shape of the BASELINE.json configurations -- take the
// coarse-to-fine kernel in row_select_fast.cuh.  The kernel below is the shape-generic path (ragged C,
// unaligned views): a plain 31-round bisection on 32-bit keys, bit-exact but far from the HBM roofline.
//
// Replaces  W_metric = |W| * sqrt(scaler_row);  sort(W_metric, dim=-1, stable=True);
//           indices[:, :k];  scatter_;  W[mask] = 0
// (LAVIS/lavis/compression/pruners/wanda_pruner.py:260,272-279; CoOp wanda_pruner.py:357,379-383;
//  UPop wanda_pruner.py:243,253-260): 12*R*C bytes of temporaries become zero -- a row is read
// once, selected in registers and written once.
//
// One GROUP of G lanes (G = 32 ... 512, a power of two) owns one row; every lane keeps up to NV
// 8-element vectors of the row in registers as order-preserving uint32 keys of the exact fp32 score.
// The k-th smallest key is found by bisection on the key bits: each round is one compare+add per
// element in registers and one group-wide integer reduction (REDUX inside a warp, a named barrier
// across the warps of a group).  Shared-memory histograms are deliberately not used: shared atomics
// retire ~1 element/clk/SM while an HBM-bound select needs ~6 elements/clk/SM.
// Ties at the threshold are resolved by ascending column index with a group-wide prefix scan, which
// reproduces torch.sort(stable=True)[:, :k] exactly.
// Bound: HBM.  Algorithmic bytes per call: 2*R*C*sizeof(w) + 4*C.
#include <cstdlib>

#include "row_select_fast.cuh"

namespace ecf {

constexpr int kRsMaxWarpsPerGroup = 32;

// sum of an int over the G lanes of a row group
template <bool MULTI_WARP>
__device__ __forceinline__ int group_sum(int v, int G, int group, int warp_in_group, int lane, int* slots /*[groups][32]*/,
                                         int parity) {
  int w = warp_sum(v);
  if constexpr (!MULTI_WARP) return w;
  int* s = slots + (parity * 8 + group) * kRsMaxWarpsPerGroup;
  if (lane == 0) s[warp_in_group] = w;
  named_bar_sync(1 + group, G);
  int tot = 0;
  const int nw = G >> 5;
  for (int j = 0; j < nw; ++j) tot += s[j];
  return tot;
}

template <int DT, int NV, bool ALIGNED, bool MULTI_WARP, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
    row_select_kernel(void* __restrict__ W, int64_t R, int64_t C, int64_t ld, const float* __restrict__ scaler_row,
                      int64_t k64, int G, uint8_t* __restrict__ mask_bits, int64_t mask_ld,
                      unsigned long long* __restrict__ n_zero) {
  constexpr int E = 8 * NV;  // elements per lane
  __shared__ int slots[2 * 8 * kRsMaxWarpsPerGroup];
  __shared__ int scan_slots[8 * kRsMaxWarpsPerGroup];

  const int tid = threadIdx.x;
  const int group = tid / G;           // row group inside the CTA (<= 8)
  const int gl = tid - group * G;      // lane inside the group
  const int lane = tid & 31;
  const int warp_in_group = gl >> 5;
  const int rows_per_cta = blockDim.x / G;
  const int64_t row = (int64_t)blockIdx.x * rows_per_cta + group;
  if (row >= R) return;  // whole group leaves together (barriers are per group)
  const int k = (int)k64;

  char* wrow = reinterpret_cast<char*>(W) + row * ld * DType<DT>::kBytes;

  uint32_t key[E];
  uint32_t raw[DT == ECF_F32 ? 1 : 4 * NV];  // 16-bit weights stay packed; fp32 rows are re-read (L2) before the store

  // ---- load the row, compute scores -----------------------------------------------------------
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int64_t c0 = ((int64_t)i * G + gl) * 8;
    float w[8];
    if (ALIGNED) {
      if (c0 < C) {
        if constexpr (DT == ECF_F32) {
          const uint4 a = ldg_stream(wrow + c0 * 4), b = ldg_stream(wrow + c0 * 4 + 16);
          w[0] = __uint_as_float(a.x); w[1] = __uint_as_float(a.y); w[2] = __uint_as_float(a.z); w[3] = __uint_as_float(a.w);
          w[4] = __uint_as_float(b.x); w[5] = __uint_as_float(b.y); w[6] = __uint_as_float(b.z); w[7] = __uint_as_float(b.w);
        } else {
          const uint4 a = ldg_stream(wrow + c0 * 2);
          raw[4 * i + 0] = a.x; raw[4 * i + 1] = a.y; raw[4 * i + 2] = a.z; raw[4 * i + 3] = a.w;
#pragma unroll
          for (int j = 0; j < 4; ++j) unpack2<DT>(raw[4 * i + j], w[2 * j], w[2 * j + 1]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = (c0 + j < C) ? load_elem<DT>(wrow, c0 + j) : 0.f;
      if constexpr (DT != ECF_F32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t lo, hi;
          if constexpr (DT == ECF_BF16) {
            lo = __float_as_uint(w[2 * j]) >> 16; hi = __float_as_uint(w[2 * j + 1]) >> 16;
          } else {
            lo = __half_as_ushort(__float2half_rn(w[2 * j])); hi = __half_as_ushort(__float2half_rn(w[2 * j + 1]));
          }
          raw[4 * i + j] = lo | (hi << 16);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t c = c0 + j;
      if (c < C) {
        const float sq = sqrtf(scaler_row[c]);
        key[8 * i + j] = score_key(wanda_score(w[j], sq));
      } else {
        key[8 * i + j] = 0xffffffffu;  // padding: sorts after every real key, never counted
      }
    }
  }

  // ---- k-th smallest key by bisection on the bits ----------------------------------------------
  uint32_t tkey = 0;
  int cnt_lt = 0;
  int parity = 0;
  const bool select = (k > 0) && (k < C);
  if (select) {
    for (int bit = 30; bit >= 0; --bit) {
      const uint32_t cand = tkey | (1u << bit);
      int c = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) c += (key[e] < cand) ? 1 : 0;
      c = group_sum<MULTI_WARP>(c, G, group, warp_in_group, lane, slots, parity);
      parity ^= 1;
      if (c < k) {
        tkey = cand;
        cnt_lt = c;
      }
    }
  }
  // invariant: #{key < tkey} = cnt_lt < k <= #{key <= tkey}
  int need = k - cnt_lt;  // ties (key == tkey) that must go, lowest column first
  int nties = 0;
  uint32_t tie_off[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) tie_off[i] = 0;
  bool ordered = false;
  if (select) {
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (key[e] == tkey) ? 1 : 0;
    nties = group_sum<MULTI_WARP>(c, G, group, warp_in_group, lane, slots, parity);
    parity ^= 1;
    ordered = need < nties;
    if (ordered) {
      // rank of every tie by column: vectors are interleaved (vector i of lane l = i*G + l)
      int base = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int ci = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) ci += (key[8 * i + j] == tkey) ? 1 : 0;
        // inclusive warp scan
        int inc = ci;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        int excl = inc - ci;
        int total = __shfl_sync(0xffffffffu, inc, 31);
        if constexpr (MULTI_WARP) {
          int* s = scan_slots + group * kRsMaxWarpsPerGroup;
          if (lane == 31) s[warp_in_group] = inc;
          named_bar_sync(1 + group, G);
          const int nw = G >> 5;
          int before = 0, tot = 0;
          for (int j = 0; j < nw; ++j) {
            const int v = s[j];
            if (j < warp_in_group) before += v;
            tot += v;
          }
          named_bar_sync(1 + group, G);  // slots are reused by the next vector
          excl += before;
          total = tot;
        }
        tie_off[i] = base + excl;
        base += total;
      }
    }
  }

  // ---- apply + store ----------------------------------------------------------------------------
  int zeros = 0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int64_t c0 = ((int64_t)i * G + gl) * 8;
    if (c0 >= C) continue;
    uint32_t m = 0;
    int rank = (int)tie_off[i];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t kk = key[8 * i + j];
      bool p;
      if (k >= C)
        p = (c0 + j < C);
      else if (!select)
        p = false;
      else if (kk < tkey)
        p = true;
      else if (kk == tkey) {
        p = (!ordered) || (rank < need);
        ++rank;
      } else
        p = false;
      m |= (p ? 1u : 0u) << j;
    }
    if constexpr (DT == ECF_F32) {
      if (ALIGNED) {
        uint4 a = ldg_v4(wrow + c0 * 4), b = ldg_v4(wrow + c0 * 4 + 16);
        uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (m >> j & 1) v[j] = 0;
          zeros += ((v[j] & 0x7fffffffu) == 0) ? 1 : 0;
        }
        if (m) {
          stg_v4(wrow + c0 * 4, make_uint4(v[0], v[1], v[2], v[3]));
          stg_v4(wrow + c0 * 4 + 16, make_uint4(v[4], v[5], v[6], v[7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (c0 + j < C) {
            const bool p = m >> j & 1;
            if (p) store_zero<DT>(wrow, c0 + j);
            zeros += (p || load_elem<DT>(wrow, c0 + j) == 0.f) ? 1 : 0;
          }
        }
      }
    } else {
      uint32_t v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t x = raw[4 * i + j];
        if (m >> (2 * j) & 1) x &= 0xffff0000u;
        if (m >> (2 * j + 1) & 1) x &= 0x0000ffffu;
        v[j] = x;
        zeros += ((x & 0x00007fffu) == 0 && c0 + 2 * j < C) ? 1 : 0;
        zeros += ((x & 0x7fff0000u) == 0 && c0 + 2 * j + 1 < C) ? 1 : 0;
      }
      if (ALIGNED) {
        if (m) stg_v4(wrow + c0 * 2, make_uint4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < C && (m >> j & 1)) store_zero<DT>(wrow, c0 + j);
      }
    }
    if (mask_bits != nullptr) mask_bits[row * mask_ld + (c0 >> 3)] = (uint8_t)m;
  }
  if (n_zero != nullptr) {
    const int z = warp_sum(zeros);
    if (lane == 0 && z) atomicAdd(n_zero, (unsigned long long)z);
  }
}

template <int DT, int NV, bool ALIGNED>
static int launch_rs(void* W, int64_t R, int64_t C, int64_t ld, const float* s, int64_t k, int G, uint8_t* mask,
                     int64_t mask_ld, unsigned long long* nz, cudaStream_t stream) {
  const int block = G > 256 ? 512 : 256;
  const int rows_per_cta = block / G;
  const int64_t grid = (R + rows_per_cta - 1) / rows_per_cta;
  if (G == 32)
    row_select_kernel<DT, NV, ALIGNED, false, 256><<<(unsigned)grid, block, 0, stream>>>(W, R, C, ld, s, k, G, mask, mask_ld, nz);
  else if (G <= 256)
    row_select_kernel<DT, NV, ALIGNED, true, 256><<<(unsigned)grid, block, 0, stream>>>(W, R, C, ld, s, k, G, mask, mask_ld, nz);
  else
    row_select_kernel<DT, NV, ALIGNED, true, 512><<<(unsigned)grid, block, 0, stream>>>(W, R, C, ld, s, k, G, mask, mask_ld, nz);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

template <int DT, bool ALIGNED>
static int dispatch_nv(int nv, void* W, int64_t R, int64_t C, int64_t ld, const float* s, int64_t k, int G,
                       uint8_t* mask, int64_t mask_ld, unsigned long long* nz, cudaStream_t stream) {
  switch (nv) {
#define ECF_CASE(N) \
  case N: return launch_rs<DT, N, ALIGNED>(W, R, C, ld, s, k, G, mask, mask_ld, nz, stream);
    ECF_CASE(1) ECF_CASE(2) ECF_CASE(3) ECF_CASE(4)
    case 5: case 6: case 7: case 8:
      if constexpr (DT == ECF_F32) break; else {
        switch (nv) { ECF_CASE(5) ECF_CASE(6) ECF_CASE(7) ECF_CASE(8) }
      }
#undef ECF_CASE
  }
  set_error("row_select: unsupported vectors-per-lane %d", nv);
  return ECF_ERR_INVALID;
}

template <int DT>
static int run_row_select(void* W, int64_t R, int64_t C, int64_t ld, const float* s, int64_t k, uint8_t* mask,
                          int64_t mask_ld, unsigned long long* nz, cudaStream_t stream) {
  const int nv_max = (DT == ECF_F32) ? 4 : 8;
  const int64_t nvec = (C + 7) / 8;
  int G = 32;
  while (G < 512 && (nvec + G - 1) / G > nv_max) G <<= 1;
  const int nv = (int)((nvec + G - 1) / G);
  ECF_REQUIRE(nv <= nv_max, ECF_ERR_INVALID, "row_select: C=%lld exceeds the supported row length %d",
              (long long)C, 8 * 512 * nv_max);
  const int V = DType<DT>::kVec;
  const bool aligned = (C % 8 == 0) && (ld % V == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (aligned) return dispatch_nv<DT, true>(nv, W, R, C, ld, s, k, G, mask, mask_ld, nz, stream);
  return dispatch_nv<DT, false>(nv, W, R, C, ld, s, k, G, mask, mask_ld, nz, stream);
}

int row_select_fast_f16(RfBatch&, int, bool, cudaStream_t);
int row_select_fast_bf16(RfBatch&, int, bool, cudaStream_t);
int row_select_fast_f32(RfBatch&, int, bool, cudaStream_t);
int row_select_tma_f16(RfBatch&, int, int, size_t, cudaStream_t);
int row_select_tma_bf16(RfBatch&, int, int, size_t, cudaStream_t);

// tuning / A-B switches (read once): ECF_RS_NVMAX = vectors per lane cap of the fast kernel (1..8),
// ECF_RS_KEEP=0 makes 16-bit rows with > 4 vectors per lane re-read the weights in the apply pass (fewer registers),
// ECF_RS_GENERIC=1 forces the generic kernel.
static int rs_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v != nullptr && *v ? atoi(v) : dflt;
}

}  // namespace ecf

namespace ecf {

static int rs_check(const ecf_row_desc& d, int i) {
  ECF_REQUIRE(d.W != nullptr && d.scaler_row != nullptr, ECF_ERR_INVALID, "row_select: null pointer (matrix %d)", i);
  ECF_REQUIRE(d.R >= 0 && d.C > 0 && d.ld >= d.C, ECF_ERR_INVALID, "row_select: bad shape R=%lld C=%lld ld=%lld (matrix %d)",
              (long long)d.R, (long long)d.C, (long long)d.ld, i);
  ECF_REQUIRE(d.k_per_row >= 0, ECF_ERR_INVALID, "row_select: negative k (matrix %d)", i);
  ECF_REQUIRE(d.mask_bits == nullptr || d.mask_ld >= (d.C + 7) / 8, ECF_ERR_INVALID, "row_select: mask_ld too small (matrix %d)", i);
  ECF_REQUIRE(d.dtype >= 0 && d.dtype <= 2, ECF_ERR_INVALID, "row_select: unknown dtype %d (matrix %d)", d.dtype, i);
  return ECF_OK;
}

// the bulk-copy kernel (row_select_tma.cuh): 16-bit weights, whole 256-column tiles, no mask / zero-count outputs
static bool rs_tma_ok(const ecf_row_desc& d) {
  return d.dtype != ECF_F32 && (d.C == 768 || d.C == 1024 || d.C == 2048 || d.C == 3072 || d.C == 4096 || d.C == 5120) &&
         (d.ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(d.W) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d.scaler_row) & 15) == 0) &&
         d.mask_bits == nullptr && d.n_zero == nullptr;
}

// Second stream + events for running two bulk-copy launches of one call side by side (fork / join on the caller's stream:
// legal under stream capture, where it becomes two parallel branches of the graph).  Created on the first call that is NOT
// being captured (stream creation is not allowed while a capture is open); until then the launches stay in order.
struct RsAux {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false;
};
static RsAux* rs_aux(cudaStream_t s) {
  constexpr int kMaxDev = 32;  // one second stream per device (a process may drive several)
  static RsAux auxes[kMaxDev];
  static bool tried_dev[kMaxDev] = {};
  int devi = 0;
  if (cudaGetDevice(&devi) != cudaSuccess || devi < 0 || devi >= kMaxDev) {
    (void)cudaGetLastError();
    return nullptr;
  }
  RsAux& aux = auxes[devi];
  bool& tried = tried_dev[devi];
  if (!tried) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      (void)cudaGetLastError();
      return nullptr;
    }
    tried = true;
    aux.ok = cudaStreamCreateWithFlags(&aux.stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&aux.fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&aux.join, cudaEventDisableTiming) == cudaSuccess;
    if (!aux.ok) (void)cudaGetLastError();
  }
  return aux.ok ? &aux : nullptr;
}

static bool rs_fast_ok(const ecf_row_desc& d) {
  const int vec = d.dtype == ECF_F32 ? 4 : 8;
  return (d.C % 8 == 0) && (d.ld % vec == 0) && ((reinterpret_cast<uintptr_t>(d.W) & 15) == 0) && d.C <= 32768;
}

}  // namespace ecf

extern "C" int ecf_wanda_row_select_apply_batched(const ecf_row_desc* descs, int n, void* ws, size_t ws_bytes, ecf_stream_t stream) {
  using namespace ecf;
  (void)ws;
  (void)ws_bytes;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(descs != nullptr && n >= 1 && n <= ECF_ROW_MAX_BATCH, ECF_ERR_INVALID, "row_select: batch size %d outside [1, %d]", n,
              ECF_ROW_MAX_BATCH);
  for (int i = 0; i < n; ++i) {
    if ((st = rs_check(descs[i], i)) != ECF_OK) return st;
    for (int j = 0; j < i; ++j)
      ECF_REQUIRE(descs[j].W != descs[i].W, ECF_ERR_INVALID, "row_select: matrices %d and %d are the same tensor", j, i);
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static const int nv_max = [] { int v = rs_env("ECF_RS_NVMAX", 8); return v < 1 ? 1 : (v > 8 ? 8 : v); }();
  static const bool force_generic = rs_env("ECF_RS_GENERIC", 0) != 0;
  static const bool keep = rs_env("ECF_RS_KEEP", 1) != 0;
  static const int prefetch = rs_env("ECF_RS_PREFETCH", 0);
  static const int tma = rs_env("ECF_RS_TMA", 1);  // 0: round-2 kernels only; 1 / 2: bulk-copy kernel with that many stages (C <= 2048)
  static const int corun_short = rs_env("ECF_RS_CORUN_SHORT", 3);  // CTAs per SM of the short-row kernel in a side-by-side pair
  static const int corun_pad_kb = rs_env("ECF_RS_CORUN_PAD_KB", 76);
  static const int corun = rs_env("ECF_RS_CORUN", 1);  // 1: a long-row and a short-row bulk-copy launch of one call share the SMs
  bool done[ECF_ROW_MAX_BATCH];
  for (int i = 0; i < n; ++i) done[i] = descs[i].R == 0;
  RfBatch tma_groups[ECF_ROW_MAX_BATCH];
  int tma_dtype[ECF_ROW_MAX_BATCH];
  int n_tma = 0;
  for (int i = 0; i < n; ++i) {
    if (done[i]) continue;
    const ecf_row_desc& d0 = descs[i];
    if (force_generic || !rs_fast_ok(d0)) {
      // shape-generic kernel, one launch per matrix
      const int64_t k = d0.k_per_row > d0.C ? d0.C : d0.k_per_row;  // sort_res[1][:, :k] clamps like python slicing
      switch (d0.dtype) {
        case ECF_F32: st = run_row_select<ECF_F32>(d0.W, d0.R, d0.C, d0.ld, d0.scaler_row, k, d0.mask_bits, d0.mask_ld, d0.n_zero, s); break;
        case ECF_F16: st = run_row_select<ECF_F16>(d0.W, d0.R, d0.C, d0.ld, d0.scaler_row, k, d0.mask_bits, d0.mask_ld, d0.n_zero, s); break;
        default: st = run_row_select<ECF_BF16>(d0.W, d0.R, d0.C, d0.ld, d0.scaler_row, k, d0.mask_bits, d0.mask_ld, d0.n_zero, s); break;
      }
      if (st != ECF_OK) return st;
      done[i] = true;
      continue;
    }
    // every remaining matrix with this row length and dtype (and the same kernel) joins the launch
    const bool use_tma = tma != 0 && rs_tma_ok(d0);
    RfBatch& tb = use_tma ? tma_groups[n_tma] : tma_groups[ECF_ROW_MAX_BATCH - 1];  // (the last slot doubles as scratch: n_tma < n)
    tb.n = 0;
    tb.C = (int)d0.C;
    tb.prefetch = prefetch;
    for (int j = i; j < n; ++j) {
      const ecf_row_desc& d = descs[j];
      if (done[j] || d.C != d0.C || d.dtype != d0.dtype || !rs_fast_ok(d) || (tma != 0 && rs_tma_ok(d)) != use_tma) continue;
      RfMat& M = tb.m[tb.n++];
      M.W = d.W; M.s = d.scaler_row; M.mask = d.mask_bits; M.n_zero = d.n_zero; M.R = d.R; M.ld = d.ld; M.mask_ld = d.mask_ld;
      M.k = (int)(d.k_per_row > d.C ? d.C : d.k_per_row);
      M.batch_begin = 0;
      done[j] = true;
    }
    if (use_tma) {  // launched below, once every group of the call is known
      tma_dtype[n_tma++] = d0.dtype;
      continue;
    }
    switch (d0.dtype) {
      case ECF_F32: st = row_select_fast_f32(tb, nv_max, keep, s); break;
      case ECF_F16: st = row_select_fast_f16(tb, nv_max, keep, s); break;
      default: st = row_select_fast_bf16(tb, nv_max, keep, s); break;
    }
    if (st != ECF_OK) return st;
  }
  auto launch_tma = [&](int g, int share, size_t pad, cudaStream_t on) {
    return tma_dtype[g] == ECF_F16 ? row_select_tma_f16(tma_groups[g], tma, share, pad, on) : row_select_tma_bf16(tma_groups[g], tma, share, pad, on);
  };
  // A T5 block is one launch of short rows (C = 2048: q, k, v, o, wi) and one of long rows (wo, C = 5120: only 2 048 rows, one
  // per warp -- a latency-bound 17 us on its own).  Side by side the long-row kernel keeps ONE four-warp CTA per SM (its
  // dynamic shared memory is padded so that a second one does not fit next to the others) and the short-row kernel three
  // CTAs: 28 warps per SM in all, and the long rows' latency hides behind the short rows' work.
  RsAux* aux = nullptr;
  if (corun != 0 && n_tma == 2 && (tma_groups[0].C >= 3072) != (tma_groups[1].C >= 3072)) aux = rs_aux(s);
  if (aux != nullptr) {
    const int big = tma_groups[0].C >= 3072 ? 0 : 1;
    ECF_CUDA_OK(cudaEventRecord(aux->fork, s));
    ECF_CUDA_OK(cudaStreamWaitEvent(aux->stream, aux->fork, 0));
    if ((st = launch_tma(big, 1, (size_t)corun_pad_kb * 1024, aux->stream)) != ECF_OK) return st;
    if ((st = launch_tma(1 - big, corun_short, 0, s)) != ECF_OK) return st;
    ECF_CUDA_OK(cudaEventRecord(aux->join, aux->stream));
    ECF_CUDA_OK(cudaStreamWaitEvent(s, aux->join, 0));
  } else {
    for (int g = 0; g < n_tma; ++g)
      if ((st = launch_tma(g, 0, 0, s)) != ECF_OK) return st;
  }
  return ECF_OK;
}

extern "C" int ecf_wanda_row_select_apply(void* W, int w_dtype, int64_t R, int64_t C, int64_t ld,
                                          const float* scaler_row, int64_t k_per_row, uint8_t* mask_bits,
                                          int64_t mask_ld, unsigned long long* n_zero, void* ws, size_t ws_bytes,
                                          ecf_stream_t stream) {
  ecf_row_desc d;
  d.W = W; d.scaler_row = scaler_row; d.R = R; d.C = C; d.ld = ld; d.dtype = w_dtype; d.k_per_row = k_per_row;
  d.mask_bits = mask_bits; d.mask_ld = mask_ld; d.n_zero = n_zero;
  return ecf_wanda_row_select_apply_batched(&d, 1, ws, ws_bytes, stream);
}
