// A6 -- n:m structured Wanda select, fused score / select / apply.
//
// Replaces  for ii in range(0, C, m): tmp = W_metric[:, ii:ii+m];  W_mask.scatter_(1, ii + topk(tmp, n, largest=False)[1])
//           W[W_mask] = 0
// (LAVIS/lavis/compression/pruners/wanda_pruner.py:265-270 and :546-551; the LLaMA CLI's 2:4 / 4:8,
//  LLaMA/main.py:35,55-58).  In every group of m consecutive columns of a row the n smallest scores
// fp32(|w|) * sqrtf(scaler_row) are zeroed in place; ties go to the lower column (torch.topk leaves the tie order
// unspecified; the library-wide rule of the masks is used).  A last group shorter than n makes torch.topk raise: the
// host entry returns ECF_ERR_RANGE for it.
//
// One thread owns one group: m <= 32 scores in registers, rank by m^2 compares (16 for 2:4, 64 for 4:8).  The 2:4 / 4:8
// modes on aligned rows take a column-stationary fast path: sqrt(scaler_row) once per thread, 128-bit loads / stores,
// four rows in flight (an IEEE sqrtf per element would cost more issue slots than the rest of the kernel).
// Bound: HBM.  Algorithmic bytes: 2*R*C*sizeof(w) + 4*C.
#include "common.cuh"

namespace ecf {

constexpr int kNmThreads = 256;
constexpr int kNmMaxM = 32;

template <int DT>
__device__ __forceinline__ void nm_store_zero_or_keep(void* wrow, int64_t c, bool prune) {
  if (prune) store_zero<DT>(wrow, c);
}

// generic path: any 1 <= m <= 32, ragged C, unaligned views
template <int DT>
__global__ void __launch_bounds__(kNmThreads)
    nm_select_kernel(void* __restrict__ W, int64_t R, int64_t C, int64_t ld, const float* __restrict__ s, int n, int m,
                     uint8_t* __restrict__ mask_bits, int64_t mask_ld, unsigned long long* __restrict__ n_zero) {
  const int64_t gpr = (C + m - 1) / m;  // groups per row
  const int64_t total = R * gpr;
  int zeros = 0;
  for (int64_t g = (int64_t)blockIdx.x * kNmThreads + threadIdx.x; g < total; g += (int64_t)gridDim.x * kNmThreads) {
    const int64_t row = g / gpr, c0 = (g - row * gpr) * m;
    const int len = (int)min((int64_t)m, C - c0);
    char* wrow = reinterpret_cast<char*>(W) + row * ld * DType<DT>::kBytes;
    uint32_t key[kNmMaxM];
    float w[kNmMaxM];
#pragma unroll
    for (int j = 0; j < kNmMaxM; ++j) {
      if (j < len) {
        w[j] = load_elem<DT>(wrow, c0 + j);
        key[j] = score_key(wanda_score(w[j], __fadd_rn(sqrtf(s[c0 + j]), 0.f)));
      } else {
        w[j] = 0.f;
        key[j] = 0xffffffffu;
      }
    }
    uint32_t pm = 0;
#pragma unroll
    for (int j = 0; j < kNmMaxM; ++j) {
      if (j < len) {
        int rank = 0;
#pragma unroll
        for (int i = 0; i < kNmMaxM; ++i)
          if (i < len) rank += (key[i] < key[j] || (key[i] == key[j] && i < j)) ? 1 : 0;
        if (rank < n) pm |= 1u << j;
      }
    }
#pragma unroll
    for (int j = 0; j < kNmMaxM; ++j) {
      if (j < len) {
        const bool p = pm >> j & 1;
        if (p) store_zero<DT>(wrow, c0 + j);
        zeros += (p || w[j] == 0.f) ? 1 : 0;
        if (p && mask_bits != nullptr) {
          // byte-granular OR (neighbouring groups may share a mask byte when m < 8)
          const uintptr_t addr = reinterpret_cast<uintptr_t>(mask_bits + row * mask_ld + ((c0 + j) >> 3));
          atomicOr(reinterpret_cast<unsigned*>(addr & ~uintptr_t(3)), 1u << ((addr & 3) * 8 + ((c0 + j) & 7)));
        }
      }
    }
  }
  if (n_zero != nullptr) {
    const int z = warp_sum(zeros);
    if ((threadIdx.x & 31) == 0 && z) atomicAdd(n_zero, (unsigned long long)z);
  }
}

// fast path: aligned rows, C % 8 == 0, m in {4, 8}.  A CTA owns a tile of 256 eight-element vectors (2048 columns) and a
// range of rows: every thread computes sqrt(scaler_row) of ITS eight columns once and keeps it in registers, then
// streams its column down the rows with kNmUnroll independent 128-bit loads in flight.
constexpr int kNmUnroll = 4;

template <int DT>
__device__ __forceinline__ void nm_load_vec(const char* p, uint32_t (&raw)[8]) {
  if constexpr (DT == ECF_F32) {
    const uint4 a = ldg_v4(p), b = ldg_v4(p + 16);
    raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w; raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
  } else {
    const uint4 a = ldg_v4(p);
    raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w;
    raw[4] = raw[5] = raw[6] = raw[7] = 0;
  }
}

template <int DT, int M>
__global__ void __launch_bounds__(kNmThreads)
    nm_select_vec_kernel(void* __restrict__ W, int64_t R, int64_t C, int64_t ld, const float* __restrict__ s, int n,
                         int64_t rows_per_cta, uint8_t* __restrict__ mask_bits, int64_t mask_ld,
                         unsigned long long* __restrict__ n_zero) {
  const int64_t c0 = ((int64_t)blockIdx.x * kNmThreads + threadIdx.x) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  int zeros = 0;
  if (c0 < C) {
    const float4 sa = *reinterpret_cast<const float4*>(s + c0), sb = *reinterpret_cast<const float4*>(s + c0 + 4);
    const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
    float q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = __fadd_rn(sqrtf(sv[j]), 0.f);
    char* col = reinterpret_cast<char*>(W) + c0 * DType<DT>::kBytes;
    const int64_t row_bytes = ld * DType<DT>::kBytes;
    for (int64_t rb = r0; rb < r1; rb += kNmUnroll) {
      uint32_t raw[kNmUnroll][8];
#pragma unroll
      for (int u = 0; u < kNmUnroll; ++u)
        if (rb + u < r1) nm_load_vec<DT>(col + (rb + u) * row_bytes, raw[u]);
#pragma unroll
      for (int u = 0; u < kNmUnroll; ++u) {
        if (rb + u >= r1) break;
        float w[8];
        if constexpr (DT == ECF_F32) {
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = __uint_as_float(raw[u][j]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) unpack2<DT>(raw[u][j], w[2 * j], w[2 * j + 1]);
        }
        uint32_t key[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) key[j] = score_key(wanda_score(w[j], q[j]));
        uint32_t pm = 0;
#pragma unroll
        for (int g0 = 0; g0 < 8; g0 += M) {
#pragma unroll
          for (int j = 0; j < M; ++j) {
            int rank = 0;
#pragma unroll
            for (int i = 0; i < M; ++i) rank += (key[g0 + i] < key[g0 + j] || (key[g0 + i] == key[g0 + j] && i < j)) ? 1 : 0;
            if (rank < n) pm |= 1u << (g0 + j);
          }
        }
        char* p = col + (rb + u) * row_bytes;
        if constexpr (DT == ECF_F32) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (pm >> j & 1) raw[u][j] = 0;
            zeros += (raw[u][j] & 0x7fffffffu) == 0 ? 1 : 0;
          }
          if (pm & 0x0fu) stg_v4(p, make_uint4(raw[u][0], raw[u][1], raw[u][2], raw[u][3]));
          if (pm & 0xf0u) stg_v4(p + 16, make_uint4(raw[u][4], raw[u][5], raw[u][6], raw[u][7]));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t keep = ((pm >> (2 * j) & 1) ? 0u : 0x0000ffffu) | ((pm >> (2 * j + 1) & 1) ? 0u : 0xffff0000u);
            raw[u][j] &= keep;
            zeros += ((raw[u][j] & 0x00007fffu) == 0 ? 1 : 0) + ((raw[u][j] & 0x7fff0000u) == 0 ? 1 : 0);
          }
          if (pm) stg_v4(p, make_uint4(raw[u][0], raw[u][1], raw[u][2], raw[u][3]));
        }
        if (mask_bits != nullptr) mask_bits[(rb + u) * mask_ld + (c0 >> 3)] = (uint8_t)pm;
      }
    }
  }
  if (n_zero != nullptr) {
    const int z = warp_sum(zeros);
    if ((threadIdx.x & 31) == 0 && z) atomicAdd(n_zero, (unsigned long long)z);
  }
}

template <int DT>
static int nm_launch(void* W, int64_t R, int64_t C, int64_t ld, const float* s, int n, int m, uint8_t* mask, int64_t mask_ld,
                     unsigned long long* nz, cudaStream_t stream) {
  const int V = DType<DT>::kVec;
  const bool aligned = (C % 8 == 0) && (ld % V == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(s) & 15) == 0);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (aligned && (m == 4 || m == 8)) {
    const int64_t col_tiles = (C / 8 + kNmThreads - 1) / kNmThreads;
    int64_t row_splits = (cap + col_tiles - 1) / col_tiles;           // ~8 CTAs per SM in total
    int64_t rows_per_cta = (R + row_splits - 1) / row_splits;
    rows_per_cta = (rows_per_cta + kNmUnroll - 1) / kNmUnroll * kNmUnroll;
    if (rows_per_cta < 4 * kNmUnroll) rows_per_cta = 4 * kNmUnroll;
    row_splits = (R + rows_per_cta - 1) / rows_per_cta;
    ECF_REQUIRE(row_splits <= 65535, ECF_ERR_INVALID, "nm_select: too many rows");
    const dim3 grid((unsigned)col_tiles, (unsigned)row_splits);
    if (m == 4)
      nm_select_vec_kernel<DT, 4><<<grid, kNmThreads, 0, stream>>>(W, R, C, ld, s, n, rows_per_cta, mask, mask_ld, nz);
    else
      nm_select_vec_kernel<DT, 8><<<grid, kNmThreads, 0, stream>>>(W, R, C, ld, s, n, rows_per_cta, mask, mask_ld, nz);
  } else {
    const int64_t total = R * ((C + m - 1) / m);
    int64_t grid = (total + kNmThreads - 1) / kNmThreads;
    if (grid > cap) grid = cap;
    nm_select_kernel<DT><<<(unsigned)grid, kNmThreads, 0, stream>>>(W, R, C, ld, s, n, m, mask, mask_ld, nz);
  }
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

}  // namespace ecf

extern "C" int ecf_wanda_nm_select_apply(void* W, int w_dtype, int64_t R, int64_t C, int64_t ld, const float* scaler_row,
                                         int n, int m, uint8_t* mask_bits, int64_t mask_ld, unsigned long long* n_zero,
                                         ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(W != nullptr && scaler_row != nullptr, ECF_ERR_INVALID, "nm_select: null pointer");
  ECF_REQUIRE(R >= 0 && C > 0 && ld >= C, ECF_ERR_INVALID, "nm_select: bad shape R=%lld C=%lld ld=%lld", (long long)R, (long long)C,
              (long long)ld);
  ECF_REQUIRE(m >= 1 && m <= kNmMaxM && n >= 0, ECF_ERR_INVALID, "nm_select: n=%d m=%d outside 0 <= n, 1 <= m <= %d", n, m, kNmMaxM);
  ECF_REQUIRE(mask_bits == nullptr || mask_ld >= (C + 7) / 8, ECF_ERR_INVALID, "nm_select: mask_ld too small");
  // torch.topk(tmp, n) raises when a group holds fewer than n columns (n > m, or a short last group)
  const int64_t last = C % m == 0 ? m : C % m;
  ECF_REQUIRE(n <= m && n <= last, ECF_ERR_RANGE, "nm_select: a group of %lld columns cannot give its %d smallest (torch.topk raises)",
              (long long)(n > m ? m : last), n);
  if (R == 0 || n == 0) return ECF_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (w_dtype) {
    case ECF_F32: return nm_launch<ECF_F32>(W, R, C, ld, scaler_row, n, m, mask_bits, mask_ld, n_zero, s);
    case ECF_F16: return nm_launch<ECF_F16>(W, R, C, ld, scaler_row, n, m, mask_bits, mask_ld, n_zero, s);
    case ECF_BF16: return nm_launch<ECF_BF16>(W, R, C, ld, scaler_row, n, m, mask_bits, mask_ld, n_zero, s);
  }
  set_error("nm_select: unknown dtype %d", w_dtype);
  return ECF_ERR_INVALID;
}
