// A8 -- SparseGPT Hessian accumulation  H = beta*H + alpha * X^T X  on tcgen05 tensor cores.
//
// Replaces SparseGPT.add_batch (LAVIS/lavis/compression/pruners/sparsegpt_pruner.py:71-82; CoOp
// sparsegpt_pruner.py:160-171): the reference materialises an fp32 copy of X^T, rescales the whole
// C x C matrix and runs an fp32 SGEMM for every calibration batch.
//
// Mapping.  X is [T, C] row-major, so both operands of X^T X are "MN-major" (the contraction index t is
// the slow one).  TMA loads 64-column x BK-row boxes of X with the 128-byte swizzle; a box is exactly one
// column of MN-major SWIZZLE_128B atoms, so the shared-memory descriptors use LBO = BK*128 B (next
// 64-column atom) and SBO = 1024 B (next 8 rows of t).  One CTA owns a 128 x 256 tile of H:
// tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16) accumulates in fp32 in TMEM (2 x 256 columns,
// double buffered so the epilogue of one tile overlaps the MMAs of the next).  Warp roles: warp 0 TMA
// producer, warp 1 MMA issuer (one elected lane), warp 2 TMEM allocator, warps 4-7 epilogue
// (tcgen05.ld 32x32b, alpha/beta, 128-bit stores).  Only tiles that touch the upper triangle are
// computed; a second small kernel mirrors them.  Small C (few tiles) is split along T and combined
// with vector red.global.add.
// fp16/bf16 products are exact in the fp32 accumulator.  fp32 activations (LayerNorm outputs under
// autocast) are split into bf16 hi + mid terms and accumulated as hi*hi + hi*mid + mid*hi (rel. error
// ~2^-16): plain TF32 truncation would bias the diagonal by ~1e-3.
// Bound: tensor pipe.  Algorithmic flops per call: 2*T*C^2.
#include "common.cuh"
#include "umma.cuh"

namespace ecf {

using namespace umma;

constexpr int kHBM = 128;
constexpr int kHBN = 256;
constexpr int kHStages = 4;
constexpr int kHStageBytes = (kHBM + kHBN) * 64 * 2;  // 48 KB for either (1 term, BK=64) or (2 terms, BK=32)
constexpr int kHThreads = 256;
constexpr int kHSmemBytes = kHStages * kHStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct HessParams {
  float* H;
  int64_t ldh;
  int64_t T;
  int C;
  float alpha, beta;
  int MT, NT, num_tiles, splits;
  int64_t rows_per_split;
  int atomic;   // 1: red.add into a pre-scaled H (split-K); 0: read-modify-write
  int vec_ok;   // H rows 16-byte aligned
};

__device__ __forceinline__ void decode_tile(int t, int MT, int& mi, int& nj) {
  nj = 0;
  for (;; ++nj) {
    const int cnt = min(MT, 2 * nj + 2);
    if (t < cnt) break;
    t -= cnt;
  }
  mi = t;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int FMT, int NTERMS>
__global__ void __launch_bounds__(kHThreads, 1)
    hessian_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1, HessParams p) {
  constexpr int BK = 64 / NTERMS;
  constexpr int kBoxBytes = BK * 128;              // one 64-column x BK-row box
  constexpr int kATermBytes = (kHBM / 64) * kBoxBytes;
  constexpr int kBTermBytes = (kHBN / 64) * kBoxBytes;
  constexpr uint32_t kIdesc = make_idesc(FMT, kMajorMN, kMajorMN, kHBM, kHBN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kHStages * kHStageBytes);
  uint64_t* empty_bar = full_bar + kHStages;
  uint64_t* tmem_full = empty_bar + kHStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap0);
    if (NTERMS == 2) prefetch_tmap(&tmap1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kHStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int total_work = p.num_tiles * p.splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w / p.num_tiles;
        int mi, nj;
        decode_tile(w - split * p.num_tiles, p.MT, mi, nj);
        const int64_t t_begin = (int64_t)split * p.rows_per_split;
        const int64_t t_end = min(p.T, t_begin + p.rows_per_split);
        const int nkb = (int)((t_end - t_begin + BK - 1) / BK);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], kHStageBytes);
          uint8_t* st = smem + stage * kHStageBytes;
          const int32_t row = (int32_t)(t_begin + (int64_t)kb * BK);
#pragma unroll
          for (int term = 0; term < NTERMS; ++term) {
            const CUtensorMap* tm = term == 0 ? &tmap0 : &tmap1;
            uint8_t* a = st + term * kATermBytes;
            uint8_t* b = st + NTERMS * kATermBytes + term * kBTermBytes;
#pragma unroll
            for (int j = 0; j < kHBM / 64; ++j) tma_load_2d(a + j * kBoxBytes, tm, &full_bar[stage], mi * kHBM + 64 * j, row);
#pragma unroll
            for (int j = 0; j < kHBN / 64; ++j) tma_load_2d(b + j * kBoxBytes, tm, &full_bar[stage], nj * kHBN + 64 * j, row);
          }
          if (++stage == kHStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w / p.num_tiles;
        const int64_t t_begin = (int64_t)split * p.rows_per_split;
        const int64_t t_end = min(p.T, t_begin + p.rows_per_split);
        const int nkb = (int)((t_end - t_begin + BK - 1) / BK);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kHBN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * kHStageBytes);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            // one K=16 step = 16 rows of t = 2048 bytes inside every 64-column box
            const uint32_t a0 = st + ks * 2048;
            const uint32_t b0 = st + NTERMS * kATermBytes + ks * 2048;
            if constexpr (NTERMS == 1) {
              mma_f16_ss(d_tmem, make_smem_desc_sw128(a0, kBoxBytes, 1024), make_smem_desc_sw128(b0, kBoxBytes, 1024), kIdesc,
                         (kb | ks) != 0);
            } else {
              const uint64_t a_hi = make_smem_desc_sw128(a0, kBoxBytes, 1024);
              const uint64_t a_mid = make_smem_desc_sw128(a0 + kATermBytes, kBoxBytes, 1024);
              const uint64_t b_hi = make_smem_desc_sw128(b0, kBoxBytes, 1024);
              const uint64_t b_mid = make_smem_desc_sw128(b0 + kBTermBytes, kBoxBytes, 1024);
              mma_f16_ss(d_tmem, a_hi, b_hi, kIdesc, (kb | ks) != 0);
              mma_f16_ss(d_tmem, a_hi, b_mid, kIdesc, 1);
              mma_f16_ss(d_tmem, a_mid, b_hi, kIdesc, 1);
            }
          }
          mma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == kHStages) { stage = 0; phase ^= 1; }
        }
        mma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (TMEM -> registers -> H) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int split = w / p.num_tiles;
      int mi, nj;
      decode_tile(w - split * p.num_tiles, p.MT, mi, nj);
      const int64_t t_begin = (int64_t)split * p.rows_per_split;
      const bool has_work = t_begin < p.T;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int r = mi * kHBM + q * 32 + lane;
      float* hrow = p.H + (int64_t)r * p.ldh;
#pragma unroll 1
      for (int chunk = 0; chunk < kHBN / 32; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kHBN + chunk * 32), v);
        tmem_ld_wait();
        const int c0 = nj * kHBN + chunk * 32;
        if (r < p.C && has_work) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int c = c0 + j;
            if (c >= p.C) break;
            const float d0 = p.alpha * __uint_as_float(v[j]), d1 = p.alpha * __uint_as_float(v[j + 1]);
            const float d2 = p.alpha * __uint_as_float(v[j + 2]), d3 = p.alpha * __uint_as_float(v[j + 3]);
            if (p.vec_ok && c + 3 < p.C) {
              if (p.atomic) {
                red_add_v4(hrow + c, d0, d1, d2, d3);
              } else if (p.beta == 0.f) {
                *reinterpret_cast<float4*>(hrow + c) = make_float4(d0, d1, d2, d3);
              } else {
                float4 h = *reinterpret_cast<const float4*>(hrow + c);
                h.x = fmaf(p.beta, h.x, d0); h.y = fmaf(p.beta, h.y, d1); h.z = fmaf(p.beta, h.z, d2); h.w = fmaf(p.beta, h.w, d3);
                *reinterpret_cast<float4*>(hrow + c) = h;
              }
            } else {
              const float d[4] = {d0, d1, d2, d3};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (c + e < p.C) {
                  if (p.atomic) atomicAdd(hrow + c + e, d[e]);
                  else hrow[c + e] = (p.beta == 0.f ? 0.f : p.beta * hrow[c + e]) + d[e];
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// H[c][r] = H[r][c] for c > r (32x32 tiles through shared memory, coalesced both ways)
__global__ void __launch_bounds__(256) symmetrize_kernel(float* __restrict__ H, int64_t ldh, int C) {
  __shared__ float tile[32][33];
  const int nb = (C + 31) / 32;
  // linear index over pairs (bi <= bj)
  int idx = blockIdx.x, bi = 0;
  for (;; ++bi) {
    const int cnt = nb - bi;
    if (idx < cnt) break;
    idx -= cnt;
  }
  const int bj = bi + idx;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = bi * 32 + i, c = bj * 32 + tx;
    if (r < C && c < C) tile[i][tx] = H[(int64_t)r * ldh + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int r = bj * 32 + i, c = bi * 32 + tx;  // destination (lower part)
    if (r < C && c < C && r > c) H[(int64_t)r * ldh + c] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(256) scale_matrix_kernel(float* __restrict__ H, int64_t ldh, int C, float beta) {
  const int64_t total = (int64_t)C * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C, c = i - r * C;
    float* p = H + r * ldh + c;
    *p = beta == 0.f ? 0.f : *p * beta;
  }
}

// fp32 -> bf16 hi + bf16 mid (x ~= hi + mid, |err| <= 2^-17 |x|)
__global__ void __launch_bounds__(256)
    split_bf16_kernel(const float* __restrict__ x, int64_t T, int C, int64_t ld, __nv_bfloat16* __restrict__ hi,
                      __nv_bfloat16* __restrict__ mid, int64_t ldo) {
  const int64_t total = T * (int64_t)ldo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / ldo;
    const int c = (int)(i - t * ldo);
    float v = c < C ? x[t * ld + c] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float rem = v - __bfloat162float(h);
    hi[i] = h;
    mid[i] = __float2bfloat16_rn(rem);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, int dtype_code, uint64_t inner, uint64_t outer,
                   uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return ECF_ERR_CUDA;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(sym);
  }
  CUtensorMapDataType dt = dtype_code == ECF_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : dtype_code == ECF_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                   : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  (void)elem_bytes;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes);
    return ECF_ERR_CUDA;
  }
  return ECF_OK;
}

static int64_t round8(int64_t c) { return (c + 7) / 8 * 8; }

size_t hessian_workspace_bytes(int64_t T, int64_t C) {
  // T > 0 means "fp32 activations": room for the bf16 hi/mid split.  16-bit inputs need no scratch (pass T = 0).
  return 256 + (size_t)T * (size_t)round8(C) * 2 * sizeof(__nv_bfloat16);
}

template <int FMT, int NTERMS>
static int launch_hessian(const CUtensorMap& m0, const CUtensorMap& m1, const HessParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    ECF_CUDA_OK(cudaFuncSetAttribute(hessian_kernel<FMT, NTERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHSmemBytes));
    attr_set = true;
  }
  const int total = p.num_tiles * p.splits;
  const int grid = total < sm_count() ? total : sm_count();
  hessian_kernel<FMT, NTERMS><<<grid, kHThreads, kHSmemBytes, stream>>>(m0, m1, p);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}

int hessian_run(const void* x, int x_dtype, int64_t T, int64_t C, int64_t ld, float* H, int64_t ldh, float alpha, float beta,
                void* ws, size_t ws_bytes, bool mirror, cudaStream_t stream) {
  ECF_REQUIRE(C % 8 == 0, ECF_ERR_INVALID, "hessian: C=%lld must be a multiple of 8 (TMA row pitch)", (long long)C);
  ECF_REQUIRE(T < (1ll << 31) && C < (1 << 24), ECF_ERR_INVALID, "hessian: shape too large");
  HessParams p;
  p.H = H; p.ldh = ldh; p.T = T; p.C = (int)C; p.alpha = alpha; p.beta = beta;
  p.MT = (int)((C + kHBM - 1) / kHBM);
  p.NT = (int)((C + kHBN - 1) / kHBN);
  p.num_tiles = 0;
  for (int nj = 0; nj < p.NT; ++nj) p.num_tiles += (p.MT < 2 * nj + 2 ? p.MT : 2 * nj + 2);
  const int nterms = x_dtype == ECF_F32 ? 2 : 1;
  const int BK = 64 / nterms;
  const int sms = sm_count();
  int splits = 1;
  if (p.num_tiles < sms) {
    splits = (sms + p.num_tiles - 1) / p.num_tiles;
    const int64_t max_splits = (T + 4 * BK - 1) / (4 * BK);  // at least 4 k-blocks per split
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
  }
  int64_t rows = (T + splits - 1) / splits;
  rows = (rows + BK - 1) / BK * BK;
  p.rows_per_split = rows;
  p.splits = (int)((T + rows - 1) / rows);
  p.atomic = p.splits > 1;
  p.vec_ok = (ldh % 4 == 0) && ((reinterpret_cast<uintptr_t>(H) & 15) == 0);
  if (p.atomic) {
    if (beta != 1.f) {
      scale_matrix_kernel<<<sms * 4, 256, 0, stream>>>(H, ldh, (int)C, beta);
      ECF_CUDA_OK(cudaGetLastError());
    }
  }
  CUtensorMap m0, m1;
  int st;
  if (x_dtype == ECF_F32) {
    const int64_t ldo = round8(C);
    const size_t need = hessian_workspace_bytes(T, C);
    ECF_REQUIRE(ws != nullptr && ws_bytes >= need, ECF_ERR_WORKSPACE, "hessian: fp32 input needs %zu workspace bytes, got %zu",
                need, ws_bytes);
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(ws) + 256);
    __nv_bfloat16* mid = hi + T * ldo;
    split_bf16_kernel<<<sms * 8, 256, 0, stream>>>(reinterpret_cast<const float*>(x), T, (int)C, ld, hi, mid, ldo);
    ECF_CUDA_OK(cudaGetLastError());
    if ((st = encode_tmap_2d(&m0, hi, 2, ECF_BF16, C, T, ldo * 2, 64, BK)) != ECF_OK) return st;
    if ((st = encode_tmap_2d(&m1, mid, 2, ECF_BF16, C, T, ldo * 2, 64, BK)) != ECF_OK) return st;
    if ((st = launch_hessian<kFmtBF16, 2>(m0, m1, p, stream)) != ECF_OK) return st;
  } else {
    ECF_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, ECF_ERR_INVALID,
                "hessian: activation rows must be 16-byte aligned (ld %% 8 == 0)");
    if ((st = encode_tmap_2d(&m0, x, 2, x_dtype, C, T, ld * 2, 64, BK)) != ECF_OK) return st;
    m1 = m0;
    if (x_dtype == ECF_F16) st = launch_hessian<kFmtF16, 1>(m0, m1, p, stream);
    else st = launch_hessian<kFmtBF16, 1>(m0, m1, p, stream);
    if (st != ECF_OK) return st;
  }
  if (mirror) {
    const int nb = (int)((C + 31) / 32);
    symmetrize_kernel<<<nb * (nb + 1) / 2, 256, 0, stream>>>(H, ldh, (int)C);
    ECF_CUDA_OK(cudaGetLastError());
  }
  return ECF_OK;
}

}  // namespace ecf

extern "C" int ecf_hessian_accum(const void* x, int x_dtype, int64_t T, int64_t C, int64_t ld, float* H, int64_t ldh,
                                 float alpha, float beta, void* ws, size_t ws_bytes, ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(x != nullptr && H != nullptr, ECF_ERR_INVALID, "hessian: null pointer");
  ECF_REQUIRE(T > 0 && C > 0 && ld >= C && ldh >= C, ECF_ERR_INVALID, "hessian: bad shape T=%lld C=%lld ld=%lld ldh=%lld",
              (long long)T, (long long)C, (long long)ld, (long long)ldh);
  ECF_REQUIRE(x_dtype >= 0 && x_dtype <= 2, ECF_ERR_INVALID, "hessian: unknown dtype %d", x_dtype);
  return hessian_run(x, x_dtype, T, C, ld, H, ldh, alpha, beta, ws, ws_bytes, /*mirror=*/true,
                     reinterpret_cast<cudaStream_t>(stream));
}
