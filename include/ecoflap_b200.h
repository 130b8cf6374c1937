/*
 * ecoflap_b200 -- C ABI of the B200 (sm_100a) implementation of ECoFLaP's pruning hot path.
 *
 * The reference (ylsung/ECoFLaP) is pure Python/PyTorch and has no FFI boundary of its own; each
 * entry point below replaces one ATen call sequence inside the reference's pruner classes (cited
 * per function, paths relative to the reference root).  A maintainer binds these with ctypes --
 * see INTEGRATION.md for the stubs that replace the bodies of WrappedGPT.add_batch,
 * SparseGPT.add_batch/fasterprune, the score/sort/mask block of *_prune() and
 * LayerSparsity.zo_perturb_parameters.
 *
 * Conventions
 *   - every pointer is a BORROWED DEVICE pointer (e.g. torch tensor.data_ptr()); the library never
 *     allocates or frees caller memory.  Scratch comes from a caller-supplied workspace whose size
 *     is returned by ecf_workspace_bytes().
 *   - every call is asynchronous on the supplied CUDA stream and re-entrant across streams
 *     (one workspace per stream).  The only global state is the thread-local error string.
 *   - return value: 0 = OK, negative = error (text through ecf_last_error()).
 *   - dtype enum: 0 = fp32, 1 = fp16, 2 = bf16.  Sizes are int64_t.  Matrices are row-major with a
 *     leading dimension `ld` counted in ELEMENTS.
 *   - k / kth_index are computed ON THE HOST with the reference's exact Python expression
 *     (int(C * s), int(numel * s)); the kernels never see a sparsity ratio.
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute entry point
 *     returns ECF_ERR_NO_DEVICE.
 */
#ifndef ECOFLAP_B200_H_
#define ECOFLAP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECF_ABI_VERSION 6 /* 2: batched per-row select, n:m select, peer-memory norm exchange; 3: cutoff-path per-layer select, flag offset; 4: global select / apply; 5: ecf_obs_prune takes n:m; 6: first-order accumulate / score sums */

#if defined(__GNUC__)
#define ECF_API __attribute__((visibility("default")))
#else
#define ECF_API
#endif

typedef struct CUstream_st* ecf_stream_t; /* == cudaStream_t */

enum ecf_dtype { ECF_F32 = 0, ECF_F16 = 1, ECF_BF16 = 2 };

enum ecf_status {
  ECF_OK = 0,
  ECF_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, unsupported shape)   */
  ECF_ERR_WORKSPACE = -2, /* workspace too small                                              */
  ECF_ERR_CUDA = -3,      /* a CUDA runtime call failed                                       */
  ECF_ERR_NO_DEVICE = -4, /* no sm_100 device                                                 */
  ECF_ERR_RANGE = -5      /* kth_index >= numel (the reference raises IndexError here)        */
};

enum ecf_op {
  ECF_OP_SQNORM = 0,
  ECF_OP_ROW_SELECT = 1,
  ECF_OP_LAYER_THRESH = 2,
  ECF_OP_GROUP_REDUCE = 3,
  ECF_OP_HESSIAN = 4,
  ECF_OP_OBS = 5,
  ECF_OP_GLOBAL_SELECT = 6 /* R = number of segments (1 = one global threshold, n = one per tensor), C unused */
};

ECF_API int ecf_version(void);
ECF_API const char* ecf_last_error(void);
/* number of SMs of the current device, or a negative ecf_status */
ECF_API int ecf_device_sm_count(void);

/* Scratch bytes needed by `op` for an R x C problem (T x C for SQNORM/HESSIAN; for GROUP_REDUCE
 * R = number of tensors and C = total number of chunks, see ecf_group_reduce_chunk_elems). */
ECF_API size_t ecf_workspace_bytes(int op, int64_t R, int64_t C);

/* A1 -- WrappedGPT.add_batch, LAVIS/lavis/compression/pruners/wanda_pruner.py:71-84
 * (CoOp/trainers/pruners/wanda_pruner.py:159-172, UPop/pruners/wanda_pruner.py:65-78).
 *   scaler_row[c] = scaler_row[c] * rescale + (sum_t x[t,c]^2) * inv_n
 * x is the hook input flattened to [T, C]; the host passes rescale = n/(n+B), inv_n = 1/(n+B).
 * fp32 accumulation, deterministic summation order.  The first 64 KB of `ws` hold self-resetting tickets: zero
 * them once before the first call and do not share this workspace with other ops. */
ECF_API int ecf_sqnorm_accum(const void* x, int x_dtype, int64_t T, int64_t C, int64_t ld,
                     float* scaler_row, float rescale, float inv_n,
                     void* ws, size_t ws_bytes, ecf_stream_t stream);

/* A1, batched -- many hook calls in ONE launch (the reference fires one hook per Linear per calibration batch:
 * wanda_pruner.py:238-253 registers them, :71-84 is the body).  `descs` is a HOST array of n <= ECF_SQNORM_MAX_BATCH
 * descriptors; each is one (hook input, accumulator) pair with the same meaning as the arguments of
 * ecf_sqnorm_accum.  Descriptors may share x (q/k/v, wi_0/wi_1 see the same input).  Descriptors may also share
 * scaler_row (successive calibration batches of one Linear; same C and dtype; at most 32 distinct accumulators per
 * launch): they are applied in array order, i.e. the result equals the sequential calls up to fp32 rounding.
 * Workspace: ecf_sqnorm_batched_workspace_bytes(descs, n); its first 64 KB hold self-resetting tickets (zero
 * them once before the first call; do not share the workspace with other ops or streams). */
#define ECF_SQNORM_MAX_BATCH 256
typedef struct ecf_sqnorm_desc {
  const void* x;      /* [T, C] activations, row-major, leading dimension ld (elements) */
  float* scaler_row;  /* [C] fp32 accumulator, updated in place                         */
  int64_t T, C, ld;
  int32_t dtype;      /* enum ecf_dtype of x */
  float rescale;      /* n / (n + B)  */
  float inv_n;        /* 1 / (n + B)  */
} ecf_sqnorm_desc;
ECF_API size_t ecf_sqnorm_batched_workspace_bytes(const ecf_sqnorm_desc* descs, int n);
ECF_API int ecf_sqnorm_accum_batched(const ecf_sqnorm_desc* descs, int n,
                             void* ws, size_t ws_bytes, ecf_stream_t stream);

/* A3+A4+A7 -- per-ROW Wanda select, wanda_pruner.py:260,272-279 (T5); CoOp wanda_pruner.py:357,
 * 379-383 (CLIP); UPop wanda_pruner.py:243,253-260 (BERT); LLaMA/image_classifiers/prune_utils.py:35-38.
 * For every row: score = fp32(|w|) * sqrtf(scaler_row[c]); the k_per_row smallest scores (ties ->
 * lower column index, i.e. torch.sort(stable=True)[:, :k]) are zeroed IN PLACE.
 *   mask_bits (nullable): packed mask, bit (c & 7) of byte mask_bits[r * mask_ld + (c >> 3)] = pruned
 *   n_zero    (nullable): += number of zero-valued weights after the call (check_sparsity, :139-163) */
ECF_API int ecf_wanda_row_select_apply(void* W, int w_dtype, int64_t R, int64_t C, int64_t ld,
                               const float* scaler_row, int64_t k_per_row,
                               uint8_t* mask_bits, int64_t mask_ld, unsigned long long* n_zero,
                               void* ws, size_t ws_bytes, ecf_stream_t stream);

/* A3+A4+A7, batched -- the per-row select of all the Linears of a block (T5: q, k, v, o, wi_0, wi_1, wo; the
 * reference prunes them one after the other, wanda_pruner.py:255-279) in as few launches as their shapes allow:
 * matrices with the same row length C and dtype share ONE persistent launch (a 2048 x 2048 matrix alone is a single
 * 15 us wave on 148 SMs).  Every matrix is pruned exactly as by ecf_wanda_row_select_apply.  `descs` is a HOST array
 * of n <= ECF_ROW_MAX_BATCH distinct matrices. */
#define ECF_ROW_MAX_BATCH 16
typedef struct ecf_row_desc {
  void* W;                    /* [R, C] weights, row-major, leading dimension ld (elements); pruned in place */
  const float* scaler_row;    /* [C] fp32 */
  int64_t R, C, ld;
  int32_t dtype;              /* enum ecf_dtype of W */
  int64_t k_per_row;          /* int(C * s), host-computed */
  uint8_t* mask_bits;         /* nullable: packed mask, row stride mask_ld bytes */
  int64_t mask_ld;
  unsigned long long* n_zero; /* nullable: += zero-valued weights after the call */
} ecf_row_desc;
ECF_API int ecf_wanda_row_select_apply_batched(const ecf_row_desc* descs, int n,
                                       void* ws, size_t ws_bytes, ecf_stream_t stream);

/* A6 -- n:m structured Wanda select, wanda_pruner.py:265-270 (T5), :546-551 (ViT); the 2:4 / 4:8 modes of the LLaMA
 * CLI (LLaMA/main.py:35,55-58).  In every group of m consecutive columns of a row the n smallest scores are zeroed in
 * place (ties -> lower column; torch.topk leaves the tie order unspecified).  1 <= m <= 32.  A last group shorter
 * than n (or n > m) returns ECF_ERR_RANGE: torch.topk raises there.  n == 0 is a no-op. */
ECF_API int ecf_wanda_nm_select_apply(void* W, int w_dtype, int64_t R, int64_t C, int64_t ld,
                              const float* scaler_row, int n, int m,
                              uint8_t* mask_bits, int64_t mask_ld, unsigned long long* n_zero,
                              ecf_stream_t stream);

/* A3+A5+A7 -- per-LAYER threshold select, wanda_pruner.py:541,553-558 (ViT); UPop :502,512-517;
 * prune_utils.py:28-31.  thres = kth_index-th (0-based) smallest score of the whole matrix; every
 * entry with score <= thres is zeroed in place (>= kth_index+1 entries, more on ties).
 *   thres_out (nullable): device float receiving thres. */
ECF_API int ecf_wanda_layer_thresh_apply(void* W, int w_dtype, int64_t R, int64_t C, int64_t ld,
                                 const float* scaler_row, int64_t kth_index, float* thres_out,
                                 uint8_t* mask_bits, int64_t mask_ld, unsigned long long* n_zero,
                                 void* ws, size_t ws_bytes, ecf_stream_t stream);

/* A3+A5+A7, batched -- the per-layer select of all the Linears of a block (ViT: qkv, proj, fc1, fc2) in one call:
 * 16-bit aligned matrices take the cutoff path (sample, count, apply and fix-up kernels working on per-column magnitude
 * cutoffs, with an exact cluster radix select as its fallback), fp32 / ragged / unaligned matrices ONE cooperative
 * kernel; every matrix gets its own exact threshold exactly as in ecf_wanda_layer_thresh_apply
 * (the reference prunes them one after the other: wanda_pruner.py:536-558).  `descs` is a HOST array of
 * n <= ECF_LAYER_MAX_BATCH distinct matrices.  Workspace: ecf_layer_thresh_batched_workspace_bytes(descs, n),
 * zeroed once before its first use, not shared with other ops or streams. */
#define ECF_LAYER_MAX_BATCH 8
typedef struct ecf_layer_desc {
  void* W;                    /* [R, C] weights, row-major, leading dimension ld (elements); pruned in place */
  const float* scaler_row;    /* [C] fp32 */
  int64_t R, C, ld;
  int32_t dtype;              /* enum ecf_dtype of W */
  int64_t kth_index;          /* int(numel * s), host-computed */
  float* thres_out;           /* nullable: device float receiving the threshold */
  uint8_t* mask_bits;         /* nullable: packed mask, row stride mask_ld bytes */
  int64_t mask_ld;
  unsigned long long* n_zero; /* nullable: += zero-valued weights after the call */
} ecf_layer_desc;
ECF_API size_t ecf_layer_thresh_batched_workspace_bytes(const ecf_layer_desc* descs, int n);
/* Test / profiling aid: byte offset, inside the per-layer select's workspace, of a 32-bit word that is non-zero when the LAST
 * launch on that workspace left the sampled cutoff path for the exact cluster radix select (k-th score outside the sampled
 * bracket, heavy ties, non-finite norms).  Results are exact either way; read the word after synchronising the stream. */
ECF_API size_t ecf_layer_thresh_flag_offset(void);
ECF_API int ecf_wanda_layer_thresh_apply_batched(const ecf_layer_desc* descs, int n,
                                         void* ws, size_t ws_bytes, ecf_stream_t stream);

/* A1 / A8 exchange step on one NVSwitch box (the reference is single-GPU): average a packed fp32
 * vector (the norm accumulators of a block) over the ranks through PEER MEMORY, in place -- a fused copy / signal /
 * gather-sum kernel per rank instead of an NCCL all-reduce.  `peer_staging` is a HOST array of `world` device pointers:
 * rank r's staging buffer as mapped into this process (ecf_norm_exchange_staging_bytes(max_floats) bytes each, zeroed
 * once, symmetric across ranks; the host obtains the mappings from torch's symmetric-memory rendezvous).  Every rank
 * must make the same sequence of calls (same n).  All ranks end with bit-identical results (rank-ordered sums). */
#define ECF_EXCHANGE_MAX_WORLD 16
#define ECF_EXCHANGE_MAX_CTAS 32
ECF_API size_t ecf_norm_exchange_staging_bytes(int64_t max_floats);
ECF_API int ecf_norm_exchange_p2p(float* local, int64_t n, void* const* peer_staging, int64_t max_floats,
                                  int rank, int world, ecf_stream_t stream);

/* A14 -- group aggregation, layer_single_base_pruner.py:361-377 together with the |W| / W^2 factors
 * of :467,:556-559.  One launch over a DEVICE table of tensors; per tensor i:
 *   sum_abs[i] = sum |w|,  sum_sq[i] = sum w^2      (fp32 partials, fp64 final combine)
 * The host multiplies by g-hat and adds per group.  chunk_begin is the running count of
 * ecf_group_reduce_chunk_elems()-sized chunks before tensor i (table[n-1] end = total_chunks). */
typedef struct ecf_tensor_desc {
  const void* ptr;
  int64_t numel;
  int32_t dtype;
  int32_t reserved;
  int64_t chunk_begin;
} ecf_tensor_desc;
ECF_API int64_t ecf_group_reduce_chunk_elems(void);
ECF_API int ecf_group_abs_reduce(const ecf_tensor_desc* d_table, int n_tensors, int64_t total_chunks,
                         double* sum_abs, double* sum_sq,
                         void* ws, size_t ws_bytes, ecf_stream_t stream);

/* N3 -- global-pruner baselines: BLIPT5GlobalPruner.get_mask / get_layerwise_mask, global_pruner.py:116-157, and
 * LayerSparsity.get_mask of the 'Real*' ratio oracle, layer_single_base_pruner.py:156-181.
 *   thr = topk(cat(flatten(score_t)), int(p * numel), largest=False)[-1];   w_t *= (score_t > thr)
 * The scores are recomputed from W (and the accumulated |grad| / grad^2 sums G, n_batches) inside a 3-digit radix select
 * over a DEVICE table of tensors; nothing is materialised.  `segmented` != 0: one select / threshold per tensor
 * (get_layerwise_mask; the per-tensor protection thresholds of get_mask).  d_ranks[seg]: 0-based ascending rank of the
 * wanted score (host: int(p * numel) - 1; numel_t - int(numel_t * (1 - max_sparsity)) for a protection threshold).
 * d_protect (nullable): per-tensor key at or above which a score counts as finfo.max (output of a previous segmented
 * select).  d_tkeys [segments]: order-preserving uint32 key of the threshold (ecf_global_select writes, _apply reads).
 * Pruned weights keep their sign bit (w * 0.0), like `v.data *= mask`.  chunk_begin as in ecf_tensor_desc, with
 * ecf_global_chunk_elems()-sized chunks. */
enum { ECF_GLOBAL_MAG = 0, ECF_GLOBAL_GRAD_MAG_ABS = 1, ECF_GLOBAL_GRAD_MAG_SQ = 2, ECF_GLOBAL_GRAD_ONLY = 3 };
typedef struct ecf_global_desc {
  void* W;
  const float* G; /* fp32 [numel] sum over the batches of |grad| (grad^2 for GRAD_MAG_SQ); unused for ECF_GLOBAL_MAG */
  int64_t numel;
  int32_t dtype;
  int32_t reserved;
  int64_t chunk_begin;
} ecf_global_desc;
ECF_API int64_t ecf_global_chunk_elems(void);
ECF_API int ecf_global_select(const ecf_global_desc* d_table, int n_tensors, int64_t total_chunks, int mode, double n_batches,
                      int segmented, const uint32_t* d_protect, const long long* d_ranks, uint32_t* d_tkeys,
                      void* ws, size_t ws_bytes, ecf_stream_t stream);
ECF_API int ecf_global_apply(const ecf_global_desc* d_table, int n_tensors, int64_t total_chunks, int mode, double n_batches,
                     int segmented, const uint32_t* d_protect, const uint32_t* d_tkeys, unsigned long long* d_n_pruned,
                     ecf_stream_t stream);

/* A13 -- first-order scores, layer_single_base_pruner.py:416-471 (global_pruner.py:256-300).
 * ecf_grad_accum:  G += |g| (square == 0) or g^2, G fp32, g in the parameter dtype -- replaces the per-batch
 *   `gradients_dict[k] += v.cpu().data.float().abs()` (:447-450) with a device-resident accumulator.
 * ecf_global_score_sum:  d_sums[t] += sum over the elements of tensor t of the per-element score (same table and score
 *   modes as ecf_global_select) -- all return_sparsity (:361-370) reads of a first-order score tensor; d_sums: fp64 [n]. */
ECF_API int ecf_grad_accum(float* G, const void* g, int g_dtype, int64_t numel, int square, ecf_stream_t stream);
ECF_API int ecf_global_score_sum(const ecf_global_desc* d_table, int n_tensors, int64_t total_chunks, int mode, double n_batches,
                         double* d_sums, ecf_stream_t stream);

/* A11 -- LayerSparsity.zo_perturb_parameters, layer_single_base_pruner.py:473-486.
 *   w = rn(w + rn(rn(scaling * z) * eps)), each rounding in the parameter dtype (torch evaluates
 * `scaling_factor * z * zo_eps` left to right in w's dtype).  z is drawn by the caller with
 * torch.normal after torch.manual_seed so the RNG stream is the reference's own. */
ECF_API int ecf_zo_perturb(void* W, int w_dtype, int64_t numel, const void* z,
                   double scaling, double eps, ecf_stream_t stream);

/* A17 -- (W == 0).sum(), wanda_pruner.py:154; evaluate_blip.py:432-436.  *n_zero += count. */
ECF_API int ecf_count_zero(const void* W, int w_dtype, int64_t numel, unsigned long long* n_zero,
                   ecf_stream_t stream);

/* A8 -- SparseGPT.add_batch, sparsegpt_pruner.py:71-82 (CoOp sparsegpt_pruner.py:160-171).
 *   H = beta * H + alpha * X^T X          X: [T, C] row-major, H: [C, C] fp32 row-major (ldh)
 * tcgen05 tensor cores, fp32 accumulation in TMEM.  fp16/bf16 products are exact; fp32 inputs are
 * split into bf16 hi/mid terms inside the workspace (error ~2^-16).  The host passes
 * beta = n/(n+B), alpha = 2/(n+B). */
ECF_API int ecf_hessian_accum(const void* x, int x_dtype, int64_t T, int64_t C, int64_t ld,
                      float* H, int64_t ldh, float alpha, float beta,
                      void* ws, size_t ws_bytes, ecf_stream_t stream);

/* A10 -- SparseGPT.fasterprune block loop, sparsegpt_pruner.py:172-213.
 * W: [R, C] fp32 working copy (dead columns already zeroed), Hinv: [C, C] fp32 upper Cholesky
 * factor from the prologue (:96-163, cuSOLVER through torch.linalg).  For each `blocksize`-column
 * block: per-tile threshold at index kth_per_block[b] (= int(R*count*s), host-computed), the
 * in-block sequential OBS sweep, then the trailing update W[:, i2:] -= Err @ Hinv[i1:i2, i2:].
 * prune_n != 0: the n:m branch (:182-198) -- no tile threshold; whenever the sweep reaches a column i with
 * i % prune_m == 0 the prune_n smallest w^2 / diag(Hinv)^2 of columns i .. i+prune_m-1 (weights as updated so far) are
 * pruned; prune_m must be a power of two <= 32; kth_per_block may be NULL. */
ECF_API int ecf_obs_prune(float* W, int64_t R, int64_t C, int64_t ldw,
                  const float* Hinv, int64_t ldh,
                  const int64_t* kth_per_block /*host array, ceil(C/blocksize) entries*/,
                  int blocksize, int prune_n, int prune_m, void* ws, size_t ws_bytes, ecf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ECOFLAP_B200_H_ */
