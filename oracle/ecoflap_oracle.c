/* TEST INFRASTRUCTURE ONLY -- plain-C (OpenMP) restatement of the Wanda part of the hot path, used as the
 * multi-threaded CPU baseline of bench.py and cross-checked against the numpy oracle in tests/.
 * Never linked into, loaded by or called from the product (ecoflap_b200/).
 *
 *   ecf_ref_sqnorm_accum        WrappedGPT.add_batch          LAVIS/lavis/compression/pruners/wanda_pruner.py:71-84
 *   ecf_ref_wanda_row_prune     |W|*sqrt(s), stable row sort, first k -> 0     wanda_pruner.py:260,272-279
 *   ecf_ref_wanda_layer_prune   thres = sort(flatten)[idx]; W[M <= thres] = 0  wanda_pruner.py:541,553-558
 *
 * dtype codes: 0 fp32, 1 fp16, 2 bf16 (16-bit values travel as uint16_t bit patterns).
 * Parity: pinned against tests/golden (generated from the reference) through tests/test_oracle_c.py.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static inline float fp16_to_f32(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu, u;
  if (exp == 0) {
    if (man == 0) {
      u = sign;
    } else { /* subnormal */
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) {
    u = sign | 0x7f800000u | (man << 13);
  } else {
    u = sign | ((exp + 112) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static inline float load_as_f32(const void* p, int dtype, int64_t i) {
  if (dtype == 0) return ((const float*)p)[i];
  if (dtype == 1) return fp16_to_f32(((const uint16_t*)p)[i]);
  return bf16_to_f32(((const uint16_t*)p)[i]);
}

static inline void store_zero(void* p, int dtype, int64_t i) {
  if (dtype == 0) ((float*)p)[i] = 0.0f; else ((uint16_t*)p)[i] = 0;
}

/* scaler_row = scaler_row * n/(n+b) + colsum(x^2)/(n+b);  x is [T, C] row-major */
void ecf_ref_sqnorm_accum(const void* x, int dtype, int64_t T, int64_t C, float* scaler_row, int64_t n_old, int64_t b) {
  const float rescale = (float)((double)n_old / (double)(n_old + b));
  const float n = (float)(n_old + b);
#pragma omp parallel for schedule(static)
  for (int64_t c0 = 0; c0 < C; c0 += 64) {
    const int64_t c1 = c0 + 64 < C ? c0 + 64 : C;
    float acc[64];
    for (int j = 0; j < 64; ++j) acc[j] = 0.0f;
    for (int64_t t = 0; t < T; ++t)
      for (int64_t c = c0; c < c1; ++c) {
        const float v = load_as_f32(x, dtype, t * C + c);
        acc[c - c0] += v * v;
      }
    for (int64_t c = c0; c < c1; ++c) {
      const float nrm = sqrtf(acc[c - c0]); /* torch.norm(p=2) ** 2 */
      scaler_row[c] = scaler_row[c] * rescale + (nrm * nrm) / n;
    }
  }
}

typedef struct { float s; int32_t i; } key_t_;

static int cmp_key(const void* a, const void* b) {
  const key_t_* x = (const key_t_*)a; const key_t_* y = (const key_t_*)b;
  if (x->s < y->s) return -1;
  if (x->s > y->s) return 1;
  return (x->i > y->i) - (x->i < y->i); /* stable: ties -> lower column */
}

/* per row: zero the k entries with the smallest |w|*sqrt(s), ties to the lower column index */
void ecf_ref_wanda_row_prune(void* W, int dtype, int64_t R, int64_t C, const float* scaler_row, int64_t k) {
  if (k <= 0) return;
  if (k > C) k = C;
#pragma omp parallel
  {
    key_t_* keys = (key_t_*)malloc((size_t)C * sizeof(key_t_));
    float* sq = (float*)malloc((size_t)C * sizeof(float));
    for (int64_t c = 0; c < C; ++c) sq[c] = sqrtf(scaler_row[c]);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < R; ++r) {
      for (int64_t c = 0; c < C; ++c) {
        keys[c].s = fabsf(load_as_f32(W, dtype, r * C + c)) * sq[c];
        keys[c].i = (int32_t)c;
      }
      qsort(keys, (size_t)C, sizeof(key_t_), cmp_key);
      for (int64_t j = 0; j < k; ++j) store_zero(W, dtype, r * C + keys[j].i);
    }
    free(keys);
    free(sq);
  }
}

static int cmp_f32(const void* a, const void* b) {
  const float x = *(const float*)a, y = *(const float*)b;
  return (x > y) - (x < y);
}

/* thres = idx-th smallest score of the whole matrix; zero every entry with score <= thres. returns thres */
float ecf_ref_wanda_layer_prune(void* W, int dtype, int64_t R, int64_t C, const float* scaler_row, int64_t idx) {
  const int64_t n = R * C;
  float* m = (float*)malloc((size_t)n * sizeof(float));
  float* sorted = (float*)malloc((size_t)n * sizeof(float));
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < R; ++r)
    for (int64_t c = 0; c < C; ++c) m[r * C + c] = fabsf(load_as_f32(W, dtype, r * C + c)) * sqrtf(scaler_row[c]);
  memcpy(sorted, m, (size_t)n * sizeof(float));
  qsort(sorted, (size_t)n, sizeof(float), cmp_f32);
  const float thres = sorted[idx];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    if (m[i] <= thres) store_zero(W, dtype, i);
  free(m);
  free(sorted);
  return thres;
}


/* host threads used by the OpenMP loops above (torchrun exports OMP_NUM_THREADS=1 to its workers: the benchmark's
 * reference arm sets the count explicitly) */
void ecf_ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ecf_ref_max_threads(void) { return omp_get_max_threads(); }
