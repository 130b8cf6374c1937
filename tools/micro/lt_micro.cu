// Micro-benchmark: what can the per-layer select's streaming COUNT pass reach as a stand-alone kernel?
// (#(coarse key < lo) + "has a bracket element" per 8-element fp16 vector; q table from global memory through L1)
// Variants: occupancy via __launch_bounds__, loads in flight per thread.  Build: nvcc -arch=sm_100a -O3 -o lt_micro.bin lt_micro.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint4 ldg_noalloc(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ __half2 h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

__device__ __forceinline__ bool count_vec(const uint4& raw, const float* __restrict__ q8, __half2 pl, __half2 ph, int& pc16) {
  const float4 qa = __ldg(reinterpret_cast<const float4*>(q8)), qb = __ldg(reinterpret_cast<const float4*>(q8 + 4));
  const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h2(w[j]));
    const uint32_t u0 = __float_as_uint(__fmul_rn(fabsf(f.x), q[2 * j])), u1 = __float_as_uint(__fmul_rn(fabsf(f.y), q[2 * j + 1]));
    const uint32_t co = __vminu2(__byte_perm(u0, u1, 0x7632) & 0x7fff7fffu, 0x7bff7bffu);
    const uint32_t ml = __hlt2_mask(h2(co), pl), mh = __hlt2_mask(h2(co), ph);
    x |= ml ^ mh;
    pc16 += __popc(ml);
  }
  return x != 0;
}

template <int DEPTH, int MINB>
__global__ void __launch_bounds__(256, MINB)
    count_kernel(const char* __restrict__ W, uint32_t nvec, uint32_t nvpr, const float* __restrict__ q, uint32_t lo, uint32_t hi,
                 unsigned long long* out, uint32_t* list, unsigned* list_n) {
  const uint32_t gtid = blockIdx.x * 256 + threadIdx.x, gthreads = gridDim.x * 256, lane = threadIdx.x & 31;
  const __half2 pl = h2(lo | (lo << 16)), ph = h2(hi | (hi << 16));
  int pc16 = 0;
  unsigned cands = 0;
  for (uint32_t base = gtid - lane; base < nvec; base += DEPTH * gthreads) {
    uint4 r[DEPTH];
#pragma unroll
    for (int k = 0; k < DEPTH; ++k) {
      const uint32_t v = base + lane + k * gthreads;
      r[k] = v < nvec ? ldg_noalloc(W + (size_t)v * 16) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < DEPTH; ++k) {
      const uint32_t v = base + lane + k * gthreads;
      bool c = false;
      if (v < nvec) c = count_vec(r[k], q + (v % nvpr) * 8, pl, ph, pc16);
      const unsigned bal = __ballot_sync(0xffffffffu, c);
      if (bal) {
        if (c) list[(blockIdx.x * 8 + (threadIdx.x >> 5)) * 512 + ((cands + __popc(bal & ((1u << lane) - 1u))) & 511)] = v;
        cands += __popc(bal);
      }
    }
  }
  __shared__ unsigned s_cnt, s_cand;
  if (threadIdx.x == 0) s_cnt = s_cand = 0;
  __syncthreads();
  const int w = __reduce_add_sync(0xffffffffu, pc16 >> 4);
  if (lane == 0) { atomicAdd(&s_cnt, (unsigned)w); atomicAdd(&s_cand, cands); }
  __syncthreads();
  if (threadIdx.x == 0) { atomicAdd(out, (unsigned long long)s_cnt); atomicAdd(list_n, s_cand); }
}

template <int DEPTH, int MINB>
float run(const char* W, uint32_t nvec, uint32_t nvpr, const float* q, uint32_t lo, uint32_t hi, unsigned long long* out, uint32_t* list,
          unsigned* list_n, char* flush, size_t flush_bytes, int sms) {
  float best = 1e9f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaMemsetAsync(flush, rep, flush_bytes);
    cudaMemsetAsync(out, 0, 8);
    cudaMemsetAsync(list_n, 0, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    count_kernel<DEPTH, MINB><<<sms * MINB, 256>>>(W, nvec, nvpr, q, lo, hi, out, list, list_n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  const uint32_t R = argc > 1 ? (uint32_t)atoi(argv[1]) : 17920, C = 1408;  // 25.2 M fp16 elements = 50 MB (one ViT-g block worth of weights)
  const uint32_t nvpr = C / 8, nvec = R * nvpr;
  std::vector<__half> hw((size_t)R * C);
  srand(1);
  for (auto& x : hw) x = __float2half(((rand() % 20001) - 10000) * 2e-6f);
  std::vector<float> hq(C);
  for (auto& x : hq) x = 0.3f + (rand() % 1000) * 1e-3f;
  char *W, *flush; float* q; unsigned long long* out; uint32_t* list; unsigned* list_n;
  const size_t flush_bytes = 256u << 20;
  cudaMalloc(&W, (size_t)R * C * 2); cudaMalloc(&q, C * 4); cudaMalloc(&out, 8); cudaMalloc(&flush, flush_bytes);
  cudaMalloc(&list, (size_t)148 * 8 * 8 * 512 * 4); cudaMalloc(&list_n, 4);
  cudaMemcpy(W, hw.data(), (size_t)R * C * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(q, hq.data(), C * 4, cudaMemcpyHostToDevice);
  int sms = 148;
  // bracket around the median score: |w| ~ U(0, 0.02) * q ~ 0.8 -> median ~ 0.008; coarse = upper 16 bits of the fp32 pattern
  float lo_f = 0.0078f, hi_f = 0.0080f;
  uint32_t lo = (*(uint32_t*)&lo_f) >> 16, hi = ((*(uint32_t*)&hi_f) >> 16) + 1;
#define RUN(D, B) { float ms = run<D, B>(W, nvec, nvpr, q, lo, hi, out, list, list_n, flush, flush_bytes, sms); \
    unsigned long long c; unsigned ln; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&ln, list_n, 4, cudaMemcpyDeviceToHost); \
    printf("depth %d, %d CTAs/SM of 256: %.1f us  (%.0f GB/s)  count_lt %llu  bracket vectors %u of %u\n", D, B, ms * 1e3, (double)R * C * 2 / ms / 1e6, c, ln, nvec); }
  RUN(1, 4) RUN(2, 4) RUN(4, 4) RUN(2, 6) RUN(1, 8) RUN(2, 8)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
