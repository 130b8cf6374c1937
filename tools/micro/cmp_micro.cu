// Micro-benchmark: issue cost of "count packed 16-bit keys below a pivot" formulations on sm_100a.
// Every thread holds 32 packed key pairs in registers and counts them against a pivot that changes per iteration.
//   V0  HSET2.BF + HADD2            (the round-1/2 per-row select)
//   V1  HFMA2.SAT(x, -BIG, p*BIG) + HADD2   (FMA pipe only)
//   V2  IADD (P|0x8000.. - x) + LOP3 + LEA.HI  (ALU pipe only)
//   V3  half the pairs V1, half V2
//   V4  HSETP2 + 2 predicated IADD
//   V5  key build: HMUL2.BF16 |w|*q    V6 key build: 2 cvt + 2 FMUL + PRMT + VIMNMX (current)
//   V7  apply mask: IADD + PRMT(sign replicate) + LOP3     V8 apply mask: HSET2 mask + LOP3
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cmp_micro.bin cmp_micro.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ __half2 h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t u2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ uint32_t fma_sat_h2(uint32_t x, uint32_t nbig, uint32_t pb) {
  uint32_t d;
  asm("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(nbig), "r"(pb));
  return d;
}
__device__ __forceinline__ uint32_t lea_hi(uint32_t a, uint32_t b, int sh) {  // (a >> (32 - sh)) + b
  return (a >> (32 - sh)) + b;
}

constexpr int NP = 32;

template <int V>
__global__ void __launch_bounds__(256, 2) k(const uint32_t* __restrict__ in, uint32_t* out, int iters, uint32_t p0) {
  uint32_t x[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) x[i] = in[(blockIdx.x * 256 + threadIdx.x) * NP + i];
  uint32_t total = 0;
  uint32_t p = p0;
  for (int it = 0; it < iters; ++it) {
    if constexpr (V == 0) {
      const __half2 pv = h2(p | (p << 16));
      __half2 a[4] = {h2(0), h2(0), h2(0), h2(0)};
#pragma unroll
      for (int i = 0; i < NP; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = __hadd2(a[u], __hlt2(h2(x[i + u]), pv));
      const float2 f = __half22float2(__hadd2(__hadd2(a[0], a[1]), __hadd2(a[2], a[3])));
      total += (uint32_t)(f.x + f.y);
    } else if constexpr (V == 1) {
      const uint32_t pb = p | (p << 16), nb = 0xe800e800u ^ (it & 1);  // -2048 (the real kernel derives BIG from p)
      __half2 a[4] = {h2(0), h2(0), h2(0), h2(0)};
#pragma unroll
      for (int i = 0; i < NP; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = __hadd2(a[u], h2(fma_sat_h2(x[i + u], nb, pb)));
      const float2 f = __half22float2(__hadd2(__hadd2(a[0], a[1]), __hadd2(a[2], a[3])));
      total += (uint32_t)(f.x + f.y);
    } else if constexpr (V == 2) {
      const uint32_t pb = (p | (p << 16)) | 0x80008000u;
      uint32_t a[4] = {0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < NP; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = lea_hi((pb - x[i + u]) & 0x80008000u, a[u], 17);
      const uint32_t s = a[0] + a[1] + a[2] + a[3];
      total += (s & 0xffff) + (s >> 16);
    } else if constexpr (V == 3) {
      const uint32_t pbf = p | (p << 16), nb = 0xe800e800u ^ (it & 1);
      const uint32_t pbi = pbf | 0x80008000u;
      __half2 a[2] = {h2(0), h2(0)};
      uint32_t b[2] = {0, 0};
#pragma unroll
      for (int i = 0; i < NP; i += 4) {
        a[0] = __hadd2(a[0], h2(fma_sat_h2(x[i], nb, pbf)));
        b[0] = lea_hi((pbi - x[i + 1]) & 0x80008000u, b[0], 17);
        a[1] = __hadd2(a[1], h2(fma_sat_h2(x[i + 2], nb, pbf)));
        b[1] = lea_hi((pbi - x[i + 3]) & 0x80008000u, b[1], 17);
      }
      const float2 f = __half22float2(__hadd2(a[0], a[1]));
      const uint32_t s = b[0] + b[1];
      total += (uint32_t)(f.x + f.y) + (s & 0xffff) + (s >> 16);
    } else if constexpr (V == 4) {
      const __half2 pv = h2(p | (p << 16));
      uint32_t a0 = 0, a1 = 0;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        bool lo, hi;
        asm("{ .reg .pred a, b; setp.lt.f16x2 a|b, %2, %3; selp.u32 %0, 1, 0, a; selp.u32 %1, 1, 0, b; }"
            : "=r"(*(uint32_t*)&lo), "=r"(*(uint32_t*)&hi) : "r"(x[i]), "r"(u2(pv)));
        (void)lo; (void)hi;
      }
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        uint32_t l, h;
        asm("{ .reg .pred a, b; setp.lt.f16x2 a|b, %2, %3; selp.u32 %0, 1, 0, a; selp.u32 %1, 1, 0, b; }"
            : "=r"(l), "=r"(h) : "r"(x[i]), "r"(u2(pv)));
        a0 += l; a1 += h;
      }
      total += a0 + a1;
    } else if constexpr (V == 5) {
      // key build in bf16: |w| * q  (x = weights, q derived from p)
      const uint32_t q = 0x3f803f80u + (p & 0x7f);
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        uint32_t d;
        asm("{ .reg .b32 t; and.b32 t, %1, 0x7fff7fff; mul.rn.bf16x2 %0, t, %2; }" : "=r"(d) : "r"(x[i]), "r"(q));
        acc ^= d;
      }
      total += acc;
    } else if constexpr (V == 6) {
      const float q0 = __uint_as_float(0x3f800000u + p), q1 = __uint_as_float(0x3f900000u + p);
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float w0 = __uint_as_float(x[i] << 16), w1 = __uint_as_float(x[i] & 0xffff0000u);
        const uint32_t b0 = __float_as_uint(__fmul_rn(fabsf(w0), q0)), b1 = __float_as_uint(__fmul_rn(fabsf(w1), q1));
        acc ^= __vminu2(__byte_perm(b0, b1, 0x7632), 0x7bff7bffu);
      }
      total += acc;
    } else if constexpr (V == 7) {
      const uint32_t pb = (p | (p << 16)) | 0x80008000u;
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const uint32_t t = pb - x[i];
        const uint32_t m = __byte_perm(t, 0, 0xbb99);  // sign of bytes 1 / 3 replicated
        acc ^= x[i] & ~m;
      }
      total += acc;
    } else if constexpr (V == 9) {
      // negated keys in x: flag = min(max(P + (-x), 0), 1) on 16-bit lanes (VIADDMNMX + VIMNMX + VIADD)
      const uint32_t pb = p | (p << 16);
      uint32_t a[4] = {0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < NP; i += 4)
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = __vadd2(a[u], __vminu2(__vmaxs2(__vadd2(pb, x[i + u]), 0u), 0x00010001u));
      const uint32_t s = __vadd2(__vadd2(a[0], a[1]), __vadd2(a[2], a[3]));
      total += (s & 0xffff) + (s >> 16);
    } else if constexpr (V == 10) {
      // 5 of 8 pairs on the FMA pipe (HFMA2.SAT + HADD2), 3 of 8 on the ALU pipe
      const uint32_t pbf = p | (p << 16), nb = 0xe800e800u ^ (it & 1);
      const uint32_t pbi = pbf | 0x80008000u;
      __half2 a[2] = {h2(0), h2(0)};
      uint32_t b[2] = {0, 0};
#pragma unroll
      for (int i = 0; i < NP; i += 8) {
        a[0] = __hadd2(a[0], h2(fma_sat_h2(x[i], nb, pbf)));
        b[0] = lea_hi((pbi - x[i + 1]) & 0x80008000u, b[0], 17);
        a[1] = __hadd2(a[1], h2(fma_sat_h2(x[i + 2], nb, pbf)));
        a[0] = __hadd2(a[0], h2(fma_sat_h2(x[i + 3], nb, pbf)));
        b[1] = lea_hi((pbi - x[i + 4]) & 0x80008000u, b[1], 17);
        a[1] = __hadd2(a[1], h2(fma_sat_h2(x[i + 5], nb, pbf)));
        b[0] = lea_hi((pbi - x[i + 6]) & 0x80008000u, b[0], 17);
        a[0] = __hadd2(a[0], h2(fma_sat_h2(x[i + 7], nb, pbf)));
      }
      const float2 f = __half22float2(__hadd2(a[0], a[1]));
      const uint32_t s = b[0] + b[1];
      total += (uint32_t)(f.x + f.y) + (s & 0xffff) + (s >> 16);
    } else if constexpr (V == 11) {
      // two pivots at once on the FMA pipe: 2 x (HFMA2.SAT + HADD2)
      const uint32_t pb = p | (p << 16), pb2 = pb + 0x00100010u, nb = 0xe800e800u ^ (it & 1);
      __half2 a[2] = {h2(0), h2(0)}, b[2] = {h2(0), h2(0)};
#pragma unroll
      for (int i = 0; i < NP; i += 2) {
        a[0] = __hadd2(a[0], h2(fma_sat_h2(x[i], nb, pb)));
        b[0] = __hadd2(b[0], h2(fma_sat_h2(x[i], nb, pb2)));
        a[1] = __hadd2(a[1], h2(fma_sat_h2(x[i + 1], nb, pb)));
        b[1] = __hadd2(b[1], h2(fma_sat_h2(x[i + 1], nb, pb2)));
      }
      const float2 f = __half22float2(__hadd2(a[0], a[1])), g = __half22float2(__hadd2(b[0], b[1]));
      total += (uint32_t)(f.x + f.y) + ((uint32_t)(g.x + g.y) << 16);
    } else if constexpr (V == 8) {
      const __half2 pv = h2(p | (p << 16));
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < NP; ++i) acc ^= x[i] & ~__hlt2_mask(h2(x[i]), pv);
      total += acc;
    }
    p = (p * 5 + 1) & 0x3fff;
  }
  out[blockIdx.x * 256 + threadIdx.x] = total;
}

template <int V>
static void run(const char* name, const uint32_t* in, uint32_t* out, int grid, int instr_per_pair_hint) {
  const int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<grid, 256>>>(in, out, 10, 1234);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<V><<<grid, 256>>>(in, out, iters, 1234);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  // cycles per SMSP per warp per iteration: grid = 2 CTAs per SM -> 16 warps per SM -> 4 per SMSP
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * 1965e6;
  const double per_warp_iter = cyc / iters / 4.0;
  printf("%-44s %8.3f ms  %7.1f cycles per warp-iteration (32 pairs) = %5.2f cycles per pair  [%s]\n", name, ms, per_warp_iter,
         per_warp_iter / NP, cudaGetErrorString(cudaGetLastError()));
  (void)instr_per_pair_hint;
}

int main() {
  const int grid = 148 * 2;
  const size_t n = (size_t)grid * 256 * NP;
  std::vector<uint32_t> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = ((uint32_t)(rand() & 0x3fff) << 16) | (uint32_t)(rand() & 0x3fff);
  uint32_t *in, *out;
  cudaMalloc(&in, n * 4); cudaMalloc(&out, (size_t)grid * 256 * 4);
  cudaMemcpy(in, h.data(), n * 4, cudaMemcpyHostToDevice);
  run<0>("V0 HSET2 + HADD2", in, out, grid, 2);
  run<1>("V1 HFMA2.SAT + HADD2", in, out, grid, 2);
  run<2>("V2 IADD + LOP3 + LEA.HI", in, out, grid, 3);
  run<3>("V3 half V1, half V2", in, out, grid, 2);
  run<4>("V4 HSETP2 + SEL + IADD", in, out, grid, 3);
  run<5>("V5 key build: LOP3 + HMUL2.BF16", in, out, grid, 2);
  run<6>("V6 key build: fp32 scores, PRMT, VIMNMX", in, out, grid, 6);
  run<7>("V7 apply mask: IADD + PRMT + LOP3", in, out, grid, 3);
  run<8>("V8 apply mask: HSET2 mask + LOP3", in, out, grid, 2);
  run<9>("V9 VIADDMNMX + VIMNMX + VIADD (16x2)", in, out, grid, 3);
  run<10>("V10 5/8 V1, 3/8 V2", in, out, grid, 2);
  run<11>("V11 two pivots, 2 x (HFMA2.SAT + HADD2)", in, out, grid, 4);
  // exactness of the HFMA2.SAT compare is checked on the host side of the real kernel's tests
  return 0;
}
