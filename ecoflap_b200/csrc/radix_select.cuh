// Shared pieces of the multi-pass radix select used by the per-layer Wanda threshold and the SparseGPT
// per-tile threshold: the device-resident select state and the single-CTA histogram scan.
#pragma once

#include "common.cuh"

namespace ecf {

struct LtState {
  uint32_t prefix;          // key bits fixed so far (right aligned)
  uint32_t pad;
  unsigned long long rem;   // 0-based rank still to resolve inside the prefix bucket
};

// one CTA: find the bin holding rank `rem`, extend the prefix, clear the histogram for the next pass
template <int BITS, bool FIRST>
__global__ void __launch_bounds__(1024) lt_scan_kernel(LtState* state, unsigned* hist, unsigned long long rem_init) {
  constexpr int NB = 1 << BITS;
  __shared__ unsigned long long warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned long long rem = FIRST ? rem_init : state->rem;
  const uint32_t old_prefix = FIRST ? 0u : state->prefix;
  // each thread owns NB/1024 consecutive bins (NB = 2048 -> 2, 1024 -> 1)
  constexpr int PER = NB / 1024 > 0 ? NB / 1024 : 1;
  unsigned long long mine[PER];
  unsigned long long sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int b = tid * PER + j;
    mine[j] = b < NB ? hist[b] : 0ull;
    sum += mine[j];
  }
  unsigned long long inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    unsigned long long w = warp_tot[lane];
    unsigned long long winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;  // exclusive
  }
  __syncthreads();
  unsigned long long run = warp_tot[wid] + inc - sum;  // exclusive prefix of this thread's first bin
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int b = tid * PER + j;
    if (b < NB && rem >= run && rem < run + mine[j]) {
      state->prefix = (old_prefix << BITS) | (uint32_t)b;
      state->rem = rem - run;
    }
    run += mine[j];
  }
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int b = tid * PER + j;
    if (b < NB) hist[b] = 0;
  }
}


}  // namespace ecf
