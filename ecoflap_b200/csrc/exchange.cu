// A1 / A8 exchange step across the GPUs of one NVSwitch box (new -- the reference runs on one
// GPU): the packed per-block norm vectors (10-30 k floats) of all ranks are averaged THROUGH PEER MEMORY by one small
// kernel per rank, instead of an NCCL all-reduce whose ~25-40 us of launch + protocol latency per block is the whole
// multi-GPU overhead of the hot path (87 blocks per BLIP-2 pass).
//
// Every rank owns a symmetric staging buffer (mapped into all peers; allocated and exchanged by the host through
// torch's symmetric-memory rendezvous -- plumbing) laid out as  [2][max_floats] fp32 staging halves | flag words.
// A CTA owns a slice of the vector and is independent of the other CTAs:
//   1. copy the slice of the local vector into the staging half of this call's parity, CTA barrier
//   2. store-release (system scope) this call's epoch into flag[cta][my rank] of EVERY peer (remote stores over NVLink)
//   3. spin (bounded) with load-acquire until flag[cta][r] >= epoch for every rank r in the local flags
//   4. read the slice from every peer's staging half (volatile loads: NVLink, not cached), add in RANK ORDER -- all
//      ranks compute bit-identical sums -- scale by 1/world, write the local vector in place
// The epoch lives in device memory (flag[cta][world]) and is advanced by the kernel itself, so a CUDA-graph replay of
// the launch keeps working.  Parity double-buffering is safe because a rank can only reach call e+2 (which rewrites
// the half of call e) after every peer has signalled call e+1, i.e. has finished reading call e.
#include "common.cuh"

namespace ecf {

constexpr int kExMaxWorld = ECF_EXCHANGE_MAX_WORLD;
constexpr int kExThreads = 512;
constexpr int kExMaxCtas = ECF_EXCHANGE_MAX_CTAS;
constexpr int kExFlagStride = kExMaxWorld + 1;  // per CTA: one flag per source rank + the CTA's epoch counter

struct ExParams {
  float* peer[kExMaxWorld];  // staging buffers of all ranks (own included), as mapped into THIS process
  float* local;              // the vector to average, in place
  int64_t n, max_floats;
  int rank, world;
};

__device__ __forceinline__ unsigned* ex_flags(float* staging, int64_t max_floats) {
  return reinterpret_cast<unsigned*>(staging + 2 * max_floats);
}

__global__ void __launch_bounds__(kExThreads) norm_exchange_kernel(const __grid_constant__ ExParams p) {
  const int tid = threadIdx.x, cta = blockIdx.x;
  unsigned* my_flags = ex_flags(p.peer[p.rank], p.max_floats) + cta * kExFlagStride;
  __shared__ unsigned s_epoch;
  if (tid == 0) s_epoch = my_flags[kExMaxWorld] + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  // The slicing depends on the STAGING size only, never on n: a CTA always owns the same index range of the halves and
  // takes part in every call (with an empty slice when n is short), so all CTAs carry the same epoch and the parity
  // argument above holds per CTA when vectors of different lengths share one staging buffer (round-1 advice).
  const int64_t per = ((p.max_floats + gridDim.x - 1) / gridDim.x + 3) & ~int64_t(3);
  const int64_t i0 = min(p.n, (int64_t)cta * per), i1 = min(p.n, i0 + per);
  const int64_t half = (int64_t)(epoch & 1u) * p.max_floats;
  const bool vec = ((reinterpret_cast<uintptr_t>(p.local) & 15) == 0);  // staging halves are 16-byte aligned by contract
  const int64_t v1 = vec ? i0 + ((i1 - i0) & ~int64_t(3)) : i0;       // [i0, v1): float4 path, [v1, i1): scalar tail

  // 1. publish my slice
  float* mine = p.peer[p.rank] + half;
  for (int64_t i = i0 + 4 * tid; i < v1; i += 4 * kExThreads)
    *reinterpret_cast<float4*>(mine + i) = *reinterpret_cast<const float4*>(p.local + i);
  for (int64_t i = v1 + tid; i < i1; i += kExThreads) mine[i] = p.local[i];
  __syncthreads();
  // 2. signal every peer (own flags included: the wait below is uniform).  The flag store is a RELEASE at system scope
  // issued after the CTA barrier: it is cumulative over the slice stores of all the CTA's threads (they happen before it
  // through the barrier), which replaces a __threadfence_system() per thread -- the fence was the larger part of the
  // exchange's ~20 us (round 2).
  if (tid < p.world) {
    unsigned* f = ex_flags(p.peer[tid], p.max_floats) + cta * kExFlagStride + p.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
  // 3. wait for every rank's slice (epochs only grow; wrap-around after 2^32 calls is not handled).  The wait is
  // bounded (~4 s of SM clocks): a peer that died must not leave this GPU spinning forever -- the slice is then
  // poisoned with NaN so that the failure is loud downstream instead of a hang or a silently partial mean.
  __shared__ int s_timeout;
  if (tid == 0) s_timeout = 0;
  __syncthreads();
  if (tid < p.world) {
    const unsigned* f = my_flags + tid;
    const long long t0 = clock64();
    for (;;) {  // ACQUIRE loads at system scope: pair with the peers' release stores
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if (clock64() - t0 > 8000000000ll) {
        s_timeout = 1;
        break;
      }
    }
  }
  __syncthreads();  // the other threads' reads below happen after the acquires through this barrier
  if (s_timeout) {
    for (int64_t i = i0 + tid; i < i1; i += kExThreads) p.local[i] = __int_as_float(0x7fffffff);
    if (tid == 0) my_flags[kExMaxWorld] = epoch;
    return;
  }
  // 4. rank-ordered sum of the peers' slices: all the remote 128-bit loads of a thread are in flight together
  const float inv = 1.0f / (float)p.world;
  for (int64_t i = i0 + 4 * tid; i < v1; i += 4 * kExThreads) {
    float4 v[kExMaxWorld];
#pragma unroll
    for (int r = 0; r < kExMaxWorld; ++r)
      if (r < p.world) v[r] = __ldcv(reinterpret_cast<const float4*>(p.peer[r] + half + i));
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kExMaxWorld; ++r)
      if (r < p.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    *reinterpret_cast<float4*>(p.local + i) = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
  }
  for (int64_t i = v1 + tid; i < i1; i += kExThreads) {
    float s = 0.f;
    for (int r = 0; r < p.world; ++r) s += __ldcv(p.peer[r] + half + i);
    p.local[i] = s * inv;
  }
  if (tid == 0) my_flags[kExMaxWorld] = epoch;
}

}  // namespace ecf

extern "C" size_t ecf_norm_exchange_staging_bytes(int64_t max_floats) {
  using namespace ecf;
  if (max_floats <= 0) return 0;
  return (size_t)(2 * max_floats) * sizeof(float) + (size_t)kExMaxCtas * kExFlagStride * sizeof(unsigned);
}

extern "C" int ecf_norm_exchange_p2p(float* local, int64_t n, void* const* peer_staging, int64_t max_floats, int rank, int world,
                                     ecf_stream_t stream) {
  using namespace ecf;
  int st = check_device();
  if (st != ECF_OK) return st;
  ECF_REQUIRE(local != nullptr && peer_staging != nullptr, ECF_ERR_INVALID, "norm_exchange: null pointer");
  ECF_REQUIRE(world >= 1 && world <= kExMaxWorld && rank >= 0 && rank < world, ECF_ERR_INVALID, "norm_exchange: rank %d / world %d (max %d)",
              rank, world, kExMaxWorld);
  ECF_REQUIRE(n >= 0 && n <= max_floats && max_floats % 4 == 0, ECF_ERR_INVALID, "norm_exchange: n=%lld exceeds the staging halves (%lld floats)",
              (long long)n, (long long)max_floats);
  if (n == 0 || world == 1) return ECF_OK;
  ExParams p;
  for (int r = 0; r < world; ++r) {
    ECF_REQUIRE(peer_staging[r] != nullptr, ECF_ERR_INVALID, "norm_exchange: staging pointer of rank %d is null", r);
    p.peer[r] = reinterpret_cast<float*>(peer_staging[r]);
  }
  p.local = local; p.n = n; p.max_floats = max_floats; p.rank = rank; p.world = world;
  // a fixed CTA count per staging buffer (a function of max_floats alone): flags and epochs are per CTA
  int ctas = (int)((max_floats + 4 * kExThreads - 1) / (4 * kExThreads));
  if (ctas > kExMaxCtas) ctas = kExMaxCtas;
  if (ctas < 1) ctas = 1;
  norm_exchange_kernel<<<ctas, kExThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  ECF_CUDA_OK(cudaGetLastError());
  return ECF_OK;
}
