// Gradient averaging as OUR kernel over NVLink peer memory (replaces hvd.DistributedOptimizer's allreduce,
// run_video_retrieval.py:320-323,444, and the NCCL kernel of round 1).
//
// Why not NCCL here (profiles/r02c_scaling_probe_n2.md): the bucket reductions run on a side stream UNDER the backward
// GEMMs. An NCCL CTA (512+ threads, ~100 registers, shared memory) cannot share an SM with a 220 KB-smem GEMM CTA, so
// every SM it sits on is missing from the 148-CTA persistent GEMM that launches next: that GEMM's cluster waits for the
// collective to finish and the overlap buys nothing (N=2: exposed 0.5 ms instead of 2.7 ms, but the GEMMs of the step
// got 2.1 ms slower). This kernel is built to CO-RESIDE: 128 threads, no shared memory, < 40 registers, a handful of
// CTAs; a GEMM CTA still fits next to it on the same SM, so it only takes memory-pipe slots.
//
// Data path, two variants, both "two-shot": rank r owns the slice [r*chunk, (r+1)*chunk) of the bucket,
//   NVLS (multicast mapping available):  multimem.ld_reduce.add.v4.f32 pulls the slice of ALL ranks, added inside the
//        NVSwitch; the mean is written to ALL ranks with one multimem.st.v4.f32 — 1/W of the bucket read and written
//        per rank over its own links;
//   P2P  (no multicast): W peer loads per 16 bytes, sum in rank order, W peer stores.
// Every element is reduced by exactly one rank and broadcast, so all ranks end with bit-identical averages.
// Cross-rank ordering (all gradients of the bucket final before anyone reads; all means landed before anyone uses them)
// is the caller's: a signal-pad barrier before and after the launch on the same stream (alpro_b200/comm.py).
#include "common.h"

namespace alpro {
namespace {

constexpr int AR_THREADS = 128;
constexpr int AR_UNROLL = 8;   // 8 x 16 B in flight per thread: the multimem round trip is ~4.5 us (r02d probe)

__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// n4 = number of float4 of this rank's slice, starting at float4 index i0 of the (multicast) buffer
__global__ void __launch_bounds__(AR_THREADS) nvls_allreduce_kernel(float* __restrict__ mc, long long i0, long long n4,
                                                                    float scale) {
  const long long stride = static_cast<long long>(gridDim.x) * AR_THREADS;
  long long i = static_cast<long long>(blockIdx.x) * AR_THREADS + threadIdx.x;
  for (; i + (AR_UNROLL - 1) * stride < n4; i += AR_UNROLL * stride) {
    float4 v[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) v[u] = mc_ld_reduce(mc + 4 * (i0 + i + u * stride));
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) {
      v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
      mc_st(mc + 4 * (i0 + i + u * stride), v[u]);
    }
  }
  for (; i < n4; i += stride) {
    float4 v = mc_ld_reduce(mc + 4 * (i0 + i));
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    mc_st(mc + 4 * (i0 + i), v);
  }
  __threadfence_system();
}

struct PeerPtrs { float* p[16]; };

// peer memory changes under us between launches: a relaxed system-scope load, never the read-only / non-coherent path
__device__ __forceinline__ float4 peer_ld(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(AR_THREADS) p2p_allreduce_kernel(PeerPtrs peers, int world, long long i0, long long n4,
                                                                   float scale) {
  const long long stride = static_cast<long long>(gridDim.x) * AR_THREADS;
  for (long long i = static_cast<long long>(blockIdx.x) * AR_THREADS + threadIdx.x; i < n4; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (r < world) v[r] = peer_ld(reinterpret_cast<const float4*>(peers.p[r]) + i0 + i);   // all loads in flight first
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (r < world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(peers.p[r])[i0 + i] = acc;
  }
  __threadfence_system();
}

// Copy-engine variant (comm.CeGradReducer): the slices of the other ranks arrive in `stage` by cudaMemcpyAsync over
// NVLink (no SM involved); this in-stream kernel finishes the owner's slice: own = (own + sum_k stage[k]) * scale.
__global__ void __launch_bounds__(256) sum_slices_kernel(float* __restrict__ own, const float* __restrict__ stage,
                                                         int nparts, long long n4, long long part_stride4, float scale) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = reinterpret_cast<const float4*>(own)[i];
    for (int k = 0; k < nparts; ++k) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(stage) + k * part_stride4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    reinterpret_cast<float4*>(own)[i] = acc;
  }
}

}  // namespace
}  // namespace alpro

using namespace alpro;

extern "C" int alpro_sum_slices(float* own, const float* stage, int nparts, int64_t count, int64_t part_stride,
                                float scale, void* stream) {
  ALPRO_REQUIRE(own && (stage || nparts == 0) && nparts >= 0 && count > 0 && (count & 3) == 0 && (part_stride & 3) == 0,
                "alpro_sum_slices: bad args (count and part_stride must be multiples of 4 floats)");
  ALPRO_REQUIRE(aligned16(own) && (nparts == 0 || aligned16(stage)), "alpro_sum_slices: alignment");
  const long long n4 = count / 4;
  long long g = cdiv(n4, 256);
  if (g > num_sms() * 4) g = num_sms() * 4;
  sum_slices_kernel<<<static_cast<unsigned>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(own, stage, nparts, n4,
                                                                                             part_stride / 4, scale);
  ALPRO_CHECK_LAUNCH("alpro_sum_slices");
  return 0;
}

// stream-ordered copy between any two device-visible addresses (local or peer-mapped): the driver's copy engines
extern "C" int alpro_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  ALPRO_REQUIRE(dst && src && bytes > 0, "alpro_memcpy_async: bad args");
  const cudaError_t e = cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDeviceToDevice,
                                        static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    set_last_error("alpro_memcpy_async: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

// Averages (scale = 1/world) or sums (scale = 1) buf[offset, offset + count) over `world` ranks in place; this call
// processes the slice owned by `rank`. peer_ptrs: HOST array of `world` device pointers to every rank's buffer base
// (peer-mapped; peer_ptrs[rank] is the local one); mc_ptr: multicast mapping of the same buffer or null.
// offset and count are in floats; offset % 4 == 0 and the bases are 16-byte aligned.
extern "C" int alpro_nvl_allreduce(const void* const* peer_ptrs, void* mc_ptr, int world, int rank, int64_t offset,
                                   int64_t count, float scale, int num_ctas, void* stream) {
  ALPRO_REQUIRE(peer_ptrs && world >= 1 && world <= 16 && rank >= 0 && rank < world && count > 0 && offset >= 0,
                "alpro_nvl_allreduce: bad args");
  ALPRO_REQUIRE((offset & 3) == 0, "alpro_nvl_allreduce: offset must be a multiple of 4 floats");
  for (int r = 0; r < world; ++r)
    ALPRO_REQUIRE(peer_ptrs[r] && aligned16(peer_ptrs[r]), "alpro_nvl_allreduce: peer pointer %d null / unaligned", r);
  // slice of this rank, in float4 units (a ragged tail of < 4 floats cannot occur: GradStore pads regions to 4)
  const long long n4_total = (count + 3) / 4;
  const long long chunk4 = (n4_total + world - 1) / world;
  const long long lo4 = static_cast<long long>(rank) * chunk4;
  const long long n4 = lo4 >= n4_total ? 0 : (lo4 + chunk4 > n4_total ? n4_total - lo4 : chunk4);
  if (n4 <= 0) return 0;
  const long long i0 = offset / 4 + lo4;
  if (num_ctas <= 0) num_ctas = 16;
  long long need = cdiv(n4, AR_THREADS * AR_UNROLL);
  const int grid = static_cast<int>(need < num_ctas ? (need < 1 ? 1 : need) : num_ctas);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mc_ptr) {
    ALPRO_REQUIRE(aligned16(mc_ptr), "alpro_nvl_allreduce: multicast pointer unaligned");
    nvls_allreduce_kernel<<<grid, AR_THREADS, 0, st>>>(static_cast<float*>(mc_ptr), i0, n4, scale);
  } else {
    PeerPtrs pp;
    for (int r = 0; r < 16; ++r) pp.p[r] = r < world ? static_cast<float*>(const_cast<void*>(peer_ptrs[r])) : nullptr;
    p2p_allreduce_kernel<<<grid, AR_THREADS, 0, st>>>(pp, world, i0, n4, scale);
  }
  ALPRO_CHECK_LAUNCH("alpro_nvl_allreduce");
  return 0;
}
