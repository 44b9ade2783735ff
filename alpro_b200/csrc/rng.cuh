// Counter-based dropout randomness: a 32-bit integer mixer (lowbias32) over (seed, element-pair index). Stateless, so
// the backward pass regenerates exactly the masks of the forward pass without storing them. 16-bit resolution per
// element (p = 0.1 -> threshold 6554/65536).
#pragma once
#include <stdint.h>

namespace alpro {

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
// two 16-bit uniforms for the element pair `pair_idx` of the dropout site identified by `seed`
__host__ __device__ __forceinline__ uint32_t rand16x2(uint32_t seed, uint64_t pair_idx) {
  return mix32(seed ^ mix32(static_cast<uint32_t>(pair_idx) + 0x9e3779b9U * static_cast<uint32_t>(pair_idx >> 32)));
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) { return static_cast<uint32_t>(p * 65536.0f + 0.5f); }

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the generator behind
// torch.multinomial on CUDA): counter-based, so a draw is a pure function of (key, counter) and needs no state.
// Known-answer vectors of Random123 are checked in tests/test_gpu_sampler.py.
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0, n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1, n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

}  // namespace alpro
