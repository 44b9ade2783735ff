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

}  // namespace alpro
