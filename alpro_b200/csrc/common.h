// Host-side helpers shared by all translation units of libalpro_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/alpro_b200.h"

namespace alpro {

void set_last_error(const char* fmt, ...);
int num_sms();

#define ALPRO_REQUIRE(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::alpro::set_last_error(__VA_ARGS__);      \
      return ALPRO_EINVAL;                       \
    }                                            \
  } while (0)

// Check the launch; kernels are asynchronous so this reports configuration errors only.
#define ALPRO_CHECK_LAUNCH(name)                                                        \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      ::alpro::set_last_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
      return static_cast<int>(e__);                                                     \
    }                                                                                   \
  } while (0)

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// A training step is ~1000 back-to-back launches of 20-170 us kernels on one stream; the 1-2 us between the end of one
// kernel and the first instruction of the next (profiles/r02o_timeline.md: 2.96 ms idle per 83.7 ms step) plus each
// kernel's own prologue (barrier init, TMEM allocation, tensor-map fetch) are serial time. Every kernel of the library
// is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may become resident while the previous
// kernel of the stream is still draining, run their prologue, and block in pdl_grid_sync() (griddepcontrol.wait) until
// the previous grid has completed and flushed its memory operations. Rules that keep this safe:
//   * every thread executes pdl_grid_sync() before its first global-memory access (read OR write) and before any exit;
//   * griddepcontrol.launch_dependents is issued only AFTER the wait, so at most two generations are ever in flight
//     (the next kernel's CTAs can only be scheduled once this grid has seen its predecessor complete).
// ALPRO_PDL=0 launches without the attribute (the device-side instructions are then no-ops).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
// Wait for the previous grid of the stream (no-op when launched without the attribute), then allow the next one in.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace alpro
