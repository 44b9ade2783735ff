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

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace alpro
