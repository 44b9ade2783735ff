// Micro-benchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+), and of a GELU-like mix with MUFU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/micro/ffma2_bench.cu -o gpurun_out/ffma2_bench && ./gpurun_out/ffma2_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ex2f(float a) { float d; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }

constexpr int CH = 8;
template <int MODE>   // 0: FFMA x16 per iter, 1: FFMA2 x8 per iter (same flops), 2: FFMA x16 + 2 MUFU, 3: FFMA2 x8 + 2 MUFU
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float v[2 * CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) v[i] = threadIdx.x * 1e-3f + i;
  uint64_t p[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) p[i] = pk(v[2 * i], v[2 * i + 1]);
  const uint64_t A = pk(a, a), B = pk(b, b);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) v[i] = fma1(v[i], a, b);
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i) p[i] = fma2(p[i], A, B);
    }
    if (MODE >= 2) {
      if (MODE == 2) { v[0] = ex2f(v[0]); v[1] = ex2f(v[1]); }
      else { float x, y; upk(p[0], x, y); x = ex2f(x); y = ex2f(y); p[0] = pk(x, y); }
    }
  }
  float s = 0.f;
  if (MODE == 0 || MODE == 2) {
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) s += v[i];
  } else {
#pragma unroll
    for (int i = 0; i < CH; ++i) { float x, y; upk(p[i], x, y); s += x + y; }
  }
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, float* out) {
  const int iters = 20000, grid = 148 * 8;
  k<MODE><<<grid, 256>>>(out, 100, 1.0001f, 1e-7f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 1e-7f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = double(grid) * 256 * iters * 16;
  printf("%-28s %8.3f ms  %7.1f GFMA/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", name, ms, fma / ms * 1e-6, fma / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
  float* out; cudaMalloc(&out, 4);
  run<0>("FFMA  x16", out);
  run<1>("FFMA2 x8", out);
  run<2>("FFMA  x16 + 2 MUFU.EX2", out);
  run<3>("FFMA2 x8  + 2 MUFU.EX2", out);
  return 0;
}
