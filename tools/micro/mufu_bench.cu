// Micro-benchmark: MUFU.EX2 issue rate as a function of the number of warps per SM sub-partition (1, 2, 4), with 32
// independent exponentials in flight per warp — does ONE warp reach the pipe's rate (4 lanes/clk/SMSP = 8 cycles per
// warp instruction)?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/micro/mufu_bench.cu -o /tmp/mufu && /tmp/mufu
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ float ex2f(float a) { float d; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }
template <int MIX>   // 0: MUFU only; 1: FFMA + MUFU + FADD per element (the softmax inner loop)
__global__ void k(float* out, int iters, float a, long long* cyc) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -1e-3f * (threadIdx.x + i);
  float l0 = 0.f, l1 = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (MIX) {
        const float e = ex2f(fmaf(v[i], a, -0.5f));
        if (i & 1) l1 += e; else l0 += e;
        v[i] = v[i] * 0.999f;
      } else {
        v[i] = ex2f(v[i]);
      }
    }
  }
  const long long t1 = clock64();
  float s = l0 + l1;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MIX>
void run(const char* name, int threads, float* out, long long* cyc) {
  const int iters = 2000;
  k<MIX><<<148, threads>>>(out, 10, 1.0001f, cyc);
  k<MIX><<<148, threads>>>(out, iters, 1.0001f, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %2d warps/SMSP: %6.2f cycles per warp-level ex2 (per SMSP: %5.2f)\n", name, threads / 128,
         double(c) / (iters * 32.0), double(c) / (iters * 32.0 * (threads / 128)));
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  for (int t : {128, 256, 512}) run<0>("MUFU.EX2 only", t, out, cyc);
  for (int t : {128, 256, 512}) run<1>("FFMA + MUFU.EX2 + FADD + FMUL", t, out, cyc);
  return 0;
}
