// Fused optimizer step over the flat parameter / gradient buffers: global-norm gradient clipping
// (torch.nn.utils.clip_grad_norm_, run_video_retrieval.py:473-476) + AdamW (src/optimization/adamw.py:40-103) in two
// launches instead of a ~300-tensor Python loop with ~10 elementwise passes.
#include "common.h"
#include "ptx.cuh"

namespace alpro {
namespace {

__global__ void sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ float red[32];
  float s = 0.f;
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += x[i] * x[i];
  s = warp_sum(s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    s = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(out, s);
  }
}

// adamw.py:73-98:  m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= step_size * m / (sqrt(v) + eps) ;
//                  p -= lr*wd * p   (decoupled decay applied to the already-updated p, :96-98)
// g is first scaled by the clip coefficient min(1, max_norm / (||g|| + 1e-6)) when max_norm > 0.
__global__ void adamw_prepare_kernel(const float* __restrict__ gnorm_sq, float lr, float beta1, float beta2,
                                     int correct_bias, int* __restrict__ step_count, float* __restrict__ out) {
  pdl_grid_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (!isfinite(*gnorm_sq)) { *out = 0.f; return; }
  const int t = ++*step_count;
  float ss = lr;
  if (correct_bias) ss = lr * sqrtf(1.f - powf(beta2, static_cast<float>(t))) / (1.f - powf(beta1, static_cast<float>(t)));
  *out = ss;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float beta1, float beta2, float eps, float step_size,
                             const float* __restrict__ step_size_dev, float lr_wd, const float* __restrict__ gnorm_sq,
                             float max_norm) {
  pdl_grid_sync();
  float clip = 1.f;
  if (gnorm_sq && !isfinite(*gnorm_sq)) return;   // fp16 gradient overflow: skip this update (dynamic loss scaling)
  if (step_size_dev) step_size = *step_size_dev;
  if (max_norm > 0.f && gnorm_sq) {
    const float c = max_norm / (sqrtf(*gnorm_sq) + 1e-6f);
    clip = c < 1.f ? c : 1.f;
  }
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gg = gp[j] * clip;
      mp[j] = mp[j] * beta1 + (1.f - beta1) * gg;
      vp[j] = vp[j] * beta2 + (1.f - beta2) * gg * gg;
      float q = pp[j] - step_size * mp[j] / (sqrtf(vp[j]) + eps);
      q -= lr_wd * q;
      pp[j] = q;
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

}  // namespace
}  // namespace alpro

using namespace alpro;

extern "C" int alpro_sumsq(const float* x, int64_t n, float* out, void* stream) {
  ALPRO_REQUIRE(x && out && n > 0 && aligned16(x), "alpro_sumsq: bad args");
  long long g = cdiv(cdiv(n, 4), 256);
  if (g > num_sms() * 8) g = num_sms() * 8;
  launch_k(sumsq_kernel, static_cast<unsigned>(g), 256, 0, static_cast<cudaStream_t>(stream), x, n, out);
  ALPRO_CHECK_LAUNCH("alpro_sumsq");
  return 0;
}

extern "C" int alpro_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2,
                                float eps, float step_size, float lr_wd, const float* gnorm_sq, float max_norm,
                                void* stream) {
  ALPRO_REQUIRE(p && g && m && v && n > 0 && (n % 4) == 0, "alpro_adamw_step: n must be a positive multiple of 4");
  ALPRO_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "alpro_adamw_step: alignment");
  long long gr = cdiv(n / 4, 256);
  if (gr > num_sms() * 8) gr = num_sms() * 8;
  launch_k(adamw_kernel, static_cast<unsigned>(gr), 256, 0, static_cast<cudaStream_t>(stream), p, g, m, v, n, beta1, beta2, eps,
                                                                                        step_size, nullptr, lr_wd,
                                                                                        gnorm_sq, max_norm);
  ALPRO_CHECK_LAUNCH("alpro_adamw_step");
  return 0;
}

extern "C" int alpro_adamw_prepare(const float* gnorm_sq, float lr, float beta1, float beta2, int correct_bias,
                                   int* step_count, float* step_size_out, void* stream) {
  ALPRO_REQUIRE(gnorm_sq && step_count && step_size_out, "alpro_adamw_prepare: bad args");
  launch_k(adamw_prepare_kernel, 1, 32, 0, static_cast<cudaStream_t>(stream), gnorm_sq, lr, beta1, beta2, correct_bias,
                                                                       step_count, step_size_out);
  ALPRO_CHECK_LAUNCH("alpro_adamw_prepare");
  return 0;
}

extern "C" int alpro_adamw_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2,
                                    float eps, const float* step_size_dev, float lr_wd, const float* gnorm_sq,
                                    float max_norm, void* stream) {
  ALPRO_REQUIRE(p && g && m && v && step_size_dev && n > 0 && (n % 4) == 0, "alpro_adamw_step_dev: bad args");
  ALPRO_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "alpro_adamw_step_dev: alignment");
  long long gr = cdiv(n / 4, 256);
  if (gr > num_sms() * 8) gr = num_sms() * 8;
  launch_k(adamw_kernel, static_cast<unsigned>(gr), 256, 0, static_cast<cudaStream_t>(stream), p, g, m, v, n, beta1, beta2, eps,
                                                                                        0.f, step_size_dev, lr_wd,
                                                                                        gnorm_sq, max_norm);
  ALPRO_CHECK_LAUNCH("alpro_adamw_step_dev");
  return 0;
}
