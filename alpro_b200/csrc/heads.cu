// Small fp32 kernels for the task heads (VTC / VTM / MLM-CE / MPM; alpro_models.py:103-128, 209-232, 288-344, 346-373).
// These operate on O(batch) rows, are latency- not throughput-bound, and are kept in fp32 on CUDA cores: the VTC head
// amplifies errors by 1/temp ~ 14x, so its projections, normalisation and similarities do not go through 16-bit
// operands (SURVEY.md §7 "Precision vs the 1e-3 gate").
#include "common.h"
#include "ptx.cuh"
#include "rng.cuh"

namespace alpro {
namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (warp == 0) {
    r = warp_max(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

__device__ __forceinline__ float resolve_alpha(float alpha, const float* alpha_dev, int mode) {
  if (mode == 1) return alpha * (*alpha_dev);
  if (mode == 2) return alpha / (*alpha_dev);
  return alpha;
}

// ------------------------------------------------------------------------------------------------ small linear
// y[m,n] = act(alpha * sum_k x[m,k] W[n,k] + b[n]); one warp per output, lanes over k.
__global__ void small_linear_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ W,
                                        long long ldw, const float* __restrict__ b, float* __restrict__ y,
                                        long long ldy, int M, int N, int K, float alpha, const float* alpha_dev,
                                        int alpha_mode, int relu) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long widx = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (widx >= static_cast<long long>(M) * N) return;
  const int m = static_cast<int>(widx / N), n = static_cast<int>(widx - static_cast<long long>(m) * N);
  const float* xr = x + m * ldx;
  const float* wr = W + n * ldw;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += xr[k] * wr[k];
  s = warp_sum(s);
  if (lane == 0) {
    float v = s * resolve_alpha(alpha, alpha_dev, alpha_mode) + (b ? b[n] : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    y[m * ldy + n] = v;
  }
}

// dx[m,k] (+)= alpha * sum_n dy'[m,n] W[n,k],  dy' = dy * (yact > 0) when yact != null
__global__ void small_linear_dx_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ yact,
                                       long long ldya, const float* __restrict__ W, long long ldw,
                                       float* __restrict__ dx, long long lddx, int M, int N, int K, float alpha,
                                       const float* alpha_dev, int alpha_mode, int accumulate) {
  pdl_grid_sync();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (k >= K) return;
  float s = 0.f;
  for (int n = 0; n < N; ++n) {
    float g = dy[m * lddy + n];
    if (yact && !(yact[m * ldya + n] > 0.f)) g = 0.f;
    s += g * W[n * ldw + k];
  }
  s *= resolve_alpha(alpha, alpha_dev, alpha_mode);
  float* o = dx + m * lddx + k;
  *o = accumulate ? *o + s : s;
}

// Same contraction for long N (the 768-1536-term sums of the composed-Linear un-mixing, the MPM head, the VTC sims): the
// kernel above walks N serially in 1-3 blocks (r02g launch list: 19 launches, 1.7 ms of the step). Here a block owns
// 128 columns k (a float4 per lane) and one slab of 64 rows n, its 8 warps take every 8th row of the slab, partial
// sums meet in shared memory and are added to dx with one atomic per element (dx is zeroed first unless accumulating).
__global__ void __launch_bounds__(256) small_linear_dx_split_kernel(const float* __restrict__ dy, long long lddy,
                                                                    const float* __restrict__ yact, long long ldya,
                                                                    const float* __restrict__ W, long long ldw,
                                                                    float* __restrict__ dx, long long lddx, int N, int K,
                                                                    float alpha, const float* alpha_dev, int alpha_mode) {
  pdl_grid_sync();
  __shared__ float4 part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.y;
  const int k = (blockIdx.x * 32 + lane) * 4;
  const int n0 = blockIdx.z * 64, n1 = min(N, n0 + 64);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K) {
#pragma unroll 4
    for (int n = n0 + warp; n < n1; n += 8) {
      float g = dy[m * lddy + n];
      if (yact && !(yact[m * ldya + n] > 0.f)) g = 0.f;
      const float4 w = *reinterpret_cast<const float4*>(W + n * ldw + k);
      acc.x = fmaf(g, w.x, acc.x); acc.y = fmaf(g, w.y, acc.y); acc.z = fmaf(g, w.z, acc.z); acc.w = fmaf(g, w.w, acc.w);
    }
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && k < K) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 o = part[w][lane];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    const float a = resolve_alpha(alpha, alpha_dev, alpha_mode);
    float* o = dx + m * lddx + k;
    atomicAdd(o + 0, acc.x * a); atomicAdd(o + 1, acc.y * a); atomicAdd(o + 2, acc.z * a); atomicAdd(o + 3, acc.w * a);
  }
}

// dW[n,k] (+)= alpha * sum_m dy'[m,n] x[m,k];  db[n] (+)= sum_m dy'[m,n]   (bias handled by the k == 0 thread row)
__global__ void small_linear_dw_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ yact,
                                       long long ldya, const float* __restrict__ x, long long ldx,
                                       float* __restrict__ dW, long long lddw, float* __restrict__ db, int M, int N,
                                       int K, float alpha, const float* alpha_dev, int alpha_mode, float out_scale,
                                       int accumulate) {
  pdl_grid_sync();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (k >= K) return;
  float s = 0.f, sb = 0.f;
  for (int m = 0; m < M; ++m) {
    float g = dy[m * lddy + n];
    if (yact && !(yact[m * ldya + n] > 0.f)) g = 0.f;
    s += g * x[m * ldx + k];
    sb += g;
  }
  s *= resolve_alpha(alpha, alpha_dev, alpha_mode) * out_scale;
  float* o = dW + n * lddw + k;
  *o = accumulate ? *o + s : s;
  if (db && k == 0) db[n] = accumulate ? db[n] + sb * out_scale : sb * out_scale;
}

// ------------------------------------------------------------------------------------------------ L2 normalise
// y = x / max(||x||, eps)   (F.normalize, alpro_models.py:103,205,750,761); one warp per row
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norm, int M,
                                  int d, float eps) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += x[m * d + c] * x[m * d + c];
  const float nrm = fmaxf(sqrtf(warp_sum(s)), eps);
  for (int c = lane; c < d; c += 32) y[m * d + c] = x[m * d + c] / nrm;
  if (lane == 0) norm[m] = nrm;
}
// dx = (dy - y (y . dy)) / norm       (exact when norm > eps, which holds for any non-degenerate feature)
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                  const float* __restrict__ norm, float* __restrict__ dx, int M, int d) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += dy[m * d + c] * y[m * d + c];
  s = warp_sum(s);
  const float inv = 1.f / norm[m];
  for (int c = lane; c < d; c += 32) dx[m * d + c] = (dy[m * d + c] - y[m * d + c] * s) * inv;
}

// ------------------------------------------------------------------------------------------------ softmax CE
// One block per row. hard labels (int64, negative = ignored) or soft labels (fp32 rows); optional row_ignore bytes.
// row_loss = -sum_c t_c log_softmax(x)_c ; row_valid in {0,1}; row_tsum = sum_c t_c ; row_lse = logsumexp(x).
__global__ void softmax_ce_fwd_kernel(const float* __restrict__ logits, long long ld, int C,
                                      const long long* __restrict__ hard, const float* __restrict__ soft,
                                      long long ld_soft, const uint8_t* __restrict__ row_ignore,
                                      float* __restrict__ row_loss, float* __restrict__ row_lse,
                                      float* __restrict__ row_valid, float* __restrict__ row_tsum) {
  pdl_grid_sync();
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float* x = logits + r * ld;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, x[c]);
  mx = block_max(mx, red);
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += __expf(x[c] - mx);
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  float tx = 0.f, ts = 0.f;
  bool valid = !(row_ignore && row_ignore[r]);
  if (hard) {
    const long long lab = hard[r];
    if (lab < 0 || lab >= C) valid = false;
    if (valid && threadIdx.x == 0) { tx = x[lab]; ts = 1.f; }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float t = soft[r * ld_soft + c];
      tx += t * x[c];
      ts += t;
    }
  }
  tx = block_sum(tx, red);
  ts = block_sum(ts, red);
  if (threadIdx.x == 0) {
    row_loss[r] = valid ? (ts * lse - tx) : 0.f;
    row_lse[r] = lse;
    row_valid[r] = valid ? 1.f : 0.f;
    row_tsum[r] = ts;
  }
}

// loss = sum_r row_loss / denom,  denom = denom_mode 0: #valid rows ; 1: R (plain mean over all rows)
__global__ void loss_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_valid, int R,
                                   int denom_mode, float* __restrict__ loss_out, float* __restrict__ denom_out) {
  pdl_grid_sync();
  __shared__ float red[32];
  float s = 0.f, v = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    s += row_loss[r];
    v += row_valid[r];
  }
  s = block_sum(s, red);
  v = block_sum(v, red);
  if (threadIdx.x == 0) {
    const float den = denom_mode == 1 ? static_cast<float>(R) : v;
    *denom_out = den;
    *loss_out = s / den;
  }
}

// dlogits[r,c] = g * valid_r / denom * (softmax_c * tsum_r - t_c);  g = gscale * (*gptr or 1)
__global__ void softmax_ce_bwd_kernel(const float* __restrict__ logits, long long ld, int C,
                                      const long long* __restrict__ hard, const float* __restrict__ soft,
                                      long long ld_soft, const float* __restrict__ row_lse,
                                      const float* __restrict__ row_valid, const float* __restrict__ row_tsum,
                                      const float* __restrict__ denom, const float* __restrict__ gptr, float gscale,
                                      float* __restrict__ out32, uint16_t* __restrict__ out16, int fmt,
                                      long long ld_out, int C_out) {
  pdl_grid_sync();
  const int r = blockIdx.x;
  const float* x = logits + r * ld;
  const float coef = gscale * (gptr ? *gptr : 1.f) * row_valid[r] / (*denom);
  const float lse = row_lse[r], ts = row_tsum[r];
  const long long lab = hard ? hard[r] : -1;
  for (int c = threadIdx.x; c < C_out; c += blockDim.x) {
    float v = 0.f;
    if (c < C && coef != 0.f) {
      const float t = hard ? (c == lab ? 1.f : 0.f) : soft[r * ld_soft + c];
      v = coef * (__expf(x[c] - lse) * ts - t);
    }
    if (out32) out32[r * ld_out + c] = v;
    if (out16) out16[r * ld_out + c] = f32_to_16(v, fmt);
  }
}

// dtemp += coef * (-1/temp) * (sum dsa*sa + sum dsb*sb)          (sim = raw / temp  =>  d/dtemp = -sim / temp)
__global__ void temp_grad_kernel(const float* __restrict__ dsa, const float* __restrict__ sa, long long na,
                                 const float* __restrict__ dsb, const float* __restrict__ sb, long long nb,
                                 const float* __restrict__ temp, float* __restrict__ dtemp, float coef) {
  pdl_grid_sync();
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = threadIdx.x; i < na; i += blockDim.x) s += dsa[i] * sa[i];
  for (long long i = threadIdx.x; i < nb; i += blockDim.x) s += dsb[i] * sb[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *dtemp += -coef * s / (*temp);
}

__global__ void clamp_scalar_kernel(float* p, float lo, float hi) {
  pdl_grid_sync(); *p = fminf(fmaxf(*p, lo), hi); }

// ------------------------------------------------------------------------------------------------ MPM pooling
// pooled[b] = sum_n w[b,n] x[b, row0 + n] / sum_n w[b,n],  w = 1 - patch_mask   (alpro_models.py:214-224)
__global__ void masked_mean_fwd_kernel(const float* __restrict__ x, long long seq_stride, int row0,
                                       const float* __restrict__ patch_mask, int Np, int h, float* __restrict__ out) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  float cnt = 0.f;
  for (int n = 0; n < Np; ++n) cnt += 1.f - patch_mask[b * Np + n];
  for (int c = threadIdx.x; c < h; c += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < Np; ++n) s += (1.f - patch_mask[b * Np + n]) * x[b * seq_stride + static_cast<long long>(row0 + n) * h + c];
    out[b * h + c] = s / cnt;
  }
}
__global__ void masked_mean_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ patch_mask, int Np,
                                       int h, float* __restrict__ dx, long long seq_stride, int row0) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  float cnt = 0.f;
  for (int n = 0; n < Np; ++n) cnt += 1.f - patch_mask[b * Np + n];
  for (int c = threadIdx.x; c < h; c += blockDim.x) {
    const float g = dout[b * h + c] / cnt;
    for (int n = 0; n < Np; ++n)
      dx[b * seq_stride + static_cast<long long>(row0 + n) * h + c] += (1.f - patch_mask[b * Np + n]) * g;
  }
}

// ------------------------------------------------------------------------------------------------ row-block copies
// out[i, l, :] = src[(s0 + i), l, :] for l < L (first L rows of each R-row sequence)     (mlm txt_output, :366-367)
__global__ void take_rows_fwd_kernel(const float* __restrict__ src, int R, int s0, int n, int L, int h,
                                     float* __restrict__ out32, uint16_t* __restrict__ out16, int fmt) {
  pdl_grid_sync();
  const long long total = static_cast<long long>(n) * L * h;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / h;
    const int c = static_cast<int>(i - row * h);
    const int s = static_cast<int>(row / L), l = static_cast<int>(row - static_cast<long long>(s) * L);
    const float v = src[((static_cast<long long>(s0 + s)) * R + l) * h + c];
    if (out32) out32[i] = v;
    if (out16) out16[i] = f32_to_16(v, fmt);
  }
}
__global__ void take_rows_bwd_kernel(const float* __restrict__ dout, int R, int s0, int n, int L, int h,
                                     float* __restrict__ dsrc) {
  pdl_grid_sync();
  const long long total = static_cast<long long>(n) * L * h;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / h;
    const int c = static_cast<int>(i - row * h);
    const int s = static_cast<int>(row / L), l = static_cast<int>(row - static_cast<long long>(s) * L);
    dsrc[((static_cast<long long>(s0 + s)) * R + l) * h + c] += dout[i];
  }
}

// hard-negative sampling weights: softmax over the local [b,b] block of sim with -inf on the diagonal
// (alpro_models.py:288-299, 819-828); one warp per row
__global__ void neg_weights_kernel(const float* __restrict__ sim, long long ld, int col0, int b,
                                   float* __restrict__ w) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= b) return;
  float mx = -INFINITY;
  for (int c = lane; c < b; c += 32)
    if (c != r) mx = fmaxf(mx, sim[r * ld + col0 + c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < b; c += 32)
    if (c != r) s += __expf(sim[r * ld + col0 + c] - mx);
  s = warp_sum(s);
  for (int c = lane; c < b; c += 32) w[r * b + c] = c == r ? 0.f : __expf(sim[r * ld + col0 + c] - mx) / s;
}

// Hard-negative DRAW fused with the weights (alpro_models.py:301-316, 833-844: `torch.multinomial(weights[b], 1).item()`
// per row, 2*B host synchronisations per step in the reference): one warp per row computes the softmax weights of the
// local block with the diagonal excluded and samples one index by inverse CDF with a Philox4x32-10 uniform
// (counter = (row, draw, 0, 0), key = seed). The index is a device int64; nothing returns to the host.
__global__ void neg_sample_kernel(const float* __restrict__ sim, long long ld, int col0, int b, uint32_t seed_lo,
                                  uint32_t seed_hi, uint32_t draw, float* __restrict__ w, long long* __restrict__ idx) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= b) return;
  const float* x = sim + r * ld + col0;
  float mx = -INFINITY;
  for (int c = lane; c < b; c += 32)
    if (c != r) mx = fmaxf(mx, x[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < b; c += 32)
    if (c != r) s += __expf(x[c] - mx);
  s = warp_sum(s);
  if (w)
    for (int c = lane; c < b; c += 32) w[r * b + c] = c == r ? 0.f : __expf(x[c] - mx) / s;
  const Philox4 rnd = philox4x32_10(static_cast<uint32_t>(r), draw, 0u, 0u, seed_lo, seed_hi);
  const float u = (static_cast<float>(rnd.x >> 8) + 0.5f) * (1.0f / 16777216.0f);   // 24-bit uniform in (0, 1)
  const float target = u * s;
  float run = 0.f;
  int pick = -1, last = -1;
  for (int c0 = 0; c0 < b && pick < 0; c0 += 32) {
    const int c = c0 + lane;
    const float e = (c < b && c != r) ? __expf(x[c] - mx) : 0.f;
    float sc = e;   // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, sc, o);
      if (lane >= o) sc += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, e > 0.f && run + sc > target);
    const unsigned any = __ballot_sync(0xffffffffu, e > 0.f);
    if (any) last = c0 + 31 - __clz(any);
    if (hit) pick = c0 + __ffs(hit) - 1;
    run += __shfl_sync(0xffffffffu, sc, 31);
  }
  if (pick < 0) pick = last >= 0 ? last : (r == 0 && b > 1 ? 1 : 0);   // rounding at the top of the CDF
  if (lane == 0) idx[r] = pick;
}

__global__ void philox_kat_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Philox4 v = philox4x32_10(in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3], in[6 * i + 4], in[6 * i + 5]);
  out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
}

// out16 = dy * dact, dact = gelu'(pre) saved by the forward GEMM epilogue   (MLM transform backward, xbert.py:659-661)
__global__ void gelu_grad_mul_kernel(const float* __restrict__ dy, const uint16_t* __restrict__ pre, int pre_fmt,
                                     uint16_t* __restrict__ out, int out_fmt, long long n) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = f32_to_16(dy[i] * f16_to_32(pre[i], pre_fmt), out_fmt);
}

// Prompter._compute_soft_labels (alpro_models.py:525-529): soft = softmax(sim) per row;
// ignore = (argmax index < 0.2)  i.e. argmax == 0 (reference behaviour kept bug-compatible).
__global__ void pseudo_labels_kernel(const float* __restrict__ sim, int C, float* __restrict__ soft,
                                     uint8_t* __restrict__ ignore) {
  pdl_grid_sync();
  __shared__ float red[32];
  __shared__ int redi[32];
  const int r = blockIdx.x;
  const float* x = sim + static_cast<long long>(r) * C;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    if (x[c] > mx) { mx = x[c]; arg = c; }
  // block arg-max with first-index tie break (torch.max returns the first maximal index on CPU)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  if (lane == 0) { red[warp] = mx; redi[warp] = arg; }
  __syncthreads();
  if (warp == 0) {
    float m2 = lane < nw ? red[lane] : -INFINITY;
    int a2 = lane < nw ? redi[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m2, o);
      const int oa = __shfl_xor_sync(0xffffffffu, a2, o);
      if (om > m2 || (om == m2 && oa < a2)) { m2 = om; a2 = oa; }
    }
    if (lane == 0) { red[0] = m2; redi[0] = a2; }
  }
  __syncthreads();
  mx = red[0];
  arg = redi[0];
  __syncthreads();
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += __expf(x[c] - mx);
  se = block_sum(se, red);
  const float inv = 1.f / se;
  for (int c = threadIdx.x; c < C; c += blockDim.x) soft[static_cast<long long>(r) * C + c] = __expf(x[c] - mx) * inv;
  if (threadIdx.x == 0) ignore[r] = static_cast<float>(arg) < 0.2f ? 1 : 0;
}

inline int grid_for(long long work_items, int block) {
  long long g = cdiv(work_items, block);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace
}  // namespace alpro

using namespace alpro;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int alpro_small_linear_fwd(const float* x, int64_t ldx, const float* W, int64_t ldw, const float* b, float* y,
                                      int64_t ldy, int M, int N, int K, float alpha, const float* alpha_dev,
                                      int alpha_mode, int relu, void* stream) {
  ALPRO_REQUIRE(x && W && y && M > 0 && N > 0 && K > 0, "alpro_small_linear_fwd: bad args");
  ALPRO_REQUIRE(alpha_mode == 0 || alpha_dev, "alpro_small_linear_fwd: alpha_dev missing");
  const long long warps = static_cast<long long>(M) * N;
  launch_k(small_linear_fwd_kernel, static_cast<unsigned>(cdiv(warps, 8)), 256, 0, ST, x, ldx, W, ldw, b, y, ldy, M, N, K,
                                                                                 alpha, alpha_dev, alpha_mode, relu);
  ALPRO_CHECK_LAUNCH("alpro_small_linear_fwd");
  return 0;
}

extern "C" int alpro_small_linear_bwd(const float* dy, int64_t lddy, const float* yact, int64_t ldya, const float* x,
                                      int64_t ldx, const float* W, int64_t ldw, float* dx, int64_t lddx, int dx_accumulate,
                                      float* dW, int64_t lddw, float* db, int dw_accumulate, int M, int N, int K,
                                      float alpha, const float* alpha_dev, int alpha_mode, float dw_scale,
                                      void* stream) {
  ALPRO_REQUIRE(dy && M > 0 && N > 0 && K > 0, "alpro_small_linear_bwd: bad args");
  if (dx) {
    ALPRO_REQUIRE(W, "alpro_small_linear_bwd: W needed for dx");
    const bool vec = N >= 128 && (K % 4) == 0 && (ldw % 4) == 0 && aligned16(W) && M <= 65535;
    if (vec) {
      if (!dx_accumulate) {
        const cudaError_t e = cudaMemset2DAsync(dx, static_cast<size_t>(lddx) * 4, 0, static_cast<size_t>(K) * 4, M, ST);
        if (e != cudaSuccess) {
          set_last_error("alpro_small_linear_bwd: memset failed: %s", cudaGetErrorString(e));
          return static_cast<int>(e);
        }
      }
      dim3 grid(static_cast<unsigned>(cdiv(K, 128)), M, static_cast<unsigned>(cdiv(N, 64)));
      launch_k(small_linear_dx_split_kernel, grid, 256, 0, ST, dy, lddy, yact, ldya, W, ldw, dx, lddx, N, K, alpha, alpha_dev,
                                                         alpha_mode);
    } else {
      dim3 grid(static_cast<unsigned>(cdiv(K, 128)), M);
      launch_k(small_linear_dx_kernel, grid, 128, 0, ST, dy, lddy, yact, ldya, W, ldw, dx, lddx, M, N, K, alpha, alpha_dev,
                                                   alpha_mode, dx_accumulate);
    }
    ALPRO_CHECK_LAUNCH("alpro_small_linear_bwd(dx)");
  }
  if (dW) {
    ALPRO_REQUIRE(x, "alpro_small_linear_bwd: x needed for dW");
    dim3 grid(static_cast<unsigned>(cdiv(K, 128)), N);
    launch_k(small_linear_dw_kernel, grid, 128, 0, ST, dy, lddy, yact, ldya, x, ldx, dW, lddw, db, M, N, K, alpha, alpha_dev,
                                                 alpha_mode, dw_scale, dw_accumulate);
    ALPRO_CHECK_LAUNCH("alpro_small_linear_bwd(dW)");
  }
  return 0;
}

extern "C" int alpro_l2norm_fwd(const float* x, float* y, float* norm, int M, int d, float eps, void* stream) {
  ALPRO_REQUIRE(x && y && norm && M > 0, "alpro_l2norm_fwd: bad args");
  launch_k(l2norm_fwd_kernel, static_cast<unsigned>(cdiv(M, 4)), 128, 0, ST, x, y, norm, M, d, eps);
  ALPRO_CHECK_LAUNCH("alpro_l2norm_fwd");
  return 0;
}
extern "C" int alpro_l2norm_bwd(const float* dy, const float* y, const float* norm, float* dx, int M, int d,
                                void* stream) {
  ALPRO_REQUIRE(dy && y && norm && dx && M > 0, "alpro_l2norm_bwd: bad args");
  launch_k(l2norm_bwd_kernel, static_cast<unsigned>(cdiv(M, 4)), 128, 0, ST, dy, y, norm, dx, M, d);
  ALPRO_CHECK_LAUNCH("alpro_l2norm_bwd");
  return 0;
}

extern "C" int alpro_softmax_ce_fwd(const float* logits, int64_t ld, int R, int C, const int64_t* hard,
                                    const float* soft, int64_t ld_soft, const uint8_t* row_ignore, float* row_loss,
                                    float* row_lse, float* row_valid, float* row_tsum, int denom_mode, float* loss_out,
                                    float* denom_out, void* stream) {
  ALPRO_REQUIRE(logits && (hard || soft) && row_loss && row_lse && row_valid && row_tsum && loss_out && denom_out && R > 0,
                "alpro_softmax_ce_fwd: bad args");
  launch_k(softmax_ce_fwd_kernel, R, 256, 0, ST, logits, ld, C, reinterpret_cast<const long long*>(hard), soft, ld_soft,
                                           row_ignore, row_loss, row_lse, row_valid, row_tsum);
  ALPRO_CHECK_LAUNCH("alpro_softmax_ce_fwd");
  launch_k(loss_reduce_kernel, 1, 256, 0, ST, row_loss, row_valid, R, denom_mode, loss_out, denom_out);
  ALPRO_CHECK_LAUNCH("alpro_softmax_ce_fwd(reduce)");
  return 0;
}

extern "C" int alpro_softmax_ce_bwd(const float* logits, int64_t ld, int R, int C, const int64_t* hard,
                                    const float* soft, int64_t ld_soft, const float* row_lse, const float* row_valid,
                                    const float* row_tsum, const float* denom, const float* gptr, float gscale,
                                    float* out32, void* out16, int out16_fmt, int64_t ld_out, int C_out, void* stream) {
  ALPRO_REQUIRE(logits && (hard || soft) && row_lse && row_valid && row_tsum && denom && (out32 || out16) && C_out >= C,
                "alpro_softmax_ce_bwd: bad args");
  launch_k(softmax_ce_bwd_kernel, R, 256, 0, ST, logits, ld, C, reinterpret_cast<const long long*>(hard), soft, ld_soft,
                                           row_lse, row_valid, row_tsum, denom, gptr, gscale, out32,
                                           static_cast<uint16_t*>(out16), out16_fmt, ld_out, C_out);
  ALPRO_CHECK_LAUNCH("alpro_softmax_ce_bwd");
  return 0;
}

extern "C" int alpro_temp_grad(const float* dsa, const float* sa, int64_t na, const float* dsb, const float* sb,
                               int64_t nb, const float* temp, float* dtemp, float coef, void* stream) {
  ALPRO_REQUIRE(dsa && sa && temp && dtemp, "alpro_temp_grad: bad args");
  launch_k(temp_grad_kernel, 1, 256, 0, ST, dsa, sa, na, dsb, sb, dsb ? nb : 0, temp, dtemp, coef);
  ALPRO_CHECK_LAUNCH("alpro_temp_grad");
  return 0;
}

extern "C" int alpro_clamp_scalar(float* p, float lo, float hi, void* stream) {
  ALPRO_REQUIRE(p, "alpro_clamp_scalar: null");
  launch_k(clamp_scalar_kernel, 1, 1, 0, ST, p, lo, hi);
  ALPRO_CHECK_LAUNCH("alpro_clamp_scalar");
  return 0;
}

extern "C" int alpro_masked_mean_fwd(const float* x, int64_t seq_stride, int row0, const float* patch_mask, int B,
                                     int Np, int h, float* out, void* stream) {
  ALPRO_REQUIRE(x && patch_mask && out && B > 0, "alpro_masked_mean_fwd: bad args");
  launch_k(masked_mean_fwd_kernel, B, 256, 0, ST, x, seq_stride, row0, patch_mask, Np, h, out);
  ALPRO_CHECK_LAUNCH("alpro_masked_mean_fwd");
  return 0;
}
extern "C" int alpro_masked_mean_bwd(const float* dout, const float* patch_mask, int B, int Np, int h, float* dx,
                                     int64_t seq_stride, int row0, void* stream) {
  ALPRO_REQUIRE(dout && patch_mask && dx && B > 0, "alpro_masked_mean_bwd: bad args");
  launch_k(masked_mean_bwd_kernel, B, 256, 0, ST, dout, patch_mask, Np, h, dx, seq_stride, row0);
  ALPRO_CHECK_LAUNCH("alpro_masked_mean_bwd");
  return 0;
}

extern "C" int alpro_take_rows_fwd(const float* src, int R, int s0, int n, int L, int h, float* out32, void* out16,
                                   int fmt, void* stream) {
  ALPRO_REQUIRE(src && (out32 || out16) && n > 0, "alpro_take_rows_fwd: bad args");
  launch_k(take_rows_fwd_kernel, grid_for(static_cast<long long>(n) * L * h, 256), 256, 0, ST, 
      src, R, s0, n, L, h, out32, static_cast<uint16_t*>(out16), fmt);
  ALPRO_CHECK_LAUNCH("alpro_take_rows_fwd");
  return 0;
}
extern "C" int alpro_take_rows_bwd(const float* dout, int R, int s0, int n, int L, int h, float* dsrc, void* stream) {
  ALPRO_REQUIRE(dout && dsrc && n > 0, "alpro_take_rows_bwd: bad args");
  launch_k(take_rows_bwd_kernel, grid_for(static_cast<long long>(n) * L * h, 256), 256, 0, ST, dout, R, s0, n, L, h, dsrc);
  ALPRO_CHECK_LAUNCH("alpro_take_rows_bwd");
  return 0;
}

extern "C" int alpro_neg_weights(const float* sim, int64_t ld, int col0, int b, float* w, void* stream) {
  ALPRO_REQUIRE(sim && w && b > 0, "alpro_neg_weights: bad args");
  launch_k(neg_weights_kernel, static_cast<unsigned>(cdiv(b, 4)), 128, 0, ST, sim, ld, col0, b, w);
  ALPRO_CHECK_LAUNCH("alpro_neg_weights");
  return 0;
}

extern "C" int alpro_neg_sample(const float* sim, int64_t ld, int col0, int b, uint32_t seed_lo, uint32_t seed_hi,
                                uint32_t draw, float* w, int64_t* idx, void* stream) {
  ALPRO_REQUIRE(sim && idx && b > 0, "alpro_neg_sample: bad args");
  launch_k(neg_sample_kernel, static_cast<unsigned>(cdiv(b, 4)), 128, 0, ST, sim, ld, col0, b, seed_lo, seed_hi, draw, w,
                                                                      reinterpret_cast<long long*>(idx));
  ALPRO_CHECK_LAUNCH("alpro_neg_sample");
  return 0;
}

extern "C" int alpro_philox4x32_10(const uint32_t* ctr_key, uint32_t* out, int n, void* stream) {
  ALPRO_REQUIRE(ctr_key && out && n > 0, "alpro_philox4x32_10: bad args");
  launch_k(philox_kat_kernel, static_cast<unsigned>(cdiv(n, 128)), 128, 0, ST, ctr_key, out, n);
  ALPRO_CHECK_LAUNCH("alpro_philox4x32_10");
  return 0;
}

extern "C" int alpro_gelu_grad_mul(const float* dy, const void* pre, int pre_fmt, void* out, int out_fmt, int64_t n,
                                   void* stream) {
  ALPRO_REQUIRE(dy && pre && out && n > 0, "alpro_gelu_grad_mul: bad args");
  launch_k(gelu_grad_mul_kernel, grid_for(n, 256), 256, 0, ST, dy, static_cast<const uint16_t*>(pre), pre_fmt,
                                                         static_cast<uint16_t*>(out), out_fmt, n);
  ALPRO_CHECK_LAUNCH("alpro_gelu_grad_mul");
  return 0;
}

extern "C" int alpro_pseudo_labels(const float* sim, int R, int C, float* soft, uint8_t* ignore, void* stream) {
  ALPRO_REQUIRE(sim && soft && ignore && R > 0 && C > 0, "alpro_pseudo_labels: bad args");
  launch_k(pseudo_labels_kernel, R, 256, 0, ST, sim, C, soft, ignore);
  ALPRO_CHECK_LAUNCH("alpro_pseudo_labels");
  return 0;
}
