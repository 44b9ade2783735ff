// HBM-bound kernels of the ALPRO path: casts, LayerNorm fwd/bwd, bias-gradient column sums, patch gathering,
// TimeSformer embedding assembly / temporal pooling, BERT embedding gather/scatter, fusion-input gather/concat.
// All are written for coalesced 128-bit accesses with one warp per token row where a row reduction is needed.
#include "common.h"
#include "ptx.cuh"
#include "rng.cuh"

namespace alpro {
namespace {

constexpr int LN_MAX_V4 = 8;  // per-lane float4 registers -> d <= 1024

__device__ __forceinline__ float load_any(const void* p, int kind, long long idx) {
  if (kind == 0) return reinterpret_cast<const float*>(p)[idx];
  return f16_to_32(reinterpret_cast<const uint16_t*>(p)[idx], kind - 1);
}

// ------------------------------------------------------------------------------------------------ cast
__global__ void cast_f32_to_16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n, int fmt) {
  pdl_grid_sync();
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    uint2 w;
    w.x = pack2_16(v.x, v.y, fmt);
    w.y = pack2_16(v.z, v.w, fmt);
    *reinterpret_cast<uint2*>(dst + i) = w;
  }
  if (i < n) {  // tail (n % 4 != 0): handled by the thread whose window crosses n
    for (long long j = i; j < n; ++j) dst[j] = f32_to_16(src[j], fmt);
  }
}

// Batched form for the once-per-step refresh of every 16-bit weight operand (OperandCache.refresh): `table` holds one
// (src, dst, count) triple per chunk of at most CAST_CHUNK elements (count % 4 == 0 except for a tensor's last chunk);
// one block per chunk, so ~200 per-tensor launches become one.
constexpr int CAST_CHUNK = 16384;
__global__ void __launch_bounds__(256) cast_multi_kernel(const long long* __restrict__ table, int fmt) {
  pdl_grid_sync();
  const long long* e = table + 3LL * blockIdx.x;
  const float* __restrict__ src = reinterpret_cast<const float*>(e[0]);
  uint16_t* __restrict__ dst = reinterpret_cast<uint16_t*>(e[1]);
  const int n = static_cast<int>(e[2]);
  const int n4 = n >> 2;
  for (int i = threadIdx.x; i < n4; i += 256) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    uint2 w;
    w.x = pack2_16(v.x, v.y, fmt);
    w.y = pack2_16(v.z, v.w, fmt);
    reinterpret_cast<uint2*>(dst)[i] = w;
  }
  for (int j = (n4 << 2) + threadIdx.x; j < n; j += 256) dst[j] = f32_to_16(src[j], fmt);
}

// ------------------------------------------------------------------------------------------------ dropout mask
// out[i] = keep_i / (1 - p) with keep_i ~ Bernoulli(1 - p) from the counter hash (nn.Dropout, xbert.py:178,331,358,436)
__global__ void dropout_mask_kernel(uint16_t* __restrict__ out, int fmt, long long n, uint32_t thr, float scale,
                                    uint32_t seed) {
  pdl_grid_sync();
  const long long pairs = (n + 1) >> 1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < pairs;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t r = rand16x2(seed, static_cast<uint64_t>(i));
    const float a = (r & 0xffff) >= thr ? scale : 0.f, b = (r >> 16) >= thr ? scale : 0.f;
    if (2 * i + 1 < n) *reinterpret_cast<uint32_t*>(out + 2 * i) = pack2_16(a, b, fmt);
    else out[2 * i] = f32_to_16(a, fmt);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm fwd
// One warp per row; two-pass statistics held in registers (matches torch: var = mean((x-mean)^2), biased).
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, long long M, int d,
                                     float* __restrict__ out32, long long ld32, uint16_t* __restrict__ out16,
                                     long long ld16, int fmt, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, const uint16_t* __restrict__ mul16,
                                     long long ldmul) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nv = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[LN_MAX_V4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      v[i] = xr[c];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
      q += (a * a + b * b) + (e * e + f * f);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 y;
      y.x = (v[i].x - mean) * rstd * g.x + b.x;
      y.y = (v[i].y - mean) * rstd * g.y + b.y;
      y.z = (v[i].z - mean) * rstd * g.z + b.z;
      y.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (mul16) {   // dropout applied to the normalised output (BertEmbeddings, xbert.py:211-212)
        const uint2 mw = reinterpret_cast<const uint2*>(mul16 + row * ldmul)[c];
        y.x *= f16_to_32(static_cast<uint16_t>(mw.x & 0xffff), fmt); y.y *= f16_to_32(static_cast<uint16_t>(mw.x >> 16), fmt);
        y.z *= f16_to_32(static_cast<uint16_t>(mw.y & 0xffff), fmt); y.w *= f16_to_32(static_cast<uint16_t>(mw.y >> 16), fmt);
      }
      if (out32) reinterpret_cast<float4*>(out32 + row * ld32)[c] = y;
      if (out16) {
        uint2 w;
        w.x = pack2_16(y.x, y.y, fmt);
        w.y = pack2_16(y.z, y.w, fmt);
        reinterpret_cast<uint2*>(out16 + row * ld16)[c] = w;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm bwd
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma += dy*xhat; dbeta += dy.
// Each warp walks rows with a grid stride and keeps its dgamma/dbeta partials in registers; one block-level
// reduction + fp32 atomics at the end.
template <int NV4>
__global__ void __launch_bounds__(192, 3) layernorm_bwd_kernel(const void* __restrict__ dy, int dy_kind, long long lddy,
                                     const float* __restrict__ x, long long ldx, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, const float* __restrict__ gamma, long long M,
                                     int d, float* __restrict__ dx32, long long lddx, int accumulate,
                                     uint16_t* __restrict__ dx16, long long lddx16, int fmt, int zero_period,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, float param_scale,
                                     float* __restrict__ colsum, int colsum_zero_period,
                                     const uint16_t* __restrict__ dy_mul16, long long lddymul,
                                     const uint16_t* __restrict__ dx16_mul16, long long lddxmul,
                                     const float* __restrict__ dx16_row_scale,
                                     const float* __restrict__ colsum_row_scale) {
  pdl_grid_sync();
  // Per-warp accumulators for dgamma / dbeta / bias column sums live in shared memory ([warp][3][d], each lane owns its
  // float4 slots -> conflict-free, no atomics) so that registers stay free for more resident warps.
  extern __shared__ float red[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int nv = d >> 2;
  float4* ag = reinterpret_cast<float4*>(red + (warp * 3 + 0) * d);
  float4* ab = reinterpret_cast<float4*>(red + (warp * 3 + 1) * d);
  float4* ac = reinterpret_cast<float4*>(red + (warp * 3 + 2) * d);
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) ag[c] = ab[c] = ac[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (long long row = static_cast<long long>(blockIdx.x) * nwarps + warp; row < M;
       row += static_cast<long long>(gridDim.x) * nwarps) {
    const float mu = mean[row], rs = rstd[row];
    {
      // Pull the NEXT row this warp will visit towards L2 (one prefetch per 128-byte line, spread over the lanes): the
      // kernel is bound by exposed load latency (ncu r01i: long-scoreboard stalls, 28% warps active, 2.6 TB/s), and a
      // row ahead turns DRAM round trips into L2 hits.
      const long long nrow = row + static_cast<long long>(gridDim.x) * nwarps;
      if (nrow < M) {
        const int xlines = (d * 4 + 127) >> 7;
        for (int l = lane; l < xlines; l += 32) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(x + nrow * ldx) + (l << 7)));
          if (accumulate)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(dx32 + nrow * lddx) + (l << 7)));
        }
        const int esz = dy_kind == 0 ? 4 : 2;
        const char* dyn = reinterpret_cast<const char*>(dy) + nrow * lddy * esz;
        const int dlines = (d * esz + 127) >> 7;
        for (int l = lane; l < dlines; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(dyn + (l << 7)));
      }
    }
    // Only the raw x / dy values stay live across the two passes (xhat and g = dy*gamma are recomputed) so that the
    // kernel fits 2 blocks per SM next to its 3 x d/32 accumulator registers.
    float4 xv[NV4], dv[NV4];
    float s1 = 0.f, s2 = 0.f;
    float4* dxrow = reinterpret_cast<float4*>(dx32 + row * lddx);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        if (accumulate) asm volatile("prefetch.global.L1 [%0];" ::"l"(dxrow + c));   // read-modify-write target
        xv[i] = reinterpret_cast<const float4*>(x + row * ldx)[c];
        if (dy_kind == 0) {
          dv[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * lddy)[c];
        } else {
          const uint2 w = reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(dy) + row * lddy)[c];
          dv[i].x = f16_to_32(static_cast<uint16_t>(w.x & 0xffff), dy_kind - 1);
          dv[i].y = f16_to_32(static_cast<uint16_t>(w.x >> 16), dy_kind - 1);
          dv[i].z = f16_to_32(static_cast<uint16_t>(w.y & 0xffff), dy_kind - 1);
          dv[i].w = f16_to_32(static_cast<uint16_t>(w.y >> 16), dy_kind - 1);
        }
        if (dy_mul16) {   // upstream dropout on this LayerNorm's output
          const uint2 mw = reinterpret_cast<const uint2*>(dy_mul16 + row * lddymul)[c];
          dv[i].x *= f16_to_32(static_cast<uint16_t>(mw.x & 0xffff), fmt); dv[i].y *= f16_to_32(static_cast<uint16_t>(mw.x >> 16), fmt);
          dv[i].z *= f16_to_32(static_cast<uint16_t>(mw.y & 0xffff), fmt); dv[i].w *= f16_to_32(static_cast<uint16_t>(mw.y >> 16), fmt);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float hx = (xv[i].x - mu) * rs, hy = (xv[i].y - mu) * rs, hz = (xv[i].z - mu) * rs, hw = (xv[i].w - mu) * rs;
        const float gx = dv[i].x * gm.x, gy = dv[i].y * gm.y, gz = dv[i].z * gm.z, gw = dv[i].w * gm.w;
        s1 += (gx + gy) + (gz + gw);
        s2 += (gx * hx + gy * hy) + (gz * hz + gw * hw);
        float4 a1 = ag[c], a2 = ab[c];
        a1.x += dv[i].x * hx; a1.y += dv[i].y * hy; a1.z += dv[i].z * hz; a1.w += dv[i].w * hw;
        a2.x += dv[i].x; a2.y += dv[i].y; a2.z += dv[i].z; a2.w += dv[i].w;
        ag[c] = a1; ab[c] = a2;
      }
    }
    const float c1 = warp_sum(s1) / d, c2 = warp_sum(s2) / d;
    const bool zero16 = zero_period > 0 && (row % zero_period) == 0;
    const bool cs_on = colsum != nullptr && !(colsum_zero_period > 0 && (row % colsum_zero_period) == 0);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        float4 o;
        o.x = rs * (dv[i].x * gm.x - c1 - (xv[i].x - mu) * rs * c2);
        o.y = rs * (dv[i].y * gm.y - c1 - (xv[i].y - mu) * rs * c2);
        o.z = rs * (dv[i].z * gm.z - c1 - (xv[i].z - mu) * rs * c2);
        o.w = rs * (dv[i].w * gm.w - c1 - (xv[i].w - mu) * rs * c2);
        if (accumulate) {
          const float4 pv = dxrow[c];
          o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
        }
        dxrow[c] = o;
        // the 16-bit copy / bias column sums are the gradient of the *branch output* that feeds this residual sum:
        // it carries that branch's dropout mask and stochastic-depth row scale
        if (dx16_mul16) {
          const uint2 mw = reinterpret_cast<const uint2*>(dx16_mul16 + row * lddxmul)[c];
          o.x *= f16_to_32(static_cast<uint16_t>(mw.x & 0xffff), fmt); o.y *= f16_to_32(static_cast<uint16_t>(mw.x >> 16), fmt);
          o.z *= f16_to_32(static_cast<uint16_t>(mw.y & 0xffff), fmt); o.w *= f16_to_32(static_cast<uint16_t>(mw.y >> 16), fmt);
        }
        if (cs_on) {
          const float cs = colsum_row_scale ? colsum_row_scale[row] : (dx16_row_scale ? dx16_row_scale[row] : 1.f);
          float4 a3 = ac[c];
          a3.x += o.x * cs; a3.y += o.y * cs; a3.z += o.z * cs; a3.w += o.w * cs;
          ac[c] = a3;
        }
        if (dx16) {
          uint2 w;
          if (zero16) {
            w.x = w.y = 0u;
          } else {
            const float rsv = dx16_row_scale ? dx16_row_scale[row] : 1.f;
            w.x = pack2_16(o.x * rsv, o.y * rsv, fmt);
            w.y = pack2_16(o.z * rsv, o.w * rsv, fmt);
          }
          reinterpret_cast<uint2*>(dx16 + row * lddx16)[c] = w;
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float sg = 0.f, sb = 0.f, sc = 0.f;
    for (int w = 0; w < nwarps; ++w) {
      sg += red[(w * 3 + 0) * d + c];
      sb += red[(w * 3 + 1) * d + c];
      sc += red[(w * 3 + 2) * d + c];
    }
    if (dgamma) atomicAdd(dgamma + c, sg * param_scale);
    if (dbeta) atomicAdd(dbeta + c, sb * param_scale);
    if (colsum) atomicAdd(colsum + c, sc * param_scale);   // bias gradient of the Linear this dx belongs to
  }
}

// LayerNorm bwd for large M (the TimeSformer stream: 50k rows x 768): same math and hooks as layernorm_bwd_kernel, built
// for memory-level parallelism instead of occupancy.
//   * ONE block of 12 warps per SM; every lane stages ITS OWN 16-byte chunks of the next row's x / dy / dx-accumulate
//     with cp.async into a private double buffer while it works on the current row (a lane only ever reads what it
//     copied itself, so cp.async.wait_group is the only synchronisation) -> ~100 KB of loads in flight per SM, where the
//     register version exposed a DRAM round trip per row (ncu r01i: long-scoreboard bound, 2.6 TB/s);
//   * dgamma / dbeta / bias-colsum partials live in registers (255 are available at this occupancy): no shared-memory
//     read-modify-write per element; one block-level reduction + fp32 atomics at the end.
constexpr int LNB_WARPS = 12;
template <int NV4>
__global__ void __launch_bounds__(LNB_WARPS * 32, 1)
layernorm_bwd_async_kernel(const void* __restrict__ dy, int dy_kind, long long lddy, const float* __restrict__ x,
                           long long ldx, const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float* __restrict__ gamma, long long M, int d, float* __restrict__ dx32, long long lddx,
                           int accumulate, uint16_t* __restrict__ dx16, long long lddx16, int fmt, int zero_period,
                           float* __restrict__ dgamma, float* __restrict__ dbeta, float param_scale,
                           float* __restrict__ colsum, int colsum_zero_period, const uint16_t* __restrict__ dy_mul16,
                           long long lddymul, const uint16_t* __restrict__ dx16_mul16, long long lddxmul,
                           const float* __restrict__ dx16_row_scale, const float* __restrict__ colsum_row_scale) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t lnb_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nv = d >> 2;
  constexpr int ARR = NV4 * 32;                                  // float4 slots per staged row
  float4* wbase = reinterpret_cast<float4*>(lnb_smem) + static_cast<size_t>(warp) * (2 * 3 * ARR);
  // stage s: [x | dy | dx]  (dy rows of 16-bit inputs use the first half of their slot array as uint2)
  auto sx = [&](int s) { return wbase + (s * 3 + 0) * ARR; };
  auto sdy = [&](int s) { return wbase + (s * 3 + 1) * ARR; };
  auto sdx = [&](int s) { return wbase + (s * 3 + 2) * ARR; };
  // gamma is the same for every row: one copy in shared memory behind the staging buffers (it was re-read through
  // __ldg twice per row and element: long-scoreboard stalls on L1/L2 hits, ncu r01p)
  float4* sgam = reinterpret_cast<float4*>(lnb_smem) + static_cast<size_t>(LNB_WARPS) * (2 * 3 * ARR);
  for (int c = threadIdx.x; c < nv; c += blockDim.x) sgam[c] = __ldg(reinterpret_cast<const float4*>(gamma) + c);
  __syncthreads();
  auto issue = [&](long long row, int s) {
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sx(s) + c)),
                     "l"(reinterpret_cast<const float4*>(x + row * ldx) + c)
                     : "memory");
        if (dy_kind == 0)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdy(s) + c)),
                       "l"(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * lddy) + c)
                       : "memory");
        else
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(reinterpret_cast<uint2*>(sdy(s)) + c)),
                       "l"(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(dy) + row * lddy) + c)
                       : "memory");
        if (accumulate)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdx(s) + c)),
                       "l"(reinterpret_cast<const float4*>(dx32 + row * lddx) + c)
                       : "memory");
        // the 16-bit multiplier rows are read straight from global when their row is processed: pull them into L2 now
        if (dx16_mul16)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint2*>(dx16_mul16 + row * lddxmul) + c));
        if (dy_mul16)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint2*>(dy_mul16 + row * lddymul) + c));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // per-row scalars of the NEXT row are fetched one iteration ahead (a dependent global load per row otherwise)
  struct RowScalars { float mu, rs, cs, rsv; };
  auto row_scalars = [&](long long row) {
    RowScalars r;
    r.mu = mean[row];
    r.rs = rstd[row];
    r.rsv = dx16_row_scale ? dx16_row_scale[row] : 1.f;
    r.cs = colsum_row_scale ? colsum_row_scale[row] : r.rsv;
    return r;
  };

  float4 ag[NV4], ab[NV4], ac[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) ag[i] = ab[i] = ac[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const long long stride = static_cast<long long>(gridDim.x) * LNB_WARPS;
  long long row = static_cast<long long>(blockIdx.x) * LNB_WARPS + warp;
  RowScalars cur{0.f, 0.f, 1.f, 1.f}, nxt{0.f, 0.f, 1.f, 1.f};
  if (row < M) {
    issue(row, 0);
    cur = row_scalars(row);
  }
  int s = 0;
  for (; row < M; row += stride, s ^= 1, cur = nxt) {
    const long long nrow = row + stride;
    if (nrow < M) {
      issue(nrow, s ^ 1);
      nxt = row_scalars(nrow);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    const float mu = cur.mu, rs = cur.rs;
    float4 xv[NV4], dv[NV4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        xv[i] = sx(s)[c];
        if (dy_kind == 0) {
          dv[i] = sdy(s)[c];
        } else {
          const uint2 w = reinterpret_cast<const uint2*>(sdy(s))[c];
          dv[i].x = f16_to_32(static_cast<uint16_t>(w.x & 0xffff), dy_kind - 1);
          dv[i].y = f16_to_32(static_cast<uint16_t>(w.x >> 16), dy_kind - 1);
          dv[i].z = f16_to_32(static_cast<uint16_t>(w.y & 0xffff), dy_kind - 1);
          dv[i].w = f16_to_32(static_cast<uint16_t>(w.y >> 16), dy_kind - 1);
        }
        if (dy_mul16) {   // upstream dropout on this LayerNorm's output
          const uint2 mw = reinterpret_cast<const uint2*>(dy_mul16 + row * lddymul)[c];
          dv[i].x *= f16_to_32(static_cast<uint16_t>(mw.x & 0xffff), fmt); dv[i].y *= f16_to_32(static_cast<uint16_t>(mw.x >> 16), fmt);
          dv[i].z *= f16_to_32(static_cast<uint16_t>(mw.y & 0xffff), fmt); dv[i].w *= f16_to_32(static_cast<uint16_t>(mw.y >> 16), fmt);
        }
        const float4 gm = sgam[c];
        const float hx = (xv[i].x - mu) * rs, hy = (xv[i].y - mu) * rs, hz = (xv[i].z - mu) * rs, hw = (xv[i].w - mu) * rs;
        const float gx = dv[i].x * gm.x, gy = dv[i].y * gm.y, gz = dv[i].z * gm.z, gw = dv[i].w * gm.w;
        s1 += (gx + gy) + (gz + gw);
        s2 += (gx * hx + gy * hy) + (gz * hz + gw * hw);
        ag[i].x += dv[i].x * hx; ag[i].y += dv[i].y * hy; ag[i].z += dv[i].z * hz; ag[i].w += dv[i].w * hw;
        ab[i].x += dv[i].x; ab[i].y += dv[i].y; ab[i].z += dv[i].z; ab[i].w += dv[i].w;
      }
    }
    const float c1 = warp_sum(s1) / d, c2 = warp_sum(s2) / d;
    const bool zero16 = zero_period > 0 && (row % zero_period) == 0;
    const bool cs_on = colsum != nullptr && !(colsum_zero_period > 0 && (row % colsum_zero_period) == 0);
    const float cs = cur.cs, rsv = cur.rsv;
    float4* dxrow = reinterpret_cast<float4*>(dx32 + row * lddx);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        const float4 gm = sgam[c];
        float4 o;
        o.x = rs * (dv[i].x * gm.x - c1 - (xv[i].x - mu) * rs * c2);
        o.y = rs * (dv[i].y * gm.y - c1 - (xv[i].y - mu) * rs * c2);
        o.z = rs * (dv[i].z * gm.z - c1 - (xv[i].z - mu) * rs * c2);
        o.w = rs * (dv[i].w * gm.w - c1 - (xv[i].w - mu) * rs * c2);
        if (accumulate) {
          const float4 pv = sdx(s)[c];
          o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
        }
        dxrow[c] = o;
        if (dx16_mul16) {
          const uint2 mw = reinterpret_cast<const uint2*>(dx16_mul16 + row * lddxmul)[c];
          o.x *= f16_to_32(static_cast<uint16_t>(mw.x & 0xffff), fmt); o.y *= f16_to_32(static_cast<uint16_t>(mw.x >> 16), fmt);
          o.z *= f16_to_32(static_cast<uint16_t>(mw.y & 0xffff), fmt); o.w *= f16_to_32(static_cast<uint16_t>(mw.y >> 16), fmt);
        }
        if (cs_on) {
          ac[i].x += o.x * cs; ac[i].y += o.y * cs; ac[i].z += o.z * cs; ac[i].w += o.w * cs;
        }
        if (dx16) {
          uint2 w;
          if (zero16) {
            w.x = w.y = 0u;
          } else {
            w.x = pack2_16(o.x * rsv, o.y * rsv, fmt);
            w.y = pack2_16(o.z * rsv, o.w * rsv, fmt);
          }
          reinterpret_cast<uint2*>(dx16 + row * lddx16)[c] = w;
        }
      }
    }
  }
  // block reduction of the per-warp register partials through the (now idle) staging memory: red[warp][3][d]
  __syncthreads();
  float* red = reinterpret_cast<float*>(lnb_smem);
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      reinterpret_cast<float4*>(red + (warp * 3 + 0) * d)[c] = ag[i];
      reinterpret_cast<float4*>(red + (warp * 3 + 1) * d)[c] = ab[i];
      reinterpret_cast<float4*>(red + (warp * 3 + 2) * d)[c] = ac[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float sg = 0.f, sb = 0.f, sc = 0.f;
    for (int w = 0; w < LNB_WARPS; ++w) {
      sg += red[(w * 3 + 0) * d + c];
      sb += red[(w * 3 + 1) * d + c];
      sc += red[(w * 3 + 2) * d + c];
    }
    if (dgamma) atomicAdd(dgamma + c, sg * param_scale);
    if (dbeta) atomicAdd(dbeta + c, sb * param_scale);
    if (colsum) atomicAdd(colsum + c, sc * param_scale);
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[n] += alpha * sum_m x[m,n]   (bias gradients). Each warp streams whole rows slices with 128-bit loads (a warp
// covers 256 columns of a row), 4 rows in flight; block = 8 warps on different rows, smem reduce, one atomic per column.
__global__ void __launch_bounds__(256) colsum16_kernel(const uint16_t* __restrict__ x, int fmt, long long ld,
                                                       long long M, int N, float* __restrict__ out, float alpha,
                                                       int zero_period) {
  pdl_grid_sync();
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const bool vec = (col + 8 <= N) && ((ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (col < N) {
    const long long stride = static_cast<long long>(gridDim.y) * 8;
    for (long long r0 = static_cast<long long>(blockIdx.y) * 8 + warp; r0 < M; r0 += stride * 4) {
      uint4 w[4];
      bool use[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long r = r0 + u * stride;
        use[u] = r < M && !(zero_period > 0 && (r % zero_period) == 0);
        w[u] = make_uint4(0u, 0u, 0u, 0u);
        if (use[u]) {
          if (vec) {
            w[u] = *reinterpret_cast<const uint4*>(x + r * ld + col);
          } else {
            uint16_t t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = (col + j < N) ? x[r * ld + col + j] : static_cast<uint16_t>(0);
            w[u].x = t[0] | (static_cast<uint32_t>(t[1]) << 16); w[u].y = t[2] | (static_cast<uint32_t>(t[3]) << 16);
            w[u].z = t[4] | (static_cast<uint32_t>(t[5]) << 16); w[u].w = t[6] | (static_cast<uint32_t>(t[7]) << 16);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t ww[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += f16_to_32(static_cast<uint16_t>(ww[j] & 0xffff), fmt);
          acc[2 * j + 1] += f16_to_32(static_cast<uint16_t>(ww[j] >> 16), fmt);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < N) {
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][c];
    atomicAdd(out + blockIdx.x * 256 + c, s * alpha);
  }
}

__global__ void colsum32_kernel(const float* __restrict__ x, long long ld, long long M, int N, float* __restrict__ out,
                                float alpha) {
  pdl_grid_sync();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float s = 0.f;
  for (long long r = blockIdx.y; r < M; r += gridDim.y) s += x[r * ld + col];
  atomicAdd(out + col, s * alpha);
}

// ------------------------------------------------------------------------------------------------ patch gather
// frames fp32 [B,T,3,H,W] -> 16-bit patch matrix [B*(1+N*T), 3*P*P] in the canonical token order
// (row b*(1+N*T) + 1 + n*T + t; row b*(1+N*T) is the clip's cls slot and is zero-filled). Column index
// c*P*P + ky*P + kx matches Conv2d weight [d,3,P,P].view(d,-1) (PatchEmbed, vit.py:230-238).
__global__ void patchify_kernel(const float* __restrict__ frames, uint16_t* __restrict__ out, int fmt, int B, int T,
                                int H, int W, int P) {
  pdl_grid_sync();
  const int gw = W / P, gh = H / P;
  const int N = gw * gh;
  const int Kd = 3 * P * P;
  const int k4 = Kd >> 2;
  const long long rows = static_cast<long long>(B) * (1 + N * T);
  const long long total = rows * k4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / k4;
    const int k = static_cast<int>(i - row * k4) * 4;
    const int b = static_cast<int>(row / (1 + N * T));
    const int j = static_cast<int>(row - static_cast<long long>(b) * (1 + N * T));
    uint2 w = make_uint2(0u, 0u);
    if (j > 0) {
      const int n = (j - 1) / T, t = (j - 1) - n * T;
      const int py = n / gw, px = n - py * gw;
      const int c = k / (P * P);
      const int rem = k - c * P * P;
      const int ky = rem / P, kx = rem - ky * P;
      const float4 v = *reinterpret_cast<const float4*>(
          frames + (((static_cast<long long>(b) * T + t) * 3 + c) * H + (py * P + ky)) * W + px * P + kx);
      w.x = pack2_16(v.x, v.y, fmt);
      w.y = pack2_16(v.z, v.w, fmt);
    }
    *reinterpret_cast<uint2*>(out + row * Kd + k) = w;
  }
}

// Same gather from RAW uint8 frames [B,T,3,H,W] with the input normalisation fused in:
// x = (u8 / 255 - mean[c]) / std[c]   (ImageNorm, src/datasets/data_utils.py:437-457). 4x fewer H2D / HBM bytes.
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ frames, uint16_t* __restrict__ out, int fmt, int B,
                                   int T, int H, int W, int P, float s0, float s1, float s2, float o0, float o1,
                                   float o2) {
  pdl_grid_sync();
  const int gw = W / P, gh = H / P;
  const int N = gw * gh;
  const int Kd = 3 * P * P;
  const int k4 = Kd >> 2;
  const long long rows = static_cast<long long>(B) * (1 + N * T);
  const long long total = rows * k4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / k4;
    const int k = static_cast<int>(i - row * k4) * 4;
    const int b = static_cast<int>(row / (1 + N * T));
    const int j = static_cast<int>(row - static_cast<long long>(b) * (1 + N * T));
    uint2 w = make_uint2(0u, 0u);
    if (j > 0) {
      const int n = (j - 1) / T, t = (j - 1) - n * T;
      const int py = n / gw, px = n - py * gw;
      const int c = k / (P * P);
      const int rem = k - c * P * P;
      const int ky = rem / P, kx = rem - ky * P;
      const uchar4 v = *reinterpret_cast<const uchar4*>(
          frames + (((static_cast<long long>(b) * T + t) * 3 + c) * H + (py * P + ky)) * W + px * P + kx);
      const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2), of = c == 0 ? o0 : (c == 1 ? o1 : o2);
      w.x = pack2_16(fmaf(v.x, sc, of), fmaf(v.y, sc, of), fmt);
      w.y = pack2_16(fmaf(v.z, sc, of), fmaf(v.w, sc, of), fmt);
    }
    *reinterpret_cast<uint2*>(out + row * Kd + k) = w;
  }
}

// x[b,0] = cls + pos[0];  x[b,1+n*T+t] = proj[b,1+n*T+t] + pos[1+n] + time[t]      (vit.py:324-361)
__global__ void vit_embed_fwd_kernel(const float* __restrict__ proj, const float* __restrict__ cls,
                                     const float* __restrict__ pos, const float* __restrict__ tim,
                                     float* __restrict__ x, int B, int N, int T, int d) {
  pdl_grid_sync();
  const int d4 = d >> 2;
  const long long total = static_cast<long long>(B) * (1 + N * T) * d4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / d4;
    const int c = static_cast<int>(i - row * d4);
    const int j = static_cast<int>(row % (1 + N * T));
    float4 o;
    if (j == 0) {
      const float4 a = reinterpret_cast<const float4*>(cls)[c];
      const float4 p = reinterpret_cast<const float4*>(pos)[c];
      o = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    } else {
      const int n = (j - 1) / T, t = (j - 1) - n * T;
      const float4 a = reinterpret_cast<const float4*>(proj)[i];
      const float4 p = reinterpret_cast<const float4*>(pos + static_cast<long long>(1 + n) * d)[c];
      const float4 q = reinterpret_cast<const float4*>(tim + static_cast<long long>(t) * d)[c];
      o = make_float4(a.x + p.x + q.x, a.y + p.y + q.y, a.z + p.z + q.z, a.w + p.w + q.w);
    }
    reinterpret_cast<float4*>(x)[i] = o;
  }
}

// Gradients of cls_token / pos_embed / time_embed: reductions of dx over (b,t), (b,n), b.
// grid = (1 + N + T output rows, B): each block reduces its clip's rows for one output row and adds it with atomics
// (outputs zero-initialised by the caller); threads over d.
__global__ void vit_embed_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dcls, float* __restrict__ dpos,
                                     float* __restrict__ dtim, int B, int N, int T, int d, float alpha) {
  pdl_grid_sync();
  const int which = blockIdx.x;
  const int b = blockIdx.y;
  const long long S = 1 + static_cast<long long>(N) * T;
  const float* base = dx + b * S * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f;
    if (which == 0) {
      s = base[c] * alpha;
      atomicAdd(dcls + c, s);
      atomicAdd(dpos + c, s);
    } else if (which <= N) {
      const int n = which - 1;
      for (int t = 0; t < T; ++t) s += base[(1 + static_cast<long long>(n) * T + t) * d + c];
      atomicAdd(dpos + static_cast<long long>(1 + n) * d + c, s * alpha);
    } else {
      const int t = which - 1 - N;
      for (int n = 0; n < N; ++n) s += base[(1 + static_cast<long long>(n) * T + t) * d + c];
      atomicAdd(dtim + static_cast<long long>(t) * d + c, s * alpha);
    }
  }
}

// video_embeds[b,0] = xn[b,0]; video_embeds[b,1+n] = mean_t xn[b,1+n*T+t]     (TimeSformer.forward_features :484-492)
__global__ void temporal_pool_fwd_kernel(const float* __restrict__ xn, float* __restrict__ out, int B, int N, int T,
                                         int d) {
  pdl_grid_sync();
  const int d4 = d >> 2;
  const long long total = static_cast<long long>(B) * (1 + N) * d4;
  const float inv = 1.f / T;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / d4;
    const int c = static_cast<int>(i - row * d4);
    const int b = static_cast<int>(row / (1 + N));
    const int j = static_cast<int>(row - static_cast<long long>(b) * (1 + N));
    const float4* src = reinterpret_cast<const float4*>(xn + (static_cast<long long>(b) * (1 + N * T)) * d) + c;
    float4 o;
    if (j == 0) {
      o = src[0];
    } else {
      o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int t = 0; t < T; ++t) {
        const float4 v = src[(1 + static_cast<long long>(j - 1) * T + t) * d4];
        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
      }
      o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

__global__ void temporal_pool_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dxn, int B, int N, int T,
                                         int d, float alpha) {
  pdl_grid_sync();
  const int d4 = d >> 2;
  const long long total = static_cast<long long>(B) * (1 + N * T) * d4;
  const float inv = alpha / T;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / d4;
    const int c = static_cast<int>(i - row * d4);
    const int b = static_cast<int>(row / (1 + N * T));
    const int j = static_cast<int>(row - static_cast<long long>(b) * (1 + N * T));
    const int jo = j == 0 ? 0 : 1 + (j - 1) / T;
    const float4 v = reinterpret_cast<const float4*>(dout + (static_cast<long long>(b) * (1 + N) + jo) * d)[c];
    const float s = j == 0 ? alpha : inv;
    reinterpret_cast<float4*>(dxn)[i] = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
  }
}

// ------------------------------------------------------------------------------------------------ BERT embeddings
// e[b,l] = word[ids[b,l]] + type[0] + pos[l]            (BertEmbeddings.forward xbert.py:186-210, before LayerNorm)
__global__ void bert_embed_gather_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                         const float* __restrict__ pos, const float* __restrict__ type,
                                         float* __restrict__ out, long long BL, int L, int h) {
  pdl_grid_sync();
  const int h4 = h >> 2;
  const long long total = BL * h4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long tok = i / h4;
    const int c = static_cast<int>(i - tok * h4);
    const int l = static_cast<int>(tok % L);
    const float4 a = reinterpret_cast<const float4*>(word + ids[tok] * h)[c];
    const float4 p = reinterpret_cast<const float4*>(pos + static_cast<long long>(l) * h)[c];
    const float4 t = reinterpret_cast<const float4*>(type)[c];
    reinterpret_cast<float4*>(out)[i] = make_float4(a.x + p.x + t.x, a.y + p.y + t.y, a.z + p.z + t.z, a.w + p.w + t.w);
  }
}

__global__ void bert_embed_scatter_kernel(const long long* __restrict__ ids, const float* __restrict__ de,
                                          float* __restrict__ dword, float* __restrict__ dpos,
                                          float* __restrict__ dtype, long long BL, int L, int h, float alpha) {
  pdl_grid_sync();
  const long long total = BL * h;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long tok = i / h;
    const int c = static_cast<int>(i - tok * h);
    const int l = static_cast<int>(tok % L);
    const float v = de[i] * alpha;
    atomicAdd(dword + ids[tok] * h + c, v);
    atomicAdd(dpos + static_cast<long long>(l) * h + c, v);
    atomicAdd(dtype + c, v);
  }
}

// ------------------------------------------------------------------------------------------------ fusion input
// out[s] = cat(text[ti[s]] (L rows), video[vi[s]] (Nv rows));  add_mask[s] = (1 - cat(tmask[ti[s]], 1)) * -10000
// (compute_vtm / compute_mlm embedding_output + get_extended_attention_mask; alpro_models.py:273-275,318-331,
//  xbert.py:878-938). Index lists let positives, hard negatives and the MLM pass share one batched fusion pass.
__global__ void fusion_gather_fwd_kernel(const float* __restrict__ text, const float* __restrict__ video,
                                         const long long* __restrict__ tmask, const int* __restrict__ ti,
                                         const int* __restrict__ vi, float* __restrict__ out32,
                                         uint16_t* __restrict__ out16, int fmt, float* __restrict__ add_mask, int S,
                                         int L, int Nv, int h) {
  pdl_grid_sync();
  const int h4 = h >> 2;
  const int R = L + Nv;
  const long long total = static_cast<long long>(S) * R * h4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / h4;
    const int c = static_cast<int>(i - row * h4);
    const int s = static_cast<int>(row / R);
    const int j = static_cast<int>(row - static_cast<long long>(s) * R);
    float4 v;
    if (j < L) {
      v = reinterpret_cast<const float4*>(text + (static_cast<long long>(ti[s]) * L + j) * h)[c];
      if (c == 0 && add_mask) add_mask[row] = (1.0f - static_cast<float>(tmask[static_cast<long long>(ti[s]) * L + j])) * -10000.0f;
    } else {
      v = reinterpret_cast<const float4*>(video + (static_cast<long long>(vi[s]) * Nv + (j - L)) * h)[c];
      if (c == 0 && add_mask) add_mask[row] = 0.f;
    }
    reinterpret_cast<float4*>(out32)[i] = v;
    if (out16) {
      uint2 w;
      w.x = pack2_16(v.x, v.y, fmt);
      w.y = pack2_16(v.z, v.w, fmt);
      reinterpret_cast<uint2*>(out16)[i] = w;
    }
  }
}

__global__ void fusion_gather_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ ti,
                                         const int* __restrict__ vi, float* __restrict__ dtext,
                                         float* __restrict__ dvideo, int S, int L, int Nv, int h) {
  pdl_grid_sync();
  const int R = L + Nv;
  const long long total = static_cast<long long>(S) * R * h;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / h;
    const int c = static_cast<int>(i - row * h);
    const int s = static_cast<int>(row / R);
    const int j = static_cast<int>(row - static_cast<long long>(s) * R);
    const float v = dout[i];
    if (j < L) atomicAdd(dtext + (static_cast<long long>(ti[s]) * L + j) * h + c, v);
    else atomicAdd(dvideo + (static_cast<long long>(vi[s]) * Nv + (j - L)) * h + c, v);
  }
}

// mean over t of the T per-frame cls rows -> the clip's canonical cls row (Block.forward vit.py:184-187, moved in front
// of the linear `proj`, with which the mean commutes).
__global__ void cls_mean_fwd_kernel(const uint16_t* __restrict__ cls_t, uint16_t* __restrict__ o, long long ldo,
                                    int fmt, int B, int T, int S, int d, const float* __restrict__ wgt) {
  pdl_grid_sync();
  // cls_t: [B, T, d]; o row b*S gets the mean
  const long long total = static_cast<long long>(B) * d;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / d);
    const int c = static_cast<int>(i - static_cast<long long>(b) * d);
    float s = 0.f;
    for (int t = 0; t < T; ++t)   // wgt: per-frame stochastic-depth factor mask/keep (vit.py:181-187 in train mode)
      s += f16_to_32(cls_t[(static_cast<long long>(b) * T + t) * d + c], fmt) * (wgt ? wgt[b * T + t] : 1.f);
    o[static_cast<long long>(b) * S * ldo + c] = f32_to_16(s / T, fmt);
  }
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

namespace alpro {
int layernorm_fwd_bulk(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int64_t M, int d,
                       float* out32, int64_t ld32, void* out16, int64_t ld16, int out16_fmt, float* mean, float* rstd,
                       const void* mul16, cudaStream_t st);
int layernorm_bwd_bulk(const void* dy, int dy_kind, int64_t lddy, const float* x, int64_t ldx, const float* mean,
                       const float* rstd, const float* gamma, int64_t M, int d, float* dx32, int64_t lddx, int accumulate,
                       void* dx16, int64_t lddx16, int dx16_fmt, int zero_period, float* dgamma, float* dbeta,
                       float param_scale, float* colsum, int colsum_zero_period, const void* dy_mul16,
                       const void* dx16_mul16, const float* dx16_row_scale, const float* colsum_row_scale,
                       cudaStream_t st);
}

extern "C" int alpro_cast_f32_to_16(const float* src, void* dst, int64_t n, int fmt, void* stream) {
  ALPRO_REQUIRE(src && dst && n >= 0, "alpro_cast_f32_to_16: bad args");
  if (n == 0) return 0;
  ALPRO_REQUIRE(aligned16(src) && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "alpro_cast_f32_to_16: alignment");
  launch_k(cast_f32_to_16_kernel, grid_for(cdiv(n, 4), 256), 256, 0, static_cast<cudaStream_t>(stream), 
      src, static_cast<uint16_t*>(dst), n, fmt);
  ALPRO_CHECK_LAUNCH("alpro_cast_f32_to_16");
  return 0;
}

extern "C" int alpro_cast_f32_to_16_multi(const int64_t* table, int num_chunks, int fmt, void* stream) {
  ALPRO_REQUIRE(table && num_chunks > 0, "alpro_cast_f32_to_16_multi: bad args");
  launch_k(cast_multi_kernel, static_cast<unsigned>(num_chunks), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const long long*>(table), fmt);
  ALPRO_CHECK_LAUNCH("alpro_cast_f32_to_16_multi");
  return 0;
}

extern "C" int alpro_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                                   int64_t M, int d, float* out32, int64_t ld32, void* out16, int64_t ld16,
                                   int out16_fmt, float* mean, float* rstd, const void* mul16, int64_t ldmul,
                                   void* stream) {
  ALPRO_REQUIRE(x && gamma && beta && M > 0, "alpro_layernorm_fwd: bad args");
  ALPRO_REQUIRE(d % 4 == 0 && d <= LN_MAX_V4 * 128, "alpro_layernorm_fwd: d=%d unsupported (multiple of 4, <= 1024)", d);
  ALPRO_REQUIRE(ldx % 4 == 0 && (!out32 || ld32 % 4 == 0) && (!out16 || ld16 % 4 == 0), "alpro_layernorm_fwd: ld");
  if (layernorm_fwd_bulk(x, ldx, gamma, beta, eps, M, d, out32, ld32, out16, ld16, out16_fmt, mean, rstd, mul16,
                         static_cast<cudaStream_t>(stream)) == ALPRO_OK) {
    ALPRO_CHECK_LAUNCH("alpro_layernorm_fwd(bulk)");
    return 0;
  }
  const int wpb = 8;
  launch_k(layernorm_fwd_kernel, static_cast<unsigned>(cdiv(M, wpb)), wpb * 32, 0, static_cast<cudaStream_t>(stream), 
      x, ldx, gamma, beta, eps, M, d, out32, ld32, static_cast<uint16_t*>(out16), ld16, out16_fmt, mean, rstd,
      static_cast<const uint16_t*>(mul16), ldmul);
  ALPRO_CHECK_LAUNCH("alpro_layernorm_fwd");
  return 0;
}

extern "C" int alpro_layernorm_bwd(const void* dy, int dy_kind, int64_t lddy, const float* x, int64_t ldx,
                                   const float* mean, const float* rstd, const float* gamma, int64_t M, int d,
                                   float* dx32, int64_t lddx, int accumulate, void* dx16, int64_t lddx16,
                                   int dx16_fmt, int zero_period, float* dgamma, float* dbeta, float param_scale,
                                   float* colsum, int colsum_zero_period, const void* dy_mul16, int64_t lddymul,
                                   const void* dx16_mul16, int64_t lddxmul, const float* dx16_row_scale,
                                   const float* colsum_row_scale, void* stream) {
  ALPRO_REQUIRE(dy && x && mean && rstd && gamma && dx32 && M > 0, "alpro_layernorm_bwd: bad args");
  ALPRO_REQUIRE(d % 4 == 0 && d <= LN_MAX_V4 * 128, "alpro_layernorm_bwd: d=%d unsupported", d);
  ALPRO_REQUIRE(dy_kind >= 0 && dy_kind <= 2, "alpro_layernorm_bwd: dy_kind");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // d = 768 token streams without multiplier rows: bulk-copy pipeline with packed fp32 math (layernorm_bulk.cu)
  if (layernorm_bwd_bulk(dy, dy_kind, lddy, x, ldx, mean, rstd, gamma, M, d, dx32, lddx, accumulate, dx16, lddx16, dx16_fmt,
                         zero_period, dgamma, dbeta, param_scale, colsum, colsum_zero_period, dy_mul16, dx16_mul16,
                         dx16_row_scale, colsum_row_scale, st) == ALPRO_OK) {
    ALPRO_CHECK_LAUNCH("alpro_layernorm_bwd(bulk)");
    return 0;
  }
  // Large row counts (the video token stream): one 12-warp block per SM with cp.async double buffering.
  // ALPRO_LN_BWD_ASYNC=0 keeps the register/occupancy version (read per call).
  {
    const char* e = getenv("ALPRO_LN_BWD_ASYNC");
    const bool use_async = !(e && e[0] == '0') && M >= 4096 && d <= 768 && aligned16(x) && aligned16(dx32) && (ldx % 4) == 0 &&
                           (lddx % 4) == 0 && aligned16(dy) && (lddy % 4) == 0;
    if (use_async) {
      const int nv4 = d <= 256 ? 2 : 6;
      const size_t smem_a = static_cast<size_t>(LNB_WARPS) * 2 * 3 * nv4 * 32 * sizeof(float4) +
                            static_cast<size_t>(nv4) * 32 * sizeof(float4);   // staging + one copy of gamma
      int grid_a = static_cast<int>(cdiv(M, LNB_WARPS));
      if (grid_a > num_sms()) grid_a = num_sms();
#define ALPRO_LN_BWD_A(NV)                                                                                             \
  do {                                                                                                                 \
    cudaFuncSetAttribute(layernorm_bwd_async_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,                  \
                         static_cast<int>(smem_a));                                                                    \
    launch_k(layernorm_bwd_async_kernel<NV>, grid_a, LNB_WARPS * 32, smem_a, st,                                             \
        dy, dy_kind, lddy, x, ldx, mean, rstd, gamma, M, d, dx32, lddx, accumulate, static_cast<uint16_t*>(dx16),      \
        lddx16, dx16_fmt, zero_period, dgamma, dbeta, param_scale, colsum, colsum_zero_period,                         \
        static_cast<const uint16_t*>(dy_mul16), lddymul, static_cast<const uint16_t*>(dx16_mul16), lddxmul,            \
        dx16_row_scale, colsum_row_scale);                                                                             \
  } while (0)
      if (nv4 == 2) ALPRO_LN_BWD_A(2);
      else ALPRO_LN_BWD_A(6);   // (d > 768 would need 288 KB of staging: stays on the register version)
#undef ALPRO_LN_BWD_A
      ALPRO_CHECK_LAUNCH("alpro_layernorm_bwd(async)");
      return 0;
    }
  }
  const int wpb = 6;
  int grid = static_cast<int>(cdiv(M, wpb));
  const int cap = num_sms() * 3;   // one resident wave (3 blocks/SM): fewest same-address atomics at the end
  if (grid > cap) grid = cap;
  const size_t smem = static_cast<size_t>(3) * wpb * d * sizeof(float);
#define ALPRO_LN_BWD(NV)                                                                                              \
  do {                                                                                                                \
  cudaFuncSetAttribute(layernorm_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));  \
  cudaFuncSetAttribute(layernorm_bwd_kernel<NV>, cudaFuncAttributePreferredSharedMemoryCarveout,                      \
                       cudaSharedmemCarveoutMaxShared);                                                                \
  launch_k(layernorm_bwd_kernel<NV>, grid, wpb * 32, smem, st, dy, dy_kind, lddy, x, ldx, mean, rstd, gamma, M, d, dx32, lddx, \
                                                         accumulate, static_cast<uint16_t*>(dx16), lddx16, dx16_fmt,   \
                                                         zero_period, dgamma, dbeta, param_scale, colsum,              \
                                                         colsum_zero_period, static_cast<const uint16_t*>(dy_mul16),   \
                                                         lddymul, static_cast<const uint16_t*>(dx16_mul16), lddxmul,   \
                                                         dx16_row_scale, colsum_row_scale);            \
  } while (0)
  if (d <= 256) ALPRO_LN_BWD(2);
  else if (d <= 768) ALPRO_LN_BWD(6);
  else ALPRO_LN_BWD(8);
#undef ALPRO_LN_BWD
  ALPRO_CHECK_LAUNCH("alpro_layernorm_bwd");
  return 0;
}

extern "C" int alpro_colsum(const void* x, int kind, int64_t ld, int64_t M, int N, float* out, float alpha,
                            int zero_period, void* stream) {
  ALPRO_REQUIRE(x && out && M > 0 && N > 0, "alpro_colsum: bad args");
  int rows = static_cast<int>(M < 512 ? M : 512);
  if (kind == 0) {
    dim3 grid(static_cast<unsigned>(cdiv(N, 128)), rows);
    launch_k(colsum32_kernel, grid, 128, 0, static_cast<cudaStream_t>(stream), static_cast<const float*>(x), ld, M, N, out,
                                                                          alpha);
  } else {
    int gy = static_cast<int>(cdiv(M, 8 * 16));   // >= 16 rows per warp
    const int cap = (num_sms() * 8) / static_cast<int>(cdiv(N, 256));
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    dim3 grid(static_cast<unsigned>(cdiv(N, 256)), gy);
    launch_k(colsum16_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), static_cast<const uint16_t*>(x), kind - 1, ld,
                                                                          M, N, out, alpha, zero_period);
  }
  ALPRO_CHECK_LAUNCH("alpro_colsum");
  return 0;
}

extern "C" int alpro_patchify(const float* frames, void* out16, int fmt, int B, int T, int H, int W, int P,
                              void* stream) {
  ALPRO_REQUIRE(frames && out16 && B > 0 && T > 0, "alpro_patchify: bad args");
  ALPRO_REQUIRE(P % 4 == 0 && H % P == 0 && W % P == 0 && W % 4 == 0, "alpro_patchify: P=%d H=%d W=%d unsupported", P, H, W);
  const long long total = static_cast<long long>(B) * (1 + (H / P) * (W / P) * T) * (3 * P * P / 4);
  launch_k(patchify_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      frames, static_cast<uint16_t*>(out16), fmt, B, T, H, W, P);
  ALPRO_CHECK_LAUNCH("alpro_patchify");
  return 0;
}

extern "C" int alpro_vit_embed_fwd(const float* proj, const float* cls, const float* pos, const float* tim, float* x,
                                   int B, int N, int T, int d, void* stream) {
  ALPRO_REQUIRE(proj && cls && pos && tim && x && d % 4 == 0, "alpro_vit_embed_fwd: bad args");
  const long long total = static_cast<long long>(B) * (1 + N * T) * (d / 4);
  launch_k(vit_embed_fwd_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), proj, cls, pos, tim, x, B,
                                                                                           N, T, d);
  ALPRO_CHECK_LAUNCH("alpro_vit_embed_fwd");
  return 0;
}

extern "C" int alpro_vit_embed_bwd(const float* dx, float* dcls, float* dpos, float* dtim, int B, int N, int T, int d,
                                   float alpha, void* stream) {
  ALPRO_REQUIRE(dx && dcls && dpos && dtim, "alpro_vit_embed_bwd: bad args");
  launch_k(vit_embed_bwd_kernel, dim3(1 + N + T, B), 256, 0, static_cast<cudaStream_t>(stream), dx, dcls, dpos, dtim, B, N,
                                                                                          T, d, alpha);
  ALPRO_CHECK_LAUNCH("alpro_vit_embed_bwd");
  return 0;
}

extern "C" int alpro_temporal_pool_fwd(const float* xn, float* out, int B, int N, int T, int d, void* stream) {
  ALPRO_REQUIRE(xn && out && d % 4 == 0, "alpro_temporal_pool_fwd: bad args");
  const long long total = static_cast<long long>(B) * (1 + N) * (d / 4);
  launch_k(temporal_pool_fwd_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), xn, out, B, N, T, d);
  ALPRO_CHECK_LAUNCH("alpro_temporal_pool_fwd");
  return 0;
}

extern "C" int alpro_temporal_pool_bwd(const float* dout, float* dxn, int B, int N, int T, int d, float alpha,
                                       void* stream) {
  ALPRO_REQUIRE(dout && dxn && d % 4 == 0, "alpro_temporal_pool_bwd: bad args");
  const long long total = static_cast<long long>(B) * (1 + N * T) * (d / 4);
  launch_k(temporal_pool_bwd_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), dout, dxn, B, N, T, d,
                                                                                               alpha);
  ALPRO_CHECK_LAUNCH("alpro_temporal_pool_bwd");
  return 0;
}

extern "C" int alpro_bert_embed_gather(const int64_t* ids, const float* word, const float* pos, const float* type,
                                       float* out, int64_t BL, int L, int h, void* stream) {
  ALPRO_REQUIRE(ids && word && pos && type && out && h % 4 == 0, "alpro_bert_embed_gather: bad args");
  launch_k(bert_embed_gather_kernel, grid_for(BL * (h / 4), 256), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const long long*>(ids), word, pos, type, out, BL, L, h);
  ALPRO_CHECK_LAUNCH("alpro_bert_embed_gather");
  return 0;
}

extern "C" int alpro_bert_embed_scatter(const int64_t* ids, const float* de, float* dword, float* dpos, float* dtype,
                                        int64_t BL, int L, int h, float alpha, void* stream) {
  ALPRO_REQUIRE(ids && de && dword && dpos && dtype, "alpro_bert_embed_scatter: bad args");
  launch_k(bert_embed_scatter_kernel, grid_for(BL * h, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const long long*>(ids), de, dword, dpos, dtype, BL, L, h, alpha);
  ALPRO_CHECK_LAUNCH("alpro_bert_embed_scatter");
  return 0;
}

extern "C" int alpro_fusion_gather_fwd(const float* text, const float* video, const int64_t* tmask, const int32_t* ti,
                                       const int32_t* vi, float* out32, void* out16, int fmt, float* add_mask, int S,
                                       int L, int Nv, int h, void* stream) {
  ALPRO_REQUIRE(text && video && tmask && ti && vi && out32 && h % 4 == 0, "alpro_fusion_gather_fwd: bad args");
  const long long total = static_cast<long long>(S) * (L + Nv) * (h / 4);
  launch_k(fusion_gather_fwd_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      text, video, reinterpret_cast<const long long*>(tmask), ti, vi, out32, static_cast<uint16_t*>(out16), fmt,
      add_mask, S, L, Nv, h);
  ALPRO_CHECK_LAUNCH("alpro_fusion_gather_fwd");
  return 0;
}

extern "C" int alpro_fusion_gather_bwd(const float* dout, const int32_t* ti, const int32_t* vi, float* dtext,
                                       float* dvideo, int S, int L, int Nv, int h, void* stream) {
  ALPRO_REQUIRE(dout && ti && vi && dtext && dvideo, "alpro_fusion_gather_bwd: bad args");
  const long long total = static_cast<long long>(S) * (L + Nv) * h;
  launch_k(fusion_gather_bwd_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), dout, ti, vi, dtext,
                                                                                               dvideo, S, L, Nv, h);
  ALPRO_CHECK_LAUNCH("alpro_fusion_gather_bwd");
  return 0;
}

extern "C" int alpro_cls_mean_fwd(const void* cls_t, void* o, int64_t ldo, int fmt, int B, int T, int S, int d,
                                  const float* frame_weight, void* stream) {
  ALPRO_REQUIRE(cls_t && o, "alpro_cls_mean_fwd: bad args");
  launch_k(cls_mean_fwd_kernel, grid_for(static_cast<long long>(B) * d, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const uint16_t*>(cls_t), static_cast<uint16_t*>(o), ldo, fmt, B, T, S, d, frame_weight);
  ALPRO_CHECK_LAUNCH("alpro_cls_mean_fwd");
  return 0;
}

extern "C" int alpro_dropout_mask(void* out16, int fmt, int64_t n, float p, uint32_t seed, void* stream) {
  ALPRO_REQUIRE(out16 && n > 0 && p >= 0.f && p < 1.f, "alpro_dropout_mask: bad args");
  launch_k(dropout_mask_kernel, grid_for((n + 1) / 2, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<uint16_t*>(out16), fmt, n, drop_threshold(p), 1.0f / (1.0f - p), seed);
  ALPRO_CHECK_LAUNCH("alpro_dropout_mask");
  return 0;
}

extern "C" int alpro_patchify_u8(const uint8_t* frames, void* out16, int fmt, int B, int T, int H, int W, int P,
                                 const float* mean3, const float* std3, void* stream) {
  ALPRO_REQUIRE(frames && out16 && mean3 && std3 && B > 0 && T > 0, "alpro_patchify_u8: bad args (mean3/std3 are HOST pointers)");
  ALPRO_REQUIRE(P % 4 == 0 && H % P == 0 && W % P == 0 && W % 4 == 0, "alpro_patchify_u8: P=%d H=%d W=%d unsupported", P, H, W);
  float sc[3], of[3];
  for (int c = 0; c < 3; ++c) {
    sc[c] = 1.0f / (255.0f * std3[c]);
    of[c] = -mean3[c] / std3[c];
  }
  const long long total = static_cast<long long>(B) * (1 + (H / P) * (W / P) * T) * (3 * P * P / 4);
  launch_k(patchify_u8_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      frames, static_cast<uint16_t*>(out16), fmt, B, T, H, W, P, sc[0], sc[1], sc[2], of[0], of[1], of[2]);
  ALPRO_CHECK_LAUNCH("alpro_patchify_u8");
  return 0;
}
