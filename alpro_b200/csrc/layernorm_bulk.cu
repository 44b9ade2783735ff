// LayerNorm backward for the big token streams (d = 768, M >= 4096), third generation.
//
// The cp.async version (elementwise.cu) moves 4.6 TB/s at 996 W: ncu shows 1400 warp instructions per row of which
// half are address arithmetic, predicates and divergence bookkeeping for its per-lane 16-byte copies, and the board
// power limit, not HBM, sets its pace (profiles/r02q_energy_probe.md: 173 pJ per byte against 85 for a plain copy).
// This kernel does the same math with ~4x fewer instructions:
//   * rows arrive by 1-D bulk copies (cp.async.bulk + mbarrier complete_tx): lane 0 issues THREE instructions per row
//     (x, dy, dx-accumulate) where 32 lanes issued 18 predicated copies each; three stages per warp keep two rows
//     (15 KB) per warp = 123 KB per SM in flight;
//   * d is a template constant (six float4 per lane): no column predicates, gamma lives in registers;
//   * the fp32 arithmetic is packed (fma/mul/add.rn.f32x2 = FFMA2/FMUL2/FADD2, one issue slot per two lanes of a
//     float4), x-hat and gamma*dy are kept in registers between the statistics pass and the output pass;
//   * row-period tests (cls rows) are carried incrementally instead of a 64-bit modulo per row.
// Reference semantics: nn.LayerNorm of the TimeSformer blocks (vit.py:140-154, eps 1e-6) and of BertSelfOutput /
// BertOutput (xbert.py:349-360, 425-438, eps 1e-12): y = (x - mean) * rstd * gamma + beta with the biased variance;
// backward dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
// Same contract as layernorm_bwd_async_kernel except for the 16-bit multiplier rows (BERT hidden-dropout masks), which
// stay on the older kernel. ALPRO_LN_BWD_BULK=0 disables it (read per call).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace alpro {
namespace lnbulk {

constexpr int NV4 = 6;              // float4 per lane: d = 768
constexpr int D = NV4 * 128;
constexpr int WARPS = 8;
constexpr int NS = 3;               // stages per warp
constexpr int X_BYTES = D * 4;      // 3072

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// packed fp32 pairs
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk(f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

struct Params {
  const void* dy;          // [M, d] fp32 (DYK 0) or 16-bit (DYK 1 = fp16, 2 = bf16)
  long long lddy;
  const float* x;
  long long ldx;
  const float* mean;
  const float* rstd;
  const float* gamma;
  long long M;
  float* dx32;
  long long lddx;
  int accumulate;
  uint16_t* dx16;
  long long lddx16;
  int fmt;                 // dx16 format: 0 fp16, 1 bf16
  int zero_period;
  float* dgamma;
  float* dbeta;
  float param_scale;
  float* colsum;
  int colsum_zero_period;
  const float* dx16_row_scale;
  const float* colsum_row_scale;
};

template <int DYK>
struct Stage {
  static constexpr int DY_BYTES = DYK == 0 ? D * 4 : D * 2;
  static constexpr int BYTES = 2 * X_BYTES + DY_BYTES;   // x | dx | dy
};

template <int DYK>
__global__ void __launch_bounds__(WARPS * 32, 1) layernorm_bwd_bulk_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int SB = Stage<DYK>::BYTES;
  constexpr int DYB = Stage<DYK>::DY_BYTES;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint8_t* wbase = smem + static_cast<size_t>(warp) * NS * SB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(WARPS) * NS * SB) + warp * NS;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncwarp();
  pdl_grid_sync();

  const long long stride = static_cast<long long>(gridDim.x) * WARPS;
  const long long row0 = static_cast<long long>(blockIdx.x) * WARPS + warp;
  const uint32_t tx_bytes = static_cast<uint32_t>(X_BYTES + DYB + (p.accumulate ? X_BYTES : 0));
  auto issue = [&](long long row, int s) {   // lane 0 only
    uint8_t* st = wbase + s * SB;
    mbar_arrive_expect_tx(&bars[s], tx_bytes);
    bulk_load_1d(st, p.x + row * p.ldx, X_BYTES, &bars[s]);
    if (p.accumulate) bulk_load_1d(st + X_BYTES, p.dx32 + row * p.lddx, X_BYTES, &bars[s]);
    if (DYK == 0) bulk_load_1d(st + 2 * X_BYTES, reinterpret_cast<const float*>(p.dy) + row * p.lddy, DYB, &bars[s]);
    else bulk_load_1d(st + 2 * X_BYTES, reinterpret_cast<const uint16_t*>(p.dy) + row * p.lddy, DYB, &bars[s]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS - 1; ++s)
      if (row0 + s * stride < p.M) issue(row0 + s * stride, s);
  }

  f2 gam[NV4][2];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + lane + 32 * i);
    gam[i][0] = pk(g.x, g.y);
    gam[i][1] = pk(g.z, g.w);
  }
  f2 ag[NV4][2], ab[NV4][2], ac[NV4][2];
#pragma unroll
  for (int i = 0; i < NV4; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) ag[i][h] = ab[i][h] = ac[i][h] = 0ull;   // (+0.f, +0.f)

  struct RowScalars { float mu, rs, cs, rsv; };
  auto row_scalars = [&](long long row) {
    RowScalars r;
    r.mu = __ldg(p.mean + row);
    r.rs = __ldg(p.rstd + row);
    r.rsv = p.dx16_row_scale ? __ldg(p.dx16_row_scale + row) : 1.f;
    r.cs = p.colsum_row_scale ? __ldg(p.colsum_row_scale + row) : r.rsv;
    return r;
  };
  // row % period, carried incrementally (row advances by `stride` per iteration)
  int zr = 0, zstep = 0, cr = 0, cstep = 0;
  if (p.zero_period > 0) {
    zr = static_cast<int>(row0 % p.zero_period);
    zstep = static_cast<int>(stride % p.zero_period);
  }
  if (p.colsum_zero_period > 0) {
    cr = static_cast<int>(row0 % p.colsum_zero_period);
    cstep = static_cast<int>(stride % p.colsum_zero_period);
  }

  RowScalars cur{0.f, 0.f, 1.f, 1.f}, nxt{0.f, 0.f, 1.f, 1.f};
  if (row0 < p.M) cur = row_scalars(row0);
  int s = 0;
  uint32_t parity = 0;
  for (long long row = row0; row < p.M; row += stride, cur = nxt) {
    {
      const long long prow = row + (NS - 1) * stride;   // refill the stage consumed in the previous iteration
      int ps = s + NS - 1;
      if (ps >= NS) ps -= NS;
      if (lane == 0 && prow < p.M) issue(prow, ps);
      const long long nrow = row + stride;
      if (nrow < p.M) nxt = row_scalars(nrow);
    }
    mbar_wait(&bars[s], parity);
    const uint8_t* st = wbase + s * SB;
    const float4* sx = reinterpret_cast<const float4*>(st);
    const float4* sdx = reinterpret_cast<const float4*>(st + X_BYTES);
    const float rs = cur.rs;
    const f2 rs2 = pk(rs, rs), nmu = pk(-cur.mu, -cur.mu);
    f2 xh[NV4][2], gd[NV4][2];
    f2 s1 = 0ull, s2 = 0ull;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      const float4 xv = sx[c];
      f2 d0, d1;
      if (DYK == 0) {
        const float4 dv = reinterpret_cast<const float4*>(st + 2 * X_BYTES)[c];
        d0 = pk(dv.x, dv.y);
        d1 = pk(dv.z, dv.w);
      } else {
        const uint2 w = reinterpret_cast<const uint2*>(st + 2 * X_BYTES)[c];
        if (DYK == 1) {
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
          const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
          d0 = pk(a.x, a.y);
          d1 = pk(b.x, b.y);
        } else {
          d0 = pk(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u));
          d1 = pk(__uint_as_float(w.y << 16), __uint_as_float(w.y & 0xffff0000u));
        }
      }
      xh[i][0] = mul2(add2(pk(xv.x, xv.y), nmu), rs2);   // (x - mu) * rs: the difference first, no cancellation
      xh[i][1] = mul2(add2(pk(xv.z, xv.w), nmu), rs2);
      gd[i][0] = mul2(d0, gam[i][0]);
      gd[i][1] = mul2(d1, gam[i][1]);
      s1 = add2(s1, add2(gd[i][0], gd[i][1]));
      s2 = fma2(gd[i][0], xh[i][0], s2);
      s2 = fma2(gd[i][1], xh[i][1], s2);
      ag[i][0] = fma2(d0, xh[i][0], ag[i][0]);
      ag[i][1] = fma2(d1, xh[i][1], ag[i][1]);
      ab[i][0] = add2(ab[i][0], d0);
      ab[i][1] = add2(ab[i][1], d1);
    }
    float s1a, s1b, s2a, s2b;
    upk(s1, s1a, s1b);
    upk(s2, s2a, s2b);
    float t1 = s1a + s1b, t2 = s2a + s2b;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    // dx = rs * (g - c1 - xh * c2) = fma(xh, -rs*c2, fma(g, rs, -rs*c1))
    const float c1 = t1 * (1.f / D), c2 = t2 * (1.f / D);
    const f2 k1 = pk(-rs * c1, -rs * c1), k2 = pk(-rs * c2, -rs * c2);
    const bool zero16 = p.zero_period > 0 && zr == 0;
    const bool cs_on = p.colsum != nullptr && !(p.colsum_zero_period > 0 && cr == 0);
    const f2 cs2 = cs_on ? pk(cur.cs, cur.cs) : 0ull;
    const f2 rsv2 = pk(cur.rsv, cur.rsv);
    float4* dxrow = reinterpret_cast<float4*>(p.dx32 + row * p.lddx);
    uint2* d16row = p.dx16 ? reinterpret_cast<uint2*>(p.dx16 + row * p.lddx16) : nullptr;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      f2 o0 = fma2(xh[i][0], k2, fma2(gd[i][0], rs2, k1));
      f2 o1 = fma2(xh[i][1], k2, fma2(gd[i][1], rs2, k1));
      if (p.accumulate) {
        const float4 pv = sdx[c];
        o0 = add2(o0, pk(pv.x, pv.y));
        o1 = add2(o1, pk(pv.z, pv.w));
      }
      float4 o;
      upk(o0, o.x, o.y);
      upk(o1, o.z, o.w);
      dxrow[c] = o;
      ac[i][0] = fma2(o0, cs2, ac[i][0]);
      ac[i][1] = fma2(o1, cs2, ac[i][1]);
      if (d16row) {
        uint2 w;
        if (zero16) {
          w.x = w.y = 0u;
        } else {
          float a, b, cc, dd;
          upk(mul2(o0, rsv2), a, b);
          upk(mul2(o1, rsv2), cc, dd);
          if (p.fmt) {
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.x) : "f"(b), "f"(a));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.y) : "f"(dd), "f"(cc));
          } else {
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w.x) : "f"(b), "f"(a));
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w.y) : "f"(dd), "f"(cc));
          }
        }
        d16row[c] = w;
      }
    }
    __syncwarp();   // every lane has read its part of stage s: lane 0 may refill it in the next iteration
    if (++s == NS) { s = 0; parity ^= 1; }
    if (p.zero_period > 0) { zr += zstep; if (zr >= p.zero_period) zr -= p.zero_period; }
    if (p.colsum_zero_period > 0) { cr += cstep; if (cr >= p.colsum_zero_period) cr -= p.colsum_zero_period; }
  }

  // block reduction of the per-warp register partials through the (now idle) staging memory: red[warp][3][D]
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + 32 * i;
    float4 v;
    upk(ag[i][0], v.x, v.y); upk(ag[i][1], v.z, v.w);
    reinterpret_cast<float4*>(red + (warp * 3 + 0) * D)[c] = v;
    upk(ab[i][0], v.x, v.y); upk(ab[i][1], v.z, v.w);
    reinterpret_cast<float4*>(red + (warp * 3 + 1) * D)[c] = v;
    upk(ac[i][0], v.x, v.y); upk(ac[i][1], v.z, v.w);
    reinterpret_cast<float4*>(red + (warp * 3 + 2) * D)[c] = v;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sg = 0.f, sb = 0.f, sc = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      sg += red[(w * 3 + 0) * D + c];
      sb += red[(w * 3 + 1) * D + c];
      sc += red[(w * 3 + 2) * D + c];
    }
    if (p.dgamma) atomicAdd(p.dgamma + c, sg * p.param_scale);
    if (p.dbeta) atomicAdd(p.dbeta + c, sb * p.param_scale);
    if (p.colsum) atomicAdd(p.colsum + c, sc * p.param_scale);
  }
}

template <int DYK>
static int launch(const Params& p, cudaStream_t st) {
  constexpr size_t smem = static_cast<size_t>(WARPS) * NS * Stage<DYK>::BYTES + WARPS * NS * sizeof(uint64_t);
  static_assert(smem <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
  static_assert(static_cast<size_t>(WARPS) * 3 * D * sizeof(float) <= static_cast<size_t>(WARPS) * NS * Stage<DYK>::BYTES,
                "reduction scratch must fit the staging memory");
  cudaFuncSetAttribute(layernorm_bwd_bulk_kernel<DYK>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int grid = static_cast<int>(cdiv(p.M, WARPS));
  if (grid > num_sms()) grid = num_sms();
  launch_k(layernorm_bwd_bulk_kernel<DYK>, grid, WARPS * 32, smem, st, p);
  return 0;
}

// ------------------------------------------------------------------------------------------------ forward
// Same pipeline for the forward: one bulk copy per row, two-pass statistics on the register copy (torch's biased
// variance), gamma / beta in registers, packed arithmetic, 16-bit and/or fp32 output straight from registers.
constexpr int FW_WARPS = 8;
constexpr int FW_NS = 4;

struct FwdParams {
  const float* x;
  long long ldx;
  const float* gamma;
  const float* beta;
  float eps;
  long long M;
  float* out32;
  long long ld32;
  uint16_t* out16;
  long long ld16;
  int fmt;
  float* mean;
  float* rstd;
};

__global__ void __launch_bounds__(FW_WARPS * 32, 2) layernorm_fwd_bulk_kernel(const FwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint8_t* wbase = smem + static_cast<size_t>(warp) * FW_NS * X_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(FW_WARPS) * FW_NS * X_BYTES) + warp * FW_NS;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < FW_NS; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncwarp();
  pdl_grid_sync();
  const long long stride = static_cast<long long>(gridDim.x) * FW_WARPS;
  const long long row0 = static_cast<long long>(blockIdx.x) * FW_WARPS + warp;
  auto issue = [&](long long row, int s) {   // lane 0 only
    mbar_arrive_expect_tx(&bars[s], X_BYTES);
    bulk_load_1d(wbase + s * X_BYTES, p.x + row * p.ldx, X_BYTES, &bars[s]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < FW_NS - 1; ++s)
      if (row0 + s * stride < p.M) issue(row0 + s * stride, s);
  }
  f2 gam[NV4][2], bet[NV4][2];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + lane + 32 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta) + lane + 32 * i);
    gam[i][0] = pk(g.x, g.y); gam[i][1] = pk(g.z, g.w);
    bet[i][0] = pk(b.x, b.y); bet[i][1] = pk(b.z, b.w);
  }
  int s = 0;
  uint32_t parity = 0;
  for (long long row = row0; row < p.M; row += stride) {
    {
      const long long prow = row + (FW_NS - 1) * stride;
      int ps = s + FW_NS - 1;
      if (ps >= FW_NS) ps -= FW_NS;
      if (lane == 0 && prow < p.M) issue(prow, ps);
    }
    mbar_wait(&bars[s], parity);
    const float4* sx = reinterpret_cast<const float4*>(wbase + s * X_BYTES);
    f2 v[NV4][2];
    f2 acc = 0ull;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const float4 xv = sx[lane + 32 * i];
      v[i][0] = pk(xv.x, xv.y);
      v[i][1] = pk(xv.z, xv.w);
      acc = add2(acc, add2(v[i][0], v[i][1]));
    }
    __syncwarp();   // stage s is in registers: lane 0 may refill it in the next iteration
    float a, b;
    upk(acc, a, b);
    const float mean = warp_sum(a + b) * (1.f / D);
    const f2 nm = pk(-mean, -mean);
    f2 q = 0ull;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      v[i][0] = add2(v[i][0], nm);
      v[i][1] = add2(v[i][1], nm);
      q = fma2(v[i][0], v[i][0], q);
      q = fma2(v[i][1], v[i][1], q);
    }
    upk(q, a, b);
    const float rstd = rsqrtf(warp_sum(a + b) * (1.f / D) + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[row] = mean;
      if (p.rstd) p.rstd[row] = rstd;
    }
    const f2 r2 = pk(rstd, rstd);
    float4* o32 = p.out32 ? reinterpret_cast<float4*>(p.out32 + row * p.ld32) : nullptr;
    uint2* o16 = p.out16 ? reinterpret_cast<uint2*>(p.out16 + row * p.ld16) : nullptr;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      float4 y;
      upk(fma2(mul2(v[i][0], r2), gam[i][0], bet[i][0]), y.x, y.y);
      upk(fma2(mul2(v[i][1], r2), gam[i][1], bet[i][1]), y.z, y.w);
      if (o32) o32[c] = y;
      if (o16) {
        uint2 w;
        if (p.fmt) {
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.x) : "f"(y.y), "f"(y.x));
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.y) : "f"(y.w), "f"(y.z));
        } else {
          asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w.x) : "f"(y.y), "f"(y.x));
          asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w.y) : "f"(y.w), "f"(y.z));
        }
        o16[c] = w;
      }
    }
    if (++s == FW_NS) { s = 0; parity ^= 1; }
  }
}

}  // namespace lnbulk

// Forward counterpart of layernorm_bwd_bulk (same return convention).
int layernorm_fwd_bulk(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int64_t M, int d,
                       float* out32, int64_t ld32, void* out16, int64_t ld16, int out16_fmt, float* mean, float* rstd,
                       const void* mul16, cudaStream_t st) {
  const char* e = getenv("ALPRO_LN_FWD_BULK");
  if (e && e[0] == '0') return ALPRO_ENOTSUP;
  if (d != lnbulk::D || M < 4096 || mul16) return ALPRO_ENOTSUP;
  if (!aligned16(x) || !aligned16(gamma) || !aligned16(beta) || (ldx % 4) || (out32 && (!aligned16(out32) || (ld32 % 4))) ||
      (out16 && ((reinterpret_cast<uintptr_t>(out16) & 7) || (ld16 % 4))))
    return ALPRO_ENOTSUP;
  lnbulk::FwdParams p;
  p.x = x; p.ldx = ldx; p.gamma = gamma; p.beta = beta; p.eps = eps; p.M = M; p.out32 = out32; p.ld32 = ld32;
  p.out16 = static_cast<uint16_t*>(out16); p.ld16 = ld16; p.fmt = out16_fmt; p.mean = mean; p.rstd = rstd;
  constexpr size_t smem = static_cast<size_t>(lnbulk::FW_WARPS) * lnbulk::FW_NS * lnbulk::X_BYTES +
                          lnbulk::FW_WARPS * lnbulk::FW_NS * sizeof(uint64_t);
  static_assert(2 * (smem + 1024) <= 232448, "two blocks per SM");
  cudaFuncSetAttribute(lnbulk::layernorm_fwd_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int grid = static_cast<int>(cdiv(M, lnbulk::FW_WARPS));
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  launch_k(lnbulk::layernorm_fwd_bulk_kernel, grid, lnbulk::FW_WARPS * 32, smem, st, p);
  return ALPRO_OK;
}

// Returns ALPRO_OK when the bulk kernel took the problem, ALPRO_ENOTSUP when the caller should use another kernel.
int layernorm_bwd_bulk(const void* dy, int dy_kind, int64_t lddy, const float* x, int64_t ldx, const float* mean,
                       const float* rstd, const float* gamma, int64_t M, int d, float* dx32, int64_t lddx, int accumulate,
                       void* dx16, int64_t lddx16, int dx16_fmt, int zero_period, float* dgamma, float* dbeta,
                       float param_scale, float* colsum, int colsum_zero_period, const void* dy_mul16,
                       const void* dx16_mul16, const float* dx16_row_scale, const float* colsum_row_scale,
                       cudaStream_t st) {
  const char* e = getenv("ALPRO_LN_BWD_BULK");
  if (e && e[0] == '0') return ALPRO_ENOTSUP;
  if (d != lnbulk::D || M < 4096 || dy_mul16 || dx16_mul16) return ALPRO_ENOTSUP;
  // bulk copies: 16-byte aligned row starts
  const int dy_elt = dy_kind == 0 ? 4 : 2;
  if (!aligned16(x) || !aligned16(dx32) || !aligned16(dy) || !aligned16(gamma) || (ldx % 4) || (lddx % 4) ||
      ((lddy * dy_elt) % 16) || (dx16 && ((reinterpret_cast<uintptr_t>(dx16) & 7) || (lddx16 % 4))))
    return ALPRO_ENOTSUP;
  lnbulk::Params p;
  p.dy = dy; p.lddy = lddy; p.x = x; p.ldx = ldx; p.mean = mean; p.rstd = rstd; p.gamma = gamma; p.M = M;
  p.dx32 = dx32; p.lddx = lddx; p.accumulate = accumulate; p.dx16 = static_cast<uint16_t*>(dx16); p.lddx16 = lddx16;
  p.fmt = dx16_fmt; p.zero_period = zero_period; p.dgamma = dgamma; p.dbeta = dbeta; p.param_scale = param_scale;
  p.colsum = colsum; p.colsum_zero_period = colsum_zero_period; p.dx16_row_scale = dx16_row_scale;
  p.colsum_row_scale = colsum_row_scale;
  if (dy_kind == 0) return lnbulk::launch<0>(p, st);
  if (dy_kind == 1) return lnbulk::launch<1>(p, st);
  return lnbulk::launch<2>(p, st);
}

}  // namespace alpro
