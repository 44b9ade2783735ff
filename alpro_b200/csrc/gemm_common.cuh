// Shared pieces of the tcgen05 GEMM kernels (1-CTA gemm_tc.cu and 2-CTA gemm_tc2.cu): kernel parameters, the
// compile-time epilogue modes and the coalesced-side fused epilogue.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace alpro {
namespace gemm {

struct GemmKParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  int split_k, kb_per_split;
  int a_mn, b_mn;
  uint32_t idesc;
  // epilogue
  const float* bias;
  const uint16_t* aux16;
  const float* resid;
  float* out32;
  uint16_t* out16;
  uint16_t* out16b;
  long long ld32, ld16, ld16b, ldresid, ldaux;
  int out16_fmt, out16b_fmt, aux_fmt;
  int act;
  int skip_period;
  const float* rs_acc;   // optional per-row scale of (alpha*acc)   [M]  (drop_path: mask/keep of the row's sample)
  const float* rs_bias;  // optional per-row scale of the bias term  [M]  (defaults to rs_acc)
  const float* bias2;    // optional second, never row-scaled bias [N] (residual modes; skip rows keep the residual)
  int vec_ok;  // all leading dims / pointers allow vector accesses on 4-column groups
  float alpha;
};

// Compile-time epilogue modes (each is one kernel instantiation; E_GENERIC keeps every flag at run time).
enum EpiMode {
  E_OUT16 = 0,        // out16 = acc*alpha + bias
  E_GELU_SAVE = 1,    // out16b = pre-activation, out16 = gelu(pre)
  E_GELU_GRAD = 2,    // out16 = acc*alpha * gelu'(aux)
  E_RESID_OUT32 = 3,  // out32 = acc*alpha + bias + resid (skip_period rows: resid only); optional out16 copy
  E_OUT32 = 4,        // out32 = acc*alpha + bias
  E_ATOMIC = 5,       // out32 += acc*alpha   (split-K)
  E_GENERIC = 6
};

__device__ __forceinline__ void store4_16(uint16_t* o, const float (&v)[4], int fmt, bool full, int ncol) {
  if (full) {
    uint2 w;
    w.x = pack2_16(v[0], v[1], fmt);
    w.y = pack2_16(v[2], v[3], fmt);
    *reinterpret_cast<uint2*>(o) = w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < ncol) o[j] = f32_to_16(v[j], fmt);
  }
}
__device__ __forceinline__ void load4_16(const uint16_t* a, float (&u)[4], int fmt, bool full, int ncol) {
  if (full) {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(a));
    u[0] = f16_to_32(static_cast<uint16_t>(w.x & 0xffff), fmt);
    u[1] = f16_to_32(static_cast<uint16_t>(w.x >> 16), fmt);
    u[2] = f16_to_32(static_cast<uint16_t>(w.y & 0xffff), fmt);
    u[3] = f16_to_32(static_cast<uint16_t>(w.y >> 16), fmt);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = (j < ncol) ? f16_to_32(a[j], fmt) : 0.f;
  }
}
__device__ __forceinline__ void load4_32(const float* r, float (&u)[4], bool full, int ncol) {
  if (full) {
    const float4 t = *reinterpret_cast<const float4*>(r);
    u[0] = t.x; u[1] = t.y; u[2] = t.z; u[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = (j < ncol) ? r[j] : 0.f;
  }
}
__device__ __forceinline__ void store4_32(float* o, const float (&v)[4], bool full, int ncol) {
  if (full) {
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < ncol) o[j] = v[j];
  }
}

// Epilogue for NR rows x 4 consecutive columns held by one lane on the coalesced side. All global loads of the group
// are issued before any store so that they overlap (the compiler cannot hoist them itself: resid may alias out), and
// the math is unconditional straight-line code (row indices are clamped for the loads, only the stores are predicated)
// so that the NR*4 independent element chains interleave.
template <int MODE, int NR>
__device__ __forceinline__ void epilogue_rows(const GemmKParams& p, float (&v)[NR][4], const float4& b4,
                                              const long long (&row)[NR], const bool (&ok)[NR], int col) {
  const int ncol = min(4, p.N - col);
  const bool full = p.vec_ok && ncol == 4;
  if (MODE == E_ATOMIC) {
#pragma unroll
    for (int i = 0; i < NR; ++i)
      if (ok[i]) {
        float* o = p.out32 + row[i] * p.ld32 + col;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < ncol) atomicAdd(o + j, v[i][j]);  // result unused -> RED.ADD.F32
      }
    return;
  }
  const int act = (MODE == E_GENERIC) ? p.act
                  : (MODE == E_GELU_SAVE) ? ALPRO_ACT_GELU
                  : (MODE == E_GELU_GRAD) ? ALPRO_ACT_GELU_GRAD : ALPRO_ACT_NONE;
  const bool has_resid = (MODE == E_GENERIC) ? (p.resid != nullptr) : (MODE == E_RESID_OUT32);
  const bool has_out32 = (MODE == E_GENERIC) ? (p.out32 != nullptr) : (MODE == E_RESID_OUT32 || MODE == E_OUT32);
  const bool has_out16 = (MODE == E_GENERIC || MODE == E_RESID_OUT32) ? (p.out16 != nullptr)
                                                                       : (MODE != E_OUT32);
  const bool has_out16b = (MODE == E_GENERIC || MODE == E_GELU_SAVE) ? (p.out16b != nullptr) : false;
  long long lrow[NR];  // clamped row for loads (always in bounds)
#pragma unroll
  for (int i = 0; i < NR; ++i) lrow[i] = ok[i] ? row[i] : static_cast<long long>(p.M) - 1;
  float u[NR][4], rr[NR][4];
  const bool grad_act = act == ALPRO_ACT_GELU_GRAD || act == ALPRO_ACT_RELU_GRAD;
  if (grad_act) {
#pragma unroll
    for (int i = 0; i < NR; ++i) load4_16(p.aux16 + lrow[i] * p.ldaux + col, u[i], p.aux_fmt, full, ncol);
  }
  if (has_resid) {
#pragma unroll
    for (int i = 0; i < NR; ++i) load4_32(p.resid + lrow[i] * p.ldresid + col, rr[i], full, ncol);
  }
  if (p.rs_acc) {   // stochastic depth: v = rs_acc[row] * acc + rs_bias[row] * bias
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const float ra = __ldg(p.rs_acc + lrow[i]);
      const float rb = p.rs_bias ? __ldg(p.rs_bias + lrow[i]) : ra;
      v[i][0] = fmaf(v[i][0], ra, b4.x * rb); v[i][1] = fmaf(v[i][1], ra, b4.y * rb);
      v[i][2] = fmaf(v[i][2], ra, b4.z * rb); v[i][3] = fmaf(v[i][3], ra, b4.w * rb);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      v[i][0] += b4.x; v[i][1] += b4.y; v[i][2] += b4.z; v[i][3] += b4.w;
    }
  }
  if (act == ALPRO_ACT_GELU) {
    // out16b receives gelu'(pre) (NOT the pre-activation): the backward epilogue is then a plain multiply
    float dv[NR][4];
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) gelu_erf_both(v[i][j], v[i][j], dv[i][j]);
    if (has_out16b) {
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if (ok[i]) store4_16(p.out16b + row[i] * p.ld16b + col, dv[i], p.out16b_fmt, full, ncol);
    }
  } else if (act == ALPRO_ACT_RELU) {
    if (has_out16b) {
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if (ok[i]) store4_16(p.out16b + row[i] * p.ld16b + col, v[i], p.out16b_fmt, full, ncol);
    }
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = fmaxf(v[i][j], 0.f);
  } else if (act == ALPRO_ACT_GELU_GRAD) {
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] *= u[i][j];   // aux = gelu'(pre) saved by the forward epilogue
  } else if (act == ALPRO_ACT_RELU_GRAD) {
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = u[i][j] > 0.f ? v[i][j] : 0.f;
  }
  if (has_resid) {
    float c2[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.bias2) load4_32(p.bias2 + col, c2, full, ncol);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const bool skip = p.skip_period > 0 && (row[i] % p.skip_period) == 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = skip ? rr[i][j] : v[i][j] + rr[i][j] + c2[j];
    }
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    if (ok[i]) {
      if (has_out32) store4_32(p.out32 + row[i] * p.ld32 + col, v[i], full, ncol);
      if (has_out16) store4_16(p.out16 + row[i] * p.ld16 + col, v[i], p.out16_fmt, full, ncol);
    }
  }
}


}  // namespace gemm
}  // namespace alpro
