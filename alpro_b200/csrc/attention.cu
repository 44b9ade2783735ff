// Attention kernels of the ALPRO path (head_dim = 64).
//
//  * temporal attention (TimeSformer divided attention over the T frames of one patch position, T <= 8):
//    one warp per (clip, patch, head); the whole TxT problem lives in registers. Reference: Attention.forward
//    vit.py:81-100 called through Block.forward vit.py:146-157 on '(b h w) t m'.
//  * sequence attention for S <= 256 keys (TimeSformer spatial attention over cls + N patches of one frame,
//    BERT text / fusion self-attention with additive key mask): one CTA per (sequence, head), Q/K/V tiles in
//    XOR-swizzled shared memory, QK^T and PV on mma.sync m16n8k16 tensor-core atoms with fp32 softmax.
//    Reference: vit.py:81-100 via Block.forward vit.py:165-181; BertSelfAttention.forward xbert.py:263-346.
//    Token gathering for the '(b t) (h w)' view is done by index arithmetic on the canonical 'b (h w t)' layout
//    instead of the reference's rearrange/cat copies.
//
// Backward kernels recompute the probabilities from the saved log-sum-exp (no SxS tensor ever reaches HBM).
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"
#include "rng.cuh"

namespace alpro {
namespace {

constexpr int DH = 64;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {   // single MUFU.EX2 (ex2() adds range/denormal fix-up code)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ================================================================================================ temporal attention
struct TAttnParams {
  const uint16_t* qkv;  // [rows, 3*d]
  uint16_t* out;        // fwd: o [rows, d];  bwd: dqkv [rows, 3*d]
  const uint16_t* dout; // bwd: do [rows, d]
  long long ld_qkv, ld_out, ld_dout;
  int B, N, heads, d, fmt;
  float scale;
};

// Lane mapping: lane = (i, part) with i = query/frame index (T of them) and `part` one of 32/T slices of the 64 head
// dims. A lane holds q_i[slice] and k_j[slice], v_j[slice] for all j, so every score needs only log2(32/T) shuffles
// (16 per warp for T=8) and the PV product is lane-local.
// 16-bit pair <-> fp32 with the storage format as a template parameter: one instruction per element (HADD2.F32 on a
// half of the packed register / a shift), where the run-time-format helpers of ptx.cuh cost three.
template <bool BF>
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi) {
  if (BF) {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
  } else {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x;
    hi = f.y;
  }
}
template <bool BF>
__device__ __forceinline__ uint32_t pack2c(float lo, float hi) {
  uint32_t d;
  if (BF) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

template <int DPP, bool BF>
__device__ __forceinline__ void ld16(const uint16_t* ptr, float (&out)[DPP]) {
  uint32_t w[DPP / 2];
  if (DPP == 16) {
    const uint4 a = *reinterpret_cast<const uint4*>(ptr);
    const uint4 b = *reinterpret_cast<const uint4*>(ptr + 8);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4 % (DPP / 2)] = b.x; w[5 % (DPP / 2)] = b.y; w[6 % (DPP / 2)] = b.z; w[7 % (DPP / 2)] = b.w;
  } else if (DPP == 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(ptr);
    w[0] = a.x; w[1 % (DPP / 2)] = a.y; w[2 % (DPP / 2)] = a.z; w[3 % (DPP / 2)] = a.w;
  } else if (DPP == 4) {
    const uint2 a = *reinterpret_cast<const uint2*>(ptr);
    w[0] = a.x; w[1 % (DPP / 2)] = a.y;
  } else {
    w[0] = *reinterpret_cast<const uint32_t*>(ptr);
  }
#pragma unroll
  for (int j = 0; j < DPP / 2; ++j) unpack2<BF>(w[j], out[2 * j], out[2 * j + 1]);
}
template <int DPP, bool BF>
__device__ __forceinline__ void st16(uint16_t* ptr, const float (&v)[DPP]) {
  uint32_t w[DPP / 2];
#pragma unroll
  for (int j = 0; j < DPP / 2; ++j) w[j] = pack2c<BF>(v[2 * j], v[2 * j + 1]);
  if (DPP == 16) {
    *reinterpret_cast<uint4*>(ptr) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(ptr + 8) = make_uint4(w[4 % (DPP / 2)], w[5 % (DPP / 2)], w[6 % (DPP / 2)], w[7 % (DPP / 2)]);
  } else if (DPP == 8) {
    *reinterpret_cast<uint4*>(ptr) = make_uint4(w[0], w[1 % (DPP / 2)], w[2 % (DPP / 2)], w[3 % (DPP / 2)]);
  } else if (DPP == 4) {
    *reinterpret_cast<uint2*>(ptr) = make_uint2(w[0], w[1 % (DPP / 2)]);
  } else {
    *reinterpret_cast<uint32_t*>(ptr) = w[0];
  }
}
template <int PARTS>
__device__ __forceinline__ float part_sum(float v) {
#pragma unroll
  for (int o = PARTS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Stage T rows x 64 dims of one 16-bit matrix into fp32 shared memory (coalesced 128-byte row reads).
// Shared-memory rows of the temporal backward are stored PERMUTED so that the PARTS lanes of one query row read one
// contiguous 16*PARTS-byte span with 128-bit loads (conflict-free; the straightforward [t][64] layout made the four
// parts of T = 8 collide pairwise on every scalar read: 49M bank conflicts per launch, ncu r01j):
//   element e = part * DPP + VW * kv + c   lives at   kv * (VW * PARTS) + VW * part + c      (VW = min(4, DPP))
template <int T, bool BF>
__device__ __forceinline__ void tattn_stage(const uint16_t* src, long long ld, int lane, float (*dst)[DH]) {
  constexpr int PARTS = 32 / T, DPP = DH / PARTS, VW = DPP < 4 ? DPP : 4;
  const int e = lane * 2;
  const int pos = ((e % DPP) / VW) * (VW * PARTS) + (e / DPP) * VW + (e % VW);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(src + t * ld + e);
    float2 f;
    unpack2<BF>(w, f.x, f.y);
    *reinterpret_cast<float2*>(&dst[t][pos]) = f;
  }
}
// this lane's DPP-wide slice (its `part`) of one staged row
template <int DPP, int PARTS>
__device__ __forceinline__ void tattn_ld_part(const float* row, int part, float (&out)[DPP]) {
  if (DPP >= 4) {
#pragma unroll
    for (int k4 = 0; k4 < DPP / 4; ++k4) {
      const float4 v = *reinterpret_cast<const float4*>(row + k4 * (4 * PARTS) + part * 4);
      out[4 * k4] = v.x; out[(4 * k4 + 1) % DPP] = v.y; out[(4 * k4 + 2) % DPP] = v.z; out[(4 * k4 + 3) % DPP] = v.w;
    }
  } else {   // DPP == 2 (T = 1)
    const float2 v = *reinterpret_cast<const float2*>(row + part * 2);
    out[0] = v.x; out[1 % DPP] = v.y;
  }
}

// softmax row of query i against all T keys: partial dots over this lane's DPP dims + log2(PARTS) shuffles per key
template <int T, int DPP, int PARTS>
__device__ __forceinline__ void tattn_row_probs(const float (&q)[DPP], const float (*sk)[DH], int part, float scale,
                                                float (&pr)[T]) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    float s = 0.f, kk[DPP];
    tattn_ld_part<DPP, PARTS>(sk[j], part, kk);
#pragma unroll
    for (int e = 0; e < DPP; ++e) s += q[e] * kk[e];
    pr[j] = part_sum<PARTS>(s) * scale;
    mx = fmaxf(mx, pr[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    pr[j] = __expf(pr[j] - mx);
    sum += pr[j];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < T; ++j) pr[j] *= inv;
}

// scores + softmax for this lane's query row i (identical in all PARTS lanes of the row)
template <int T, int DPP, int PARTS>
__device__ __forceinline__ void tattn_row_probs_reg(const float (&q)[DPP], const float (&k)[T][DPP], float scale,
                                                float (&pr)[T]) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < DPP; ++e) s += q[e] * k[j][e];
    pr[j] = part_sum<PARTS>(s) * scale;
    mx = fmaxf(mx, pr[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    pr[j] = __expf(pr[j] - mx);
    sum += pr[j];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < T; ++j) pr[j] *= inv;
}

// grid: ceil(B*N*heads / warps_per_block); each warp = one (b, n, head). cls rows are zero-filled by the n==0 warps.
template <int T, bool BF>
__global__ void __launch_bounds__(128, 3) tattn_fwd_kernel(const TAttnParams p) {
  pdl_grid_sync();
  constexpr int PARTS = 32 / T, DPP = DH / PARTS;
  const int lane = threadIdx.x & 31;
  const long long unit = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = static_cast<long long>(p.B) * p.N * p.heads;
  if (unit >= total) return;
  const int head = static_cast<int>(unit % p.heads);
  const long long bn = unit / p.heads;
  const int n = static_cast<int>(bn % p.N);
  const long long b = bn / p.N;
  const long long clip_rows = 1 + static_cast<long long>(p.N) * T;
  const long long row0 = b * clip_rows + 1 + static_cast<long long>(n) * T;
  if (n == 0) *reinterpret_cast<uint32_t*>(p.out + b * clip_rows * p.ld_out + head * DH + lane * 2) = 0u;  // cls row
  const int i = lane / PARTS, part = lane % PARTS;
  const int coff = head * DH + part * DPP;
  float q[DPP], k[T][DPP], v[T][DPP], pr[T];
  ld16<DPP, BF>(p.qkv + (row0 + i) * p.ld_qkv + coff, q);
#pragma unroll
  for (int j = 0; j < T; ++j) {
    ld16<DPP, BF>(p.qkv + (row0 + j) * p.ld_qkv + p.d + coff, k[j]);
    ld16<DPP, BF>(p.qkv + (row0 + j) * p.ld_qkv + 2 * p.d + coff, v[j]);
  }
  tattn_row_probs_reg<T, DPP, PARTS>(q, k, p.scale, pr);
  float o[DPP];
#pragma unroll
  for (int e = 0; e < DPP; ++e) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < T; ++j) a += pr[j] * v[j][e];
    o[e] = a;
  }
  st16<DPP, BF>(p.out + (row0 + i) * p.ld_out + coff, o);
}

template <int T, bool BF>
__global__ void __launch_bounds__(128, 4) tattn_bwd_kernel(const TAttnParams p) {
  pdl_grid_sync();
  constexpr int PARTS = 32 / T, DPP = DH / PARTS, WPB = 4;
  __shared__ __align__(16) float sQ[WPB][T][DH], sK[WPB][T][DH], sV[WPB][T][DH], sG[WPB][T][DH];
  __shared__ float sP[WPB][T][T], sDS[WPB][T][T];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long unit = static_cast<long long>(blockIdx.x) * WPB + wib;
  const long long total = static_cast<long long>(p.B) * p.N * p.heads;
  if (unit >= total) return;
  const int head = static_cast<int>(unit % p.heads);
  const long long bn = unit / p.heads;
  const int n = static_cast<int>(bn % p.N);
  const long long b = bn / p.N;
  const long long clip_rows = 1 + static_cast<long long>(p.N) * T;
  const long long row0 = b * clip_rows + 1 + static_cast<long long>(n) * T;
  if (n == 0) {  // cls row receives no gradient from the temporal branch
    uint16_t* z = p.out + b * clip_rows * p.ld_out + head * DH + lane * 2;
    *reinterpret_cast<uint32_t*>(z) = 0u;
    *reinterpret_cast<uint32_t*>(z + p.d) = 0u;
    *reinterpret_cast<uint32_t*>(z + 2 * p.d) = 0u;
  }
  const uint16_t* base = p.qkv + row0 * p.ld_qkv + head * DH;
  tattn_stage<T, BF>(base, p.ld_qkv, lane, sQ[wib]);
  tattn_stage<T, BF>(base + p.d, p.ld_qkv, lane, sK[wib]);
  tattn_stage<T, BF>(base + 2 * p.d, p.ld_qkv, lane, sV[wib]);
  tattn_stage<T, BF>(p.dout + row0 * p.ld_dout + head * DH, p.ld_dout, lane, sG[wib]);
  __syncwarp();
  const int i = lane / PARTS, part = lane % PARTS;
  float q[DPP], go[DPP], pr[T];
  tattn_ld_part<DPP, PARTS>(sQ[wib][i], part, q);
  tattn_ld_part<DPP, PARTS>(sG[wib][i], part, go);
  tattn_row_probs<T, DPP, PARTS>(q, sK[wib], part, p.scale, pr);
  float dp[T], dot = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    float sdot = 0.f, vv[DPP];
    tattn_ld_part<DPP, PARTS>(sV[wib][j], part, vv);
#pragma unroll
    for (int e = 0; e < DPP; ++e) sdot += go[e] * vv[e];
    dp[j] = part_sum<PARTS>(sdot);
    dot += dp[j] * pr[j];
  }
  float dq[DPP];
#pragma unroll
  for (int e = 0; e < DPP; ++e) dq[e] = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    const float ds = pr[j] * (dp[j] - dot) * p.scale;
    if (part == 0) {
      sP[wib][i][j] = pr[j];
      sDS[wib][i][j] = ds;
    }
    float kk[DPP];
    tattn_ld_part<DPP, PARTS>(sK[wib][j], part, kk);
#pragma unroll
    for (int e = 0; e < DPP; ++e) dq[e] += ds * kk[e];
  }
  uint16_t* o = p.out + (row0 + i) * p.ld_out + head * DH + part * DPP;
  st16<DPP, BF>(o, dq);
  __syncwarp();
  // role switch: this lane now owns key/value row j = i
  // (two passes so that only one accumulator + one operand slice is live at a time: 128-register budget)
  {
    float dk[DPP];
#pragma unroll
    for (int e = 0; e < DPP; ++e) dk[e] = 0.f;
#pragma unroll
    for (int r = 0; r < T; ++r) {
      const float ds = sDS[wib][r][i];
      float qq[DPP];
      tattn_ld_part<DPP, PARTS>(sQ[wib][r], part, qq);
#pragma unroll
      for (int e = 0; e < DPP; ++e) dk[e] += ds * qq[e];
    }
    st16<DPP, BF>(o + p.d, dk);
  }
  float dv[DPP];
#pragma unroll
  for (int e = 0; e < DPP; ++e) dv[e] = 0.f;
#pragma unroll
  for (int r = 0; r < T; ++r) {
    const float pp = sP[wib][r][i];
    float gg[DPP];
    tattn_ld_part<DPP, PARTS>(sG[wib][r], part, gg);
#pragma unroll
    for (int e = 0; e < DPP; ++e) dv[e] += pp * gg[e];
  }
  st16<DPP, BF>(o + 2 * p.d, dv);
}

// ================================================================================================ sequence attention
struct SAttnParams {
  const uint16_t* qkv;   // [rows, ld_qkv]: q at col 0, k at col d, v at col 2d (head h at +h*64)
  const float* mask;     // additive key mask [nseq, S] or null
  uint16_t* o;           // fwd out [rows, ld_o]
  uint16_t* cls_o;       // fwd: if non-null, token 0 output goes to cls_o[seq, d] instead of its canonical row
  float* lse;            // [nseq, heads, S]
  // backward
  const uint16_t* dout;  // [rows, ld_o] upstream grad; with seq_div > 1 token 0 is the group's shared cls row whose
                         // forward value was the mean over the seq_div frames, so each frame receives dout / seq_div
  const uint16_t* o_fwd;   // forward output rows [rows, ld_o] (for D = rowsum(dO * O))
  const uint16_t* cls_fwd; // forward per-sequence token-0 outputs [nseq, d] when seq_div > 1
  const float* cls_weight; // optional [nseq]: d(group cls row)/d(frame cls output); default 1/seq_div (plain mean)
  uint16_t* dqkv;        // [rows, ld_qkv]
  float* dcls_qkv;       // [nseq, 3*d] fp32: per-sequence gradient of the shared cls q/k/v row (when seq_div > 1)
  long long ld_qkv, ld_o;
  int S, nseq, heads, d, fmt;
  int seq_div, stride;   // row(seq, j) = (seq / seq_div) * clip_rows + (j == 0 ? 0 : 1 + seq % seq_div + (j-1) * stride)
  long long clip_rows;
  float scale;
  // train-mode attention-probability dropout (BertSelfAttention.dropout, xbert.py:331): 0 = off
  uint32_t drop_thr, drop_seed;
  float drop_scale;
  int cls_last;          // TMA kernels: tile row S-1 holds token 0 (the strided tokens 1..S-1 are ONE tensor-map box)
  long long* trace;      // diagnostics (ALPRO_ATTN_TRACE=1): 64 clock64() stamps per CTA, null in normal runs
  int stagger;           // tcgen05 backward: first-wave CTA i starts (i % 4) * stagger cycles late (0 = off)
};

// mask/keep factors of the key pair (2*jp, 2*jp+1) for query row i of (seq, head): one counter-hash per pair
__device__ __forceinline__ void drop_pair(const SAttnParams& p, int S_pad, int seq, int head, int i, int jp, float& m0,
                                          float& m1) {
  const uint64_t idx = (static_cast<uint64_t>(seq) * p.heads + head) * S_pad * (S_pad >> 1) +
                       static_cast<uint64_t>(i) * (S_pad >> 1) + jp;
  const uint32_t h = rand16x2(p.drop_seed, idx);
  m0 = (h & 0xffff) >= p.drop_thr ? p.drop_scale : 0.f;
  m1 = (h >> 16) >= p.drop_thr ? p.drop_scale : 0.f;
}

// Canonical row of token j of one sequence: row(seq, j) = (seq / seq_div) * clip_rows + (j == 0 ? 0 : 1 + seq % seq_div +
// (j-1) * stride). The division / modulo by the run-time seq_div is done ONCE per CTA (it was 16% of the backward's
// instructions when evaluated per access, ncu r01j).
struct SeqRows {
  long long base, first;
  int stride;
  __device__ __forceinline__ long long operator()(int j) const {
    return j == 0 ? base : first + static_cast<long long>(j - 1) * stride;
  }
};
__device__ __forceinline__ SeqRows make_rows(const SAttnParams& p, int seq) {
  SeqRows r;
  r.base = static_cast<long long>(seq / p.seq_div) * p.clip_rows;
  r.first = r.base + 1 + (seq % p.seq_div);
  r.stride = p.stride;
  return r;
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <bool BF>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (BF) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
template <bool BF>
__device__ __forceinline__ uint32_t pack2(float a, float b) { return pack2_16(a, b, BF ? 1 : 0); }

// smem tile [rows][64] 16-bit, 128-byte rows, 16-byte chunk index XOR (row & 7)
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int col) {
  return base + row * 128 + ((((col >> 3) ^ (row & 7)) << 4)) + ((col & 7) << 1);
}

// Cooperative gather of one [S_pad][64] tile (q, k or v columns of qkv, or a plain [rows, ld] matrix) with 16-byte
// cp.async copies: every thread keeps all of its copies in flight (no register staging), rows >= S are zero-filled
// through the src-size operand. Completion: cp_async_wait_all() + __syncthreads().
__device__ __forceinline__ void load_tile(const SAttnParams& p, const uint16_t* src, long long ld, int col0,
                                          const SeqRows& rows, int S_pad, uint8_t* smem_tile, const uint16_t* tok0_override) {
  const uint32_t base = smem_u32(smem_tile);
  for (int idx = threadIdx.x; idx < S_pad * 8; idx += blockDim.x) {
    const int row = idx >> 3, ch = idx & 7;
    const int rr = row < p.S ? row : 0;   // clamped (the copy reads 0 bytes for padded rows)
    const uint16_t* g = (rr == 0 && tok0_override) ? tok0_override + ch * 8
                                                    : src + rows(rr) * ld + col0 + ch * 8;
    const uint32_t nbytes = row < p.S ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + row * 128 + ((ch ^ (row & 7)) << 4)),
                 "l"(g), "r"(nbytes)
                 : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ forward
// grid (heads, nseq), 128 threads. NT = number of 8-key tiles; EXACT = NT is the true tile count (no predication),
// otherwise NT is an upper bound and the loops are predicated on the runtime count.
template <bool BF, int NT, bool EXACT>
__global__ void __launch_bounds__(128) sattn_fwd_kernel(const SAttnParams p) {
  pdl_grid_sync();
  extern __shared__ __align__(128) uint8_t sm[];
  const int head = blockIdx.x, seq = blockIdx.y;
  const SeqRows rows = make_rows(p, seq);
  const int S_pad = (p.S + 15) & ~15;
  uint8_t* sQ = sm;
  uint8_t* sK = sQ + S_pad * 128;
  uint8_t* sV = sK + S_pad * 128;
  float* sMask = reinterpret_cast<float*>(sV + S_pad * 128);
  load_tile(p, p.qkv, p.ld_qkv, head * DH, rows, S_pad, sQ, nullptr);
  load_tile(p, p.qkv, p.ld_qkv, p.d + head * DH, rows, S_pad, sK, nullptr);
  load_tile(p, p.qkv, p.ld_qkv, 2 * p.d + head * DH, rows, S_pad, sV, nullptr);
  for (int j = threadIdx.x; j < S_pad; j += blockDim.x)
    sMask[j] = j < p.S ? (p.mask ? p.mask[static_cast<long long>(seq) * p.S + j] * LOG2E : 0.f) : -INFINITY;
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nt = EXACT ? NT : (S_pad >> 3);
  const float sl2 = p.scale * LOG2E;
  // Per-lane ldmatrix address pieces. Every fragment row index is (lane & 7) + multiples of 8, so the swizzle term
  // (chunk ^ (row & 7)) depends only on the lane and the 16-byte chunk -> four precomputed byte offsets per pattern.
  const int l7 = lane & 7, b3 = (lane >> 3) & 1, b4 = lane >> 4;
  uint32_t xa[4], xb[4];   // chunk = 2*i + b4 (A operands / transposed B)   and   chunk = 2*i + b3 (plain B operands)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xa[i] = static_cast<uint32_t>(((2 * i + b4) ^ l7) << 4);
    xb[i] = static_cast<uint32_t>(((2 * i + b3) ^ l7) << 4);
  }
  const uint32_t q_lane = smem_u32(sQ) + (l7 + b3 * 8) * 128;   // + qt*2048 + xa[ks]
  const uint32_t k_lane = smem_u32(sK) + (l7 + b4 * 8) * 128;   // + n2*2048 + xb[ks]
  const uint32_t v_lane = smem_u32(sV) + (l7 + b3 * 8) * 128;   // + kk*2048 + xa[dp]

  for (int qt = warp; qt < (S_pad >> 4); qt += 4) {
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], q_lane + qt * 2048 + xa[ks]);
    float s[NT][4];
#pragma unroll
    for (int n2 = 0; n2 < NT / 2; ++n2) {
      if (EXACT || n2 * 2 < nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) s[2 * n2][e] = s[2 * n2 + 1][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t kb[4];
          ldsm_x4(kb, k_lane + n2 * 2048 + xb[ks]);
          mma16816<BF>(s[2 * n2], qa[ks], kb[0], kb[1]);
          mma16816<BF>(s[2 * n2 + 1], qa[ks], kb[2], kb[3]);
        }
      }
    }
    // softmax over keys (rows g and g+8 of this query tile), base-2 domain
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (EXACT || n < nt) {
        const float2 mk = *reinterpret_cast<const float2*>(sMask + n * 8 + 2 * t);
        s[n][0] = fmaf(s[n][0], sl2, mk.x); s[n][1] = fmaf(s[n][1], sl2, mk.y);
        s[n][2] = fmaf(s[n][2], sl2, mk.x); s[n][3] = fmaf(s[n][3], sl2, mk.y);
        m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
        m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (EXACT || n < nt) {
        s[n][0] = ex2(s[n][0] - m0); s[n][1] = ex2(s[n][1] - m0);
        s[n][2] = ex2(s[n][2] - m1); s[n][3] = ex2(s[n][3] - m1);
        l0 += s[n][0] + s[n][1];
        l1 += s[n][2] + s[n][3];
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    if (p.drop_thr) {   // dropout acts on the probabilities that multiply V; the row sums above stay undropped
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (EXACT || n < nt) {
          float a0, a1, b0, b1;
          drop_pair(p, S_pad, seq, head, qt * 16 + g, n * 4 + t, a0, a1);
          drop_pair(p, S_pad, seq, head, qt * 16 + g + 8, n * 4 + t, b0, b1);
          s[n][0] *= a0; s[n][1] *= a1; s[n][2] *= b0; s[n][3] *= b1;
        }
      }
    }
    // O = P V
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      if (EXACT || kk * 2 < nt) {
        uint32_t pa[4];
        pa[0] = pack2<BF>(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack2<BF>(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack2<BF>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack2<BF>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vb[4];
          ldsm_x4_t(vb, v_lane + kk * 2048 + xa[dp]);
          mma16816<BF>(o[2 * dp], pa, vb[0], vb[1]);
          mma16816<BF>(o[2 * dp + 1], pa, vb[2], vb[3]);
        }
      }
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = qt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? r1 : r0;
      if (r < p.S) {
        const float inv = half ? i1 : i0;
        uint16_t* dst = (r == 0 && p.cls_o) ? p.cls_o + static_cast<long long>(seq) * p.d + head * DH
                                            : p.o + rows(r) * p.ld_o + head * DH;
#pragma unroll
        for (int n = 0; n < 8; ++n)
          *reinterpret_cast<uint32_t*>(dst + n * 8 + 2 * t) =
              pack2<BF>(o[n][half * 2] * inv, o[n][half * 2 + 1] * inv);
        if (t == 0 && p.lse)
          p.lse[(static_cast<long long>(seq) * p.heads + head) * p.S + r] = ((half ? m1 : m0) + log2f(half ? l1 : l0));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// grid (heads, nseq), 128 threads. smem: Q, K, V, dO tiles + lse2[S_pad] (base-2 log-sum-exp) + D[S_pad] + mask.
// Phase A: each warp owns 16-query tiles -> dQ.   Phase B: each warp owns 16-key tiles -> dK, dV (recomputing S^T).
template <bool BF>
__global__ void __launch_bounds__(256, 2) sattn_bwd_kernel(const SAttnParams p) {
  pdl_grid_sync();
  extern __shared__ __align__(128) uint8_t sm[];
  const int head = blockIdx.x, seq = blockIdx.y;
  const SeqRows rows = make_rows(p, seq);
  const int S_pad = (p.S + 15) & ~15;
  uint8_t* sQ = sm;
  uint8_t* sK = sQ + S_pad * 128;
  uint8_t* sV = sK + S_pad * 128;
  uint8_t* sG = sV + S_pad * 128;  // dO
  float* sMask = reinterpret_cast<float*>(sG + S_pad * 128);
  float* sLse = sMask + S_pad;
  float* sD = sLse + S_pad;
  load_tile(p, p.qkv, p.ld_qkv, head * DH, rows, S_pad, sQ, nullptr);
  load_tile(p, p.qkv, p.ld_qkv, p.d + head * DH, rows, S_pad, sK, nullptr);
  load_tile(p, p.qkv, p.ld_qkv, 2 * p.d + head * DH, rows, S_pad, sV, nullptr);
  load_tile(p, p.dout, p.ld_o, head * DH, rows, S_pad, sG, nullptr);
  const float gscale0 = p.cls_weight ? p.cls_weight[seq] : 1.f / p.seq_div;  // d(weighted mean_t cls_t) / d cls_t
  for (int j = threadIdx.x; j < S_pad; j += blockDim.x) {
    sMask[j] = j < p.S ? (p.mask ? p.mask[static_cast<long long>(seq) * p.S + j] * LOG2E : 0.f) : -INFINITY;
    sLse[j] = j < p.S ? p.lse[(static_cast<long long>(seq) * p.heads + head) * p.S + j] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // token-0 upstream gradient scaling (mean over frames) applied in place, then D = rowsum(dO * O) recomputed as
  // sum_j P_ij dP_ij is avoided: O is not stored per frame for the cls row, so D is computed from P and dP below.
  if (p.seq_div > 1 && warp == 0) {
    // scale row 0 of dO by 1/seq_div
    for (int c = lane; c < 64; c += 32) {
      uint16_t* e = reinterpret_cast<uint16_t*>(sG + 0 * 128 + ((((c >> 3) ^ 0) << 4)) + ((c & 7) << 1));
      *e = f32_to_16(f16_to_32(*e, BF ? 1 : 0) * gscale0, BF ? 1 : 0);
    }
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const uint32_t q_base = smem_u32(sQ), k_base = smem_u32(sK), v_base = smem_u32(sV), g_base = smem_u32(sG);
  // per-lane ldmatrix address pieces (see sattn_fwd_kernel): fragment rows are (lane & 7) + multiples of 8
  const int l7 = lane & 7, b3 = (lane >> 3) & 1, b4 = lane >> 4;
  uint32_t xa[4], xb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xa[i] = static_cast<uint32_t>(((2 * i + b4) ^ l7) << 4);   // A operands / transposed B operands
    xb[i] = static_cast<uint32_t>(((2 * i + b3) ^ l7) << 4);   // plain B operands
  }
  const uint32_t rowA = (l7 + b3 * 8) * 128, rowB = (l7 + b4 * 8) * 128;
  const float sl2 = p.scale * LOG2E;
  const int nkb = S_pad >> 4;  // 16-wide blocks along either sequence axis

  // ---------------- D_i = rowsum(dO_i * O_i): O from global (coalesced 128-byte rows, 8 rows in flight per warp),
  // dO from the staged tile
  for (int r0 = warp; r0 < S_pad; r0 += 64) {
    uint32_t wo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + 8 * u;
      wo[u] = 0u;
      if (r < p.S) {
        const uint16_t* orow = (r == 0 && p.seq_div > 1) ? p.cls_fwd + static_cast<long long>(seq) * p.d + head * DH
                                                          : p.o_fwd + rows(r) * p.ld_o + head * DH;
        wo[u] = *reinterpret_cast<const uint32_t*>(orow + lane * 2);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + 8 * u;
      if (r < S_pad) {   // warp-uniform
        const int ch = lane >> 2;   // 16-byte chunk of the dO row holding columns 2*lane, 2*lane+1
        const uint32_t wg = *reinterpret_cast<const uint32_t*>(sG + r * 128 + ((ch ^ (r & 7)) << 4) + ((lane & 3) << 2));
        float dsum = f16_to_32(static_cast<uint16_t>(wo[u] & 0xffff), BF ? 1 : 0) * f16_to_32(static_cast<uint16_t>(wg & 0xffff), BF ? 1 : 0) +
                     f16_to_32(static_cast<uint16_t>(wo[u] >> 16), BF ? 1 : 0) * f16_to_32(static_cast<uint16_t>(wg >> 16), BF ? 1 : 0);
        dsum = warp_sum(dsum);
        if (lane == 0) sD[r] = dsum;
      }
    }
  }
  __syncthreads();

  // ---------------- phase A: dQ (warp per 16-query tile)
  for (int qt = warp; qt < nkb; qt += 8) {
    uint32_t qa[4][4], ga[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldsm_x4(qa[ks], q_base + rowA + qt * 2048 + xa[ks]);
      ldsm_x4(ga[ks], g_base + rowA + qt * 2048 + xa[ks]);
    }
    const float ls0 = sLse[qt * 16 + g], ls1 = sLse[qt * 16 + g + 8];
    const float D0 = sD[qt * 16 + g], D1 = sD[qt * 16 + g + 8];
    float dq[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[n][e] = 0.f;
    for (int kb = 0; kb < nkb; ++kb) {
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t kf[4], vf[4];
        ldsm_x4(kf, k_base + rowB + kb * 2048 + xb[ks]);
        ldsm_x4(vf, v_base + rowB + kb * 2048 + xb[ks]);
        mma16816<BF>(s[0], qa[ks], kf[0], kf[1]);
        mma16816<BF>(s[1], qa[ks], kf[2], kf[3]);
        mma16816<BF>(dp[0], ga[ks], vf[0], vf[1]);
        mma16816<BF>(dp[1], ga[ks], vf[2], vf[3]);
      }
      float ds[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const float mk0 = sMask[kb * 16 + n * 8 + 2 * t], mk1 = sMask[kb * 16 + n * 8 + 2 * t + 1];
        if (p.drop_thr) {   // dP flows through the dropout mask of the forward pass
          float a0, a1, b0, b1;
          drop_pair(p, S_pad, seq, head, qt * 16 + g, kb * 8 + n * 4 + t, a0, a1);
          drop_pair(p, S_pad, seq, head, qt * 16 + g + 8, kb * 8 + n * 4 + t, b0, b1);
          dp[n][0] *= a0; dp[n][1] *= a1; dp[n][2] *= b0; dp[n][3] *= b1;
        }
        ds[n][0] = ex2(s[n][0] * sl2 + mk0 - ls0) * (dp[n][0] - D0) * p.scale;
        ds[n][1] = ex2(s[n][1] * sl2 + mk1 - ls0) * (dp[n][1] - D0) * p.scale;
        ds[n][2] = ex2(s[n][2] * sl2 + mk0 - ls1) * (dp[n][2] - D1) * p.scale;
        ds[n][3] = ex2(s[n][3] * sl2 + mk1 - ls1) * (dp[n][3] - D1) * p.scale;
      }
      uint32_t da[4] = {pack2<BF>(ds[0][0], ds[0][1]), pack2<BF>(ds[0][2], ds[0][3]), pack2<BF>(ds[1][0], ds[1][1]),
                        pack2<BF>(ds[1][2], ds[1][3])};
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {  // dQ += dS K   (B = K[key][dh], k = key -> transposed ldmatrix)
        uint32_t kb4[4];
        ldsm_x4_t(kb4, k_base + rowA + kb * 2048 + xa[dpair]);
        mma16816<BF>(dq[2 * dpair], da, kb4[0], kb4[1]);
        mma16816<BF>(dq[2 * dpair + 1], da, kb4[2], kb4[3]);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = qt * 16 + g + half * 8;
      if (r < p.S) {
        if (r == 0 && p.dcls_qkv) {
          float* dst = p.dcls_qkv + static_cast<long long>(seq) * 3 * p.d + head * DH;
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            dst[n * 8 + 2 * t] = dq[n][half * 2];
            dst[n * 8 + 2 * t + 1] = dq[n][half * 2 + 1];
          }
        } else {
          uint16_t* dst = p.dqkv + rows(r) * p.ld_qkv + head * DH;
#pragma unroll
          for (int n = 0; n < 8; ++n)
            *reinterpret_cast<uint32_t*>(dst + n * 8 + 2 * t) = pack2<BF>(dq[n][half * 2], dq[n][half * 2 + 1]);
        }
      }
    }
  }

  // ---------------- phase B: dK, dV (warp per 16-key tile); S^T = K Q^T, dP^T = V dO^T
  for (int kt = warp; kt < nkb; kt += 8) {
    uint32_t ka[4][4], va[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldsm_x4(ka[ks], k_base + rowA + kt * 2048 + xa[ks]);
      ldsm_x4(va[ks], v_base + rowA + kt * 2048 + xa[ks]);
    }
    const float mk0 = sMask[kt * 16 + g], mk1 = sMask[kt * 16 + g + 8];  // key = fragment row now
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) dk[n][e] = dv[n][e] = 0.f;
    for (int qb = 0; qb < nkb; ++qb) {
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t qf[4], gf[4];
        ldsm_x4(qf, q_base + rowB + qb * 2048 + xb[ks]);
        ldsm_x4(gf, g_base + rowB + qb * 2048 + xb[ks]);
        mma16816<BF>(st[0], ka[ks], qf[0], qf[1]);
        mma16816<BF>(st[1], ka[ks], qf[2], qf[3]);
        mma16816<BF>(dpt[0], va[ks], gf[0], gf[1]);
        mma16816<BF>(dpt[1], va[ks], gf[2], gf[3]);
      }
      float pt[2][4], dst_[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int qc = qb * 16 + n * 8 + 2 * t;  // query index = fragment column
        const float l0 = sLse[qc], l1 = sLse[qc + 1], D0 = sD[qc], D1 = sD[qc + 1];
        const bool v0 = qc < p.S, v1 = qc + 1 < p.S;  // padded query rows contribute nothing
        pt[n][0] = v0 ? ex2(st[n][0] * sl2 + mk0 - l0) : 0.f;
        pt[n][1] = v1 ? ex2(st[n][1] * sl2 + mk0 - l1) : 0.f;
        pt[n][2] = v0 ? ex2(st[n][2] * sl2 + mk1 - l0) : 0.f;
        pt[n][3] = v1 ? ex2(st[n][3] * sl2 + mk1 - l1) : 0.f;
        float mq[4] = {1.f, 1.f, 1.f, 1.f};   // mask(query, key) for the 4 fragment elements (key rows g / g+8)
        if (p.drop_thr) {
          const int j0 = kt * 16 + g, j1 = j0 + 8;
          float lo, hi;
          drop_pair(p, S_pad, seq, head, qc, j0 >> 1, lo, hi);     mq[0] = (j0 & 1) ? hi : lo;
          drop_pair(p, S_pad, seq, head, qc + 1, j0 >> 1, lo, hi); mq[1] = (j0 & 1) ? hi : lo;
          drop_pair(p, S_pad, seq, head, qc, j1 >> 1, lo, hi);     mq[2] = (j1 & 1) ? hi : lo;
          drop_pair(p, S_pad, seq, head, qc + 1, j1 >> 1, lo, hi); mq[3] = (j1 & 1) ? hi : lo;
        }
        dst_[n][0] = pt[n][0] * (dpt[n][0] * mq[0] - D0) * p.scale;
        dst_[n][1] = pt[n][1] * (dpt[n][1] * mq[1] - D1) * p.scale;
        dst_[n][2] = pt[n][2] * (dpt[n][2] * mq[2] - D0) * p.scale;
        dst_[n][3] = pt[n][3] * (dpt[n][3] * mq[3] - D1) * p.scale;
        pt[n][0] *= mq[0]; pt[n][1] *= mq[1]; pt[n][2] *= mq[2]; pt[n][3] *= mq[3];   // dV uses the dropped probabilities
      }
      uint32_t pa[4] = {pack2<BF>(pt[0][0], pt[0][1]), pack2<BF>(pt[0][2], pt[0][3]), pack2<BF>(pt[1][0], pt[1][1]),
                        pack2<BF>(pt[1][2], pt[1][3])};
      uint32_t da[4] = {pack2<BF>(dst_[0][0], dst_[0][1]), pack2<BF>(dst_[0][2], dst_[0][3]),
                        pack2<BF>(dst_[1][0], dst_[1][1]), pack2<BF>(dst_[1][2], dst_[1][3])};
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {
        uint32_t gb[4], qb4[4];
        ldsm_x4_t(gb, g_base + rowA + qb * 2048 + xa[dpair]);    // dV += P^T dO
        ldsm_x4_t(qb4, q_base + rowA + qb * 2048 + xa[dpair]);   // dK += dS^T Q
        mma16816<BF>(dv[2 * dpair], pa, gb[0], gb[1]);
        mma16816<BF>(dv[2 * dpair + 1], pa, gb[2], gb[3]);
        mma16816<BF>(dk[2 * dpair], da, qb4[0], qb4[1]);
        mma16816<BF>(dk[2 * dpair + 1], da, qb4[2], qb4[3]);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = kt * 16 + g + half * 8;
      if (r < p.S) {
        if (r == 0 && p.dcls_qkv) {
          float* dst = p.dcls_qkv + static_cast<long long>(seq) * 3 * p.d + head * DH;
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            dst[p.d + n * 8 + 2 * t] = dk[n][half * 2];
            dst[p.d + n * 8 + 2 * t + 1] = dk[n][half * 2 + 1];
            dst[2 * p.d + n * 8 + 2 * t] = dv[n][half * 2];
            dst[2 * p.d + n * 8 + 2 * t + 1] = dv[n][half * 2 + 1];
          }
        } else {
          uint16_t* dst = p.dqkv + rows(r) * p.ld_qkv + head * DH;
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            *reinterpret_cast<uint32_t*>(dst + p.d + n * 8 + 2 * t) = pack2<BF>(dk[n][half * 2], dk[n][half * 2 + 1]);
            *reinterpret_cast<uint32_t*>(dst + 2 * p.d + n * 8 + 2 * t) = pack2<BF>(dv[n][half * 2], dv[n][half * 2 + 1]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward, tcgen05
// Same contract as sattn_fwd_kernel, but both contractions run on the 5th-generation tensor cores:
//   S = Q K^T : tcgen05.mma M=128, N=S32 (<= 256), K=64   -> fp32 scores in TMEM (one TMEM lane per query row)
//   softmax   : thread r owns query row r (tcgen05.ld of its TMEM lane), base-2 exp, fp32 row statistics
//   O = P V   : P written as a 16-bit K-major 128B-swizzled A tile in shared memory (128 keys at a time),
//               V consumed in place as an MN-major B operand; fp32 O accumulator in TMEM columns [0,64)
// The gathered Q/K/V tiles already use the UMMA 128B-swizzle K-major layout, so the cp.async gather needs no repacking.
// One CTA (4 warps) per (sequence, head), ~105 KB smem and 256 TMEM columns -> two CTAs per SM overlap each other's
// load / MMA / softmax phases. 197 queries = two M=128 tiles (the second has 69 valid rows).
// r01p revision (lessons of the backward kernel's phase trace): MMAs issued by an elected lane in warp-uniform control
// flow (back-to-back UTCHMMA instead of one ELECT/R2UR loop per MMA), the next query tile is gathered while the current
// one is processed, TMEM chunks are loaded two at a time, O rows leave through a swizzled staging block so that every
// store instruction writes four full 128-byte rows, dropout is a template parameter.
#define ATTN_TRACE(cond, slot)                                                                             \
  do {                                                                                                   \
    if (p.trace && (cond))                                                                               \
      p.trace[(static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * 64 + (slot)] = clock64() - t_start; \
  } while (0)
// TMA = true (r02): the operand tiles arrive by tensor-map loads instead of the per-thread cp.async gather. The strided
// token rows of one sequence (row = clip*clip_rows + 1 + frame + (j-1)*stride) are ONE box of a 4-D map
// (column, token, frame, clip) over the qkv matrix, written by the TMA unit straight into the 128B-swizzled K-major
// layout: one elected thread issues three instructions where 256 threads issued ~4600 predicated 16-byte copies with
// their address arithmetic (3.1 k of the CTA's 24 k cycles, profiles/r01p_sattn_tc_traces.md). The shared cls token
// is not part of that box; it is placed LAST (tile row S-1, cp.async by 8 threads) so that the box lands at the
// 1024-byte-aligned tile base: attention is invariant under a common permutation of keys/queries, only the row <->
// token map of the mask, the log-sum-exp and the output rows changes (tok()).
template <bool BF, bool DROP, bool MASK, bool TMA>
__global__ void __launch_bounds__(256, 2) sattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmAll,
                                                              const __grid_constant__ CUtensorMap tmQ0,
                                                              const __grid_constant__ CUtensorMap tmQ1,
                                                              const SAttnParams p) {
  const long long t_start = p.trace ? clock64() : 0;
  extern __shared__ uint8_t sm_raw[];
  const uint32_t raw_addr = smem_u32(sm_raw);
  uint8_t* sm = sm_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int head = blockIdx.x, seq = blockIdx.y;
  const SeqRows rows = make_rows(p, seq);
  const int S = p.S;
  const int S_pad = (S + 15) & ~15;  // dropout-mask indexing (shared with the mma.sync kernels)
  const int S32 = (S + 31) & ~31;    // keys padded to whole 32-column TMEM chunks; padded K/V rows are zero
  uint8_t* sQ = sm;                    // [128][64]  one query tile
  uint8_t* sK = sQ + 128 * 128;        // [S32][64]
  uint8_t* sV = sK + S32 * 128;        // [S32][64]
  uint8_t* sP = sV + S32 * 128;        // 2 blocks of [128][64]: probabilities of 128 keys; O staging in the epilogue
  float* sMask = reinterpret_cast<float*>(sP + 2 * 16384);
  float* sStat = sMask + 256;          // [2][128] row maxima, then [2][128] row sums of the two column halves
  uint64_t* bar = reinterpret_cast<uint64_t*>(sStat + 512);
  uint64_t* ldbar = bar + 1;           // TMA: K, V and the first query tile have landed
  uint64_t* qbar = bar + 2;            // TMA: the next query tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  // Two threads per query row: warp w and warp w + 4 share TMEM lane quarter w % 4; group `wg` takes the 32-key chunks
  // with (chunk & 1) == wg, and columns [32 wg, 32 wg + 32) of the output row.
  const int wg = warp >> 2, lq = warp & 3;
  const int rit = lq * 32 + lane;      // row inside the query tile = TMEM lane
  // tile row -> token of the sequence (identity unless the cls token sits last)
  const bool cls_last = TMA && p.cls_last;
  auto tok = [&](int j) { return cls_last ? (j == S - 1 ? 0 : j + 1) : j; };
  const int n_box = cls_last ? S - 1 : S;                       // rows delivered by the tensor-map boxes
  const int c_t = seq % p.seq_div, c_b = seq / p.seq_div;       // frame / clip coordinates of the 4-D map

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(ldbar, 1);
    mbar_init(qbar, 1);
    fence_barrier_init();
    if (TMA) {
      tma_prefetch_desc(&tmAll);
      tma_prefetch_desc(&tmQ0);
    }
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  pdl_grid_sync();   // setup above overlapped the previous kernel's tail (common.h)
  // thread t owns the 16-byte chunk ch = t & 7 of rows (t >> 3) + 32 u of every tile
  const int ch = tid & 7, r0 = tid >> 3;
  const uint32_t swz = static_cast<uint32_t>((ch ^ (r0 & 7)) << 4);   // (row & 7) == (r0 & 7): rows advance by 32
  // rows [row_lo, row_lo + nrows) of the sequence -> tile rows [0, nrows); rows beyond S are zero-filled
  // (token j >= 1 lives at canonical row first + (j-1) * stride: one pointer increment per copy; the 64-bit
  // row(j) * ld products of the first version were 22% of the kernel's instructions, ncu r01p)
  auto gather = [&](int col0, int row_lo, int nrows, uint8_t* tile) {
    const uint32_t dst = smem_u32(tile) + r0 * 128 + swz;
    const uint16_t* src = p.qkv + col0 + ch * 8;
    const uint16_t* tok0 = src + rows.base * p.ld_qkv;
    const uint16_t* g = src + (rows.first + static_cast<long long>(row_lo + r0 - 1) * rows.stride) * p.ld_qkv;
    const long long step = 32LL * rows.stride * p.ld_qkv;
    for (int u = 0, gr = row_lo + r0; u * 32 + r0 < nrows; ++u, gr += 32, g += step) {
      const bool valid = gr < S;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + u * 4096),
                   "l"((gr == 0 || !valid) ? tok0 : g), "r"(valid ? 16u : 0u)
                   : "memory");
    }
  };
  // one 16-byte chunk of the cls token's row (token 0 lives at the clip's base row) -> tile row `trow_` of `tile`
  auto cls_chunk = [&](int col0, uint8_t* tile, int trow_, int c8) {
    const uint16_t* src = p.qkv + rows.base * p.ld_qkv + col0 + c8 * 8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile) + trow_ * 128 +
                                                                   ((c8 ^ (trow_ & 7)) << 4)), "l"(src) : "memory");
  };
  if (TMA) {
    __syncthreads();                   // barrier initialisation visible before the first arrive / complete_tx
    if (tid == 0) {
      const int q0 = min(128, n_box);
      mbar_arrive_expect_tx(ldbar, static_cast<uint32_t>((2 * n_box + q0) * 128));
      tma_load_4d(sK, &tmAll, ldbar, p.d + head * DH, 0, c_t, c_b);
      tma_load_4d(sQ, &tmQ0, ldbar, head * DH, 0, c_t, c_b);
      tma_load_4d(sV, &tmAll, ldbar, 2 * p.d + head * DH, 0, c_t, c_b);
    }
    // rows the boxes do not write: the cls token (last valid row) and the zero padding up to S32 (padded keys are
    // masked with -inf, but 0 * NaN of stale shared memory would still poison P V)
    if (cls_last) {
      if (tid < 8) cls_chunk(p.d + head * DH, sK, S - 1, tid);
      else if (tid < 16) cls_chunk(2 * p.d + head * DH, sV, S - 1, tid - 8);
      else if (tid < 24 && S - 1 < 128) cls_chunk(head * DH, sQ, S - 1, tid - 16);
    }
    for (int i = tid; i < (S32 - S) * 16; i += 256) {            // (S32 - S) rows x 8 chunks x {K, V}
      const int rr = S + (i >> 4), c8 = i & 7;
      uint8_t* tile = (i & 8) ? sV : sK;
      *reinterpret_cast<uint4*>(tile + rr * 128 + (c8 << 4)) = make_uint4(0u, 0u, 0u, 0u);   // all-zero row: any swizzle
    }
  } else {
    gather(p.d + head * DH, 0, S32, sK);
    gather(head * DH, 0, 128, sQ);
    gather(2 * p.d + head * DH, 0, S32, sV);
  }
  sMask[tid] = tid < S ? (p.mask ? p.mask[static_cast<long long>(seq) * S + tok(tid)] * LOG2E : 0.f) : -INFINITY;
  ATTN_TRACE(tid == 0, 1);
  cp_async_wait_all();
  fence_proxy_async();               // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  if (TMA) mbar_wait(ldbar, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  ATTN_TRACE(tid == 0, 2);
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + (static_cast<uint32_t>(lq * 32) << 16);   // this warp's TMEM lane quarter
  const float sl2 = p.scale * LOG2E;
  constexpr int fmt = BF ? 1 : 0;
  const uint32_t idesc_s = make_idesc_f16(fmt, fmt, 0, 0, 128, S32);
  const uint32_t idesc_o = make_idesc_f16(fmt, fmt, 0, 1, 128, DH);
  const int nks = S32 >> 4;            // 16-key steps of the PV contraction
  const int nchunk = S32 >> 5;
  const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), va = smem_u32(sV), pa = smem_u32(sP);
  uint32_t phase = 0;

  for (int qt = 0; qt * 128 < S; ++qt) {
    if (warp == 0) {   // uniform branch; one elected lane issues
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem, make_smem_desc_sw128(qa + ks * 32, 16, 1024), make_smem_desc_sw128(ka + ks * 32, 16, 1024),
                   idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    ATTN_TRACE(tid == 0 && qt < 2, 4 + 10 * qt);
    const bool more = (qt + 1) * 128 < S;
    if (more) {                        // sQ is free again: next query tile lands under the softmax
      if (TMA) {
        const int lo = (qt + 1) * 128;
        if (tid == 0 && n_box > lo) {
          mbar_arrive_expect_tx(qbar, static_cast<uint32_t>((n_box - lo) * 128));
          tma_load_4d(sQ, &tmQ1, qbar, head * DH, lo, c_t, c_b);
        }
        if (cls_last && tid < 8 && S - 1 >= lo) cls_chunk(head * DH, sQ, S - 1 - lo, tid);
      } else {
        gather(head * DH, (qt + 1) * 128, 128, sQ);
      }
    }

    const int row = qt * 128 + rit;    // query row owned by this thread (shared with its partner in the other group)
    // ---- pass 1: maximum over this thread's chunks (base-2 domain, 4 independent chains), then over both groups
    float m;
    {
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      for (int c = wg; c < nchunk; c += 2) {
        uint32_t r[32];
        tmem_ld_32x32(trow + c * 32, r);
        tmem_ld_wait();
        if (MASK || (c + 1) * 32 > S) {   // additive mask, or the chunk that holds the padded keys (mask = -inf)
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 mk = *reinterpret_cast<const float4*>(sMask + c * 32 + j4 * 4);
            m0 = fmaxf(m0, fmaf(__uint_as_float(r[j4 * 4 + 0]), sl2, mk.x));
            m1 = fmaxf(m1, fmaf(__uint_as_float(r[j4 * 4 + 1]), sl2, mk.y));
            m2 = fmaxf(m2, fmaf(__uint_as_float(r[j4 * 4 + 2]), sl2, mk.z));
            m3 = fmaxf(m3, fmaf(__uint_as_float(r[j4 * 4 + 3]), sl2, mk.w));
          }
        } else {                          // no mask: maximum of the raw scores, scaled once (sl2 > 0)
          float x0 = -INFINITY, x1 = -INFINITY, x2 = -INFINITY, x3 = -INFINITY;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            x0 = fmaxf(x0, __uint_as_float(r[j4 * 4 + 0]));
            x1 = fmaxf(x1, __uint_as_float(r[j4 * 4 + 1]));
            x2 = fmaxf(x2, __uint_as_float(r[j4 * 4 + 2]));
            x3 = fmaxf(x3, __uint_as_float(r[j4 * 4 + 3]));
          }
          m0 = fmaxf(m0, fmaxf(fmaxf(x0, x1), fmaxf(x2, x3)) * sl2);
        }
      }
      m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      sStat[wg * 128 + rit] = m;
      __syncthreads();
      m = fmaxf(m, sStat[(wg ^ 1) * 128 + rit]);   // token 0 is never masked: the row maximum is finite
    }
    ATTN_TRACE(tid == 0 && qt < 2, 5 + 10 * qt);
    // ---- pass 2: probabilities, 128 keys at a time -> sP -> O (+)= P V
    float l0 = 0.f, l1 = 0.f;
    for (int half = 0; half * 128 < S32; ++half) {
      const int c0 = half * 4, c1 = min(nchunk, c0 + 4);
      for (int c = c0 + wg; c < c1; c += 2) {
        uint32_t r[32];
        tmem_ld_32x32(trow + c * 32, r);
        tmem_ld_wait();
        float pv[32];
        if (MASK || (c + 1) * 32 > S) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 mk = *reinterpret_cast<const float4*>(sMask + c * 32 + j4 * 4);
            const float mkv[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = j4 * 4 + e;
              pv[j] = ex2(fmaf(__uint_as_float(r[j]), sl2, mkv[e]) - m);   // keys >= S: mask = -inf -> 0
              if (e & 1) l1 += pv[j]; else l0 += pv[j];
            }
          }
        } else {
          const float nm = -m;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            pv[j] = ex2(fmaf(__uint_as_float(r[j]), sl2, nm));
            if (j & 1) l1 += pv[j]; else l0 += pv[j];
          }
        }
        if (DROP) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a0, a1;
            drop_pair(p, S_pad, seq, head, row, c * 16 + j, a0, a1);
            pv[2 * j] *= a0;
            pv[2 * j + 1] *= a1;
          }
        }
        // 32 keys = 4 chunks of 16 bytes in block (c - c0) / 2 of this half
        uint8_t* blk = sP + ((c - c0) >> 1) * 16384 + rit * 128;
        const int cb0 = ((c - c0) & 1) * 4;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 w;
          w.x = pack2<BF>(pv[q4 * 8 + 0], pv[q4 * 8 + 1]); w.y = pack2<BF>(pv[q4 * 8 + 2], pv[q4 * 8 + 3]);
          w.z = pack2<BF>(pv[q4 * 8 + 4], pv[q4 * 8 + 5]); w.w = pack2<BF>(pv[q4 * 8 + 6], pv[q4 * 8 + 7]);
          *reinterpret_cast<uint4*>(blk + (((cb0 + q4) ^ (rit & 7)) << 4)) = w;
        }
      }
      sStat[256 + wg * 128 + rit] = l0 + l1;   // running partial row sum (complete after the last half)
      ATTN_TRACE(tid == 0 && qt < 2 && half < 2, 6 + 10 * qt + 2 * half);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          const int k0 = half * 8, k1 = min(nks, k0 + 8);
          for (int ks = k0; ks < k1; ++ks) {
            const int kl = ks - k0;
            umma_f16(tmem, make_smem_desc_sw128(pa + (kl >> 2) * 16384 + (kl & 3) * 32, 16, 1024),
                     make_smem_desc_sw128(va + ks * 2048, 8192, 1024), idesc_o, ks > 0 ? 1u : 0u);
          }
          umma_commit(bar);
        }
        __syncwarp();
      }
      mbar_wait(bar, phase);           // sP may be overwritten / O may be read after this
      phase ^= 1;
      tc_fence_after();
      ATTN_TRACE(tid == 0 && qt < 2 && half < 2, 7 + 10 * qt + 2 * half);
    }
    // ---- epilogue: this thread's 32 columns of its O row / l -> 16-bit -> staging rows in sP block 0 (the PV MMAs have
    // retired) -> global with 8 lanes per 128-byte row; log-sum-exp for the backward.
    {
      const float l = sStat[256 + rit] + sStat[256 + 128 + rit];
      const float inv = 1.f / l;
      uint32_t r[32];
      tmem_ld_32x32(trow + wg * 32, r);
      tmem_ld_wait();
      uint8_t* myrow = sP + rit * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 w;
        w.x = pack2<BF>(__uint_as_float(r[q4 * 8 + 0]) * inv, __uint_as_float(r[q4 * 8 + 1]) * inv);
        w.y = pack2<BF>(__uint_as_float(r[q4 * 8 + 2]) * inv, __uint_as_float(r[q4 * 8 + 3]) * inv);
        w.z = pack2<BF>(__uint_as_float(r[q4 * 8 + 4]) * inv, __uint_as_float(r[q4 * 8 + 5]) * inv);
        w.w = pack2<BF>(__uint_as_float(r[q4 * 8 + 6]) * inv, __uint_as_float(r[q4 * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(myrow + (((wg * 4 + q4) ^ (rit & 7)) << 4)) = w;
      }
      if (wg == 0 && row < S && p.lse) p.lse[(static_cast<long long>(seq) * p.heads + head) * S + tok(row)] = m + log2f(l);
      tc_fence_before();
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = r0 + 32 * i;
        const int jj = qt * 128 + rr;
        const uint4 v = *reinterpret_cast<const uint4*>(sP + rr * 128 + swz);
        if (jj < S) {
          const int tk = tok(jj);
          uint16_t* dst = (tk == 0 && p.cls_o) ? p.cls_o + static_cast<long long>(seq) * p.d + head * DH
                                               : p.o + rows(tk) * p.ld_o + head * DH;
          *reinterpret_cast<uint4*>(dst + ch * 8) = v;
        }
      }
    }
    ATTN_TRACE(tid == 0 && qt < 2, 10 + 10 * qt);
    if (more) {
      cp_async_wait_all();             // next query tile has landed
      fence_proxy_async();
      if (TMA && n_box > (qt + 1) * 128) mbar_wait(qbar, static_cast<uint32_t>(qt & 1));
    }
    tc_fence_before();
    __syncthreads();                   // every thread is done with TMEM / sP before the next tile reuses them
    tc_fence_after();
    ATTN_TRACE(tid == 0 && qt < 2, 11 + 10 * qt);
  }
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 256);
  }
}

// ------------------------------------------------------------------------------------------------ backward, tcgen05
// Same contract as sattn_bwd_kernel with all five contractions on tcgen05 and every probability computed ONCE (the
// mma.sync kernel recomputes S and dP in two phases: 7 contractions and twice the exponentials).
//
// Orientation: the MMA M dimension is the KEY index, so a softmax thread owns one key row j (TMEM lane) and walks over
// queries; the softmax statistics (lse, D) are per-column broadcasts from shared memory and no row reduction exists:
//   S^T  = K_kt  Q_c^T   M=128 keys, N<=64 queries, K=64     (A = K tile rows, B = Q rows, both K-major, in place)
//   dP^T = V_kt dO_c^T   same shapes                         (A = V tile rows, B = dO rows)
//   P^T  = exp2(S^T*c + mask_j - lse_q),  dS^T = P^T (dP^T - D_q) * scale      -> 16-bit, staged in shared memory as
//          128B-swizzled [128 keys][64 queries] atoms (thread j writes row j)
//   dV_kt += P^T  dO_c   M=128 keys, N=64, K=queries of the chunk   (A = staged atom, K-major; B = dO, MN-major in place)
//   dK_kt += dS^T Q_c    same                                       (A = staged atom;          B = Q,  MN-major in place)
//   dQ_pair += dS K_kt   M=128 queries (two adjacent atoms), N=64, K=keys of the tile
//                        (A = the SAME dS^T atoms read MN-major; B = K tile rows, MN-major in place)
// Steps run over (key tile kt, query chunk c); TMEM (512 columns): two S^T/dP^T stages of 64+64 columns, dV, dK (64+64)
// for the current key tile, dQ for the two 128-query halves (64+64, accumulated over key tiles).
// Warp roles: warps 0..7 = softmax/dS (warp w: TMEM lane quarter w%4, column half w/4) and epilogues; warp 8 lane 0 =
// MMA issuer. The issuer runs one step ahead with S^T/dP^T so the tensor pipe works on step s+1 (and on the gradient
// MMAs of step s-1) while the softmax warps transform step s.
// Shared memory: Q, dO [S16][64]; K, V [nkt*128][64] (zero rows beyond S); P^T ring 2 atoms; dS^T 4 atoms (one per query
// chunk of the current key tile) -> ~215 KB at S=197, one CTA per SM; S <= 240.
// TMA = true (r02): Q, dO, K, V arrive as four tensor-map boxes issued by one thread (see the forward kernel for the
// map and the cls-last tile order); the per-thread gather took 8 k cycles just to ISSUE its ~7 k predicated copies
// (profiles/r01p_sattn_tc_traces.md: "loads issued" at 7951 of 31 k cycles).
template <bool BF, bool DROP, bool TMA>
__global__ void __launch_bounds__(288, 1) sattn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                              const __grid_constant__ CUtensorMap tmDO,
                                                              const SAttnParams p) {
  const long long t_start = p.trace ? clock64() : 0;
  extern __shared__ uint8_t sm_raw[];
  const uint32_t raw_addr = smem_u32(sm_raw);
  uint8_t* sm = sm_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int head = blockIdx.x, seq = blockIdx.y;
  const SeqRows rows = make_rows(p, seq);
  const int S = p.S;
  const int S16 = (S + 15) & ~15;      // also the row pitch of the dropout-mask index (shared with the other kernels)
  const int nkt = (S + 127) >> 7;      // key tiles of 128 rows
  const int nqc = (S16 + 63) >> 6;     // query chunks of 64
  const int nsteps = nkt * nqc;
  const int krows = nkt * 128;
  uint8_t* sQ = sm;
  uint8_t* sG = sQ + S16 * 128;        // dO
  uint8_t* sK = sG + S16 * 128;
  uint8_t* sV = sK + krows * 128;
  uint8_t* sP = sV + krows * 128;      // 2 atoms of [128 keys][64 queries]
  uint8_t* sDS = sP + 2 * 16384;       // 4 atoms (query chunk c of the current key tile)
  float* sMask = reinterpret_cast<float*>(sDS + 4 * 16384);   // [256] additive key mask * log2(e); -inf beyond S
  float* sNl = sMask + 256;            // [256] -(base-2 log-sum-exp) of query q; -inf beyond S
  float* sD = sNl + 256;               // [256] scale * D_q, D_q = rowsum(dO_q * O_q)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 256);
  uint64_t* s_full = bars;             // [2] MMA -> softmax: S^T/dP^T stage written
  uint64_t* s_free = bars + 2;         // [2] softmax -> MMA: stage drained to registers (8 warp arrivals)
  uint64_t* p_ready = bars + 4;        // [2] softmax -> MMA: P^T / dS^T atoms of the step staged (8 warp arrivals)
  uint64_t* g_done = bars + 6;         // [2] MMA -> softmax: gradient MMAs of the step retired (staging reusable)
  uint64_t* acc_full = bars + 8;       //     MMA -> softmax: dV/dK of the key tile complete
  uint64_t* acc_free = bars + 9;       //     softmax -> MMA: dV/dK read out (8 warp arrivals)
  uint64_t* ld_bar = bars + 10;        //     TMA: the four operand boxes have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: role branches stay uniform
  constexpr int fmt = BF ? 1 : 0;
  // tile row -> token (identity unless the shared cls token sits last, see the forward kernel)
  const bool cls_last = TMA && p.cls_last;
  auto tok = [&](int j) { return cls_last ? (j == S - 1 ? 0 : j + 1) : j; };
  const int cls_row = cls_last ? S - 1 : 0;                  // tile row of token 0
  const int n_box = cls_last ? S - 1 : S;
  if (TMA) {
    if (tid == 0) {
      mbar_init(ld_bar, 1);
      fence_barrier_init();
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
    }
    __syncthreads();
  }
  // One CTA per SM and every unit costs the same, so all SMs load, compute and store in lockstep: HBM idles while
  // they compute and is oversubscribed while they load. ALPRO_ATTN_STAGGER=<cycles> starts the first wave in four
  // phases so that the SMs stay out of step for the whole launch (experiment; off by default).
  if (p.stagger > 0) {
    const int lin = blockIdx.y * gridDim.x + blockIdx.x;
    if (lin < 148) {
      const long long until = clock64() + static_cast<long long>(lin & 3) * p.stagger;
      while (clock64() < until) {
      }
    }
  }
  pdl_grid_sync();   // barrier setup above overlapped the previous kernel's tail (common.h)

  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_free[i], 8);
        mbar_init(&p_ready[i], 8);
        mbar_init(&g_done[i], 1);
      }
      mbar_init(acc_full, 1);
      mbar_init(acc_free, 8);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  } else {
    // Thread t owns the 16-byte chunk ch = t & 7 of rows (t >> 3) + 32 u of every tile. ALL global reads of the unit
    // (four gathered tiles, the O rows for D, mask / lse tables) are issued before the first wait: one DRAM round trip.
    const int ch = tid & 7, r0 = tid >> 3;
    const uint32_t swz = static_cast<uint32_t>((ch ^ (r0 & 7)) << 4);   // (row & 7) == (r0 & 7): rows advance by 32
    auto gather = [&](const uint16_t* src, long long ld, int col0, int nrows, uint8_t* tile) {
      const uint32_t dst = smem_u32(tile) + r0 * 128 + swz;
      const uint16_t* g0 = src + rows(r0 < S ? r0 : 0) * ld + col0 + ch * 8;
      const uint16_t* g1 = src + rows(r0 + 32) * ld + col0 + ch * 8;     // only dereferenced when r0 + 32 < S
      const long long step = 32LL * rows.stride * ld;
      for (int u = 0, row = r0; row < nrows; ++u, row += 32) {
        const bool valid = row < S;                                      // rows beyond S: zero-fill (src-size 0)
        const uint16_t* g = (u == 0 || !valid) ? g0 : g1 + (u - 1) * step;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + u * 4096), "l"(g), "r"(valid ? 16u : 0u)
                     : "memory");
      }
    };
    if (TMA) {
      const int c_t = seq % p.seq_div, c_b = seq / p.seq_div;
      if (tid == 0) {
        mbar_arrive_expect_tx(ld_bar, static_cast<uint32_t>(4 * n_box * 128));
        tma_load_4d(sQ, &tmQKV, ld_bar, head * DH, 0, c_t, c_b);
        tma_load_4d(sG, &tmDO, ld_bar, head * DH, 0, c_t, c_b);
        tma_load_4d(sK, &tmQKV, ld_bar, p.d + head * DH, 0, c_t, c_b);
        tma_load_4d(sV, &tmQKV, ld_bar, 2 * p.d + head * DH, 0, c_t, c_b);
      }
      // token 0 (not part of the boxes): the eight threads that own tile row cls_row copy its four rows, so that the
      // D computation below reads chunks this very thread has loaded
      if (cls_last && r0 == (cls_row & 31)) {
        const uint32_t off = cls_row * 128 + ((ch ^ (cls_row & 7)) << 4);
        const uint16_t* q0 = p.qkv + rows.base * p.ld_qkv + head * DH + ch * 8;
        const uint16_t* g0 = p.dout + rows.base * p.ld_o + head * DH + ch * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sQ) + off), "l"(q0) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sG) + off), "l"(g0) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sK) + off), "l"(q0 + p.d) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sV) + off), "l"(q0 + 2 * p.d) : "memory");
      }
      // rows no box writes: zero (all-zero rows look the same under any swizzle)
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (int i = tid; i < (S16 - S) * 8; i += 256) {
        *reinterpret_cast<uint4*>(sQ + (S + (i >> 3)) * 128 + ((i & 7) << 4)) = z;
        *reinterpret_cast<uint4*>(sG + (S + (i >> 3)) * 128 + ((i & 7) << 4)) = z;
      }
      for (int i = tid; i < (krows - S) * 8; i += 256) {
        *reinterpret_cast<uint4*>(sK + (S + (i >> 3)) * 128 + ((i & 7) << 4)) = z;
        *reinterpret_cast<uint4*>(sV + (S + (i >> 3)) * 128 + ((i & 7) << 4)) = z;
      }
    } else {
      gather(p.qkv, p.ld_qkv, head * DH, S16, sQ);
      gather(p.dout, p.ld_o, head * DH, S16, sG);
      gather(p.qkv, p.ld_qkv, p.d + head * DH, krows, sK);
      gather(p.qkv, p.ld_qkv, 2 * p.d + head * DH, krows, sV);
    }
    uint4 ov[8];   // this thread's chunks of the forward outputs O (for D_q = rowsum(dO_q * O_q))
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int row = r0 + 32 * u;
      ov[u] = make_uint4(0u, 0u, 0u, 0u);
      if (row < S) {
        const int tk = tok(row);
        const uint16_t* orow = (tk == 0 && p.seq_div > 1) ? p.cls_fwd + static_cast<long long>(seq) * p.d + head * DH
                                                           : p.o_fwd + rows(tk) * p.ld_o + head * DH;
        ov[u] = *reinterpret_cast<const uint4*>(orow + ch * 8);
      }
    }
    {
      const int j = tid;   // 256 gather threads = 256 table entries
      sMask[j] = j < S ? (p.mask ? p.mask[static_cast<long long>(seq) * S + tok(j)] * LOG2E : 0.f) : -INFINITY;
      sNl[j] = j < S ? -p.lse[(static_cast<long long>(seq) * p.heads + head) * S + tok(j)] : -INFINITY;
    }
    ATTN_TRACE(tid == 0, 1);
    cp_async_wait_all();
    if (TMA) mbar_wait(ld_bar, 0);
    ATTN_TRACE(tid == 0, 2);
    // D from this thread's own chunks of dO (cp.async data of the issuing thread is visible after the wait). The
    // token-0 upstream gradient is rescaled in place first: the group's cls output was the (weighted) mean over frames.
    const float gscale0 = p.seq_div > 1 ? (p.cls_weight ? p.cls_weight[seq] : 1.f / p.seq_div) : 1.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int row = r0 + 32 * u;
      if (row < S16) {   // warp-uniform (a warp covers 4 consecutive rows, S16 is a multiple of 16)
        uint4* gp = reinterpret_cast<uint4*>(sG + row * 128 + swz);
        uint4 gv = *gp;
        uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
        const uint32_t ow[4] = {ov[u].x, ov[u].y, ov[u].z, ov[u].w};
        float dsum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float g0f, g1f, o0f, o1f;
          unpack2<BF>(gw[k], g0f, g1f);
          unpack2<BF>(ow[k], o0f, o1f);
          if (row == cls_row && p.seq_div > 1) {
            gw[k] = pack2c<BF>(g0f * gscale0, g1f * gscale0);
            unpack2<BF>(gw[k], g0f, g1f);
          }
          dsum = fmaf(g0f, o0f, dsum);
          dsum = fmaf(g1f, o1f, dsum);
        }
        if (row == cls_row && p.seq_div > 1) *gp = make_uint4(gw[0], gw[1], gw[2], gw[3]);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 4);
        if (ch == 0) sD[row] = dsum * p.scale;   // pre-scaled: dS = P * (dP * scale - D * scale)
      }
    }
  }
  fence_proxy_async();   // cp.async tiles and the rescaled dO row: generic-proxy writes -> visible to the tensor cores
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t COL_DV = 256, COL_DK = 320, COL_DQ = 384;
  ATTN_TRACE(tid == 0, 3);

  if (warp == 8) {
    // ---------------------------------------------------------------------------------------- MMA issuer
    // The whole warp walks the (uniform) loop and waits on the barriers; one elected lane issues.
    {
      const uint32_t qa = smem_u32(sQ), ga = smem_u32(sG), ka = smem_u32(sK), va = smem_u32(sV);
      const uint32_t pa = smem_u32(sP), da = smem_u32(sDS);
      const uint32_t idesc_g = make_idesc_f16(fmt, fmt, 0, 1, 128, DH);   // A staged (K-major), B in place (MN-major)
      const uint32_t idesc_q = make_idesc_f16(fmt, fmt, 1, 1, 128, DH);   // A = dS^T atoms read MN-major
      auto issue_sdp = [&](int s, int kt, int qc) {
        const int st = s & 1;
        const int N = min(64, S16 - qc * 64);
        const uint32_t idesc = make_idesc_f16(fmt, fmt, 0, 0, 128, N);
        const uint32_t tS = tmem + st * 128, tDP = tS + 64;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tS, make_smem_desc_sw128(ka + kt * 16384 + ks * 32, 16, 1024),
                     make_smem_desc_sw128(qa + qc * 8192 + ks * 32, 16, 1024), idesc, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tDP, make_smem_desc_sw128(va + kt * 16384 + ks * 32, 16, 1024),
                     make_smem_desc_sw128(ga + qc * 8192 + ks * 32, 16, 1024), idesc, ks > 0 ? 1u : 0u);
          umma_commit(&s_full[st]);
        }
        __syncwarp();
      };
      issue_sdp(0, 0, 0);
      for (int s = 0, kt = 0, qc = 0; s < nsteps; ++s) {
        const int st = s & 1;
        const int qn = qc + 1 == nqc ? 0 : qc + 1, ktn = qc + 1 == nqc ? kt + 1 : kt;   // step s + 1
        if (s + 1 < nsteps) {
          if (s + 1 >= 2) {
            mbar_wait(&s_free[(s + 1) & 1], (((s + 1) >> 1) - 1) & 1);
            tc_fence_after();
          }
          issue_sdp(s + 1, ktn, qn);
        }
        mbar_wait(&p_ready[st], (s >> 1) & 1);
        tc_fence_after();
        ATTN_TRACE(lane == 0 && s < 8, 48 + 2 * s);
        if (qc == 0 && kt > 0) {   // dV/dK accumulators of the previous key tile must have been read out
          mbar_wait(acc_free, (kt - 1) & 1);
          tc_fence_after();
        }
        const int N = min(64, S16 - qc * 64);
        const int nks = N >> 4;
        if (elect_one()) {
        for (int ks = 0; ks < nks; ++ks)   // dV += P^T dO_c
          umma_f16(tmem + COL_DV, make_smem_desc_sw128(pa + st * 16384 + ks * 32, 16, 1024),
                   make_smem_desc_sw128(ga + (qc * 64 + ks * 16) * 128, 8192, 1024), idesc_g, (qc > 0 || ks > 0) ? 1u : 0u);
        for (int ks = 0; ks < nks; ++ks)   // dK += dS^T Q_c
          umma_f16(tmem + COL_DK, make_smem_desc_sw128(da + qc * 16384 + ks * 32, 16, 1024),
                   make_smem_desc_sw128(qa + (qc * 64 + ks * 16) * 128, 8192, 1024), idesc_g, (qc > 0 || ks > 0) ? 1u : 0u);
        if ((qc & 1) || qc == nqc - 1) {   // dQ[128-query half] += dS K_kt  (both atoms of the half are staged)
          const int pair = qc >> 1;
          const int nkk = min(8, (S - kt * 128 + 15) >> 4);
          for (int ks = 0; ks < nkk; ++ks)
            umma_f16(tmem + COL_DQ + pair * 64, make_smem_desc_sw128(da + pair * 32768 + ks * 2048, 16384, 1024),
                     make_smem_desc_sw128(ka + (kt * 128 + ks * 16) * 128, 8192, 1024), idesc_q, (kt > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&g_done[st]);
        if (qc == nqc - 1) umma_commit(acc_full);
        }
        __syncwarp();
        ATTN_TRACE(lane == 0 && s < 8, 49 + 2 * s);
        qc = qn;
        kt = ktn;
      }
    }
  } else {
    // ---------------------------------------------------------------------------------------- softmax / dS warps
    const int lq = warp & 3, grp = warp >> 2;
    const int rloc = lq * 32 + lane;                                   // key row inside the tile = TMEM lane
    const uint32_t tlane = tmem + (static_cast<uint32_t>(lq * 32) << 16);
    const float sl2 = p.scale * LOG2E;
    const int fbase = (static_cast<int>(p.dcls_qkv != nullptr));      // token 0 goes to the fp32 cls scratch
    // One 32-row x 64-column fp32 accumulator block of this warp: TMEM -> 16-bit -> this warp's 4 KB staging block
    // (swizzled) -> global with 8 lanes per 128-byte row (4 full lines per store instruction; a lane storing its own
    // row touched 32 lines per instruction and the LSU serialised them: 2-4K cycles per read-out, trace r01p).
    // Token 0 of the shared-cls layout goes to the fp32 scratch straight from the registers.
    auto store_block = [&](uint32_t tacc, uint8_t* stg, int jbase, int coff, uint64_t* release) {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tacc, r0);
      tmem_ld_32x32(tacc + 32, r1);
      tmem_ld_wait();
      if (release) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(release);
      }
      if (jbase + lane == cls_row && fbase) {
        float* dst = p.dcls_qkv + static_cast<long long>(seq) * 3 * p.d + coff;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          dst[c] = __uint_as_float(r0[c]);
          dst[32 + c] = __uint_as_float(r1[c]);
        }
      }
      uint8_t* myrow = stg + lane * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 w, x;
        w.x = pack2<BF>(__uint_as_float(r0[q4 * 8 + 0]), __uint_as_float(r0[q4 * 8 + 1]));
        w.y = pack2<BF>(__uint_as_float(r0[q4 * 8 + 2]), __uint_as_float(r0[q4 * 8 + 3]));
        w.z = pack2<BF>(__uint_as_float(r0[q4 * 8 + 4]), __uint_as_float(r0[q4 * 8 + 5]));
        w.w = pack2<BF>(__uint_as_float(r0[q4 * 8 + 6]), __uint_as_float(r0[q4 * 8 + 7]));
        x.x = pack2<BF>(__uint_as_float(r1[q4 * 8 + 0]), __uint_as_float(r1[q4 * 8 + 1]));
        x.y = pack2<BF>(__uint_as_float(r1[q4 * 8 + 2]), __uint_as_float(r1[q4 * 8 + 3]));
        x.z = pack2<BF>(__uint_as_float(r1[q4 * 8 + 4]), __uint_as_float(r1[q4 * 8 + 5]));
        x.w = pack2<BF>(__uint_as_float(r1[q4 * 8 + 6]), __uint_as_float(r1[q4 * 8 + 7]));
        *reinterpret_cast<uint4*>(myrow + ((q4 ^ (lane & 7)) << 4)) = w;
        *reinterpret_cast<uint4*>(myrow + (((4 + q4) ^ (lane & 7)) << 4)) = x;
      }
      __syncwarp();
      const int c = lane & 7;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3);
        const int jj = jbase + rr;
        const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((c ^ (rr & 7)) << 4));
        if (jj < S && !(jj == cls_row && fbase))
          *reinterpret_cast<uint4*>(p.dqkv + rows(tok(jj)) * p.ld_qkv + coff + c * 8) = v;
      }
      __syncwarp();
    };
    // dV (warps 0-3) / dK (warps 4-7) rows of key tile k. Staging: rows [0,128) of the V / K tiles, which no MMA reads
    // any more when this runs (tile 0 is read out during tile 1; the last tile after every MMA has retired).
    auto store_dkv = [&](int k) {
      ATTN_TRACE(tid == 0, 36 + 3 * (k & 1));
      mbar_wait(acc_full, k & 1);
      tc_fence_after();
      ATTN_TRACE(tid == 0, 37 + 3 * (k & 1));
      store_block(tlane + (grp == 0 ? COL_DV : COL_DK), (grp == 0 ? sV : sK) + lq * 4096, k * 128 + lq * 32,
                  (grp == 0 ? 2 * p.d : p.d) + head * DH, acc_free);
      ATTN_TRACE(tid == 0, 38 + 3 * (k & 1));
    };
    for (int s = 0, kt = 0, qc = 0; s < nsteps; ++s) {
      const int st = s & 1;
      const int N = min(64, S16 - qc * 64);
      const int j = kt * 128 + rloc;                                   // key index of this thread
      const float mk = sMask[j];
      mbar_wait(&s_full[st], (s >> 1) & 1);
      tc_fence_after();
      ATTN_TRACE(tid == 0 && s < 8, 4 + 4 * s);
      uint32_t sv[2][16], dv[2][16];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int sc = grp * 2 + u;                                    // 16-query sub-chunk of the 64-query chunk
        if (sc * 16 < N) {                                             // warp-uniform
          tmem_ld_32x16(tlane + st * 128 + sc * 16, sv[u]);
          tmem_ld_32x16(tlane + st * 128 + 64 + sc * 16, dv[u]);
        }
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[st]);
      ATTN_TRACE(tid == 0 && s < 8, 5 + 4 * s);
      if (s >= 2) mbar_wait(&g_done[st], ((s >> 1) - 1) & 1);          // staging of step s-2 (and older) consumed
      ATTN_TRACE(tid == 0 && s < 8, 6 + 4 * s);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int sc = grp * 2 + u;
        if (sc * 16 < N) {
          const int q0 = qc * 64 + sc * 16;
          float pv[16], ds[16];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 nl = *reinterpret_cast<const float4*>(sNl + q0 + e4 * 4);
            const float4 dd = *reinterpret_cast<const float4*>(sD + q0 + e4 * 4);
            const float nlv[4] = {nl.x, nl.y, nl.z, nl.w}, ddv[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = e4 * 4 + e;
              float pe = ex2(fmaf(__uint_as_float(sv[u][i]), sl2, nlv[e]) + mk);
              float dp = __uint_as_float(dv[u][i]);
              if (DROP) {   // dP flows through the dropout mask of the forward pass; dV uses the dropped P
                float lo, hi;
                drop_pair(p, S16, seq, head, q0 + i, j >> 1, lo, hi);
                const float mq = (j & 1) ? hi : lo;
                ds[i] = pe * fmaf(dp * mq, p.scale, -ddv[e]);
                pe *= mq;
              } else {
                ds[i] = pe * fmaf(dp, p.scale, -ddv[e]);
              }
              pv[i] = pe;
            }
          }
          uint8_t* prow = sP + st * 16384 + rloc * 128;
          uint8_t* drow = sDS + qc * 16384 + rloc * 128;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 w, x;
            w.x = pack2<BF>(pv[h * 8 + 0], pv[h * 8 + 1]); w.y = pack2<BF>(pv[h * 8 + 2], pv[h * 8 + 3]);
            w.z = pack2<BF>(pv[h * 8 + 4], pv[h * 8 + 5]); w.w = pack2<BF>(pv[h * 8 + 6], pv[h * 8 + 7]);
            x.x = pack2<BF>(ds[h * 8 + 0], ds[h * 8 + 1]); x.y = pack2<BF>(ds[h * 8 + 2], ds[h * 8 + 3]);
            x.z = pack2<BF>(ds[h * 8 + 4], ds[h * 8 + 5]); x.w = pack2<BF>(ds[h * 8 + 6], ds[h * 8 + 7]);
            const int off = ((sc * 2 + h) ^ (rloc & 7)) << 4;
            *reinterpret_cast<uint4*>(prow + off) = w;
            *reinterpret_cast<uint4*>(drow + off) = x;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[st]);
      ATTN_TRACE(tid == 0 && s < 8, 7 + 4 * s);

      // The read-out of the previous key tile's dV/dK is deferred until this tile's first chunk has been transformed:
      // the tensor pipe finishes the last gradient MMAs of tile kt-1 meanwhile and the issuer already has work queued.
      if (qc == 0 && kt > 0) store_dkv(kt - 1);
      if (++qc == nqc) {
        qc = 0;
        ++kt;
      }
    }
    store_dkv(nkt - 1);
    // ---- dQ: every MMA has retired (acc_full of the last key tile was committed after the last dQ MMA); staging in
    // the P^T ring
    if (grp * 128 < S)   // warp-uniform
      store_block(tlane + COL_DQ + grp * 64, sP + grp * 16384 + lq * 4096, grp * 128 + lq * 32, head * DH, nullptr);
  }
  ATTN_TRACE(tid == 0, 42);
  tc_fence_before();
  __syncthreads();
  ATTN_TRACE(tid == 0, 43);
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
#undef ATTN_TRACE

// dqkv[group cls row] = sum over the group's seq_div frames of the per-sequence cls-row gradients
__global__ void cls_qkv_reduce_kernel(const float* __restrict__ part, uint16_t* __restrict__ dqkv, long long ld,
                                      long long clip_rows, int groups, int seq_div, int d3, int fmt) {
  pdl_grid_sync();
  const long long total = static_cast<long long>(groups) * d3;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int gidx = static_cast<int>(i / d3);
    const int c = static_cast<int>(i - static_cast<long long>(gidx) * d3);
    float s = 0.f;
    for (int t = 0; t < seq_div; ++t) s += part[(static_cast<long long>(gidx) * seq_div + t) * d3 + c];
    dqkv[gidx * clip_rows * ld + c] = f32_to_16(s, fmt);
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  // ask for the full shared-memory carveout so that two ~100 KB CTAs fit one SM
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(%zu bytes): %s", bytes, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
  }
  return 0;
}

}  // namespace
}  // namespace alpro

using namespace alpro;

namespace alpro {
namespace tattn {   // tcgen05 path for T = 8 (tattn_tc.cu); ALPRO_ENOTSUP = shape / alignment does not fit, fall back
int forward(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int B, int N, int T, int heads, int fmt,
            float scale, cudaStream_t st);
int backward(const void* qkv, int64_t ld_qkv, const void* dout, int64_t ld_dout, void* dqkv, int64_t ld_dqkv, int B,
             int N, int T, int heads, int fmt, float scale, cudaStream_t st);
}  // namespace tattn
int sattn_fwd_ps(const void* qkv, int64_t ld_qkv, void* o, int64_t ld_o, void* cls_o, float* lse, int S, int nseq, int heads,
                 int fmt, int seq_div, int stride, int64_t clip_rows, float scale, cudaStream_t st);
}  // namespace alpro
// The tcgen05 temporal-attention kernels are the default (T = 8: 0.085 / 0.104 ms vs 0.132 / 0.329 ms fwd / bwd at 32
// clips); ALPRO_TATTN_TC=0 selects the CUDA-core kernels (read per call so tests can switch)
static bool want_tattn_tc() {
  const char* e = getenv("ALPRO_TATTN_TC");
  return !(e && e[0] == '0');
}

extern "C" int alpro_temporal_attn_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int B, int N, int T,
                                       int heads, int fmt, float scale, void* stream) {
  ALPRO_REQUIRE(qkv && out && B > 0 && N > 0 && heads > 0, "alpro_temporal_attn_fwd: bad args");
  if (want_tattn_tc()) {
    const int rc = tattn::forward(qkv, ld_qkv, out, ld_out, B, N, T, heads, fmt, scale, static_cast<cudaStream_t>(stream));
    if (rc != ALPRO_ENOTSUP) return rc;
  }
  TAttnParams p{};
  p.qkv = static_cast<const uint16_t*>(qkv); p.out = static_cast<uint16_t*>(out);
  p.ld_qkv = ld_qkv; p.ld_out = ld_out; p.B = B; p.N = N; p.heads = heads; p.d = heads * DH; p.fmt = fmt; p.scale = scale;
  const long long units = static_cast<long long>(B) * N * heads;
  const int wpb = 4;
  const unsigned grid = static_cast<unsigned>(cdiv(units, wpb));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (T) {
    case 1:
      if (fmt) launch_k(tattn_fwd_kernel<1, true>, grid, wpb * 32, 0, st, p);
      else launch_k(tattn_fwd_kernel<1, false>, grid, wpb * 32, 0, st, p);
      break;
    case 2:
      if (fmt) launch_k(tattn_fwd_kernel<2, true>, grid, wpb * 32, 0, st, p);
      else launch_k(tattn_fwd_kernel<2, false>, grid, wpb * 32, 0, st, p);
      break;
    case 4:
      if (fmt) launch_k(tattn_fwd_kernel<4, true>, grid, wpb * 32, 0, st, p);
      else launch_k(tattn_fwd_kernel<4, false>, grid, wpb * 32, 0, st, p);
      break;
    case 8:
      if (fmt) launch_k(tattn_fwd_kernel<8, true>, grid, wpb * 32, 0, st, p);
      else launch_k(tattn_fwd_kernel<8, false>, grid, wpb * 32, 0, st, p);
      break;
    default: set_last_error("alpro_temporal_attn_fwd: T=%d unsupported (1,2,4,8)", T); return ALPRO_ENOTSUP;
  }
  ALPRO_CHECK_LAUNCH("alpro_temporal_attn_fwd");
  return 0;
}

extern "C" int alpro_temporal_attn_bwd(const void* qkv, int64_t ld_qkv, const void* dout, int64_t ld_dout, void* dqkv,
                                       int64_t ld_dqkv, int B, int N, int T, int heads, int fmt, float scale,
                                       void* stream) {
  ALPRO_REQUIRE(qkv && dout && dqkv, "alpro_temporal_attn_bwd: bad args");
  if (want_tattn_tc()) {
    const int rc = tattn::backward(qkv, ld_qkv, dout, ld_dout, dqkv, ld_dqkv, B, N, T, heads, fmt, scale,
                                   static_cast<cudaStream_t>(stream));
    if (rc != ALPRO_ENOTSUP) return rc;
  }
  TAttnParams p{};
  p.qkv = static_cast<const uint16_t*>(qkv); p.out = static_cast<uint16_t*>(dqkv);
  p.dout = static_cast<const uint16_t*>(dout);
  p.ld_qkv = ld_qkv; p.ld_out = ld_dqkv; p.ld_dout = ld_dout;
  p.B = B; p.N = N; p.heads = heads; p.d = heads * DH; p.fmt = fmt; p.scale = scale;
  const long long units = static_cast<long long>(B) * N * heads;
  const int wpb = 4;
  const unsigned grid = static_cast<unsigned>(cdiv(units, wpb));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (T) {
    case 1:
      if (fmt) launch_k(tattn_bwd_kernel<1, true>, grid, wpb * 32, 0, st, p);
      else launch_k(tattn_bwd_kernel<1, false>, grid, wpb * 32, 0, st, p);
      break;
    case 2:
      if (fmt) launch_k(tattn_bwd_kernel<2, true>, grid, wpb * 32, 0, st, p);
      else launch_k(tattn_bwd_kernel<2, false>, grid, wpb * 32, 0, st, p);
      break;
    case 4:
      if (fmt) launch_k(tattn_bwd_kernel<4, true>, grid, wpb * 32, 0, st, p);
      else launch_k(tattn_bwd_kernel<4, false>, grid, wpb * 32, 0, st, p);
      break;
    case 8:
      if (fmt) launch_k(tattn_bwd_kernel<8, true>, grid, wpb * 32, 0, st, p);
      else launch_k(tattn_bwd_kernel<8, false>, grid, wpb * 32, 0, st, p);
      break;
    default: set_last_error("alpro_temporal_attn_bwd: T=%d unsupported (1,2,4,8)", T); return ALPRO_ENOTSUP;
  }
  ALPRO_CHECK_LAUNCH("alpro_temporal_attn_bwd");
  return 0;
}

static long long* g_trace = nullptr;   // ALPRO_ATTN_TRACE=1 diagnostics buffer (64 stamps per CTA of the last traced launch)
static size_t g_trace_len = 0;

extern "C" int alpro_debug_attn_trace(void* host_out, int64_t max_values) {
  ALPRO_REQUIRE(host_out && max_values > 0, "alpro_debug_attn_trace: bad args");
  if (!g_trace) return 0;
  const size_t n = g_trace_len < static_cast<size_t>(max_values) ? g_trace_len : static_cast<size_t>(max_values);
  cudaError_t e = cudaMemcpy(host_out, g_trace, n * sizeof(long long), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    set_last_error("alpro_debug_attn_trace: %s", cudaGetErrorString(e));
    return -1;
  }
  return static_cast<int>(n / 64);
}

// ALPRO_ATTN_TRACE=1 (diagnostics only): zeroed per-CTA stamp buffer for the next tcgen05 attention launch, else null
static long long* trace_buffer(size_t ctas, cudaStream_t st) {
  const char* tr_env = getenv("ALPRO_ATTN_TRACE");
  if (!(tr_env && tr_env[0] == '1')) return nullptr;
  const size_t need = ctas * 64;
  if (need > g_trace_len) {
    if (g_trace) cudaFree(g_trace);
    g_trace = nullptr;
    g_trace_len = 0;
    if (cudaMalloc(&g_trace, need * sizeof(long long)) == cudaSuccess) g_trace_len = need;
  }
  if (g_trace) cudaMemsetAsync(g_trace, 0, g_trace_len * sizeof(long long), st);
  return g_trace;
}

// ---------------------------------------------------------------------------------------------- tensor maps (r02)
typedef CUresult (*SeqEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static SeqEncodeFn seq_encode_fn() {
  static SeqEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<SeqEncodeFn>(f);
    return static_cast<SeqEncodeFn>(nullptr);
  }();
  return fn;
}

// 4-D map (column, token, frame, clip) over a row-major 16-bit matrix [rows, ld] holding the sequences of the
// SAttnParams row rule: token i of the box is row  clip*clip_rows + first_row + frame + i*stride. Box = 64 columns x
// box_rows tokens of one (frame, clip), 128-byte swizzle (= the UMMA K-major SW128 layout of a [box_rows][64] tile).
static int make_seq_map(CUtensorMap* m, const uint16_t* mat, int64_t ld, int cols, int first_row, int n_tok, int stride,
                        int seq_div, int groups, int64_t clip_rows, int box_rows) {
  SeqEncodeFn enc = seq_encode_fn();
  if (!enc) return ALPRO_ENOTSUP;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(n_tok),
                        static_cast<cuuint64_t>(seq_div), static_cast<cuuint64_t>(groups)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(stride) * ld * 2, static_cast<cuuint64_t>(ld) * 2,
                           static_cast<cuuint64_t>(clip_rows) * ld * 2};
  cuuint32_t box[4] = {DH, static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  void* base = const_cast<uint16_t*>(mat + static_cast<int64_t>(first_row) * ld);
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(4d) failed (%d): cols=%d tok=%d div=%d groups=%d ld=%lld box=%d", (int)r, cols,
                   n_tok, seq_div, groups, (long long)ld, box_rows);
    return ALPRO_EINVAL;
  }
  return 0;
}

// Tensor-map path usable? (ALPRO_ATTN_TMA=0 switches it off; read per call so tests can compare both paths)
// Measured (r02h, tools/check_attn_tc.py, 256x12 sequences of 197 / 128x12 of 237): strided ViT layout 0.3758 -> 0.3648 ms
// backward, 0.1425 -> 0.1417 ms forward; contiguous BERT layout 0.3054 -> 0.3113 / 0.1723 -> 0.1764 ms. The gain is
// small because the kernels are bound by the lock-step DRAM round trip of the whole wave, not by issuing the copies;
// default: tensor maps for the strided layout, per-thread gather for contiguous rows; ALPRO_ATTN_TMA=1 forces maps.
static bool seq_tma_wanted(const SAttnParams& p) {
  const char* e = getenv("ALPRO_ATTN_TMA");
  if (e && e[0] == '0') return false;
  if (!seq_encode_fn()) return false;
  if (!aligned16(p.qkv) || (p.ld_qkv % 8) != 0 || 3 * p.d > p.ld_qkv || p.nseq % p.seq_div != 0) return false;
  const bool cls_last = !(p.seq_div == 1 && p.stride == 1);
  if (cls_last && (p.drop_thr != 0 || p.S < 2)) return false;   // dropout counters are defined in token order
  if (!cls_last && !(e && e[0] == '1')) return false;
  return true;
}

static int fill_sattn(SAttnParams& p, const void* qkv, int64_t ld_qkv, const float* mask, int S, int nseq, int heads,
                      int fmt, int seq_div, int stride, int64_t clip_rows, float scale, float drop_p,
                      uint32_t drop_seed) {
  ALPRO_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "alpro_seq_attn: dropout probability out of range");
  p.drop_thr = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  p.drop_seed = drop_seed;
  p.drop_scale = 1.0f / (1.0f - drop_p);
  ALPRO_REQUIRE(qkv && S > 0 && S <= 256 && nseq > 0 && heads > 0, "alpro_seq_attn: bad args (S=%d must be <= 256)", S);
  ALPRO_REQUIRE(seq_div >= 1 && stride >= 1 && (ld_qkv % 8) == 0, "alpro_seq_attn: bad layout");
  p.qkv = static_cast<const uint16_t*>(qkv); p.ld_qkv = ld_qkv; p.mask = mask; p.S = S; p.nseq = nseq; p.heads = heads;
  p.d = heads * DH; p.fmt = fmt; p.seq_div = seq_div; p.stride = stride; p.clip_rows = clip_rows; p.scale = scale;
  return 0;
}

extern "C" int alpro_seq_attn_fwd(const void* qkv, int64_t ld_qkv, const float* mask, void* o, int64_t ld_o,
                                  void* cls_o, float* lse, int S, int nseq, int heads, int fmt, int seq_div, int stride,
                                  int64_t clip_rows, float scale, float drop_p, uint32_t drop_seed, void* stream) {
  SAttnParams p{};
  int rc = fill_sattn(p, qkv, ld_qkv, mask, S, nseq, heads, fmt, seq_div, stride, clip_rows, scale, drop_p, drop_seed);
  if (rc) return rc;
  ALPRO_REQUIRE(o && (ld_o % 8) == 0, "alpro_seq_attn_fwd: bad output");
  p.o = static_cast<uint16_t*>(o); p.ld_o = ld_o; p.cls_o = static_cast<uint16_t*>(cls_o); p.lse = lse;
  const int S_pad = (S + 15) & ~15;
  dim3 grid(heads, nseq);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // tcgen05 / TMEM forward (same outputs and lse format as the mma.sync kernel; default for S >= 96: 0.143 vs 0.197 ms
  // at 256x12 sequences of 197). ALPRO_ATTN_TC=0 forces mma.sync, =1 forces tcgen05; read per call so tests can switch.
  const char* tc_env = getenv("ALPRO_ATTN_TC");
  if (tc_env && (tc_env[0] == '0' || tc_env[0] == '1') ? tc_env[0] == '1' : S >= 96) {
    // persistent warp-specialised kernel for the strided two-tile layout without mask / dropout (sattn_ps.cu)
    if (!mask && !p.drop_thr &&
        sattn_fwd_ps(qkv, ld_qkv, o, ld_o, cls_o, lse, S, nseq, heads, fmt, seq_div, stride, clip_rows, scale, st) == ALPRO_OK) {
      ALPRO_CHECK_LAUNCH("alpro_seq_attn_fwd(persistent)");
      return 0;
    }
    const size_t S32 = (S + 31) & ~31;
    const size_t smem_tc = 1024 + 128 * 128 + S32 * 128 * 2 + 2 * 16384 + (256 + 512) * sizeof(float) + 64;
    p.trace = trace_buffer(static_cast<size_t>(heads) * nseq, st);
    CUtensorMap tmAll, tmQ0, tmQ1;
    memset(&tmAll, 0, sizeof(tmAll));
    bool tma = seq_tma_wanted(p);
    if (tma) {
      p.cls_last = !(seq_div == 1 && stride == 1);
      const int n_box = p.cls_last ? S - 1 : S, first = p.cls_last ? 1 : 0, groups = nseq / seq_div;
      rc = make_seq_map(&tmAll, p.qkv, ld_qkv, 3 * p.d, first, n_box, stride, seq_div, groups, clip_rows, n_box);
      if (!rc) rc = make_seq_map(&tmQ0, p.qkv, ld_qkv, 3 * p.d, first, n_box, stride, seq_div, groups, clip_rows,
                                 n_box < 128 ? n_box : 128);
      if (!rc && n_box > 128)
        rc = make_seq_map(&tmQ1, p.qkv, ld_qkv, 3 * p.d, first, n_box, stride, seq_div, groups, clip_rows, n_box - 128);
      if (rc) { tma = false; p.cls_last = 0; rc = 0; }          // descriptor not encodable: per-thread gather
      else if (n_box <= 128) tmQ1 = tmQ0;
    }
    if (!tma) tmQ0 = tmQ1 = tmAll;
#define LAUNCH_FWD_TC(BF, DR, MK)                                                                   \
  do {                                                                                              \
    if (tma) {                                                                                      \
      rc = set_smem(sattn_fwd_tc_kernel<BF, DR, MK, true>, smem_tc);                                \
      if (rc) return rc;                                                                            \
      launch_k(sattn_fwd_tc_kernel<BF, DR, MK, true>, grid, 256, smem_tc, st, tmAll, tmQ0, tmQ1, p);      \
    } else {                                                                                        \
      rc = set_smem(sattn_fwd_tc_kernel<BF, DR, MK, false>, smem_tc);                               \
      if (rc) return rc;                                                                            \
      launch_k(sattn_fwd_tc_kernel<BF, DR, MK, false>, grid, 256, smem_tc, st, tmAll, tmQ0, tmQ1, p);     \
    }                                                                                               \
  } while (0)
#define LAUNCH_FWD_TC_F(BF)                                                          \
  do {                                                                               \
    if (p.drop_thr) { if (mask) LAUNCH_FWD_TC(BF, true, true); else LAUNCH_FWD_TC(BF, true, false); }   \
    else            { if (mask) LAUNCH_FWD_TC(BF, false, true); else LAUNCH_FWD_TC(BF, false, false); } \
  } while (0)
    if (fmt == 1) LAUNCH_FWD_TC_F(true); else LAUNCH_FWD_TC_F(false);
#undef LAUNCH_FWD_TC_F
#undef LAUNCH_FWD_TC
    ALPRO_CHECK_LAUNCH("alpro_seq_attn_fwd(tcgen05)");
    return 0;
  }
  const size_t smem = static_cast<size_t>(S_pad) * 128 * 3 + S_pad * sizeof(float);
#define LAUNCH_FWD(BF, NT, EX)                                               \
  do {                                                                       \
    rc = set_smem(sattn_fwd_kernel<BF, NT, EX>, smem);                       \
    if (rc) return rc;                                                       \
    launch_k(sattn_fwd_kernel<BF, NT, EX>, grid, 128, smem, st, p);                \
  } while (0)
#define LAUNCH_FWD_T(BF)                                                                              \
  do {                                                                                                \
    const int ntiles = S_pad >> 3;                                                                    \
    if (ntiles == 26) LAUNCH_FWD(BF, 26, true);        /* 1 + 196 patches (224^2)            */       \
    else if (ntiles == 30) LAUNCH_FWD(BF, 30, true);   /* fusion: 40 text + 197 video tokens */       \
    else if (ntiles == 6) LAUNCH_FWD(BF, 6, true);     /* text, L = 40                       */       \
    else if (ntiles <= 8) LAUNCH_FWD(BF, 8, false);                                                   \
    else if (ntiles <= 16) LAUNCH_FWD(BF, 16, false);                                                 \
    else LAUNCH_FWD(BF, 32, false);                                                                   \
  } while (0)
  if (fmt == 1) LAUNCH_FWD_T(true); else LAUNCH_FWD_T(false);
#undef LAUNCH_FWD_T
#undef LAUNCH_FWD
  ALPRO_CHECK_LAUNCH("alpro_seq_attn_fwd");
  return 0;
}

extern "C" int alpro_seq_attn_bwd(const void* qkv, int64_t ld_qkv, const float* mask, const float* lse, const void* o_fwd,
                                  const void* cls_fwd, const float* cls_weight, const void* dout, int64_t ld_o,
                                  void* dqkv, float* dcls_qkv_scratch, int S, int nseq,
                                  int heads, int fmt, int seq_div, int stride, int64_t clip_rows, float scale,
                                  float drop_p, uint32_t drop_seed, void* stream) {
  SAttnParams p{};
  int rc = fill_sattn(p, qkv, ld_qkv, mask, S, nseq, heads, fmt, seq_div, stride, clip_rows, scale, drop_p, drop_seed);
  if (rc) return rc;
  ALPRO_REQUIRE(lse && dout && dqkv && o_fwd && (ld_o % 8) == 0, "alpro_seq_attn_bwd: bad args");
  ALPRO_REQUIRE(seq_div == 1 || cls_fwd, "alpro_seq_attn_bwd: shared-cls layout needs the per-sequence cls outputs");
  p.o_fwd = static_cast<const uint16_t*>(o_fwd);
  p.cls_fwd = static_cast<const uint16_t*>(cls_fwd);
  p.cls_weight = cls_weight;
  ALPRO_REQUIRE(seq_div == 1 || dcls_qkv_scratch, "alpro_seq_attn_bwd: shared-cls layout needs the fp32 scratch");
  p.lse = const_cast<float*>(lse); p.dout = static_cast<const uint16_t*>(dout); p.ld_o = ld_o;
  p.dqkv = static_cast<uint16_t*>(dqkv);
  p.dcls_qkv = seq_div > 1 ? dcls_qkv_scratch : nullptr;
  const int S_pad = (S + 15) & ~15;
  const size_t smem = static_cast<size_t>(S_pad) * 128 * 4 + 3 * S_pad * sizeof(float);
  dim3 grid(heads, nseq);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // tcgen05 / TMEM backward (default for 96 <= S <= 240: 0.58 vs 0.75 ms at 256x12 sequences of 197, 0.44 vs 0.81 ms at
  // 128x12 of 237 with dropout; the per-CTA fixed cost makes it slower than mma.sync for short text sequences).
  // ALPRO_ATTN_BWD_TC=0 forces the mma.sync kernel, =1 forces tcgen05 wherever it fits (S <= 240); read per call so
  // tests can switch implementations.
  const char* tc_env = getenv("ALPRO_ATTN_BWD_TC");
  const bool use_tc = S <= 240 && (tc_env && (tc_env[0] == '0' || tc_env[0] == '1') ? tc_env[0] == '1' : S >= 96);
  if (use_tc) {   // same inputs, outputs and dropout stream as the mma.sync kernel
    p.trace = trace_buffer(static_cast<size_t>(heads) * nseq, st);
    {
      // first wave in four phases, 8000 cycles apart: the SMs stay out of step for the whole launch instead of loading
      // and computing in lockstep (r02n, 256x12 sequences of 197: 0.370 -> 0.342 ms; 3000 / 5000 cycles: 0.350)
      const char* sg = getenv("ALPRO_ATTN_STAGGER");
      p.stagger = sg ? atoi(sg) : 8000;
    }
    const size_t krows = static_cast<size_t>((S + 127) / 128) * 128;
    const size_t smem_tc = 1024 + 2 * static_cast<size_t>(S_pad) * 128 + 2 * krows * 128 + 6 * 16384 +
                           3 * 256 * sizeof(float) + 11 * sizeof(uint64_t) + 16;
    CUtensorMap tmQKV, tmDO;
    memset(&tmQKV, 0, sizeof(tmQKV));
    bool tma = seq_tma_wanted(p) && aligned16(dout);
    if (tma) {
      p.cls_last = !(seq_div == 1 && stride == 1);
      const int n_box = p.cls_last ? S - 1 : S, first = p.cls_last ? 1 : 0, groups = nseq / seq_div;
      rc = make_seq_map(&tmQKV, p.qkv, ld_qkv, 3 * p.d, first, n_box, stride, seq_div, groups, clip_rows, n_box);
      if (!rc) rc = make_seq_map(&tmDO, p.dout, ld_o, p.d, first, n_box, stride, seq_div, groups, clip_rows, n_box);
      if (rc) { tma = false; p.cls_last = 0; rc = 0; }
    }
    if (!tma) tmDO = tmQKV;
#define LAUNCH_BWD_TC(BF, DR)                                                             \
  do {                                                                                    \
    if (tma) {                                                                            \
      rc = set_smem(sattn_bwd_tc_kernel<BF, DR, true>, smem_tc);                          \
      if (rc) return rc;                                                                  \
      launch_k(sattn_bwd_tc_kernel<BF, DR, true>, grid, 288, smem_tc, st, tmQKV, tmDO, p);      \
    } else {                                                                              \
      rc = set_smem(sattn_bwd_tc_kernel<BF, DR, false>, smem_tc);                         \
      if (rc) return rc;                                                                  \
      launch_k(sattn_bwd_tc_kernel<BF, DR, false>, grid, 288, smem_tc, st, tmQKV, tmDO, p);     \
    }                                                                                     \
  } while (0)
    if (fmt == 1) {
      if (p.drop_thr) LAUNCH_BWD_TC(true, true); else LAUNCH_BWD_TC(true, false);
    } else {
      if (p.drop_thr) LAUNCH_BWD_TC(false, true); else LAUNCH_BWD_TC(false, false);
    }
#undef LAUNCH_BWD_TC
    ALPRO_CHECK_LAUNCH("alpro_seq_attn_bwd(tcgen05)");
  } else if (fmt == 1) {
    rc = set_smem(sattn_bwd_kernel<true>, smem);
    if (rc) return rc;
    launch_k(sattn_bwd_kernel<true>, grid, 256, smem, st, p);
  } else {
    rc = set_smem(sattn_bwd_kernel<false>, smem);
    if (rc) return rc;
    launch_k(sattn_bwd_kernel<false>, grid, 256, smem, st, p);
  }
  ALPRO_CHECK_LAUNCH("alpro_seq_attn_bwd");
  if (seq_div > 1) {
    const int groups = nseq / seq_div;
    const long long total = static_cast<long long>(groups) * 3 * p.d;
    long long g = cdiv(total, 256);
    launch_k(cls_qkv_reduce_kernel, static_cast<unsigned>(g), 256, 0, st, dcls_qkv_scratch, p.dqkv, ld_qkv, clip_rows, groups,
                                                                    seq_div, 3 * p.d, fmt);
    ALPRO_CHECK_LAUNCH("alpro_seq_attn_bwd(cls reduce)");
  }
  return 0;
}

namespace alpro {
namespace {
__global__ void attn_drop_mask_kernel(float* __restrict__ out, int S, int S_pad, int nseq, int heads, uint32_t thr,
                                      uint32_t seed, float scale) {
  pdl_grid_sync();
  SAttnParams p{};
  p.heads = heads; p.drop_thr = thr; p.drop_seed = seed; p.drop_scale = scale;
  const long long total = static_cast<long long>(nseq) * heads * S * S;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % S);
    const int q = static_cast<int>((i / S) % S);
    const long long sh = i / (static_cast<long long>(S) * S);
    float lo, hi;
    drop_pair(p, S_pad, static_cast<int>(sh / heads), static_cast<int>(sh % heads), q, j >> 1, lo, hi);
    out[i] = (j & 1) ? hi : lo;
  }
}
}  // namespace
}  // namespace alpro

extern "C" int alpro_attn_dropout_mask(float* out, int S, int nseq, int heads, float drop_p, uint32_t drop_seed,
                                       void* stream) {
  ALPRO_REQUIRE(out && S > 0 && nseq > 0 && heads > 0 && drop_p > 0.f && drop_p < 1.f, "alpro_attn_dropout_mask: bad args");
  const int S_pad = (S + 15) & ~15;
  launch_k(attn_drop_mask_kernel, num_sms() * 8, 256, 0, static_cast<cudaStream_t>(stream), 
      out, S, S_pad, nseq, heads, drop_threshold(drop_p), drop_seed, 1.0f / (1.0f - drop_p));
  ALPRO_CHECK_LAUNCH("alpro_attn_dropout_mask");
  return 0;
}
