// Temporal attention (TimeSformer divided attention over the T = 8 frames of one patch position) on tcgen05.
//
// Reference: Attention.forward vit.py:81-100 called through Block.forward vit.py:146-157 on '(b h w) t m'.
//
// The 8x8 problem of one (clip, patch, head) unit is far too small for a tensor-core tile, but 16 units are 128
// CONSECUTIVE rows of the canonical [B (1 + N T), 3d] qkv matrix, so one TMA box [128 rows x 64 columns] per operand
// stages 16 units at once and
//     S = Q K^T   (M = 128, N = 128, K = 64)
// holds the 16 wanted 8x8 score blocks on its diagonal. The other 15/16 of the tile are wasted tensor work
// (2 MFLOP = 128 cycles per 16 units: nothing next to the 64 KB of HBM traffic the group costs), and in exchange the
// CUDA cores only touch 8 scores per row: thread r owns row r, reads its diagonal block from TMEM, does the softmax in
// registers and writes the 8 NORMALISED probabilities as one 16-byte chunk of a block-diagonal 128x128 16-bit tile in
// shared memory (the off-diagonal chunks are zeroed once per kernel). Then
//     O = P V     (M = 128, N = 64, K = 128; V in place as an MN-major B operand)
// Backward (all five contractions, same trick): S = Q K^T and dP = dO V^T give thread q its 8 probabilities and
// D_q = sum_j P dP in registers; P and dS = P (dP - D) scale are staged as block-diagonal tiles and
//     dQ = dS K,   dV = P^T dO,   dK = dS^T Q
// read them K-major (dQ) or MN-major (dV, dK: the transposed views of the same bytes); K, dO, Q are MN-major B operands
// in place. The mma.sync-free CUDA-core kernels in attention.cu did this with ~1100 instructions per unit and reached
// 2.5 / 1.6 TB/s; this version is a persistent TMA -> tcgen05 -> TMA-store pipeline bound by HBM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (elected lane, warp-uniform control flow),
// warps 2..5 = softmax + epilogue (TMEM lane quarter = warp % 4).
#include <cuda.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace alpro {
namespace tattn {

constexpr int DH = 64;
constexpr int ROWS = 128;                 // rows of one group = 16 units x 8 frames
constexpr int TILE = ROWS * DH * 2;       // 16 KB: one [128][64] 16-bit operand tile
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <bool BF>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  if (BF) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// this lane's 8 diagonal values out of the 32 columns its warp loaded (block index k = lane / 8)
__device__ __forceinline__ void pick8(const uint32_t (&r)[32], int k, float (&v)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const uint32_t a = k & 1 ? r[8 + e] : r[e];
    const uint32_t b = k & 1 ? r[24 + e] : r[16 + e];
    v[e] = __uint_as_float(k & 2 ? b : a);
  }
}

struct Params {
  int B, G, heads, d;       // clips, groups of 128 rows per clip, heads, model width
  int n_items;              // B * G * heads
  long long clip_rows;      // 1 + N * T
  uint16_t* out;            // forward: o;  backward: dqkv   (cls rows are zero-filled here)
  long long ld_out;
  float scale;
};

// item -> (clip b, group g, head h): heads fastest so that concurrently running CTAs read adjacent 128-byte slices of
// the same qkv rows
__device__ __forceinline__ void decode_item(const Params& p, int item, int& b, int& g, int& h) {
  h = item % p.heads;
  const int bg = item / p.heads;
  g = bg % p.G;
  b = bg / p.G;
}

// ------------------------------------------------------------------------------------------------ forward
namespace fwd {
constexpr int NST = 3;                                  // operand stages (Q, K, V tiles = 48 KB each): two items of
                                                        // loads in flight while a third is consumed
constexpr int STAGE_BYTES = 3 * TILE;
constexpr int P_BYTES = 2 * TILE;                       // block-diagonal [128][128] 16-bit = two 64-key atoms
constexpr int OFF_P = NST * STAGE_BYTES;                // 2 P buffers
constexpr int OFF_O = OFF_P + 2 * P_BYTES;              // 1 O staging tile (TMA store source)
constexpr int OFF_BAR = OFF_O + TILE;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
constexpr int THREADS = 192;

template <bool BF>
__global__ void __launch_bounds__(THREADS, 1)
tattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = sm + OFF_P;
  uint8_t* sO = sm + OFF_O;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* full = bars;            // [NST] TMA -> MMA
  uint64_t* empty = bars + NST;     // [NST] MMA (PV retired) -> TMA
  uint64_t* s_full = bars + 2 * NST;  // [2] MMA -> softmax: scores in TMEM
  uint64_t* p_ready = s_full + 2;   // [2] softmax -> MMA: P tile staged (4 warp arrivals)
  uint64_t* o_full = p_ready + 2;   // [2] MMA -> epilogue: O in TMEM
  uint64_t* o_free = o_full + 2;    // [2] epilogue -> MMA: TMEM stage read out (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int fmt = BF ? 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmO);
      for (int i = 0; i < NST; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_ready[i], 4);
        mbar_init(&o_full[i], 1);
        mbar_init(&o_free[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // the P tiles are block diagonal: zero them once, the softmax threads only ever rewrite the diagonal chunks
  for (int i = tid; i < 2 * P_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_grid_sync();   // prologue done; global memory is touched below only (common.h)

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    for (int item = blockIdx.x, k = 0; item < p.n_items; item += gridDim.x, ++k) {
      const int s = k % NST;
      int b, g, h;
      decode_item(p, item, b, g, h);
      mbar_wait(&empty[s], ((k / NST) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* st = sm + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        const int row = 1 + g * ROWS;
        tma_load_3d(st, &tmQKV, &full[s], h * DH, row, b);
        tma_load_3d(st + TILE, &tmQKV, &full[s], p.d + h * DH, row, b);
        tma_load_3d(st + 2 * TILE, &tmQKV, &full[s], 2 * p.d + h * DH, row, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc_s = make_idesc_f16(fmt, fmt, 0, 0, ROWS, ROWS);   // S = Q K^T
    const uint32_t idesc_o = make_idesc_f16(fmt, fmt, 0, 1, ROWS, DH);     // O = P V (V MN-major in place)
    const uint32_t sm_a = smem_u32(sm), p_a = smem_u32(sP);
    auto issue_qk = [&](int k) {
      const int s = k % NST, a = k & 1;
      mbar_wait(&full[s], (k / NST) & 1);
      if (k >= 2) mbar_wait(&o_free[a], ((k >> 1) - 1) & 1);   // TMEM stage a read out by item k-2
      tc_fence_after();
      if (elect_one()) {
        const uint32_t qa = sm_a + s * STAGE_BYTES, ka = qa + TILE;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem + a * 256, make_smem_desc_sw128(qa + ks * 32, 16, 1024),
                   make_smem_desc_sw128(ka + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(&s_full[a]);
      }
      __syncwarp();
    };
    if (static_cast<int>(blockIdx.x) < p.n_items) issue_qk(0);
    for (int item = blockIdx.x, k = 0; item < p.n_items; item += gridDim.x, ++k) {
      const int s = k % NST, a = k & 1;
      if (item + static_cast<int>(gridDim.x) < p.n_items) issue_qk(k + 1);   // scores of the next item under this softmax
      mbar_wait(&p_ready[a], (k >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t va = sm_a + s * STAGE_BYTES + 2 * TILE, pa = p_a + a * P_BYTES;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_f16(tmem + a * 256 + 128, make_smem_desc_sw128(pa + (ks >> 2) * TILE + (ks & 3) * 32, 16, 1024),
                   make_smem_desc_sw128(va + ks * 2048, 8192, 1024), idesc_o, ks > 0 ? 1u : 0u);
        umma_commit(&o_full[a]);
        umma_commit(&empty[s]);   // all three operand tiles of the stage have been consumed
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue (warps 2..5)
    const int lq = warp & 3;                      // TMEM lane quarter
    const int r = lq * 32 + lane;                 // row of the group
    const int etid = tid - 64;                    // 0..127 inside the group of four warps
    const uint32_t tlane = tmem + (static_cast<uint32_t>(lq * 32) << 16);
    const float sl2 = p.scale * LOG2E;
    const uint32_t p_off = (r >> 6) * TILE + r * 128 + ((((r >> 3) & 7) ^ (r & 7)) << 4);   // this row's diagonal chunk
    // scores of item k -> normalised probabilities -> block-diagonal P tile
    auto softmax_item = [&](int item, int k) {
      const int a = k & 1;
      int b, g, h;
      decode_item(p, item, b, g, h);
      if (g == 0 && etid < 8)   // cls row of the clip: the temporal branch leaves it at zero
        *reinterpret_cast<uint4*>(p.out + b * p.clip_rows * p.ld_out + h * DH + etid * 8) = make_uint4(0u, 0u, 0u, 0u);
      mbar_wait(&s_full[a], (k >> 1) & 1);
      tc_fence_after();
      uint32_t sr[32];
      tmem_ld_32x32(tlane + a * 256 + lq * 32, sr);
      tmem_ld_wait();
      float v[8];
      pick8(sr, lane >> 3, v);
      float mx = v[0];
#pragma unroll
      for (int e = 1; e < 8; ++e) mx = fmaxf(mx, v[e]);
      float sum = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] = ex2((v[e] - mx) * sl2);
        sum += v[e];
      }
      const float inv = 1.f / sum;
      uint4 w;
      w.x = pack2<BF>(v[0] * inv, v[1] * inv); w.y = pack2<BF>(v[2] * inv, v[3] * inv);
      w.z = pack2<BF>(v[4] * inv, v[5] * inv); w.w = pack2<BF>(v[6] * inv, v[7] * inv);
      // P buffer a was last read by the PV MMAs of item k-2, whose completion (o_full) this thread observed in that
      // item's epilogue, which precedes this call in program order
      *reinterpret_cast<uint4*>(sP + a * P_BYTES + p_off) = w;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[a]);
    };
    // O row of item k -> 16-bit -> swizzled staging tile -> TMA store (clips the rows beyond the clip's end)
    auto store_item = [&](int item, int k) {
      const int a = k & 1;
      int b, g, h;
      decode_item(p, item, b, g, h);
      mbar_wait(&o_full[a], (k >> 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(tlane + a * 256 + 128, o0);
      tmem_ld_32x32(tlane + a * 256 + 160, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[a]);
      if (etid == 0) bulk_wait_read<0>();   // the store of the previous item has drained the staging tile
      named_bar_sync(1, 128);
      uint8_t* orow = sO + r * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 x, y;
        x.x = pack2<BF>(__uint_as_float(o0[q4 * 8 + 0]), __uint_as_float(o0[q4 * 8 + 1]));
        x.y = pack2<BF>(__uint_as_float(o0[q4 * 8 + 2]), __uint_as_float(o0[q4 * 8 + 3]));
        x.z = pack2<BF>(__uint_as_float(o0[q4 * 8 + 4]), __uint_as_float(o0[q4 * 8 + 5]));
        x.w = pack2<BF>(__uint_as_float(o0[q4 * 8 + 6]), __uint_as_float(o0[q4 * 8 + 7]));
        y.x = pack2<BF>(__uint_as_float(o1[q4 * 8 + 0]), __uint_as_float(o1[q4 * 8 + 1]));
        y.y = pack2<BF>(__uint_as_float(o1[q4 * 8 + 2]), __uint_as_float(o1[q4 * 8 + 3]));
        y.z = pack2<BF>(__uint_as_float(o1[q4 * 8 + 4]), __uint_as_float(o1[q4 * 8 + 5]));
        y.w = pack2<BF>(__uint_as_float(o1[q4 * 8 + 6]), __uint_as_float(o1[q4 * 8 + 7]));
        *reinterpret_cast<uint4*>(orow + ((q4 ^ (r & 7)) << 4)) = x;
        *reinterpret_cast<uint4*>(orow + (((4 + q4) ^ (r & 7)) << 4)) = y;
      }
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (etid == 0) {
        tma_store_3d(&tmO, sO, h * DH, 1 + g * ROWS, b);
        bulk_commit();
      }
    };
    // software pipeline: the probabilities of item k+1 are produced while the tensor pipe runs P V of item k
    if (static_cast<int>(blockIdx.x) < p.n_items) softmax_item(blockIdx.x, 0);
    for (int item = blockIdx.x, k = 0; item < p.n_items; item += gridDim.x, ++k) {
      const int next = item + static_cast<int>(gridDim.x);
      if (next < p.n_items) softmax_item(next, k + 1);
      store_item(item, k);
    }
    if (etid == 0) bulk_wait<0>();   // every store has completed before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
}  // namespace fwd

// ------------------------------------------------------------------------------------------------ backward
namespace bwd {
constexpr int NST = 2;                                  // operand stages (Q, K, V, dO tiles = 64 KB each)
constexpr int STAGE_BYTES = 4 * TILE;
constexpr int P_BYTES = 2 * TILE;
constexpr int OFF_P = NST * STAGE_BYTES;                // P  (block-diagonal [128 q][128 keys])
constexpr int OFF_DS = OFF_P + P_BYTES;                 // dS (same layout)
constexpr int OFF_BAR = OFF_DS + P_BYTES;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
constexpr int THREADS = 192;
// TMEM columns (single-buffered): S, dP (128 each), dQ, dK, dV (64 each)
constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DQ = 256, COL_DK = 320, COL_DV = 384;

template <bool BF>
__global__ void __launch_bounds__(THREADS, 1)
tattn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const __grid_constant__ CUtensorMap tmDQKV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = sm + OFF_P;
  uint8_t* sDS = sm + OFF_DS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* full = bars;              // [NST] TMA -> MMA
  uint64_t* empty = bars + NST;       // [NST] epilogue (gradient tiles stored from the stage buffers) -> TMA
  uint64_t* s_full = bars + 2 * NST;  // MMA -> softmax: S and dP in TMEM
  uint64_t* s_free = s_full + 1;      // softmax -> MMA: S / dP read (4 warp arrivals)
  uint64_t* p_ready = s_free + 1;     // softmax -> MMA: P and dS tiles staged (4 warp arrivals)
  uint64_t* g_full = p_ready + 1;     // MMA -> epilogue: dQ, dK, dV in TMEM
  uint64_t* g_free = g_full + 1;      // epilogue -> MMA: gradient accumulators read (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_free + 1);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int fmt = BF ? 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      tma_prefetch_desc(&tmDQKV);
      for (int i = 0; i < NST; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(s_free, 4);
      mbar_init(p_ready, 4);
      mbar_init(g_full, 1);
      mbar_init(g_free, 4);
      fence_barrier_init();
    }
    __syncwarp();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = tid; i < 2 * P_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_grid_sync();   // prologue done; global memory is touched below only (common.h)

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    for (int item = blockIdx.x, k = 0; item < p.n_items; item += gridDim.x, ++k) {
      const int s = k % NST;
      int b, g, h;
      decode_item(p, item, b, g, h);
      mbar_wait(&empty[s], ((k / NST) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* st = sm + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        const int row = 1 + g * ROWS;
        tma_load_3d(st, &tmQKV, &full[s], h * DH, row, b);
        tma_load_3d(st + TILE, &tmQKV, &full[s], p.d + h * DH, row, b);
        tma_load_3d(st + 2 * TILE, &tmQKV, &full[s], 2 * p.d + h * DH, row, b);
        tma_load_3d(st + 3 * TILE, &tmDO, &full[s], h * DH, row, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc_s = make_idesc_f16(fmt, fmt, 0, 0, ROWS, ROWS);   // S = Q K^T, dP = dO V^T
    const uint32_t idesc_q = make_idesc_f16(fmt, fmt, 0, 1, ROWS, DH);     // dQ = dS K   (A K-major, B MN-major)
    const uint32_t idesc_t = make_idesc_f16(fmt, fmt, 1, 1, ROWS, DH);     // dV = P^T dO, dK = dS^T Q (A read MN-major)
    const uint32_t sm_a = smem_u32(sm), pa = smem_u32(sP), da = smem_u32(sDS);
    auto issue_sdp = [&](int k) {
      const int s = k % NST;
      mbar_wait(&full[s], (k / NST) & 1);
      if (k >= 1) mbar_wait(s_free, (k - 1) & 1);   // S / dP of item k-1 are in registers
      tc_fence_after();
      if (elect_one()) {
        const uint32_t qa = sm_a + s * STAGE_BYTES, ka = qa + TILE, va = qa + 2 * TILE, ga = qa + 3 * TILE;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem + COL_S, make_smem_desc_sw128(qa + ks * 32, 16, 1024),
                   make_smem_desc_sw128(ka + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem + COL_DP, make_smem_desc_sw128(ga + ks * 32, 16, 1024),
                   make_smem_desc_sw128(va + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    if (static_cast<int>(blockIdx.x) < p.n_items) issue_sdp(0);
    for (int item = blockIdx.x, k = 0; item < p.n_items; item += gridDim.x, ++k) {
      const int s = k % NST;
      mbar_wait(p_ready, k & 1);
      if (k >= 1) mbar_wait(g_free, (k - 1) & 1);   // gradient accumulators of item k-1 read out
      tc_fence_after();
      if (elect_one()) {
        const uint32_t qa = sm_a + s * STAGE_BYTES, ka = qa + TILE, ga = qa + 3 * TILE;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dQ = dS K        (K = 128 keys)
          umma_f16(tmem + COL_DQ, make_smem_desc_sw128(da + (ks >> 2) * TILE + (ks & 3) * 32, 16, 1024),
                   make_smem_desc_sw128(ka + ks * 2048, 8192, 1024), idesc_q, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dV = P^T dO      (K = 128 queries; A = the P tile read MN-major)
          umma_f16(tmem + COL_DV, make_smem_desc_sw128(pa + ks * 2048, TILE, 1024),
                   make_smem_desc_sw128(ga + ks * 2048, 8192, 1024), idesc_t, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dK = dS^T Q
          umma_f16(tmem + COL_DK, make_smem_desc_sw128(da + ks * 2048, TILE, 1024),
                   make_smem_desc_sw128(qa + ks * 2048, 8192, 1024), idesc_t, ks > 0 ? 1u : 0u);
        umma_commit(g_full);
      }
      __syncwarp();
      if (item + static_cast<int>(gridDim.x) < p.n_items) issue_sdp(k + 1);
    }
  } else {
    // ------------------------------------------------------------ softmax / dS + epilogue (warps 2..5)
    const int lq = warp & 3;
    const int r = lq * 32 + lane;
    const int etid = tid - 64;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(lq * 32) << 16);
    const float sl2 = p.scale * LOG2E;
    const uint32_t p_off = (r >> 6) * TILE + r * 128 + ((((r >> 3) & 7) ^ (r & 7)) << 4);
    for (int item = blockIdx.x, k = 0; item < p.n_items; item += gridDim.x, ++k) {
      const int s = k % NST;
      int b, g, h;
      decode_item(p, item, b, g, h);
      if (g == 0 && etid < 24) {   // cls row of the clip receives no gradient from the temporal branch
        uint16_t* z = p.out + b * p.clip_rows * p.ld_out + (etid >> 3) * p.d + h * DH + (etid & 7) * 8;
        *reinterpret_cast<uint4*>(z) = make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(s_full, k & 1);
      tc_fence_after();
      uint32_t sr[32], dr[32];
      tmem_ld_32x32(tlane + COL_S + lq * 32, sr);
      tmem_ld_32x32(tlane + COL_DP + lq * 32, dr);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      float v[8], dp[8];
      pick8(sr, lane >> 3, v);
      pick8(dr, lane >> 3, dp);
      float mx = v[0];
#pragma unroll
      for (int e = 1; e < 8; ++e) mx = fmaxf(mx, v[e]);
      float sum = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] = ex2((v[e] - mx) * sl2);
        sum += v[e];
      }
      const float inv = 1.f / sum;
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] *= inv;
        dot = fmaf(v[e], dp[e], dot);
      }
      float ds[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) ds[e] = v[e] * (dp[e] - dot) * p.scale;
      uint4 w, x;
      w.x = pack2<BF>(v[0], v[1]); w.y = pack2<BF>(v[2], v[3]); w.z = pack2<BF>(v[4], v[5]); w.w = pack2<BF>(v[6], v[7]);
      x.x = pack2<BF>(ds[0], ds[1]); x.y = pack2<BF>(ds[2], ds[3]); x.z = pack2<BF>(ds[4], ds[5]); x.w = pack2<BF>(ds[6], ds[7]);
      // P / dS were last read by the gradient MMAs of item k-1, whose completion (g_full) this thread has observed
      *reinterpret_cast<uint4*>(sP + p_off) = w;
      *reinterpret_cast<uint4*>(sDS + p_off) = x;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // ---- epilogue: dQ, dK, dV rows -> 16-bit -> the (consumed) Q, K, V buffers of this stage -> TMA stores
      mbar_wait(g_full, k & 1);
      tc_fence_after();
      uint8_t* st = sm + s * STAGE_BYTES;
#pragma unroll
      for (int t3 = 0; t3 < 3; ++t3) {
        uint32_t o0[32], o1[32];
        const uint32_t col = t3 == 0 ? COL_DQ : (t3 == 1 ? COL_DK : COL_DV);
        tmem_ld_32x32(tlane + col, o0);
        tmem_ld_32x32(tlane + col + 32, o1);
        tmem_ld_wait();
        uint8_t* orow = st + t3 * TILE + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 a4, b4;
          a4.x = pack2<BF>(__uint_as_float(o0[q4 * 8 + 0]), __uint_as_float(o0[q4 * 8 + 1]));
          a4.y = pack2<BF>(__uint_as_float(o0[q4 * 8 + 2]), __uint_as_float(o0[q4 * 8 + 3]));
          a4.z = pack2<BF>(__uint_as_float(o0[q4 * 8 + 4]), __uint_as_float(o0[q4 * 8 + 5]));
          a4.w = pack2<BF>(__uint_as_float(o0[q4 * 8 + 6]), __uint_as_float(o0[q4 * 8 + 7]));
          b4.x = pack2<BF>(__uint_as_float(o1[q4 * 8 + 0]), __uint_as_float(o1[q4 * 8 + 1]));
          b4.y = pack2<BF>(__uint_as_float(o1[q4 * 8 + 2]), __uint_as_float(o1[q4 * 8 + 3]));
          b4.z = pack2<BF>(__uint_as_float(o1[q4 * 8 + 4]), __uint_as_float(o1[q4 * 8 + 5]));
          b4.w = pack2<BF>(__uint_as_float(o1[q4 * 8 + 6]), __uint_as_float(o1[q4 * 8 + 7]));
          *reinterpret_cast<uint4*>(orow + ((q4 ^ (r & 7)) << 4)) = a4;
          *reinterpret_cast<uint4*>(orow + (((4 + q4) ^ (r & 7)) << 4)) = b4;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_free);
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (etid == 0) {
        const int row = 1 + g * ROWS;
        tma_store_3d(&tmDQKV, st, h * DH, row, b);
        tma_store_3d(&tmDQKV, st + TILE, p.d + h * DH, row, b);
        tma_store_3d(&tmDQKV, st + 2 * TILE, 2 * p.d + h * DH, row, b);
        bulk_commit();
        bulk_wait_read<0>();        // the stage buffers have been read: hand the stage back to the producer
        mbar_arrive(&empty[s]);
      }
    }
    if (etid == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
}  // namespace bwd

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(f);
    return static_cast<EncodeTiledFn>(nullptr);
  }();
  return fn;
}

// 3-D map over a [clips][clip_rows][cols] view of a row-major [clips * clip_rows, ld] 16-bit matrix; box = 64 columns x
// 128 rows of one clip, 128-byte swizzle. Rows past the clip's end are out of bounds: zero-filled on load, skipped on
// store.
static int make_map3(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t clip_rows, uint64_t clips, uint64_t ld) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable");
    return ALPRO_EDRIVER;
  }
  cuuint64_t dims[3] = {cols, clip_rows, clips};
  cuuint64_t strides[2] = {ld * 2, clip_rows * ld * 2};
  cuuint32_t box[3] = {DH, ROWS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(3d) failed (%d): cols=%llu rows=%llu clips=%llu ld=%llu ptr=%p", (int)r,
                   (unsigned long long)cols, (unsigned long long)clip_rows, (unsigned long long)clips,
                   (unsigned long long)ld, ptr);
    return ALPRO_EINVAL;
  }
  return 0;
}

static bool tma_ok(const void* ptr, int64_t ld) { return aligned16(ptr) && (ld % 8) == 0; }

// Returns 0 on success, ALPRO_ENOTSUP when the shape / alignment does not fit this path (caller falls back).
int forward(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int B, int N, int T, int heads, int fmt,
            float scale, cudaStream_t st) {
  if (T != 8 || !tma_ok(qkv, ld_qkv) || !tma_ok(out, ld_out)) return ALPRO_ENOTSUP;
  Params p{};
  p.B = B; p.heads = heads; p.d = heads * DH; p.clip_rows = 1 + static_cast<long long>(N) * T;
  p.G = static_cast<int>(cdiv(static_cast<int64_t>(N) * T, ROWS));
  p.n_items = B * p.G * heads;
  p.out = static_cast<uint16_t*>(out); p.ld_out = ld_out; p.scale = scale;
  CUtensorMap tq, to;
  int rc = make_map3(&tq, qkv, 3ull * p.d, p.clip_rows, B, ld_qkv);
  if (rc) return rc;
  rc = make_map3(&to, out, p.d, p.clip_rows, B, ld_out);
  if (rc) return rc;
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  if (fmt == 1) {
    cudaFuncSetAttribute(fwd::tattn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::SMEM_BYTES);
    launch_k(fwd::tattn_fwd_tc_kernel<true>, grid, fwd::THREADS, fwd::SMEM_BYTES, st, tq, to, p);
  } else {
    cudaFuncSetAttribute(fwd::tattn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::SMEM_BYTES);
    launch_k(fwd::tattn_fwd_tc_kernel<false>, grid, fwd::THREADS, fwd::SMEM_BYTES, st, tq, to, p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("tattn_fwd_tc: launch failed: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

int backward(const void* qkv, int64_t ld_qkv, const void* dout, int64_t ld_dout, void* dqkv, int64_t ld_dqkv, int B,
             int N, int T, int heads, int fmt, float scale, cudaStream_t st) {
  if (T != 8 || !tma_ok(qkv, ld_qkv) || !tma_ok(dout, ld_dout) || !tma_ok(dqkv, ld_dqkv)) return ALPRO_ENOTSUP;
  Params p{};
  p.B = B; p.heads = heads; p.d = heads * DH; p.clip_rows = 1 + static_cast<long long>(N) * T;
  p.G = static_cast<int>(cdiv(static_cast<int64_t>(N) * T, ROWS));
  p.n_items = B * p.G * heads;
  p.out = static_cast<uint16_t*>(dqkv); p.ld_out = ld_dqkv; p.scale = scale;
  CUtensorMap tq, tg, td;
  int rc = make_map3(&tq, qkv, 3ull * p.d, p.clip_rows, B, ld_qkv);
  if (rc) return rc;
  rc = make_map3(&tg, dout, p.d, p.clip_rows, B, ld_dout);
  if (rc) return rc;
  rc = make_map3(&td, dqkv, 3ull * p.d, p.clip_rows, B, ld_dqkv);
  if (rc) return rc;
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  if (fmt == 1) {
    cudaFuncSetAttribute(bwd::tattn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::SMEM_BYTES);
    launch_k(bwd::tattn_bwd_tc_kernel<true>, grid, bwd::THREADS, bwd::SMEM_BYTES, st, tq, tg, td, p);
  } else {
    cudaFuncSetAttribute(bwd::tattn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::SMEM_BYTES);
    launch_k(bwd::tattn_bwd_tc_kernel<false>, grid, bwd::THREADS, bwd::SMEM_BYTES, st, tq, tg, td, p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("tattn_bwd_tc: launch failed: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace tattn
}  // namespace alpro
