// gemm16_2cta_tma_epi_kernel: the CTA-pair tcgen05 GEMM of gemm_tc2.cu with a TMA-store epilogue for the 16-bit output
// modes (E_OUT16, E_GELU_SAVE, E_GELU_GRAD) and the fp32 residual mode (E_RESID_OUT32) -- the modes whose K = 768
// problems were bound by epilogue instruction issue, not by the tensor pipe (profiles/r01f: fc1+GELU 648 TFLOP/s,
// fc2-dgrad*gelu' 511 TFLOP/s, the 768x768 projections+residual ~660 TFLOP/s).
//
// Differences from gemm_tc2.cu:
//   * 16 epilogue warps (4 per TMEM lane quarter, 64 accumulator columns each) instead of 8: four resident epilogue
//     warps per scheduler hide the TMEM-load / MUFU / shared-memory latencies of each other;
//   * the math runs in the ROW-OWNER layout tcgen05.ld delivers (lane i = row i, 32 consecutive columns in registers):
//     no fp32 transpose through shared memory, no per-element address arithmetic or store predication;
//   * results are packed to 16 bit, written once to a 64B-swizzled 32x32 staging tile (conflict-free 16-byte stores)
//     and shipped by ONE cp.async.bulk.tensor store per tile chunk; the TMA unit clips the ragged M edge;
//   * E_GELU_GRAD reads the saved gelu'(pre) tile with a TMA load one chunk ahead (double-buffered, in-place multiply
//     in packed 16-bit arithmetic); E_RESID_OUT32 does the same with 32x16 fp32 residual tiles (x + branch in place);
//   * per-row scalars (stochastic-depth row scales, the "cls rows keep the residual" rule) cost one load per ROW here
//     because a lane owns a row;
//   * the TMEM accumulator is released to the MMA warp as soon as its last chunk sits in registers.
// Pipeline: 5 smem stages of 32 KB per CTA (the sixth stage of gemm_tc2.cu pays for the second set of epilogue warps).
#include <mutex>

#include "common.h"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace alpro {
namespace gemm3 {

constexpr int BM = 128;   // rows per CTA (256 per pair)
constexpr int BN = 256;   // columns per pair
constexpr int BNH = 128;  // B rows staged per CTA
constexpr int BK = 64;
constexpr int STAGES = 5;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int B_STAGE_BYTES = BNH * BK * 2;  // 16 KiB
constexpr int MN_BOX_BYTES = 64 * BK * 2;
constexpr int NUM_EPI_WARPS = 16;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner
constexpr int TMEM_COLS = 512;
constexpr int EPI_CHUNK = 32;                          // columns per tcgen05.ld / per TMA store box
constexpr int EPI_BUF_BYTES = 32 * EPI_CHUNK * 2;      // one 32x32 16-bit staging tile
constexpr int EPI_WARP_BYTES = 2 * EPI_BUF_BYTES;
constexpr int PIPE_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
constexpr int NUM_BARS = 2 * STAGES + 4 + 2 * NUM_EPI_WARPS;
constexpr int SMEM_BYTES = 1024 + PIPE_BYTES + NUM_EPI_WARPS * EPI_WARP_BYTES + NUM_BARS * 8 + 64;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

using namespace gemm;

template <bool BF>
__device__ __forceinline__ uint32_t cvt_pack(float lo, float hi) {
  uint32_t d;
  if (BF) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool BF>
__device__ __forceinline__ uint32_t mul_pack(uint32_t a, uint32_t b) {
  uint32_t d;
  if (BF) asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// Byte offset of 16-byte chunk `c` (0..3) of row `row` in a 64B-swizzled 32-row x 64-byte staging tile (the layout a
// CU_TENSOR_MAP_SWIZZLE_64B box expects: chunk index XOR address bits [7,9)).
__device__ __forceinline__ uint32_t stage_off(int row, int c) { return row * 64 + ((c ^ ((row >> 1) & 3)) << 4); }

// Per-row epilogue scalars of this lane's row: v = ra * acc + rb * bias (ra already includes alpha).
struct RowScale {
  float ra, rb;
  float rc;   // factor of the second (never row-scaled) bias: 1, or 0 on rows that keep the plain residual
};

// One 32-column chunk of one accumulator row (this lane's): r[] raw fp32 accumulators -> packed 16-bit results in pk[]
// (E_OUT16: pk[0..16); E_GELU_SAVE: gelu in pk[0..16), gelu' in pk[16..32)) or, for E_GELU_GRAD, the in-place product
// with the saved-derivative tile in buf0. RS: the row carries stochastic-depth scales (otherwise ra = alpha, rb = 1).
template <int MODE, bool BF, bool RS>
__device__ __forceinline__ void chunk_math(const GemmKParams& p, const uint32_t (&r)[32], int col0, int lane,
                                           uint8_t* buf0, const RowScale rs, uint32_t (&pk)[32]) {
  if (MODE == E_GELU_GRAD) {
    // out = (alpha * acc) * gelu'(pre): the saved derivative tile was TMA-loaded into buf0; multiply in place
    const float alpha = rs.ra;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 a = *reinterpret_cast<const uint4*>(buf0 + stage_off(lane, g));
      a.x = mul_pack<BF>(cvt_pack<BF>(__uint_as_float(r[8 * g + 0]) * alpha, __uint_as_float(r[8 * g + 1]) * alpha), a.x);
      a.y = mul_pack<BF>(cvt_pack<BF>(__uint_as_float(r[8 * g + 2]) * alpha, __uint_as_float(r[8 * g + 3]) * alpha), a.y);
      a.z = mul_pack<BF>(cvt_pack<BF>(__uint_as_float(r[8 * g + 4]) * alpha, __uint_as_float(r[8 * g + 5]) * alpha), a.z);
      a.w = mul_pack<BF>(cvt_pack<BF>(__uint_as_float(r[8 * g + 6]) * alpha, __uint_as_float(r[8 * g + 7]) * alpha), a.w);
      *reinterpret_cast<uint4*>(buf0 + stage_off(lane, g)) = a;
    }
  } else {
    // bias of the 32 columns: the same addresses in every lane (one broadcast transaction per float4)
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * g));
      if (RS) { b.x *= rs.rb; b.y *= rs.rb; b.z *= rs.rb; b.w *= rs.rb; }
      float x0 = fmaf(__uint_as_float(r[4 * g + 0]), rs.ra, b.x), x1 = fmaf(__uint_as_float(r[4 * g + 1]), rs.ra, b.y);
      float x2 = fmaf(__uint_as_float(r[4 * g + 2]), rs.ra, b.z), x3 = fmaf(__uint_as_float(r[4 * g + 3]), rs.ra, b.w);
      if (MODE == E_GELU_SAVE) {   // (without out16b the derivative half is dead code the compiler drops per branch)
        float d0, d1, d2, d3;
        gelu_erf_both(x0, x0, d0); gelu_erf_both(x1, x1, d1); gelu_erf_both(x2, x2, d2); gelu_erf_both(x3, x3, d3);
        pk[16 + 2 * g] = cvt_pack<BF>(d0, d1);
        pk[16 + 2 * g + 1] = cvt_pack<BF>(d2, d3);
      }
      pk[2 * g] = cvt_pack<BF>(x0, x1);
      pk[2 * g + 1] = cvt_pack<BF>(x2, x3);
    }
  }
}

// E_RESID_OUT32: one 16-column fp32 chunk; the residual tile sits in buf (TMA-loaded), result written in place:
//   out = resid + ra * acc + rb * bias      (rows that keep the plain residual arrive with ra = rb = 0)
//   with a 16-bit multiplier tile (p.aux16: the hidden-dropout mask of BertSelfOutput / BertOutput, xbert.py:358,436):
//   out = resid + (ra * acc + rb * bias) * aux      — the lane reads the 32 bytes of its own row straight from global
__device__ __forceinline__ void chunk_resid(const GemmKParams& p, const uint32_t (&r)[16], int col0, int lane,
                                            uint8_t* buf, const RowScale rs, long long row) {
  if (p.aux16) {   // warp-uniform
    const uint4* ap = reinterpret_cast<const uint4*>(p.aux16 + row * p.ldaux + col0);
    const uint4 a0 = __ldg(ap), a1 = __ldg(ap + 1);
    const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * g));
      float4 x = *reinterpret_cast<const float4*>(buf + stage_off(lane, g));
      const float m0 = f16_to_32(static_cast<uint16_t>(aw[2 * g] & 0xffff), p.aux_fmt);
      const float m1 = f16_to_32(static_cast<uint16_t>(aw[2 * g] >> 16), p.aux_fmt);
      const float m2 = f16_to_32(static_cast<uint16_t>(aw[2 * g + 1] & 0xffff), p.aux_fmt);
      const float m3 = f16_to_32(static_cast<uint16_t>(aw[2 * g + 1] >> 16), p.aux_fmt);
      x.x = fmaf(fmaf(__uint_as_float(r[4 * g + 0]), rs.ra, b.x * rs.rb), m0, x.x);
      x.y = fmaf(fmaf(__uint_as_float(r[4 * g + 1]), rs.ra, b.y * rs.rb), m1, x.y);
      x.z = fmaf(fmaf(__uint_as_float(r[4 * g + 2]), rs.ra, b.z * rs.rb), m2, x.z);
      x.w = fmaf(fmaf(__uint_as_float(r[4 * g + 3]), rs.ra, b.w * rs.rb), m3, x.w);
      *reinterpret_cast<float4*>(buf + stage_off(lane, g)) = x;
    }
    return;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * g));
    float4 x = *reinterpret_cast<const float4*>(buf + stage_off(lane, g));
    x.x = fmaf(__uint_as_float(r[4 * g + 0]), rs.ra, fmaf(b.x, rs.rb, x.x));
    x.y = fmaf(__uint_as_float(r[4 * g + 1]), rs.ra, fmaf(b.y, rs.rb, x.y));
    x.z = fmaf(__uint_as_float(r[4 * g + 2]), rs.ra, fmaf(b.z, rs.rb, x.z));
    x.w = fmaf(__uint_as_float(r[4 * g + 3]), rs.ra, fmaf(b.w, rs.rb, x.w));
    if (p.bias2) {   // warp-uniform
      const float4 c = __ldg(reinterpret_cast<const float4*>(p.bias2 + col0 + 4 * g));
      x.x = fmaf(c.x, rs.rc, x.x); x.y = fmaf(c.y, rs.rc, x.y); x.z = fmaf(c.z, rs.rc, x.z); x.w = fmaf(c.w, rs.rc, x.w);
    }
    *reinterpret_cast<float4*>(buf + stage_off(lane, g)) = x;
  }
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm16_2cta_tma_epi_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
                           const __grid_constant__ CUtensorMap tmAux, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sEpi = smem + PIPE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + NUM_EPI_WARPS * EPI_WARP_BYTES);
  uint64_t* full_bar = bars;                // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;      // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;  // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;     // [2] epilogue -> MMA
  uint64_t* aux_bar = tempty_bar + 2;       // [NUM_EPI_WARPS][2] TMA load of the saved-derivative tile -> epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 2 * NUM_EPI_WARPS);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform (uniform role branches)
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();   // 0 = leader (issues the MMAs), 1 = peer

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (MODE == E_GELU_SAVE) tma_prefetch_desc(&tmO2);
    if (MODE == E_GELU_GRAD || MODE == E_RESID_OUT32) tma_prefetch_desc(&tmAux);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * NUM_EPI_WARPS);  // epilogue warps of both CTAs arrive on the leader's barrier
    }
    for (int i = 0; i < 2 * NUM_EPI_WARPS; ++i) mbar_init(&aux_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation visible in both CTAs before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();   // everything above overlapped the tail of the previous kernel; operands are touched below only

  // work item = (tile of (2*BM) x BN, K split); split_k > 1 only in E_ATOMIC (out32 += partial products)
  const int num_work = p.num_m_tiles * p.num_n_tiles * p.split_k;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // the whole warp walks the loop (uniform control flow); one elected lane issues the TMA instructions
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const int tile = w / p.split_k;
        const int split = w - tile * p.split_k;
        const int m_blk = tile / p.num_n_tiles;
        const int n_blk = tile - m_blk * p.num_n_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);   // own smem slot free (multicast commit of the leader's MMAs)
          if (elect_one()) {
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B_STAGE_BYTES));
            uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
            uint8_t* b_dst = sB + stage * B_STAGE_BYTES;
            const int m0 = m_blk * (2 * BM) + static_cast<int>(cta_rank) * BM;
            const int n0 = n_blk * BN + static_cast<int>(cta_rank) * BNH;
            if (!p.a_mn) {
              tma_load_2d_2cta(a_dst, &tmA, &full_bar[stage], kb * BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d_2cta(a_dst + j * MN_BOX_BYTES, &tmA, &full_bar[stage], m0 + j * 64, kb * BK);
            }
            if (!p.b_mn) {
              tma_load_2d_2cta(b_dst, &tmB, &full_bar[stage], kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BNH / 64; ++j)
                tma_load_2d_2cta(b_dst + j * MN_BOX_BYTES, &tmB, &full_bar[stage], n0 + j * 64, kb * BK);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread of the leader CTA)
    // the leader CTA's whole warp walks the loop; one elected lane issues MMAs and commits (see elect_one, ptx.cuh)
    if (cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t a_lbo = p.a_mn ? MN_BOX_BYTES : 16, a_kstep = p.a_mn ? UMMA_K * 128 : UMMA_K * 2;
      const uint32_t b_lbo = p.b_mn ? MN_BOX_BYTES : 16, b_kstep = p.b_mn ? UMMA_K * 128 : UMMA_K * 2;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const int split = w % p.split_k;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * B_STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
              umma_f16_2cta(tmem_acc, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2cta(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_2cta(&tfull_bar[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (16 warps)
    // TMEM lane quarter q = warp % 4 (hardware restriction); the four warps of a quarter take 64 columns each.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int slot = ew >> 2;
    uint8_t* const ebase = sEpi + ew * EPI_WARP_BYTES;
    auto bufp = [&](uint32_t i) { return ebase + (i & 1u) * EPI_BUF_BYTES; };   // the warp's two staging tiles
    uint64_t* abar = aux_bar + 2 * ew;
    const bool bf = p.out16_fmt != 0;
    constexpr bool LOADS = (MODE == E_GELU_GRAD || MODE == E_RESID_OUT32);   // epilogue input tile via TMA load
    constexpr int CW = (MODE == E_RESID_OUT32 || MODE == E_ATOMIC) ? 16 : 32;   // chunk width: 64-byte rows either way
    constexpr int NCH = 64 / CW;                                             // chunks per warp and tile
    int acc = 0;
    uint32_t acc_phase = 0;
    // chunk (w, ch) of this warp covers columns n_blk*BN + slot*64 + ch*CW .. +CW ; valid while it starts below N
    auto chunk_col = [&](int w, int ch) { return ((w / p.split_k) % p.num_n_tiles) * BN + slot * 64 + ch * CW; };
    auto chunk_row = [&](int w) {
      return ((w / p.split_k) / p.num_n_tiles) * (2 * BM) + static_cast<int>(cta_rank) * BM + q * 32;
    };
    // LOADS: cursor of the NEXT chunk whose input tile has to be requested, one chunk ahead of its use
    int pw = cluster_id, pch = 0;
    uint32_t nload = 0, nuse = 0;   // chunks requested / consumed (buffer = n & 1, barrier parity = (n >> 1) & 1)
    auto request_next = [&]() {
      while (pw < num_work && chunk_col(pw, pch) >= p.N) {
        if (++pch == NCH) { pch = 0; pw += num_clusters; }
      }
      if (pw >= num_work) return;
      if (lane == 0) {
        bulk_wait_read<0>();   // the store that last read this buffer has drained it
        mbar_arrive_expect_tx(&abar[nload & 1], EPI_BUF_BYTES);
        tma_load_2d(bufp(nload), &tmAux, &abar[nload & 1], chunk_col(pw, pch), chunk_row(pw));
      }
      ++nload;
      if (++pch == NCH) { pch = 0; pw += num_clusters; }
    };
    if (LOADS) request_next();

    for (int w = cluster_id; w < num_work; w += num_clusters) {
      const int row0 = chunk_row(w);
      // per-row scalars of this lane's row (clamped row index for the loads; the TMA store clips rows >= M)
      RowScale rs;
      rs.ra = p.alpha;
      rs.rb = 1.f;
      rs.rc = 1.f;
      if (MODE != E_GELU_GRAD && MODE != E_ATOMIC) {
        const int row = min(row0 + lane, p.M - 1);
        if (p.rs_acc) {
          const float ra = __ldg(p.rs_acc + row);
          rs.rb = p.rs_bias ? __ldg(p.rs_bias + row) : ra;
          rs.ra *= ra;
        }
        if (MODE == E_RESID_OUT32 && p.skip_period > 0 && (row % p.skip_period) == 0) rs.ra = rs.rb = rs.rc = 0.f;
      }
      const bool row_scaled = p.rs_acc != nullptr;   // warp-uniform
      const int nvalid = min(NCH, max(0, (p.N - chunk_col(w, 0) + CW - 1) / CW));   // warp-uniform
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN) +
                             static_cast<uint32_t>(slot * 64);
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        const bool valid = ch < nvalid;
        const bool last = ch == max(nvalid, 1) - 1;   // exactly one release per tile, after this warp's last TMEM read
        uint32_t r[CW], pk[32];
        if (valid) {
          if (CW == 32) tmem_ld_32x32(taddr + ch * CW, reinterpret_cast<uint32_t (&)[32]>(r));
          else          tmem_ld_32x16(taddr + ch * CW, reinterpret_cast<uint32_t (&)[16]>(r));
          if (LOADS) request_next();     // keep the next input tile in flight
          tmem_ld_wait();
        }
        if (last) {   // every column this warp needs from the accumulator is in registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
        }
        if (!valid) {
          if (last) break;
          continue;
        }
        const int col0 = chunk_col(w, ch);
        uint8_t* b0;
        if (LOADS) {
          b0 = bufp(nuse);
          mbar_wait(&abar[nuse & 1], (nuse >> 1) & 1);
          if (MODE == E_RESID_OUT32) {
            chunk_resid(p, reinterpret_cast<const uint32_t (&)[16]>(r), col0, lane, b0, rs,
                        min(static_cast<long long>(row0) + lane, static_cast<long long>(p.M) - 1));
          } else {
            if (bf) chunk_math<MODE, true, false>(p, reinterpret_cast<const uint32_t (&)[32]>(r), col0, lane, b0, rs, pk);
            else    chunk_math<MODE, false, false>(p, reinterpret_cast<const uint32_t (&)[32]>(r), col0, lane, b0, rs, pk);
          }
          ++nuse;
        } else if (MODE == E_ATOMIC) {
          // split-K partial product: fp32 tile -> staging -> TMA reduce-add into out32 (the L2 performs the adds in
          // bulk; the per-element red.global of gemm_tc2.cu issued 64 K atomics per tile)
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          b0 = bufp(nuse);
          ++nuse;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(b0 + stage_off(lane, g)) =
                make_float4(__uint_as_float(r[4 * g + 0]) * rs.ra, __uint_as_float(r[4 * g + 1]) * rs.ra,
                            __uint_as_float(r[4 * g + 2]) * rs.ra, __uint_as_float(r[4 * g + 3]) * rs.ra);
        } else {
          if (row_scaled) {
            if (bf) chunk_math<MODE, true, true>(p, reinterpret_cast<const uint32_t (&)[32]>(r), col0, lane, nullptr, rs, pk);
            else    chunk_math<MODE, false, true>(p, reinterpret_cast<const uint32_t (&)[32]>(r), col0, lane, nullptr, rs, pk);
          } else {
            if (bf) chunk_math<MODE, true, false>(p, reinterpret_cast<const uint32_t (&)[32]>(r), col0, lane, nullptr, rs, pk);
            else    chunk_math<MODE, false, false>(p, reinterpret_cast<const uint32_t (&)[32]>(r), col0, lane, nullptr, rs, pk);
          }
          // staging tiles free again? E_OUT16 alternates its two tiles, E_GELU_SAVE fills both per chunk
          if (lane == 0) {
            if (MODE == E_GELU_SAVE) bulk_wait_read<0>();
            else bulk_wait_read<1>();
          }
          __syncwarp();
          b0 = (MODE == E_GELU_SAVE) ? bufp(0) : bufp(nuse);
          ++nuse;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(b0 + stage_off(lane, g)) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          if (MODE == E_GELU_SAVE && p.out16b) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(bufp(1) + stage_off(lane, g)) =
                  make_uint4(pk[16 + 4 * g], pk[16 + 4 * g + 1], pk[16 + 4 * g + 2], pk[16 + 4 * g + 3]);
          }
        }
        fence_proxy_async();   // generic-proxy writes of every lane -> visible to the TMA unit
        __syncwarp();
        if (lane == 0) {
          if (MODE == E_ATOMIC) tma_reduce_add_2d(&tmO, b0, col0, row0);
          else tma_store_2d(&tmO, b0, col0, row0);
          if (MODE == E_GELU_SAVE && p.out16b) tma_store_2d(&tmO2, bufp(1), col0, row0);
          bulk_commit();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();   // all stores of this warp have completed before the CTA may retire
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA may exit (or free TMEM) while its peer can still signal / read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

template <int MODE>
static int launch_mode(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmO2,
                       const CUtensorMap& tmAux, const GemmKParams& p, int grid, cudaStream_t st) {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(gemm16_2cta_tma_epi_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  launch_k(gemm16_2cta_tma_epi_kernel<MODE>, grid, NUM_THREADS, SMEM_BYTES, st, tmA, tmB, tmO, tmO2, tmAux, p);
  return 0;
}

// Called by alpro_gemm16 (gemm_tc.cu) for E_OUT16 / E_GELU_SAVE / E_GELU_GRAD problems that satisfy the TMA
// constraints (see tma_epilogue_ok there). `p` holds tile counts for 256x256 pair tiles; grid = 2 * clusters.
int launch_2cta_tma_epi(int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                        const CUtensorMap& tmO2, const CUtensorMap& tmAux, const GemmKParams& p, int grid,
                        cudaStream_t st) {
  switch (mode) {
    case E_OUT16: return launch_mode<E_OUT16>(tmA, tmB, tmO, tmO2, tmAux, p, grid, st);
    case E_GELU_SAVE: return launch_mode<E_GELU_SAVE>(tmA, tmB, tmO, tmO2, tmAux, p, grid, st);
    case E_GELU_GRAD: return launch_mode<E_GELU_GRAD>(tmA, tmB, tmO, tmO2, tmAux, p, grid, st);
    case E_RESID_OUT32: return launch_mode<E_RESID_OUT32>(tmA, tmB, tmO, tmO2, tmAux, p, grid, st);
    case E_ATOMIC: return launch_mode<E_ATOMIC>(tmA, tmB, tmO, tmO2, tmAux, p, grid, st);
    default: return -1;
  }
}

}  // namespace gemm3
}  // namespace alpro
