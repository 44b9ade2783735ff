// gemm16_2cta_kernel: the CTA-pair (tcgen05 cta_group::2) variant of the persistent GEMM in gemm_tc.cu.
//
// Two CTAs of one cluster (= the two SMs of a TPC) compute one 256x256 output tile: each CTA stages its own 128 rows of
// A and its own 128-row half of the B tile (32 KB per k-block instead of 48 KB -> a 6-stage ring fits), the leader CTA's
// single MMA thread issues tcgen05.mma.cta_group::2 (M=256, N=256, K=16) that reads both CTAs' shared memory, and each
// CTA keeps its 128x256 fp32 accumulator half in its own TMEM and runs its own fused epilogue.
//   * TMA loads of both CTAs complete on the LEADER's full barrier (peer bit of the barrier address cleared);
//   * tcgen05.commit ... multicast::cluster releases the smem slot / publishes the accumulator in both CTAs;
//   * epilogue warps of both CTAs arrive on the leader's TMEM-empty barrier (mapa + mbarrier.arrive.shared::cluster).
#include <mutex>

#include "common.h"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace alpro {
namespace gemm2 {

constexpr int BM = 128;   // rows per CTA (256 per pair)
constexpr int BN = 256;   // columns per pair
constexpr int BNH = 128;  // B rows staged per CTA
constexpr int BK = 64;
constexpr int STAGES = 6;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int B_STAGE_BYTES = BNH * BK * 2;  // 16 KiB
constexpr int MN_BOX_BYTES = 64 * BK * 2;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int EPI_CHUNK = 32;
constexpr int EPI_WARP_BYTES = 32 * EPI_CHUNK * 4;
constexpr int PIPE_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
constexpr int SMEM_BYTES = 1024 + PIPE_BYTES + NUM_EPI_WARPS * EPI_WARP_BYTES + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

using namespace gemm;

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment in the shared window.
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform (uniform role branches)
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();   // 0 = leader (issues the MMAs), 1 = peer

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * NUM_EPI_WARPS);  // epilogue warps of both CTAs arrive on the leader's barrier
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation visible in both CTAs before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();   // the prologue above overlapped the tail of the previous kernel (common.h)

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;   // tiles of (2*BM) x BN, one per CTA pair
  const int num_work = num_tiles * p.split_k;
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
            // the leader's full barrier collects the bytes of BOTH CTAs' loads
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B_STAGE_BYTES));
            uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
            uint8_t* b_dst = sB + stage * B_STAGE_BYTES;
            const int m0 = m_blk * (2 * BM) + static_cast<int>(cta_rank) * BM;   // this CTA's 128 rows of the 256-row tile
            const int n0 = n_blk * BN + static_cast<int>(cta_rank) * BNH;        // this CTA's half of the B tile
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
      // K-major SW128: 8-row groups 1024 B apart (SBO); +32 B per UMMA_K inside the swizzle row.
      // MN-major SW128: 64-wide MN blocks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO); +2048 B per UMMA_K.
      const uint32_t a_lbo = p.a_mn ? MN_BOX_BYTES : 16, a_kstep = p.a_mn ? UMMA_K * 128 : UMMA_K * 2;
      const uint32_t b_lbo = p.b_mn ? MN_BOX_BYTES : 16, b_kstep = p.b_mn ? UMMA_K * 128 : UMMA_K * 2;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const int tile = w / p.split_k;
        const int split = w - tile * p.split_k;
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
            umma_commit_2cta(&empty_bar[stage]);  // frees the slot in both CTAs once these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_2cta(&tfull_bar[acc]);  // accumulator halves complete -> both CTAs' epilogues
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    // TMEM lane quarter q = warp % 4 (hardware restriction); the two warps of a quarter split the 256 columns.
    // Row-owner side: tcgen05.ld gives lane i the 32-column chunk of row q*32+i -> XOR-swizzled smem.
    // Coalesced side: 8 lanes cover the 32 columns of one row (float4 each), 4 rows per warp instruction.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    float* stg = reinterpret_cast<float*>(sEpi + (warp - 4) * EPI_WARP_BYTES);
    int acc = 0;
    uint32_t acc_phase = 0;
    const int crow = lane >> 3;  // 0..3
    const int cch = lane & 7;    // float4 index within the 32-column chunk
    const float alpha = p.alpha;
    // Pull this warp's slice of the residual / saved-derivative tile of work item `pw` towards L2 one tile ahead of
    // its use, so that the epilogue's global loads are L2 hits instead of exposed DRAM latency.
    auto prefetch_tile = [&](int pw) {
      if (pw >= num_work) return;
      const int ptile = pw / p.split_k;
      const int pm = ptile / p.num_n_tiles;
      const int pn = ptile - pm * p.num_n_tiles;
      const long long prow = static_cast<long long>(pm) * (2 * BM) + cta_rank * BM + q * 32 + lane;
      const int pcol = pn * BN + half * (BN / 2);
      if (prow >= p.M) return;
      if (MODE == E_GELU_GRAD || (MODE == E_GENERIC && p.aux16)) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (pcol + j * 64 < p.N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.aux16 + prow * p.ldaux + pcol + j * 64));
      }
      if (MODE == E_RESID_OUT32 || (MODE == E_GENERIC && p.resid)) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (pcol + j * 32 < p.N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + prow * p.ldresid + pcol + j * 32));
      }
    };
    prefetch_tile(cluster_id);
    for (int w = cluster_id; w < num_work; w += num_clusters) {
      const int tile = w / p.split_k;
      const int m_blk = tile / p.num_n_tiles;
      const int n_blk = tile - m_blk * p.num_n_tiles;
      prefetch_tile(w + num_clusters);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const long long row0 = static_cast<long long>(m_blk) * (2 * BM) + cta_rank * BM + q * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      const int c_begin = half * (BN / 2), c_end = (half + 1) * (BN / 2);
      uint32_t r[32];
      bool have = n_blk * BN + c_begin < p.N;  // warp-uniform
      if (have) tmem_ld_32x32(taddr + c_begin, r);
#pragma unroll 1
      for (int c = c_begin; c < c_end && have; c += EPI_CHUNK) {
        const int col0 = n_blk * BN + c;
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 t = make_float4(__uint_as_float(r[4 * j]) * alpha, __uint_as_float(r[4 * j + 1]) * alpha,
                                 __uint_as_float(r[4 * j + 2]) * alpha, __uint_as_float(r[4 * j + 3]) * alpha);
          *reinterpret_cast<float4*>(stg + lane * EPI_CHUNK + ((j ^ (lane & 7)) << 2)) = t;
        }
        __syncwarp();
        // prefetch the next chunk's accumulators while this one is written out
        have = (c + EPI_CHUNK < c_end) && (col0 + EPI_CHUNK < p.N);
        if (have) tmem_ld_32x32(taddr + c + EPI_CHUNK, r);
        const int col = col0 + cch * 4;
        if (col < p.N) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (MODE != E_ATOMIC && MODE != E_GELU_GRAD && p.bias) {
            if (p.vec_ok && col + 4 <= p.N) {
              b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            } else {
              b4.x = __ldg(p.bias + col);
              if (col + 1 < p.N) b4.y = __ldg(p.bias + col + 1);
              if (col + 2 < p.N) b4.z = __ldg(p.bias + col + 2);
              if (col + 3 < p.N) b4.w = __ldg(p.bias + col + 3);
            }
          }
          {
            float v[8][4];
            long long rows[8];
            bool ok[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = crow + 4 * i;
              rows[i] = row0 + rl;
              ok[i] = rows[i] < p.M;
              const float4 t = *reinterpret_cast<const float4*>(stg + rl * EPI_CHUNK + ((cch ^ (rl & 7)) << 2));
              v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
            }
            epilogue_rows<MODE, 8>(p, v, b4, rows, ok, col);   // all 8 rows at once: 8 global loads in flight per lane
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);   // leader CTA's barrier
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA may exit (or free TMEM) while its peer can still signal / read it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}


template <int MODE>
static int launch_mode(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p, int grid, cudaStream_t st) {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(gemm16_2cta_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  launch_k(gemm16_2cta_kernel<MODE>, grid, NUM_THREADS, SMEM_BYTES, st, tmA, tmB, p);
  return 0;
}

// Called by alpro_gemm16 (gemm_tc.cu). `p` holds tile counts for 256x256 pair tiles; grid = 2 * clusters.
int launch_2cta(int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p, int grid,
                cudaStream_t st) {
  switch (mode) {
    case E_OUT16: return launch_mode<E_OUT16>(tmA, tmB, p, grid, st);
    case E_GELU_SAVE: return launch_mode<E_GELU_SAVE>(tmA, tmB, p, grid, st);
    case E_GELU_GRAD: return launch_mode<E_GELU_GRAD>(tmA, tmB, p, grid, st);
    case E_RESID_OUT32: return launch_mode<E_RESID_OUT32>(tmA, tmB, p, grid, st);
    case E_OUT32: return launch_mode<E_OUT32>(tmA, tmB, p, grid, st);
    case E_ATOMIC: return launch_mode<E_ATOMIC>(tmA, tmB, p, grid, st);
    default: return launch_mode<E_GENERIC>(tmA, tmB, p, grid, st);
  }
}

}  // namespace gemm2
}  // namespace alpro
