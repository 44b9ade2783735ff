// alpro_gemm16: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0 (1 lane) : TMA producer   — cp.async.bulk.tensor 128B-swizzled operand tiles into a 4-stage smem ring
//   warp 1 (1 lane) : MMA issuer     — tcgen05.mma.cta_group::1.kind::f16, 128x256x16 atoms, fp32 accumulators in TMEM
//   warp 2          : TMEM allocator — 512 columns = two 128x256 fp32 accumulator buffers (epilogue/mainloop overlap)
//   warps 4..11     : epilogue       — tcgen05.ld TMEM->regs -> swizzled smem -> coalesced fused bias/GELU/residual stores
//
// Operands are 16-bit (fp16 or bf16, chosen per operand in the instruction descriptor); both K-major and MN-major
// operand storage are supported through the UMMA smem-descriptor / TMA box shapes, so the same kernel runs
//   fwd   y  = x W^T      (A K-major,  B K-major)
//   dgrad dx = dy W       (A K-major,  B MN-major)
//   wgrad dW = dy^T x     (A MN-major, B MN-major, split-K with fp32 red.global.add)
// without any transposed copies in HBM.
//
// Reference call sites replaced: nn.Linear in src/modeling/timesformer/vit.py:60,63,84,98,161 and
// src/modeling/xbert.py:273-292,357,422,435,659,681 (and their autograd-derived backward GEMMs).
#include <stdlib.h>

#include <mutex>

#include "common.h"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace alpro {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 x 16-bit = 128 bytes = one swizzle row
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int MN_BOX_BYTES = 64 * BK * 2;   // one [64 k][64 mn] box of an MN-major operand = 8 KiB
constexpr int NUM_EPI_WARPS = 8;                        // 2 per TMEM lane quarter (each takes half of the columns)
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;   // warps 0..3: TMA / MMA / TMEM-alloc / spare
constexpr int TMEM_COLS = 512;
constexpr int EPI_CHUNK = 32;                           // columns staged per step
constexpr int EPI_WARP_BYTES = 32 * EPI_CHUNK * 4;      // 32 rows x 32 fp32, XOR-swizzled (no padding)
constexpr int PIPE_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
constexpr int SMEM_BYTES = 1024 + PIPE_BYTES + NUM_EPI_WARPS * EPI_WARP_BYTES + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

using namespace gemm;

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
      mbar_init(&tempty_bar[a], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();   // the prologue above overlapped the tail of the previous kernel (common.h)

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_work = num_tiles * p.split_k;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int tile = w / p.split_k;
        const int split = w - tile * p.split_k;
        const int m_blk = tile / p.num_n_tiles;
        const int n_blk = tile - m_blk * p.num_n_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
          uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
          uint8_t* b_dst = sB + stage * B_STAGE_BYTES;
          if (!p.a_mn) {
            tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(a_dst + j * MN_BOX_BYTES, &tmA, &full_bar[stage], m_blk * BM + j * 64, kb * BK);
          }
          if (!p.b_mn) {
            tma_load_2d(b_dst, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(b_dst + j * MN_BOX_BYTES, &tmB, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // K-major SW128: 8-row groups 1024 B apart (SBO); +32 B per UMMA_K inside the swizzle row.
      // MN-major SW128: 64-wide MN blocks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO); +2048 B per UMMA_K.
      const uint32_t a_lbo = p.a_mn ? MN_BOX_BYTES : 16, a_kstep = p.a_mn ? UMMA_K * 128 : UMMA_K * 2;
      const uint32_t b_lbo = p.b_mn ? MN_BOX_BYTES : 16, b_kstep = p.b_mn ? UMMA_K * 128 : UMMA_K * 2;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
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
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
            umma_f16(tmem_acc, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
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
      const long long prow = static_cast<long long>(pm) * BM + q * 32 + lane;
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
    prefetch_tile(static_cast<int>(blockIdx.x));
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int tile = w / p.split_k;
      const int m_blk = tile / p.num_n_tiles;
      const int n_blk = tile - m_blk * p.num_n_tiles;
      prefetch_tile(w + static_cast<int>(gridDim.x));
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const long long row0 = static_cast<long long>(m_blk) * BM + q * 32;
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
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// 2-D tensor map over 16-bit (or fp32) elements: dims {inner, outer}, 128B swizzle by default, zero fill out of bounds.
// The encoded descriptor is a pure function of the arguments, and a training step presents the same few hundred
// (buffer, shape) combinations again and again, so descriptors are kept in a small direct-mapped cache
// (cuTensorMapEncodeTiled was ~1/3 of the host time of a launch).
struct MapKey {
  const void* ptr;
  uint64_t inner, outer, pitch;
  uint32_t box_inner, box_outer;
  int swizzle, elem_bytes;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && pitch == o.pitch && box_inner == o.box_inner &&
           box_outer == o.box_outer && swizzle == o.swizzle && elem_bytes == o.elem_bytes;
  }
};
struct MapSlot {
  MapKey key;
  CUtensorMap map;
  bool valid;
};
constexpr int MAP_CACHE_SLOTS = 2048;
MapSlot g_map_cache[MAP_CACHE_SLOTS];
std::mutex g_map_mutex;

int make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner,
             uint32_t box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, int elem_bytes = 2) {
  const MapKey key{ptr, inner, outer, pitch_elems, box_inner, box_outer, static_cast<int>(swizzle), elem_bytes};
  uint64_t h = reinterpret_cast<uint64_t>(ptr) * 0x9E3779B97F4A7C15ull;
  h ^= (inner * 0xC2B2AE3D27D4EB4Full) ^ (outer * 0x165667B19E3779F9ull) ^ (pitch_elems << 17) ^
       (static_cast<uint64_t>(box_inner) << 40) ^ (static_cast<uint64_t>(box_outer) << 48) ^
       (static_cast<uint64_t>(swizzle) << 56) ^ static_cast<uint64_t>(elem_bytes);
  h ^= h >> 29;
  MapSlot& slot = g_map_cache[h % MAP_CACHE_SLOTS];
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    if (slot.valid && slot.key == key) {
      *m = slot.map;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable");
    return ALPRO_EDRIVER;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu pitch=%llu box=%ux%u ptr=%p", (int)r,
                   (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner,
                   box_outer, ptr);
    return ALPRO_EINVAL;
  }
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    slot.key = key;
    slot.map = *m;
    slot.valid = true;
  }
  return 0;
}

}  // namespace

namespace gemm2 {
int launch_2cta(int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const gemm::GemmKParams& p, int grid,
                cudaStream_t st);
}

namespace gemm3 {
int launch_2cta_tma_epi(int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                        const CUtensorMap& tmO2, const CUtensorMap& tmAux, const gemm::GemmKParams& p, int grid,
                        cudaStream_t st);
}

// ALPRO_GEMM_TMA_EPI=0 keeps the 16-bit output modes on the register/ST.G epilogue of gemm_tc2.cu (read per call).
static bool use_tma_epilogue() {
  const char* e = getenv("ALPRO_GEMM_TMA_EPI");
  return !(e && e[0] == '0');
}

// ALPRO_GEMM_2CTA=0 selects the single-CTA kernel (default: CTA-pair kernel, cta_group::2).
static bool use_2cta() {   // read per call (like every other switch) so that tests can run both kernels in one process
  const char* e = getenv("ALPRO_GEMM_2CTA");
  return !(e && e[0] == '0');
}
}  // namespace alpro

using namespace alpro;

extern "C" int alpro_gemm16(const void* A, const void* B, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                            int a_layout, int b_layout, int a_fmt, int b_fmt, const AlproGemmEpilogue* ep,
                            void* stream) {
  ALPRO_REQUIRE(A && B && ep, "alpro_gemm16: null operand");
  ALPRO_REQUIRE(M > 0 && N > 0 && K > 0, "alpro_gemm16: empty problem M=%lld N=%lld K=%lld", (long long)M,
                (long long)N, (long long)K);
  ALPRO_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "alpro_gemm16: dimension too large");
  ALPRO_REQUIRE(aligned16(A) && aligned16(B), "alpro_gemm16: operands must be 16-byte aligned");
  ALPRO_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0, "alpro_gemm16: lda/ldb must be multiples of 8 (got %lld, %lld)",
                (long long)lda, (long long)ldb);
  ALPRO_REQUIRE(lda >= (a_layout == ALPRO_KMAJOR ? K : M) && ldb >= (b_layout == ALPRO_KMAJOR ? K : N),
                "alpro_gemm16: leading dimension smaller than the contiguous extent");
  ALPRO_REQUIRE(ep->out32 || ep->out16, "alpro_gemm16: no output");
  ALPRO_REQUIRE((a_fmt | 1) == 1 && (b_fmt | 1) == 1, "alpro_gemm16: fmt must be 0 (fp16) or 1 (bf16)");
  if (ep->act == ALPRO_ACT_GELU_GRAD || ep->act == ALPRO_ACT_RELU_GRAD)
    ALPRO_REQUIRE(ep->aux16, "alpro_gemm16: *_GRAD activation needs aux16");

  const bool pair = use_2cta();
  GemmKParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.num_m_tiles = (int)cdiv(M, pair ? 2 * BM : BM);
  p.num_n_tiles = (int)cdiv(N, BN);
  p.num_k_blocks = (int)cdiv(K, BK);
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int split = ep->split_k;
  if (split == -1) {
    split = 1;
    const int units = pair ? num_sms() / 2 : num_sms();
    if (tiles < units) split = (2 * units) / tiles;
    if (split > p.num_k_blocks / 4) split = p.num_k_blocks / 4;
    // Default since r02 (tools/probe_gemm.py, isolated launches: 768x768x50208 1037 -> 1217 TFLOP/s, 768x3072x50208
    // 1270 -> 1431, 2304x768x50176 unchanged): pick the split that minimises waves * (k-blocks per item + epilogue),
    // the epilogue of a split-K item (TMA reduce-add of a 128x256 fp32 tile per CTA) costed as 8 k-blocks. The rule
    // above filled two waves regardless of how unevenly. ALPRO_GEMM_SPLITK=two_waves restores it (read once).
    static const char* sk_env = getenv("ALPRO_GEMM_SPLITK");
    if (!(sk_env && sk_env[0] == 't') && tiles < 4 * units) {
      const int epi = 8;
      long best = -1;
      int best_s = 1;
      const int smax = p.num_k_blocks / 4 > 0 ? p.num_k_blocks / 4 : 1;
      for (int sct = 1; sct <= smax && sct <= 64; ++sct) {
        const long kb = cdiv(p.num_k_blocks, sct);
        const long items = static_cast<long>(tiles) * cdiv(p.num_k_blocks, kb);
        const long waves = cdiv(items, units);
        const long cost = waves * (kb + (sct > 1 ? epi : 2));
        if (best < 0 || cost < best) { best = cost; best_s = sct; }
      }
      split = best_s;
    }
  }
  if (split < 1) split = 1;
  if (split > p.num_k_blocks) split = p.num_k_blocks;
  p.kb_per_split = (int)cdiv(p.num_k_blocks, split);
  split = (int)cdiv(p.num_k_blocks, p.kb_per_split);  // no empty splits
  p.split_k = split;
  if (ep->split_k != 0 && ep->split_k != 1) {
    ALPRO_REQUIRE(ep->out32 && !ep->out16 && !ep->out16b && !ep->bias && !ep->resid && ep->act == ALPRO_ACT_NONE,
                  "alpro_gemm16: split-K supports only out32 += alpha*acc");
    ALPRO_REQUIRE(!ep->row_scale_acc, "alpro_gemm16: split-K does not take row scales");
  }
  p.a_mn = a_layout == ALPRO_MNMAJOR;
  p.b_mn = b_layout == ALPRO_MNMAJOR;
  p.idesc = make_idesc_f16(a_fmt, b_fmt, p.a_mn, p.b_mn, pair ? 2 * BM : BM, BN);
  p.bias = ep->bias;
  p.aux16 = static_cast<const uint16_t*>(ep->aux16);
  p.resid = ep->resid;
  p.out32 = ep->out32;
  p.out16 = static_cast<uint16_t*>(ep->out16);
  p.out16b = static_cast<uint16_t*>(ep->out16b);
  p.ld32 = ep->ld32; p.ld16 = ep->ld16; p.ld16b = ep->ld16b; p.ldresid = ep->ldresid; p.ldaux = ep->ldaux;
  p.out16_fmt = ep->out16_fmt; p.out16b_fmt = ep->out16b_fmt; p.aux_fmt = ep->aux_fmt;
  p.act = ep->act;
  p.skip_period = ep->skip_period;
  p.alpha = ep->alpha;
  p.rs_acc = ep->row_scale_acc;
  p.rs_bias = ep->row_scale_bias;
  p.bias2 = ep->bias2;
  ALPRO_REQUIRE(!p.bias2 || (p.resid && p.out32 && p.act == ALPRO_ACT_NONE),
                "alpro_gemm16: bias2 is an option of the residual epilogue (resid + out32, no activation)");
  bool vec = true;
  if (p.bias) vec = vec && aligned16(p.bias);
  if (p.bias2) vec = vec && aligned16(p.bias2);
  if (p.out32) vec = vec && aligned16(p.out32) && (p.ld32 % 4) == 0;
  if (p.resid) vec = vec && aligned16(p.resid) && (p.ldresid % 4) == 0;
  if (p.out16) vec = vec && aligned16(p.out16) && (p.ld16 % 4) == 0;
  if (p.out16b) vec = vec && aligned16(p.out16b) && (p.ld16b % 4) == 0;
  if (p.aux16) vec = vec && aligned16(p.aux16) && (p.ldaux % 4) == 0;
  p.vec_ok = vec ? 1 : 0;

  CUtensorMap tmA, tmB;
  int rc;
  if (!p.a_mn) rc = make_map(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  else         rc = make_map(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  if (rc) return rc;
  if (!p.b_mn) rc = make_map(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, pair ? BN / 2 : BN);
  else         rc = make_map(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return rc;

  int mode = E_GENERIC;
  if (p.split_k > 1) mode = E_ATOMIC;
  else if (p.act == ALPRO_ACT_NONE && !p.resid && !p.out32 && p.out16 && !p.out16b) mode = E_OUT16;
  else if (p.act == ALPRO_ACT_GELU && !p.resid && !p.out32 && p.out16) mode = E_GELU_SAVE;
  else if (p.act == ALPRO_ACT_GELU_GRAD && !p.resid && !p.out32 && p.out16 && !p.out16b) mode = E_GELU_GRAD;
  else if (p.act == ALPRO_ACT_NONE && p.resid && p.out32 && !p.out16b) mode = E_RESID_OUT32;
  else if (p.act == ALPRO_ACT_NONE && !p.resid && p.out32 && !p.out16 && !p.out16b) mode = E_OUT32;
  const int work = tiles * p.split_k;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pair) {
    const int clusters = work < num_sms() / 2 ? work : num_sms() / 2;
    // TMA-store epilogue (gemm_tc3.cu): outputs (and the residual / saved-derivative input) whose rows the TMA unit can
    // address (16-byte aligned base and pitch), whole 32-column chunks, one 16-bit storage format.
    bool tma_epi = use_tma_epilogue() && (N % 32) == 0 && (!p.bias || aligned16(p.bias)) &&
                   (!p.bias2 || aligned16(p.bias2));
    if (mode == E_OUT16 || mode == E_GELU_SAVE || mode == E_GELU_GRAD) {
      tma_epi = tma_epi && aligned16(p.out16) && (p.ld16 % 8) == 0;
      if (mode == E_GELU_SAVE && p.out16b)
        tma_epi = tma_epi && aligned16(p.out16b) && (p.ld16b % 8) == 0 && p.out16b_fmt == p.out16_fmt;
      if (mode == E_GELU_GRAD)
        tma_epi = tma_epi && !p.rs_acc && aligned16(p.aux16) && (p.ldaux % 8) == 0 && p.aux_fmt == p.out16_fmt;
    } else if (mode == E_RESID_OUT32) {
      tma_epi = tma_epi && !p.out16 && aligned16(p.out32) && (p.ld32 % 4) == 0 && aligned16(p.resid) &&
                (p.ldresid % 4) == 0;
    } else if (mode == E_GENERIC && p.act == ALPRO_ACT_GELU_GRAD && p.resid && p.out32 && !p.out16 && !p.out16b &&
               p.aux16 && !p.bias2 && p.skip_period == 0) {
      // (acc + bias) * aux + resid — the dropout-mask form of the BERT output projections: the residual kernel with a
      // multiplier tile (the generic epilogue of gemm_tc2.cu ran these at 257 / 845 TFLOP/s, r02g shape table)
      tma_epi = tma_epi && aligned16(p.out32) && (p.ld32 % 4) == 0 && aligned16(p.resid) && (p.ldresid % 4) == 0 &&
                aligned16(p.aux16) && (p.ldaux % 8) == 0 && (N % 16) == 0;
      if (tma_epi) mode = E_RESID_OUT32;
    } else if (mode == E_ATOMIC) {   // split-K partial sums: TMA reduce-add instead of red.global per element
      tma_epi = tma_epi && aligned16(p.out32) && (p.ld32 % 4) == 0;
    } else {
      tma_epi = false;
    }
    if (tma_epi) {
      CUtensorMap tmO, tmO2, tmAux;
      if (mode == E_ATOMIC) {
        rc = make_map(&tmO, p.out32, (uint64_t)N, (uint64_t)M, (uint64_t)p.ld32, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B, 4);
        if (rc) return rc;
        tmO2 = tmO;
        tmAux = tmO;
      } else if (mode == E_RESID_OUT32) {
        rc = make_map(&tmO, p.out32, (uint64_t)N, (uint64_t)M, (uint64_t)p.ld32, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B, 4);
        if (rc) return rc;
        rc = make_map(&tmAux, p.resid, (uint64_t)N, (uint64_t)M, (uint64_t)p.ldresid, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B, 4);
        if (rc) return rc;
        tmO2 = tmO;
      } else {
        rc = make_map(&tmO, p.out16, (uint64_t)N, (uint64_t)M, (uint64_t)p.ld16, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
        tmO2 = tmO;
        tmAux = tmO;
        if (mode == E_GELU_SAVE && p.out16b)
          rc = make_map(&tmO2, p.out16b, (uint64_t)N, (uint64_t)M, (uint64_t)p.ld16b, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        if (mode == E_GELU_GRAD)
          rc = make_map(&tmAux, p.aux16, (uint64_t)N, (uint64_t)M, (uint64_t)p.ldaux, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
      }
      gemm3::launch_2cta_tma_epi(mode, tmA, tmB, tmO, tmO2, tmAux, p, 2 * clusters, st);
      ALPRO_CHECK_LAUNCH("alpro_gemm16(2cta, tma epilogue)");
      return 0;
    }
    gemm2::launch_2cta(mode, tmA, tmB, p, 2 * clusters, st);
    ALPRO_CHECK_LAUNCH("alpro_gemm16(2cta)");
    return 0;
  }
  const int grid = work < num_sms() ? work : num_sms();
#define ALPRO_LAUNCH_GEMM(M_)                                                                            \
  case M_: {                                                                                             \
    static std::once_flag once;                                                                          \
    std::call_once(once, [] {                                                                            \
      cudaFuncSetAttribute(gemm16_kernel<M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);  \
    });                                                                                                  \
    launch_k(gemm16_kernel<M_>, grid, NUM_THREADS, SMEM_BYTES, st, tmA, tmB, p);                               \
  } break;
  switch (mode) {
    ALPRO_LAUNCH_GEMM(E_OUT16)
    ALPRO_LAUNCH_GEMM(E_GELU_SAVE)
    ALPRO_LAUNCH_GEMM(E_GELU_GRAD)
    ALPRO_LAUNCH_GEMM(E_RESID_OUT32)
    ALPRO_LAUNCH_GEMM(E_OUT32)
    ALPRO_LAUNCH_GEMM(E_ATOMIC)
    default:
    ALPRO_LAUNCH_GEMM(E_GENERIC)
  }
#undef ALPRO_LAUNCH_GEMM
  ALPRO_CHECK_LAUNCH("alpro_gemm16");
  return 0;
}
