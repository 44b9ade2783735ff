// alpro_gemm16: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0 (1 lane) : TMA producer   — cp.async.bulk.tensor 128B-swizzled operand tiles into a 4-stage smem ring
//   warp 1 (1 lane) : MMA issuer     — tcgen05.mma.cta_group::1.kind::f16, 128x256x16 atoms, fp32 accumulators in TMEM
//   warp 2          : TMEM allocator — 512 columns = two 128x256 fp32 accumulator buffers (epilogue/mainloop overlap)
//   warps 4..7      : epilogue       — tcgen05.ld TMEM->registers, fused bias/GELU/GELU'/residual, 128-bit stores
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
#include <mutex>

#include "common.h"
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
constexpr int NUM_THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = 1024 + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256;

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
  int vec_ok;  // all leading dims / pointers allow 128-bit accesses
  float alpha;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Epilogue for 8 consecutive columns of one row.
__device__ __forceinline__ void epilogue8(const GemmKParams& p, float (&v)[8], long long row, int col, bool atomic) {
  if (atomic) {
    float* o = p.out32 + row * p.ld32 + col;
    if (p.vec_ok && col + 8 <= p.N) {
      red_add_v4(o, v[0], v[1], v[2], v[3]);
      red_add_v4(o + 4, v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (col + j < p.N) atomicAdd(o + j, v[j]);
    }
    return;
  }
  const bool full = p.vec_ok && (col + 8 <= p.N);
  if (p.bias) {
    if (full) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (col + j < p.N) v[j] += __ldg(p.bias + col + j);
    }
  }
  if (p.act != ALPRO_ACT_NONE) {
    if (p.act == ALPRO_ACT_GELU || p.act == ALPRO_ACT_RELU) {
      if (p.out16b) {  // save the pre-activation for the backward pass
        uint16_t* o = p.out16b + row * p.ld16b + col;
        if (full) {
          uint4 w;
          w.x = pack2_16(v[0], v[1], p.out16b_fmt); w.y = pack2_16(v[2], v[3], p.out16b_fmt);
          w.z = pack2_16(v[4], v[5], p.out16b_fmt); w.w = pack2_16(v[6], v[7], p.out16b_fmt);
          *reinterpret_cast<uint4*>(o) = w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (col + j < p.N) o[j] = f32_to_16(v[j], p.out16b_fmt);
        }
      }
      if (p.act == ALPRO_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
      }
    } else {  // *_GRAD: multiply by f'(aux)
      float u[8];
      const uint16_t* a = p.aux16 + row * p.ldaux + col;
      if (full) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(a));
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u[2 * j] = f16_to_32(static_cast<uint16_t>(ww[j] & 0xffff), p.aux_fmt);
          u[2 * j + 1] = f16_to_32(static_cast<uint16_t>(ww[j] >> 16), p.aux_fmt);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = (col + j < p.N) ? f16_to_32(a[j], p.aux_fmt) : 0.f;
      }
      if (p.act == ALPRO_ACT_GELU_GRAD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= gelu_erf_grad(u[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = u[j] > 0.f ? v[j] : 0.f;
      }
    }
  }
  if (p.resid) {
    const bool skip = p.skip_period > 0 && (row % p.skip_period) == 0;
    const float* r = p.resid + row * p.ldresid + col;
    if (full) {
      const float4 r0 = *reinterpret_cast<const float4*>(r);
      const float4 r1 = *reinterpret_cast<const float4*>(r + 4);
      if (skip) {
        v[0] = r0.x; v[1] = r0.y; v[2] = r0.z; v[3] = r0.w; v[4] = r1.x; v[5] = r1.y; v[6] = r1.z; v[7] = r1.w;
      } else {
        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (col + j < p.N) v[j] = skip ? r[j] : v[j] + r[j];
    }
  }
  if (p.out32) {
    float* o = p.out32 + row * p.ld32 + col;
    if (full) {
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (col + j < p.N) o[j] = v[j];
    }
  }
  if (p.out16) {
    uint16_t* o = p.out16 + row * p.ld16 + col;
    if (full) {
      uint4 w;
      w.x = pack2_16(v[0], v[1], p.out16_fmt); w.y = pack2_16(v[2], v[3], p.out16_fmt);
      w.z = pack2_16(v[4], v[5], p.out16_fmt); w.w = pack2_16(v[6], v[7], p.out16_fmt);
      *reinterpret_cast<uint4*>(o) = w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (col + j < p.N) o[j] = f32_to_16(v[j], p.out16_fmt);
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment in the shared window.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE_BYTES);
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
      mbar_init(&tempty_bar[a], 128);
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
    // ------------------------------------------------------------ epilogue (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool atomic = p.split_k > 1;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int tile = w / p.split_k;
      const int m_blk = tile / p.num_n_tiles;
      const int n_blk = tile - m_blk * p.num_n_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const long long row = static_cast<long long>(m_blk) * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        const int col0 = n_blk * BN + c;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
        if (row < p.M) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            if (col < p.N) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) * p.alpha;
              epilogue8(p, v, row, col, atomic);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
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

// 2-D tensor map over 16-bit elements: dims {inner, outer}, 128B swizzle, zero fill out of bounds.
int make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner,
             uint32_t box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable");
    return ALPRO_EDRIVER;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu pitch=%llu box=%ux%u ptr=%p", (int)r,
                   (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner,
                   box_outer, ptr);
    return ALPRO_EINVAL;
  }
  return 0;
}

}  // namespace
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

  GemmKParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.num_m_tiles = (int)cdiv(M, BM);
  p.num_n_tiles = (int)cdiv(N, BN);
  p.num_k_blocks = (int)cdiv(K, BK);
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int split = ep->split_k;
  if (split == -1) {
    split = 1;
    if (tiles < num_sms()) split = (2 * num_sms()) / tiles;
    if (split > p.num_k_blocks / 4) split = p.num_k_blocks / 4;
  }
  if (split < 1) split = 1;
  if (split > p.num_k_blocks) split = p.num_k_blocks;
  p.kb_per_split = (int)cdiv(p.num_k_blocks, split);
  split = (int)cdiv(p.num_k_blocks, p.kb_per_split);  // no empty splits
  p.split_k = split;
  if (ep->split_k != 0 && ep->split_k != 1) {
    ALPRO_REQUIRE(ep->out32 && !ep->out16 && !ep->out16b && !ep->bias && !ep->resid && ep->act == ALPRO_ACT_NONE,
                  "alpro_gemm16: split-K supports only out32 += alpha*acc");
  }
  p.a_mn = a_layout == ALPRO_MNMAJOR;
  p.b_mn = b_layout == ALPRO_MNMAJOR;
  p.idesc = make_idesc_f16(a_fmt, b_fmt, p.a_mn, p.b_mn, BM, BN);
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
  bool vec = true;
  if (p.bias) vec = vec && aligned16(p.bias);
  if (p.out32) vec = vec && aligned16(p.out32) && (p.ld32 % 4) == 0;
  if (p.resid) vec = vec && aligned16(p.resid) && (p.ldresid % 4) == 0;
  if (p.out16) vec = vec && aligned16(p.out16) && (p.ld16 % 8) == 0;
  if (p.out16b) vec = vec && aligned16(p.out16b) && (p.ld16b % 8) == 0;
  if (p.aux16) vec = vec && aligned16(p.aux16) && (p.ldaux % 8) == 0;
  p.vec_ok = vec ? 1 : 0;

  CUtensorMap tmA, tmB;
  int rc;
  if (!p.a_mn) rc = make_map(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  else         rc = make_map(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  if (rc) return rc;
  if (!p.b_mn) rc = make_map(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN);
  else         rc = make_map(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return rc;

  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  const int work = tiles * p.split_k;
  const int grid = work < num_sms() ? work : num_sms();
  gemm16_kernel<<<grid, NUM_THREADS, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  ALPRO_CHECK_LAUNCH("alpro_gemm16");
  return 0;
}
