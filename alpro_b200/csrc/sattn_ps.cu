// Sequence-attention forward, persistent warp-specialised version for the TimeSformer spatial attention
// (129 <= S <= 256 tokens per (frame, head) unit, strided token rows with a shared cls token, no mask, no dropout).
//
// sattn_fwd_tc_kernel (attention.cu) runs one CTA per unit, two CTAs per SM, and every CTA walks load -> QK^T -> max ->
// exp -> PV -> store as serial phases behind __syncthreads: 12 k cycles per unit where the exponentials need 3.6 k
// (MUFU) and the instruction stream 5.4 k; ncu: issue slots 45 % busy, 64 % of the instructions are loop / address /
// barrier-spin code (profiles/r02r). This kernel keeps ONE CTA per SM resident over all its units:
//   * warp 0 streams the operands of the NEXT unit (three 4-D tensor-map boxes + the cls token's rows as 16-byte bulk
//     copies placed chunk by chunk into the swizzled layout) into a two-slot ring while the current unit is processed;
//   * a unit is two 128-query tiles; softmax warpgroup w (4 warps, one thread per query row, the whole row of <= 256
//     scores) takes tile (i + w) & 1 of unit i, so both groups do the same work over two units, each with its own 256
//     TMEM columns and its own MMA-issuing warp: one group exponentiates while the other waits for its PV product;
//   * P never touches shared memory: the 16-bit probabilities are written back by tcgen05.st over the score columns
//     they were computed from and PV reads its A operand from TMEM (tcgen05.mma [d], [a_tmem], b_desc), O accumulates
//     in dead score columns; TMEM loads are issued one chunk ahead of the arithmetic;
//   * O rows leave through a swizzled staging tile and ONE tensor-map store per tile.
// Same outputs as the other forward kernels (o, per-sequence cls outputs, base-2 log-sum-exp in token order).
// Reference semantics: Attention.forward, /root/reference/src/modeling/timesformer/vit.py:81-100, applied per frame to
// the '(b t) (1 + h w) m' rearrangement of Block.forward's spatial branch (vit.py:165-196): softmax(q k^T * 64^-0.5) v
// per head, with the clip's cls token prepended to every frame.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"

namespace alpro {
namespace sattn_ps {

constexpr int DH = 64;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int THREADS = 384;   // warp 0 TMA, 1 / 2 MMA issuers of group A / B, 3 TMEM owner, 4-7 group A, 8-11 group B

struct Params {
  const uint16_t* qkv;    // [rows, ld_qkv] (the cls token's rows are copied from here; everything else through tmBox)
  long long ld_qkv;
  uint16_t* cls_o;        // [nseq, d] token-0 outputs, or null (then the cls row goes to its canonical row of o)
  uint16_t* o;
  float* lse;             // [nseq, heads, S]
  long long ld_o, clip_rows;
  int S, heads, d, seq_div, n_units;
  float scale;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bulk_load_16(void* smem_dst, const void* gsrc, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(smem_u32(bar))
               : "memory");
}
// mbarrier wait with a watchdog: a protocol error in this many-barrier pipeline must end in a trap (reported as a
// launch failure), never in a kernel that spins until the box is reclaimed (2 s of SM clocks).
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s
  }
}
template <bool BF>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  if (BF) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

template <bool BF>
__global__ void __launch_bounds__(THREADS, 1)
sattn_fwd_ps_kernel(const __grid_constant__ CUtensorMap tmBox, const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1, const Params p) {
  extern __shared__ uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
  const int S = p.S;
  const int S16 = (S + 15) & ~15, S32 = (S + 31) & ~31;
  const int n_box = S - 1;                       // tokens 1..S-1 arrive as one box; token 0 (cls) sits in tile row S-1
  const uint32_t QB = static_cast<uint32_t>((S16 * 128 + 1023) & ~1023), KB = static_cast<uint32_t>(S32 * 128);
  const uint32_t SLOT = QB + 2 * KB;
  uint8_t* stage_o = sm + 2 * SLOT;              // 2 x 16 KB: O staging tile of each softmax group
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_o + 2 * 16384);
  uint64_t* full = bars;         // [2] TMA -> MMA: operands of the unit in slot s have landed
  uint64_t* empty = bars + 2;    // [2] MMA -> TMA: both tiles of the unit in slot s are done with it (2 commits)
  uint64_t* sfull = bars + 4;    // [2] MMA -> softmax group w: scores in TMEM
  uint64_t* pready = bars + 6;   // [2] softmax group w -> MMA: probabilities in TMEM (4 warp arrivals)
  uint64_t* ofull = bars + 8;    // [2] MMA -> softmax group w: O in TMEM
  uint64_t* tfree = bars + 10;   // [2] softmax group w -> MMA: O read out, the group's TMEM region is free (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 2);
      mbar_init(&sfull[i], 1);
      mbar_init(&pready[i], 4);
      mbar_init(&ofull[i], 1);
      mbar_init(&tfree[i], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmBox);
    tma_prefetch_desc(&tmO0);
    tma_prefetch_desc(&tmO1);
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // K / V rows [S, S32) are never written by a box: zero them once (padded keys must contribute exactly 0 to P V)
  for (int i = tid; i < 2 * 2 * (S32 - S) * 8; i += THREADS) {
    const int c8 = i & 7, rr = S + ((i >> 3) % (S32 - S)), which = (i >> 3) / (S32 - S);   // which: slot * 2 + {K, V}
    uint8_t* tile = sm + (which >> 1) * SLOT + QB + (which & 1) * KB;
    *reinterpret_cast<uint4*>(tile + rr * 128 + (c8 << 4)) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_grid_sync();

  const int n_local = (p.n_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  constexpr int fmt = BF ? 1 : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    for (int i = 0; i < n_local; ++i) {
      const int slot = i & 1;
      const int g = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const int head = g % p.heads, seq = g / p.heads;
      const int c_t = seq % p.seq_div, c_b = seq / p.seq_div;
      mbar_wait_wd(&empty[slot], ((i >> 1) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* q = sm + slot * SLOT;
        uint8_t* k = q + QB;
        uint8_t* v = k + KB;
        mbar_arrive_expect_tx(&full[slot], static_cast<uint32_t>(3 * S * 128));
        tma_load_4d(k, &tmBox, &full[slot], p.d + head * DH, 0, c_t, c_b);
        tma_load_4d(q, &tmBox, &full[slot], head * DH, 0, c_t, c_b);
        tma_load_4d(v, &tmBox, &full[slot], 2 * p.d + head * DH, 0, c_t, c_b);
        // token 0 -> tile row S-1: eight 16-byte chunks per matrix, each to its 128B-swizzle position
        const uint16_t* cls = p.qkv + c_b * p.clip_rows * p.ld_qkv + head * DH;
        const uint32_t roff = static_cast<uint32_t>(n_box * 128), rx = static_cast<uint32_t>(n_box & 7);
#pragma unroll
        for (uint32_t c8 = 0; c8 < 8; ++c8) {
          const uint32_t off = roff + ((c8 ^ rx) << 4);
          bulk_load_16(q + off, cls + c8 * 8, &full[slot]);
          bulk_load_16(k + off, cls + p.d + c8 * 8, &full[slot]);
          bulk_load_16(v + off, cls + 2 * p.d + c8 * 8, &full[slot]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------ MMA issuer of softmax group w
    const int w = warp - 1;
    const uint32_t region = tmem + static_cast<uint32_t>(w * 256);
    const uint32_t idesc_s = make_idesc_f16(fmt, fmt, 0, 0, 128, S32);
    const uint32_t idesc_o = make_idesc_f16(fmt, fmt, 0, 1, 128, DH);
    const int nks = S32 >> 4;
    for (int i = 0; i < n_local; ++i) {
      const int slot = i & 1, t = (i + w) & 1;
      const uint32_t qa = smem_u32(sm + slot * SLOT) + static_cast<uint32_t>(t * 16384);
      const uint32_t ka = smem_u32(sm + slot * SLOT + QB), va = ka + KB;
      mbar_wait_wd(&full[slot], (i >> 1) & 1);
      mbar_wait_wd(&tfree[w], (i & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(region, make_smem_desc_sw128(qa + ks * 32, 16, 1024), make_smem_desc_sw128(ka + ks * 32, 16, 1024),
                   idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(&sfull[w]);
      }
      __syncwarp();
      mbar_wait_wd(&pready[w], i & 1);
      tc_fence_after();
      if (elect_one()) {
        for (int ks = 0; ks < nks; ++ks)
          umma_f16_ts(region + 128, region + static_cast<uint32_t>(ks * 8), make_smem_desc_sw128(va + ks * 2048, 8192, 1024),
                      idesc_o, ks > 0 ? 1u : 0u);
        umma_commit(&ofull[w]);
        umma_commit(&empty[slot]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax group w: thread = one query row
    const int w = (warp - 4) >> 2, quarter = warp & 3;
    const int rit = quarter * 32 + lane;                      // row inside the tile = TMEM lane
    const uint32_t region = tmem + static_cast<uint32_t>(w * 256);
    const uint32_t trow = region + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* stg = stage_o + w * 16384;
    const bool leader = (quarter == 0) && (lane == 0);        // issues the group's tensor-map stores
    const float sl2 = p.scale * LOG2E;
    const int nchunk = S32 >> 5;
    for (int i = 0; i < n_local; ++i) {
      const int t = (i + w) & 1;
      const int g = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const int head = g % p.heads, seq = g / p.heads;
      const int rows_valid = t == 0 ? 128 : S - 128;
      const bool live = quarter * 32 < rows_valid;            // warp-uniform: a warp without valid rows only signals
      const int row = t * 128 + rit;                          // tile row of the unit; valid while < S
      float m = 0.f, l = 0.f;
      mbar_wait_wd(&sfull[w], i & 1);
      tc_fence_after();
      if (live) {
        uint32_t ra[32], rb[32];
        // ---- pass 1: row maximum of the raw scores (scale > 0: scaled once)
        float x0 = -INFINITY, x1 = -INFINITY;
        tmem_ld_32x32(trow, ra);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < nchunk) {
            tmem_ld_wait();
            uint32_t (&cur)[32] = (c & 1) ? rb : ra;
            if (c + 1 < nchunk) tmem_ld_32x32(trow + (c + 1) * 32, (c & 1) ? ra : rb);
            if ((c + 1) * 32 <= S) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                x0 = fmaxf(x0, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
                x1 = fmaxf(x1, fmaxf(__uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3])));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c * 32 + j < S) x0 = fmaxf(x0, __uint_as_float(cur[j]));
            }
          }
        }
        m = fmaxf(x0, x1) * sl2;
        const float nm = -m;
        // ---- pass 2: probabilities -> 16-bit -> back into TMEM over the consumed score columns
        float l0 = 0.f, l1 = 0.f;
        tmem_ld_32x32(trow, ra);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < nchunk) {
            tmem_ld_wait();
            uint32_t (&cur)[32] = (c & 1) ? rb : ra;
            if (c + 1 < nchunk) tmem_ld_32x32(trow + (c + 1) * 32, (c & 1) ? ra : rb);
            uint32_t pk[16];
            const bool fullc = (c + 1) * 32 <= S;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float a = ex2(fmaf(__uint_as_float(cur[j]), sl2, nm));
              float b = ex2(fmaf(__uint_as_float(cur[j + 1]), sl2, nm));
              if (!fullc) {
                if (c * 32 + j >= S) a = 0.f;
                if (c * 32 + j + 1 >= S) b = 0.f;
              }
              l0 += a;
              l1 += b;
              pk[j >> 1] = pack2<BF>(a, b);
            }
            tmem_st_32x16(trow + c * 16, pk);
          }
        }
        l = l0 + l1;
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pready[w]);
      // ---- O = P V of this tile -> registers; the group's TMEM region is free again after that
      mbar_wait_wd(&ofull[w], i & 1);
      tc_fence_after();
      uint32_t ow[32];   // 64 output columns as 32 packed pairs
      if (live) {
        const float inv = 1.f / l;
        uint32_t r[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tmem_ld_32x32(trow + 128 + h * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            ow[h * 16 + (j >> 1)] = pack2<BF>(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tfree[w]);
      // ---- staging tile -> one tensor-map store (tokens 1..S-1); the cls row and the log-sum-exp go out directly
      if (leader) bulk_wait_read<0>();           // the previous store of this group has drained the staging tile
      named_bar_sync(1 + w, 128);
      if (live) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4)
          *reinterpret_cast<uint4*>(stg + rit * 128 + ((q4 ^ (rit & 7)) << 4)) =
              make_uint4(ow[4 * q4], ow[4 * q4 + 1], ow[4 * q4 + 2], ow[4 * q4 + 3]);
        if (row < S) {
          const int tok = row == S - 1 ? 0 : row + 1;
          p.lse[(static_cast<long long>(seq) * p.heads + head) * S + tok] = m + log2f(l);
          if (row == S - 1) {
            uint16_t* dst = p.cls_o ? p.cls_o + static_cast<long long>(seq) * p.d + head * DH
                                    : p.o + (seq / p.seq_div) * p.clip_rows * p.ld_o + head * DH;
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4)
              reinterpret_cast<uint4*>(dst)[q4] = make_uint4(ow[4 * q4], ow[4 * q4 + 1], ow[4 * q4 + 2], ow[4 * q4 + 3]);
          }
        }
      }
      fence_proxy_async();
      named_bar_sync(1 + w, 128);
      if (leader) {
        const int c_t = seq % p.seq_div, c_b = seq / p.seq_div;
        if (t == 0) tma_store_4d(&tmO0, stg, head * DH, 0, c_t, c_b);
        else tma_store_4d(&tmO1, stg, head * DH, 128, c_t, c_b);
        bulk_commit();
      }
    }
    if (leader) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeFn>(f);
    return static_cast<EncodeFn>(nullptr);
  }();
  return fn;
}

// 4-D map (column, token, frame, clip): token i of a box is row clip*clip_rows + 1 + frame + i*stride of `mat`
static bool map4(CUtensorMap* m, const uint16_t* mat, int64_t ld, int cols, int n_tok, int stride, int seq_div, int groups,
                 int64_t clip_rows, int box_rows) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(n_tok), static_cast<cuuint64_t>(seq_div),
                        static_cast<cuuint64_t>(groups)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(stride) * ld * 2, static_cast<cuuint64_t>(ld) * 2,
                           static_cast<cuuint64_t>(clip_rows) * ld * 2};
  cuuint32_t box[4] = {DH, static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<uint16_t*>(mat + ld), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace sattn_ps

// ALPRO_OK: launched; ALPRO_ENOTSUP: the problem is outside this kernel's scope (caller falls back).
int sattn_fwd_ps(const void* qkv, int64_t ld_qkv, void* o, int64_t ld_o, void* cls_o, float* lse, int S, int nseq, int heads,
                 int fmt, int seq_div, int stride, int64_t clip_rows, float scale, cudaStream_t st) {
  using namespace sattn_ps;
  const char* e = getenv("ALPRO_ATTN_PS");   // default on; ALPRO_ATTN_PS=0 keeps the per-unit kernel (read per call)
  if (e && e[0] == '0') return ALPRO_ENOTSUP;
  if (S <= 129 || S > 256 || (seq_div == 1 && stride == 1) || !encode_fn() || !lse) return ALPRO_ENOTSUP;
  const int d = heads * DH;
  if (!aligned16(qkv) || !aligned16(o) || (ld_qkv % 8) || (ld_o % 8) || 3 * d > ld_qkv || d > ld_o || nseq % seq_div ||
      (cls_o && !aligned16(cls_o)))
    return ALPRO_ENOTSUP;
  const int groups = nseq / seq_div, n_box = S - 1;
  const uint16_t* q16 = static_cast<const uint16_t*>(qkv);
  uint16_t* o16 = static_cast<uint16_t*>(o);
  CUtensorMap tmBox, tmO0, tmO1;
  if (!map4(&tmBox, q16, ld_qkv, 3 * d, n_box, stride, seq_div, groups, clip_rows, n_box)) return ALPRO_ENOTSUP;
  if (!map4(&tmO0, o16, ld_o, d, n_box, stride, seq_div, groups, clip_rows, 128)) return ALPRO_ENOTSUP;
  if (!map4(&tmO1, o16, ld_o, d, n_box, stride, seq_div, groups, clip_rows, n_box - 128)) return ALPRO_ENOTSUP;
  Params p;
  p.qkv = q16; p.ld_qkv = ld_qkv; p.cls_o = static_cast<uint16_t*>(cls_o); p.o = o16; p.lse = lse; p.ld_o = ld_o; p.clip_rows = clip_rows; p.S = S;
  p.heads = heads; p.d = d; p.seq_div = seq_div; p.n_units = nseq * heads; p.scale = scale;
  const int S16 = (S + 15) & ~15, S32 = (S + 31) & ~31;
  const size_t QB = static_cast<size_t>((S16 * 128 + 1023) & ~1023), KB = static_cast<size_t>(S32) * 128;
  const size_t smem = 1024 + 2 * (QB + 2 * KB) + 2 * 16384 + 12 * 8 + 64;
  if (smem > 232448) return ALPRO_ENOTSUP;
  int grid = p.n_units < num_sms() ? p.n_units : num_sms();
  if (fmt == 1) {
    cudaFuncSetAttribute(sattn_fwd_ps_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    launch_k(sattn_fwd_ps_kernel<true>, grid, THREADS, smem, st, tmBox, tmO0, tmO1, p);
  } else {
    cudaFuncSetAttribute(sattn_fwd_ps_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    launch_k(sattn_fwd_ps_kernel<false>, grid, THREADS, smem, st, tmBox, tmO0, tmO1, p);
  }
  return ALPRO_OK;
}

}  // namespace alpro
