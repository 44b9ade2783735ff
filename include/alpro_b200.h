/*
 * alpro_b200 C-ABI — the drop-in boundary underneath the Python task models.
 *
 * The reference (salesforce/ALPRO) has no FFI: its hot path is torch.nn.functional calls made from
 *   src/modeling/timesformer/vit.py   (Mlp :59-65, Attention :81-100, Block :136-213, PatchEmbed :233-239,
 *                                      VisionTransformer.forward_features :321-377, TimeSformer :475-503)
 *   src/modeling/xbert.py             (BertEmbeddings :186-213, BertSelfAttention :263-346, BertSelfOutput :349-360,
 *                                      BertIntermediate/BertOutput :412-438, BertLMPredictionHead :665-682)
 *   src/modeling/alpro_models.py      (VTC :103-128/:750-779, VTM :269-344/:800-872, MLM :346-373, MPM :209-232)
 * Each entry point below replaces one of those library calls (cited per function) with a hand-written sm_100a kernel.
 * Plain pointers + sizes + a cudaStream_t passed as void*; no torch types. All pointers are DEVICE pointers unless
 * stated otherwise. Every function returns 0 on success or a negative ALPRO_E* / positive cudaError_t code; the last
 * error text is available from alpro_last_error().
 *
 * 16-bit formats:  fmt 0 = IEEE fp16, fmt 1 = bf16.
 */
#ifndef ALPRO_B200_H
#define ALPRO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALPRO_OK 0
#define ALPRO_EINVAL (-1)   /* bad argument (shape / alignment / null pointer) */
#define ALPRO_EDRIVER (-2)  /* driver entry point (TMA descriptor encode) unavailable */
#define ALPRO_ENOTSUP (-3)

#define ALPRO_FMT_F16 0
#define ALPRO_FMT_BF16 1

#define ALPRO_ACT_NONE 0
#define ALPRO_ACT_GELU 1      /* out = gelu_erf(acc + bias)                 (nn.GELU, vit.py:50,61; ACT2FN['gelu'] xbert.py:417) */
#define ALPRO_ACT_GELU_GRAD 2 /* out = (acc) * gelu'(aux)                   (autograd of the above) */
#define ALPRO_ACT_RELU 3
#define ALPRO_ACT_RELU_GRAD 4 /* out = acc * (aux > 0) */

/* Operand storage: K-major  = the contraction index is contiguous in memory (x[M,K] row-major, nn.Linear weight [N,K]);
 *                  MN-major = the non-contracted index is contiguous (the same buffers read "transposed").          */
#define ALPRO_KMAJOR 0
#define ALPRO_MNMAJOR 1

const char* alpro_last_error(void);
int alpro_version(void);
int alpro_num_sms(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 tensor cores (TMA-staged operands, TMEM fp32 accumulators, fused epilogue).
 *   acc[m,n] = sum_k A(m,k) * B(n,k)
 *   v        = alpha * acc + bias[n]
 *   act      : see ALPRO_ACT_*  (pre-activation v optionally saved to out16b for the backward pass)
 *   v       += resid[m,n]       (fp32; rows with m % skip_period == 0 pass resid through unchanged when skip_period>0)
 *   out32[m,n] = v (fp32) and/or out16[m,n] = v (fmt out16_fmt)
 * Replaces F.linear / nn.Linear.forward on the path (vit.py:60,63,84,98,161; xbert.py:273-292,357,422,435,659,681) and
 * the dgrad / wgrad GEMMs autograd derives from them.
 *   A: a_layout K-major -> stored [M, lda]; MN-major -> stored [K, lda].   B likewise with N.
 *   lda/ldb in elements, multiples of 8; base pointers 16-byte aligned. M, N, K arbitrary (tails are masked).
 */
typedef struct AlproGemmEpilogue {
  const float* bias;     /* [N] or NULL */
  const void* aux16;     /* [M, ldaux] 16-bit pre-activation for *_GRAD acts, or NULL */
  const float* resid;    /* [M, ldresid] fp32 or NULL */
  float* out32;          /* [M, ld32] or NULL */
  void* out16;           /* [M, ld16] or NULL */
  void* out16b;          /* [M, ld16b] pre-activation copy (GELU/RELU) or NULL */
  int64_t ld32, ld16, ld16b, ldresid, ldaux;
  int32_t out16_fmt, out16b_fmt, aux_fmt;
  int32_t act;
  int32_t skip_period;
  int32_t split_k;       /* 0/1: none. >1: split the K range over that many CTAs. -1: auto. When split-K is active the
                            epilogue is out32 += alpha*acc with fp32 red.global.add (caller zeroes / pre-loads out32);
                            bias/act/resid/out16 must be unset. Used for weight-gradient contractions (K = #tokens). */
  float alpha;
} AlproGemmEpilogue;

int alpro_gemm16(const void* A, const void* B, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                 int a_layout, int b_layout, int a_fmt, int b_fmt, const AlproGemmEpilogue* ep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALPRO_B200_H */
