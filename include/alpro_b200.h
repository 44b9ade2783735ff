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
#define ALPRO_ACT_GELU 1      /* out = gelu_erf(acc + bias); out16b (optional) = gelu'(acc + bias)   (nn.GELU, vit.py:50,61; ACT2FN['gelu'] xbert.py:417) */
#define ALPRO_ACT_GELU_GRAD 2 /* out = acc * aux, aux = the gelu' saved by ALPRO_ACT_GELU               (autograd of the above) */
#define ALPRO_ACT_MUL_AUX 2   /* alias: out = (acc + bias) * aux      (nn.Dropout mask/keep multiply, xbert.py:358,436) */
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
 *   act      : see ALPRO_ACT_*  (GELU: derivative gelu'(v) / RELU: pre-activation v optionally saved to out16b for the backward pass)
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
  const float* row_scale_acc;  /* optional [M]: v = row_scale_acc[m]*alpha*acc + row_scale_bias[m]*bias   (DropPath: the
                                  per-sample mask/keep_prob of vit_utils.py:137-162 expanded to token rows) */
  const float* row_scale_bias; /* optional [M]; defaults to row_scale_acc */
  const float* bias2;          /* optional [N], residual modes only: a second bias that is NOT row-scaled,
                                  out = resid + row_scale_acc*alpha*acc + row_scale_bias*bias + bias2 (skip_period rows
                                  keep the plain residual). Lets one GEMM stand for Linear -> DropPath -> Linear:
                                  temporal_attn.proj -> drop_path -> temporal_fc, vit.py:157-161 */
} AlproGemmEpilogue;

int alpro_gemm16(const void* A, const void* B, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                 int a_layout, int b_layout, int a_fmt, int b_fmt, const AlproGemmEpilogue* ep, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * HBM-bound row kernels (alpro_b200/csrc/elementwise.cu)
 */
/* fp32 -> 16-bit cast of a contiguous buffer (weight operand copies; replaces nothing in the reference, which runs fp32) */
int alpro_cast_f32_to_16(const float* src, void* dst, int64_t n, int fmt, void* stream);
/* the same cast for MANY tensors in one launch: `table` (device, int64) holds num_chunks triples (src address, dst
 * address, element count <= 16384); src chunks 16-byte aligned, dst chunks 8-byte aligned. Used for the once-per-step
 * refresh of all 16-bit weight operands. */
int alpro_cast_f32_to_16_multi(const int64_t* table, int num_chunks, int fmt, void* stream);

/* LayerNorm over the last dim (nn.LayerNorm: vit.py:113,119,127,279 eps 1e-6; xbert.py:177,354,433,658 eps 1e-12).
 * x fp32 [M,d] -> out32 (optional) and/or out16 (optional); per-row mean / rstd saved for the backward. d%4==0, d<=1024.
 * mul16 (optional, same 16-bit format as out16): dropout mask/keep multiplied into the outputs. */
int alpro_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int64_t M, int d,
                        float* out32, int64_t ld32, void* out16, int64_t ld16, int out16_fmt, float* mean, float* rstd,
                        const void* mul16, int64_t ldmul, void* stream);
/* dy_kind 0 fp32 / 1 fp16 / 2 bf16. dx32 = (accumulate ? dx32 : 0) + LN'(dy); dx16 = 16-bit copy of the resulting dx32
 * (rows with row % zero_period == 0 written as zero when zero_period > 0). dgamma/dbeta += param_scale * sums (atomics).
 * colsum (optional) += param_scale * column sums of the resulting dx over rows with row % colsum_zero_period != 0: the
 * bias gradient of the Linear layer whose output gradient this dx is (saves a separate pass over dx).
 * Train-mode hooks (all optional): dy_mul16 = dropout mask on this LayerNorm's output; dx16_mul16 / dx16_row_scale =
 * dropout mask / DropPath row factor of the branch whose output gradient dx16 is; colsum_row_scale overrides the row
 * factor for the bias column sums. Masks use the 16-bit format of dx16. */
int alpro_layernorm_bwd(const void* dy, int dy_kind, int64_t lddy, const float* x, int64_t ldx, const float* mean,
                        const float* rstd, const float* gamma, int64_t M, int d, float* dx32, int64_t lddx,
                        int accumulate, void* dx16, int64_t lddx16, int dx16_fmt, int zero_period, float* dgamma,
                        float* dbeta, float param_scale, float* colsum, int colsum_zero_period, const void* dy_mul16,
                        int64_t lddymul, const void* dx16_mul16, int64_t lddxmul, const float* dx16_row_scale,
                        const float* colsum_row_scale, void* stream);
/* out[n] += alpha * sum_m x[m,n]; kind 0 fp32 / 1 fp16 / 2 bf16 (bias gradients of every nn.Linear on the path) */
int alpro_colsum(const void* x, int kind, int64_t ld, int64_t M, int N, float* out, float alpha, int zero_period,
                 void* stream);
/* PatchEmbed im2col (vit.py:233-239): frames fp32 [B,T,3,H,W] -> 16-bit [B*(1+N*T), 3*P*P], canonical token order
 * row = b*(1+N*T) + 1 + n*T + t, with a zero row in every clip's cls slot */
int alpro_patchify(const float* frames, void* out16, int fmt, int B, int T, int H, int W, int P, void* stream);
/* same from raw uint8 frames [B,T,3,H,W] with ImageNorm fused: (u8/255 - mean[c]) / std[c] (src/datasets/data_utils.py:437-457).
 * mean3 / std3 are HOST pointers to 3 floats. */
int alpro_patchify_u8(const uint8_t* frames, void* out16, int fmt, int B, int T, int H, int W, int P,
                      const float* mean3, const float* std3, void* stream);
/* x = cat(cls + pos[0], proj + pos[1+n] + time[t]) in 'b (n t)' order (vit.py:324-361) and its parameter gradients */
int alpro_vit_embed_fwd(const float* proj, const float* cls, const float* pos, const float* tim, float* x, int B, int N,
                        int T, int d, void* stream);
int alpro_vit_embed_bwd(const float* dx, float* dcls, float* dpos, float* dtim, int B, int N, int T, int d, float alpha,
                        void* stream);
/* temporal mean pooling (TimeSformer.forward_features vit.py:484-492): [B,1+N*T,d] -> [B,1+N,d] */
int alpro_temporal_pool_fwd(const float* xn, float* out, int B, int N, int T, int d, void* stream);
int alpro_temporal_pool_bwd(const float* dout, float* dxn, int B, int N, int T, int d, float alpha, void* stream);
/* BertEmbeddings (xbert.py:186-210) before its LayerNorm: word[ids] + type[0] + pos[l]; and the scatter-add backward */
int alpro_bert_embed_gather(const int64_t* ids, const float* word, const float* pos, const float* type, float* out,
                            int64_t BL, int L, int h, void* stream);
int alpro_bert_embed_scatter(const int64_t* ids, const float* de, float* dword, float* dpos, float* dtype, int64_t BL,
                             int L, int h, float alpha, void* stream);
/* fusion-encoder input: out[s] = cat(text[ti[s]], video[vi[s]]), additive key mask (1-mask)*-1e4
 * (alpro_models.py:273-275,318-331,354-357; xbert.py:878-938). ti/vi are int32 index lists on the device. */
int alpro_fusion_gather_fwd(const float* text, const float* video, const int64_t* tmask, const int32_t* ti,
                            const int32_t* vi, float* out32, void* out16, int fmt, float* add_mask, int S, int L, int Nv,
                            int h, void* stream);
int alpro_fusion_gather_bwd(const float* dout, const int32_t* ti, const int32_t* vi, float* dtext, float* dvideo, int S,
                            int L, int Nv, int h, void* stream);
/* mean over the T per-frame cls outputs [B,T,d] -> canonical cls row b*S of o (Block.forward vit.py:184-187) */
int alpro_cls_mean_fwd(const void* cls_t, void* o, int64_t ldo, int fmt, int B, int T, int S, int d,
                       const float* frame_weight, void* stream);
/* out[i] = keep_i/(1-p), keep_i ~ Bernoulli(1-p) from a stateless counter hash of (seed, i) (nn.Dropout sites xbert.py:178,358,436) */
int alpro_dropout_mask(void* out16, int fmt, int64_t n, float p, uint32_t seed, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Attention (alpro_b200/csrc/attention.cu), head_dim = 64
 */
/* temporal attention over the T frames of each patch position (vit.py:81-100 via :146-157); T in {1,2,4,8}.
 * qkv [B*(1+N*T), 3d] canonical rows; cls rows are skipped (zero-filled in the outputs).
 * T = 8 with 16-byte-aligned rows runs the TMA + tcgen05 kernels of tattn_tc.cu (16 units per 128-row tile);
 * ALPRO_TATTN_TC=0 in the environment selects the CUDA-core kernels (read on every call). */
int alpro_temporal_attn_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int B, int N, int T, int heads,
                            int fmt, float scale, void* stream);
int alpro_temporal_attn_bwd(const void* qkv, int64_t ld_qkv, const void* dout, int64_t ld_dout, void* dqkv,
                            int64_t ld_dqkv, int B, int N, int T, int heads, int fmt, float scale, void* stream);
/* sequence attention, S <= 256 keys, optional additive key mask [nseq,S] (vit.py:81-100 via :165-181;
 * xbert.py:263-346). Token j of sequence s lives at row (s/seq_div)*clip_rows + (j==0 ? 0 : 1 + s%seq_div + (j-1)*stride):
 *   BERT: seq_div=1, stride=1, clip_rows=S.   TimeSformer spatial: seq_div=T, stride=T, clip_rows=1+N*T, S=1+N.
 * With seq_div>1 the per-frame cls outputs go to cls_o [nseq,d] (mean taken by alpro_cls_mean_fwd). lse: [nseq,heads,S]. */
/* drop_p > 0: train-mode dropout of the attention probabilities (xbert.py:331) from the stateless hash (drop_seed).
 * Implementations (same outputs, lse format and dropout stream; the environment is read on every call):
 *   forward : tcgen05/TMEM kernel for S >= 96, mma.sync kernel below; ALPRO_ATTN_TC=0 / 1 forces mma.sync / tcgen05
 *   backward: tcgen05/TMEM kernel for 96 <= S <= 240, mma.sync otherwise; ALPRO_ATTN_BWD_TC=0 / 1 forces one of them */
int alpro_seq_attn_fwd(const void* qkv, int64_t ld_qkv, const float* mask, void* o, int64_t ld_o, void* cls_o, float* lse,
                       int S, int nseq, int heads, int fmt, int seq_div, int stride, int64_t clip_rows, float scale,
                       float drop_p, uint32_t drop_seed, void* stream);
/* backward: o_fwd = the forward output rows, cls_fwd = the forward per-sequence token-0 outputs (cls_o, seq_div > 1 only),
 * cls_weight (optional [nseq]) = weight of each frame's cls output in the group's cls row (default 1/seq_div) */
int alpro_seq_attn_bwd(const void* qkv, int64_t ld_qkv, const float* mask, const float* lse, const void* o_fwd,
                       const void* cls_fwd, const float* cls_weight, const void* dout, int64_t ld_o, void* dqkv,
                       float* dcls_qkv_scratch, int S,
                       int nseq, int heads, int fmt, int seq_div, int stride, int64_t clip_rows, float scale, float drop_p,
                       uint32_t drop_seed, void* stream);
/* test utility: materialises the mask/keep factors [nseq, heads, S, S] the two functions above apply for (drop_p, seed) */
int alpro_attn_dropout_mask(float* out, int S, int nseq, int heads, float drop_p, uint32_t drop_seed, void* stream);

/* Diagnostics (no reference counterpart): with ALPRO_ATTN_TRACE=1 in the environment the tcgen05 backward kernel
 * records 64 clock64() stamps per CTA (phase boundaries relative to CTA start). Copies the stamps of the last traced
 * launch to host_out (int64 values, synchronises the device); returns the number of CTAs copied, 0 if nothing was
 * traced, -1 on error. tools/check_attn_tc.py --cases trace:vit prints the per-phase medians. */
int alpro_debug_attn_trace(void* host_out, int64_t max_values);

/* ------------------------------------------------------------------------------------------------------------------
 * Task-head kernels, fp32 (alpro_b200/csrc/heads.cu)
 */
/* y = act(alpha' * x W^T + b); alpha' = alpha (mode 0), alpha * *alpha_dev (1), alpha / *alpha_dev (2). Replaces the
 * nn.Linear heads vision_proj/text_proj/itm_head/mpm_head and the sim = feat @ feats.t() / temp products
 * (alpro_models.py:103,115-116,205,337,226). */
int alpro_small_linear_fwd(const float* x, int64_t ldx, const float* W, int64_t ldw, const float* b, float* y, int64_t ldy,
                           int M, int N, int K, float alpha, const float* alpha_dev, int alpha_mode, int relu,
                           void* stream);
int alpro_small_linear_bwd(const float* dy, int64_t lddy, const float* yact, int64_t ldya, const float* x, int64_t ldx,
                           const float* W, int64_t ldw, float* dx, int64_t lddx, int dx_accumulate, float* dW,
                           int64_t lddw, float* db, int dw_accumulate, int M, int N, int K, float alpha,
                           const float* alpha_dev, int alpha_mode, float dw_scale, void* stream);
int alpro_l2norm_fwd(const float* x, float* y, float* norm, int M, int d, float eps, void* stream);
int alpro_l2norm_bwd(const float* dy, const float* y, const float* norm, float* dx, int M, int d, void* stream);
/* row-wise softmax cross-entropy with hard (int64, <0 ignored) or soft labels, loss = sum rows / denom
 * (denom_mode 0: #valid rows, 1: R). F.cross_entropy / -sum(log_softmax*targets) at alpro_models.py:125-128,228-232,339,371. */
int alpro_softmax_ce_fwd(const float* logits, int64_t ld, int R, int C, const int64_t* hard, const float* soft,
                         int64_t ld_soft, const uint8_t* row_ignore, float* row_loss, float* row_lse, float* row_valid,
                         float* row_tsum, int denom_mode, float* loss_out, float* denom_out, void* stream);
int alpro_softmax_ce_bwd(const float* logits, int64_t ld, int R, int C, const int64_t* hard, const float* soft,
                         int64_t ld_soft, const float* row_lse, const float* row_valid, const float* row_tsum,
                         const float* denom, const float* gptr, float gscale, float* out32, void* out16, int out16_fmt,
                         int64_t ld_out, int C_out, void* stream);
int alpro_temp_grad(const float* dsa, const float* sa, int64_t na, const float* dsb, const float* sb, int64_t nb,
                    const float* temp, float* dtemp, float coef, void* stream);
int alpro_clamp_scalar(float* p, float lo, float hi, void* stream); /* temp.clamp_(0.001, 0.5), alpro_models.py:80-81 */
int alpro_masked_mean_fwd(const float* x, int64_t seq_stride, int row0, const float* patch_mask, int B, int Np, int h,
                          float* out, void* stream);
int alpro_masked_mean_bwd(const float* dout, const float* patch_mask, int B, int Np, int h, float* dx, int64_t seq_stride,
                          int row0, void* stream);
int alpro_take_rows_fwd(const float* src, int R, int s0, int n, int L, int h, float* out32, void* out16, int fmt,
                        void* stream);
int alpro_take_rows_bwd(const float* dout, int R, int s0, int n, int L, int h, float* dsrc, void* stream);
/* out16 = dy * dact, dact = the activation derivative saved by an ALPRO_ACT_GELU epilogue (backward of
 * BertPredictionHeadTransform's activation, xbert.py:659-661) */
int alpro_gelu_grad_mul(const float* dy, const void* dact, int dact_fmt, void* out, int out_fmt, int64_t n, void* stream);
/* Prompter._compute_soft_labels (alpro_models.py:525-529): soft = softmax(sim); ignore = (argmax index < 0.2) */
int alpro_pseudo_labels(const float* sim, int R, int C, float* soft, uint8_t* ignore, void* stream);
/* hard-negative sampling weights: softmax of the local sim block with -inf diagonal (alpro_models.py:288-299) */
int alpro_neg_weights(const float* sim, int64_t ld, int col0, int b, float* w, void* stream);
/* the weights above AND the draw `torch.multinomial(weights[r], 1)` of every row r (alpro_models.py:301-316, 833-844) in
 * one launch: idx[r] (int64, device) ~ Categorical(w[r, :]) by inverse CDF on a Philox4x32-10 uniform with
 * counter (r, draw, 0, 0) and key (seed_lo, seed_hi). w may be null. No host synchronisation. */
int alpro_neg_sample(const float* sim, int64_t ld, int col0, int b, uint32_t seed_lo, uint32_t seed_hi, uint32_t draw,
                     float* w, int64_t* idx, void* stream);
/* known-answer access to the generator: n x (4 counter words, 2 key words) -> n x 4 output words */
int alpro_philox4x32_10(const uint32_t* ctr_key, uint32_t* out, int n, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused optimizer step over flat buffers (alpro_b200/csrc/optim.cu)
 */
/* *out += sum_i x[i]^2   (global gradient norm for clip_grad_norm_, run_video_retrieval.py:473-476) */
int alpro_sumsq(const float* x, int64_t n, float* out, void* stream);
/* AdamW of src/optimization/adamw.py:40-103 on flat p/g/m/v (n % 4 == 0); g is scaled by min(1, max_norm/(sqrt(*gnorm_sq)+1e-6))
 * when max_norm > 0; step_size = lr*sqrt(1-b2^t)/(1-b1^t) (or lr without bias correction); lr_wd = lr*weight_decay */
int alpro_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2, float eps,
                     float step_size, float lr_wd, const float* gnorm_sq, float max_norm, void* stream);
/* Device-side step bookkeeping for alpro_adamw_step: when *gnorm_sq is finite, ++*step_count (the number of updates
 * actually APPLIED; an fp16-overflow step is skipped and not counted) and *step_size_out = lr*sqrt(1-b2^t)/(1-b1^t)
 * (lr when correct_bias == 0); otherwise *step_size_out = 0. Pass step_size_out as `step_size_dev` below. */
int alpro_adamw_prepare(const float* gnorm_sq, float lr, float beta1, float beta2, int correct_bias, int* step_count,
                        float* step_size_out, void* stream);
/* alpro_adamw_step with the step size read from device memory (written by alpro_adamw_prepare) */
int alpro_adamw_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2, float eps,
                         const float* step_size_dev, float lr_wd, const float* gnorm_sq, float max_norm, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Collectives of the data-parallel path over NCCL / NVLink (alpro_b200/csrc/comm.cu). One communicator per process and
 * GPU (the caller's CURRENT device at alpro_comm_init). NCCL is bound at run time (dlopen libnccl.so.2);
 * ALPRO_ENOTSUP when it cannot be loaded. dtype: 0 = f32, 1 = f16, 2 = bf16. All calls are stream-ordered.
 * Replaces hvd.allgather (alpro_models.py:110-111, 764-765), its gradient (sum over ranks, local slice) and the
 * gradient averaging of hvd.DistributedOptimizer (run_video_retrieval.py:320-323, 444).
 */
/* rank 0 creates the 128-byte id and ships it to the other ranks by any out-of-band means */
int alpro_comm_unique_id(void* id128);
int alpro_comm_init(void** comm_out, int world, int rank, const void* id128);
int alpro_comm_destroy(void* comm);
int alpro_comm_rank(void* comm);
int alpro_comm_world(void* comm);
/* recv[world * count_per_rank] = concatenation of every rank's send[count_per_rank] in rank order */
int alpro_comm_allgather(void* comm, const void* send, void* recv, int64_t count_per_rank, int dtype, void* stream);
/* recv[count_per_rank] = slice `rank` of the element-wise SUM over ranks of send[world * count_per_rank] */
int alpro_comm_reduce_scatter(void* comm, const void* send, void* recv, int64_t count_per_rank, int dtype, void* stream);
/* buf[count] = sum (average != 0: mean) over ranks, in place */
int alpro_comm_allreduce(void* comm, void* buf, int64_t count, int dtype, int average, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Gradient averaging as a thin kernel over NVLink peer memory (alpro_b200/csrc/allreduce.cu): 128-thread CTAs without
 * shared memory that co-reside with the backward GEMMs. buf[offset, offset+count) (floats, offset % 4 == 0) is replaced
 * on EVERY rank by scale * (sum over ranks); this call reduces and broadcasts the 1/world slice owned by `rank`.
 * peer_ptrs: host array of `world` (<= 16) peer-mapped device pointers to each rank's buffer base; mc_ptr: NVSwitch
 * multicast mapping of the same buffer (in-switch reduction, multimem.ld_reduce / multimem.st) or null (P2P loads and
 * stores). The caller brackets the launch with a cross-rank barrier on the same stream.
 */
int alpro_nvl_allreduce(const void* const* peer_ptrs, void* mc_ptr, int world, int rank, int64_t offset, int64_t count,
                        float scale, int num_ctas, void* stream);
/* copy-engine variant of the same reduction (comm.CeGradReducer): peers' slices are pulled into `stage`
 * ([nparts] x part_stride floats) with alpro_memcpy_async, then own[0,count) = (own + sum_k stage[k]) * scale */
int alpro_sum_slices(float* own, const float* stage, int nparts, int64_t count, int64_t part_stride, float scale,
                     void* stream);
int alpro_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALPRO_B200_H */
