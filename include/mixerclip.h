/* libmixerclip -- C ABI of the B200-native Mixer-CLIP training hot path.
 *
 * The reference (corentin-ryr/CLIP-mixer) is pure Python/PyTorch and has no FFI; this is the
 * boundary a maintainer would bind (ctypes stub in INTEGRATION.md).  Every entry point replaces a
 * group of library-call sites of the reference, cited per function as training/...:line.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers owned by the caller
 *     (PyTorch's caching allocator); the library never allocates or frees device memory;
 *   - every function launches asynchronously on `stream` (a cudaStream_t passed as void*) and
 *     never synchronises, so the calls are CUDA-graph capturable;
 *   - return 0 on success, a negative MC_ERR_* otherwise; mc_last_error() gives the message of
 *     the last failure on the calling thread.  No exception crosses the boundary;
 *   - "act" tensors are the GEMM operand copies: bf16 for the tensor-core engine, fp32 for the
 *     SIMT validation engine.  The residual stream, LayerNorm statistics, gradients of
 *     parameters, features and the contrastive head are always fp32.
 */
#ifndef MIXERCLIP_H_
#define MIXERCLIP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC_OK 0
#define MC_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define MC_ERR_CUDA (-2)    /* a CUDA runtime / driver call failed */

#define MC_F32 0
#define MC_BF16 1

#define MC_MAJOR_K 0  /* the K (reduction) index is contiguous in memory */
#define MC_MAJOR_MN 1 /* the M (for A) / N (for B) index is contiguous in memory */

#define MC_BIAS_NONE 0
#define MC_BIAS_N 1 /* bias[n] added to every row    (nn.Linear on rows)         */
#define MC_BIAS_M 2 /* bias[m] added to every column (token-mixing orientation)  */

#define MC_ACT_NONE 0
#define MC_ACT_GELU 1     /* QuickGELU, model.py:175-177                         */
#define MC_ACT_GELU_BWD 2 /* multiply by QuickGELU'(zin[m,n])                     */

int mc_version(void);
const char* mc_last_error(void);
/* Persistent kernels size their grids for min(sms, SM count) SMs (0 = all; also the MC_SM_LIMIT environment variable).
 * Data-parallel runs leave a few SMs to the NCCL all-reduce that overlaps the backward pass (training.py:93,170). */
int mc_set_sm_limit(int sms);
/* sm count / compute capability of the current device */
int mc_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * Batched GEMM with fused epilogue:  for every batch b
 *     acc[m,n] = sum_k A_b[m,k] * B_b[n,k]
 *     x = acc (+ bias)            ; optional store of x to zout (pre-activation, act dtype)
 *     x = act(x)                  ; GELU, or x * GELU'(zin[m,n])
 *     x = x + R_b[m,n]            ; optional fp32 residual
 *     C_b[m',n] (=, += or atomic+=) x        with m' = m (+ m / row_remap + 1 if row_remap > 0)
 *     rowsum_out[m] += sum_{b,n} x           (optional)
 * Replaces the cuBLAS/ATen call sites k1, k4-k11, k16 of SURVEY.md 2.4:
 *   nn.Linear lin1..lin4 + QuickGELU + residual   training/clip/model.py:206-222
 *   patch convolution as an im2col GEMM           training/clip/model.py:258,272
 *   x @ proj, x @ text_projection                 training/clip/model.py:288,424
 *   and the dgrad / wgrad GEMMs autograd derives from them (training/training.py:170).
 *
 * Operand A is logically [batch][M][K], B is [batch][N][K]; `*_major` says which logical index is
 * contiguous, `ld*` is the element stride of the other index, `*_batch_stride` the element stride
 * between batches (0 = shared).  With k_spans_batch != 0 the reduction also runs over the batch
 * index (weight gradients of the token-mixing MLP: K = batch x D) and C has no batch dimension.
 * ------------------------------------------------------------------------------------------ */
typedef struct mc_gemm_params {
    int64_t M, N, K, batch;
    const void* A;
    int32_t a_major;
    int64_t lda, a_batch_stride;
    const void* B;
    int32_t b_major;
    int64_t ldb, b_batch_stride;
    int32_t k_spans_batch;
    /* output */
    void* C;
    int32_t c_dtype; /* MC_F32 or MC_BF16 (act dtype engines only write their own act dtype or fp32) */
    int64_t ldc, c_batch_stride;
    int32_t accumulate; /* C += x (fp32 C only) */
    int32_t split_k;    /* >1: K is split over CTAs and C is updated with atomic adds (fp32 C, implies +=) */
    int32_t row_remap;  /* >0: destination row = m + m / row_remap + 1 (patch rows behind a class token) */
    /* epilogue */
    const float* bias;
    int32_t bias_mode;
    void* zout; /* pre-activation, same shape as C: fp16 for the tensor-core engine, fp32 for SIMT */
    int64_t ldz, z_batch_stride;
    const void* zin; /* same format as zout */
    int64_t ldzin, zin_batch_stride;
    int32_t act;
    const float* R;
    int64_t ldr, r_batch_stride;
    /* optional fused reduction: rowsum_out[m] += sum_{batch, n} of the value written to C (before rounding);
     * the bias gradient of token-mixing lin1 rides on the dZ1 GEMM this way (model.py:207) */
    float* rowsum_out;
    /* != 0: C and R are stored TRANSPOSED: element (m, n) of batch b lives at b*batch_stride + n*ld + m (fp32 C only).
     * Lets a GEMM whose result is a [P x D] slab run with D as its M dimension (full 128-row tiles, coalesced
     * 128-byte stores).  Measured slower than the [P x D]-row orientation for the token-mixing shapes in round 1
     * (profiles/r1g), so the engine does not use it by default; kept as an option of the ABI. */
    int32_t c_transposed;
    /* Optional second operand pair with the same M, N, K, batch (tensor-core engine, act == MC_ACT_GELU_BWD and
     * zin == NULL): acc2[m,n] = sum_k A2_b[m,k] * B2_b[n,k] is computed next to acc and the epilogue uses
     * z = acc2 + bias2[m] instead of a saved pre-activation:  C = acc * QuickGELU'(acc2 + bias2[m]).
     * Token-mixing backward recomputes Z1 = W1 U + b1 this way (K = P is tiny), so the forward pass never stores
     * Z1 and the backward pass never reads it ("backward must recompute Z1/H1"). */
    const void* A2;
    int32_t a2_major;
    int64_t lda2, a2_batch_stride;
    const void* B2;
    int32_t b2_major;
    int64_t ldb2, b2_batch_stride;
    const float* bias2;
    /* optional, residual epilogue of the tensor-core engine only (act NONE, fp32 C, R set, batch 1):
     * rowstat_out[2*m] += sum_n x, rowstat_out[2*m + 1] += sum_n x*x over the values written to C.  Gives the LayerNorm that
     * follows a channel-mixing lin4 (model.py:216 on the output of :217) its row statistics from the producing GEMM, so
     * that the fused token-mixing kernel can normalise in its prologue (mc_token_mix_params.ln_sums). */
    float* rowstat_out;
} mc_gemm_params;

/* tcgen05 / TMEM / TMA engine: bf16 operands, fp32 accumulation in tensor memory; zout / zin fp16. */
int mc_gemm_bf16_tc(const mc_gemm_params* p, void* stream);
/* SIMT FFMA engine: fp32 operands (the 1e-5 validation precision of north_star). */
int mc_gemm_f32_simt(const mc_gemm_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused token-mixing MLP of one Mixer block (tcgen05 engine only), training/clip/model.py:206-208,216,220-222:
 *     y = x + ( lin2( QuickGELU( lin1( u.permute(0,2,1) ) ) ) ).permute(0,2,1),      u = LN1(x)  (bf16 [B,P,D])
 * One persistent kernel per call; the [B, D, 4P] hidden activation never reaches HBM and no transposed copy of
 * the activation is made (the permute lives in the UMMA descriptor).  Per sample, with W1 [4P x P], W2 [P x 4P]:
 *   mc_token_mix_fwd    y  = x + W2 g(W1 u + b1) + b2                                  (x, y fp32 [B,P,D], x != y)
 *   mc_token_mix_dgrad  y  = W1^T ( (W2^T dy) * g'(W1 u + b1) )                         (dy bf16 [B,P,D]; y = dU fp32)
 *   mc_token_mix_wgrad  gw2 += sum_b dy H1^T, gw1 += sum_b dZ1 u^T, gb1 += rowsum(dZ1)  (H1, dZ1 recomputed on chip)
 * w1 / w2 are the bf16 operand copies with row pitches ld1 / ld2 (multiples of 8, pad elements ignored).
 * Supported shapes: mc_token_mix_supported(P, D) != 0  (P <= 80 tokens, D a multiple of 128); other shapes
 * (B/16: 197 tokens) run the same math as separate mc_gemm_bf16_tc calls.
 * ------------------------------------------------------------------------------------------ */
typedef struct mc_token_mix_params {
    int64_t B, P, D;
    const void* u;
    const void* w1;
    int64_t ld1;
    /* W1^T as its own bf16 matrix [P x ld1t] (ld1t >= 4P, multiple of 8; mc_transpose_bf16 of w1): both resident weight
     * tiles are then fetched by TMA.  Required. */
    const void* w1t;
    int64_t ld1t;
    const float* b1;
    const void* w2;
    int64_t ld2;
    const float* b2;
    const float* x;
    float* y;
    const void* dy;
    float* gw1;
    int64_t ldg1;
    float* gw2;
    int64_t ldg2;
    float* gb1;
    /* mc_token_mix_fwd only, optional: LayerNorm in the prologue (model.py:216 folded into :220-222).  With ln_sums != NULL
     * the kernel ignores `u`: it reads the fp32 block input x, normalises it with the per-row (sum, sum of squares) over D
     * that the producing GEMM left in ln_sums [B*P][2] (mc_gemm_params.rowstat_out) and ln_gamma / ln_beta [D], feeds the
     * bf16 result to its tensor-core GEMMs and also writes it to u_out [B,P,D] (the backward kernels read it) together
     * with the row statistics ln_mean / ln_rstd [B*P] (mc_ln_bwd reads them).  x, ln_gamma, ln_beta, u_out: 16-byte
     * aligned; ln_sums: 8-byte aligned.  Variance = E[x^2] - mean^2 in fp32, clamped at 0, eps 1e-5. */
    const float* ln_sums;
    const float* ln_gamma;
    const float* ln_beta;
    void* u_out;
    float* ln_mean;
    float* ln_rstd;
} mc_token_mix_params;
int mc_token_mix_supported(int64_t P, int64_t D);
int mc_token_mix_fwd(const mc_token_mix_params* p, void* stream);
int mc_token_mix_dgrad(const mc_token_mix_params* p, void* stream);
int mc_token_mix_wgrad(const mc_token_mix_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm (always fp32 math), training/clip/model.py:166-172; instances ln_pre :263,
 * layerNorm1/2 :205,210, ln_post :268, ln_final :344.
 *   row r reads x + (row_index ? row_index[r] : r) * x_row_stride; rows with
 *   r % cls_period == 0 read `cls` instead when cls != NULL (class token of model.py:275-277).
 *   Writes y (y_dtype, dense [rows, D], or strided by y_row_stride) and mean / rstd [rows].
 * ------------------------------------------------------------------------------------------ */
int mc_ln_fwd(const float* x, int64_t x_row_stride, const int32_t* row_index, const float* cls, int64_t cls_period,
              const float* gamma, const float* beta, void* y, int32_t y_dtype, int64_t y_row_stride, float* mean,
              float* rstd, int64_t rows, int64_t D, void* stream);

/* LayerNorm backward fused with the residual-gradient add and the reductions that ride on it:
 *   dx[r] = (dres ? dres[r] : 0) + LNbwd(dy[r])           (fp32, plus an optional act copy dx_act)
 *   dgamma += sum_r dy*xhat ; dbeta += sum_r dy           (atomic accumulation into fp32 [D])
 *   colsum_out[d]   += sum_r dx[r,d]                      (bias grad of the preceding lin4, may be NULL)
 *   rowsum_out[r%P] += sum_d dx[r,d]                      (bias grad of token-mix lin2, may be NULL)
 * x rows are addressed like mc_ln_fwd (row_index / x_row_stride); dx rows likewise
 * (dx_row_stride); dy is dense [rows, D] fp32.  Rows with r % cls_period == 0 (cls != NULL) take
 * x from `cls` and accumulate their dx into dcls[D] instead of writing dx.
 * Aliasing contract: dres MAY be the same buffer as dx (in-place residual-gradient add; each element is read and
 * then written by one thread).  No other pair of arguments may overlap. */
int mc_ln_bwd(const float* dy, const float* x, int64_t x_row_stride, const int32_t* row_index, const float* cls,
              int64_t cls_period, const float* mean, const float* rstd, const float* gamma, const float* dres,
              float* dx, int64_t dx_row_stride, void* dx_act, int32_t act_dtype, float* dgamma, float* dbeta,
              float* colsum_out, float* rowsum_out, int64_t rowsum_period, float* dcls, int64_t rows, int64_t D,
              void* stream);

/* Column sums (bias grads of lin3, model.py:212) and per-row-group sums (bias grads of lin1):
 *   colsum:  out[c] += sum_r x[r, c]                      x: [rows, cols] act dtype
 *   rowsum:  out[r % period] += sum_c x[r, c]                                                */
int mc_colsum(const void* x, int32_t dtype, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream);
int mc_rowsum(const void* x, int32_t dtype, int64_t rows, int64_t cols, int64_t ld, int64_t period, float* out,
              void* stream);

/* fp32 -> act copy (dst dense), optional pad of the leading dimension (token-mix weights). */
int mc_cast_pad(const float* src, int64_t rows, int64_t cols, int64_t src_ld, void* dst, int32_t dst_dtype,
                int64_t dst_ld, void* stream);

/* dst[b][c][r] = src[b][r][c] for `batch` small bf16 matrices [rows x cols] (pitch ld_src) -> [cols x ld_dst], pad columns
 * r >= rows written as zero: W1^T copies of the token-mixing lin1 weights, refreshed once per step. */
int mc_transpose_bf16(const void* src, int64_t rows, int64_t cols, int64_t ld_src, int64_t src_batch_stride, void* dst,
                      int64_t ld_dst, int64_t dst_batch_stride, int64_t batch, void* stream);

/* im2col of the stride==kernel patch convolution, model.py:258,272:
 *   image [B,3,R,R] (fp32, or uint8 with the /255 + Normalize of training.py:115,149 fused) ->
 *   act [B*g*g, 3*p*p], row (b,gy,gx), column (c,py,px). */
int mc_im2col(const void* image, int32_t image_is_u8, int64_t B, int64_t R, int64_t patch, void* out,
              int32_t out_dtype, void* stream);

/* Token embedding, model.py:414: x[b,t,:] = table[text[b,t],:]; and its gradient (run-length
 * pre-reduced scatter-add: padding tokens repeat). text is int64 [B, C]. */
int mc_embed_fwd(const int64_t* text, const float* table, float* x, int64_t B, int64_t C, int64_t W, int64_t vocab,
                 void* stream);
int mc_embed_bwd(const int64_t* text, const float* dx, float* dtable, int64_t B, int64_t C, int64_t W, int64_t vocab,
                 void* stream);
/* eot_row[b] = b*C + argmax_t text[b,t]  (first maximum, like torch.argmax; model.py:424) */
int mc_eot_rows(const int64_t* text, int32_t* eot_row, int64_t B, int64_t C, void* stream);

/* L2 normalisation of the features, model.py:433-434, and its backward
 *   u = f / ||f||;   df = (du - u (u.du)) / ||f|| */
int mc_l2norm_fwd(const float* f, float* u, float* inv_norm, int64_t rows, int64_t E, void* stream);
int mc_l2norm_bwd(const float* du, const float* u, const float* inv_norm, float* df, void* df_act, int32_t act_dtype,
                  int64_t rows, int64_t E, void* stream);

/* ------------------------------------------------------------------------------------------
 * Contrastive head, training/training.py:158-168 (+ its backward through autograd, :170), with
 * the gathered features detached: two row-softmax cross-entropies over [n x N] logits that are
 * never written to memory (online softmax over column tiles).
 *   ui, ut     [n, E]  local normalised features;  ui_all, ut_all [N, E] gathered (rank order)
 *   log_scale  device scalar t (logit_scale parameter); labels g_i = rank*n + i
 * Outputs (fp32): loss[1] (+=, pre-zeroed by the caller), dui, dut [n, E] (overwritten),
 *   dlog_scale[1] (+=), all already divided for the mean over n and the /2.
 * workspace: mc_head_workspace_bytes(n, N, E) bytes, 256-byte aligned.
 * Two implementations behind the same entry point (MC_HEAD_TC = auto | 0 | 1): fp32 FFMA kernels (online softmax over
 * 32-column tiles) for per-GPU batches, and from 2^24 logits per direction a tensor-core path: both contractions as
 * bf16 x 3 split GEMMs on the tcgen05 engine (fp32-class accuracy), online softmax per 4096-column slab - one
 * [n x 4096] fp32 slab of logits exists at a time, never the [n x N] matrix (n = 4096, N = 32768: 2.6 ms vs 57.6 ms).
 * ------------------------------------------------------------------------------------------ */
int64_t mc_head_workspace_bytes(int64_t n, int64_t N, int64_t E);
int mc_head_fwd_bwd(const float* ui, const float* ut, const float* ui_all, const float* ut_all, const float* log_scale,
                    int64_t n, int64_t N, int64_t E, int64_t rank, float grad_scale, float* loss, float* dui,
                    float* dut, float* dlog_scale, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer step of training/training.py:73-82,181,185 over flat fp32 buffers:
 *   mc_sumsq: out[0] += sum g^2           (global grad norm for clip_grad_norm_(.., 20))
 *   mc_adamw: g' = g * grad_mul * min(1, max_norm / (sqrt(sumsq[0]) * grad_mul + 1e-6));
 *             AdamW (decoupled decay on the 64-element chunks whose decay_flags byte is 1);
 *             optional bf16 mirror of the new weights.  The per-step scalars come from DEVICE
 *             memory so a captured CUDA graph can be replayed while the schedule advances:
 *             hyper = {lr, 1 - beta1^t, 1 - beta2^t}.  sumsq may be NULL (no clipping).
 *   mc_sched_step: the scheduler / step-count side of the same lines, ON THE DEVICE (so that a captured graph
 *             replays with no host-written scalar that a later step could overwrite): state = int64 {t, s}
 *             (Adam step count, scheduler step; training.py:185-186); writes hyper = {lr(s), 1 - beta1^(t+1),
 *             1 - beta2^(t+1)} and stores {t+1, s+1}.  lr(s) = CosineAnnealingWarmupRestarts(first_cycle_steps,
 *             max_lr, min_lr, warmup_steps, cycle_mult 1, gamma 1) of training.py:83-89, or fixed_lr when >= 0.
 *   mc_sumsq is deterministic (per-block partials reduced in fixed order by the last block), so data-parallel
 *             replicas holding identical all-reduced gradients clip with bit-identical coefficients.
 * ------------------------------------------------------------------------------------------ */
int mc_sumsq(const float* g, int64_t n, float* out, void* stream);
int mc_sched_step(int64_t* state, float* hyper, int64_t first_cycle_steps, double max_lr, double min_lr,
                  int64_t warmup_steps, double beta1, double beta2, double fixed_lr, void* stream);
int mc_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, const uint8_t* decay_flags, int64_t n,
             const float* sumsq, const float* hyper, float grad_mul, float max_norm, float beta1, float beta2,
             float eps, float weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIXERCLIP_H_ */
