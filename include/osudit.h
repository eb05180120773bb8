/* osudit — C ABI of the B200-native DiT denoising hot path (libosudit.so).
 *
 * The reference (OliBomby/osu-diffusion) is pure PyTorch and has no FFI of its own,
 * so each entry point below cites the reference Python it replaces; the host-side binding a
 * maintainer adds is the ctypes stub in INTEGRATION.md (shipped as osu-diffusion_b200/osudit/_lib.py).
 *
 * Conventions (all entry points):
 *   - return 0 on success, <0 on error; the message is in osudit_last_error() (thread-local).
 *   - every pointer is a borrowed DEVICE pointer (the caller's allocator owns it and keeps it alive
 *     until the work enqueued on `stream` has run); shapes/strides are explicit; `stream` is a
 *     cudaStream_t passed as void*.
 *   - no allocation, no host synchronisation, no mutable global state except a mutex-guarded cache
 *     of TMA descriptors: calls are re-entrant and CUDA-graph capturable.
 *   - bf16 tensors are row-major uint16 storage; "split-bf16" means a (hi, lo) pair with
 *     hi = bf16(v), lo = bf16(v - hi).
 */
#ifndef OSUDIT_H_
#define OSUDIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSUDIT_VERSION 2

int osudit_version(void); /* library ABI version (no reference counterpart: the reference has no FFI) */

/* Caps the grid of every persistent kernel (GEMMs, windowed attention) at `n` CTAs instead of one per SM (0 = no cap;
 * rounded down to an even number for the CTA-pair GEMM; n = -1 only queries); returns the previous cap.  Data-parallel training sets it so
 * that NCCL's all-reduce kernels (train.py:152,257) find free SMs next to the backward instead of queueing behind a
 * full-machine persistent grid.  Process-wide; takes effect at the next launch (captured graphs keep their grids). */
int osudit_set_sm_limit(int n);
const char* osudit_last_error(void);

/* Epilogues of osudit_gemm_bf16. */
#define OSUDIT_EPI_F32 0       /* out fp32  = acc + bias                                  */
#define OSUDIT_EPI_BF16 1      /* out bf16  = acc + bias                                  */
#define OSUDIT_EPI_BF16_GELU 2 /* out bf16  = gelu_tanh(acc + bias)   (models.py:138)     */

/* out[M,N] = sum_{s<nseg} A_s[M,K_s] . B_s[N,K_s]^T + bias, tcgen05/TMEM/TMA GEMM.
 * Replaces every nn.Linear on the path: models.py:233-234 (first layer), :35-38 (t-MLP),
 * :152-159,:193 (adaLN), :164-170 (packed QKV in_proj, out_proj), :112-119 (fc1+GELU, fc2).
 * A_s, B_s bf16 row-major with leading dimensions lda/ldb (elements, multiple of 8); K_s % 8 == 0;
 * N % 8 == 0; bias fp32[N] or NULL; out leading dimension ldo (elements). */
int osudit_gemm_bf16(int nseg, const void* const* a, const int64_t* lda, const void* const* b,
                     const int64_t* ldb, const int64_t* k, int64_t M, int64_t N, const float* bias,
                     int epilogue, void* out, int64_t ldo, void* stream);

/* Banded / full / generically masked multi-head self-attention on packed QKV.
 * Replaces nn.MultiheadAttention's core (models.py:164-170) with the band mask of sample.py:81-84:
 * query j attends key i iff -w_left <= i - j <= w_right (w_left = W-1, w_right = W); pass -1/-1 for
 * no mask.  `mask` (T*T bytes, non-zero = blocked) is an optional generic mask applied on top.
 * qkv bf16 [B*T, 3*H*head_dim], out bf16 [B*T, H*head_dim].
 * algo: AUTO picks the tcgen05 window kernel (128 queries x 384-key window, S and O in TMEM) when
 * the allowed keys of every 128-query tile fit its window (band within +-128, or T <= 256; no
 * generic mask), else the mma.sync flash kernel; the other two values force one (tests). */
#define OSUDIT_ATTN_AUTO 0
#define OSUDIT_ATTN_MMA_SYNC 1
#define OSUDIT_ATTN_TCGEN05 2
#define OSUDIT_ATTN_FA 3
#define OSUDIT_ATTN_STREAM 4
/* STREAM: the double-buffered streaming tcgen05 kernel (attn_stream.cu): 128-key slabs, scores two slabs ahead of
 * the softmax in three TMEM buffers, probabilities written back to TMEM (PV reads its A operand there), epilogue
 * warps with a TMA store; head_dim 64, any T, band or full, optional lse.  AUTO prefers it where it applies.
 * FA: the two-slot streaming kernel (attn_fa.cu), same coverage (round 2a; kept selectable).
 * lse (optional, fp32 [B, H, T]): log2-domain log-sum-exp of every row, saved for the backward. */
int osudit_attn_band(const void* qkv, void* out, int B, int T, int H, int head_dim, int w_left,
                     int w_right, const uint8_t* mask, int algo, float* lse, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward pass (training, train.py:249-257: autograd through DiT.forward under the loss of
 * gaussian_diffusion.py:785-874).  Data- and weight-gradient GEMMs reuse osudit_gemm_bf16 on
 * operands transposed by osudit_transpose_bf16.
 * ---------------------------------------------------------------------------------------------- */

/* Weight gradient of y = x W^T — autograd of every nn.Linear on the path (models.py:35-38,112-119,152-159,164-170,193,
 * 233-234) under loss.backward() (train.py:257): out[M,N] (fp32, ACCUMULATED: zero it first) += dy[rows,M]^T . x[rows,N],
 * both operands read token-major as MN-major tcgen05 operands; the token dimension is split across
 * CTAs and combined with TMA reduce-add.  ld_* in elements (multiples of 8). */
int osudit_gemm_wgrad(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, int64_t rows, int64_t M,
                      int64_t N, float* out, int64_t ldo, void* stream);

/* Autograd of nn.MultiheadAttention under the band / generic mask (models.py:130-135,164-170; train.py:257):
 * dqkv (bf16 [B*T, 3*H*hd]) from dout (bf16 [B*T, H*hd]), the forward's qkv / out / lse.
 * delta is fp32 [B, H, T] scratch.  Band semantics as in osudit_attn_band; head_dim 64 or 72.
 * dbias_qkv (fp32 [3*H*hd], ACCUMULATED, may be NULL) += column sums of dqkv: the in_proj_bias gradient. */
int osudit_attn_band_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                         float* delta, void* dqkv, int B, int T, int H, int head_dim, int w_left,
                         int w_right, float* dbias_qkv, void* stream);

/* Operand re-layout for the data-gradient GEMMs (dx = dy W, i.e. autograd of F.linear, models.py:112-119,164-170):
 * out[cols, out_ld] (bf16, out_ld >= rows: pad the GEMM K dimension to a multiple of 8) = in[rows, cols]^T;
 * in is bf16, or fp32 when in_is_f32. */
int osudit_transpose_bf16(const void* in, void* out, int64_t rows, int64_t cols, int64_t out_ld, int in_is_f32,
                          void* stream);

/* Every GEMM-ready copy of every fp32 weight matrix in one launch: what the reference gets from autocast's per-call
 * casts (train.py:249-255) and what the native backward needs transposed.  `segments` is a DEVICE array of `nseg`
 * records, sorted by tile0:
 *   struct { const float* src; void* copy; void* trans; void* hi; void* lo; int64_t ld_trans;
 *            int32_t rows, cols, tile0, tiles_x; }            (64 bytes)
 * src fp32 [rows, cols]; copy bf16 [rows, cols]; trans bf16 [cols, ld_trans] with trans[c][r] = src[r][c]; hi / lo
 * split-bf16 [rows, cols]; any destination may be NULL.  A segment owns tiles [tile0, tile0 + tiles_x * ceil(rows/64))
 * of 64 x 64 elements, tiles_x = ceil(cols/64); total_tiles is the launch grid. */
int osudit_repack_weights(const void* segments, int nseg, int total_tiles, void* stream);

/* nn.GELU(approximate="tanh") of the Mlp and its derivative (models.py:112-119):
 * backward == 0: out = gelu_tanh(pre);  backward == 1: out = dy * gelu_tanh'(pre).  bf16, n % 8 == 0. */
int osudit_gelu(const void* pre, const void* dy, void* out, int64_t n, int backward, void* stream);

/* The gated residual update of a DiT block done by the GEMM that produces the branch (inference):
 *   x[M, N] (fp32) += gate[row / rows_per_batch, :] * (A[M, K] B[N, K]^T + bias)
 * Replaces `x = x + gate_msa.unsqueeze(1) * self.attn(...)` and `x = x + gate_mlp.unsqueeze(1) * self.mlp(...)`
 * (models.py:164-174) for the out-projection / fc2 Linear: the epilogue multiplies by the gate and adds into the
 * residual stream with an fp32 TMA reduce-add, so the branch is never written and the LayerNorm kernel that follows
 * only reads x (6 D instead of 12 D bytes per token).  gate: fp32, row b at gate + b * gate_ld (elements).
 * CTA-pair kernel only: osudit_gemm_gated_residual_applicable(M, N, rows_per_batch) != 0 says whether the shape is
 * taken (N % 256 == 0 or N % 192 == 0, rows_per_batch % 128 == 0 and dividing M, >= 37 output tiles); otherwise
 * use osudit_gemm_bf16 + osudit_ln_modulate with a branch. */
int osudit_gemm_gated_residual_applicable(int64_t M, int64_t N, int64_t rows_per_batch);
int osudit_gemm_gated_residual(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t K, int64_t M, int64_t N,
                               const float* bias, const float* gate, int64_t gate_ld, int64_t rows_per_batch, float* x,
                               int64_t ldx, void* stream);

/* Single-segment bf16 GEMM (as osudit_gemm_bf16) whose epilogue also touches a second bf16 [M, N] tensor `aux`:
 *   OSUDIT_EPI_BF16_GELU_SAVE: out = gelu_tanh(A B^T + bias), aux = gelu_tanh'(A B^T + bias): all the backward needs
 *                              from the fc1 pre-activation                     (training forward, models.py:112-119)
 *   OSUDIT_EPI_BF16_DGELU:     out = (A B^T) * aux                             (the gradient through that GELU)
 * Fused into the CTA-pair kernel's epilogue where that kernel applies; two launches with the same result otherwise. */
#define OSUDIT_EPI_BF16_GELU_SAVE 3
#define OSUDIT_EPI_BF16_DGELU 4
int osudit_gemm_bf16_aux(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t K, int64_t M, int64_t N,
                         const float* bias, int epilogue, void* out, int64_t ldo, void* aux, int64_t ld_aux,
                         void* stream);

/* Backward through the Mlp's GELU (models.py:112-119), for shapes the fused DGELU epilogue does not take:
 * out[rows, N] = dy * gelu_tanh'(pre) and dbias[N] (fp32, ACCUMULATED, may be NULL) += column sums of out:
 * the fc1 pre-activation gradient together with the fc1 bias gradient.  bf16, N % 8 == 0. */
int osudit_gelu_bwd(const void* pre, const void* dy, void* out, int64_t rows, int N, float* dbias, void* stream);

/* out[N] (fp32, ACCUMULATED: caller zeroes) += column sums of in[rows, N] (bf16 or fp32): the bias gradients of the
 * nn.Linear layers (models.py:35-38,112-119,152-159,164-170,193,233-234). */
int osudit_colsum(const void* in, int in_is_f32, int64_t rows, int N, float* out, void* stream);

/* Backward of x_out = x + gate[b] * y (models.py:161-163,172):
 * dy (bf16) = gate[b] * dx;  dgate[b] (fp32, accumulated) += sum_t dx * y;
 * dbias[D] (fp32, ACCUMULATED, may be NULL) += column sums of dy: the bias gradient of the Linear that made y. */
int osudit_gate_residual_bwd(const float* dx, const void* y, const float* gate, float* dgate,
                             int64_t mod_ld, int B, int T, int D, void* dy, float* dbias, void* stream);

/* Backward of h = LayerNorm(x) * (1 + scale[b]) + shift[b] (models.py:12-13,160,173):
 * dshift[b], dscale[b] accumulated (fp32); dx written (accumulate == 0) or added to (== 1). */
int osudit_ln_modulate_bwd(const float* x, const void* dh, const float* scale, float* dshift,
                           float* dscale, int64_t mod_ld, int B, int T, int D, float* dx,
                           int accumulate, void* stream);

/* osudit_ln_modulate_bwd followed by osudit_gate_residual_bwd on the freshly updated dx, in one pass over the
 * residual-stream gradient (the order the two occur in the backward of models.py:160-163,172-174):
 * dx += LN-modulate-backward(x, dh);  then, when y != NULL (the output of the branch that was added to the
 * residual right before this LayerNorm): dy (bf16) = gate[b] * dx, dgate[b] += sum_t dx * y,
 * dbias[D] (may be NULL) += column sums of dy. */
int osudit_ln_gate_bwd(const float* x, const void* dh, const float* scale, float* dshift, float* dscale,
                       int64_t mod_ld, int B, int T, int D, float* dx, int accumulate, const void* y,
                       const float* gate, float* dgate, void* dy, float* dbias, void* stream);

/* Backward of FinalLayer (models.py:192-196) given dout fp32 [B,4,T] and x = the layer's input
 * (last gated residual already applied): dw [4,D], dbias [4], dshift/dscale accumulated; dx written. */
int osudit_final_layer_bwd(const float* x, const float* dout, const float* shift, const float* scale,
                           float* dshift, float* dscale, int64_t mod_ld, int B, int T, int D,
                           const float* w, float* dw, float* dbias, float* dx, void* stream);

/* dcond[r] = ds[r] * SiLU'(a[r] + table[y[r]]);  dtable[y[r]] += dcond[r] when dtable != NULL
 * (dense label-embedding gradient, models.py:73). */
int osudit_silu_bwd(const float* a, const float* table, const int64_t* y, const float* ds, int64_t rows,
                    int D, float* dcond, float* dtable, void* stream);

/* x (fp32 [rows, D], in place) += gate[b] * branch (bf16 [rows, D]) when branch != NULL, then
 * h (bf16 [rows, D]) = LayerNorm(x) * (1 + scale[b]) + shift[b], b = row / T.
 * gate/shift/scale point at column 0 of their chunk in the adaLN output, `mod_ld` floats per batch
 * row.  Replaces modulate(norm(x), shift, scale) and the gated residual adds
 * (models.py:12-13,160-163,172-174). */
/* x_out (optional): write the updated residual there instead of in place (training keeps every
 * LayerNorm input for the backward). */
int osudit_ln_modulate(float* x, const void* branch, const float* gate, const float* shift,
                       const float* scale, int64_t mod_ld, int64_t rows, int T, int D, void* h,
                       float* x_out, void* stream);

/* FinalLayer (models.py:192-196) fused with the last gated residual add and the output transpose
 * (models.py:323-324): out fp32 [B, out_channels, T]; w fp32 [out_channels, D]. */
int osudit_final_layer(float* x, const void* branch, const float* gate, const float* shift,
                       const float* scale, int64_t mod_ld, int64_t rows, int T, int D,
                       const float* w, const float* bias, int out_channels, float* out,
                       void* stream);

/* FirstLayer input (models.py:227-233, positional_embedding.py:29-77): split-bf16
 * [sincos(x*pf_x) | sincos(y*pf_y) | sincos(o/10) | c] of shape [B*T, 384+E].
 * x fp32 [xrows, 2, T] (row b reads x[b % xrows]: forward_with_cfg, models.py:332-333),
 * o fp32 [B, T], c fp32 [B, E, T], freqs64 = exp(-ln(1e4) k/64). */
int osudit_embed_xoc(const float* x, const float* o, const float* c, const float* freqs64,
                     float pf_x, float pf_y, int B, int xrows, int T, int E, void* a_hi, void* a_lo,
                     void* stream);

/* The x columns only (first 256 of the 384 + E): within one sampling loop o and c are the same at every denoising step
 * (sample.py:87-108 builds them once; gaussian_diffusion.py:514-561 only updates x), so after one full osudit_embed_xoc
 * into the same a_hi / a_lo the other 128 + E columns are left untouched. */
int osudit_embed_x(const float* x, const float* freqs64, float pf_x, float pf_y, int B, int xrows, int T, int E,
                   void* a_hi, void* a_lo, void* stream);

/* timestep_embedding(t, 256) (positional_embedding.py:29-49) as split-bf16 [rows, 256]. */
int osudit_timestep_features(const int64_t* t, const float* freqs128, int rows, void* hi, void* lo,
                             void* stream);

/* out[r] = SiLU(a[a_index ? a_index[r] : r] + (table ? table[y[r]] : 0)) as split-bf16 [rows, D]:
 * the SiLU inside the t-MLP (models.py:30) and, with the label table, "SiLU(t_emb + y_emb)" in
 * front of every adaLN Linear (models.py:73,320,148,189). */
int osudit_silu_split(const float* a, const int32_t* a_index, const float* table, const int64_t* y,
                      int64_t rows, int D, void* hi, void* lo, void* stream);

/* Device-side assertion that every class label lies in [0, table_rows): the launch traps (as the reference's CUDA
 * embedding lookup asserts, models.py:73) instead of gathering / scattering outside the table.  Issued in front of
 * every launch that indexes the embedding table with `y`. */
int osudit_check_labels(const int64_t* y, int64_t n, int64_t table_rows, void* stream);

/* fp32 -> split-bf16 (lo may be NULL): packs nn.Linear weights for the GEMM (the reference keeps fp32 parameters and
 * casts per call under autocast, train.py:249-255; sample.py runs them in fp32). */
int osudit_split_bf16(const float* a, int64_t n, void* hi, void* lo, void* stream);

/* One reverse-diffusion update, optionally with the classifier-free-guidance combine.
 * Replaces models.py:338-343 + gaussian_diffusion.py:312-324,341-358,454-466.
 * model_out fp32 [B,4,T]; x, noise, sample, pred_xstart fp32 [B,2,T]; t int64 [B] (respaced index);
 * coef_table fp32 [K,6] = {log beta, posterior_log_variance_clipped, sqrt_recip_alphas_cumprod,
 * sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2}.
 * cfg_half = B/2 to guide (eps rows b and b+B/2 combined, variance channels untouched), 0 for none.
 * phase 0: whole step.  phase 1: only the unclamped x0 prediction into pred_xstart (the host then
 * applies denoised_fn).  phase 2: x0 taken from x0_in, then clamp/mean/sample.
 * sample / mean / log_variance may be NULL (p_mean_variance wants mean+log_variance, no sample). */
int osudit_diffusion_step(const float* model_out, const float* x, const float* noise,
                          const float* x0_in, const int64_t* t, const float* coef_table, int B, int T,
                          int cfg_half, float cfg_scale, int clip_denoised, int phase, float* sample,
                          float* pred_xstart, float* mean, float* log_variance, void* stream);

/* forward_with_cfg's output on its own (models.py:338-343): [B,4,T] -> [B,4,T]. */
int osudit_cfg_combine(const float* model_out, int B, int T, float cfg_scale, float* out,
                       void* stream);

/* q_sample (gaussian_diffusion.py:231-247): out = sqrt_acp[t] x0 + sqrt_1m_acp[t] noise. */
int osudit_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_acp,
                    const float* sqrt_1m_acp, int B, int64_t per_row, float* out, void* stream);

/* Training loss values and gradient (gaussian_diffusion.py:785-874, EPSILON + LEARNED_RANGE):
 * term_main[b] = mean |noise-eps| (use_l1) or (noise-eps)^2, term_vb[b] = VB term in bits (eps
 * detached), dmodel_out [B,4,T] = d(term_main + term_vb)[b] / d model_out[b].  coef_table as in
 * osudit_diffusion_step; t indexes it. */
int osudit_diffusion_loss(const float* model_out, const float* x0, const float* x_t, const float* noise,
                          const int64_t* t, const float* coef_table, int B, int T, int use_l1,
                          float* term_main, float* term_vb, float* dmodel_out, void* stream);

/* out[b, :] = in[b, :] * g[b]: chain rule from `loss = terms["loss"].mean()` (train.py:255-257) through the per-sample
 * terms of training_losses (gaussian_diffusion.py:785-874). */
int osudit_scale_rows(const float* in, const float* g, int B, int64_t per_row, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Beatmap feature builder: data_loading.py:146-151 (calc_distances), :172-187
 * (split_and_process_sequence_no_augment), :195-203 / sample.py:64-65 (relative time), on the device.
 * seq fp32 [B, R, T], rows = x px, y px, time ms, then R-3 one-hot type rows (R = 19 in the reference).
 * o [B, T] = time - time[0] (+ o_shift[b] when o_shift != NULL: the random offset of training windows);
 * c [B, 128 + R - 3, T] = [cos | sin](dist * freqs64) over the distance to the previous object (the first one
 * measured from the playfield centre), then the type rows; x [B, 2, T] = pos / (512, 384), or NULL to skip it.
 * ---------------------------------------------------------------------------------------------- */
int osudit_beatmap_features(const float* seq, int B, int R, int T, const float* freqs64, const float* o_shift,
                            float* x, float* o, float* c, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step: one multi-tensor launch for AdamW + EMA + gradient unscale + skip on
 * non-finite gradients.  Opt-in replacement of `scaler.step(opt)` with torch.optim.AdamW (train.py:154,
 * 258-259) followed by `update_ema(ema, model.module)` (train.py:36-45,261).
 * ---------------------------------------------------------------------------------------------- */
typedef struct OsuditOptSeg {
  float* p;       /* parameter (fp32), updated in place                         */
  const float* g; /* its gradient (fp32), possibly still multiplied by grad_scale */
  float* m;       /* exp_avg                                                    */
  float* v;       /* exp_avg_sq                                                 */
  float* ema;     /* the EMA copy of p, or NULL                                 */
  long long n;    /* elements                                                   */
} OsuditOptSeg;

/* (train.py:154,258-261.)  Elements per chunk: chunk c = (segment index, chunk offset within the segment) covers elements
 * [offset * E, min((offset + 1) * E, n)). */
int osudit_opt_chunk_elems(void);

/* torch.optim.AdamW.step + update_ema (train.py:36-45,154,258-261) in one launch.
 * segs: DEVICE array of OsuditOptSeg; chunks: DEVICE array of nchunks (int32 segment, int32 offset) pairs.
 * step (device fp32 scalar) is incremented first unless *found_inf != 0; then, unless *found_inf != 0, every
 * element gets torch.optim.AdamW's update with g / *grad_scale and ema = ema*ema_decay + p_new*(1-ema_decay).
 * grad_scale / found_inf may be NULL (no scaling / never skip). */
int osudit_adamw_ema_step(const void* segs, const int32_t* chunks, int nchunks, float lr, float beta1,
                          float beta2, float eps, float weight_decay, float ema_decay, float* step,
                          const float* grad_scale, const float* found_inf, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 mode (north star: eps within 1e-5 relative L2 of the fp32 reference).  Activations stay fp32;
 * a GEMM operand is the three-way bf16 split v = hi + mid + lo stored as one row [hi(K)|mid(K)|lo(K)]
 * ("split3", bf16 [rows, 3K]); weights are stored [hi hi hi mid mid lo] ([N, 6K]) so that three
 * K-segments of osudit_gemm_bf16 (widths 3K, 2K, K) accumulate the six significant products.
 * ---------------------------------------------------------------------------------------------- */

/* fp32-mode form of every nn.Linear forward (models.py:35-38,112-119,152-159,164-170,193,233-234):
 * osudit_gemm_bf16 with EPI_F32 and "precise" accumulation: the tensor core adds into its fp32 accumulator
 * with truncation, a bias that grows with the length of the accumulation chain (measured 4e-6 relative at
 * 200 MMAs).  Here the concatenated K range of all segments is cut into chains of kb_per_split 64-wide
 * k-blocks; each chain's partial tile is added into `out` (zeroed by this call) with round-to-nearest fp32
 * reduce-adds (cp.reduce.async.bulk).  bias is added once. */
int osudit_gemm_bf16_splitk(int nseg, const void* const* a, const int64_t* lda, const void* const* b,
                            const int64_t* ldb, const int64_t* k, int64_t M, int64_t N, const float* bias,
                            int kb_per_split, float* out, int64_t ldo, void* stream);

/* out3[r] = split3(act(in[r, :] (+ table[y[r], :]))), in fp32 [rows, K]; act 0 identity, 1 GELU(tanh)
 * (models.py:138), 2 SiLU (models.py:30,148,189); table/y: the label-embedding add of models.py:320. */
int osudit_split3_bf16(const float* in, int64_t rows, int K, int act, const float* table, const int64_t* y,
                       void* out3, void* stream);

/* osudit_ln_modulate with an fp32 branch (x updated in place) and a split3 result h3 bf16 [rows, 3D];
 * any D % 4 == 0; the reference's operation order without contraction (models.py:12-13,160-163). */
int osudit_ln_modulate_f32(float* x, const float* branch, const float* gate, const float* shift,
                           const float* scale, int64_t mod_ld, int64_t rows, int T, int D, void* h3,
                           void* stream);

/* osudit_final_layer with an fp32 branch (models.py:192-196,323-324). */
int osudit_final_layer_f32(float* x, const float* branch, const float* gate, const float* shift,
                           const float* scale, int64_t mod_ld, int64_t rows, int T, int D, const float* w,
                           const float* bias, int out_channels, float* out, void* stream);

/* fp32-mode attention (models.py:164-170 with the mask of sample.py:81-84):
 * osudit_attn_band on fp32 qkv [B*T, 3*H*head_dim] -> fp32 out [B*T, H*head_dim], computed in fp32 on the
 * CUDA cores (head_dim 64 or 72; band, full (w = -1) or generic mask as osudit_attn_band). */
int osudit_attn_band_f32(const float* qkv, float* out, int B, int T, int H, int head_dim, int w_left,
                         int w_right, const uint8_t* mask, void* stream);

/* fp32-mode feature builders (models.py:227-233, 35-36; positional_embedding.py:29-77):
 * osudit_embed_xoc / osudit_timestep_features writing the fp32 feature rows unsplit:
 * a fp32 [B*T, 384 + E]; out fp32 [rows, 256]. */
int osudit_embed_xoc_f32(const float* x, const float* o, const float* c, const float* freqs64, float pf_x,
                         float pf_y, int B, int xrows, int T, int E, float* a, void* stream);
int osudit_timestep_features_f32(const int64_t* t, const float* freqs128, int rows, float* out,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSUDIT_H_ */
