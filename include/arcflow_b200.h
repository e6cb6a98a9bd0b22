/* arcflow_b200 — C ABI of the B200-native ArcFlow denoising hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types, no exceptions across it.
 * Every entry point returns 0 (AFB_OK) or a negative error code; afb_last_error() gives the message.
 * All work is enqueued asynchronously on the caller's CUDA stream (passed as void* = cudaStream_t);
 * no entry point allocates device memory except afb_engine_create / afb_engine_reserve.
 * All device buffers are bf16 unless stated, row-major, and owned by the caller.
 *
 * What each entry replaces in the reference (paths relative to the upstream repo root):
 *   afb_gemm            torch.nn.Linear (+ peft LoRA branch) inside the diffusers MMDiT blocks built at
 *                       lakonlab/models/architecture/arcflow/arcflux.py:60-88 and called at :158-249
 *   afb_attention       F.scaled_dot_product_attention reached through FluxAttention /
 *                       QwenDoubleStreamAttnProcessor2_0 (arcflux.py:180-230, arcqwen.py:136-155)
 *   afb_ln_modulate     AdaLayerNormZero / ZeroSingle / Continuous (arcflux.py:85, blocks' norm1/norm2)
 *   afb_rmsnorm_rope    per-head RMSNorm(q,k) + apply_rotary_emb (pos_embed: arcflux.py:171-173)
 *   afb_small_linear    the M = batch Linears: time_text_embed MLPs (arcflux.py:163-168) and every
 *                       AdaLN modulation Linear (one batched call per forward)
 *   afb_timestep_embed  diffusers Timesteps(256, flip_sin_to_cos=True)
 *   afb_sampler_step    _unpack_mp + ArcFlowPolicy + momentum_integration + _pack_latents
 *                       (lakonlab/pipelines/arcflux_pipeline.py:135-249, :482-510;
 *                        lakonlab/models/diffusions/policies/arcflow.py:25-50)
 *   afb_engine_*        _ArcFluxTransformer2DModel.forward (arcflux.py:134-257) /
 *                       _ArcQwenImageTransformer2DModel.forward (arcqwen.py:106-174) and the denoising
 *                       loop of ArcFluxPipeline.__call__ (arcflux_pipeline.py:453-524)
 */
#ifndef ARCFLOW_B200_H_
#define ARCFLOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct afb_engine afb_engine;

#define AFB_OK 0
#define AFB_ERR_INVALID (-1)
#define AFB_ERR_CUDA (-2)
#define AFB_ERR_UNSUPPORTED (-3)

/* ABI version; bumped on any incompatible change of the structs below. */
#define AFB_ABI_VERSION 3
int afb_abi_version(void);
/* Message of the last failing call on this thread ("" if none). */
const char* afb_last_error(void);
/* Number of kernels this library has launched in this process (all entry points). */
uint64_t afb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM: out[b, r, n] = epi( sum_s sum_k A_s[b, r, k] * W[n, koff_s + k] )
 * ---------------------------------------------------------------------------------------------- */
enum {
  AFB_EPI_BIAS = 0,          /* y + bias                                   (bias may be NULL)        */
  AFB_EPI_BIAS_GELU = 1,     /* gelu_tanh(y + bias)                                                   */
  AFB_EPI_BIAS_GATE_RES = 2, /* res + gate[b, n] * (y + bias)              (AdaLN-Zero gate+residual) */
  AFB_EPI_BIAS_RES = 3,      /* res + y + bias                             (gradient accumulation)   */
  AFB_EPI_BIAS_QKNORM_ROPE = 4 /* fused QKV projection: columns [0, qk_cols) are q heads then k heads (128 each):
                                  RMSNorm over the head (x norm_q / norm_k) + rotary embedding; the rest: y + bias.
                                  diffusers FluxAttention / QwenDoubleStreamAttnProcessor q/k norm + apply_rotary_emb
                                  (reached from arcflux.py:191-197, arcqwen.py:147-155), same bf16 rounding chain */
};

typedef struct afb_gemm_desc {
  /* A operand as up to 3 K-segments (LoRA / concat folded in as extra K). a_k[s] == 0 ends the list.
   * Each segment: bf16 [batches, rows_per_batch, a_k[s]] with leading dim a_ld[s] (elements) and
   * batch stride a_batch_stride[s] (elements). K of every segment must be a multiple of 64. */
  const void* a[3];
  int64_t a_ld[3];
  int64_t a_batch_stride[3];
  int32_t a_k[3];
  int32_t batches;
  int32_t rows_per_batch;
  /* W: bf16 [n, sum(a_k)] K-major (torch Linear.weight layout), leading dim w_ld. */
  const void* w;
  int64_t w_ld;
  int32_t n; /* multiple of 8 */
  int32_t epilogue;
  /* out: bf16 [batches, rows_per_batch, n], leading dim out_ld, batch stride out_batch_stride. */
  void* out;
  int64_t out_ld;
  int64_t out_batch_stride;
  const void* bias; /* bf16 [n] or NULL */
  const void* gate; /* bf16, gate[b * gate_batch_stride + n] (AFB_EPI_BIAS_GATE_RES) */
  int64_t gate_batch_stride;
  const void* res; /* bf16, same shape as out; may alias out */
  int64_t res_ld;
  int64_t res_batch_stride;
  /* Activation-gradient form (dX = dY W): w_transposed != 0 reads W as bf16 [K, n] row-major (leading dim w_ld >= n),
   * i.e. the forward's [out, in] Linear weight used without a transposed copy. The K rows may continue in a second
   * buffer w2 (leading dim w2_ld) after the first w_k rows (dX = dY W + dT A_lora in one accumulation). */
  int32_t w_transposed;
  int32_t w_k;
  const void* w2;
  int64_t w2_ld;
  /* The accumulator is multiplied by alpha before bias / epilogue (0 means 1): the LoRA A-projection uses it for a
   * runtime adapter scale, t = scale * x A^T (peft `scaling`). */
  float alpha;
  /* AFB_EPI_BIAS_QKNORM_ROPE only: RMSNorm weights bf16 [128] for the q and the k heads, rotary table in the afb_rope_pack
   * layout, table position of the first output row of a batch (the image stream of a
   * joint sequence starts at txt_len), number of leading q + k columns, RMSNorm eps (0 -> 1e-6). */
  const void* norm_q;
  const void* norm_k;
  const void* rope;
  int32_t rope_row0;
  int32_t qk_cols;
  float norm_eps;
} afb_gemm_desc;

int afb_gemm(const afb_gemm_desc* desc, void* stream);
/* fp32 [rows, 128] cos / sin tables (adjacent-pair layout, each value repeated twice) -> the rotary table the
 * AFB_EPI_BIAS_QKNORM_ROPE epilogue reads: fp32 [ceil(rows / 32)][2][16][32][4], i.e. for each block of 32 positions and each
 * half of the head the (cos, sin, cos, sin) of pair 2 i, 2 i + 1 of the 32 positions side by side (warp-coalesced for a
 * one-row-per-thread reader). out holds ceil(rows / 32) * 32 * 128 floats. */
int afb_rope_pack(const float* cos_tab, const float* sin_tab, float* out, int64_t rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Joint (text+image) non-causal attention, head_dim 128:
 *   o[b, s, h, :] = softmax(q[b, s, h, :] . k[b, :, h, :] / sqrt(128)) v[b, :, h, :]
 * q/k/v: bf16 [batch, seq, heads*128] views with leading dim *_ld and batch stride *_batch_stride
 * (they may live interleaved in one fused QKV buffer). o: bf16 [batch, seq, heads*128].
 * ---------------------------------------------------------------------------------------------- */
typedef struct afb_attn_desc {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  int64_t q_ld, k_ld, v_ld, o_ld;
  int64_t q_batch_stride, k_batch_stride, v_batch_stride, o_batch_stride;
  int32_t batch, seq, heads;
  float scale; /* 0 -> 1/sqrt(128) */
  float* lse;  /* optional fp32 [batch, heads, seq]: log2-domain logsumexp of the scaled scores, saved for afb_attention_backward */
  float score_bound; /* optional, > 0: the caller guarantees |scale * q_i . k_j| <= score_bound for every pair (e.g. from
                        per-head RMSNorm weights: sqrt(128) * max|w_q| * max|w_k|). The kernel then evaluates
                        P = exp(s - score_bound) against this FIXED reference — same softmax, no running row maximum, no
                        O rescale. 0 = unknown: running-max kernel. Bounds above ~41 (2^60) fall back to it too. */
} afb_attn_desc;

int afb_attention(const afb_attn_desc* desc, void* stream);

/* Backward of afb_attention: dq/dk/dv from d_o, given q, k, v, o and the forward's lse (fp32 [batch, heads, seq]).
 * q/k/v share one leading dim / batch stride (views of a fused QKV buffer), o and d_o another; dq/dk/dv are written with
 * dqkv_ld / dqkv_batch_stride (they may be the three thirds of one fused gradient buffer). delta_ws: fp32 scratch
 * [batch, heads, seq]. Replaces torch autograd through F.scaled_dot_product_attention (arcflux.py:180-230). */
typedef struct afb_attn_bwd_desc {
  const void *q, *k, *v;
  int64_t qkv_ld, qkv_batch_stride;
  const void *o, *d_o;
  int64_t o_ld, o_batch_stride;
  const float* lse;
  float* delta_ws;
  void *dq, *dk, *dv;
  int64_t dqkv_ld, dqkv_batch_stride;
  int32_t batch, seq, heads;
  float scale; /* 0 -> 1/sqrt(128) */
} afb_attn_bwd_desc;
int afb_attention_backward(const afb_attn_bwd_desc* desc, void* stream);
/* Developer instrumentation: with env AFB_ATTN_DEBUG_MODE=7 the kernel records SM-clock stamps of the softmax
 * hand-shake of CTA 0 ([warpgroup 2][iteration 64][stamp 8] int64); this copies them out. Not a product path. */
int afb_debug_attention_trace(int64_t* out, int32_t n);

/* ------------------------------------------------------------------------------------------------
 * y[b, r, :] = LayerNorm(x[b, r, :]; eps, no affine) * (1 + scale[b, :]) + shift[b, :]
 * x, y: bf16 [batches, rows_per_batch, dim] (ld = dim, batch stride given); scale/shift: bf16 vectors
 * addressed as ptr[b * mod_batch_stride + j]. dim must be a multiple of 256 and <= 8192.
 * ---------------------------------------------------------------------------------------------- */
int afb_ln_modulate(const void* x, int64_t x_batch_stride, void* y, int64_t y_batch_stride,
                    const void* scale, const void* shift, int64_t mod_batch_stride,
                    int32_t batches, int32_t rows_per_batch, int32_t dim, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * In place on a fused QKV buffer [batches, seq, ld]: for every head of q (cols q_off + h*128) and
 * k (cols k_off + h*128): x <- RoPE(RMSNorm(x; eps) * w). Rows s < txt_rows use (wq_txt, wk_txt),
 * the rest (wq_img, wk_img) (bf16 [128] each). cos/sin: fp32 [seq, 128] (adjacent-pair convention:
 * out[2j] = x[2j]*cos[2j] - x[2j+1]*sin[2j]; out[2j+1] = x[2j+1]*cos[2j+1] + x[2j]*sin[2j+1]).
 * ---------------------------------------------------------------------------------------------- */
int afb_rmsnorm_rope(void* qkv, int64_t ld, int64_t batch_stride, int32_t q_off, int32_t k_off,
                     int32_t batches, int32_t seq, int32_t heads, int32_t txt_rows,
                     const void* wq_txt, const void* wk_txt, const void* wq_img, const void* wk_img,
                     const float* cos_tab, const float* sin_tab, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Small-M Linear (M = batch <= 8): y[m, n] (+)= sum_k act(x[m, k]) * W[n, k] + bias[n]
 * x: bf16 [m, k] (ld x_ld); W: bf16 [n, k] (ld w_ld); y: bf16 [m, n] (ld y_ld). HBM-bound on W.
 * flags: bit0 = apply SiLU to x on load; bit1 = accumulate into existing y. k % 256 == 0, k <= 4096.
 * ---------------------------------------------------------------------------------------------- */
#define AFB_SL_SILU_IN 1
#define AFB_SL_ACCUMULATE 2
int afb_small_linear(const void* x, int64_t x_ld, const void* w, int64_t w_ld, const void* bias,
                     void* y, int64_t y_ld, int32_t m, int32_t n, int32_t k, int32_t flags,
                     void* stream);

/* out[m, 0:128] = cos(t[m] * f_j), out[m, 128:256] = sin(t[m] * f_j), f_j = exp(-ln(1e4) j / 128).
 * t: fp32 [m] (already in the units the embedder sees, e.g. 1000*sigma); out: bf16 [m, 256]. */
int afb_timestep_embed(const float* t, void* out, int32_t m, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One analytic momentum-integration step in packed token layout.
 * head: bf16 [tokens, head_ld] = [means K*64 | logit K*4 | loggamma (K-1)*4 | pad], the raw output of
 * the three ArcFlow heads (logits NOT yet log-softmaxed). x_in/x_out: fp32 [tokens, 64] packed latents
 * (channel index c*4 + ph*2 + pw); x_out_bf16 (optional): bf16 copy for the next network call.
 *   x_out = x_in - sum_k softmax_k(bf16(log_softmax_k(logit))) * mean_k * exp(lam_k * dt_past)
 *                        * dt_step * phi(lam_k * dt_step),   lam_0 = 0, phi(z) = expm1(z~)/z~
 * dt_past = sigma_src - sigma_start, dt_step = sigma_start - sigma_end, eps = 1e-4 clamp on |z|.
 * ---------------------------------------------------------------------------------------------- */
int afb_sampler_step(const void* head, int64_t head_ld, const float* x_in, float* x_out,
                     void* x_out_bf16, int64_t tokens, int32_t num_gaussians, float sigma_src,
                     float sigma_start, float sigma_end, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training-side policy evaluation (trajectory distillation roll-out) with PER-SAMPLE times, from the raw
 * head tensor of one student call, in packed-token layout. Replaces, per call, the reference's
 *   INTEGRATE  ArcFlowImitationBase.momentum_integration   lakonlab/models/diffusions/arcflow.py:28-79
 *   VELOCITY   ArcFlowPolicy.velocity                       lakonlab/models/diffusions/policies/arcflow.py:52-76
 *   AVERAGE_U  policy_average_u_momentum                    lakonlab/models/diffusions/arcflow.py:81-110
 * plus ArcFlowPolicy.dropout_ (policies/arcflow.py:96-106) through drop_mask. sigma_* / drop_mask / small are
 * HOST arrays (batch <= 64), passed to the kernel by value.
 * ---------------------------------------------------------------------------------------------- */
enum { AFB_POLICY_INTEGRATE = 0, AFB_POLICY_VELOCITY = 1, AFB_POLICY_AVERAGE_U = 2 };

typedef struct afb_policy_args {
  const void* head;         /* bf16 [batch*tokens, head_ld] raw heads (means | logits | loggamma) */
  int64_t head_ld;
  int32_t batch, tokens;    /* tokens per sample */
  int32_t num_gaussians;    /* 16 */
  int32_t mode;
  const float* sigma_src;   /* host [batch] */
  const float* sigma_start; /* host [batch] (VELOCITY: the evaluation time) */
  const float* sigma_end;   /* host [batch] (unused for VELOCITY) */
  const uint8_t* drop_mask; /* host [batch, 16] or NULL: 1 = component dropped */
  const uint8_t* small;     /* host [batch] or NULL: AVERAGE_U falls back to the local velocity where set */
  const float* x_in;        /* fp32 [batch*tokens, 64] (INTEGRATE) */
  float* out;               /* fp32 [batch*tokens, 64]: x_end (INTEGRATE) or u */
  void* out_bf16;           /* optional bf16 copy of out */
  float eps;                /* 1e-4 */
} afb_policy_args;
int afb_policy_eval(const afb_policy_args* args, void* stream);

/* Backward of the velocity-matching loss through AVERAGE_U (args as afb_policy_eval with mode AVERAGE_U; drop_mask is
 * ignored — the grad-carrying policy is never dropped, arcflow.py:183-188):
 *   L = coef/2 * sum (policy_average_u(head) - tgt)^2   ->   dhead[tokens, dh_ld] (fp32, same column layout as head)
 * gets dL/d(means | logits | loggamma); accumulate != 0 adds to the existing contents (the 4 roll-out states of one
 * student step share one head tensor). The rounding of the bf16 log-softmax is treated as straight-through. */
int afb_policy_backward(const afb_policy_args* args, const void* tgt, float* dhead, int64_t dh_ld, float coef,
                        int32_t accumulate, int32_t tgt_is_f32, void* stream);
/* out[n] += sum_t x[t, n]; x fp32 [rows, ld] (bias gradients). out must be initialised by the caller. */
int afb_colsum_f32(const float* x, int64_t ld, float* out, int64_t rows, int32_t n, void* stream);

/* Activation-gradient kernels of the streaming ops (the trunk is frozen: modulation vectors are constants here).
 *  afb_ln_modulate_bwd   dh[b,r,:] (+)= d/dx [LN(x) * (1 + scale[b]) + shift[b]] applied to dy
 *  afb_rowscale          out[b,r,:] = vec[b,:] * x[b,r,:]                         (du = gate (.) dh')
 *  afb_gelu_bwd          dm[r,:] *= gelu_tanh'(pre[r,:])                          (in place)
 *  afb_rmsnorm_rope_bwd  in place on the q|k columns of dqkv: gradient w.r.t. the projection output `raw` */
int afb_ln_modulate_bwd(const void* x, int64_t x_batch_stride, const void* dy, int64_t dy_batch_stride, void* dh,
                        int64_t dh_batch_stride, const void* scale, int64_t mod_batch_stride, int32_t batches,
                        int32_t rows_per_batch, int32_t dim, float eps, int32_t accumulate, void* stream);
int afb_rowscale(const void* x, int64_t x_ld, int64_t x_batch_stride, const void* vec, int64_t vec_batch_stride, void* out,
                 int64_t out_ld, int64_t out_batch_stride, int32_t batches, int32_t rows_per_batch, int32_t cols, void* stream);
int afb_gelu_bwd(void* dm, int64_t dm_ld, const void* pre, int64_t pre_ld, int64_t rows, int32_t cols, void* stream);
int afb_rmsnorm_rope_bwd(void* dqkv, const void* raw, int64_t ld, int64_t batch_stride, int32_t q_off, int32_t k_off,
                         int32_t batches, int32_t seq, int32_t heads, int32_t txt_rows, const void* wq_txt,
                         const void* wk_txt, const void* wq_img, const void* wk_img, const float* cos_tab,
                         const float* sin_tab, float eps, void* stream);

/* out (=|+=) keep_mask (.) act(x) / (1 - p) on a bf16 [batches, rows, cols] view — the LoRA-branch input dropout. The mask
 * is a counter-based hash of (seed, layer_id, logical index (batch*rows + row) * logical_cols + col0 + col); silu_in applies
 * SiLU (rounded to bf16) first; accumulate adds into out (the backward's dx += mask (.) (dT A) / keep). */
int afb_dropout_rows(const void* x, int64_t x_ld, int64_t x_batch_stride, void* out, int64_t out_ld, int64_t out_batch_stride,
                     int32_t batches, int32_t rows_per_batch, int32_t cols, int32_t logical_cols, int32_t col0, uint64_t seed,
                     uint32_t layer_id, float p, int32_t silu_in, int32_t accumulate, void* stream);

/* Weight-gradient ("TN") GEMM: out[m, n] += sum_t a[t, m] * b[t, n]; a bf16 [tokens, m] (ld a_ld), b bf16 [tokens, n],
 * out fp32 [m, n] (must be initialised; results are ADDED). dW = dY^T X for the heads / LoRA pairs. */
int afb_gemm_tn(const void* a, int64_t a_ld, const void* b, int64_t b_ld, float* out, int64_t out_ld, int64_t tokens,
                int32_t m, int32_t n, void* stream);
/* Parameter gradients of y = LN(x) * (1 + scale[b]) + shift[b]: dscale/dshift fp32 [batches, dim] are ADDED to.
 * stats_ws: fp32 scratch of 2 * batches * rows_per_batch floats. x, dy: bf16 [batches, rows, dim] with batch strides. */
int afb_ln_mod_param_grad(const void* x, int64_t x_batch_stride, const void* dy, int64_t dy_batch_stride, float* stats_ws,
                          float* dscale, float* dshift, int32_t batches, int32_t rows_per_batch, int32_t dim, float eps,
                          void* stream);
/* Gradients of a batch-row Linear e = W act(t) + bias (AdaLN modulation Linears): dw[j, d] += sum_b de[b, j] act(t[b, d]),
 * dbias[j] += sum_b de[b, j]. de fp32 [m, n_out], t bf16 [m, k_in], silu_in: act = SiLU (rounded to bf16 as in the forward). */
int afb_rowlinear_param_grad(const float* de, int64_t de_ld, const void* t, int64_t t_ld, float* dw, int64_t dw_ld,
                             float* dbias, int32_t m, int32_t n_out, int32_t k_in, int32_t silu_in, void* stream);
/* Copies an internal activation of the LAST afb_engine_forward to caller memory (training saves them for the backward):
 * which = 0 final image hidden states [batch, img_len, dim], 1 head input (norm_out output) [batch, img_len, dim], 2 temb [batch, dim]. */
int afb_engine_export(afb_engine* e, int32_t which, void* dst, int32_t batch, int32_t txt_len, int32_t img_len, void* stream);

/* Teacher targets u / tgt below are bf16 (a network output, FLUX) or fp32 when *_is_f32 != 0 (the true-CFG combination). */
/* out[b, :] = x[b, :] + coef[b] * u[b, :]  (teacher Euler step, arcflow.py:190). x/out fp32, coef host. */
int afb_axpy_rows(const float* x, const void* u, const float* coef, float* out, void* out_bf16,
                  int32_t batch, int64_t per_sample, int32_t u_is_f32, void* stream);
/* out[b] = mean_i (pred[b, i] - tgt[b, i])^2 — mmgen mse_loss(reduction='flatmean'); pred fp32, out device fp32 [batch]. */
int afb_mse_rows(const float* pred, const void* tgt, float* out, int32_t batch, int64_t per_sample, int32_t tgt_is_f32,
                 void* stream);
/* True classifier-free guidance on a batch-doubled bf16 network output [neg; pos] (each `half` elements):
 * out = pos + (pos - neg) * (guidance_scale - 1), fp32 (GaussianFlow.forward_u, gaussian_flow.py:18-26, 224-254). */
int afb_cfg_combine(const void* both_bf16, float* out, int64_t half, float guidance_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FLUX VAE decoder building blocks. Replaces what the reference reaches through diffusers'
 * AutoencoderKL.decode (lakonlab/pipelines/arcflux_pipeline.py:531-534; lakonlab/models/architecture/diffusers/
 * pretrained.py:69-76): 3x3 convolutions as implicit GEMMs on the tcgen05 kernel, GroupNorm(32) + swish, nearest 2x
 * upsampling, the row softmax of the single-head attention block. Activations are NHWC bf16 (pixel-major rows).
 * ---------------------------------------------------------------------------------------------- */
typedef struct afb_conv_desc {
  const void* x;      /* bf16 [n, h, w_px, c_in] (pixel stride x_ld, 0 -> c_in); c_in % 64 == 0 */
  const void* w;      /* bf16 [c_out, 3, 3, c_in] = [c_out, 9 * c_in]: tap-major, channel-minor rows (pack of the torch
                         [c_out, c_in, 3, 3] weight via permute(0, 2, 3, 1)) */
  const void* bias;   /* bf16 [c_out] or NULL */
  const void* res;    /* bf16 [n, h, w_px, c_out] added to the output (AFB_EPI_BIAS_RES), may alias out */
  void* out;          /* bf16 [n, h, w_px, c_out] (pixel stride out_ld, 0 -> c_out); c_out % 8 == 0 */
  int64_t x_ld, out_ld, res_ld;
  int32_t n, h, w_px, c_in, c_out;
  int32_t epilogue;   /* AFB_EPI_BIAS or AFB_EPI_BIAS_RES */
} afb_conv_desc;
int afb_conv3x3(const afb_conv_desc* desc, void* stream);
/* GroupNorm(32 groups, affine) over NHWC bf16 [n, hw, c] (+ swish when silu != 0): y may alias x. gamma / beta fp32 [c];
 * ws: fp32 scratch of afb_groupnorm_ws_floats(n, hw) floats. c in {64, 128, 256, 512, 1024, 2048}. Deterministic. */
int afb_groupnorm_ws_floats(int32_t n, int64_t hw);
int afb_groupnorm(const void* x, void* y, const float* gamma, const float* beta, float* ws, int64_t ws_floats, int32_t n,
                  int64_t hw, int32_t c, float eps, int32_t silu, void* stream);
/* nearest-neighbour 2x upsampling, NHWC bf16 [n, h, w, c] -> [n, 2h, 2w, c] */
int afb_upsample2x(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, void* stream);
/* in-place softmax over the last dim of a bf16 matrix [rows, cols] (row stride ld) */
int afb_softmax_rows(void* x, int64_t ld, int64_t rows, int32_t cols, void* stream);
/* latents fp32 NCHW [n, c_in, h, w] -> (z / scale + shift) bf16 NHWC [n, h, w, c_pad], channels >= c_in zero-filled */
int afb_vae_pre(const float* z, void* out, int32_t n, int32_t c_in, int32_t h, int32_t w, int32_t c_pad, float scale,
                float shift, void* stream);
/* bf16 NHWC rows (pixel stride x_ld >= 8) -> fp32 NCHW [n, c_out, h, w], first c_out <= 8 channels */
int afb_vae_post(const void* x, int64_t x_ld, float* out, int32_t n, int32_t c_out, int32_t h, int32_t w, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Step glue over ONE flat fp32 arena of all trainable adapter tensors (reference
 * lakonlab/models/base.py:76-103 + configs/flux/_ddp_train.py:13-26 + lakonlab/runner/hooks/ema_hook.py:86-121).
 * ---------------------------------------------------------------------------------------------- */
/* out[0] = sum g^2 (device scalar, fp32). */
int afb_grad_norm_sq(const float* grads, int64_t n, float* out, void* stream);
/* Same with caller-owned scratch (afb_grad_norm_scratch_floats() floats, zeroed once): one scratch per optimizer instance
 * makes concurrent launches on different streams safe and the call CUDA-graph capturable (no allocation inside). */
int afb_grad_norm_scratch_floats(void);
int afb_grad_norm_sq_ws(const float* grads, int64_t n, float* out, float* scratch, int64_t scratch_floats, void* stream);

typedef struct afb_adamw_args {
  float* params;          /* fp32 [n], updated in place */
  const float* grads;     /* fp32 [n] (already all-reduced / averaged) */
  float* exp_avg;         /* fp32 [n] */
  float* exp_avg_sq;      /* fp32 [n] */
  float* ema;             /* fp32 [n] or NULL */
  void* bf16_shadow;      /* bf16 [n] or NULL: rounded copy of the new params for the engine's packed weights */
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay;
  int32_t step;           /* 1-based optimizer step (bias correction) */
  float max_norm;         /* > 0: clip_grad_norm_ with *grad_norm_sq; a non-finite norm skips the update */
  const float* grad_norm_sq; /* device scalar from afb_grad_norm_sq */
  int32_t* skipped;       /* device flag, set to 1 when the update was skipped */
  float ema_momentum;     /* ema = m * ema + (1 - m) * p; < 0 leaves ema untouched */
  int32_t ema_copy;       /* 1: ema = p (before the EMA start iteration) */
  int64_t lr_mult_begin, lr_mult_end; /* element range using lr * lr_mult (proj_out_loggamma: 0.1) */
  float lr_mult;
  float skip_norm;        /* > 0 (with max_norm > 0): a gradient norm above it skips the update like a non-finite one —
                             `<k>_grad_clip_skip_ratio` x `<k>_grad_clip`, lakonlab/models/base.py:81,91-95 */
} afb_adamw_args;
int afb_adamw_ema_step(const afb_adamw_args* args, void* stream);

/* AdamW with block-wise 8-bit moment state — what `optimizer=dict(type='AdamW8bit')` (configs/flux/_ddp_train.py:18-26,
 * registered from bitsandbytes.optim by lakonlab/runner/optimizer/builder.py:11-24) keeps per tensor of >= 4096 elements:
 * both moments as one byte per element indexing a 256-entry "dynamic" code book (signed for exp_avg, unsigned for
 * exp_avg_sq) times one fp32 absmax per block of `blocksize` consecutive elements. Per block and step: de-quantise,
 * update the moments in fp32, take the new absmax, update the parameter, re-quantise to the nearest code (exp_avg keeps its
 * sign). bitsandbytes is not vendored and its version is not pinned by the reference (requirements.txt:11): this restates
 * the published block-wise algorithm (Dettmers et al., "8-bit Optimizers via Block-wise Quantization", and the
 * kOptimizerStatic8bit2StateBlockwise kernel of bitsandbytes >= 0.44: block size 256). Clip / skip / EMA / bf16 shadow are
 * the same as afb_adamw_ema_step. `n` must be a multiple of `blocksize` (the host pads every tensor's slot). */
typedef struct afb_adamw8bit_args {
  afb_adamw_args base;     /* exp_avg / exp_avg_sq ignored */
  uint8_t* state1;         /* [n] codes of exp_avg */
  uint8_t* state2;         /* [n] codes of exp_avg_sq */
  float* absmax1;          /* [n / blocksize] */
  float* absmax2;          /* [n / blocksize] */
  const float* qmap1;      /* [256] ascending signed code book */
  const float* qmap2;      /* [256] ascending unsigned code book */
  int32_t blocksize;       /* 256 */
  int32_t reserved0;
} afb_adamw8bit_args;
int afb_adamw8bit_ema_step(const afb_adamw8bit_args* args, void* stream);

/* fp32 -> bf16 cast of a contiguous buffer. */
int afb_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Engine: whole-transformer forward + 2-NFE denoise loop over packed weights.
 * ---------------------------------------------------------------------------------------------- */
#define AFB_ARCH_FLUX 0
#define AFB_ARCH_QWEN 1

typedef struct afb_model_desc {
  int32_t arch;         /* AFB_ARCH_FLUX | AFB_ARCH_QWEN */
  int32_t num_double;   /* FLUX 19, Qwen 60 */
  int32_t num_single;   /* FLUX 38, Qwen 0 */
  int32_t dim;          /* 3072 */
  int32_t heads;        /* 24 (head_dim fixed at 128) */
  int32_t mlp_dim;      /* 12288 */
  int32_t in_channels;  /* 64 */
  int32_t txt_dim;      /* FLUX 4096, Qwen 3584 */
  int32_t pooled_dim;   /* FLUX 768, Qwen 0 */
  int32_t guidance;     /* FLUX.1-dev: 1 */
  int32_t num_gaussians; /* 16 */
  int32_t lora_rank;    /* 256; 0 = no adapter branches (teacher trunk) */
  int32_t head_mode;    /* 0 = ArcFlow 3 heads (means|logits|loggamma), 1 = stock proj_out (teacher) */
  int32_t ignore_lora;  /* 1: frozen-trunk forward (teacher tied to the student's packed weights): LoRA A/B never read */
} afb_model_desc;

/* Packed per-block weights. All bf16. W*: [out, in(+rank)] K-major with the LoRA B matrix appended
 * along K when lora_rank > 0 (so ld = in + rank); la_*: LoRA A matrices [rank, in]. NULL la_* = no LoRA
 * on that Linear (then ld = in). */
typedef struct afb_double_block {
  /* image stream */
  const void *img_qkv_w, *img_qkv_b;         /* [3D, D], [3D] */
  const void *img_nq, *img_nk;               /* RMSNorm weights [128] */
  const void *img_out_w, *img_out_b;         /* [D, D] */
  const void *img_up_w, *img_up_b, *img_up_la;       /* [M, D(+r)], [M], [r, D] */
  const void *img_down_w, *img_down_b, *img_down_la; /* [D, M(+r)], [D], [r, M] */
  /* text stream */
  const void *txt_qkv_w, *txt_qkv_b;
  const void *txt_nq, *txt_nk;
  const void *txt_out_w, *txt_out_b;
  const void *txt_up_w, *txt_up_b, *txt_up_la;
  const void *txt_down_w, *txt_down_b, *txt_down_la;
  int64_t img_mod_off, txt_mod_off; /* column offset of this block's 6*D modulation chunk */
  float qk_bound;  /* > 0: upper bound on |q . k| / sqrt(128) after RMSNorm + RoPE over BOTH streams of the joint sequence
                      (afb_attn_desc.score_bound); 0 = not provided */
  float reserved0;
} afb_double_block;

typedef struct afb_single_block {
  const void *qkv_w, *qkv_b;              /* [3D, D] */
  const void *nq, *nk;
  const void *mlp_w, *mlp_b, *mlp_la;     /* [M, D(+r)], [M], [r, D] */
  const void *out_w, *out_b, *out_la;     /* [D, D+M(+r)], [D], [r, D+M] */
  int64_t mod_off;                        /* 3*D chunk */
  float qk_bound;                         /* as in afb_double_block */
  float reserved0;
} afb_single_block;

typedef struct afb_weights {
  const void *x_emb_w, *x_emb_b;          /* [D, in_channels] */
  const void *ctx_w, *ctx_b;              /* [D, txt_dim] */
  const void *txt_norm_w;                 /* Qwen: RMSNorm(txt_dim) weight, else NULL */
  /* timestep / guidance / pooled-text embedders: Linear(256->D), Linear(D->D) each */
  const void *t1_w, *t1_b, *t1_la, *t1_lb;   /* LoRA kept separate here: la [r,256], lb [D,r] */
  const void *t2_w, *t2_b, *t2_la, *t2_lb;
  const void *g1_w, *g1_b, *g2_w, *g2_b;
  const void *p1_w, *p1_b, *p2_w, *p2_b;
  /* all AdaLN modulation Linears of the model concatenated along the output dim */
  const void *mod_w, *mod_b;              /* [mod_total, D], [mod_total] */
  int64_t mod_total;
  int64_t norm_out_mod_off;               /* 2*D chunk (scale, shift) */
  const void *head_w, *head_b;            /* [head_n, D], [head_n]; head_n padded to a multiple of 8 */
  const void *alt_norm_out_w, *alt_norm_out_b; /* optional [2D, D], [2D]: norm_out Linear used instead of the mod_w chunk */
  int32_t head_n;
  const afb_double_block* dbl;
  const afb_single_block* sgl;
} afb_weights;


int afb_engine_create(const afb_model_desc* desc, afb_engine** out);
void afb_engine_destroy(afb_engine* e);
/* Borrows the packed weight pointers (caller keeps the storage alive). */
int afb_engine_bind(afb_engine* e, const afb_weights* w);
/* Runtime LoRA scale (joint_attention_kwargs['scale'] in the reference); 1.0 by default. The packed
 * [W | B] weights assume scale 1; other values are applied by scaling the A-projection output. */
int afb_engine_set_lora_scale(afb_engine* e, float scale);
/* Skip every LoRA branch at run time (the A-projection launches and the K-extension columns of the packed [W | B]
 * weights are not read). Used after the adapter has been merged into the base weights, W <- W + scale * B A — diffusers'
 * `fuse_lora()`; SURVEY.md §8f rank 4 — which the host does with afb_gemm (transposed-W mode, residual epilogue, in place
 * on the packed buffers). Inference only: the training entry points need the separate branch. */
int afb_engine_set_ignore_lora(afb_engine* e, int32_t on);
/* Bytes of workspace the engine needs for (batch, txt_len, img_len); afb_engine_reserve allocates it. */
size_t afb_engine_workspace_bytes(const afb_engine* e, int32_t batch, int32_t txt_len, int32_t img_len);
int afb_engine_reserve(afb_engine* e, int32_t batch, int32_t txt_len, int32_t img_len);

typedef struct afb_forward_args {
  int32_t batch, txt_len, img_len;
  const void* latents;        /* bf16 [batch, img_len, in_channels] packed tokens */
  const void* txt;            /* bf16 [batch, txt_len, txt_dim] */
  const void* pooled;         /* bf16 [batch, pooled_dim] or NULL */
  const float* timestep;      /* fp32 [batch]: value the time embedder sees (FLUX: bf16(bf16(sigma)*1000)) */
  const float* guidance;      /* fp32 [batch] or NULL (FLUX: bf16(bf16(g)*1000)) */
  const float* rope_cos;      /* fp32 [txt_len + img_len, 128] */
  const float* rope_sin;
  void* head_out;             /* bf16 [batch*img_len, head_n] raw heads (means | logits | loggamma) */
} afb_forward_args;

int afb_engine_forward(afb_engine* e, const afb_forward_args* args, void* stream);

typedef struct afb_denoise_args {
  afb_forward_args fwd;       /* latents / head_out fields are ignored (engine-internal buffers) */
  int32_t nfe;
  const float* sigmas;        /* host fp32 [nfe + 1]: sigma at each network call, then the final sigma (0) */
  const float* timesteps;     /* host fp32 [nfe]: value fed to the time embedder at each call */
  float* x;                   /* fp32 [batch, img_len, 64] packed latents, updated in place */
  float eps;                  /* 1e-4 */
} afb_denoise_args;

int afb_engine_denoise(afb_engine* e, const afb_denoise_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Adapter-only backward through the frozen trunk (train_fwd_bwd, lakonlab/models/base_diffusion.py:14-62: the
 * reference gets these from torch autograd + gradient checkpointing over the diffusers blocks).
 *   afb_engine_train_reserve   training workspace: one residual-stream checkpoint per block (+ the final one) and the
 *                              recompute / gradient buffers
 *   afb_engine_forward_train   afb_engine_forward that also stores the checkpoints
 *   afb_engine_backward        walks the blocks in reverse: recompute the block from its checkpoint, then
 *                              dX GEMMs with the forward weights read transposed, tcgen05 attention backward,
 *                              LN / GELU / RMSNorm+RoPE backward, and the LoRA gradients dB = dY^T T, dA = dT^T X
 *                              accumulated (+=) in fp32 into the caller's buffers (NULL = skip that tensor).
 * d_head_in: bf16 [batch, img_len, dim] = gradient w.r.t. the norm_out output (the heads' input). With d_mod (fp32
 * [batch, mod_total], accumulated into) the gradients of every AdaLN shift / scale / gate vector are produced too (one
 * extra GEMM per gate: the un-gated branch output is recomputed); afb_engine_backward_embed turns them into the
 * timestep-embedder LoRA gradients. Without d_mod the modulation vectors are treated as constants.
 * ---------------------------------------------------------------------------------------------- */
typedef struct afb_double_block_grads {
  float *img_up_la, *img_up_lb, *img_down_la, *img_down_lb, *txt_up_la, *txt_up_lb, *txt_down_la, *txt_down_lb;
} afb_double_block_grads;
typedef struct afb_single_block_grads {
  float *mlp_la, *mlp_lb, *out_la, *out_lb;
} afb_single_block_grads;
typedef struct afb_backward_args {
  afb_forward_args fwd;      /* the forward's arguments (latents / head_out unused) */
  const void* d_head_in;     /* bf16 [batch, img_len, dim] */
  const afb_double_block_grads* dbl; /* [num_double] */
  const afb_single_block_grads* sgl; /* [num_single] */
  float* d_mod;              /* optional fp32 [batch, mod_total] */
} afb_backward_args;
/* peft lora_dropout (configs/flux/arcflux_2nfe_k16.py:40-48: 0.05, train only): the LoRA branches of
 * afb_engine_forward_train / afb_engine_backward(_embed) see dropout(x) with the keep mask
 *   16 bits of hash(seed, layer id, logical element index / 2) >= round(p * 2^16): the low half of the 32-bit hash for the
 *   even element of a pair, the high half for the odd one   (counter-based: recompute and backward regenerate it).
 * Set a fresh seed before every afb_engine_forward_train; p = 0 (default) disables it. afb_engine_forward / _denoise
 * never drop. */
int afb_engine_set_lora_dropout(afb_engine* e, float p, uint64_t seed);
int afb_engine_train_reserve(afb_engine* e, int32_t batch, int32_t txt_len, int32_t img_len);
/* Activation stash: with 180 GB of HBM the train forward can keep each block's expensive outputs (raw QKV, attention
 * output + log-sum-exp, MLP pre-activation, the un-gated branch outputs and the mid-block residual stream: ~62 GB for FLUX
 * at batch 4, 1024 px) so that afb_engine_backward copies them back instead of recomputing the block (no QKV / MLP-up /
 * branch GEMMs and no attention forward in the backward) — what the reference's `checkpointing=True`
 * (configs/flux/arcflux_2nfe_k16.py:38, torch.utils.checkpoint re-running every block) trades the other way on 80 GB
 * parts. Gradients are identical to the recompute path up to the un-fused epilogues' extra bf16 rounding.
 * afb_engine_stash_bytes: size for a shape. afb_engine_set_activation_stash(on = 1, shape): allocate (AFB_ERR_CUDA if the
 * device cannot hold it; the engine then keeps recomputing) — on = 0 frees it. */
int64_t afb_engine_stash_bytes(afb_engine* e, int32_t batch, int32_t txt_len, int32_t img_len);
int afb_engine_set_activation_stash(afb_engine* e, int32_t on, int32_t batch, int32_t txt_len, int32_t img_len);
int afb_engine_forward_train(afb_engine* e, const afb_forward_args* args, void* stream);
int afb_engine_backward(afb_engine* e, const afb_backward_args* args, void* stream);
/* d_mod (fp32 [batch, mod_total]: afb_engine_backward's output plus the caller-written norm_out slot) -> temb ->
 * gradients of the timestep embedder's LoRA pairs (time_text_embed.timestep_embedder.linear_1/2), accumulated (+=). */
typedef struct afb_embed_grads {
  float *t1_la, *t1_lb, *t2_la, *t2_lb;
} afb_embed_grads;
int afb_engine_backward_embed(afb_engine* e, const afb_forward_args* fwd, const float* d_mod, const afb_embed_grads* grads,
                              void* stream);

/* Optional instrumentation: when on, every tensor-core launch of forward/denoise is bracketed by CUDA
 * events on the caller's stream. afb_engine_read_profile synchronises the device, returns the totals
 * since the last read (algorithmic FLOPs: 2*M*N*K per GEMM incl. LoRA K-extension, 4*B*H*S^2*128 per
 * attention launch) and clears them. Off by default; the timed path carries no events. */
typedef struct afb_profile {
  double gemm_ms, attn_ms;
  double gemm_flops, attn_flops;
  int64_t gemm_launches, attn_launches;
  /* of the GEMM figures above: the launches whose epilogue also does the per-head RMSNorm + RoPE (AFB_EPI_BIAS_QKNORM_ROPE) —
   * they carry the work of the former rmsnorm_rope kernel, so their time per FLOP is not a pure-GEMM figure */
  double gemm_fused_qk_ms, gemm_fused_qk_flops;
  int64_t gemm_fused_qk_launches;
} afb_profile;
int afb_engine_set_profiling(afb_engine* e, int32_t on);
int afb_engine_read_profile(afb_engine* e, afb_profile* out);

#ifdef __cplusplus
}
#endif
#endif /* ARCFLOW_B200_H_ */
