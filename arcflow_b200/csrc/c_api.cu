// arcflow_b200 — extern "C" surface (see include/arcflow_b200.h). Nothing but argument forwarding:
// errors become return codes, never exceptions.
#include "common.cuh"
#include "../../include/arcflow_b200.h"

namespace afb {
uint64_t launch_count();
void count_launch(int n);
int gemm_launch(const afb_gemm_desc* d, cudaStream_t stream);
int attention_launch(const afb_attn_desc* d, cudaStream_t stream);
int attention_read_trace(long long* out, int n);
int attention_backward_launch(const afb_attn_bwd_desc* d, cudaStream_t stream);
int ln_modulate_launch(const void* x, int64_t x_bs, void* y, int64_t y_bs, const void* scale,
                       const void* shift, int64_t mod_bs, int batches, int rows_per_batch, int dim,
                       float eps, cudaStream_t stream);
int rmsnorm_rope_launch(void* qkv, int64_t ld, int64_t bs, int q_off, int k_off, int batches, int seq,
                        int heads, int txt_rows, const void* wq_txt, const void* wk_txt,
                        const void* wq_img, const void* wk_img, const float* cos_tab,
                        const float* sin_tab, float eps, cudaStream_t stream);
int small_linear_launch(const void* x, int64_t x_ld, const void* w, int64_t w_ld, const void* bias,
                        void* y, int64_t y_ld, int m, int n, int k, int flags, cudaStream_t stream);
int timestep_embed_launch(const float* t, void* out, int m, cudaStream_t stream);
int sampler_step_launch(const void* head, int64_t head_ld, const float* x_in, float* x_out,
                        void* x_out_bf16, int64_t tokens, int num_gaussians, float sigma_src,
                        float sigma_start, float sigma_end, float eps, cudaStream_t stream);
int cast_f32_bf16_launch(const float* in, void* out, int64_t n, cudaStream_t stream);
int policy_eval_launch(const afb_policy_args* a, cudaStream_t stream);
int policy_backward_launch(const afb_policy_args* a, const void* tgt, float* dhead, int64_t dh_ld, float coef,
                           int accumulate, int tgt_f32, cudaStream_t stream);
int dropout_rows_launch(const void* x, int64_t x_ld, int64_t x_bs, void* out, int64_t out_ld, int64_t out_bs, int batches,
                        int rows_per_batch, int cols, int logical_cols, int col0, uint64_t seed, uint32_t layer_id, float p,
                        int silu_in, int accumulate, cudaStream_t stream);
int cfg_combine_launch(const void* both_bf16, float* out, int64_t half, float guidance_scale, cudaStream_t stream);
int colsum_f32_launch(const float* x, int64_t ld, float* out, int64_t rows, int n, cudaStream_t stream);
int grad_norm_sq_launch(const float* g, int64_t n, float* out, cudaStream_t stream);
int grad_norm_scratch_floats();
int rope_pack_launch(const float* cos_tab, const float* sin_tab, float* out, int64_t rows, cudaStream_t stream);
int conv3x3_launch(const afb_conv_desc* d, cudaStream_t stream);
int vae_pre_launch(const float* z, void* out, int n, int c_in, int h, int w, int c_pad, float scale, float shift, cudaStream_t stream);
int vae_post_launch(const void* x, int64_t x_ld, float* out, int n, int c_out, int h, int w, cudaStream_t stream);
int groupnorm_ws_floats(int n, long long hw);
int groupnorm_launch(const void* x, void* y, const float* gamma, const float* beta, float* ws, int64_t ws_floats, int n,
                     long long hw, int c, float eps, int silu_on, cudaStream_t stream);
int upsample2x_launch(const void* x, void* y, int n, int h, int w, int c, cudaStream_t stream);
int softmax_rows_launch(void* x, int64_t ld, long long rows, int cols, cudaStream_t stream);
int grad_norm_sq_ws_launch(const float* g, int64_t n, float* out, float* scratch, int64_t scratch_floats, cudaStream_t stream);
int ln_modulate_bwd_launch(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, void* dh, int64_t dh_bs,
                           const void* scale, int64_t mod_bs, int batches, int rows_per_batch, int dim, float eps,
                           int accumulate, cudaStream_t stream);
int rowscale_launch(const void* x, int64_t x_ld, int64_t x_bs, const void* vec, int64_t vec_bs, void* out, int64_t out_ld,
                    int64_t out_bs, int batches, int rows_per_batch, int cols, cudaStream_t stream);
int gelu_bwd_launch(void* dm, int64_t dm_ld, const void* pre, int64_t pre_ld, int64_t rows, int cols, cudaStream_t stream);
int rmsnorm_rope_bwd_launch(void* dqkv, const void* raw, int64_t ld, int64_t bs, int q_off, int k_off, int batches, int seq,
                            int heads, int txt_rows, const void* wq_txt, const void* wk_txt, const void* wq_img,
                            const void* wk_img, const float* cos_tab, const float* sin_tab, float eps, cudaStream_t stream);
int adamw_ema_launch(const afb_adamw_args* a, cudaStream_t stream);
int adamw8bit_ema_launch(const afb_adamw8bit_args* a, cudaStream_t stream);
int gemm_tn_launch(const void* a, int64_t a_ld, const void* b, int64_t b_ld, float* out, int64_t out_ld, int64_t tokens,
                   int m, int n, cudaStream_t stream);
int ln_mod_param_grad_launch(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, float* stats_ws, float* dscale,
                             float* dshift, int batches, int rows_per_batch, int dim, float eps, cudaStream_t stream);
int rowlinear_param_grad_launch(const float* de, int64_t de_ld, const void* t, int64_t t_ld, float* dw, int64_t dw_ld,
                                float* dbias, int m, int n_out, int k_in, int silu_in, cudaStream_t stream);
int axpy_rows_launch(const float* x, const void* u, const float* coef, float* out, void* out_bf16, int batch,
                     int64_t per_sample, int u_f32, cudaStream_t stream);
int mse_rows_launch(const float* pred, const void* tgt, float* out, int batch, int64_t per_sample, int tgt_f32,
                    cudaStream_t stream);
}  // namespace afb

extern "C" {

int afb_abi_version(void) { return AFB_ABI_VERSION; }
const char* afb_last_error(void) { return afb::get_last_error(); }
uint64_t afb_launch_count(void) { return afb::launch_count(); }

int afb_gemm(const afb_gemm_desc* desc, void* stream) {
  int rc = afb::gemm_launch(desc, static_cast<cudaStream_t>(stream));
  if (rc == AFB_OK) afb::count_launch(1);
  return rc;
}
int afb_attention_backward(const afb_attn_bwd_desc* desc, void* stream) {
  return afb::attention_backward_launch(desc, static_cast<cudaStream_t>(stream));
}
int afb_debug_attention_trace(int64_t* out, int32_t n) {
  return afb::attention_read_trace(reinterpret_cast<long long*>(out), n);
}
int afb_attention(const afb_attn_desc* desc, void* stream) {
  return afb::attention_launch(desc, static_cast<cudaStream_t>(stream));
}
int afb_ln_modulate(const void* x, int64_t x_batch_stride, void* y, int64_t y_batch_stride,
                    const void* scale, const void* shift, int64_t mod_batch_stride, int32_t batches,
                    int32_t rows_per_batch, int32_t dim, float eps, void* stream) {
  return afb::ln_modulate_launch(x, x_batch_stride, y, y_batch_stride, scale, shift, mod_batch_stride,
                                 batches, rows_per_batch, dim, eps, static_cast<cudaStream_t>(stream));
}
int afb_rmsnorm_rope(void* qkv, int64_t ld, int64_t batch_stride, int32_t q_off, int32_t k_off,
                     int32_t batches, int32_t seq, int32_t heads, int32_t txt_rows, const void* wq_txt,
                     const void* wk_txt, const void* wq_img, const void* wk_img, const float* cos_tab,
                     const float* sin_tab, float eps, void* stream) {
  return afb::rmsnorm_rope_launch(qkv, ld, batch_stride, q_off, k_off, batches, seq, heads, txt_rows,
                                  wq_txt, wk_txt, wq_img, wk_img, cos_tab, sin_tab, eps,
                                  static_cast<cudaStream_t>(stream));
}
int afb_small_linear(const void* x, int64_t x_ld, const void* w, int64_t w_ld, const void* bias,
                     void* y, int64_t y_ld, int32_t m, int32_t n, int32_t k, int32_t flags,
                     void* stream) {
  return afb::small_linear_launch(x, x_ld, w, w_ld, bias, y, y_ld, m, n, k, flags,
                                  static_cast<cudaStream_t>(stream));
}
int afb_timestep_embed(const float* t, void* out, int32_t m, void* stream) {
  return afb::timestep_embed_launch(t, out, m, static_cast<cudaStream_t>(stream));
}
int afb_sampler_step(const void* head, int64_t head_ld, const float* x_in, float* x_out,
                     void* x_out_bf16, int64_t tokens, int32_t num_gaussians, float sigma_src,
                     float sigma_start, float sigma_end, float eps, void* stream) {
  return afb::sampler_step_launch(head, head_ld, x_in, x_out, x_out_bf16, tokens, num_gaussians,
                                  sigma_src, sigma_start, sigma_end, eps,
                                  static_cast<cudaStream_t>(stream));
}
int afb_policy_eval(const afb_policy_args* args, void* stream) {
  return afb::policy_eval_launch(args, static_cast<cudaStream_t>(stream));
}
int afb_policy_backward(const afb_policy_args* args, const void* tgt, float* dhead, int64_t dh_ld, float coef,
                        int32_t accumulate, int32_t tgt_is_f32, void* stream) {
  return afb::policy_backward_launch(args, tgt, dhead, dh_ld, coef, accumulate, tgt_is_f32, static_cast<cudaStream_t>(stream));
}
int afb_dropout_rows(const void* x, int64_t x_ld, int64_t x_bs, void* out, int64_t out_ld, int64_t out_bs, int32_t batches,
                     int32_t rows_per_batch, int32_t cols, int32_t logical_cols, int32_t col0, uint64_t seed, uint32_t layer_id,
                     float p, int32_t silu_in, int32_t accumulate, void* stream) {
  return afb::dropout_rows_launch(x, x_ld, x_bs, out, out_ld, out_bs, batches, rows_per_batch, cols, logical_cols, col0, seed,
                                  layer_id, p, silu_in, accumulate, static_cast<cudaStream_t>(stream));
}
int afb_cfg_combine(const void* both_bf16, float* out, int64_t half, float guidance_scale, void* stream) {
  return afb::cfg_combine_launch(both_bf16, out, half, guidance_scale, static_cast<cudaStream_t>(stream));
}
int afb_colsum_f32(const float* x, int64_t ld, float* out, int64_t rows, int32_t n, void* stream) {
  return afb::colsum_f32_launch(x, ld, out, rows, n, static_cast<cudaStream_t>(stream));
}
int afb_gemm_tn(const void* a, int64_t a_ld, const void* b, int64_t b_ld, float* out, int64_t out_ld, int64_t tokens,
                int32_t m, int32_t n, void* stream) {
  return afb::gemm_tn_launch(a, a_ld, b, b_ld, out, out_ld, tokens, m, n, static_cast<cudaStream_t>(stream));
}
int afb_ln_mod_param_grad(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, float* stats_ws, float* dscale,
                          float* dshift, int32_t batches, int32_t rows_per_batch, int32_t dim, float eps, void* stream) {
  return afb::ln_mod_param_grad_launch(x, x_bs, dy, dy_bs, stats_ws, dscale, dshift, batches, rows_per_batch, dim, eps,
                                       static_cast<cudaStream_t>(stream));
}
int afb_rowlinear_param_grad(const float* de, int64_t de_ld, const void* t, int64_t t_ld, float* dw, int64_t dw_ld,
                             float* dbias, int32_t m, int32_t n_out, int32_t k_in, int32_t silu_in, void* stream) {
  return afb::rowlinear_param_grad_launch(de, de_ld, t, t_ld, dw, dw_ld, dbias, m, n_out, k_in, silu_in,
                                          static_cast<cudaStream_t>(stream));
}
int afb_ln_modulate_bwd(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, void* dh, int64_t dh_bs, const void* scale,
                        int64_t mod_bs, int32_t batches, int32_t rows_per_batch, int32_t dim, float eps, int32_t accumulate,
                        void* stream) {
  return afb::ln_modulate_bwd_launch(x, x_bs, dy, dy_bs, dh, dh_bs, scale, mod_bs, batches, rows_per_batch, dim, eps,
                                     accumulate, static_cast<cudaStream_t>(stream));
}
int afb_rowscale(const void* x, int64_t x_ld, int64_t x_bs, const void* vec, int64_t vec_bs, void* out, int64_t out_ld,
                 int64_t out_bs, int32_t batches, int32_t rows_per_batch, int32_t cols, void* stream) {
  return afb::rowscale_launch(x, x_ld, x_bs, vec, vec_bs, out, out_ld, out_bs, batches, rows_per_batch, cols,
                              static_cast<cudaStream_t>(stream));
}
int afb_gelu_bwd(void* dm, int64_t dm_ld, const void* pre, int64_t pre_ld, int64_t rows, int32_t cols, void* stream) {
  return afb::gelu_bwd_launch(dm, dm_ld, pre, pre_ld, rows, cols, static_cast<cudaStream_t>(stream));
}
int afb_rmsnorm_rope_bwd(void* dqkv, const void* raw, int64_t ld, int64_t bs, int32_t q_off, int32_t k_off, int32_t batches,
                         int32_t seq, int32_t heads, int32_t txt_rows, const void* wq_txt, const void* wk_txt,
                         const void* wq_img, const void* wk_img, const float* cos_tab, const float* sin_tab, float eps,
                         void* stream) {
  return afb::rmsnorm_rope_bwd_launch(dqkv, raw, ld, bs, q_off, k_off, batches, seq, heads, txt_rows, wq_txt, wk_txt, wq_img,
                                      wk_img, cos_tab, sin_tab, eps, static_cast<cudaStream_t>(stream));
}
int afb_grad_norm_sq(const float* grads, int64_t n, float* out, void* stream) {
  return afb::grad_norm_sq_launch(grads, n, out, static_cast<cudaStream_t>(stream));
}
int afb_rope_pack(const float* cos_tab, const float* sin_tab, float* out, int64_t rows, void* stream) {
  return afb::rope_pack_launch(cos_tab, sin_tab, out, rows, static_cast<cudaStream_t>(stream));
}
int afb_conv3x3(const afb_conv_desc* desc, void* stream) {
  int rc = afb::conv3x3_launch(desc, static_cast<cudaStream_t>(stream));
  if (rc == AFB_OK) afb::count_launch(1);
  return rc;
}
int afb_groupnorm_ws_floats(int32_t n, int64_t hw) { return afb::groupnorm_ws_floats(n, hw); }
int afb_groupnorm(const void* x, void* y, const float* gamma, const float* beta, float* ws, int64_t ws_floats, int32_t n,
                  int64_t hw, int32_t c, float eps, int32_t silu, void* stream) {
  return afb::groupnorm_launch(x, y, gamma, beta, ws, ws_floats, n, hw, c, eps, silu, static_cast<cudaStream_t>(stream));
}
int afb_upsample2x(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, void* stream) {
  return afb::upsample2x_launch(x, y, n, h, w, c, static_cast<cudaStream_t>(stream));
}
int afb_softmax_rows(void* x, int64_t ld, int64_t rows, int32_t cols, void* stream) {
  return afb::softmax_rows_launch(x, ld, rows, cols, static_cast<cudaStream_t>(stream));
}
int afb_vae_pre(const float* z, void* out, int32_t n, int32_t c_in, int32_t h, int32_t w, int32_t c_pad, float scale,
                float shift, void* stream) {
  return afb::vae_pre_launch(z, out, n, c_in, h, w, c_pad, scale, shift, static_cast<cudaStream_t>(stream));
}
int afb_vae_post(const void* x, int64_t x_ld, float* out, int32_t n, int32_t c_out, int32_t h, int32_t w, void* stream) {
  return afb::vae_post_launch(x, x_ld, out, n, c_out, h, w, static_cast<cudaStream_t>(stream));
}
int afb_grad_norm_scratch_floats(void) { return afb::grad_norm_scratch_floats(); }
int afb_grad_norm_sq_ws(const float* grads, int64_t n, float* out, float* scratch, int64_t scratch_floats, void* stream) {
  return afb::grad_norm_sq_ws_launch(grads, n, out, scratch, scratch_floats, static_cast<cudaStream_t>(stream));
}
int afb_adamw_ema_step(const afb_adamw_args* args, void* stream) {
  return afb::adamw_ema_launch(args, static_cast<cudaStream_t>(stream));
}
int afb_adamw8bit_ema_step(const afb_adamw8bit_args* args, void* stream) {
  return afb::adamw8bit_ema_launch(args, static_cast<cudaStream_t>(stream));
}
int afb_axpy_rows(const float* x, const void* u, const float* coef, float* out, void* out_bf16, int32_t batch,
                  int64_t per_sample, int32_t u_is_f32, void* stream) {
  return afb::axpy_rows_launch(x, u, coef, out, out_bf16, batch, per_sample, u_is_f32, static_cast<cudaStream_t>(stream));
}
int afb_mse_rows(const float* pred, const void* tgt, float* out, int32_t batch, int64_t per_sample, int32_t tgt_is_f32,
                 void* stream) {
  return afb::mse_rows_launch(pred, tgt, out, batch, per_sample, tgt_is_f32, static_cast<cudaStream_t>(stream));
}
int afb_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  return afb::cast_f32_bf16_launch(in, out, n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
