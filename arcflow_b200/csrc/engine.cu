// arcflow_b200 — engine: the whole ArcFlow transformer forward and the N-NFE denoising loop as a
// fixed sequence of kernel launches on the caller's stream (CUDA-graph capturable: no host sync, no
// allocation after afb_engine_reserve).
//
// Restates, on packed weights and a joint [text; image] residual buffer:
//   FLUX  _ArcFluxTransformer2DModel.forward   lakonlab/models/architecture/arcflow/arcflux.py:134-257
//         (+ diffusers FluxTransformerBlock / FluxSingleTransformerBlock)
//   Qwen  _ArcQwenImageTransformer2DModel.forward lakonlab/models/architecture/arcflow/arcqwen.py:106-174
//   loop  ArcFluxPipeline.__call__ denoising loop  lakonlab/pipelines/arcflux_pipeline.py:453-524
//
// Data layout in HBM (all bf16 unless noted; B batch, St text tokens, Si image tokens, S = St + Si):
//   h     [B, S, D]    joint residual stream; the text stream is rows [0, St), the image stream rows
//                      [St, S) of each batch — double-stream blocks address them as strided views, the
//                      single-stream blocks use the whole buffer, so no concat/split copy ever happens
//   y     [B, S, D]    AdaLN-modulated activations (GEMM A operand)
//   qkv   [B, S, 3D]   fused Q|K|V; RMSNorm+RoPE in place; attention reads strided head views
//   attn  [B, S, D]    attention output (A operand of the out projections)
//   mlp   [B, S, M]    GELU(MLP-up) hidden
//   lt0/1 [B, S, r]    LoRA A-projections (the K-extension operands)
//   mod   [B, mod_total] every AdaLN vector of the forward, from ONE weight-streaming launch
//   head  [B*Si, head_n] raw ArcFlow heads (means | logits | loggamma), consumed by the sampler kernel
#include <stdlib.h>

#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: a no-op stub unless a profiler injects itself (nsys / ncu --nvtx)

#include "common.cuh"

namespace afb {
int gemm_launch(const afb_gemm_desc* d, cudaStream_t stream);
int attention_launch(const afb_attn_desc* d, cudaStream_t stream);
int ln_modulate_launch(const void* x, int64_t x_bs, void* y, int64_t y_bs, const void* scale,
                       const void* shift, int64_t mod_bs, int batches, int rows_per_batch, int dim,
                       float eps, cudaStream_t stream);
int rmsnorm_rope_launch(void* qkv, int64_t ld, int64_t bs, int q_off, int k_off, int batches, int seq,
                        int heads, int txt_rows, const void* wq_txt, const void* wk_txt,
                        const void* wq_img, const void* wk_img, const float* cos_tab,
                        const float* sin_tab, float eps, cudaStream_t stream);
int rope_pack_launch(const float* cos_tab, const float* sin_tab, float* out, int64_t rows, cudaStream_t stream);
int small_linear_launch(const void* x, int64_t x_ld, const void* w, int64_t w_ld, const void* bias,
                        void* y, int64_t y_ld, int m, int n, int k, int flags, cudaStream_t stream);
int timestep_embed_launch(const float* t, void* out, int m, cudaStream_t stream);
int sampler_step_launch(const void* head, int64_t head_ld, const float* x_in, float* x_out,
                        void* x_out_bf16, int64_t tokens, int num_gaussians, float sigma_src,
                        float sigma_start, float sigma_end, float eps, cudaStream_t stream);
int cast_f32_bf16_launch(const float* in, void* out, int64_t n, cudaStream_t stream);
int fill_f32_launch(float* p, float v, int n, cudaStream_t stream);
int rmsnorm_rows_launch(const void* x, void* y, const void* w, int64_t rows, int dim, float eps,
                        cudaStream_t stream);
// training path
int attention_backward_launch(const afb_attn_bwd_desc* d, cudaStream_t stream);
int gemm_tn_batched_launch(const void* a, int64_t a_ld, int64_t a_bs, const void* b, int64_t b_ld, int64_t b_bs, float* out,
                           int64_t out_ld, int batches, int64_t rows, int m, int n, cudaStream_t stream);
int ln_modulate_bwd_launch(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, void* dh, int64_t dh_bs,
                           const void* scale, int64_t mod_bs, int batches, int rows_per_batch, int dim, float eps,
                           int accumulate, cudaStream_t stream);
int rowscale_launch(const void* x, int64_t x_ld, int64_t x_bs, const void* vec, int64_t vec_bs, void* out, int64_t out_ld,
                    int64_t out_bs, int batches, int rows_per_batch, int cols, cudaStream_t stream);
int gelu_bwd_launch(void* dm, int64_t dm_ld, const void* pre, int64_t pre_ld, int64_t rows, int cols, cudaStream_t stream);
int gelu_fwd_launch(const void* pre, int64_t pre_ld, void* out, int64_t out_ld, int64_t rows, int cols, cudaStream_t stream);
int rmsnorm_rope_bwd_launch(void* dqkv, const void* raw, int64_t ld, int64_t bs, int q_off, int k_off, int batches, int seq,
                            int heads, int txt_rows, const void* wq_txt, const void* wk_txt, const void* wq_img,
                            const void* wk_img, const float* cos_tab, const float* sin_tab, float eps, cudaStream_t stream);
int gate_bwd_launch(const void* dh, int64_t dh_bs, const void* u, int64_t u_bs, const void* gate, int64_t gate_bs, void* du,
                    int64_t du_bs, float* dgate, int64_t dgate_bs, int batches, int rows_per_batch, int cols,
                    cudaStream_t stream);
int gate_res_launch(const void* res, int64_t res_bs, const void* u, int64_t u_bs, const void* gate, int64_t gate_bs, void* out,
                    int64_t out_bs, int batches, int rows_per_batch, int cols, cudaStream_t stream);
int rowlinear_dx_launch(const float* de, int64_t de_ld, const void* w, int64_t w_ld, float* out, int64_t out_ld, int m, int J,
                        int N, cudaStream_t stream);
int silu_bwd_launch(float* d, int64_t d_ld, const void* x, int64_t x_ld, int rows, int cols, cudaStream_t stream);
int scale_bf16_launch(void* x, int64_t n, float s, cudaStream_t stream);
int dropout_f32_add_launch(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int rows, int cols, uint64_t seed,
                           uint32_t layer_id, float p, cudaStream_t stream);
int dropout_rows_launch(const void* x, int64_t x_ld, int64_t x_bs, void* out, int64_t out_ld, int64_t out_bs, int batches,
                        int rows_per_batch, int cols, int logical_cols, int col0, uint64_t seed, uint32_t layer_id, float p,
                        int silu_in, int accumulate, cudaStream_t stream);
int ln_mod_param_grad_strided_launch(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, float* stats_ws,
                                     float* dscale, float* dshift, int64_t out_bs, int batches, int rows_per_batch, int dim,
                                     float eps, cudaStream_t stream);
int rowlinear_param_grad_launch(const float* de, int64_t de_ld, const void* t, int64_t t_ld, float* dw, int64_t dw_ld,
                                float* dbias, int m, int n_out, int k_in, int silu_in, cudaStream_t stream);
}  // namespace afb

using bf16 = __nv_bfloat16;

struct afb_engine {
  afb_model_desc desc{};
  afb_weights w{};
  std::vector<afb_double_block> dbl;
  std::vector<afb_single_block> sgl;
  bool bound = false;
  float lora_scale = 1.0f;
  // workspace
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int cap_batch = 0, cap_txt = 0, cap_img = 0;
  // carved pointers
  bf16 *h = nullptr, *y = nullptr, *qkv = nullptr, *attn = nullptr, *mlp = nullptr, *lt0 = nullptr,
       *lt1 = nullptr, *mod = nullptr, *temb = nullptr, *tmp = nullptr, *tproj = nullptr,
       *ltv = nullptr, *head = nullptr, *x_bf16 = nullptr, *txtn = nullptr, *alt_nm = nullptr;
  float *t_dev = nullptr, *g_dev = nullptr;
  float* rope_cs = nullptr;   // afb_rope_pack layout: (cos, sin) pairs for the fused QK-norm + RoPE epilogue
  bool fuse_qk_rope = true;   // inference forwards: RMSNorm + RoPE inside the QKV GEMM epilogue (AFB_ENGINE_FUSE_QK_ROPE=0: separate kernel)
  // training workspace (afb_engine_train_reserve): checkpoints + recompute / gradient buffers
  void* tws = nullptr;
  size_t tws_bytes = 0;
  int tcap_batch = 0, tcap_txt = 0, tcap_img = 0;
  int saved_batch = 0, saved_txt = 0, saved_img = 0;  // shape of the checkpoints currently held
  bf16 *ckpt = nullptr, *dh = nullptr, *qkv_raw = nullptr, *dqkv = nullptr, *mlp_pre = nullptr, *dmlp = nullptr,
       *dattn = nullptr, *du = nullptr, *dy = nullptr, *dl = nullptr, *h_mid = nullptr, *u1 = nullptr, *u2 = nullptr,
       *tproj_t = nullptr, *tmp_t = nullptr, *ltv1 = nullptr, *ltv2 = nullptr;
  float *lse = nullptr, *delta = nullptr, *stats = nullptr, *dsilu = nullptr, *dtmp = nullptr, *dltv = nullptr;
  // activation stash (afb_engine_set_activation_stash): per-block outputs of the train forward kept for the backward
  // instead of being recomputed from the checkpoint (a B200 has the HBM for it: ~62 GB at FLUX bs 4, 1024 px)
  bool stash_on = false;
  void* stash = nullptr;
  size_t stash_bytes = 0;
  int scap_batch = 0, scap_txt = 0, scap_img = 0;
  bool stash_valid = false;  // the stash holds the activations of the checkpoints currently saved
  // LoRA input dropout (train forward / backward only): p, seed, scratch for the dropped LoRA-branch inputs
  float drop_p = 0.f;
  uint64_t drop_seed = 0;
  bf16 *xd = nullptr, *xd_small = nullptr;
  float* dtmp2 = nullptr;
  // optional per-launch CUDA-event profiling of the two tensor-core kernels
  bool profiling = false;
  struct ProfRec { int cls; double flops; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  afb_profile prof_acc{};
};

namespace {

constexpr float LN_EPS = 1e-6f;

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

struct Carver {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t elems) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align_up(elems * sizeof(T));
    return p;
  }
};

size_t carve(afb_engine* e, uint8_t* base, int B, int St, int Si) {
  const afb_model_desc& d = e->desc;
  const size_t S = size_t(St) + Si, D = d.dim, M = d.mlp_dim, r = d.lora_rank > 0 ? d.lora_rank : 8;
  Carver c{base};
  e->h = c.take<bf16>(B * S * D);
  e->y = c.take<bf16>(B * S * D);
  e->qkv = c.take<bf16>(B * S * 3 * D);
  e->attn = c.take<bf16>(B * S * D);
  e->mlp = c.take<bf16>(B * S * M);
  e->lt0 = c.take<bf16>(B * S * r);
  e->lt1 = c.take<bf16>(B * S * r);
  e->mod = c.take<bf16>(size_t(B) * size_t(e->w.mod_total > 0 ? e->w.mod_total : 1));
  e->temb = c.take<bf16>(B * D);
  e->tmp = c.take<bf16>(B * D);
  e->tproj = c.take<bf16>(size_t(B) * 256);
  e->ltv = c.take<bf16>(B * r);
  e->head = c.take<bf16>(size_t(B) * Si * size_t(e->w.head_n > 0 ? e->w.head_n : 8));
  e->x_bf16 = c.take<bf16>(size_t(B) * Si * d.in_channels);
  e->txtn = c.take<bf16>(d.arch == AFB_ARCH_QWEN ? size_t(B) * St * d.txt_dim : 8);
  e->alt_nm = c.take<bf16>(size_t(B) * 2 * D);
  e->t_dev = c.take<float>(B);
  e->g_dev = c.take<float>(B);
  e->rope_cs = c.take<float>((S + 31) / 32 * 32 * 128);
  return c.off;
}

size_t carve_train(afb_engine* e, uint8_t* base, int B, int St, int Si) {
  const afb_model_desc& d = e->desc;
  const size_t S = size_t(St) + Si, D = d.dim, M = d.mlp_dim, r = d.lora_rank > 0 ? d.lora_rank : 8;
  const size_t blocks = size_t(d.num_double) + d.num_single;
  Carver c{base};
  e->ckpt = c.take<bf16>((blocks + 1) * B * S * D);
  e->dh = c.take<bf16>(B * S * D);
  e->qkv_raw = c.take<bf16>(B * S * 3 * D);
  e->dqkv = c.take<bf16>(B * S * 3 * D);
  e->mlp_pre = c.take<bf16>(B * S * M);
  e->dmlp = c.take<bf16>(B * S * M);
  e->dattn = c.take<bf16>(B * S * D);
  e->du = c.take<bf16>(B * S * D);
  e->dy = c.take<bf16>(B * S * D);
  e->dl = c.take<bf16>(B * S * r);
  e->h_mid = c.take<bf16>(B * S * D);
  e->lse = c.take<float>(size_t(B) * d.heads * S);
  e->delta = c.take<float>(size_t(B) * d.heads * S);
  // modulation-gradient path
  e->u1 = c.take<bf16>(B * S * D);
  e->u2 = c.take<bf16>(B * S * D);
  e->stats = c.take<float>(size_t(2) * B * S);
  e->tproj_t = c.take<bf16>(size_t(B) * 256);
  e->tmp_t = c.take<bf16>(B * D);
  e->ltv1 = c.take<bf16>(B * r);
  e->ltv2 = c.take<bf16>(B * r);
  e->dsilu = c.take<float>(B * D);
  e->dtmp = c.take<float>(B * D);
  e->dltv = c.take<float>(B * r);
  e->xd = c.take<bf16>(B * S * (D + M));
  e->xd_small = c.take<bf16>(B * D);
  e->dtmp2 = c.take<float>(B * D);
  return c.off;
}

// One block's slice of the activation stash: raw QKV, attention output + log-sum-exp, MLP pre-activation, the un-gated
// branch outputs (u1: attention / single-block branch, u2: MLP branch) and, for double-stream blocks, the residual
// stream between the two halves.
struct StashSlot {
  bf16 *qkv_raw, *attn, *pre, *u1, *u2, *h_mid;
  float* lse;
};
size_t stash_slot_elems(const afb_model_desc& d, bool dbl, size_t tokens) {
  const size_t D = d.dim, M = d.mlp_dim;
  return tokens * (3 * D + D + M + D + (dbl ? 2 * D : 0));
}
size_t stash_total_bytes(const afb_model_desc& d, int B, int St, int Si) {
  const size_t tokens = size_t(B) * (size_t(St) + Si);
  const size_t lse = align_up(size_t(B) * d.heads * (size_t(St) + Si) * sizeof(float));
  return size_t(d.num_double) * (align_up(stash_slot_elems(d, true, tokens) * sizeof(bf16)) + lse) +
         size_t(d.num_single) * (align_up(stash_slot_elems(d, false, tokens) * sizeof(bf16)) + lse);
}
StashSlot stash_slot(const afb_engine* e, int block, int B, int St, int Si) {
  const afb_model_desc& d = e->desc;
  const size_t tokens = size_t(B) * (size_t(St) + Si), D = d.dim, M = d.mlp_dim;
  const size_t lse = align_up(size_t(B) * d.heads * (size_t(St) + Si) * sizeof(float));
  const size_t dslot = align_up(stash_slot_elems(d, true, tokens) * sizeof(bf16)) + lse;
  const size_t sslot = align_up(stash_slot_elems(d, false, tokens) * sizeof(bf16)) + lse;
  const bool dbl = block < d.num_double;
  uint8_t* base = static_cast<uint8_t*>(e->stash) +
                  (dbl ? size_t(block) * dslot : size_t(d.num_double) * dslot + size_t(block - d.num_double) * sslot);
  StashSlot t{};
  bf16* p = reinterpret_cast<bf16*>(base);
  t.qkv_raw = p, p += tokens * 3 * D;
  t.attn = p, p += tokens * D;
  t.pre = p, p += tokens * M;
  t.u1 = p, p += tokens * D;
  if (dbl) {
    t.u2 = p, p += tokens * D;
    t.h_mid = p, p += tokens * D;
  }
  t.lse = reinterpret_cast<float*>(base + (dbl ? dslot : sslot) - lse);
  return t;
}
int copy_bf16(bf16* dst, const bf16* src, size_t elems, cudaStream_t s) {
  AFB_CHECK_CUDA(cudaMemcpyAsync(dst, src, elems * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
  return AFB_OK;
}

// [batches, rows, cols] view helper for the GEMM descriptor
struct View {
  const bf16* p;
  int64_t ld, bs;
};

cudaEvent_t prof_event(afb_engine* e) {
  cudaEvent_t ev;
  if (!e->ev_pool.empty()) {
    ev = e->ev_pool.back();
    e->ev_pool.pop_back();
  } else {
    cudaEventCreate(&ev);
  }
  return ev;
}

struct ProfScope {
  afb_engine* e;
  cudaStream_t s;
  int cls;
  double flops;
  cudaEvent_t e0{};
  ProfScope(afb_engine* eng, cudaStream_t st, int c, double f) : e(eng), s(st), cls(c), flops(f) {
    if (e && e->profiling) {
      e0 = prof_event(e);
      cudaEventRecord(e0, s);
    }
  }
  ~ProfScope() {
    if (e && e->profiling) {
      cudaEvent_t e1 = prof_event(e);
      cudaEventRecord(e1, s);
      e->prof.push_back({cls, flops, e0, e1});
    }
  }
};

// NVTX push/pop range: names the forward / backward, every transformer block and the launch kinds inside it,
// so profiler timelines and `ncu --nvtx --nvtx-include "afb_forward/single_block_3/"` filters address launches by name.
struct NvtxRange {
  bool open = true;
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  NvtxRange(const char* fmt, int i) {
    char buf[48];
    snprintf(buf, sizeof(buf), fmt, i);
    nvtxRangePushA(buf);
  }
  void end() {
    if (open) nvtxRangePop();
    open = false;
  }
  ~NvtxRange() { end(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct Gemm {
  afb_gemm_desc d{};
  int nseg = 0;
  Gemm(int batches, int rows) {
    d.batches = batches;
    d.rows_per_batch = rows;
  }
  Gemm& a(View v, int k) {
    d.a[nseg] = v.p;
    d.a_ld[nseg] = v.ld;
    d.a_batch_stride[nseg] = v.bs;
    d.a_k[nseg] = k;
    ++nseg;
    return *this;
  }
  Gemm& w(const void* wp, int64_t ld, int n, const void* bias) {
    d.w = wp;
    d.w_ld = ld;
    d.n = n;
    d.bias = bias;
    return *this;
  }
  // W read as [K, n] row-major (dX = dY W); the K rows past k0 come from `second` (leading dim ld2) when given
  Gemm& wt(const void* wp, int64_t ld, int n, int k0, const void* second = nullptr, int64_t ld2 = 0) {
    d.w = wp;
    d.w_ld = ld;
    d.n = n;
    d.w_transposed = 1;
    d.w_k = k0;
    d.w2 = second;
    d.w2_ld = ld2;
    return *this;
  }
  Gemm& alpha(float a) {
    d.alpha = a;
    return *this;
  }
  Gemm& res(View v) {
    d.res = v.p;
    d.res_ld = v.ld;
    d.res_batch_stride = v.bs;
    return *this;
  }
  Gemm& out(View v, int epi) {
    d.out = const_cast<bf16*>(v.p);
    d.out_ld = v.ld;
    d.out_batch_stride = v.bs;
    d.epilogue = epi;
    return *this;
  }
  // fused per-head RMSNorm + RoPE on the first qk_cols output columns (q heads then k heads)
  Gemm& qk_norm_rope(const void* nq, const void* nk, const float* rope, int row0, int qk_cols, float eps) {
    d.epilogue = AFB_EPI_BIAS_QKNORM_ROPE;
    d.norm_q = nq;
    d.norm_k = nk;
    d.rope = rope;
    d.rope_row0 = row0;
    d.qk_cols = qk_cols;
    d.norm_eps = eps;
    return *this;
  }
  Gemm& gate_res(const bf16* gate, int64_t gate_bs, View res) {
    d.gate = gate;
    d.gate_batch_stride = gate_bs;
    d.res = res.p;
    d.res_ld = res.ld;
    d.res_batch_stride = res.bs;
    return *this;
  }
  int run(afb_engine* e, cudaStream_t s) {
    double k = 0;
    for (int i = 0; i < nseg; ++i) k += d.a_k[i];
    ProfScope ps(e, s, d.epilogue == AFB_EPI_BIAS_QKNORM_ROPE ? 2 : 0, 2.0 * d.batches * d.rows_per_batch * double(d.n) * k);
    NvtxRange nv(d.w_transposed ? "gemm_dx" : (d.n <= 256 ? "gemm_lora_a" : "gemm"));
    int rc = afb::gemm_launch(&d, s);
    if (rc == AFB_OK) afb::count_launch(1);
    return rc;
  }
};

int run_attention(afb_engine* e, const afb_attn_desc* at, cudaStream_t s) {
  ProfScope ps(e, s, 1, 4.0 * at->batch * at->heads * double(at->seq) * at->seq * 128.0);
  NvtxRange nv("attention");
  return afb::attention_launch(at, s);
}

// LoRA layer ids for the dropout mask stream: 4 slots per block (double: img up/down, txt up/down; single: mlp, out),
// double blocks first; the timestep embedder's two Linears use ids past any block.
constexpr uint32_t DROP_ID_T1 = 0xFFFF0u, DROP_ID_T2 = 0xFFFF1u;

#define AFB_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != AFB_OK) return _rc; \
  } while (0)

// y[m, n] = act(x) W^T + b for any m (loops the <= 8-row kernel)
int small_linear_rows(const bf16* x, int64_t x_ld, const void* w, int64_t w_ld, const void* bias,
                      bf16* y, int64_t y_ld, int m, int n, int k, int flags, cudaStream_t s) {
  for (int r0 = 0; r0 < m; r0 += 8) {
    const int mm = m - r0 < 8 ? m - r0 : 8;
    AFB_TRY(afb::small_linear_launch(x + (int64_t)r0 * x_ld, x_ld, w, w_ld, bias,
                                     y + (int64_t)r0 * y_ld, y_ld, mm, n, k, flags, s));
  }
  return AFB_OK;
}

// Linear(256 -> D) -> SiLU -> Linear(D -> D), optionally with un-merged LoRA on both, result
// written (or accumulated) into e->temb.
int embed_mlp(afb_engine* e, const bf16* in, int in_dim, const void* w1, const void* b1,
              const void* la1, const void* lb1, const void* w2, const void* b2, const void* la2,
              const void* lb2, bool accumulate, int B, cudaStream_t s, bool drop = false) {
  // drop: the LoRA branches see dropout(x) (train forward of the timestep embedder; mask ids DROP_ID_T1 / T2)
  const int D = e->desc.dim, r = e->desc.lora_rank;
  AFB_TRY(small_linear_rows(in, in_dim, w1, in_dim, b1, e->tmp, D, B, D, in_dim, 0, s));
  if (la1 && lb1) {
    const bf16* xin = in;
    if (drop) {
      AFB_TRY(afb::dropout_rows_launch(in, in_dim, int64_t(B) * in_dim, e->xd_small, in_dim, int64_t(B) * in_dim, 1, B, in_dim,
                                       in_dim, 0, e->drop_seed, DROP_ID_T1, e->drop_p, 0, 0, s));
      xin = e->xd_small;
    }
    AFB_TRY(small_linear_rows(xin, in_dim, la1, in_dim, nullptr, e->ltv, r, B, r, in_dim, 0, s));
    if (e->lora_scale != 1.0f) AFB_TRY(afb::scale_bf16_launch(e->ltv, int64_t(B) * r, e->lora_scale, s));
    AFB_TRY(small_linear_rows(e->ltv, r, lb1, r, nullptr, e->tmp, D, B, D, r, AFB_SL_ACCUMULATE, s));
  }
  AFB_TRY(small_linear_rows(e->tmp, D, w2, D, b2, e->temb, D, B, D, D,
                            AFB_SL_SILU_IN | (accumulate ? AFB_SL_ACCUMULATE : 0), s));
  if (la2 && lb2) {
    if (drop) {
      AFB_TRY(afb::dropout_rows_launch(e->tmp, D, int64_t(B) * D, e->xd_small, D, int64_t(B) * D, 1, B, D, D, 0, e->drop_seed,
                                       DROP_ID_T2, e->drop_p, 1, 0, s));
      AFB_TRY(small_linear_rows(e->xd_small, D, la2, D, nullptr, e->ltv, r, B, r, D, 0, s));
    } else {
      AFB_TRY(small_linear_rows(e->tmp, D, la2, D, nullptr, e->ltv, r, B, r, D, AFB_SL_SILU_IN, s));
    }
    if (e->lora_scale != 1.0f) AFB_TRY(afb::scale_bf16_launch(e->ltv, int64_t(B) * r, e->lora_scale, s));
    AFB_TRY(small_linear_rows(e->ltv, r, lb2, r, nullptr, e->temb, D, B, D, r, AFB_SL_ACCUMULATE, s));
  }
  return AFB_OK;
}

// dropout(x) of one LoRA branch input into the contiguous scratch e->xd, viewed [B, rows, logical_cols]; x may be one
// column slice (col0, cols) of the logical input. Returns the scratch view of the whole logical tensor.
View dropped_view(afb_engine* e, int rows, int logical_cols) {
  return View{e->xd, logical_cols, int64_t(rows) * logical_cols};
}
int drop_into(afb_engine* e, View x, int B, int rows, int cols, int logical_cols, int col0, uint32_t layer, cudaStream_t s) {
  return afb::dropout_rows_launch(x.p, x.ld, x.bs, e->xd + col0, logical_cols, int64_t(rows) * logical_cols, B, rows, cols,
                                  logical_cols, col0, e->drop_seed, layer, e->drop_p, 0, 0, s);
}

// One MLP (Linear up + GELU-tanh + Linear down, both with optional LoRA K-extension), gated into res.
int mlp_branch(afb_engine* e, View yv, View hv, View mlpv, View lt0v, View lt1v, int rows, int B,
               const void* up_w, const void* up_b, const void* up_la, const void* down_w,
               const void* down_b, const void* down_la, const bf16* gate, cudaStream_t s, bool drop = false,
               uint32_t layer_up = 0, const View* pre_v = nullptr, const View* u2_v = nullptr) {
  // pre_v / u2_v (train forward with the activation stash): keep the MLP pre-activation and the un-gated branch output
  // instead of fusing GELU and gate * y + residual into the GEMM epilogues.
  // `*_la != NULL` means the packed weight is [W | lora_B] (leading dim in + rank). A teacher engine
  // (ignore_lora) shares those buffers with the student and simply never reads the extra K columns.
  const int D = e->desc.dim, M = e->desc.mlp_dim, r = e->desc.lora_rank;
  const bool lora = !e->desc.ignore_lora;
  const int64_t mod_bs = e->w.mod_total;
  const int64_t up_ld = D + (up_la ? r : 0), down_ld = M + (down_la ? r : 0);
  if (up_la && lora) {
    View xa = yv;
    if (drop) {
      AFB_TRY(drop_into(e, yv, B, rows, D, D, 0, layer_up, s));
      xa = dropped_view(e, rows, D);
    }
    AFB_TRY(Gemm(B, rows).a(xa, D).w(up_la, D, r, nullptr).alpha(e->lora_scale).out(lt0v, AFB_EPI_BIAS).run(e, s));
    AFB_TRY(Gemm(B, rows).a(yv, D).a(lt0v, r).w(up_w, up_ld, M, up_b)
                .out(pre_v ? *pre_v : mlpv, pre_v ? AFB_EPI_BIAS : AFB_EPI_BIAS_GELU).run(e, s));
  } else {
    AFB_TRY(Gemm(B, rows).a(yv, D).w(up_w, up_ld, M, up_b)
                .out(pre_v ? *pre_v : mlpv, pre_v ? AFB_EPI_BIAS : AFB_EPI_BIAS_GELU).run(e, s));
  }
  if (pre_v)
    for (int bi = 0; bi < B; ++bi)
      AFB_TRY(afb::gelu_fwd_launch(pre_v->p + int64_t(bi) * pre_v->bs, M, const_cast<bf16*>(mlpv.p) + int64_t(bi) * mlpv.bs, M,
                                   rows, M, s));
  const View down_out = u2_v ? *u2_v : hv;
  const int down_epi = u2_v ? AFB_EPI_BIAS : AFB_EPI_BIAS_GATE_RES;
  if (down_la && lora) {
    View xa = mlpv;
    if (drop) {
      AFB_TRY(drop_into(e, mlpv, B, rows, M, M, 0, layer_up + 1, s));
      xa = dropped_view(e, rows, M);
    }
    AFB_TRY(Gemm(B, rows).a(xa, M).w(down_la, M, r, nullptr).alpha(e->lora_scale).out(lt1v, AFB_EPI_BIAS).run(e, s));
    AFB_TRY(Gemm(B, rows).a(mlpv, M).a(lt1v, r).w(down_w, down_ld, D, down_b)
                .out(down_out, down_epi).gate_res(gate, mod_bs, hv).run(e, s));
  } else {
    AFB_TRY(Gemm(B, rows).a(mlpv, M).w(down_w, down_ld, D, down_b)
                .out(down_out, down_epi).gate_res(gate, mod_bs, hv).run(e, s));
  }
  if (u2_v)
    AFB_TRY(afb::gate_res_launch(hv.p, hv.bs, u2_v->p, u2_v->bs, gate, mod_bs, const_cast<bf16*>(hv.p), hv.bs, B, rows, D, s));
  return AFB_OK;
}

int forward_impl(afb_engine* e, const afb_forward_args* a, const bf16* latents, bf16* head_out,
                 cudaStream_t s, bool save_ckpt = false) {
  const afb_model_desc& d = e->desc;
  const afb_weights& w = e->w;
  const int B = a->batch, St = a->txt_len, Si = a->img_len, S = St + Si;
  const int D = d.dim, M = d.mlp_dim, H = d.heads;
  const int rpad = d.lora_rank;                    // K columns appended to LoRA-carrying packed weights
  const int r = d.ignore_lora ? 0 : d.lora_rank;   // rank actually computed (0: frozen trunk / teacher)
  const int64_t mod_bs = w.mod_total;
  const bool flux = d.arch == AFB_ARCH_FLUX;
  const bool drop = save_ckpt && e->drop_p > 0.f && r > 0;  // LoRA input dropout: train forward only
  const bool stash = save_ckpt && e->stash_on && e->stash != nullptr;  // keep block outputs for the backward
  const size_t tok = size_t(B) * S;
  NvtxRange nvtx_fwd(save_ckpt ? "afb_forward_train" : "afb_forward");
  // Inference forwards normalise + rotate q and k inside the QKV GEMM epilogue (no separate pass over the QKV buffer); the
  // train forward keeps the stand-alone kernel: the backward needs the raw projections.
  const bool fuse_qk = e->fuse_qk_rope && !save_ckpt && D % 256 == 0;
  if (fuse_qk) AFB_TRY(afb::rope_pack_launch(a->rope_cos, a->rope_sin, e->rope_cs, S, s));

  // ---- conditioning vector temb [B, D] ---------------------------------------------------------
  NvtxRange nvtx_embed("embedders");
  AFB_TRY(afb::timestep_embed_launch(a->timestep, e->tproj, B, s));
  AFB_TRY(embed_mlp(e, e->tproj, 256, w.t1_w, w.t1_b, r > 0 ? w.t1_la : nullptr, r > 0 ? w.t1_lb : nullptr,
                    w.t2_w, w.t2_b, r > 0 ? w.t2_la : nullptr, r > 0 ? w.t2_lb : nullptr, false, B, s, drop));
  if (flux && d.guidance) {
    AFB_REQUIRE(a->guidance != nullptr, "forward: model has guidance embeds but guidance is NULL");
    AFB_TRY(afb::timestep_embed_launch(a->guidance, e->tproj, B, s));
    AFB_TRY(embed_mlp(e, e->tproj, 256, w.g1_w, w.g1_b, nullptr, nullptr, w.g2_w, w.g2_b, nullptr,
                      nullptr, true, B, s));
  }
  if (flux && d.pooled_dim > 0) {
    AFB_REQUIRE(a->pooled != nullptr, "forward: pooled projections missing");
    AFB_TRY(embed_mlp(e, static_cast<const bf16*>(a->pooled), d.pooled_dim, w.p1_w, w.p1_b, nullptr,
                      nullptr, w.p2_w, w.p2_b, nullptr, nullptr, true, B, s));
  }
  // ---- every AdaLN modulation vector of this forward: one weight-streaming launch ---------------
  AFB_TRY(small_linear_rows(e->temb, D, w.mod_w, D, w.mod_b, e->mod, mod_bs, B, int(w.mod_total), D,
                            AFB_SL_SILU_IN, s));

  // ---- token embedders into the joint residual buffer ------------------------------------------
  const View h_txt{e->h, D, int64_t(S) * D};
  const View h_img{e->h + int64_t(St) * D, D, int64_t(S) * D};
  const View y_txt{e->y, D, int64_t(S) * D};
  const View y_img{e->y + int64_t(St) * D, D, int64_t(S) * D};
  const View qkv_txt{e->qkv, 3 * D, int64_t(S) * 3 * D};
  const View qkv_img{e->qkv + int64_t(St) * 3 * D, 3 * D, int64_t(S) * 3 * D};
  View at_txt{e->attn, D, int64_t(S) * D};
  View at_img{e->attn + int64_t(St) * D, D, int64_t(S) * D};
  const View mlp_txt{e->mlp, M, int64_t(S) * M};
  const View mlp_img{e->mlp + int64_t(St) * M, M, int64_t(S) * M};
  const int rr = r > 0 ? r : 8;
  const View l0_txt{e->lt0, rr, int64_t(S) * rr};
  const View l0_img{e->lt0 + int64_t(St) * rr, rr, int64_t(S) * rr};
  const View l1_txt{e->lt1, rr, int64_t(S) * rr};
  const View l1_img{e->lt1 + int64_t(St) * rr, rr, int64_t(S) * rr};

  AFB_TRY(Gemm(B, Si).a(View{latents, d.in_channels, int64_t(Si) * d.in_channels}, d.in_channels)
              .w(w.x_emb_w, d.in_channels, D, w.x_emb_b).out(h_img, AFB_EPI_BIAS).run(e, s));
  {
    const bf16* txt = static_cast<const bf16*>(a->txt);
    if (!flux) {
      AFB_TRY(afb::rmsnorm_rows_launch(txt, e->txtn, w.txt_norm_w, int64_t(B) * St, d.txt_dim, LN_EPS, s));
      txt = e->txtn;
    }
    AFB_TRY(Gemm(B, St).a(View{txt, d.txt_dim, int64_t(St) * d.txt_dim}, d.txt_dim)
                .w(w.ctx_w, d.txt_dim, D, w.ctx_b).out(h_txt, AFB_EPI_BIAS).run(e, s));
  }

  afb_attn_desc at{};
  at.q = e->qkv;
  at.k = e->qkv + D;
  at.v = e->qkv + 2 * D;
  at.o = e->attn;
  at.q_ld = at.k_ld = at.v_ld = 3 * D;
  at.o_ld = D;
  at.q_batch_stride = at.k_batch_stride = at.v_batch_stride = int64_t(S) * 3 * D;
  at.o_batch_stride = int64_t(S) * D;
  at.batch = B;
  at.seq = S;
  at.heads = H;
  at.scale = 0.f;

  View u1_txt{}, u1_img{}, u2_txt{}, u2_img{}, pre_txt{}, pre_img{}, at_all{};
  // With the activation stash the GEMMs and the attention kernel write a block's kept outputs straight into its stash
  // slice: the engine's buffer pointers are re-pointed per block and the views rebuilt (every entry point re-carves).
  bf16 *const ws_attn = e->attn, *const ws_pre = e->mlp_pre, *const ws_u1 = e->u1, *const ws_u2 = e->u2;
  auto retarget = [&](const StashSlot* sl) {
    e->attn = sl ? sl->attn : ws_attn;
    e->mlp_pre = sl ? sl->pre : ws_pre;
    e->u1 = sl ? sl->u1 : ws_u1;
    e->u2 = sl && sl->u2 ? sl->u2 : ws_u2;
    at_txt = View{e->attn, D, int64_t(S) * D};
    at_img = View{e->attn + int64_t(St) * D, D, int64_t(S) * D};
    at_all = at_txt;
    at.o = e->attn;
    if (!e->u1) return;  // inference-only engine: no training workspace, the views below are never used
    u1_txt = View{e->u1, D, int64_t(S) * D};
    u1_img = View{e->u1 + int64_t(St) * D, D, int64_t(S) * D};
    u2_txt = View{e->u2, D, int64_t(S) * D};
    u2_img = View{e->u2 + int64_t(St) * D, D, int64_t(S) * D};
    pre_txt = View{e->mlp_pre, M, int64_t(S) * M};
    pre_img = View{e->mlp_pre + int64_t(St) * M, M, int64_t(S) * M};
  };
  retarget(nullptr);
  const size_t ckpt_elems = size_t(B) * S * D;
  auto save = [&](int idx) -> int {
    if (!save_ckpt) return AFB_OK;
    AFB_CHECK_CUDA(cudaMemcpyAsync(e->ckpt + size_t(idx) * ckpt_elems, e->h, ckpt_elems * sizeof(bf16),
                                   cudaMemcpyDeviceToDevice, s));
    return AFB_OK;
  };

  nvtx_embed.end();
  // ---- double-stream blocks --------------------------------------------------------------------
  for (int i = 0; i < d.num_double; ++i) {
    NvtxRange nvtx_block("double_block_%d", i);
    const afb_double_block& k = e->dbl[i];
    AFB_TRY(save(i));
    // chunk(6): shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    const bf16* im = e->mod + k.img_mod_off;
    const bf16* tm = e->mod + k.txt_mod_off;
    AFB_TRY(afb::ln_modulate_launch(h_img.p, h_img.bs, const_cast<bf16*>(y_img.p), y_img.bs, im + D, im,
                                    mod_bs, B, Si, D, LN_EPS, s));
    AFB_TRY(afb::ln_modulate_launch(h_txt.p, h_txt.bs, const_cast<bf16*>(y_txt.p), y_txt.bs, tm + D, tm,
                                    mod_bs, B, St, D, LN_EPS, s));
    if (fuse_qk) {
      AFB_TRY(Gemm(B, Si).a(y_img, D).w(k.img_qkv_w, D, 3 * D, k.img_qkv_b).out(qkv_img, AFB_EPI_BIAS)
                  .qk_norm_rope(k.img_nq, k.img_nk, e->rope_cs, St, 2 * D, LN_EPS).run(e, s));
      AFB_TRY(Gemm(B, St).a(y_txt, D).w(k.txt_qkv_w, D, 3 * D, k.txt_qkv_b).out(qkv_txt, AFB_EPI_BIAS)
                  .qk_norm_rope(k.txt_nq, k.txt_nk, e->rope_cs, 0, 2 * D, LN_EPS).run(e, s));
    } else {
      AFB_TRY(Gemm(B, Si).a(y_img, D).w(k.img_qkv_w, D, 3 * D, k.img_qkv_b).out(qkv_img, AFB_EPI_BIAS).run(e, s));
      AFB_TRY(Gemm(B, St).a(y_txt, D).w(k.txt_qkv_w, D, 3 * D, k.txt_qkv_b).out(qkv_txt, AFB_EPI_BIAS).run(e, s));
    }
    StashSlot sl{};
    if (stash) {
      sl = stash_slot(e, i, B, St, Si);
      retarget(&sl);
      AFB_TRY(copy_bf16(sl.qkv_raw, e->qkv, tok * 3 * D, s));
      at.lse = sl.lse;
    }
    if (!fuse_qk)
      AFB_TRY(afb::rmsnorm_rope_launch(e->qkv, 3 * D, int64_t(S) * 3 * D, 0, D, B, S, H, St, k.txt_nq,
                                       k.txt_nk, k.img_nq, k.img_nk, a->rope_cos, a->rope_sin, LN_EPS, s));
    at.score_bound = k.qk_bound;
    AFB_TRY(run_attention(e, &at, s));
    const bool last_qwen_txt = !flux && i == d.num_double - 1;  // its text stream output is never read
    if (stash) {  // un-fused: the un-gated branch output and the mid-block residual stream are kept
      AFB_TRY(Gemm(B, Si).a(at_img, D).w(k.img_out_w, D, D, k.img_out_b).out(u1_img, AFB_EPI_BIAS).run(e, s));
      AFB_TRY(afb::gate_res_launch(h_img.p, h_img.bs, u1_img.p, u1_img.bs, im + 2 * D, mod_bs, const_cast<bf16*>(h_img.p),
                                   h_img.bs, B, Si, D, s));
      if (!last_qwen_txt) {
        AFB_TRY(Gemm(B, St).a(at_txt, D).w(k.txt_out_w, D, D, k.txt_out_b).out(u1_txt, AFB_EPI_BIAS).run(e, s));
        AFB_TRY(afb::gate_res_launch(h_txt.p, h_txt.bs, u1_txt.p, u1_txt.bs, tm + 2 * D, mod_bs, const_cast<bf16*>(h_txt.p),
                                     h_txt.bs, B, St, D, s));
      }
      AFB_TRY(copy_bf16(sl.h_mid, e->h, tok * D, s));
    } else {
      AFB_TRY(Gemm(B, Si).a(at_img, D).w(k.img_out_w, D, D, k.img_out_b)
                  .out(h_img, AFB_EPI_BIAS_GATE_RES).gate_res(im + 2 * D, mod_bs, h_img).run(e, s));
      if (!last_qwen_txt)
        AFB_TRY(Gemm(B, St).a(at_txt, D).w(k.txt_out_w, D, D, k.txt_out_b)
                    .out(h_txt, AFB_EPI_BIAS_GATE_RES).gate_res(tm + 2 * D, mod_bs, h_txt).run(e, s));
    }
    AFB_TRY(afb::ln_modulate_launch(h_img.p, h_img.bs, const_cast<bf16*>(y_img.p), y_img.bs, im + 4 * D,
                                    im + 3 * D, mod_bs, B, Si, D, LN_EPS, s));
    AFB_TRY(mlp_branch(e, y_img, h_img, mlp_img, l0_img, l1_img, Si, B, k.img_up_w, k.img_up_b,
                       k.img_up_la, k.img_down_w, k.img_down_b,
                       k.img_down_la, im + 5 * D, s, drop, 4u * i, stash ? &pre_img : nullptr, stash ? &u2_img : nullptr));
    if (!last_qwen_txt) {
      AFB_TRY(afb::ln_modulate_launch(h_txt.p, h_txt.bs, const_cast<bf16*>(y_txt.p), y_txt.bs, tm + 4 * D,
                                      tm + 3 * D, mod_bs, B, St, D, LN_EPS, s));
      AFB_TRY(mlp_branch(e, y_txt, h_txt, mlp_txt, l0_txt, l1_txt, St, B, k.txt_up_w, k.txt_up_b,
                         k.txt_up_la, k.txt_down_w, k.txt_down_b,
                         k.txt_down_la, tm + 5 * D, s, drop, 4u * i + 2, stash ? &pre_txt : nullptr,
                         stash ? &u2_txt : nullptr));
    }
  }

  // ---- single-stream blocks (FLUX) on the joint buffer -----------------------------------------
  const View h_all{e->h, D, int64_t(S) * D};
  const View y_all{e->y, D, int64_t(S) * D};
  const View qkv_all{e->qkv, 3 * D, int64_t(S) * 3 * D};
  const View mlp_all{e->mlp, M, int64_t(S) * M};
  const View l0_all{e->lt0, rr, int64_t(S) * rr};
  const View l1_all{e->lt1, rr, int64_t(S) * rr};
  for (int i = 0; i < d.num_single; ++i) {
    NvtxRange nvtx_block("single_block_%d", i);
    const afb_single_block& k = e->sgl[i];
    const bf16* m = e->mod + k.mod_off;  // chunk(3): shift, scale, gate
    AFB_TRY(save(d.num_double + i));
    AFB_TRY(afb::ln_modulate_launch(e->h, int64_t(S) * D, e->y, int64_t(S) * D, m + D, m, mod_bs, B, S, D,
                                    LN_EPS, s));
    if (fuse_qk)
      AFB_TRY(Gemm(B, S).a(y_all, D).w(k.qkv_w, D, 3 * D, k.qkv_b).out(qkv_all, AFB_EPI_BIAS)
                  .qk_norm_rope(k.nq, k.nk, e->rope_cs, 0, 2 * D, LN_EPS).run(e, s));
    else
      AFB_TRY(Gemm(B, S).a(y_all, D).w(k.qkv_w, D, 3 * D, k.qkv_b).out(qkv_all, AFB_EPI_BIAS).run(e, s));
    StashSlot sl{};
    if (stash) {
      sl = stash_slot(e, d.num_double + i, B, St, Si);
      retarget(&sl);
      AFB_TRY(copy_bf16(sl.qkv_raw, e->qkv, tok * 3 * D, s));
      at.lse = sl.lse;
    }
    if (!fuse_qk)
      AFB_TRY(afb::rmsnorm_rope_launch(e->qkv, 3 * D, int64_t(S) * 3 * D, 0, D, B, S, H, 0, nullptr, nullptr,
                                       k.nq, k.nk, a->rope_cos, a->rope_sin, LN_EPS, s));
    at.score_bound = k.qk_bound;
    AFB_TRY(run_attention(e, &at, s));
    const View up_out = stash ? View{e->mlp_pre, M, int64_t(S) * M} : mlp_all;
    const int up_epi = stash ? AFB_EPI_BIAS : AFB_EPI_BIAS_GELU;
    const View proj_out_v = stash ? View{e->u1, D, int64_t(S) * D} : h_all;
    const int proj_epi = stash ? AFB_EPI_BIAS : AFB_EPI_BIAS_GATE_RES;
    const int64_t mlp_ld = D + (k.mlp_la ? rpad : 0), out_ld = D + M + (k.out_la ? rpad : 0);
    const uint32_t layer = 4u * (d.num_double + i);
    if (r > 0 && k.mlp_la) {
      View xa = y_all;
      if (drop) {
        AFB_TRY(drop_into(e, y_all, B, S, D, D, 0, layer, s));
        xa = dropped_view(e, S, D);
      }
      AFB_TRY(Gemm(B, S).a(xa, D).w(k.mlp_la, D, r, nullptr).alpha(e->lora_scale).out(l0_all, AFB_EPI_BIAS).run(e, s));
      AFB_TRY(Gemm(B, S).a(y_all, D).a(l0_all, r).w(k.mlp_w, mlp_ld, M, k.mlp_b).out(up_out, up_epi).run(e, s));
    } else {
      AFB_TRY(Gemm(B, S).a(y_all, D).w(k.mlp_w, mlp_ld, M, k.mlp_b).out(up_out, up_epi).run(e, s));
    }
    if (stash) AFB_TRY(afb::gelu_fwd_launch(e->mlp_pre, M, e->mlp, M, int64_t(B) * S, M, s));
    if (r > 0 && k.out_la) {
      if (drop) {
        AFB_TRY(drop_into(e, at_all, B, S, D, D + M, 0, layer + 1, s));
        AFB_TRY(drop_into(e, mlp_all, B, S, M, D + M, D, layer + 1, s));
        AFB_TRY(Gemm(B, S).a(dropped_view(e, S, D + M), D + M).w(k.out_la, D + M, r, nullptr).alpha(e->lora_scale).out(l1_all, AFB_EPI_BIAS).run(e, s));
      } else {
        AFB_TRY(Gemm(B, S).a(at_all, D).a(mlp_all, M).w(k.out_la, D + M, r, nullptr).alpha(e->lora_scale).out(l1_all, AFB_EPI_BIAS).run(e, s));
      }
      AFB_TRY(Gemm(B, S).a(at_all, D).a(mlp_all, M).a(l1_all, r).w(k.out_w, out_ld, D, k.out_b)
                  .out(proj_out_v, proj_epi).gate_res(m + 2 * D, mod_bs, h_all).run(e, s));
    } else {
      AFB_TRY(Gemm(B, S).a(at_all, D).a(mlp_all, M).w(k.out_w, out_ld, D, k.out_b)
                  .out(proj_out_v, proj_epi).gate_res(m + 2 * D, mod_bs, h_all).run(e, s));
    }
    if (stash)
      AFB_TRY(afb::gate_res_launch(e->h, int64_t(S) * D, e->u1, int64_t(S) * D, m + 2 * D, mod_bs, e->h, int64_t(S) * D, B, S, D, s));
  }
  retarget(nullptr);
  at.lse = nullptr;

  AFB_TRY(save(d.num_double + d.num_single));
  if (save_ckpt) e->stash_valid = stash;

  // ---- norm_out (AdaLayerNormContinuous: scale first, then shift) + heads ------------------------
  NvtxRange nvtx_heads("norm_out_heads");
  const bf16* nm = e->mod + w.norm_out_mod_off;
  int64_t nm_bs = mod_bs;
  if (w.alt_norm_out_w) {  // tied teacher: shares the trunk but has its own (frozen) norm_out Linear
    AFB_TRY(small_linear_rows(e->temb, D, w.alt_norm_out_w, D, w.alt_norm_out_b, e->alt_nm, 2 * D, B, 2 * D, D,
                              AFB_SL_SILU_IN, s));
    nm = e->alt_nm;
    nm_bs = 2 * D;
  }
  AFB_TRY(afb::ln_modulate_launch(h_img.p, h_img.bs, const_cast<bf16*>(y_img.p), y_img.bs, nm, nm + D, nm_bs,
                                  B, Si, D, LN_EPS, s));
  AFB_TRY(Gemm(B, Si).a(y_img, D).w(w.head_w, D, w.head_n, w.head_b)
              .out(View{head_out, w.head_n, int64_t(Si) * w.head_n}, AFB_EPI_BIAS).run(e, s));
  return AFB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Adapter-only backward. Per block (reverse order): recompute the block's activations from its checkpoint, then push
// dh (gradient w.r.t. the residual stream, bf16 [B, S, D]) back through it. The trunk weights are frozen: every big
// GEMM is a dX product that reads the forward weight transposed in place; the LoRA pairs get dB = dY^T T, dA = dT^T X
// from the token-contraction GEMM. Modulation vectors are constants here.
// ------------------------------------------------------------------------------------------------------------------
int tn_rows(View a, int m, View b, int n, float* out, int64_t out_ld, int B, int rows, cudaStream_t s) {
  if (!out) return AFB_OK;
  return afb::gemm_tn_batched_launch(a.p, a.ld, a.bs, b.p, b.ld, b.bs, out, out_ld, B, rows, m, n, s);
}

// Backward of one LoRA-extended Linear  out = [x | t] [W | Bl]^T,  t = dropout(x) A^T  (packed weight `w`, leading dim
// w_ld = in + rank when the layer carries LoRA). dout -> dx = dout W + mask (.) (dt A) / keep, dBl += dout^T t,
// dA += dt^T dropout(x). x may be two column segments (FLUX single block: [attn | mlp]); dx is produced per segment.
// Without dropout the LoRA term of dx joins the main GEMM as a K-extension ([dout | dt] [W ; A]).
struct LoraBwd {
  afb_engine* e;
  int B, rows;
  cudaStream_t s;
  bool drop;

  struct Seg {
    View x;    // forward input segment
    View dx;   // where its gradient goes
    int k;     // columns
  };

  int run(View dout, int out_n, const bf16* w, int64_t w_ld, const bf16* la, int in, View t, View dt, Seg s0, Seg s1,
          uint32_t layer, float* g_la, float* g_lb, bool add_res = false) {
    const int r = e->desc.lora_rank;
    if (la) {
      AFB_TRY(tn_rows(dout, out_n, t, r, g_lb, r, B, rows, s));                                    // dBl += dout^T t
      AFB_TRY(Gemm(B, rows).a(dout, out_n).wt(w + in, w_ld, r, out_n).out(dt, AFB_EPI_BIAS).run(e, s));  // dt = dout Bl
      if (drop) {  // dA += dt^T dropout(x): regenerate the dropped input (same mask as the forward)
        AFB_TRY(drop_into(e, s0.x, B, rows, s0.k, in, 0, layer, s));
        if (s1.k > 0) AFB_TRY(drop_into(e, s1.x, B, rows, s1.k, in, s0.k, layer, s));
        AFB_TRY(tn_rows(dt, r, dropped_view(e, rows, in), in, g_la, in, B, rows, s));
      } else {
        AFB_TRY(tn_rows(dt, r, s0.x, s0.k, g_la, in, B, rows, s));
        if (s1.k > 0) AFB_TRY(tn_rows(dt, r, s1.x, s1.k, g_la ? g_la + s0.k : nullptr, in, B, rows, s));
      }
    }
    const bool merged = la && !drop;
    int col0 = 0;
    for (const Seg* sg : {&s0, &s1}) {
      if (sg->k == 0) continue;
      Gemm g(B, rows);
      g.a(dout, out_n);
      if (merged) {
        g.a(dt, r);
        g.wt(w + col0, w_ld, sg->k, out_n, la + col0, in);
      } else {
        g.wt(w + col0, w_ld, sg->k, out_n);
      }
      g.out(sg->dx, add_res ? AFB_EPI_BIAS_RES : AFB_EPI_BIAS);
      if (add_res) g.res(sg->dx);
      AFB_TRY(g.run(e, s));
      col0 += sg->k;
    }
    if (la && drop) {  // dx += mask (.) (dt A) / keep: the LoRA term goes through the scratch and a masked add
      const View tmp = dropped_view(e, rows, in);
      AFB_TRY(Gemm(B, rows).a(dt, r).wt(la, in, in, r).out(tmp, AFB_EPI_BIAS).run(e, s));
      col0 = 0;
      for (const Seg* sg : {&s0, &s1}) {
        if (sg->k == 0) continue;
        AFB_TRY(afb::dropout_rows_launch(tmp.p + col0, tmp.ld, tmp.bs, const_cast<bf16*>(sg->dx.p), sg->dx.ld, sg->dx.bs, B, rows,
                                         sg->k, in, col0, e->drop_seed, layer, e->drop_p, 0, 1, s));
        col0 += sg->k;
      }
    }
    return AFB_OK;
  }
};

int backward_impl(afb_engine* e, const afb_backward_args* ba, cudaStream_t s) {
  const afb_model_desc& d = e->desc;
  const afb_weights& w = e->w;
  const afb_forward_args* a = &ba->fwd;
  const int B = a->batch, St = a->txt_len, Si = a->img_len, S = St + Si;
  const int D = d.dim, M = d.mlp_dim, H = d.heads;
  const int rpad = d.lora_rank;
  const int r = d.ignore_lora ? 0 : d.lora_rank;
  const int rr = r > 0 ? r : 8;
  const int64_t mod_bs = w.mod_total;
  const size_t ckpt_elems = size_t(B) * S * D;
  const int64_t bsD = int64_t(S) * D, bs3 = int64_t(S) * 3 * D, bsM = int64_t(S) * M, bsR = int64_t(S) * rr;
  NvtxRange nvtx_bwd("afb_backward");

  auto img = [&](bf16* p, int64_t ld) { return View{p + int64_t(St) * ld, ld, int64_t(S) * ld}; };
  auto txt = [&](bf16* p, int64_t ld) { return View{p, ld, int64_t(S) * ld}; };

  afb_attn_desc at{};
  at.q = e->qkv;
  at.k = e->qkv + D;
  at.v = e->qkv + 2 * D;
  at.o = e->attn;
  at.q_ld = at.k_ld = at.v_ld = 3 * D;
  at.o_ld = D;
  at.q_batch_stride = at.k_batch_stride = at.v_batch_stride = bs3;
  at.o_batch_stride = bsD;
  at.batch = B;
  at.seq = S;
  at.heads = H;
  at.scale = 0.f;
  at.lse = e->lse;
  afb_attn_bwd_desc ab{};
  ab.q = e->qkv;
  ab.k = e->qkv + D;
  ab.v = e->qkv + 2 * D;
  ab.qkv_ld = 3 * D;
  ab.qkv_batch_stride = bs3;
  ab.o = e->attn;
  ab.d_o = e->dattn;
  ab.o_ld = D;
  ab.o_batch_stride = bsD;
  ab.lse = e->lse;
  ab.delta_ws = e->delta;
  ab.dq = e->dqkv;
  ab.dk = e->dqkv + D;
  ab.dv = e->dqkv + 2 * D;
  ab.dqkv_ld = 3 * D;
  ab.dqkv_batch_stride = bs3;
  ab.batch = B;
  ab.seq = S;
  ab.heads = H;
  ab.scale = 0.f;
  auto attention_bwd = [&]() -> int {
    ProfScope ps(e, s, 1, 10.0 * B * H * double(S) * S * 128.0);
    return afb::attention_backward_launch(&ab, s);
  };

  // ---- dh <- gradient through norm_out's LayerNorm into the image rows; text rows start at zero -----------------
  AFB_CHECK_CUDA(cudaMemsetAsync(e->dh, 0, ckpt_elems * sizeof(bf16), s));
  {
    const bf16* h_fin = e->ckpt + size_t(d.num_double + d.num_single) * ckpt_elems;
    const bf16* nm = e->mod + w.norm_out_mod_off;  // (scale, shift)
    AFB_TRY(afb::ln_modulate_bwd_launch(h_fin + int64_t(St) * D, bsD, ba->d_head_in, int64_t(Si) * D, e->dh + int64_t(St) * D,
                                        bsD, nm, mod_bs, B, Si, D, LN_EPS, 0, s));
  }

  const View y_all{e->y, D, bsD}, mlp_all{e->mlp, M, bsM};
  View at_all{e->attn, D, bsD}, pre_all{e->mlp_pre, M, bsM}, raw_all{e->qkv_raw, 3 * D, bs3};
  const View l0_all{e->lt0, rr, bsR}, l1_all{e->lt1, rr, bsR}, dl_all{e->dl, rr, bsR};
  const View dh_all{e->dh, D, bsD}, du_all{e->du, D, bsD}, dy_all{e->dy, D, bsD}, dat_all{e->dattn, D, bsD};
  const View dmlp_all{e->dmlp, M, bsM}, dqkv_all{e->dqkv, 3 * D, bs3};
  View u1_all{e->u1, D, bsD};
  const bool drop = e->drop_p > 0.f && r > 0;
  LoraBwd lb{e, B, S, s, drop};
  float* dmod = ba->d_mod;  // fp32 [B, mod_total] or NULL
  // With the activation stash the block outputs of the train forward are copied back instead of being recomputed: no
  // QKV / MLP-up / branch-output GEMMs and no attention forward in the backward.
  const bool stash = e->stash_on && e->stash != nullptr && e->stash_valid;
  // the kept outputs are read in place: the engine's buffer pointers are re-pointed at the block's stash slice
  bf16 *const ws_raw = e->qkv_raw, *const ws_attn = e->attn, *const ws_pre = e->mlp_pre, *const ws_u1 = e->u1,
             *const ws_u2 = e->u2, *const ws_hmid = e->h_mid;
  auto retarget = [&](const StashSlot* sl) {
    e->qkv_raw = sl ? sl->qkv_raw : ws_raw;
    e->attn = sl ? sl->attn : ws_attn;
    e->mlp_pre = sl ? sl->pre : ws_pre;
    e->u1 = sl ? sl->u1 : ws_u1;
    e->u2 = sl && sl->u2 ? sl->u2 : ws_u2;
    e->h_mid = sl && sl->h_mid ? sl->h_mid : ws_hmid;
    at_all = View{e->attn, D, bsD};
    pre_all = View{e->mlp_pre, M, bsM};
    raw_all = View{e->qkv_raw, 3 * D, bs3};
    u1_all = View{e->u1, D, bsD};
    ab.o = e->attn;
    ab.lse = sl ? sl->lse : e->lse;
  };

  // ---- single-stream blocks, last to first ----------------------------------------------------------------------
  for (int i = d.num_single - 1; i >= 0; --i) {
    NvtxRange nvtx_block("bwd_single_block_%d", i);
    const afb_single_block& k = e->sgl[i];
    const afb_single_block_grads* g = ba->sgl ? &ba->sgl[i] : nullptr;
    const bf16* m = e->mod + k.mod_off;  // shift, scale, gate
    const bf16* h_in = e->ckpt + size_t(d.num_double + i) * ckpt_elems;
    const bool lm = r > 0 && k.mlp_la, lo = r > 0 && k.out_la;
    const int64_t mlp_ld = D + (k.mlp_la ? rpad : 0), out_ld = D + M + (k.out_la ? rpad : 0);
    const bf16* mlp_w = static_cast<const bf16*>(k.mlp_w);
    const bf16* out_w = static_cast<const bf16*>(k.out_w);
    // -- recompute (or restore from the stash)
    AFB_TRY(afb::ln_modulate_launch(h_in, bsD, e->y, bsD, m + D, m, mod_bs, B, S, D, LN_EPS, s));
    if (stash) {
      const StashSlot sl = stash_slot(e, d.num_double + i, B, St, Si);
      retarget(&sl);
    } else {
      AFB_TRY(Gemm(B, S).a(y_all, D).w(k.qkv_w, D, 3 * D, k.qkv_b).out(raw_all, AFB_EPI_BIAS).run(e, s));
    }
    AFB_CHECK_CUDA(cudaMemcpyAsync(e->qkv, e->qkv_raw, size_t(B) * S * 3 * D * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
    AFB_TRY(afb::rmsnorm_rope_launch(e->qkv, 3 * D, bs3, 0, D, B, S, H, 0, nullptr, nullptr, k.nq, k.nk, a->rope_cos,
                                     a->rope_sin, LN_EPS, s));
    at.score_bound = k.qk_bound;
    if (!stash) AFB_TRY(run_attention(e, &at, s));
    const uint32_t layer = 4u * (d.num_double + i);
    if (lm) {
      View xa = y_all;
      if (drop) {
        AFB_TRY(drop_into(e, y_all, B, S, D, D, 0, layer, s));
        xa = dropped_view(e, S, D);
      }
      AFB_TRY(Gemm(B, S).a(xa, D).w(k.mlp_la, D, r, nullptr).alpha(e->lora_scale).out(l0_all, AFB_EPI_BIAS).run(e, s));
      if (!stash)
        AFB_TRY(Gemm(B, S).a(y_all, D).a(l0_all, r).w(k.mlp_w, mlp_ld, M, k.mlp_b).out(pre_all, AFB_EPI_BIAS).run(e, s));
    } else if (!stash) {
      AFB_TRY(Gemm(B, S).a(y_all, D).w(k.mlp_w, mlp_ld, M, k.mlp_b).out(pre_all, AFB_EPI_BIAS).run(e, s));
    }
    AFB_TRY(afb::gelu_fwd_launch(e->mlp_pre, M, e->mlp, M, int64_t(B) * S, M, s));
    if (lo) {
      if (drop) {
        AFB_TRY(drop_into(e, at_all, B, S, D, D + M, 0, layer + 1, s));
        AFB_TRY(drop_into(e, mlp_all, B, S, M, D + M, D, layer + 1, s));
        AFB_TRY(Gemm(B, S).a(dropped_view(e, S, D + M), D + M).w(k.out_la, D + M, r, nullptr).alpha(e->lora_scale).out(l1_all, AFB_EPI_BIAS).run(e, s));
      } else {
        AFB_TRY(Gemm(B, S).a(at_all, D).a(mlp_all, M).w(k.out_la, D + M, r, nullptr).alpha(e->lora_scale).out(l1_all, AFB_EPI_BIAS).run(e, s));
      }
    }
    // -- backward
    if (dmod) {  // the gate's gradient needs the branch output u = proj_out([attn | mlp]) the fused forward never stores
      if (stash) {
      } else if (lo)
        AFB_TRY(Gemm(B, S).a(at_all, D).a(mlp_all, M).a(l1_all, r).w(k.out_w, out_ld, D, k.out_b).out(u1_all, AFB_EPI_BIAS).run(e, s));
      else
        AFB_TRY(Gemm(B, S).a(at_all, D).a(mlp_all, M).w(k.out_w, out_ld, D, k.out_b).out(u1_all, AFB_EPI_BIAS).run(e, s));
      AFB_TRY(afb::gate_bwd_launch(e->dh, bsD, e->u1, bsD, m + 2 * D, mod_bs, e->du, bsD, dmod + k.mod_off + 2 * D, mod_bs, B, S,
                                   D, s));
    } else {
      AFB_TRY(afb::rowscale_launch(e->dh, D, bsD, m + 2 * D, mod_bs, e->du, D, bsD, B, S, D, s));
    }
    AFB_TRY(lb.run(du_all, D, out_w, out_ld, lo ? static_cast<const bf16*>(k.out_la) : nullptr, D + M, l1_all, dl_all,
                   {at_all, dat_all, D}, {mlp_all, dmlp_all, M}, layer + 1, g ? g->out_la : nullptr, g ? g->out_lb : nullptr));
    AFB_TRY(afb::gelu_bwd_launch(e->dmlp, M, e->mlp_pre, M, int64_t(B) * S, M, s));
    AFB_TRY(lb.run(dmlp_all, M, mlp_w, mlp_ld, lm ? static_cast<const bf16*>(k.mlp_la) : nullptr, D, l0_all, dl_all,
                   {y_all, dy_all, D}, {View{}, View{}, 0}, layer, g ? g->mlp_la : nullptr, g ? g->mlp_lb : nullptr));
    AFB_TRY(attention_bwd());
    AFB_TRY(afb::rmsnorm_rope_bwd_launch(e->dqkv, e->qkv_raw, 3 * D, bs3, 0, D, B, S, H, 0, nullptr, nullptr, k.nq, k.nk,
                                         a->rope_cos, a->rope_sin, LN_EPS, s));
    AFB_TRY(Gemm(B, S).a(dqkv_all, 3 * D).wt(k.qkv_w, D, D, 3 * D).out(dy_all, AFB_EPI_BIAS_RES).res(dy_all).run(e, s));
    if (dmod)
      AFB_TRY(afb::ln_mod_param_grad_strided_launch(h_in, bsD, e->dy, bsD, e->stats, dmod + k.mod_off + D, dmod + k.mod_off,
                                                    mod_bs, B, S, D, LN_EPS, s));
    AFB_TRY(afb::ln_modulate_bwd_launch(h_in, bsD, e->dy, bsD, e->dh, bsD, m + D, mod_bs, B, S, D, LN_EPS, 1, s));
  }

  // ---- double-stream blocks, last to first ----------------------------------------------------------------------
  for (int i = d.num_double - 1; i >= 0; --i) {
    NvtxRange nvtx_block("bwd_double_block_%d", i);
    const afb_double_block& k = e->dbl[i];
    const afb_double_block_grads* g = ba->dbl ? &ba->dbl[i] : nullptr;
    const bf16* im = e->mod + k.img_mod_off;  // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    const bf16* tm = e->mod + k.txt_mod_off;
    bf16* h_in = e->ckpt + size_t(i) * ckpt_elems;
    struct Stream {
      int rows;
      const bf16* mod;
      View h_in, h_mid, y, qkv_raw, attn, mlp, pre, l0, l1, dl, dh, du, dy, dattn, dmlp, dqkv, u1, u2;
      int64_t mod_off;
      const void *qkv_w, *qkv_b, *out_w, *out_b, *up_w, *up_b, *up_la, *down_w, *down_b, *down_la;
      float *g_up_la, *g_up_lb, *g_down_la, *g_down_lb;
    };
    auto mk = [&](bool is_img) {
      auto v = [&](bf16* p, int64_t ld) { return is_img ? img(p, ld) : txt(p, ld); };
      Stream t{};
      t.rows = is_img ? Si : St;
      t.mod = is_img ? im : tm;
      t.h_in = v(h_in, D);
      t.h_mid = v(e->h_mid, D);
      t.y = v(e->y, D);
      t.qkv_raw = v(e->qkv_raw, 3 * D);
      t.attn = v(e->attn, D);
      t.mlp = v(e->mlp, M);
      t.pre = v(e->mlp_pre, M);
      t.l0 = v(e->lt0, rr);
      t.l1 = v(e->lt1, rr);
      t.dl = v(e->dl, rr);
      t.dh = v(e->dh, D);
      t.du = v(e->du, D);
      t.dy = v(e->dy, D);
      t.dattn = v(e->dattn, D);
      t.dmlp = v(e->dmlp, M);
      t.dqkv = v(e->dqkv, 3 * D);
      t.u1 = v(e->u1, D);
      t.u2 = v(e->u2, D);
      t.mod_off = is_img ? k.img_mod_off : k.txt_mod_off;
      if (is_img) {
        t.qkv_w = k.img_qkv_w, t.qkv_b = k.img_qkv_b, t.out_w = k.img_out_w, t.out_b = k.img_out_b;
        t.up_w = k.img_up_w, t.up_b = k.img_up_b, t.up_la = k.img_up_la;
        t.down_w = k.img_down_w, t.down_b = k.img_down_b, t.down_la = k.img_down_la;
        if (g) t.g_up_la = g->img_up_la, t.g_up_lb = g->img_up_lb, t.g_down_la = g->img_down_la, t.g_down_lb = g->img_down_lb;
      } else {
        t.qkv_w = k.txt_qkv_w, t.qkv_b = k.txt_qkv_b, t.out_w = k.txt_out_w, t.out_b = k.txt_out_b;
        t.up_w = k.txt_up_w, t.up_b = k.txt_up_b, t.up_la = k.txt_up_la;
        t.down_w = k.txt_down_w, t.down_b = k.txt_down_b, t.down_la = k.txt_down_la;
        if (g) t.g_up_la = g->txt_up_la, t.g_up_lb = g->txt_up_lb, t.g_down_la = g->txt_down_la, t.g_down_lb = g->txt_down_lb;
      }
      return t;
    };
    StashSlot dsl{};
    if (stash) {
      dsl = stash_slot(e, i, B, St, Si);
      retarget(&dsl);
    }
    Stream st2[2] = {mk(true), mk(false)};
    // Qwen: the last block's text-stream output is never read (forward_impl skips its out-projection and MLP), so that
    // stream only contributes through its K / V rows: its dattn is zero and its MLP half has no gradient.
    const bool skip_txt_tail = d.arch != AFB_ARCH_FLUX && i == d.num_double - 1;
    // -- recompute (or restore from the stash): attention half
    for (Stream& t : st2) {
      if (stash) break;
      AFB_TRY(afb::ln_modulate_launch(t.h_in.p, t.h_in.bs, const_cast<bf16*>(t.y.p), t.y.bs, t.mod + D, t.mod, mod_bs, B,
                                      t.rows, D, LN_EPS, s));
      AFB_TRY(Gemm(B, t.rows).a(t.y, D).w(t.qkv_w, D, 3 * D, t.qkv_b).out(t.qkv_raw, AFB_EPI_BIAS).run(e, s));
    }
    AFB_CHECK_CUDA(cudaMemcpyAsync(e->qkv, e->qkv_raw, size_t(B) * S * 3 * D * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
    AFB_TRY(afb::rmsnorm_rope_launch(e->qkv, 3 * D, bs3, 0, D, B, S, H, St, k.txt_nq, k.txt_nk, k.img_nq, k.img_nk,
                                     a->rope_cos, a->rope_sin, LN_EPS, s));
    at.score_bound = k.qk_bound;
    if (!stash) AFB_TRY(run_attention(e, &at, s));
    // -- recompute: h_mid and the MLP half; then the MLP half's backward (per stream)
    for (Stream& t : st2) {
      if (skip_txt_tail && &t == &st2[1]) {
        AFB_CHECK_CUDA(cudaMemset2DAsync(const_cast<bf16*>(t.dattn.p), size_t(t.dattn.bs) * sizeof(bf16), 0,
                                         size_t(t.rows) * D * sizeof(bf16), B, s));
        continue;
      }
      LoraBwd sb{e, B, t.rows, s, drop};
      const uint32_t layer_up = 4u * i + (&t == &st2[0] ? 0u : 2u);
      const bool lu = r > 0 && t.up_la, ld_ = r > 0 && t.down_la;
      const int64_t up_ld = D + (t.up_la ? rpad : 0), down_ld = M + (t.down_la ? rpad : 0);
      const bf16* up_w = static_cast<const bf16*>(t.up_w);
      const bf16* down_w = static_cast<const bf16*>(t.down_w);
      if (stash) {
        // h_mid and u1 were restored
      } else if (dmod) {  // keep the un-gated attention branch output for the gate's gradient
        AFB_TRY(Gemm(B, t.rows).a(t.attn, D).w(t.out_w, D, D, t.out_b).out(t.u1, AFB_EPI_BIAS).run(e, s));
        AFB_TRY(afb::gate_res_launch(t.h_in.p, t.h_in.bs, t.u1.p, t.u1.bs, t.mod + 2 * D, mod_bs, const_cast<bf16*>(t.h_mid.p),
                                     t.h_mid.bs, B, t.rows, D, s));
      } else {
        AFB_TRY(Gemm(B, t.rows).a(t.attn, D).w(t.out_w, D, D, t.out_b).out(t.h_mid, AFB_EPI_BIAS_GATE_RES)
                    .gate_res(t.mod + 2 * D, mod_bs, t.h_in).run(e, s));
      }
      AFB_TRY(afb::ln_modulate_launch(t.h_mid.p, t.h_mid.bs, const_cast<bf16*>(t.y.p), t.y.bs, t.mod + 4 * D, t.mod + 3 * D,
                                      mod_bs, B, t.rows, D, LN_EPS, s));
      if (lu) {
        View xa = t.y;
        if (drop) {
          AFB_TRY(drop_into(e, t.y, B, t.rows, D, D, 0, layer_up, s));
          xa = dropped_view(e, t.rows, D);
        }
        AFB_TRY(Gemm(B, t.rows).a(xa, D).w(t.up_la, D, r, nullptr).alpha(e->lora_scale).out(t.l0, AFB_EPI_BIAS).run(e, s));
        if (!stash)
          AFB_TRY(Gemm(B, t.rows).a(t.y, D).a(t.l0, r).w(t.up_w, up_ld, M, t.up_b).out(t.pre, AFB_EPI_BIAS).run(e, s));
      } else if (!stash) {
        AFB_TRY(Gemm(B, t.rows).a(t.y, D).w(t.up_w, up_ld, M, t.up_b).out(t.pre, AFB_EPI_BIAS).run(e, s));
      }
      for (int bi = 0; bi < B; ++bi)
        AFB_TRY(afb::gelu_fwd_launch(t.pre.p + int64_t(bi) * t.pre.bs, M, const_cast<bf16*>(t.mlp.p) + int64_t(bi) * t.mlp.bs, M,
                                     t.rows, M, s));
      if (ld_) {
        View xa = t.mlp;
        if (drop) {
          AFB_TRY(drop_into(e, t.mlp, B, t.rows, M, M, 0, layer_up + 1, s));
          xa = dropped_view(e, t.rows, M);
        }
        AFB_TRY(Gemm(B, t.rows).a(xa, M).w(t.down_la, M, r, nullptr).alpha(e->lora_scale).out(t.l1, AFB_EPI_BIAS).run(e, s));
      }
      // backward of  h_out = h_mid + gate_mlp * down(gelu(up(LNmod2(h_mid))))
      if (dmod) {
        if (stash) {
        } else if (ld_)
          AFB_TRY(Gemm(B, t.rows).a(t.mlp, M).a(t.l1, r).w(t.down_w, down_ld, D, t.down_b).out(t.u2, AFB_EPI_BIAS).run(e, s));
        else
          AFB_TRY(Gemm(B, t.rows).a(t.mlp, M).w(t.down_w, down_ld, D, t.down_b).out(t.u2, AFB_EPI_BIAS).run(e, s));
        AFB_TRY(afb::gate_bwd_launch(t.dh.p, t.dh.bs, t.u2.p, t.u2.bs, t.mod + 5 * D, mod_bs, const_cast<bf16*>(t.du.p), t.du.bs,
                                     dmod + t.mod_off + 5 * D, mod_bs, B, t.rows, D, s));
      } else {
        AFB_TRY(afb::rowscale_launch(t.dh.p, D, t.dh.bs, t.mod + 5 * D, mod_bs, const_cast<bf16*>(t.du.p), D, t.du.bs, B,
                                     t.rows, D, s));
      }
      AFB_TRY(sb.run(t.du, D, down_w, down_ld, ld_ ? static_cast<const bf16*>(t.down_la) : nullptr, M, t.l1, t.dl,
                     {t.mlp, t.dmlp, M}, {View{}, View{}, 0}, layer_up + 1, t.g_down_la, t.g_down_lb));
      for (int bi = 0; bi < B; ++bi)
        AFB_TRY(afb::gelu_bwd_launch(const_cast<bf16*>(t.dmlp.p) + int64_t(bi) * t.dmlp.bs, M, t.pre.p + int64_t(bi) * t.pre.bs,
                                     M, t.rows, M, s));
      AFB_TRY(sb.run(t.dmlp, M, up_w, up_ld, lu ? static_cast<const bf16*>(t.up_la) : nullptr, D, t.l0, t.dl,
                     {t.y, t.dy, D}, {View{}, View{}, 0}, layer_up, t.g_up_la, t.g_up_lb));
      if (dmod)
        AFB_TRY(afb::ln_mod_param_grad_strided_launch(t.h_mid.p, t.h_mid.bs, t.dy.p, t.dy.bs, e->stats,
                                                      dmod + t.mod_off + 4 * D, dmod + t.mod_off + 3 * D, mod_bs, B, t.rows, D,
                                                      LN_EPS, s));
      AFB_TRY(afb::ln_modulate_bwd_launch(t.h_mid.p, t.h_mid.bs, t.dy.p, t.dy.bs, const_cast<bf16*>(t.dh.p), t.dh.bs,
                                          t.mod + 4 * D, mod_bs, B, t.rows, D, LN_EPS, 1, s));
      // attention half: dattn = (gate_msa * dh) W_out
      if (dmod)
        AFB_TRY(afb::gate_bwd_launch(t.dh.p, t.dh.bs, t.u1.p, t.u1.bs, t.mod + 2 * D, mod_bs, const_cast<bf16*>(t.du.p), t.du.bs,
                                     dmod + t.mod_off + 2 * D, mod_bs, B, t.rows, D, s));
      else
        AFB_TRY(afb::rowscale_launch(t.dh.p, D, t.dh.bs, t.mod + 2 * D, mod_bs, const_cast<bf16*>(t.du.p), D, t.du.bs, B,
                                     t.rows, D, s));
      AFB_TRY(Gemm(B, t.rows).a(t.du, D).wt(t.out_w, D, D, D).out(t.dattn, AFB_EPI_BIAS).run(e, s));
    }
    AFB_TRY(attention_bwd());
    AFB_TRY(afb::rmsnorm_rope_bwd_launch(e->dqkv, e->qkv_raw, 3 * D, bs3, 0, D, B, S, H, St, k.txt_nq, k.txt_nk, k.img_nq,
                                         k.img_nk, a->rope_cos, a->rope_sin, LN_EPS, s));
    for (Stream& t : st2) {
      AFB_TRY(Gemm(B, t.rows).a(t.dqkv, 3 * D).wt(t.qkv_w, D, D, 3 * D).out(t.dy, AFB_EPI_BIAS).run(e, s));
      if (dmod)
        AFB_TRY(afb::ln_mod_param_grad_strided_launch(t.h_in.p, t.h_in.bs, t.dy.p, t.dy.bs, e->stats, dmod + t.mod_off + D,
                                                      dmod + t.mod_off, mod_bs, B, t.rows, D, LN_EPS, s));
      AFB_TRY(afb::ln_modulate_bwd_launch(t.h_in.p, t.h_in.bs, t.dy.p, t.dy.bs, const_cast<bf16*>(t.dh.p), t.dh.bs, t.mod + D,
                                          mod_bs, B, t.rows, D, LN_EPS, 1, s));
    }
  }
  retarget(nullptr);
  return AFB_OK;
}

int check_shapes(afb_engine* e, int B, int St, int Si) {
  AFB_REQUIRE(e != nullptr, "engine: null handle");
  AFB_REQUIRE(e->bound, "engine: weights not bound (call afb_engine_bind)");
  AFB_REQUIRE(B >= 1 && St >= 1 && Si >= 1, "engine: empty problem (batch=%d txt=%d img=%d)", B, St, Si);
  AFB_REQUIRE(e->ws != nullptr && B <= e->cap_batch && St <= e->cap_txt && Si <= e->cap_img,
              "engine: workspace too small for (batch=%d txt=%d img=%d); call afb_engine_reserve", B, St, Si);
  return AFB_OK;
}

}  // namespace

extern "C" {

int afb_engine_create(const afb_model_desc* desc, afb_engine** out) {
  AFB_REQUIRE(desc && out, "engine_create: null argument");
  AFB_REQUIRE(desc->arch == AFB_ARCH_FLUX || desc->arch == AFB_ARCH_QWEN, "engine_create: unknown arch %d", desc->arch);
  AFB_REQUIRE(desc->dim == desc->heads * 128, "engine_create: dim=%d must equal heads*128 (heads=%d)", desc->dim, desc->heads);
  AFB_REQUIRE(desc->dim % 256 == 0 && desc->dim <= 4096, "engine_create: dim=%d must be a multiple of 256 <= 4096", desc->dim);
  AFB_REQUIRE(desc->mlp_dim % 64 == 0 && desc->in_channels % 64 == 0 && desc->txt_dim % 256 == 0,
              "engine_create: mlp_dim/in_channels must be multiples of 64 and txt_dim of 256");
  AFB_REQUIRE(desc->lora_rank == 0 || desc->lora_rank % 256 == 0, "engine_create: lora_rank must be 0 or a multiple of 256");
  AFB_REQUIRE(desc->arch != AFB_ARCH_FLUX || desc->pooled_dim % 256 == 0, "engine_create: pooled_dim must be a multiple of 256");
  AFB_REQUIRE(desc->num_double >= 0 && desc->num_single >= 0, "engine_create: negative depth");
  afb_engine* e = new (std::nothrow) afb_engine();
  AFB_REQUIRE(e != nullptr, "engine_create: out of host memory");
  e->desc = *desc;
  if (const char* env = getenv("AFB_ENGINE_FUSE_QK_ROPE")) e->fuse_qk_rope = atoi(env) != 0;  // A/B switch (default on)
  *out = e;
  return AFB_OK;
}

void afb_engine_destroy(afb_engine* e) {
  if (!e) return;
  if (e->ws) cudaFree(e->ws);
  if (e->tws) cudaFree(e->tws);
  if (e->stash) cudaFree(e->stash);
  for (auto& r : e->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  delete e;
}

int afb_engine_bind(afb_engine* e, const afb_weights* w) {
  AFB_REQUIRE(e && w, "engine_bind: null argument");
  AFB_REQUIRE(w->x_emb_w && w->ctx_w && w->t1_w && w->t2_w && w->mod_w && w->head_w, "engine_bind: missing weights");
  AFB_REQUIRE(w->head_n >= 8 && w->head_n % 8 == 0, "engine_bind: head_n=%d must be a multiple of 8", w->head_n);
  AFB_REQUIRE(w->mod_total > 0 && w->mod_total % 8 == 0, "engine_bind: bad mod_total");
  AFB_REQUIRE(e->desc.num_double == 0 || w->dbl, "engine_bind: double blocks missing");
  AFB_REQUIRE(e->desc.num_single == 0 || w->sgl, "engine_bind: single blocks missing");
  e->w = *w;
  e->dbl.assign(w->dbl, w->dbl + e->desc.num_double);
  e->sgl.assign(w->sgl, w->sgl + e->desc.num_single);
  e->w.dbl = e->dbl.data();
  e->w.sgl = e->sgl.data();
  e->bound = true;
  if (e->ws) carve(e, static_cast<uint8_t*>(e->ws), e->cap_batch, e->cap_txt, e->cap_img);
  return AFB_OK;
}

int afb_engine_set_ignore_lora(afb_engine* e, int32_t on) {
  AFB_REQUIRE(e != nullptr, "engine: null handle");
  e->desc.ignore_lora = on ? 1 : 0;
  return AFB_OK;
}

int afb_engine_set_lora_scale(afb_engine* e, float scale) {
  AFB_REQUIRE(e != nullptr, "engine: null handle");
  AFB_REQUIRE(scale == scale && scale > -1e6f && scale < 1e6f, "engine_set_lora_scale: bad scale");
  e->lora_scale = scale;
  return AFB_OK;
}

size_t afb_engine_workspace_bytes(const afb_engine* e, int32_t batch, int32_t txt_len, int32_t img_len) {
  if (!e || batch < 1 || txt_len < 1 || img_len < 1) return 0;
  afb_engine tmp = *e;
  return carve(&tmp, nullptr, batch, txt_len, img_len);
}

int afb_engine_reserve(afb_engine* e, int32_t batch, int32_t txt_len, int32_t img_len) {
  AFB_REQUIRE(e != nullptr, "engine: null handle");
  AFB_REQUIRE(e->bound, "engine_reserve: bind weights first");
  AFB_REQUIRE(batch >= 1 && txt_len >= 1 && img_len >= 1, "engine_reserve: empty problem");
  if (e->ws && batch <= e->cap_batch && txt_len <= e->cap_txt && img_len <= e->cap_img) return AFB_OK;
  if (e->ws) {
    AFB_CHECK_CUDA(cudaDeviceSynchronize());
    AFB_CHECK_CUDA(cudaFree(e->ws));
    e->ws = nullptr;
  }
  const size_t bytes = carve(e, nullptr, batch, txt_len, img_len);
  AFB_CHECK_CUDA(cudaMalloc(&e->ws, bytes));
  e->ws_bytes = bytes;
  e->cap_batch = batch;
  e->cap_txt = txt_len;
  e->cap_img = img_len;
  carve(e, static_cast<uint8_t*>(e->ws), batch, txt_len, img_len);
  return AFB_OK;
}

int afb_engine_export(afb_engine* e, int32_t which, void* dst, int32_t batch, int32_t txt_len, int32_t img_len,
                      void* stream) {
  AFB_TRY(check_shapes(e, batch, txt_len, img_len));
  AFB_REQUIRE(dst != nullptr && which >= 0 && which <= 2, "engine_export: bad arguments");
  carve(e, static_cast<uint8_t*>(e->ws), batch, txt_len, img_len);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t D = e->desc.dim, S = size_t(txt_len) + img_len;
  if (which == 2) {
    AFB_CHECK_CUDA(cudaMemcpyAsync(dst, e->temb, size_t(batch) * D * 2, cudaMemcpyDeviceToDevice, s));
    return AFB_OK;
  }
  const bf16* src = (which == 0 ? e->h : e->y) + size_t(txt_len) * D;  // image rows of the joint buffer
  AFB_CHECK_CUDA(cudaMemcpy2DAsync(dst, size_t(img_len) * D * 2, src, S * D * 2, size_t(img_len) * D * 2, batch,
                                   cudaMemcpyDeviceToDevice, s));
  return AFB_OK;
}

int afb_engine_set_profiling(afb_engine* e, int32_t on) {
  AFB_REQUIRE(e != nullptr, "engine: null handle");
  e->profiling = on != 0;
  return AFB_OK;
}

int afb_engine_read_profile(afb_engine* e, afb_profile* out) {
  AFB_REQUIRE(e && out, "engine_read_profile: null argument");
  AFB_CHECK_CUDA(cudaDeviceSynchronize());
  for (auto& r : e->prof) {
    float ms = 0.f;
    AFB_CHECK_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    if (r.cls == 0 || r.cls == 2) {
      e->prof_acc.gemm_ms += ms;
      e->prof_acc.gemm_flops += r.flops;
      e->prof_acc.gemm_launches += 1;
      if (r.cls == 2) {
        e->prof_acc.gemm_fused_qk_ms += ms;
        e->prof_acc.gemm_fused_qk_flops += r.flops;
        e->prof_acc.gemm_fused_qk_launches += 1;
      }
    } else {
      e->prof_acc.attn_ms += ms;
      e->prof_acc.attn_flops += r.flops;
      e->prof_acc.attn_launches += 1;
    }
    e->ev_pool.push_back(r.e0);
    e->ev_pool.push_back(r.e1);
  }
  e->prof.clear();
  *out = e->prof_acc;
  e->prof_acc = afb_profile{};
  return AFB_OK;
}

int afb_engine_forward(afb_engine* e, const afb_forward_args* a, void* stream) {
  AFB_REQUIRE(a != nullptr, "engine_forward: null args");
  AFB_TRY(check_shapes(e, a->batch, a->txt_len, a->img_len));
  AFB_REQUIRE(a->latents && a->txt && a->timestep && a->rope_cos && a->rope_sin && a->head_out,
              "engine_forward: null tensor argument");
  // the workspace views are carved for the reserved capacity; re-carve for the actual shape
  carve(e, static_cast<uint8_t*>(e->ws), a->batch, a->txt_len, a->img_len);
  return forward_impl(e, a, static_cast<const bf16*>(a->latents), static_cast<bf16*>(a->head_out),
                      static_cast<cudaStream_t>(stream));
}

int afb_engine_train_reserve(afb_engine* e, int32_t batch, int32_t txt_len, int32_t img_len) {
  AFB_REQUIRE(e != nullptr, "engine: null handle");
  AFB_REQUIRE(e->bound, "engine_train_reserve: bind weights first");
  AFB_REQUIRE(batch >= 1 && txt_len >= 1 && img_len >= 1, "engine_train_reserve: empty problem");
  AFB_TRY(afb_engine_reserve(e, batch, txt_len, img_len));
  if (e->tws && batch <= e->tcap_batch && txt_len <= e->tcap_txt && img_len <= e->tcap_img) return AFB_OK;
  if (e->tws) {
    AFB_CHECK_CUDA(cudaDeviceSynchronize());
    AFB_CHECK_CUDA(cudaFree(e->tws));
    e->tws = nullptr;
  }
  const size_t bytes = carve_train(e, nullptr, batch, txt_len, img_len);
  AFB_CHECK_CUDA(cudaMalloc(&e->tws, bytes));
  e->tws_bytes = bytes;
  e->tcap_batch = batch;
  e->tcap_txt = txt_len;
  e->tcap_img = img_len;
  e->saved_batch = 0;
  return AFB_OK;
}

int64_t afb_engine_stash_bytes(afb_engine* e, int32_t batch, int32_t txt_len, int32_t img_len) {
  if (!e || batch < 1 || txt_len < 1 || img_len < 1) return 0;
  return int64_t(stash_total_bytes(e->desc, batch, txt_len, img_len));
}

int afb_engine_set_activation_stash(afb_engine* e, int32_t on, int32_t batch, int32_t txt_len, int32_t img_len) {
  AFB_REQUIRE(e != nullptr, "engine: null handle");
  if (!on) {
    if (e->stash) {
      AFB_CHECK_CUDA(cudaDeviceSynchronize());
      AFB_CHECK_CUDA(cudaFree(e->stash));
    }
    e->stash = nullptr;
    e->stash_bytes = 0;
    e->stash_on = e->stash_valid = false;
    return AFB_OK;
  }
  AFB_REQUIRE(e->bound, "engine_set_activation_stash: bind weights first");
  AFB_REQUIRE(batch >= 1 && txt_len >= 1 && img_len >= 1, "engine_set_activation_stash: empty problem");
  const size_t bytes = stash_total_bytes(e->desc, batch, txt_len, img_len);
  if (!e->stash || bytes > e->stash_bytes) {
    if (e->stash) {
      AFB_CHECK_CUDA(cudaDeviceSynchronize());
      AFB_CHECK_CUDA(cudaFree(e->stash));
      e->stash = nullptr;
      e->stash_bytes = 0;
    }
    const cudaError_t err = cudaMalloc(&e->stash, bytes);
    if (err != cudaSuccess) {  // not fatal: the engine keeps recomputing; clear the error so later launch checks stay clean
      cudaGetLastError();
      e->stash = nullptr;
      e->stash_on = e->stash_valid = false;
      afb::set_last_error("engine_set_activation_stash: cannot allocate %zu bytes (%s)", bytes, cudaGetErrorString(err));
      return AFB_ERR_CUDA;
    }
    e->stash_bytes = bytes;
  }
  e->stash_on = true;
  e->stash_valid = false;
  return AFB_OK;
}

int afb_engine_forward_train(afb_engine* e, const afb_forward_args* a, void* stream) {
  AFB_REQUIRE(a != nullptr, "engine_forward_train: null args");
  AFB_TRY(check_shapes(e, a->batch, a->txt_len, a->img_len));
  AFB_REQUIRE(e->tws && a->batch <= e->tcap_batch && a->txt_len <= e->tcap_txt && a->img_len <= e->tcap_img,
              "engine_forward_train: call afb_engine_train_reserve first");
  AFB_REQUIRE(a->latents && a->txt && a->timestep && a->rope_cos && a->rope_sin && a->head_out,
              "engine_forward_train: null tensor argument");
  carve(e, static_cast<uint8_t*>(e->ws), a->batch, a->txt_len, a->img_len);
  carve_train(e, static_cast<uint8_t*>(e->tws), a->batch, a->txt_len, a->img_len);
  if (e->lora_scale != 1.0f) {
    afb::set_last_error("engine_forward_train: the training path is built for LoRA scale 1 (alpha = rank), got %f",
                        double(e->lora_scale));
    return AFB_ERR_UNSUPPORTED;
  }
  AFB_REQUIRE(!e->stash_on || stash_total_bytes(e->desc, a->batch, a->txt_len, a->img_len) <= e->stash_bytes,
              "engine_forward_train: the activation stash was sized for a smaller problem");
  e->saved_batch = a->batch;
  e->saved_txt = a->txt_len;
  e->saved_img = a->img_len;
  return forward_impl(e, a, static_cast<const bf16*>(a->latents), static_cast<bf16*>(a->head_out),
                      static_cast<cudaStream_t>(stream), true);
}

int afb_engine_backward(afb_engine* e, const afb_backward_args* ba, void* stream) {
  AFB_REQUIRE(ba != nullptr, "engine_backward: null args");
  const afb_forward_args& a = ba->fwd;
  AFB_TRY(check_shapes(e, a.batch, a.txt_len, a.img_len));
  AFB_REQUIRE(e->tws && e->saved_batch == a.batch && e->saved_txt == a.txt_len && e->saved_img == a.img_len,
              "engine_backward: no checkpoints of this shape (run afb_engine_forward_train first)");
  AFB_REQUIRE(ba->d_head_in && a.rope_cos && a.rope_sin, "engine_backward: null tensor argument");
  carve(e, static_cast<uint8_t*>(e->ws), a.batch, a.txt_len, a.img_len);
  carve_train(e, static_cast<uint8_t*>(e->tws), a.batch, a.txt_len, a.img_len);
  return backward_impl(e, ba, static_cast<cudaStream_t>(stream));
}

// d_mod (gradient of every AdaLN vector, fp32 [B, mod_total]) -> temb -> the timestep embedder's two LoRA pairs.
// temb = t2(silu(t1(sinusoid(t)))) + guidance path + pooled-text path (both frozen), mod = mod_w silu(temb) + mod_b.
int afb_engine_backward_embed(afb_engine* e, const afb_forward_args* a, const float* d_mod, const afb_embed_grads* g,
                              void* stream) {
  AFB_REQUIRE(a && d_mod && g, "engine_backward_embed: null argument");
  AFB_TRY(check_shapes(e, a->batch, a->txt_len, a->img_len));
  AFB_REQUIRE(e->tws && e->saved_batch == a->batch && e->saved_txt == a->txt_len && e->saved_img == a->img_len,
              "engine_backward_embed: run afb_engine_forward_train first");
  AFB_REQUIRE(a->timestep != nullptr, "engine_backward_embed: timestep missing");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  carve(e, static_cast<uint8_t*>(e->ws), a->batch, a->txt_len, a->img_len);
  carve_train(e, static_cast<uint8_t*>(e->tws), a->batch, a->txt_len, a->img_len);
  const afb_weights& w = e->w;
  const int B = a->batch, D = e->desc.dim, r = e->desc.ignore_lora ? 0 : e->desc.lora_rank;
  if (r == 0 || !w.t1_la || !w.t1_lb || !w.t2_la || !w.t2_lb) return AFB_OK;  // nothing trainable upstream of temb
  // dtemb = (d_mod mod_w) * silu'(temb)
  AFB_CHECK_CUDA(cudaMemsetAsync(e->dsilu, 0, size_t(B) * D * sizeof(float), s));
  AFB_TRY(afb::rowlinear_dx_launch(d_mod, w.mod_total, w.mod_w, D, e->dsilu, D, B, int(w.mod_total), D, s));
  AFB_TRY(afb::silu_bwd_launch(e->dsilu, D, e->temb, D, B, D, s));
  // recompute the timestep path's intermediates (with the forward's dropout masks on the LoRA inputs)
  const bool drop = e->drop_p > 0.f;
  AFB_TRY(afb::timestep_embed_launch(a->timestep, e->tproj_t, B, s));
  AFB_TRY(small_linear_rows(e->tproj_t, 256, w.t1_w, 256, w.t1_b, e->tmp_t, D, B, D, 256, 0, s));
  const bf16* x1 = e->tproj_t;  // LoRA input of linear_1
  if (drop) {
    AFB_TRY(afb::dropout_rows_launch(e->tproj_t, 256, int64_t(B) * 256, e->tproj, 256, int64_t(B) * 256, 1, B, 256, 256, 0,
                                     e->drop_seed, DROP_ID_T1, e->drop_p, 0, 0, s));
    x1 = e->tproj;
  }
  AFB_TRY(small_linear_rows(x1, 256, w.t1_la, 256, nullptr, e->ltv1, r, B, r, 256, 0, s));
  AFB_TRY(small_linear_rows(e->ltv1, r, w.t1_lb, r, nullptr, e->tmp_t, D, B, D, r, AFB_SL_ACCUMULATE, s));
  if (drop) {
    AFB_TRY(afb::dropout_rows_launch(e->tmp_t, D, int64_t(B) * D, e->xd_small, D, int64_t(B) * D, 1, B, D, D, 0, e->drop_seed,
                                     DROP_ID_T2, e->drop_p, 1, 0, s));
    AFB_TRY(small_linear_rows(e->xd_small, D, w.t2_la, D, nullptr, e->ltv2, r, B, r, D, 0, s));
  } else {
    AFB_TRY(small_linear_rows(e->tmp_t, D, w.t2_la, D, nullptr, e->ltv2, r, B, r, D, AFB_SL_SILU_IN, s));
  }
  // linear_2: temb_t = W2 silu(tmp) + B2 (A2 dropout(silu(tmp)))
  if (g->t2_lb) AFB_TRY(afb::rowlinear_param_grad_launch(e->dsilu, D, e->ltv2, r, g->t2_lb, r, nullptr, B, D, r, 0, s));
  AFB_CHECK_CUDA(cudaMemsetAsync(e->dltv, 0, size_t(B) * r * sizeof(float), s));
  AFB_TRY(afb::rowlinear_dx_launch(e->dsilu, D, w.t2_lb, r, e->dltv, r, B, D, r, s));
  if (g->t2_la)
    AFB_TRY(afb::rowlinear_param_grad_launch(e->dltv, r, drop ? e->xd_small : e->tmp_t, D, g->t2_la, D, nullptr, B, r, D,
                                             drop ? 0 : 1, s));
  AFB_CHECK_CUDA(cudaMemsetAsync(e->dtmp, 0, size_t(B) * D * sizeof(float), s));
  AFB_TRY(afb::rowlinear_dx_launch(e->dsilu, D, w.t2_w, D, e->dtmp, D, B, D, D, s));
  if (drop) {
    AFB_CHECK_CUDA(cudaMemsetAsync(e->dtmp2, 0, size_t(B) * D * sizeof(float), s));
    AFB_TRY(afb::rowlinear_dx_launch(e->dltv, r, w.t2_la, D, e->dtmp2, D, B, r, D, s));
    AFB_TRY(afb::dropout_f32_add_launch(e->dtmp2, D, e->dtmp, D, B, D, e->drop_seed, DROP_ID_T2, e->drop_p, s));
  } else {
    AFB_TRY(afb::rowlinear_dx_launch(e->dltv, r, w.t2_la, D, e->dtmp, D, B, r, D, s));
  }
  AFB_TRY(afb::silu_bwd_launch(e->dtmp, D, e->tmp_t, D, B, D, s));
  // linear_1: tmp = W1 p + B1 (A1 dropout(p))
  if (g->t1_lb) AFB_TRY(afb::rowlinear_param_grad_launch(e->dtmp, D, e->ltv1, r, g->t1_lb, r, nullptr, B, D, r, 0, s));
  AFB_CHECK_CUDA(cudaMemsetAsync(e->dltv, 0, size_t(B) * r * sizeof(float), s));
  AFB_TRY(afb::rowlinear_dx_launch(e->dtmp, D, w.t1_lb, r, e->dltv, r, B, D, r, s));
  if (g->t1_la) AFB_TRY(afb::rowlinear_param_grad_launch(e->dltv, r, x1, 256, g->t1_la, 256, nullptr, B, r, 256, 0, s));
  return AFB_OK;
}

int afb_engine_set_lora_dropout(afb_engine* e, float p, uint64_t seed) {
  AFB_REQUIRE(e != nullptr, "engine: null handle");
  AFB_REQUIRE(p >= 0.f && p < 1.f, "engine_set_lora_dropout: p=%f must be in [0, 1)", double(p));
  e->drop_p = p;
  e->drop_seed = seed;
  return AFB_OK;
}

int afb_engine_denoise(afb_engine* e, const afb_denoise_args* a, void* stream) {
  AFB_REQUIRE(a != nullptr, "engine_denoise: null args");
  const afb_forward_args& f = a->fwd;
  AFB_TRY(check_shapes(e, f.batch, f.txt_len, f.img_len));
  AFB_REQUIRE(a->nfe >= 1 && a->sigmas && a->timesteps && a->x, "engine_denoise: bad arguments");
  AFB_REQUIRE(f.txt && f.rope_cos && f.rope_sin, "engine_denoise: null tensor argument");
  AFB_REQUIRE(e->desc.head_mode == 0, "engine_denoise: needs the ArcFlow heads (head_mode 0)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  carve(e, static_cast<uint8_t*>(e->ws), f.batch, f.txt_len, f.img_len);
  const int64_t tokens = int64_t(f.batch) * f.img_len;
  float* x = static_cast<float*>(a->x);
  AFB_TRY(afb::cast_f32_bf16_launch(x, e->x_bf16, tokens * e->desc.in_channels, s));
  NvtxRange nvtx_loop("afb_denoise");
  for (int i = 0; i < a->nfe; ++i) {
    NvtxRange nvtx_nfe("nfe_%d", i);
    afb_forward_args fa = f;
    AFB_TRY(afb::fill_f32_launch(e->t_dev, a->timesteps[i], f.batch, s));
    fa.timestep = e->t_dev;
    // guidance stays whatever the caller passed (device fp32 [batch]) — constant over the loop
    AFB_TRY(forward_impl(e, &fa, e->x_bf16, e->head, s));
    // sigma_start == sigma_src inside the pipeline loop (arcflux_pipeline.py:495-503)
    AFB_TRY(afb::sampler_step_launch(e->head, e->w.head_n, x, x, e->x_bf16, tokens, e->desc.num_gaussians,
                                     a->sigmas[i], a->sigmas[i], a->sigmas[i + 1],
                                     a->eps > 0.f ? a->eps : 1e-4f, s));
  }
  return AFB_OK;
}

}  // extern "C"
