// arcflow_b200 — the HBM-bound kernels around the tensor-core tiles: AdaLN modulate, per-head
// RMSNorm + RoPE, the batch-row ("small-M") Linear used for every AdaLN modulation vector and the
// time/text embedders, the sinusoidal timestep embedding and the fused ArcFlow sampler step.
// All are one-warp-per-row streaming kernels with 16-byte vector accesses.
#include <math.h>

#include "common.cuh"
#include "../../include/arcflow_b200.h"

namespace afb {
void count_launch(int n);

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// ------------------------------------------------------------------------------------------------
// AdaLN: y = LN(x) * (1 + scale[b]) + shift[b].   Reference: diffusers AdaLayerNormZero /
// AdaLayerNormZeroSingle / AdaLayerNormContinuous as used by the blocks built at
// lakonlab/models/architecture/arcflow/arcflux.py:63-85.
// One warp per row, the row lives in registers (NCH chunks of 8 bf16 per lane).
// ------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256)
ln_modulate_kernel(const __nv_bfloat16* __restrict__ x, long long x_bs, __nv_bfloat16* __restrict__ y,
                   long long y_bs, const __nv_bfloat16* __restrict__ scale,
                   const __nv_bfloat16* __restrict__ shift, long long mod_bs, int batches,
                   int rows_per_batch, float eps) {
  constexpr int DIM = NCH * 256;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)batches * rows_per_batch;
  if (row >= total) return;
  const int b = int(row / rows_per_batch);
  const int r = int(row - (long long)b * rows_per_batch);
  const __nv_bfloat16* xr = x + (long long)b * x_bs + (long long)r * DIM;
  __nv_bfloat16* yr = y + (long long)b * y_bs + (long long)r * DIM;

  float v[NCH][8];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + c * 256 + lane * 8);
    unpack8(u, v[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += v[c][i];
  }
  const float mean = warp_sum(sum) * (1.0f / DIM);
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = v[c][i] - mean;
      sq += d * d;
    }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / DIM) + eps);
  const __nv_bfloat16* sc = scale + (long long)b * mod_bs;
  const __nv_bfloat16* sh = shift + (long long)b * mod_bs;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float s[8], t[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(sc + c * 256 + lane * 8), s);
    unpack8(*reinterpret_cast<const uint4*>(sh + c * 256 + lane * 8), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (v[c][i] - mean) * rstd * (1.0f + s[i]) + t[i];
    *reinterpret_cast<uint4*>(yr + c * 256 + lane * 8) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-head RMSNorm(q, k) + rotary embedding, in place on the fused QKV buffer.
// Reference semantics: diffusers RMSNorm (fp32 variance, cast to bf16, * weight) then apply_rotary_emb
// (fp32 math on adjacent pairs, cast back) — SURVEY.md Appendix A.1, A.3, A.5; rope tables are the
// ones built at arcflux.py:171-173.  One warp per token row, 4 elements per lane per head.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rmsnorm_rope_kernel(__nv_bfloat16* __restrict__ qkv, long long ld, long long bs, int q_off, int k_off,
                    int batches, int seq, int heads, int txt_rows,
                    const __nv_bfloat16* __restrict__ wq_txt, const __nv_bfloat16* __restrict__ wk_txt,
                    const __nv_bfloat16* __restrict__ wq_img, const __nv_bfloat16* __restrict__ wk_img,
                    const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, float eps) {
  // One warp per token row; a half-warp (16 lanes x 8 elements = 16 B per lane) owns one head at a time, so the
  // warp streams two heads per step with 16-byte accesses; cos/sin/weights of this lane's 8 head-columns are
  // loaded once and reused for all heads of the row.
  const int lane = threadIdx.x & 31;
  const int hl = lane & 15;    // position inside the head: elements [8 hl, 8 hl + 8)
  const int half = lane >> 4;  // which of the two heads of this step
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)batches * seq) return;
  const int b = int(row / seq);
  const int s = int(row - (long long)b * seq);
  __nv_bfloat16* base = qkv + (long long)b * bs + (long long)s * ld;
  const bool is_txt = s < txt_rows;
  float cs[8], sn[8], wq[8], wk[8];
  {
    const float4 c0 = *reinterpret_cast<const float4*>(cos_tab + (long long)s * 128 + hl * 8);
    const float4 c1 = *reinterpret_cast<const float4*>(cos_tab + (long long)s * 128 + hl * 8 + 4);
    const float4 s0 = *reinterpret_cast<const float4*>(sin_tab + (long long)s * 128 + hl * 8);
    const float4 s1 = *reinterpret_cast<const float4*>(sin_tab + (long long)s * 128 + hl * 8 + 4);
    cs[0] = c0.x; cs[1] = c0.y; cs[2] = c0.z; cs[3] = c0.w; cs[4] = c1.x; cs[5] = c1.y; cs[6] = c1.z; cs[7] = c1.w;
    sn[0] = s0.x; sn[1] = s0.y; sn[2] = s0.z; sn[3] = s0.w; sn[4] = s1.x; sn[5] = s1.y; sn[6] = s1.z; sn[7] = s1.w;
    unpack8(*reinterpret_cast<const uint4*>((is_txt ? wq_txt : wq_img) + hl * 8), wq);
    unpack8(*reinterpret_cast<const uint4*>((is_txt ? wk_txt : wk_img) + hl * 8), wk);
  }
  const int slots = 2 * heads;  // q heads then k heads
  constexpr int UNROLL = 4;     // 4 x 16-byte loads in flight per lane
  for (int s0 = 0; s0 < slots; s0 += 2 * UNROLL) {
    uint4 u[UNROLL];
    __nv_bfloat16* ptr[UNROLL];
    bool act[UNROLL], is_k[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      const int slot = s0 + 2 * i + half;
      act[i] = slot < slots;
      is_k[i] = slot >= heads;
      const int h = is_k[i] ? slot - heads : slot;
      ptr[i] = base + (is_k[i] ? k_off : q_off) + h * 128 + hl * 8;
      u[i] = act[i] ? *reinterpret_cast<const uint4*>(ptr[i]) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      float x[8];
      unpack8(u[i], x);
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(x[e], x[e], ss);
      ss += __shfl_xor_sync(0xffffffffu, ss, 8);
      ss += __shfl_xor_sync(0xffffffffu, ss, 4);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      const float rs = rsqrtf(ss * (1.0f / 128.0f) + eps);
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = round_bf16(round_bf16(x[e] * rs) * (is_k[i] ? wk[e] : wq[e]));
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        o[e] = x[e] * cs[e] - x[e + 1] * sn[e];
        o[e + 1] = x[e + 1] * cs[e + 1] + x[e] * sn[e + 1];
      }
      if (act[i]) *reinterpret_cast<uint4*>(ptr[i]) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Small-M Linear: y[m, n] (+)= act(x[m, :]) . W[n, :] + bias[n],  m <= 8.  HBM-bound on W: every
// warp streams 4 weight rows at a time with 16-byte loads; x (<= 8 rows) sits in shared memory.
// Used for all AdaLN modulation Linears of a forward in ONE launch (weights concatenated along n) and
// for the time / guidance / pooled-text embedder MLPs.
// ------------------------------------------------------------------------------------------------
constexpr int SL_ROWS = 8;
constexpr int SL_COLS = 4;

__global__ void __launch_bounds__(256)
small_linear_kernel(const __nv_bfloat16* __restrict__ x, long long x_ld,
                    const __nv_bfloat16* __restrict__ w, long long w_ld,
                    const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                    long long y_ld, int m, int n, int k, int flags) {
  extern __shared__ __align__(16) uint8_t sl_smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(sl_smem);  // [SL_ROWS][k]
  for (int i = threadIdx.x; i < SL_ROWS * k; i += blockDim.x) {
    const int r = i / k, c = i - r * k;
    float v = 0.f;
    if (r < m) {
      v = __bfloat162float(x[(long long)r * x_ld + c]);
      if (flags & AFB_SL_SILU_IN) v = silu(v);
    }
    xs[i] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int groups = (n + SL_COLS - 1) / SL_COLS;
  for (int g = blockIdx.x * warps_per_block + (threadIdx.x >> 5); g < groups;
       g += gridDim.x * warps_per_block) {
    const int n0 = g * SL_COLS;
    float acc[SL_COLS][SL_ROWS];
#pragma unroll
    for (int c = 0; c < SL_COLS; ++c)
#pragma unroll
      for (int r = 0; r < SL_ROWS; ++r) acc[c][r] = 0.f;
    for (int kc = lane * 8; kc < k; kc += 256) {
      float wv[SL_COLS][8];
#pragma unroll
      for (int c = 0; c < SL_COLS; ++c) {
        const int col = n0 + c < n ? n0 + c : n - 1;
        unpack8(*reinterpret_cast<const uint4*>(w + (long long)col * w_ld + kc), wv[c]);
      }
#pragma unroll
      for (int r = 0; r < SL_ROWS; ++r) {
        float xv[8];
        unpack8(*reinterpret_cast<const uint4*>(xs + r * k + kc), xv);
#pragma unroll
        for (int c = 0; c < SL_COLS; ++c)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[c][r] = fmaf(wv[c][i], xv[i], acc[c][r]);
      }
    }
#pragma unroll
    for (int c = 0; c < SL_COLS; ++c)
#pragma unroll
      for (int r = 0; r < SL_ROWS; ++r) acc[c][r] = warp_sum(acc[c][r]);
    if (lane < SL_COLS * SL_ROWS) {
      const int c = lane / SL_ROWS, r = lane - c * SL_ROWS;
      // select acc[c][r] without dynamic register indexing
      float v = 0.f;
#pragma unroll
      for (int cc = 0; cc < SL_COLS; ++cc)
#pragma unroll
        for (int rr = 0; rr < SL_ROWS; ++rr)
          if (cc == c && rr == r) v = acc[cc][rr];
      const int col = n0 + c;
      if (col < n && r < m) {
        if (bias) v += __bfloat162float(bias[col]);
        __nv_bfloat16* yp = y + (long long)r * y_ld + col;
        if (flags & AFB_SL_ACCUMULATE) v += __bfloat162float(*yp);
        *yp = __float2bfloat16_rn(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin] halves.
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out,
                                      int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * 128) return;
  const int r = i / 128, j = i - r * 128;
  const float freq = expf(-9.210340371976184f * float(j) / 128.0f);  // ln(10000)
  const float a = t[r] * freq;
  out[r * 256 + j] = __float2bfloat16_rn(cosf(a));
  out[r * 256 + 128 + j] = __float2bfloat16_rn(sinf(a));
}

// ------------------------------------------------------------------------------------------------
// ArcFlow sampler step, K = 16 components, packed-token layout (one warp per token).
// Fuses what the reference does in ~30 launches + 4 layout copies per NFE:
//   _unpack_latents / _unpack_mp      lakonlab/pipelines/arcflux_pipeline.py:135-193, :482-487
//   log_softmax over K (bf16)         lakonlab/models/architecture/arcflow/arcflux.py:246-247
//   ArcFlowPolicy                     lakonlab/models/diffusions/policies/arcflow.py:25-50
//   momentum_integration              lakonlab/pipelines/arcflux_pipeline.py:195-249
//   _pack_latents                     lakonlab/pipelines/arcflux_pipeline.py:506-510
// Token layout: output channel o = c*4 + ph*2 + pw; mixture weights / rates are per (k, j = o % 4).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float group_max(float v) {  // over lanes sharing lane % 4
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
  return v;
}
__device__ __forceinline__ float group_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
__device__ __forceinline__ float phi_expm1(float z, float eps) {
  const float sgn = z < 0.f ? -1.f : 1.f;  // sign(0) := +1 as in the reference
  const float zs = sgn * fmaxf(fabsf(z), eps);
  return expm1f(zs) / zs;
}

__global__ void __launch_bounds__(256)
sampler_step_k16_kernel(const __nv_bfloat16* __restrict__ head, long long head_ld,
                        const float* __restrict__ x_in, float* __restrict__ x_out,
                        __nv_bfloat16* __restrict__ x_out_bf16, long long tokens, float dt_past,
                        float dt_step, float eps) {
  constexpr int K = 16;
  __shared__ float sF[8][K * 4];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long tok = (long long)blockIdx.x * 8 + wib;
  if (tok >= tokens) return;
  const __nv_bfloat16* hrow = head + tok * head_ld;
  const __nv_bfloat16* logit = hrow + K * 64;
  const __nv_bfloat16* lgam = logit + K * 4;

  // entries e0 = lane (k 0..7), e1 = lane + 32 (k 8..15); j = lane % 4
  const float a0 = __bfloat162float(logit[lane]);
  const float a1 = __bfloat162float(logit[lane + 32]);
  // log_softmax over k in fp32, result rounded to bf16 (the network emits bf16)
  const float mx = group_max(fmaxf(a0, a1));
  const float se = group_sum(expf(a0 - mx) + expf(a1 - mx));
  const float lse = mx + logf(se);
  const float b0 = round_bf16(a0 - lse);
  const float b1 = round_bf16(a1 - lse);
  // softmax over k in fp32 (sampler side)
  const float mx2 = group_max(fmaxf(b0, b1));
  const float e0 = expf(b0 - mx2), e1 = expf(b1 - mx2);
  const float inv = 1.0f / group_sum(e0 + e1);
  const float w0 = e0 * inv, w1 = e1 * inv;
  // rates: component 0 has lambda == 0 (decay 1, phi 1)
  const float lam0 = lane >= 4 ? __bfloat162float(lgam[lane - 4]) : 0.f;
  const float lam1 = __bfloat162float(lgam[lane + 28]);
  float f0 = w0 * dt_step;
  if (lane >= 4) f0 *= expf(lam0 * dt_past) * phi_expm1(lam0 * dt_step, eps);
  const float f1 = w1 * dt_step * expf(lam1 * dt_past) * phi_expm1(lam1 * dt_step, eps);
  sF[wib][lane] = f0;
  sF[wib][lane + 32] = f1;
  __syncwarp();

  const int j0 = (2 * lane) & 3;  // channels 2*lane, 2*lane + 1 -> j0, j0 + 1
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t mv = *reinterpret_cast<const uint32_t*>(hrow + k * 64 + 2 * lane);
    acc0 = fmaf(bf16_lo(mv), sF[wib][k * 4 + j0], acc0);
    acc1 = fmaf(bf16_hi(mv), sF[wib][k * 4 + j0 + 1], acc1);
  }
  const float2 xi = *reinterpret_cast<const float2*>(x_in + tok * 64 + 2 * lane);
  const float2 xo = make_float2(xi.x - acc0, xi.y - acc1);
  *reinterpret_cast<float2*>(x_out + tok * 64 + 2 * lane) = xo;
  if (x_out_bf16)
    *reinterpret_cast<uint32_t*>(x_out_bf16 + tok * 64 + 2 * lane) = pack_bf16x2(xo.x, xo.y);
}

// ------------------------------------------------------------------------------------------------
// Plain RMSNorm over rows (Qwen-Image txt_norm, arcqwen.py:126-127): y = bf16(x * rsqrt(mean x^2 + eps)) * w
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rmsnorm_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                    const __nv_bfloat16* __restrict__ w, long long rows, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + row * dim;
  __nv_bfloat16* yr = y + row * dim;
  float ss = 0.f;
  for (int c = lane * 8; c < dim; c += 256) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
  }
  const float rs = rsqrtf(warp_sum(ss) / float(dim) + eps);
  for (int c = lane * 8; c < dim; c += 256) {
    float f[8], g[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + c), f);
    unpack8(*reinterpret_cast<const uint4*>(w + c), g);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = round_bf16(f[i] * rs) * g[i];
    *reinterpret_cast<uint4*>(yr + c) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------------
// Training-side policy evaluation with PER-SAMPLE times and an optional mixture-component dropout mask.
// One kernel, three modes, all in packed-token layout, all from the raw head tensor of one student call:
//   INTEGRATE  x_end = x_start - sum_k w_k mu_k e^{lam_k (s_src - s_start)} dt phi(lam_k dt)
//              (ArcFlowImitationBase.momentum_integration, lakonlab/models/diffusions/arcflow.py:28-79)
//   VELOCITY   u = sum_k w_k mu_k e^{lam_k (s_src - s_t)}      (ArcFlowPolicy.velocity, policies/arcflow.py:52-76)
//   AVERAGE_U  (x_start - x_end) / max(s_start - s_end, eps), or the local velocity where `small` is set
//              (policy_average_u_momentum, arcflow.py:81-110)
// drop[b][k] != 0 masks component k of sample b (ArcFlowPolicy.dropout_, policies/arcflow.py:96-106:
// logweights.masked_fill(mask, -inf) before the softmax).
// ------------------------------------------------------------------------------------------------
constexpr int POLICY_MAX_BATCH = 64;
struct PolicyParams {
  float dt_past[POLICY_MAX_BATCH];   // s_src - s_start
  float dt_step[POLICY_MAX_BATCH];   // s_start - s_end
  uint32_t drop[POLICY_MAX_BATCH];   // bit k: component k dropped
  uint8_t small[POLICY_MAX_BATCH];
};

__global__ void __launch_bounds__(256)
policy_eval_k16_kernel(const __nv_bfloat16* __restrict__ head, long long head_ld, const float* __restrict__ x_in,
                       float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16, int tokens_per_sample,
                       long long tokens, int mode, float eps, const __grid_constant__ PolicyParams pp) {
  constexpr int K = 16;
  __shared__ float sF[8][K * 4];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long tok = (long long)blockIdx.x * 8 + wib;
  if (tok >= tokens) return;
  const int b = int(tok / tokens_per_sample);
  const float dt_past = pp.dt_past[b], dt_step = pp.dt_step[b];
  const uint32_t drop = pp.drop[b];
  const bool local_u = mode == AFB_POLICY_VELOCITY || (mode == AFB_POLICY_AVERAGE_U && pp.small[b]);
  const __nv_bfloat16* hrow = head + tok * head_ld;
  const __nv_bfloat16* logit = hrow + K * 64;
  const __nv_bfloat16* lgam = logit + K * 4;

  const int k0 = lane >> 2, k1 = 8 + (lane >> 2);
  const float a0 = __bfloat162float(logit[lane]);
  const float a1 = __bfloat162float(logit[lane + 32]);
  const float mx = group_max(fmaxf(a0, a1));
  const float se = group_sum(expf(a0 - mx) + expf(a1 - mx));
  const float lse = mx + logf(se);
  float b0 = round_bf16(a0 - lse);
  float b1 = round_bf16(a1 - lse);
  if ((drop >> k0) & 1u) b0 = -INFINITY;
  if ((drop >> k1) & 1u) b1 = -INFINITY;
  const float mx2 = group_max(fmaxf(b0, b1));
  const float e0 = expf(b0 - mx2), e1 = expf(b1 - mx2);
  const float inv = 1.0f / group_sum(e0 + e1);
  const float w0 = e0 * inv, w1 = e1 * inv;
  const float lam0 = lane >= 4 ? __bfloat162float(lgam[lane - 4]) : 0.f;
  const float lam1 = __bfloat162float(lgam[lane + 28]);
  float f0, f1;
  if (local_u) {
    f0 = lane >= 4 ? w0 * expf(lam0 * dt_past) : w0;
    f1 = w1 * expf(lam1 * dt_past);
  } else {
    // AVERAGE_U divides the displacement by max(dt, eps): fold it into the factors
    const float scale = mode == AFB_POLICY_AVERAGE_U ? dt_step / fmaxf(dt_step, eps) : dt_step;
    f0 = w0 * scale;
    if (lane >= 4) f0 *= expf(lam0 * dt_past) * phi_expm1(lam0 * dt_step, eps);
    f1 = w1 * scale * expf(lam1 * dt_past) * phi_expm1(lam1 * dt_step, eps);
  }
  sF[wib][lane] = f0;
  sF[wib][lane + 32] = f1;
  __syncwarp();
  const int j0 = (2 * lane) & 3;
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t mv = *reinterpret_cast<const uint32_t*>(hrow + k * 64 + 2 * lane);
    acc0 = fmaf(bf16_lo(mv), sF[wib][k * 4 + j0], acc0);
    acc1 = fmaf(bf16_hi(mv), sF[wib][k * 4 + j0 + 1], acc1);
  }
  float2 o;
  if (mode == AFB_POLICY_INTEGRATE) {
    const float2 xi = *reinterpret_cast<const float2*>(x_in + tok * 64 + 2 * lane);
    o = make_float2(xi.x - acc0, xi.y - acc1);
  } else {
    o = make_float2(acc0, acc1);
  }
  *reinterpret_cast<float2*>(out + tok * 64 + 2 * lane) = o;
  if (out_bf16) *reinterpret_cast<uint32_t*>(out_bf16 + tok * 64 + 2 * lane) = pack_bf16x2(o.x, o.y);
}

// teacher targets are bf16 network outputs (FLUX) or their fp32 true-CFG combination (Qwen)
__device__ __forceinline__ float2 load_pair(const void* p, long long i, int is_f32) {
  if (is_f32) return *reinterpret_cast<const float2*>(static_cast<const float*>(p) + i);
  const uint32_t v = *reinterpret_cast<const uint32_t*>(static_cast<const __nv_bfloat16*>(p) + i);
  return make_float2(bf16_lo(v), bf16_hi(v));
}

// out = pos + (pos - neg) * (g - 1) on the two halves [neg; pos] of a batch-doubled bf16 network output (fp32 result):
// guidance_jit + forward_u (lakonlab/models/diffusions/gaussian_flow.py:18-26, 224-254)
__global__ void cfg_combine_kernel(const __nv_bfloat16* __restrict__ both, float* __restrict__ out, long long half, float gm1) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= half) return;
  const float2 n = load_pair(both, i, 0), q = load_pair(both, half + i, 0);
  *reinterpret_cast<float2*>(out + i) = make_float2(q.x + (q.x - n.x) * gm1, q.y + (q.y - n.y) * gm1);
}

// out[b, :] = x[b, :] + coef[b] * u[b, :]   (teacher Euler step, arcflow.py:190)
struct RowCoef {
  float c[POLICY_MAX_BATCH];
};
__global__ void axpy_rows_kernel(const float* __restrict__ x, const void* __restrict__ u, int u_f32,
                                 float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16, long long per_sample,
                                 long long total, const __grid_constant__ RowCoef rc) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= total) return;
  const float c = rc.c[int(i / per_sample)];
  const float2 xv = *reinterpret_cast<const float2*>(x + i);
  const float2 uv = load_pair(u, i, u_f32);
  const float2 o = make_float2(fmaf(c, uv.x, xv.x), fmaf(c, uv.y, xv.y));
  *reinterpret_cast<float2*>(out + i) = o;
  if (out_bf16) *reinterpret_cast<uint32_t*>(out_bf16 + i) = pack_bf16x2(o.x, o.y);
}

// per-sample mean over all elements of (pred - tgt)^2: mmgen mse_loss(reduction='flatmean')
__global__ void __launch_bounds__(256)
mse_rows_kernel(const float* __restrict__ pred, const void* __restrict__ tgt_all, int tgt_f32, float* __restrict__ out,
                long long per_sample) {
  __shared__ float red[8];
  const int b = blockIdx.y;
  const float* p = pred + (long long)b * per_sample;
  const long long t0 = (long long)b * per_sample;
  float acc = 0.f;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < per_sample;
       i += (long long)gridDim.x * blockDim.x * 2) {
    const float2 pv = *reinterpret_cast<const float2*>(p + i);
    const float2 tv = load_pair(tgt_all, t0 + i, tgt_f32);
    const float d0 = pv.x - tv.x, d1 = pv.y - tv.y;
    acc = fmaf(d0, d0, fmaf(d1, d1, acc));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
    v += __shfl_xor_sync(0xffu, v, 4);
    v += __shfl_xor_sync(0xffu, v, 2);
    v += __shfl_xor_sync(0xffu, v, 1);
    if (threadIdx.x == 0) atomicAdd(out + b, v / float(per_sample));
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of the velocity-matching loss through AVERAGE_U (the only policy path that carries grad in
// piid_segment_momentum, lakonlab/models/diffusions/arcflow.py:183-188):
//   L += coef/2 * sum_o (pred_o - tgt_o)^2,   pred = policy_average_u(head)   ->   dhead (+)= dL/dhead
// dhead rows mirror the head rows: d means | d logits (through softmax o bf16-round o log_softmax, rounding
// straight-through) | d loggamma. One warp per token, same work split as the forward kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dphi_expm1(float z) {  // d/dz [expm1(z)/z]
  if (fabsf(z) < 1e-2f) return 0.5f + z * (1.0f / 3.0f) + z * z * 0.125f;
  return (expm1f(z) * (z - 1.0f) + z) / (z * z);
}

__global__ void __launch_bounds__(256)
policy_avg_u_bwd_k16_kernel(const __nv_bfloat16* __restrict__ head, long long head_ld,
                            const void* __restrict__ tgt, int tgt_f32, float* __restrict__ dhead, long long dh_ld,
                            int tokens_per_sample, long long tokens, float coef, float eps, int accumulate,
                            const __grid_constant__ PolicyParams pp) {
  constexpr int K = 16;
  __shared__ float sWF[8][K * 4], sA[8][K * 4];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long tok = (long long)blockIdx.x * 8 + wib;
  if (tok >= tokens) return;
  const int b = int(tok / tokens_per_sample);
  const float dt_past = pp.dt_past[b], dt_step = pp.dt_step[b];
  const bool local_u = pp.small[b] != 0;
  const __nv_bfloat16* hrow = head + tok * head_ld;
  const __nv_bfloat16* logit = hrow + K * 64;
  const __nv_bfloat16* lgam = logit + K * 4;
  float* drow = dhead + tok * dh_ld;

  const float a0 = __bfloat162float(logit[lane]);
  const float a1 = __bfloat162float(logit[lane + 32]);
  const float mx = group_max(fmaxf(a0, a1));
  const float lse = mx + logf(group_sum(expf(a0 - mx) + expf(a1 - mx)));
  const float b0 = round_bf16(a0 - lse), b1 = round_bf16(a1 - lse);
  const float mx2 = group_max(fmaxf(b0, b1));
  const float e0 = expf(b0 - mx2), e1 = expf(b1 - mx2);
  const float inv = 1.0f / group_sum(e0 + e1);
  const float w0 = e0 * inv, w1 = e1 * inv;
  const bool has0 = lane >= 4;  // component 0 has no rate
  const float lam0 = has0 ? __bfloat162float(lgam[lane - 4]) : 0.f;
  const float lam1 = __bfloat162float(lgam[lane + 28]);
  float f0, f1, df0, df1;  // F and dF/dlambda
  if (local_u) {
    f0 = has0 ? expf(lam0 * dt_past) : 1.0f;
    f1 = expf(lam1 * dt_past);
    df0 = has0 ? dt_past * f0 : 0.f;
    df1 = dt_past * f1;
  } else {
    const float scale = dt_step / fmaxf(dt_step, eps);
    auto eval = [&](float lam, float& f, float& df) {
      const float z = lam * dt_step;
      const float sgn = z < 0.f ? -1.f : 1.f;
      const float zs = sgn * fmaxf(fabsf(z), eps);
      const float decay = expf(lam * dt_past);
      f = decay * scale * (expm1f(zs) / zs);
      df = dt_past * f + (fabsf(z) >= eps ? decay * scale * dphi_expm1(zs) * dt_step : 0.f);
    };
    eval(lam1, f1, df1);
    if (has0) {
      eval(lam0, f0, df0);
    } else {
      f0 = scale;
      df0 = 0.f;
    }
  }
  sWF[wib][lane] = w0 * f0;
  sWF[wib][lane + 32] = w1 * f1;
  __syncwarp();

  const int j0 = (2 * lane) & 3;
  float mu0[K], mu1[K];
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t mv = *reinterpret_cast<const uint32_t*>(hrow + k * 64 + 2 * lane);
    mu0[k] = bf16_lo(mv);
    mu1[k] = bf16_hi(mv);
    acc0 = fmaf(mu0[k], sWF[wib][k * 4 + j0], acc0);
    acc1 = fmaf(mu1[k], sWF[wib][k * 4 + j0 + 1], acc1);
  }
  const float2 tv = load_pair(tgt, tok * 64 + 2 * lane, tgt_f32);
  const float g0 = coef * (acc0 - tv.x), g1 = coef * (acc1 - tv.y);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float2 dm = make_float2(sWF[wib][k * 4 + j0] * g0, sWF[wib][k * 4 + j0 + 1] * g1);
    float2* dst = reinterpret_cast<float2*>(drow + k * 64 + 2 * lane);
    if (accumulate) {
      const float2 old = *dst;
      dm.x += old.x;
      dm.y += old.y;
    }
    *dst = dm;
    // A[k][j] = sum over the 16 channels of mu * g; lanes of equal parity share (j0, j0 + 1)
    float p0 = mu0[k] * g0, p1 = mu1[k] * g1;
#pragma unroll
    for (int o = 2; o < 32; o <<= 1) {
      p0 += __shfl_xor_sync(0xffffffffu, p0, o);
      p1 += __shfl_xor_sync(0xffffffffu, p1, o);
    }
    if (lane < 2) {
      sA[wib][k * 4 + j0] = p0;
      sA[wib][k * 4 + j0 + 1] = p1;
    }
  }
  __syncwarp();
  const float A0 = sA[wib][lane], A1 = sA[wib][lane + 32];
  const float dw0 = f0 * A0, dw1 = f1 * A1;
  const float sdot = group_sum(w0 * dw0 + w1 * dw1);
  float dl0 = w0 * (dw0 - sdot), dl1 = w1 * (dw1 - sdot);
  float dg0 = w0 * A0 * df0, dg1 = w1 * A1 * df1;
  float* dlog = drow + K * 64;
  float* dgam = dlog + K * 4;
  if (accumulate) {
    dl0 += dlog[lane];
    dl1 += dlog[lane + 32];
    if (has0) dg0 += dgam[lane - 4];
    dg1 += dgam[lane + 28];
  }
  dlog[lane] = dl0;
  dlog[lane + 32] = dl1;
  if (has0) dgam[lane - 4] = dg0;
  dgam[lane + 28] = dg1;
}

// out[n] (+)= sum_t x[t, n]   (bias gradients); x fp32 [rows, ld], out fp32 [n]
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ x, long long ld, float* __restrict__ out, long long rows, int n,
                  int rows_per_block) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += x[r * ld + col];
  atomicAdd(out + col, acc);
}

// ------------------------------------------------------------------------------------------------
// Parameter gradients of y = LN(x) * (1 + scale[b]) + shift[b]  (AdaLayerNormContinuous / Zero):
//   dscale[b, d] += sum_rows dy * xhat,   dshift[b, d] += sum_rows dy,   xhat = (x - mean) * rstd
// Step 1: per-row (mean, rstd); step 2: column reduction over row chunks with fp32 atomics.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ln_row_stats_kernel(const __nv_bfloat16* __restrict__ x, long long x_bs, float2* __restrict__ stats, int batches,
                    int rows_per_batch, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)batches * rows_per_batch) return;
  const int b = int(row / rows_per_batch);
  const __nv_bfloat16* xr = x + (long long)b * x_bs + (row - (long long)b * rows_per_batch) * dim;
  float sum = 0.f;
  for (int c = lane * 8; c < dim; c += 256) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += f[i];
  }
  const float mean = warp_sum(sum) / float(dim);
  float sq = 0.f;
  for (int c = lane * 8; c < dim; c += 256) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) sq += (f[i] - mean) * (f[i] - mean);
  }
  const float rstd = rsqrtf(warp_sum(sq) / float(dim) + eps);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}

__global__ void __launch_bounds__(256)
ln_mod_param_grad_kernel(const __nv_bfloat16* __restrict__ x, long long x_bs, const __nv_bfloat16* __restrict__ dy,
                         long long dy_bs, const float2* __restrict__ stats, float* __restrict__ dscale,
                         float* __restrict__ dshift, long long out_bs, int rows_per_batch, int dim,
                         int chunks_per_batch, int rows_per_chunk) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (col >= dim) return;
  const int b = blockIdx.y / chunks_per_batch;
  const int r0 = (blockIdx.y - b * chunks_per_batch) * rows_per_chunk;
  const int r1 = min(rows_per_batch, r0 + rows_per_chunk);
  const __nv_bfloat16* xb = x + (long long)b * x_bs;
  const __nv_bfloat16* db = dy + (long long)b * dy_bs;
  float s0 = 0.f, s1 = 0.f, h0 = 0.f, h1 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float2 st = stats[(long long)b * rows_per_batch + r];
    const uint32_t xv = *reinterpret_cast<const uint32_t*>(xb + (long long)r * dim + col);
    const uint32_t dv = *reinterpret_cast<const uint32_t*>(db + (long long)r * dim + col);
    const float d0 = bf16_lo(dv), d1 = bf16_hi(dv);
    s0 = fmaf(d0, (bf16_lo(xv) - st.x) * st.y, s0);
    s1 = fmaf(d1, (bf16_hi(xv) - st.x) * st.y, s1);
    h0 += d0;
    h1 += d1;
  }
  atomicAdd(dscale + (long long)b * out_bs + col, s0);
  atomicAdd(dscale + (long long)b * out_bs + col + 1, s1);
  atomicAdd(dshift + (long long)b * out_bs + col, h0);
  atomicAdd(dshift + (long long)b * out_bs + col + 1, h1);
}

// Gradients of a batch-row Linear  e[b, j] = sum_d W[j, d] * act(t[b, d]) + bias[j]  (the AdaLN modulation Linears):
//   dW[j, d] += sum_b de[b, j] * act(t[b, d]),   dbias[j] += sum_b de[b, j];   act = SiLU when silu != 0
__global__ void __launch_bounds__(256)
rowlinear_param_grad_kernel(const float* __restrict__ de, long long de_ld, const __nv_bfloat16* __restrict__ t,
                            long long t_ld, float* __restrict__ dw, long long dw_ld, float* __restrict__ dbias, int m,
                            int n_out, int k_in, int silu_in) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (d >= k_in) return;
  float acc = 0.f, bsum = 0.f;
  for (int b = 0; b < m; ++b) {
    float a = __bfloat162float(t[(long long)b * t_ld + d]);
    if (silu_in) a = round_bf16(silu(a));
    const float g = de[(long long)b * de_ld + j];
    acc = fmaf(g, a, acc);
    bsum += g;
  }
  dw[(long long)j * dw_ld + d] += acc;
  if (d == 0 && dbias) dbias[j] += bsum;
}



// ================================================================================================
// LoRA input dropout (peft lora_dropout = 0.05, train only: result = base(x) + B(A(dropout(x)))).
// Counter-based: one 32-bit hash per PAIR of consecutive logical elements, 16 bits each:
//   keep(idx) = half16(lowbias32(uint32(idx >> 1) ^ key ^ hi(idx)), idx & 1) >= round(p * 2^16)
// so the checkpointed recompute and the backward regenerate the same mask from (seed, layer id) with no stored mask. The
// hash is restated in numpy by the oracle. (Round 1 hashed every element twice: ~22 integer instructions per element made
// this kernel ALU-bound at a quarter of the HBM rate; a pair per hash is ~9.)
// ================================================================================================
__device__ __forceinline__ uint32_t lowbias32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x7feb352dU;
  h ^= h >> 15;
  h *= 0x846ca68bU;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t dropout_pair_bits(uint32_t key, unsigned long long pair) {
  return lowbias32(uint32_t(pair) ^ key ^ (uint32_t(pair >> 32) * 0x9E3779B1U));
}
__device__ __forceinline__ bool dropout_keep(uint32_t key, unsigned long long idx, uint32_t thresh16) {
  const uint32_t h = dropout_pair_bits(key, idx >> 1);
  return ((idx & 1ull) ? (h >> 16) : (h & 0xFFFFu)) >= thresh16;
}
__host__ __device__ __forceinline__ uint32_t dropout_thresh16(float p) { return uint32_t(double(p) * 65536.0 + 0.5); }

struct DropParams {
  int rows_per_batch, cols, logical_cols, col0;
  uint32_t key, thresh;
  float inv_keep;
  int silu_in;      // forward: apply SiLU (rounded to bf16, as small_linear's SILU_IN does) before the mask
  int accumulate;   // backward: out += mask * x / keep instead of out = ...
};

// grid = (column blocks of 256 chunks x 8 elements, groups of DROP_ROWS rows): no per-thread index division (the first
// version derived row and column from a flat 64-bit index with two 64-bit divisions per 16-byte chunk, which cost more than
// its memory traffic), and DROP_ROWS independent 16-byte loads in flight per thread.
constexpr int DROP_ROWS = 4;
__global__ void __launch_bounds__(256)
dropout_rows_kernel(const __nv_bfloat16* __restrict__ x, long long x_ld, long long x_bs, __nv_bfloat16* __restrict__ out,
                    long long out_ld, long long out_bs, int total_rows, const DropParams p) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (c >= p.cols) return;
  for (int row0 = blockIdx.y * DROP_ROWS; row0 < total_rows; row0 += gridDim.y * DROP_ROWS) {
    uint4 xin[DROP_ROWS], acc[DROP_ROWS];
    __nv_bfloat16* op[DROP_ROWS];
#pragma unroll
    for (int q = 0; q < DROP_ROWS; ++q) {
      const int row = min(row0 + q, total_rows - 1);   // clamped: the surplus rows of the last group are loaded, not stored
      const int b = row / p.rows_per_batch;
      const int r = row - b * p.rows_per_batch;
      xin[q] = *reinterpret_cast<const uint4*>(x + (long long)b * x_bs + (long long)r * x_ld + c);
      op[q] = out + (long long)b * out_bs + (long long)r * out_ld + c;
      if (p.accumulate) acc[q] = *reinterpret_cast<const uint4*>(op[q]);
    }
#pragma unroll
    for (int q = 0; q < DROP_ROWS; ++q) {
      const int row = row0 + q;
      if (row >= total_rows) break;
      float v[8], o[8];
      unpack8(xin[q], v);
      if (p.accumulate) unpack8(acc[q], o);
      const unsigned long long pair0 = ((unsigned long long)row * p.logical_cols + p.col0 + c) >> 1;  // even: all three are
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const uint32_t h = dropout_pair_bits(p.key, pair0 + (k >> 1));
        float a0 = v[k], a1 = v[k + 1];
        if (p.silu_in) {
          a0 = round_bf16(silu(a0));
          a1 = round_bf16(silu(a1));
        }
        const float m0 = (h & 0xFFFFu) >= p.thresh ? a0 * p.inv_keep : 0.f;
        const float m1 = (h >> 16) >= p.thresh ? a1 * p.inv_keep : 0.f;
        o[k] = p.accumulate ? o[k] + m0 : m0;
        o[k + 1] = p.accumulate ? o[k + 1] + m1 : m1;
      }
      *reinterpret_cast<uint4*>(op[q]) = pack8(o);
    }
  }
}

// ================================================================================================
// Modulation-vector gradients (the only path to the timestep-embedder LoRA)
// ================================================================================================
// du = gate[b] (.) dh   and   dgate[b, :] += sum_rows dh (.) u      (u = the branch output the gate multiplied)
__global__ void __launch_bounds__(128)
gate_bwd_kernel(const __nv_bfloat16* __restrict__ dh, long long dh_bs, const __nv_bfloat16* __restrict__ u, long long u_bs,
                const __nv_bfloat16* __restrict__ gate, long long gate_bs, __nv_bfloat16* __restrict__ du, long long du_bs,
                float* __restrict__ dgate, long long dgate_bs, int rows_per_batch, int cols, int chunks_per_batch,
                int rows_per_chunk) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (c >= cols) return;
  const int b = blockIdx.y / chunks_per_batch;
  const int r0 = (blockIdx.y - b * chunks_per_batch) * rows_per_chunk;
  const int r1 = min(rows_per_batch, r0 + rows_per_chunk);
  float g[8], acc[8];
  unpack8(*reinterpret_cast<const uint4*>(gate + (long long)b * gate_bs + c), g);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int r = r0; r < r1; ++r) {
    float d[8], uv[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(dh + (long long)b * dh_bs + (long long)r * cols + c), d);
    unpack8(*reinterpret_cast<const uint4*>(u + (long long)b * u_bs + (long long)r * cols + c), uv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] = fmaf(d[i], uv[i], acc[i]);
      o[i] = d[i] * g[i];
    }
    *reinterpret_cast<uint4*>(du + (long long)b * du_bs + (long long)r * cols + c) = pack8(o);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(dgate + (long long)b * dgate_bs + c + i, acc[i]);
}

// out = res + gate[b] (.) u   (the recompute's unfused form of the GEMM's gate+residual epilogue)
__global__ void __launch_bounds__(256)
gate_res_kernel(const __nv_bfloat16* __restrict__ res, long long res_bs, const __nv_bfloat16* __restrict__ u, long long u_bs,
                const __nv_bfloat16* __restrict__ gate, long long gate_bs, __nv_bfloat16* __restrict__ out, long long out_bs,
                int rows_per_batch, int cols, unsigned total_chunks) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;   // 32-bit index math (launcher checks the range): one cheap division
  if (i >= total_chunks) return;
  const unsigned cpr = unsigned(cols) / 8u;
  const unsigned row32 = i / cpr;
  const int c = int(i - row32 * cpr) * 8;
  const long long row = row32;
  const int b = int(row32 / unsigned(rows_per_batch));
  const long long r = row - (long long)b * rows_per_batch;
  float rv[8], uv[8], gv[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(res + (long long)b * res_bs + r * cols + c), rv);
  unpack8(*reinterpret_cast<const uint4*>(u + (long long)b * u_bs + r * cols + c), uv);
  unpack8(*reinterpret_cast<const uint4*>(gate + (long long)b * gate_bs + c), gv);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = fmaf(gv[k], uv[k], rv[k]);
  *reinterpret_cast<uint4*>(out + (long long)b * out_bs + r * cols + c) = pack8(o);
}

// out[b, n] += sum_j de[b, j] * W[j, n]   (dX of a batch-row Linear: W bf16 [J, N] row-major, streamed once; m <= 8)
constexpr int RDX_JC = 64;
__global__ void __launch_bounds__(128)
rowlinear_dx_kernel(const float* __restrict__ de, long long de_ld, const __nv_bfloat16* __restrict__ w, long long w_ld,
                    float* __restrict__ out, long long out_ld, int m, int J, int N) {
  __shared__ float sde[8][RDX_JC];
  const int j0 = blockIdx.y * RDX_JC;
  const int jn = min(RDX_JC, J - j0);
  for (int i = threadIdx.x; i < 8 * RDX_JC; i += blockDim.x) {
    const int b = i / RDX_JC, j = i - b * RDX_JC;
    sde[b][j] = (b < m && j < jn) ? de[(long long)b * de_ld + j0 + j] : 0.f;
  }
  __syncthreads();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (c >= N) return;
  float acc[8][8];
#pragma unroll
  for (int b = 0; b < 8; ++b)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[b][i] = 0.f;
  for (int j = 0; j < jn; ++j) {
    float wv[8];
    unpack8(*reinterpret_cast<const uint4*>(w + (long long)(j0 + j) * w_ld + c), wv);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const float g = sde[b][j];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[b][i] = fmaf(g, wv[i], acc[b][i]);
    }
  }
  for (int b = 0; b < m; ++b)
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(out + (long long)b * out_ld + c + i, acc[b][i]);
}

// d[i] *= silu'(x[i])   (fp32 gradient, bf16 pre-activation)
__global__ void __launch_bounds__(256)
silu_bwd_kernel(float* __restrict__ d, long long d_ld, const __nv_bfloat16* __restrict__ x, long long x_ld, int rows, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= cols || r >= rows) return;
  const float v = __bfloat162float(x[(long long)r * x_ld + c]);
  const float sg = 1.0f / (1.0f + __expf(-v));
  d[(long long)r * d_ld + c] *= sg * (1.0f + v * (1.0f - sg));
}

// ================================================================================================
// Backward of the streaming ops (adapter-only backward through the FROZEN trunk: only activation
// gradients are needed here; the modulation vectors are treated as constants — see DESIGN.md).
// ================================================================================================
// dh[b, r, :] += LNmod_bwd(dy):  g = dy * (1 + scale[b]);  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
template <int NCH>
__global__ void __launch_bounds__(256)
ln_modulate_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long x_bs, const __nv_bfloat16* __restrict__ dy,
                       long long dy_bs, __nv_bfloat16* __restrict__ dh, long long dh_bs,
                       const __nv_bfloat16* __restrict__ scale, long long mod_bs, int batches, int rows_per_batch,
                       float eps, int accumulate) {
  constexpr int DIM = NCH * 256;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)batches * rows_per_batch) return;
  const int b = int(row / rows_per_batch);
  const long long r = row - (long long)b * rows_per_batch;
  const __nv_bfloat16* xr = x + (long long)b * x_bs + r * DIM;
  const __nv_bfloat16* dr = dy + (long long)b * dy_bs + r * DIM;
  __nv_bfloat16* hr = dh + (long long)b * dh_bs + r * DIM;
  const __nv_bfloat16* sc = scale + (long long)b * mod_bs;
  float v[NCH][8], g[NCH][8];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    unpack8(*reinterpret_cast<const uint4*>(xr + c * 256 + lane * 8), v[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += v[c][i];
  }
  const float mean = warp_sum(sum) * (1.0f / DIM);
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[c][i] -= mean;
      sq += v[c][i] * v[c][i];
    }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / DIM) + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float s8[8];
    unpack8(*reinterpret_cast<const uint4*>(dr + c * 256 + lane * 8), g[c]);
    unpack8(*reinterpret_cast<const uint4*>(sc + c * 256 + lane * 8), s8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[c][i] *= rstd;  // xhat
      g[c][i] *= 1.0f + s8[i];
      sg += g[c][i];
      sgx = fmaf(g[c][i], v[c][i], sgx);
    }
  }
  const float mg = warp_sum(sg) * (1.0f / DIM), mgx = warp_sum(sgx) * (1.0f / DIM);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = rstd * (g[c][i] - mg - v[c][i] * mgx);
    if (accumulate) {
      float old[8];
      unpack8(*reinterpret_cast<const uint4*>(hr + c * 256 + lane * 8), old);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += old[i];
    }
    *reinterpret_cast<uint4*>(hr + c * 256 + lane * 8) = pack8(o);
  }
}

// out[b, r, :] = vec[b, :] * x[b, r, :]   (du = gate (.) dh'), generic leading dims
__global__ void __launch_bounds__(256)
rowscale_kernel(const __nv_bfloat16* __restrict__ x, long long x_ld, long long x_bs, const __nv_bfloat16* __restrict__ vec,
                long long vec_bs, __nv_bfloat16* __restrict__ out, long long out_ld, long long out_bs, int rows_per_batch,
                int cols, unsigned total_chunks) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;   // 32-bit index math (launcher checks the range): one cheap division
  if (i >= total_chunks) return;
  const unsigned cpr = unsigned(cols) / 8u;
  const unsigned row32 = i / cpr;
  const int c = int(i - row32 * cpr) * 8;
  const long long row = row32;
  const int b = int(row32 / unsigned(rows_per_batch));
  const long long r = row - (long long)b * rows_per_batch;
  float xv[8], gv[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(x + (long long)b * x_bs + r * x_ld + c), xv);
  unpack8(*reinterpret_cast<const uint4*>(vec + (long long)b * vec_bs + c), gv);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = xv[k] * gv[k];
  *reinterpret_cast<uint4*>(out + (long long)b * out_bs + r * out_ld + c) = pack8(o);
}

// dpre = dm * gelu_tanh'(pre)  (in place on dm), generic leading dims, rows x cols
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(__nv_bfloat16* __restrict__ dm, long long dm_ld, const __nv_bfloat16* __restrict__ pre, long long pre_ld,
                int cols, unsigned total_chunks) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;   // 32-bit index math (launcher checks the range): one cheap division
  if (i >= total_chunks) return;
  const unsigned cpr = unsigned(cols) / 8u;
  const unsigned row32 = i / cpr;
  const int c = int(i - row32 * cpr) * 8;
  const long long row = row32;
  float d[8], x[8];
  unpack8(*reinterpret_cast<const uint4*>(dm + row * dm_ld + c), d);
  unpack8(*reinterpret_cast<const uint4*>(pre + row * pre_ld + c), x);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float t = tanhf(k0 * (x[k] + k1 * x[k] * x[k] * x[k]));
    d[k] *= 0.5f * (1.0f + t) + 0.5f * x[k] * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x[k] * x[k]);
  }
  *reinterpret_cast<uint4*>(dm + row * dm_ld + c) = pack8(d);
}

// out = gelu_tanh(pre) (the recompute's copy of the GEMM's fused GELU epilogue; tanh.approx like gemm.cu)
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ pre, long long pre_ld, __nv_bfloat16* __restrict__ out, long long out_ld,
                int cols, unsigned total_chunks) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;   // 32-bit index math (launcher checks the range): one cheap division
  if (i >= total_chunks) return;
  const unsigned cpr = unsigned(cols) / 8u;
  const unsigned row32 = i / cpr;
  const int c = int(i - row32 * cpr) * 8;
  const long long row = row32;
  float x[8];
  unpack8(*reinterpret_cast<const uint4*>(pre + row * pre_ld + c), x);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k0 * (x[k] + k1 * x[k] * x[k] * x[k])));
    x[k] = 0.5f * x[k] * (1.0f + t);
  }
  *reinterpret_cast<uint4*>(out + row * out_ld + c) = pack8(x);
}

// In place on the q|k columns of dqkv: d(out) -> d(raw) through RoPE^T and the per-head RMSNorm (weights frozen).
// raw: the projection output BEFORE norm/rope (saved by the recompute). Same work split as rmsnorm_rope_kernel.
__global__ void __launch_bounds__(256)
rmsnorm_rope_bwd_kernel(__nv_bfloat16* __restrict__ dqkv, const __nv_bfloat16* __restrict__ raw, long long ld, long long bs,
                        int q_off, int k_off, int batches, int seq, int heads, int txt_rows,
                        const __nv_bfloat16* __restrict__ wq_txt, const __nv_bfloat16* __restrict__ wk_txt,
                        const __nv_bfloat16* __restrict__ wq_img, const __nv_bfloat16* __restrict__ wk_img,
                        const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, float eps) {
  const int lane = threadIdx.x & 31;
  const int hl = lane & 15, half = lane >> 4;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)batches * seq) return;
  const int b = int(row / seq);
  const int s = int(row - (long long)b * seq);
  const long long base = (long long)b * bs + (long long)s * ld;
  const bool is_txt = s < txt_rows;
  float cs[8], sn[8], wq[8], wk[8];
  {
    const float4 c0 = *reinterpret_cast<const float4*>(cos_tab + (long long)s * 128 + hl * 8);
    const float4 c1 = *reinterpret_cast<const float4*>(cos_tab + (long long)s * 128 + hl * 8 + 4);
    const float4 s0 = *reinterpret_cast<const float4*>(sin_tab + (long long)s * 128 + hl * 8);
    const float4 s1 = *reinterpret_cast<const float4*>(sin_tab + (long long)s * 128 + hl * 8 + 4);
    cs[0] = c0.x; cs[1] = c0.y; cs[2] = c0.z; cs[3] = c0.w; cs[4] = c1.x; cs[5] = c1.y; cs[6] = c1.z; cs[7] = c1.w;
    sn[0] = s0.x; sn[1] = s0.y; sn[2] = s0.z; sn[3] = s0.w; sn[4] = s1.x; sn[5] = s1.y; sn[6] = s1.z; sn[7] = s1.w;
    unpack8(*reinterpret_cast<const uint4*>((is_txt ? wq_txt : wq_img) + hl * 8), wq);
    unpack8(*reinterpret_cast<const uint4*>((is_txt ? wk_txt : wk_img) + hl * 8), wk);
  }
  const int slots = 2 * heads;
  for (int s0 = 0; s0 < slots; s0 += 2) {
    const int slot = s0 + half;
    const bool act = slot < slots, is_k = slot >= heads;
    const int h = is_k ? slot - heads : slot;
    const long long off = base + (is_k ? k_off : q_off) + h * 128 + hl * 8;
    float d[8], x[8];
    unpack8(act ? *reinterpret_cast<const uint4*>(dqkv + off) : make_uint4(0, 0, 0, 0), d);
    unpack8(act ? *reinterpret_cast<const uint4*>(raw + off) : make_uint4(0, 0, 0, 0), x);
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) ss = fmaf(x[e], x[e], ss);
    ss += __shfl_xor_sync(0xffffffffu, ss, 8);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    const float rs = rsqrtf(ss * (1.0f / 128.0f) + eps);
    float dyv[8], dot = 0.f;  // dy = d(x * rs) = RoPE^T(d) * w
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const float da = d[e] * cs[e] + d[e + 1] * sn[e + 1];
      const float db = -d[e] * sn[e] + d[e + 1] * cs[e + 1];
      dyv[e] = da * (is_k ? wk[e] : wq[e]);
      dyv[e + 1] = db * (is_k ? wk[e + 1] : wq[e + 1]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) dot = fmaf(dyv[e], x[e] * rs, dot);
    dot += __shfl_xor_sync(0xffffffffu, dot, 8);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot *= (1.0f / 128.0f);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = rs * (dyv[e] - x[e] * rs * dot);
    if (act) *reinterpret_cast<uint4*>(dqkv + off) = pack8(o);
  }
}

// delta[b, h, s] = sum_d dO[b, s, h, d] * O[b, s, h, d]   (attention backward pre-pass); one warp per (row, head)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long o_ld, long long o_bs, const __nv_bfloat16* __restrict__ d_o,
                  long long do_ld, long long do_bs, float* __restrict__ delta, int batch, int seq, int heads) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)batch * seq * heads) return;
  const int h = int(w % heads);
  const long long row = w / heads;
  const int b = int(row / seq);
  const int s = int(row - (long long)b * seq);
  const uint2 ov = *reinterpret_cast<const uint2*>(o + (long long)b * o_bs + (long long)s * o_ld + h * 128 + lane * 4);
  const uint2 dv = *reinterpret_cast<const uint2*>(d_o + (long long)b * do_bs + (long long)s * do_ld + h * 128 + lane * 4);
  float acc = bf16_lo(ov.x) * bf16_lo(dv.x) + bf16_hi(ov.x) * bf16_hi(dv.x) + bf16_lo(ov.y) * bf16_lo(dv.y) +
              bf16_hi(ov.y) * bf16_hi(dv.y);
  acc = warp_sum(acc);
  if (lane == 0) delta[((long long)b * heads + h) * seq + s] = acc;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                     long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i) = o;
  } else {
    for (long long e = i; e < n; ++e) out[e] = __float2bfloat16_rn(in[e]);
  }
}

__global__ void fill_f32_kernel(float* p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

int ln_modulate_launch(const void* x, int64_t x_bs, void* y, int64_t y_bs, const void* scale,
                       const void* shift, int64_t mod_bs, int batches, int rows_per_batch, int dim,
                       float eps, cudaStream_t stream) {
  AFB_REQUIRE(x && y && scale && shift, "ln_modulate: null pointer");
  AFB_REQUIRE(batches >= 1 && rows_per_batch >= 1, "ln_modulate: empty input");
  AFB_REQUIRE(dim % 256 == 0, "ln_modulate: dim=%d must be a multiple of 256", dim);
  const long long rows = (long long)batches * rows_per_batch;
  const int wpb = 8;
  const unsigned grid = unsigned((rows + wpb - 1) / wpb);
#define AFB_LN_CASE(NCH)                                                                          \
  case NCH:                                                                                       \
    ln_modulate_kernel<NCH><<<grid, wpb * 32, 0, stream>>>(                                       \
        static_cast<const __nv_bfloat16*>(x), x_bs, static_cast<__nv_bfloat16*>(y), y_bs,         \
        static_cast<const __nv_bfloat16*>(scale), static_cast<const __nv_bfloat16*>(shift),       \
        mod_bs, batches, rows_per_batch, eps);                                                    \
    break;
  switch (dim / 256) {
    AFB_LN_CASE(1)
    AFB_LN_CASE(2)
    AFB_LN_CASE(4)
    AFB_LN_CASE(8)
    AFB_LN_CASE(12)
    AFB_LN_CASE(16)
    default:
      set_last_error("ln_modulate: unsupported dim %d", dim);
      return AFB_ERR_UNSUPPORTED;
  }
#undef AFB_LN_CASE
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

// Rotary table for the fused QK epilogue, laid out for the way the GEMM epilogue reads it: one thread per output row, 32
// consecutive rows per warp, 32 (cos, sin) pairs per 64-column chunk. out[((s >> 5) * 2 + half) * 16 + i][s & 31] (float4)
// = pairs (32 half + 2 i, 32 half + 2 i + 1) of position s as (cos, sin, cos, sin): for a fixed (block of 32 positions, half,
// i) the 32 lanes of a warp read 512 contiguous bytes. (A row-major [positions, 64, 2] table made every lane read its own
// 512-byte row: 32 scattered sectors per load instruction, and the QKV launches were 35 % slower than with the plain epilogue.)
__global__ void __launch_bounds__(256) rope_pack_kernel(const float* __restrict__ c, const float* __restrict__ sn,
                                                         float4* __restrict__ out, long long rows, long long rows_pad) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 of the output
  if (idx >= rows_pad * 32) return;
  const int lane = int(idx & 31);
  const long long g = idx >> 5;          // (block * 2 + half) * 16 + i
  const int i = int(g & 15);
  const int half = int((g >> 4) & 1);
  const long long s = ((g >> 5) << 5) + lane;
  float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
  if (s < rows) {
    const int p0 = half * 32 + 2 * i;    // pair index; the source tables repeat every value twice
    v = make_float4(c[s * 128 + 2 * p0], sn[s * 128 + 2 * p0], c[s * 128 + 2 * p0 + 2], sn[s * 128 + 2 * p0 + 2]);
  }
  out[idx] = v;
}

int rope_pack_launch(const float* cos_tab, const float* sin_tab, float* out, int64_t rows, cudaStream_t stream) {
  AFB_REQUIRE(cos_tab && sin_tab && out && rows >= 1, "rope_pack: bad arguments");
  const long long rows_pad = (rows + 31) / 32 * 32;
  const long long n4 = rows_pad * 32;
  rope_pack_kernel<<<unsigned((n4 + 255) / 256), 256, 0, stream>>>(cos_tab, sin_tab, reinterpret_cast<float4*>(out), rows, rows_pad);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int rmsnorm_rope_launch(void* qkv, int64_t ld, int64_t bs, int q_off, int k_off, int batches, int seq,
                        int heads, int txt_rows, const void* wq_txt, const void* wk_txt,
                        const void* wq_img, const void* wk_img, const float* cos_tab,
                        const float* sin_tab, float eps, cudaStream_t stream) {
  AFB_REQUIRE(qkv && wq_img && wk_img && cos_tab && sin_tab, "rmsnorm_rope: null pointer");
  AFB_REQUIRE(txt_rows == 0 || (wq_txt && wk_txt), "rmsnorm_rope: text norm weights missing");
  AFB_REQUIRE(ld % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0,
              "rmsnorm_rope: q/k columns must be 16-byte aligned");
  AFB_REQUIRE(batches >= 1 && seq >= 1 && heads >= 1, "rmsnorm_rope: empty input");
  const long long rows = (long long)batches * seq;
  const int wpb = 8;
  const unsigned grid = unsigned((rows + wpb - 1) / wpb);
  rmsnorm_rope_kernel<<<grid, wpb * 32, 0, stream>>>(
      static_cast<__nv_bfloat16*>(qkv), ld, bs, q_off, k_off, batches, seq, heads, txt_rows,
      static_cast<const __nv_bfloat16*>(wq_txt ? wq_txt : wq_img),
      static_cast<const __nv_bfloat16*>(wk_txt ? wk_txt : wk_img),
      static_cast<const __nv_bfloat16*>(wq_img), static_cast<const __nv_bfloat16*>(wk_img), cos_tab,
      sin_tab, eps);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int small_linear_launch(const void* x, int64_t x_ld, const void* w, int64_t w_ld, const void* bias,
                        void* y, int64_t y_ld, int m, int n, int k, int flags, cudaStream_t stream) {
  AFB_REQUIRE(x && w && y, "small_linear: null pointer");
  AFB_REQUIRE(m >= 1 && m <= SL_ROWS, "small_linear: m=%d must be in [1, %d]", m, SL_ROWS);
  AFB_REQUIRE(n >= 1, "small_linear: n=%d", n);
  AFB_REQUIRE(k >= 256 && k % 256 == 0 && k <= 4096, "small_linear: k=%d must be a multiple of 256 <= 4096", k);
  AFB_REQUIRE(w_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
              "small_linear: W must be 16-byte aligned with ld %% 8 == 0");
  const size_t smem = size_t(SL_ROWS) * k * 2;
  static bool attr_set = false;
  if (!attr_set) {
    AFB_CHECK_CUDA(cudaFuncSetAttribute(small_linear_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096 * 2));
    attr_set = true;
  }
  const int groups = (n + SL_COLS - 1) / SL_COLS;
  int grid = (groups + 7) / 8;
  const int max_grid = device_sm_count() * 3;
  if (grid > max_grid) grid = max_grid;
  small_linear_kernel<<<grid, 256, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(x), x_ld, static_cast<const __nv_bfloat16*>(w), w_ld,
      static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(y), y_ld, m, n, k, flags);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int timestep_embed_launch(const float* t, void* out, int m, cudaStream_t stream) {
  AFB_REQUIRE(t && out && m >= 1, "timestep_embed: bad arguments");
  timestep_embed_kernel<<<(m * 128 + 127) / 128, 128, 0, stream>>>(
      t, static_cast<__nv_bfloat16*>(out), m);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int sampler_step_launch(const void* head, int64_t head_ld, const float* x_in, float* x_out,
                        void* x_out_bf16, int64_t tokens, int num_gaussians, float sigma_src,
                        float sigma_start, float sigma_end, float eps, cudaStream_t stream) {
  AFB_REQUIRE(head && x_in && x_out, "sampler_step: null pointer");
  AFB_REQUIRE(tokens >= 1, "sampler_step: no tokens");
  if (num_gaussians != 16) {
    set_last_error("sampler_step: only K=16 mixture components are built (got %d)", num_gaussians);
    return AFB_ERR_UNSUPPORTED;
  }
  AFB_REQUIRE(head_ld >= 16 * 64 + 16 * 4 + 15 * 4 && head_ld % 2 == 0, "sampler_step: head_ld=%lld too small",
              (long long)head_ld);
  const unsigned grid = unsigned((tokens + 7) / 8);
  sampler_step_k16_kernel<<<grid, 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(head), head_ld, x_in, x_out,
      static_cast<__nv_bfloat16*>(x_out_bf16), tokens, sigma_src - sigma_start,
      sigma_start - sigma_end, eps);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int cast_f32_bf16_launch(const float* in, void* out, int64_t n, cudaStream_t stream) {
  AFB_REQUIRE(in && out && n >= 1, "cast: bad arguments");
  const long long thr = (n + 3) / 4;
  cast_f32_bf16_kernel<<<unsigned((thr + 255) / 256), 256, 0, stream>>>(
      in, static_cast<__nv_bfloat16*>(out), n);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int rmsnorm_rows_launch(const void* x, void* y, const void* w, int64_t rows, int dim, float eps,
                        cudaStream_t stream) {
  AFB_REQUIRE(x && y && w && rows >= 1, "rmsnorm_rows: bad arguments");
  AFB_REQUIRE(dim % 256 == 0, "rmsnorm_rows: dim=%d must be a multiple of 256", dim);
  rmsnorm_rows_kernel<<<unsigned((rows + 7) / 8), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y),
      static_cast<const __nv_bfloat16*>(w), rows, dim, eps);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int policy_eval_launch(const afb_policy_args* a, cudaStream_t stream) {
  AFB_REQUIRE(a && a->head && a->out && a->sigma_src && a->sigma_start, "policy_eval: null argument");
  AFB_REQUIRE(a->mode >= AFB_POLICY_INTEGRATE && a->mode <= AFB_POLICY_AVERAGE_U, "policy_eval: unknown mode %d", a->mode);
  AFB_REQUIRE(a->mode == AFB_POLICY_VELOCITY || a->sigma_end, "policy_eval: sigma_end missing");
  AFB_REQUIRE(a->mode != AFB_POLICY_INTEGRATE || a->x_in, "policy_eval: x_in missing");
  AFB_REQUIRE(a->batch >= 1 && a->batch <= POLICY_MAX_BATCH, "policy_eval: batch=%d must be in [1, %d]", a->batch,
              POLICY_MAX_BATCH);
  AFB_REQUIRE(a->tokens >= 1, "policy_eval: no tokens");
  if (a->num_gaussians != 16) {
    set_last_error("policy_eval: only K=16 mixture components are built (got %d)", a->num_gaussians);
    return AFB_ERR_UNSUPPORTED;
  }
  AFB_REQUIRE(a->head_ld >= 16 * 64 + 16 * 4 + 15 * 4 && a->head_ld % 2 == 0, "policy_eval: head_ld too small");
  PolicyParams pp{};
  for (int b = 0; b < a->batch; ++b) {
    pp.dt_past[b] = a->sigma_src[b] - a->sigma_start[b];
    pp.dt_step[b] = a->mode == AFB_POLICY_VELOCITY ? 0.f : a->sigma_start[b] - a->sigma_end[b];
    uint32_t m = 0;
    if (a->drop_mask)
      for (int k = 0; k < 16; ++k) m |= (a->drop_mask[b * 16 + k] ? 1u : 0u) << k;
    AFB_REQUIRE(m != 0xFFFFu, "policy_eval: sample %d drops every mixture component", b);
    pp.drop[b] = m;
    pp.small[b] = a->small ? a->small[b] : 0;
  }
  const long long tokens = (long long)a->batch * a->tokens;
  policy_eval_k16_kernel<<<unsigned((tokens + 7) / 8), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(a->head), a->head_ld, a->x_in, a->out,
      static_cast<__nv_bfloat16*>(a->out_bf16), a->tokens, tokens, a->mode, a->eps > 0.f ? a->eps : 1e-4f, pp);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int cfg_combine_launch(const void* both_bf16, float* out, int64_t half, float guidance_scale, cudaStream_t stream) {
  AFB_REQUIRE(both_bf16 && out && half >= 2 && half % 2 == 0, "cfg_combine: bad arguments");
  cfg_combine_kernel<<<unsigned((half / 2 + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(both_bf16), out,
                                                                          half, guidance_scale - 1.0f);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int axpy_rows_launch(const float* x, const void* u, const float* coef, float* out, void* out_bf16, int batch,
                     int64_t per_sample, int u_f32, cudaStream_t stream) {
  AFB_REQUIRE(x && u && coef && out, "axpy_rows: null argument");
  AFB_REQUIRE(batch >= 1 && batch <= POLICY_MAX_BATCH && per_sample >= 2 && per_sample % 2 == 0,
              "axpy_rows: bad shape (batch=%d per_sample=%lld)", batch, (long long)per_sample);
  RowCoef rc{};
  for (int b = 0; b < batch; ++b) rc.c[b] = coef[b];
  const long long total = (long long)batch * per_sample;
  axpy_rows_kernel<<<unsigned((total / 2 + 255) / 256), 256, 0, stream>>>(
      x, u, u_f32, out, static_cast<__nv_bfloat16*>(out_bf16), per_sample, total, rc);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int mse_rows_launch(const float* pred, const void* tgt, float* out, int batch, int64_t per_sample, int tgt_f32,
                    cudaStream_t stream) {
  AFB_REQUIRE(pred && tgt && out, "mse_rows: null argument");
  AFB_REQUIRE(batch >= 1 && per_sample >= 2 && per_sample % 2 == 0, "mse_rows: bad shape");
  AFB_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * batch, stream));
  dim3 grid(64, batch);
  mse_rows_kernel<<<grid, 256, 0, stream>>>(pred, tgt, tgt_f32, out, per_sample);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int policy_backward_launch(const afb_policy_args* a, const void* tgt, float* dhead, int64_t dh_ld, float coef,
                           int accumulate, int tgt_f32, cudaStream_t stream) {
  AFB_REQUIRE(a && a->head && tgt && dhead && a->sigma_src && a->sigma_start && a->sigma_end,
              "policy_backward: null argument");
  AFB_REQUIRE(a->batch >= 1 && a->batch <= POLICY_MAX_BATCH && a->tokens >= 1, "policy_backward: bad batch/tokens");
  if (a->num_gaussians != 16) {
    set_last_error("policy_backward: only K=16 mixture components are built (got %d)", a->num_gaussians);
    return AFB_ERR_UNSUPPORTED;
  }
  AFB_REQUIRE(a->head_ld >= 1148 && dh_ld >= 1148 && dh_ld % 2 == 0, "policy_backward: leading dims too small");
  PolicyParams pp{};
  for (int b = 0; b < a->batch; ++b) {
    pp.dt_past[b] = a->sigma_src[b] - a->sigma_start[b];
    pp.dt_step[b] = a->sigma_start[b] - a->sigma_end[b];
    pp.small[b] = a->small ? a->small[b] : 0;
  }
  const long long tokens = (long long)a->batch * a->tokens;
  policy_avg_u_bwd_k16_kernel<<<unsigned((tokens + 7) / 8), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(a->head), a->head_ld, tgt, tgt_f32, dhead, dh_ld,
      a->tokens, tokens, coef, a->eps > 0.f ? a->eps : 1e-4f, accumulate, pp);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int colsum_f32_launch(const float* x, int64_t ld, float* out, int64_t rows, int n, cudaStream_t stream) {
  AFB_REQUIRE(x && out && rows >= 1 && n >= 1, "colsum: bad arguments");
  const int rpb = 256;
  dim3 grid((n + 255) / 256, unsigned((rows + rpb - 1) / rpb));
  colsum_f32_kernel<<<grid, 256, 0, stream>>>(x, ld, out, rows, n, rpb);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int ln_mod_param_grad_strided_launch(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, float* stats_ws,
                                     float* dscale, float* dshift, int64_t out_bs, int batches, int rows_per_batch, int dim,
                                     float eps, cudaStream_t stream);
int ln_mod_param_grad_launch(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, float* stats_ws, float* dscale,
                             float* dshift, int batches, int rows_per_batch, int dim, float eps, cudaStream_t stream) {
  return ln_mod_param_grad_strided_launch(x, x_bs, dy, dy_bs, stats_ws, dscale, dshift, dim, batches, rows_per_batch, dim, eps,
                                          stream);
}
// dscale / dshift rows are out_bs floats apart (slots of one [batch, mod_total] modulation-gradient buffer)
int ln_mod_param_grad_strided_launch(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, float* stats_ws,
                                     float* dscale, float* dshift, int64_t out_bs, int batches, int rows_per_batch, int dim,
                                     float eps, cudaStream_t stream) {
  AFB_REQUIRE(x && dy && stats_ws && dscale && dshift, "ln_mod_param_grad: null pointer");
  AFB_REQUIRE(batches >= 1 && rows_per_batch >= 1 && dim % 256 == 0, "ln_mod_param_grad: bad shape");
  const long long rows = (long long)batches * rows_per_batch;
  ln_row_stats_kernel<<<unsigned((rows + 7) / 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), x_bs,
                                                                   reinterpret_cast<float2*>(stats_ws), batches,
                                                                   rows_per_batch, dim, eps);
  AFB_CHECK_CUDA(cudaGetLastError());
  const int rpc = 64;
  const int cpb = (rows_per_batch + rpc - 1) / rpc;
  dim3 grid((dim / 2 + 255) / 256, unsigned(batches * cpb));
  ln_mod_param_grad_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), x_bs,
                                                     static_cast<const __nv_bfloat16*>(dy), dy_bs,
                                                     reinterpret_cast<const float2*>(stats_ws), dscale, dshift, out_bs,
                                                     rows_per_batch, dim, cpb, rpc);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return AFB_OK;
}

int rowlinear_param_grad_launch(const float* de, int64_t de_ld, const void* t, int64_t t_ld, float* dw, int64_t dw_ld,
                                float* dbias, int m, int n_out, int k_in, int silu_in, cudaStream_t stream) {
  AFB_REQUIRE(de && t && dw && m >= 1 && n_out >= 1 && k_in >= 1, "rowlinear_param_grad: bad arguments");
  dim3 grid((k_in + 255) / 256, n_out);
  rowlinear_param_grad_kernel<<<grid, 256, 0, stream>>>(de, de_ld, static_cast<const __nv_bfloat16*>(t), t_ld, dw, dw_ld,
                                                        dbias, m, n_out, k_in, silu_in);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int ln_modulate_bwd_launch(const void* x, int64_t x_bs, const void* dy, int64_t dy_bs, void* dh, int64_t dh_bs,
                           const void* scale, int64_t mod_bs, int batches, int rows_per_batch, int dim, float eps,
                           int accumulate, cudaStream_t stream) {
  AFB_REQUIRE(x && dy && dh && scale, "ln_modulate_bwd: null pointer");
  AFB_REQUIRE(batches >= 1 && rows_per_batch >= 1 && dim % 256 == 0, "ln_modulate_bwd: bad shape");
  const long long rows = (long long)batches * rows_per_batch;
  const unsigned grid = unsigned((rows + 7) / 8);
#define AFB_LNB_CASE(NCH)                                                                                       \
  case NCH:                                                                                                     \
    ln_modulate_bwd_kernel<NCH><<<grid, 256, 0, stream>>>(                                                      \
        static_cast<const __nv_bfloat16*>(x), x_bs, static_cast<const __nv_bfloat16*>(dy), dy_bs,               \
        static_cast<__nv_bfloat16*>(dh), dh_bs, static_cast<const __nv_bfloat16*>(scale), mod_bs, batches,      \
        rows_per_batch, eps, accumulate);                                                                       \
    break;
  switch (dim / 256) {
    AFB_LNB_CASE(1)
    AFB_LNB_CASE(2)
    AFB_LNB_CASE(4)
    AFB_LNB_CASE(8)
    AFB_LNB_CASE(12)
    default:
      set_last_error("ln_modulate_bwd: unsupported dim %d", dim);
      return AFB_ERR_UNSUPPORTED;
  }
#undef AFB_LNB_CASE
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int rowscale_launch(const void* x, int64_t x_ld, int64_t x_bs, const void* vec, int64_t vec_bs, void* out, int64_t out_ld,
                    int64_t out_bs, int batches, int rows_per_batch, int cols, cudaStream_t stream) {
  AFB_REQUIRE(x && vec && out && batches >= 1 && rows_per_batch >= 1 && cols % 8 == 0, "rowscale: bad arguments");
  const long long chunks = (long long)batches * rows_per_batch * (cols / 8);
  AFB_REQUIRE(chunks < (1ll << 32) - 256, "rowscale: too many elements for one launch");
  rowscale_kernel<<<unsigned((chunks + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), x_ld, x_bs, static_cast<const __nv_bfloat16*>(vec), vec_bs,
      static_cast<__nv_bfloat16*>(out), out_ld, out_bs, rows_per_batch, cols, unsigned(chunks));
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int gelu_bwd_launch(void* dm, int64_t dm_ld, const void* pre, int64_t pre_ld, int64_t rows, int cols, cudaStream_t stream) {
  AFB_REQUIRE(dm && pre && rows >= 1 && cols % 8 == 0, "gelu_bwd: bad arguments");
  const long long chunks = rows * (cols / 8);
  AFB_REQUIRE(chunks < (1ll << 32) - 256, "gelu_bwd: too many elements for one launch");
  gelu_bwd_kernel<<<unsigned((chunks + 255) / 256), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(dm), dm_ld,
                                                                     static_cast<const __nv_bfloat16*>(pre), pre_ld, cols,
                                                                     unsigned(chunks));
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int gate_bwd_launch(const void* dh, int64_t dh_bs, const void* u, int64_t u_bs, const void* gate, int64_t gate_bs, void* du,
                    int64_t du_bs, float* dgate, int64_t dgate_bs, int batches, int rows_per_batch, int cols,
                    cudaStream_t stream) {
  AFB_REQUIRE(dh && u && gate && du && dgate, "gate_bwd: null pointer");
  AFB_REQUIRE(batches >= 1 && rows_per_batch >= 1 && cols % 8 == 0, "gate_bwd: bad shape");
  const int rpc = 32;
  const int cpb = (rows_per_batch + rpc - 1) / rpc;
  dim3 grid((cols / 8 + 127) / 128, unsigned(batches * cpb));
  gate_bwd_kernel<<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(dh), dh_bs, static_cast<const __nv_bfloat16*>(u),
                                            u_bs, static_cast<const __nv_bfloat16*>(gate), gate_bs,
                                            static_cast<__nv_bfloat16*>(du), du_bs, dgate, dgate_bs, rows_per_batch, cols, cpb,
                                            rpc);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int gate_res_launch(const void* res, int64_t res_bs, const void* u, int64_t u_bs, const void* gate, int64_t gate_bs, void* out,
                    int64_t out_bs, int batches, int rows_per_batch, int cols, cudaStream_t stream) {
  AFB_REQUIRE(res && u && gate && out && batches >= 1 && rows_per_batch >= 1 && cols % 8 == 0, "gate_res: bad arguments");
  const long long chunks = (long long)batches * rows_per_batch * (cols / 8);
  AFB_REQUIRE(chunks < (1ll << 32) - 256, "gate_res: too many elements for one launch");
  gate_res_kernel<<<unsigned((chunks + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(res), res_bs, static_cast<const __nv_bfloat16*>(u), u_bs,
      static_cast<const __nv_bfloat16*>(gate), gate_bs, static_cast<__nv_bfloat16*>(out), out_bs, rows_per_batch, cols, unsigned(chunks));
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int rowlinear_dx_launch(const float* de, int64_t de_ld, const void* w, int64_t w_ld, float* out, int64_t out_ld, int m, int J,
                        int N, cudaStream_t stream) {
  AFB_REQUIRE(de && w && out && m >= 1 && J >= 1 && N >= 8 && N % 8 == 0 && w_ld % 8 == 0, "rowlinear_dx: bad arguments");
  dim3 grid((N / 8 + 127) / 128, unsigned((J + RDX_JC - 1) / RDX_JC));
  for (int r0 = 0; r0 < m; r0 += 8) {
    const int mm = m - r0 < 8 ? m - r0 : 8;
    rowlinear_dx_kernel<<<grid, 128, 0, stream>>>(de + (int64_t)r0 * de_ld, de_ld, static_cast<const __nv_bfloat16*>(w), w_ld,
                                                  out + (int64_t)r0 * out_ld, out_ld, mm, J, N);
    AFB_CHECK_CUDA(cudaGetLastError());
    count_launch(1);
  }
  return AFB_OK;
}

int silu_bwd_launch(float* d, int64_t d_ld, const void* x, int64_t x_ld, int rows, int cols, cudaStream_t stream) {
  AFB_REQUIRE(d && x && rows >= 1 && cols >= 1, "silu_bwd: bad arguments");
  dim3 grid((cols + 255) / 256, rows);
  silu_bwd_kernel<<<grid, 256, 0, stream>>>(d, d_ld, static_cast<const __nv_bfloat16*>(x), x_ld, rows, cols);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

// dst[r, c] += mask * src[r, c] / (1 - p)   (fp32, the tiny timestep-embedder path); logical index = r * cols + c
__global__ void __launch_bounds__(256)
dropout_f32_add_kernel(const float* __restrict__ src, long long src_ld, float* __restrict__ dst, long long dst_ld, int rows,
                       int cols, uint32_t key, uint32_t thresh, float inv_keep) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= cols || r >= rows) return;
  if (dropout_keep(key, (unsigned long long)r * cols + c, thresh))
    dst[(long long)r * dst_ld + c] += src[(long long)r * src_ld + c] * inv_keep;
}

__global__ void scale_bf16_kernel(__nv_bfloat16* __restrict__ x, long long n, float s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = __float2bfloat16(__bfloat162float(x[i]) * s);
}

uint32_t dropout_layer_key(uint64_t seed, uint32_t layer_id) {
  auto mix = [](uint32_t h) {
    h ^= h >> 16;
    h *= 0x7feb352dU;
    h ^= h >> 15;
    h *= 0x846ca68bU;
    h ^= h >> 16;
    return h;
  };
  return mix(uint32_t(seed) ^ mix(uint32_t(seed >> 32) + layer_id * 0x632BE5ABU + 1U));
}

// out (=|+=) dropout_mask (.) act(x) / (1 - p) on a [batches, rows, cols] view; the mask is indexed by the LOGICAL position
// (batch*rows + row) * logical_cols + col0 + col, so column slices of one logical tensor can be processed separately.
int dropout_rows_launch(const void* x, int64_t x_ld, int64_t x_bs, void* out, int64_t out_ld, int64_t out_bs, int batches,
                        int rows_per_batch, int cols, int logical_cols, int col0, uint64_t seed, uint32_t layer_id, float p,
                        int silu_in, int accumulate, cudaStream_t stream) {
  AFB_REQUIRE(x && out && batches >= 1 && rows_per_batch >= 1 && cols % 8 == 0 && x_ld % 8 == 0 && out_ld % 8 == 0,
              "dropout_rows: bad arguments");
  AFB_REQUIRE(p >= 0.f && p < 1.f && logical_cols >= col0 + cols && logical_cols % 2 == 0 && col0 % 2 == 0,
              "dropout_rows: bad p / logical layout");
  DropParams dp{};
  dp.rows_per_batch = rows_per_batch;
  dp.cols = cols;
  dp.logical_cols = logical_cols;
  dp.col0 = col0;
  dp.key = dropout_layer_key(seed, layer_id);
  dp.thresh = dropout_thresh16(p);
  dp.inv_keep = 1.0f / (1.0f - p);
  dp.silu_in = silu_in;
  dp.accumulate = accumulate;
  const long long total_rows = (long long)batches * rows_per_batch;
  AFB_REQUIRE(total_rows < (1ll << 31), "dropout_rows: too many rows");
  const long long groups = (total_rows + 3) / 4;   // DROP_ROWS rows per block
  const dim3 grid(unsigned((cols / 8 + 255) / 256), unsigned(groups < 65535 ? groups : 65535));
  dropout_rows_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), x_ld, x_bs,
                                                static_cast<__nv_bfloat16*>(out), out_ld, out_bs, int(total_rows), dp);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int scale_bf16_launch(void* x, int64_t n, float s, cudaStream_t stream) {
  AFB_REQUIRE(x && n >= 1, "scale_bf16: bad arguments");
  scale_bf16_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(x), n, s);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int dropout_f32_add_launch(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int rows, int cols, uint64_t seed,
                           uint32_t layer_id, float p, cudaStream_t stream) {
  AFB_REQUIRE(src && dst && rows >= 1 && cols >= 1 && p >= 0.f && p < 1.f, "dropout_f32_add: bad arguments");
  dim3 grid((cols + 255) / 256, rows);
  dropout_f32_add_kernel<<<grid, 256, 0, stream>>>(src, src_ld, dst, dst_ld, rows, cols, dropout_layer_key(seed, layer_id),
                                                   dropout_thresh16(p), 1.0f / (1.0f - p));
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int gelu_fwd_launch(const void* pre, int64_t pre_ld, void* out, int64_t out_ld, int64_t rows, int cols, cudaStream_t stream) {
  AFB_REQUIRE(pre && out && rows >= 1 && cols % 8 == 0, "gelu_fwd: bad arguments");
  const long long chunks = rows * (cols / 8);
  AFB_REQUIRE(chunks < (1ll << 32) - 256, "gelu_fwd: too many elements for one launch");
  gelu_fwd_kernel<<<unsigned((chunks + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pre), pre_ld,
                                                                     static_cast<__nv_bfloat16*>(out), out_ld, cols, unsigned(chunks));
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int rmsnorm_rope_bwd_launch(void* dqkv, const void* raw, int64_t ld, int64_t bs, int q_off, int k_off, int batches, int seq,
                            int heads, int txt_rows, const void* wq_txt, const void* wk_txt, const void* wq_img,
                            const void* wk_img, const float* cos_tab, const float* sin_tab, float eps, cudaStream_t stream) {
  AFB_REQUIRE(dqkv && raw && wq_img && wk_img && cos_tab && sin_tab, "rmsnorm_rope_bwd: null pointer");
  AFB_REQUIRE(txt_rows == 0 || (wq_txt && wk_txt), "rmsnorm_rope_bwd: text norm weights missing");
  AFB_REQUIRE(ld % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0, "rmsnorm_rope_bwd: misaligned layout");
  const long long rows = (long long)batches * seq;
  rmsnorm_rope_bwd_kernel<<<unsigned((rows + 7) / 8), 256, 0, stream>>>(
      static_cast<__nv_bfloat16*>(dqkv), static_cast<const __nv_bfloat16*>(raw), ld, bs, q_off, k_off, batches, seq, heads,
      txt_rows, static_cast<const __nv_bfloat16*>(wq_txt ? wq_txt : wq_img),
      static_cast<const __nv_bfloat16*>(wk_txt ? wk_txt : wk_img), static_cast<const __nv_bfloat16*>(wq_img),
      static_cast<const __nv_bfloat16*>(wk_img), cos_tab, sin_tab, eps);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int attn_delta_launch(const void* o, int64_t o_ld, int64_t o_bs, const void* d_o, int64_t do_ld, int64_t do_bs, float* delta,
                      int batch, int seq, int heads, cudaStream_t stream) {
  AFB_REQUIRE(o && d_o && delta && batch >= 1 && seq >= 1 && heads >= 1, "attn_delta: bad arguments");
  const long long warps = (long long)batch * seq * heads;
  attn_delta_kernel<<<unsigned((warps + 7) / 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(o), o_ld, o_bs,
                                                                  static_cast<const __nv_bfloat16*>(d_o), do_ld, do_bs, delta,
                                                                  batch, seq, heads);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int fill_f32_launch(float* p, float v, int n, cudaStream_t stream) {
  fill_f32_kernel<<<(n + 127) / 128, 128, 0, stream>>>(p, v, n);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

}  // namespace afb
