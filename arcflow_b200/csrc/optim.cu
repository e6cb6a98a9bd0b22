// arcflow_b200 — the step glue after the backward as three HBM-bound passes over ONE flat fp32 parameter arena
// (all trainable adapter tensors live back to back: params | grads | exp_avg | exp_avg_sq | ema):
//   grad-norm^2 -> [clip, skip on NaN/Inf] + AdamW + Karras EMA lerp + bf16 shadow for the engine, fused.
// Reference: BaseModel.step_optimizer (lakonlab/models/base.py:76-103: clip_grad_norm_(50) from iteration 100, skip the
// step on a non-finite norm), optimizer AdamW8bit lr 1e-4 betas (0.9, 0.95) wd 0 with `proj_out_loggamma` lr x 0.1
// (configs/flux/_ddp_train.py:13-26), ExponentialMovingAverageHookMod (lakonlab/runner/hooks/ema_hook.py:86-121:
// ema = m * ema + (1 - m) * net, m = min((1 - 1/t)^(gamma+1), 1); straight copy before start_iter).
// Two state modes: fp32 moments with torch.optim.AdamW arithmetic (adamw_ema_kernel), and the block-wise 8-bit moments of
// bitsandbytes' AdamW8bit — the optimizer the reference's configs name — restated from the published algorithm
// (adamw8bit_ema_kernel; bitsandbytes is neither vendored nor version-pinned by the reference).
#include <math.h>

#include "common.cuh"

namespace afb {
namespace {

// Deterministic: per-block partials in a fixed slot, the last block to finish adds them in a fixed order — every DDP rank
// gets the bit-identical norm (and clip coefficient) from the all-reduced gradients, so the replicas cannot drift apart.
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out, float* __restrict__ partials,
             unsigned int* __restrict__ ticket) {
  float acc = 0.f;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long e = i; e < n; ++e) acc += g[e] * g[e];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    *out = s;
    *ticket = 0;
  }
}

struct AdamParams {
  float lr, beta1, beta2, eps, weight_decay, bias1, bias2;  // bias_k = 1 - beta_k^step
  float max_norm;      // <= 0: no clipping
  float skip_norm;     // > 0: a norm above it skips the step like a non-finite one (grad_clip_skip_ratio * grad_clip)
  float ema_momentum;  // < 0: EMA not touched; ema = m * ema + (1 - m) * p
  int ema_copy;        // before start_iter: ema = p
  long long lo_begin, lo_end;  // element range that uses lr * lo_mult (proj_out_loggamma)
  float lo_mult;
};

__global__ void __launch_bounds__(256)
adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 float* __restrict__ ema, __nv_bfloat16* __restrict__ shadow, long long n,
                 const float* __restrict__ gnorm_sq, int* __restrict__ skipped, const AdamParams a) {
  float clip = 1.0f;
  if (a.max_norm > 0.f) {
    const float norm = sqrtf(*gnorm_sq);
    if (!isfinite(norm) || (a.skip_norm > 0.f && norm > a.skip_norm)) {  // reference (base.py:91-95): zero_grad + skip the optimizer step; EMA still runs afterwards
      if (blockIdx.x == 0 && threadIdx.x == 0) *skipped = 1;
      clip = -1.0f;
    } else {
      clip = fminf(1.0f, a.max_norm / (norm + 1e-6f));  // torch.nn.utils.clip_grad_norm_
    }
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w = p[i];
    if (clip >= 0.f) {
      const float gi = g[i] * clip;
      const float lr = (i >= a.lo_begin && i < a.lo_end) ? a.lr * a.lo_mult : a.lr;
      w *= 1.0f - lr * a.weight_decay;
      const float mi = a.beta1 * m[i] + (1.0f - a.beta1) * gi;
      const float vi = a.beta2 * v[i] + (1.0f - a.beta2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      w -= lr * (mi / a.bias1) / (sqrtf(vi / a.bias2) + a.eps);
      p[i] = w;
    }
    if (ema) {
      if (a.ema_copy) ema[i] = w;
      else if (a.ema_momentum >= 0.f) ema[i] = w + (ema[i] - w) * a.ema_momentum;
    }
    if (shadow) shadow[i] = __float2bfloat16_rn(w);
  }
}

// ---- block-wise 8-bit moment state (bitsandbytes AdamW8bit; see include/arcflow_b200.h) -------------------------------
// One warp per 256-element quantisation block, 8 consecutive elements per lane (two float4 / one uint2 per array), so the
// block absmax is a warp-shuffle reduction and nothing but the two code books lives in shared memory.
__device__ __forceinline__ int nearest_code(const float* __restrict__ q, float x) {
  int lo = 0, hi = 255;  // q[lo] < x <= q[hi] once the search has closed in
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int mid = (lo + hi) >> 1;
    if (x > q[mid]) lo = mid;
    else hi = mid;
  }
  return (x - q[lo] > q[hi] - x) ? hi : lo;
}

__global__ void __launch_bounds__(256)
adamw8bit_ema_kernel(float* __restrict__ p, const float* __restrict__ g, uint8_t* __restrict__ c1, uint8_t* __restrict__ c2,
                     float* __restrict__ absmax1, float* __restrict__ absmax2, const float* __restrict__ qmap1,
                     const float* __restrict__ qmap2, float* __restrict__ ema, __nv_bfloat16* __restrict__ shadow,
                     long long nblocks, const float* __restrict__ gnorm_sq, int* __restrict__ skipped, const AdamParams a) {
  __shared__ float q1[256], q2[256];
  q1[threadIdx.x] = qmap1[threadIdx.x];
  q2[threadIdx.x] = qmap2[threadIdx.x];
  __syncthreads();
  float clip = 1.0f;
  if (a.max_norm > 0.f) {
    const float norm = sqrtf(*gnorm_sq);
    if (!isfinite(norm) || (a.skip_norm > 0.f && norm > a.skip_norm)) {
      if (blockIdx.x == 0 && threadIdx.x == 0) *skipped = 1;
      clip = -1.0f;
    } else {
      clip = fminf(1.0f, a.max_norm / (norm + 1e-6f));
    }
  }
  const int lane = threadIdx.x & 31;
  const float corr2 = sqrtf(a.bias2);                 // bitsandbytes: correction2 = sqrt(1 - beta2^t)
  const float eps2 = corr2 * a.eps;
  for (long long blk = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); blk < nblocks; blk += (long long)gridDim.x * 8) {
    const long long i0 = blk * 256 + lane * 8;
    float w[8];
    *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(p + i0);
    *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(p + i0 + 4);
    if (clip >= 0.f) {
      float gi[8], s1[8], s2[8];
      *reinterpret_cast<float4*>(gi) = *reinterpret_cast<const float4*>(g + i0);
      *reinterpret_cast<float4*>(gi + 4) = *reinterpret_cast<const float4*>(g + i0 + 4);
      const uint2 k1 = *reinterpret_cast<const uint2*>(c1 + i0);
      const uint2 k2 = *reinterpret_cast<const uint2*>(c2 + i0);
      const float am1 = absmax1[blk], am2 = absmax2[blk];
      const float lr = (i0 >= a.lo_begin && i0 < a.lo_end) ? a.lr * a.lo_mult : a.lr;   // slots never straddle a block
      const float step_size = -lr * corr2 / a.bias1;
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t b1 = ((j < 4 ? k1.x : k1.y) >> ((j & 3) * 8)) & 0xffu;
        const uint32_t b2 = ((j < 4 ? k2.x : k2.y) >> ((j & 3) * 8)) & 0xffu;
        const float gv = gi[j] * clip;
        if (isfinite(gv)) {
          s2[j] = (q2[b2] * am2) * a.beta2 + (1.0f - a.beta2) * gv * gv;
          s1[j] = (q1[b1] * am1) * a.beta1 + (1.0f - a.beta1) * gv;
        } else {   // a non-finite gradient element resets its moments and leaves the parameter alone
          s1[j] = 0.f;
          s2[j] = 0.f;
        }
        gi[j] = gv;
        m1 = fmaxf(m1, fabsf(s1[j]));
        m2 = fmaxf(m2, fabsf(s2[j]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      }
      if (lane == 0) {
        absmax1[blk] = m1;
        absmax2[blk] = m2;
      }
      const float inv1 = m1 > 0.f ? 1.0f / m1 : 0.f, inv2 = m2 > 0.f ? 1.0f / m2 : 0.f;
      uint32_t o1[2] = {0u, 0u}, o2[2] = {0u, 0u};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (isfinite(gi[j])) {
          w[j] += step_size * (s1[j] / (sqrtf(s2[j]) + eps2));
          if (a.weight_decay > 0.f) w[j] *= 1.0f - lr * a.weight_decay;
        }
        int n1 = nearest_code(q1, s1[j] * inv1);
        if (signbit(q1[n1]) != signbit(s1[j])) n1 += s1[j] > 0.f ? 1 : -1;   // exp_avg keeps its sign through the code book
        n1 = min(max(n1, 0), 255);
        const int n2 = nearest_code(q2, s2[j] * inv2);
        o1[j >> 2] |= uint32_t(n1) << ((j & 3) * 8);
        o2[j >> 2] |= uint32_t(n2) << ((j & 3) * 8);
      }
      *reinterpret_cast<uint2*>(c1 + i0) = make_uint2(o1[0], o1[1]);
      *reinterpret_cast<uint2*>(c2 + i0) = make_uint2(o2[0], o2[1]);
      *reinterpret_cast<float4*>(p + i0) = *reinterpret_cast<const float4*>(w);
      *reinterpret_cast<float4*>(p + i0 + 4) = *reinterpret_cast<const float4*>(w + 4);
    }
    if (ema) {
      float e[8];
      if (a.ema_copy) {
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = w[j];
      } else {
        *reinterpret_cast<float4*>(e) = *reinterpret_cast<const float4*>(ema + i0);
        *reinterpret_cast<float4*>(e + 4) = *reinterpret_cast<const float4*>(ema + i0 + 4);
        if (a.ema_momentum >= 0.f) {
#pragma unroll
          for (int j = 0; j < 8; ++j) e[j] = w[j] + (e[j] - w[j]) * a.ema_momentum;
        }
      }
      if (a.ema_copy || a.ema_momentum >= 0.f) {
        *reinterpret_cast<float4*>(ema + i0) = *reinterpret_cast<const float4*>(e);
        *reinterpret_cast<float4*>(ema + i0 + 4) = *reinterpret_cast<const float4*>(e + 4);
      }
    }
    if (shadow) {
      uint4 sh;
      sh.x = pack_bf16x2(w[0], w[1]);
      sh.y = pack_bf16x2(w[2], w[3]);
      sh.z = pack_bf16x2(w[4], w[5]);
      sh.w = pack_bf16x2(w[6], w[7]);
      *reinterpret_cast<uint4*>(shadow + i0) = sh;
    }
  }
}

}  // namespace

int grad_norm_scratch_floats() { return device_sm_count() * 8 + 1; }

// Caller-owned scratch ([grad_norm_scratch_floats()] floats, zeroed once by the caller): one scratch per optimizer
// instance, so launches of different instances on different streams cannot race on the partials / ticket, and nothing is
// allocated here (CUDA-graph capturable).
int grad_norm_sq_ws_launch(const float* g, int64_t n, float* out, float* scratch, int64_t scratch_floats, cudaStream_t stream) {
  AFB_REQUIRE(g && out && scratch && n >= 1, "grad_norm_sq: bad arguments");
  const int cap = device_sm_count() * 8;
  AFB_REQUIRE(scratch_floats >= int64_t(cap) + 1, "grad_norm_sq: scratch holds %lld floats, needs %d", (long long)scratch_floats,
              cap + 1);
  int blocks = int((n / 4 + 255) / 256);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<blocks, 256, 0, stream>>>(g, n, out, scratch, reinterpret_cast<unsigned int*>(scratch + cap));
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

// Convenience form with one library-owned scratch per device: callers must not run it concurrently on two streams of the
// same device (use the _ws form for that); the first call per device allocates and is not graph-capturable.
int grad_norm_sq_launch(const float* g, int64_t n, float* out, cudaStream_t stream) {
  AFB_REQUIRE(g && out && n >= 1, "grad_norm_sq: bad arguments");
  int blocks = int((n / 4 + 255) / 256);
  const int cap = device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  int dev = 0;
  AFB_CHECK_CUDA(cudaGetDevice(&dev));
  static float* scratch[64] = {};  // per device: [cap] partials + 1 ticket, allocated once, never freed
  AFB_REQUIRE(dev >= 0 && dev < 64, "grad_norm_sq: device ordinal out of range");
  if (!scratch[dev]) {
    AFB_CHECK_CUDA(cudaMalloc(&scratch[dev], (size_t(cap) + 1) * sizeof(float)));
    AFB_CHECK_CUDA(cudaMemsetAsync(scratch[dev], 0, (size_t(cap) + 1) * sizeof(float), stream));
  }
  sumsq_kernel<<<blocks, 256, 0, stream>>>(g, n, out, scratch[dev], reinterpret_cast<unsigned int*>(scratch[dev] + cap));
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int adamw_ema_launch(const afb_adamw_args* a, cudaStream_t stream) {
  AFB_REQUIRE(a && a->params && a->grads && a->exp_avg && a->exp_avg_sq && a->n >= 1, "adamw: bad arguments");
  AFB_REQUIRE(a->step >= 1, "adamw: step counts from 1");
  AFB_REQUIRE(a->max_norm <= 0.f || (a->grad_norm_sq && a->skipped), "adamw: clipping needs grad_norm_sq and skipped");
  AdamParams k{};
  k.lr = a->lr;
  k.beta1 = a->beta1;
  k.beta2 = a->beta2;
  k.eps = a->eps;
  k.weight_decay = a->weight_decay;
  k.bias1 = 1.0f - powf(a->beta1, float(a->step));
  k.bias2 = 1.0f - powf(a->beta2, float(a->step));
  k.max_norm = a->max_norm;
  k.skip_norm = a->skip_norm;
  k.ema_momentum = a->ema_momentum;
  k.ema_copy = a->ema_copy;
  k.lo_begin = a->lr_mult_begin;
  k.lo_end = a->lr_mult_end;
  k.lo_mult = a->lr_mult;
  if (a->skipped) AFB_CHECK_CUDA(cudaMemsetAsync(a->skipped, 0, sizeof(int), stream));
  int blocks = int((a->n + 255) / 256);
  const int cap = device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adamw_ema_kernel<<<blocks, 256, 0, stream>>>(a->params, a->grads, a->exp_avg, a->exp_avg_sq, a->ema,
                                               static_cast<__nv_bfloat16*>(a->bf16_shadow), a->n, a->grad_norm_sq,
                                               a->skipped, k);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

static AdamParams adam_params(const afb_adamw_args* a) {
  AdamParams k{};
  k.lr = a->lr;
  k.beta1 = a->beta1;
  k.beta2 = a->beta2;
  k.eps = a->eps;
  k.weight_decay = a->weight_decay;
  k.bias1 = float(1.0 - pow(double(a->beta1), double(a->step)));
  k.bias2 = float(1.0 - pow(double(a->beta2), double(a->step)));
  k.max_norm = a->max_norm;
  k.skip_norm = a->skip_norm;
  k.ema_momentum = a->ema_momentum;
  k.ema_copy = a->ema_copy;
  k.lo_begin = a->lr_mult_begin;
  k.lo_end = a->lr_mult_end;
  k.lo_mult = a->lr_mult;
  return k;
}

int adamw8bit_ema_launch(const afb_adamw8bit_args* x, cudaStream_t stream) {
  AFB_REQUIRE(x != nullptr, "adamw8bit: null arguments");
  const afb_adamw_args* a = &x->base;
  AFB_REQUIRE(a->params && a->grads && a->n >= 1, "adamw8bit: bad arguments");
  AFB_REQUIRE(x->state1 && x->state2 && x->absmax1 && x->absmax2 && x->qmap1 && x->qmap2, "adamw8bit: null state pointer");
  AFB_REQUIRE(x->blocksize == 256, "adamw8bit: block size %d (only 256, the bitsandbytes >= 0.44 value, is built)", x->blocksize);
  AFB_REQUIRE(a->n % 256 == 0, "adamw8bit: n = %lld is not a multiple of the block size", (long long)a->n);
  AFB_REQUIRE(a->lr_mult_begin % 256 == 0 && a->lr_mult_end % 256 == 0, "adamw8bit: lr-multiplier range must be block aligned");
  AFB_REQUIRE(a->step >= 1, "adamw8bit: step counts from 1");
  AFB_REQUIRE(a->max_norm <= 0.f || (a->grad_norm_sq && a->skipped), "adamw8bit: clipping needs grad_norm_sq and skipped");
  const AdamParams k = adam_params(a);
  if (a->skipped) AFB_CHECK_CUDA(cudaMemsetAsync(a->skipped, 0, sizeof(int), stream));
  const long long nblocks = a->n / 256;
  long long blocks = (nblocks + 7) / 8;
  const int cap = device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adamw8bit_ema_kernel<<<int(blocks), 256, 0, stream>>>(a->params, a->grads, x->state1, x->state2, x->absmax1, x->absmax2,
                                                        x->qmap1, x->qmap2, a->ema, static_cast<__nv_bfloat16*>(a->bf16_shadow),
                                                        nblocks, a->grad_norm_sq, a->skipped, k);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

}  // namespace afb
