// arcflow_b200 — streaming kernels of the FLUX VAE decoder, NHWC bf16 activations.
//
// The reference decodes the final latents with diffusers' AutoencoderKL (`self.vae.decode(latents / scaling_factor +
// shift_factor)`, lakonlab/pipelines/arcflux_pipeline.py:531-534; lakonlab/models/architecture/diffusers/pretrained.py:69-76),
// i.e. the black-forest-labs autoencoder: 3x3 convolutions, GroupNorm(32, eps 1e-6) + swish, nearest 2x upsampling, one
// single-head attention block at the lowest resolution. The convolutions and the attention products run on the tcgen05
// GEMM (csrc/gemm.cu: implicit-GEMM mode, 9 shifted TMA boxes per K sweep); this file holds what is HBM-bound:
//   vae_pre        latents fp32 NCHW [N,16,h,w] -> z / scale + shift -> bf16 NHWC [N,h,w,64] (channels zero-padded to 64)
//   groupnorm      two passes: per-(sample, row-slab) partial sums (deterministic, no atomics) -> normalise * gamma + beta
//                  (+ swish) in bf16; mean / variance over (H*W, C/32) in fp32, combined in double
//   upsample2x     nearest neighbour, [N,H,W,C] -> [N,2H,2W,C]
//   softmax_rows   in-place row softmax of the bf16 score matrix of the attention block (scale folded into the GEMM alpha)
//   vae_post       bf16 NHWC [N,H,W,8] (3 used) -> fp32 NCHW [N,3,H,W]
#include "common.cuh"

namespace afb {
namespace {

__device__ __forceinline__ void unpack8v(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(w[i]);
    f[2 * i + 1] = bf16_hi(w[i]);
  }
}
__device__ __forceinline__ uint4 pack8v(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vae_pre_kernel(const float* __restrict__ z, __nv_bfloat16* __restrict__ out, int n, int c_in, int hw, int c_pad, float inv_scale,
               float shift) {
  // one thread per (sample, pixel, 8-channel group) of the padded output
  const int groups = c_pad / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * hw * groups;
  if (idx >= total) return;
  const int g = int(idx % groups);
  const long long pix = idx / groups;
  const int b = int(pix / hw);
  const int p = int(pix - (long long)b * hw);
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = g * 8 + i;
    f[i] = c < c_in ? z[((long long)b * c_in + c) * hw + p] * inv_scale + shift : 0.f;
  }
  *reinterpret_cast<uint4*>(out + pix * c_pad + g * 8) = pack8v(f);
}

__global__ void __launch_bounds__(256)
vae_post_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld, float* __restrict__ out, int n, int c_out, long long hw) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * hw) return;
  const int b = int(idx / hw);
  const long long p = idx - (long long)b * hw;
  float f[8];
  unpack8v(*reinterpret_cast<const uint4*>(x + idx * x_ld), f);
  for (int c = 0; c < c_out; ++c) out[((long long)b * c_out + c) * hw + p] = f[c];
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm pass 1: grid (slabs, N). A block walks rows [slab * rows_per_slab, ...) of one sample; thread t owns the
// 8-channel vector (t % vec_per_row) and every (256 / vec_per_row)-th row. Per-thread sums -> shared per-group sums ->
// partial[n][slab][group][2]. Deterministic: fixed ownership, fixed reduction order, no atomics across blocks.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GN_GROUPS = 32;

__global__ void __launch_bounds__(256)
groupnorm_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ partial, long long hw, int c, int rows_per_slab) {
  const int vec_per_row = c / 8;
  const int lanes_rows = 256 / vec_per_row;  // rows handled concurrently by the block (c <= 2048)
  const int v = threadIdx.x % vec_per_row;
  const int rl = threadIdx.x / vec_per_row;
  const int slab = blockIdx.x, b = blockIdx.y, slabs = gridDim.x;
  const long long r0 = (long long)slab * rows_per_slab;
  long long r1 = r0 + rows_per_slab;
  if (r1 > hw) r1 = hw;
  const __nv_bfloat16* xb = x + (long long)b * hw * c;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  for (long long r = r0 + rl; r < r1; r += lanes_rows) {
    float f[8];
    unpack8v(*reinterpret_cast<const uint4*>(xb + r * c + v * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i] += f[i];
      q[i] += f[i] * f[i];
    }
  }
  // per-thread sums -> shared [row lane][channel] -> per-group totals in a fixed order (deterministic, no atomics)
  __shared__ float sh[2][2048];  // lanes_rows * c == 256 * 8
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sh[0][rl * c + v * 8 + i] = s[i];
    sh[1][rl * c + v * 8 + i] = q[i];
  }
  __syncthreads();
  if (threadIdx.x < 2 * GN_GROUPS) {
    const int which = threadIdx.x / GN_GROUPS, g = threadIdx.x % GN_GROUPS;
    const int cpg = c / GN_GROUPS;
    float t = 0.f;
    for (int r = 0; r < lanes_rows; ++r)
      for (int i = 0; i < cpg; ++i) t += sh[which][r * c + g * cpg + i];
    partial[(((long long)b * slabs + slab) * GN_GROUPS + g) * 2 + which] = t;
  }
}

// pass 2: every block first folds the slab partials of its sample into mean / rstd per group (double), then streams rows.
__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ partial,
                       const float* __restrict__ gamma, const float* __restrict__ beta, long long hw, int c, int slabs,
                       int rows_per_block, float eps, int silu_on) {
  __shared__ float mean_s[GN_GROUPS], rstd_s[GN_GROUPS];
  const int b = blockIdx.y;
  const int cpg = c / GN_GROUPS;
  if (threadIdx.x < GN_GROUPS) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < slabs; ++i) {
      s += double(partial[(((long long)b * slabs + i) * GN_GROUPS + threadIdx.x) * 2 + 0]);
      q += double(partial[(((long long)b * slabs + i) * GN_GROUPS + threadIdx.x) * 2 + 1]);
    }
    const double cnt = double(hw) * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_s[threadIdx.x] = float(mean);
    rstd_s[threadIdx.x] = float(1.0 / sqrt(var + double(eps)));
  }
  __syncthreads();
  const int vec_per_row = c / 8;
  const int lanes_rows = 256 / vec_per_row;
  const int v = threadIdx.x % vec_per_row;
  const int rl = threadIdx.x / vec_per_row;
  if (rl >= lanes_rows) return;
  float ga[8], be[8], mu[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = v * 8 + i;
    ga[i] = gamma[ch];
    be[i] = beta[ch];
    mu[i] = mean_s[ch / cpg];
    rs[i] = rstd_s[ch / cpg];
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > hw) r1 = hw;
  const __nv_bfloat16* xb = x + (long long)b * hw * c;
  __nv_bfloat16* yb = y + (long long)b * hw * c;
  for (long long r = r0 + rl; r < r1; r += lanes_rows) {
    float f[8];
    unpack8v(*reinterpret_cast<const uint4*>(xb + r * c + v * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = (f[i] - mu[i]) * rs[i] * ga[i] + be[i];
      if (silu_on) t = t / (1.0f + __expf(-t));
      f[i] = t;
    }
    *reinterpret_cast<uint4*>(yb + r * c + v * 8) = pack8v(f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h, int w, int c) {
  // one thread per 16-byte vector of the OUTPUT
  const int vec = c / 8;
  const long long total = (long long)n * (2 * h) * (2 * w) * vec;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int v = int(idx % vec);
    long long p = idx / vec;
    const int ox = int(p % (2 * w));
    p /= (2 * w);
    const int oy = int(p % (2 * h));
    const int b = int(p / (2 * h));
    const uint4 u = *reinterpret_cast<const uint4*>(x + (((long long)b * h + (oy >> 1)) * w + (ox >> 1)) * c + v * 8);
    *reinterpret_cast<uint4*>(y + idx * 8) = u;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// in-place row softmax, bf16, one block per row; the row is cached in shared memory as fp32 (cols <= 16384 -> 64 KiB)
__global__ void __launch_bounds__(256)
softmax_rows_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, int cols) {
  extern __shared__ float row[];
  __nv_bfloat16* xr = x + (long long)blockIdx.x * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x * 8; i < cols; i += 256 * 8) {
    float f[8];
    unpack8v(*reinterpret_cast<const uint4*>(xr + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      row[i + k] = f[k];
      mx = fmaxf(mx, f[k]);
    }
  }
  __shared__ float red[8];
  __shared__ float bcast;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
    bcast = t;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int i = threadIdx.x * 8; i < cols; i += 256 * 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float e = __expf(row[i + k] - mx);
      row[i + k] = e;
      sum += e;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = bcast;
  for (int i = threadIdx.x * 8; i < cols; i += 256 * 8) {
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = row[i + k] * inv;
    *reinterpret_cast<uint4*>(xr + i) = pack8v(f);
  }
}

}  // namespace

int vae_pre_launch(const float* z, void* out, int n, int c_in, int h, int w, int c_pad, float scale, float shift,
                   cudaStream_t stream) {
  AFB_REQUIRE(z && out && n >= 1 && h >= 1 && w >= 1, "vae_pre: bad arguments");
  AFB_REQUIRE(c_pad % 8 == 0 && c_pad >= c_in && c_in >= 1, "vae_pre: c_pad=%d must be a multiple of 8 >= c_in=%d", c_pad, c_in);
  AFB_REQUIRE(scale != 0.f, "vae_pre: scale factor is zero");
  const long long total = (long long)n * h * w * (c_pad / 8);
  vae_pre_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(z, static_cast<__nv_bfloat16*>(out), n, c_in, h * w, c_pad,
                                                                     1.0f / scale, shift);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int vae_post_launch(const void* x, int64_t x_ld, float* out, int n, int c_out, int h, int w, cudaStream_t stream) {
  AFB_REQUIRE(x && out && n >= 1 && h >= 1 && w >= 1, "vae_post: bad arguments");
  AFB_REQUIRE(c_out >= 1 && c_out <= 8 && x_ld % 8 == 0 && x_ld >= 8, "vae_post: c_out=%d (<= 8), x_ld=%lld", c_out, (long long)x_ld);
  const long long total = (long long)n * h * w;
  vae_post_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), x_ld, out, n, c_out,
                                                                      (long long)h * w);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int groupnorm_ws_floats(int n, long long hw) {
  long long slabs = (hw + 1023) / 1024;
  if (slabs > 256) slabs = 256;
  if (slabs < 1) slabs = 1;
  return int(n * slabs * GN_GROUPS * 2);
}

int groupnorm_launch(const void* x, void* y, const float* gamma, const float* beta, float* ws, int64_t ws_floats, int n,
                     long long hw, int c, float eps, int silu_on, cudaStream_t stream) {
  AFB_REQUIRE(x && y && gamma && beta && ws, "groupnorm: null pointer");
  AFB_REQUIRE(n >= 1 && hw >= 1, "groupnorm: empty input");
  AFB_REQUIRE(c % 64 == 0 && c <= 2048 && 2048 % c == 0, "groupnorm: channels=%d must be 64, 128, 256, 512, 1024 or 2048", c);
  long long slabs = (hw + 1023) / 1024;
  if (slabs > 256) slabs = 256;
  if (slabs < 1) slabs = 1;
  AFB_REQUIRE(ws_floats >= (long long)n * slabs * GN_GROUPS * 2, "groupnorm: workspace too small");
  const int rows_per_slab = int((hw + slabs - 1) / slabs);
  groupnorm_partial_kernel<<<dim3(unsigned(slabs), unsigned(n)), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ws, hw, c,
                                                                                     rows_per_slab);
  AFB_CHECK_CUDA(cudaGetLastError());
  const int sms = device_sm_count();
  long long blocks = (long long)sms * 8 / n;
  if (blocks < 1) blocks = 1;
  if (blocks > hw) blocks = hw;
  const int rows_per_block = int((hw + blocks - 1) / blocks);
  blocks = (hw + rows_per_block - 1) / rows_per_block;
  groupnorm_apply_kernel<<<dim3(unsigned(blocks), unsigned(n)), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), ws, gamma, beta, hw, c, int(slabs), rows_per_block,
      eps, silu_on);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return AFB_OK;
}

int upsample2x_launch(const void* x, void* y, int n, int h, int w, int c, cudaStream_t stream) {
  AFB_REQUIRE(x && y && n >= 1 && h >= 1 && w >= 1 && c % 8 == 0, "upsample2x: bad arguments");
  const long long total = (long long)n * 4 * h * w * (c / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 32;
  if (blocks > cap) blocks = cap;
  upsample2x_kernel<<<unsigned(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n,
                                                           h, w, c);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

int softmax_rows_launch(void* x, int64_t ld, long long rows, int cols, cudaStream_t stream) {
  AFB_REQUIRE(x && rows >= 1 && cols >= 8 && cols % 8 == 0 && ld % 8 == 0 && ld >= cols, "softmax_rows: bad arguments");
  AFB_REQUIRE(cols <= 49152, "softmax_rows: %d columns exceed the shared-memory row cache", cols);
  const size_t smem = size_t(cols) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    AFB_CHECK_CUDA(cudaFuncSetAttribute(softmax_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    attr = smem;
  }
  softmax_rows_kernel<<<unsigned(rows), 256, smem, stream>>>(static_cast<__nv_bfloat16*>(x), ld, cols);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

}  // namespace afb
