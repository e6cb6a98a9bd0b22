// A/B driver for historical versions of csrc/gemm.cu: links ONE version's gemm.cu + host.cu and times the named launches of
// tools/profile_gemm_shapes.py (batch 8, 1024 px) with CUDA events, L2 flushed between launches. Built per version by
// tools/ubench/build_gemm_ab.sh from `git show <rev>:...`; the binaries print one JSON line each.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "include/arcflow_b200.h"

namespace afb {
int gemm_launch(const afb_gemm_desc* d, cudaStream_t stream);
const char* get_last_error();
}  // namespace afb

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__global__ void fill_kernel(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = uint32_t(i) * 2654435761u ^ seed;
    h ^= h >> 15;
    h *= 0x2c1b3c6dU;
    h ^= h >> 12;
    p[i] = __float2bfloat16((float(h & 0xFFFF) / 32768.0f - 1.0f) * scale);
  }
}
static __nv_bfloat16* alloc_fill(size_t n, uint32_t seed, float scale) {
  __nv_bfloat16* p;
  CK(cudaMalloc(&p, n * 2));
  fill_kernel<<<2048, 256>>>(p, n, seed, scale);
  return p;
}

int main(int argc, char** argv) {
  const char* tag = argc > 1 ? argv[1] : "?";
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  const int B = 8, St = 512, Si = 4096, S = St + Si, D = 3072, M = 12288, r = 256;
  __nv_bfloat16* y = alloc_fill((size_t)B * S * D, 1, 1.0f);
  __nv_bfloat16* attn = alloc_fill((size_t)B * S * D, 2, 1.0f);
  __nv_bfloat16* mlp = alloc_fill((size_t)B * S * M, 3, 0.3f);
  __nv_bfloat16* lt = alloc_fill((size_t)B * S * r, 4, 0.1f);
  __nv_bfloat16* h = alloc_fill((size_t)B * S * D, 5, 1.0f);
  __nv_bfloat16* gate = alloc_fill((size_t)B * D, 6, 0.3f);
  __nv_bfloat16* out = alloc_fill((size_t)B * S * M, 7, 0.0f);
  __nv_bfloat16* w = alloc_fill((size_t)M * (D + M + r), 8, 0.02f);
  __nv_bfloat16* bias = alloc_fill(M, 9, 0.02f);
  void* flush;
  CK(cudaMalloc(&flush, 512u << 20));
  CK(cudaDeviceSynchronize());

  struct Case {
    const char* name;
    int nseg;
    const __nv_bfloat16* a[3];
    int ak[3];
    long long ald[3];
    int rows, N, epi;
    bool gated;
  };
  const __nv_bfloat16* yi = y + (size_t)St * D;   // image rows of the joint buffer
  std::vector<Case> cases = {
      {"img_qkv", 1, {yi, nullptr, nullptr}, {D, 0, 0}, {D, 0, 0}, Si, 3 * D, AFB_EPI_BIAS, false},
      {"img_attn_out", 1, {attn + (size_t)St * D, nullptr, nullptr}, {D, 0, 0}, {D, 0, 0}, Si, D, AFB_EPI_BIAS_GATE_RES, true},
      {"img_mlp_up", 2, {yi, lt + (size_t)St * r, nullptr}, {D, r, 0}, {D, r, 0}, Si, M, AFB_EPI_BIAS_GELU, false},
      {"img_mlp_down", 2, {mlp + (size_t)St * M, lt + (size_t)St * r, nullptr}, {M, r, 0}, {M, r, 0}, Si, D, AFB_EPI_BIAS_GATE_RES, true},
      {"single_mlp_up", 2, {y, lt, nullptr}, {D, r, 0}, {D, r, 0}, S, M, AFB_EPI_BIAS_GELU, false},
      {"single_proj_out", 3, {attn, mlp, lt}, {D, M, r}, {D, M, r}, S, D, AFB_EPI_BIAS_GATE_RES, true},
  };
  printf("{\"version\": \"%s\", \"launches\": {", tag);
  bool first = true;
  for (const Case& c : cases) {
    afb_gemm_desc d;
    memset(&d, 0, sizeof(d));
    int K = 0;
    for (int s = 0; s < c.nseg; ++s) {
      d.a[s] = c.a[s];
      d.a_k[s] = c.ak[s];
      d.a_ld[s] = c.ald[s];
      d.a_batch_stride[s] = (long long)S * c.ald[s];
      K += c.ak[s];
    }
    d.batches = B;
    d.rows_per_batch = c.rows;
    d.w = w;
    d.w_ld = K;
    d.n = c.N;
    d.epilogue = c.epi;
    d.bias = bias;
    if (c.gated) {
      __nv_bfloat16* o = h + (c.rows == Si ? (size_t)St * D : 0);
      d.out = o;
      d.out_ld = D;
      d.out_batch_stride = (long long)S * D;
      d.res = o;
      d.res_ld = D;
      d.res_batch_stride = (long long)S * D;
      d.gate = gate;
      d.gate_batch_stride = D;
    } else {
      d.out = out;
      d.out_ld = c.N;
      d.out_batch_stride = (long long)c.rows * c.N;
    }
    if (afb::gemm_launch(&d, 0) != 0) {
      fprintf(stderr, "%s: %s\n", c.name, afb::get_last_error());
      return 1;
    }
    CK(cudaDeviceSynchronize());
    std::vector<float> ts;
    for (int i = 0; i < reps; ++i) {
      CK(cudaMemsetAsync(flush, i, 512u << 20, 0));
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, 0);
      afb::gemm_launch(&d, 0);
      cudaEventRecord(e1, 0);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    const double fl = 2.0 * B * c.rows * (double)c.N * K;
    printf("%s\"%s\": {\"ms_min\": %.4f, \"ms_med\": %.4f, \"tflops\": %.1f}", first ? "" : ", ", c.name, ts[0], ts[ts.size() / 2],
           fl / (ts[0] * 1e9));
    first = false;
  }
  printf("}}\n");
  return 0;
}
