// Micro-benchmark: per-SMSP throughput of MUFU.EX2, FFMA2 and the mixed softmax inner loop on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 0.001f + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ex2(v[i]);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float2 r = __ffma2_rn(make_float2(v[i], v[i + 1]), make_float2(1.0001f, 1.0001f), make_float2(0.5f, 0.5f));
        v[i] = r.x; v[i + 1] = r.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  for (int warps : {4, 8, 16}) {
    long long h;
    k<0><<<148, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.EX2  warps/SM %2d: %.2f cyc per warp-instr per SMSP\n", warps, double(h) / (iters * 32.0 * (warps / 4)));
    k<1><<<148, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("FFMA2     warps/SM %2d: %.2f cyc per warp-instr per SMSP\n", warps, double(h) / (iters * 16.0 * (warps / 4)));
    k<2><<<148, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("FFMA      warps/SM %2d: %.2f cyc per warp-instr per SMSP\n", warps, double(h) / (iters * 32.0 * (warps / 4)));
  }
  return 0;
}
