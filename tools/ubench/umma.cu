// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16, M = 128, K = 16) by operand source and N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/ubench/umma tools/ubench/umma.cu -lcuda
// One CTA per SM, one thread issues `iters` groups of 8 MMAs (one K = 128 slab) back to back and commits once at the
// end; cycles per MMA = (clock after the commit's mbarrier completes - clock before the first issue) / count.
#include <cstdio>
#include <cstdlib>
#include "../../arcflow_b200/csrc/common.cuh"
using namespace afb;

namespace afb {
void set_last_error(const char*, ...) {}
}

template <int N, bool A_TMEM, bool B_MN>
__global__ void __launch_bounds__(128, 1) umma_rate(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // 128 x 128 bf16, two SW128 halves (32 KiB)
  uint8_t* sB = smem + 32768;      // up to 256 x 128 bf16 (64 KiB)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, false, B_MN);
    const uint64_t a_desc = make_sw128_desc(smem_u32(sA), 16, 1024);
    const uint64_t b_desc = B_MN ? make_sw128_desc(smem_u32(sB), 16384, 1024) : make_sw128_desc(smem_u32(sB), 16, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = B_MN ? uint32_t((kk * 2048) >> 4) : uint32_t(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
        const uint32_t aoff = uint32_t(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
        const uint32_t d = tm + 256 + (it & 1) * 0;  // accumulate into one tile
        if (A_TMEM)
          umma_ts(d, tm + kk * 8, b_desc + off, idesc, 1u);
        else
          umma_ss(d, a_desc + aoff, b_desc + off, idesc, 1u);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

template <int N, bool A_TMEM, bool B_MN>
void run(const char* name, long long* d_out) {
  const int iters = 2048;
  const size_t smem = 1024 + 32768 + 65536;
  cudaFuncSetAttribute(umma_rate<N, A_TMEM, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long best = 1LL << 60;
  for (int rep = 0; rep < 3; ++rep) {
    umma_rate<N, A_TMEM, B_MN><<<sms, 128, smem>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      return;
    }
    long long h;
    cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    if (h < best) best = h;
  }
  const double per = double(best) / (iters * 8.0);
  printf("%-34s N=%3d  %7.1f clk/MMA  (ideal %3d)  smem operand bytes/clk %.0f\n", name, N, per, N / 2,
         ((A_TMEM ? 0 : 4096) + N * 32) / per);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<128, false, false>("SS  A smem K-major, B smem K-major", d_out);
  run<64, false, false>("SS  A smem K-major, B smem K-major", d_out);
  run<256, false, false>("SS  A smem K-major, B smem K-major", d_out);
  run<128, true, false>("TS  A tmem, B smem K-major", d_out);
  run<64, true, false>("TS  A tmem, B smem K-major", d_out);
  run<256, true, false>("TS  A tmem, B smem K-major", d_out);
  run<128, true, true>("TS  A tmem, B smem MN-major", d_out);
  run<64, true, true>("TS  A tmem, B smem MN-major", d_out);
  run<128, false, true>("SS  A smem K-major, B smem MN-major", d_out);
  return 0;
}
