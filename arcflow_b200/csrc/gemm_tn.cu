// arcflow_b200 — "TN" weight-gradient GEMM on tcgen05:   out[m, n] += sum_t A[t, m] * B[t, n]
//
// The contraction runs over TOKENS (rows of both operands), which is what every weight gradient of the adapter-only
// backward is: dW = dY^T X for the ArcFlow heads / norm_out and dB = dY^T T, dA = dT^T X for each LoRA pair (the
// reference gets these from torch autograd through nn.Linear / peft; parameters listed in
// configs/flux/arcflux_2nfe_k16.py:20-25). Both operands are row-major [tokens, features], i.e. MN-major for UMMA:
// TMA drops [64 tokens x 64 features] boxes (128-byte swizzle) and the MMA reads them through MN-major descriptors
// (LBO = next 64-feature chunk, SBO = next 8 token rows) — no transpose pass over HBM.
// One CTA per 128 x 256 output tile and token split; fp32 result added with 16-byte red.global.add.v4.f32 (the outputs
// are small: [1152, 3072] heads, [12288, 256] / [256, 3072] LoRA), 4-stage ring, roles as in gemm.cu. The operands are
// [batches, rows, features] views (3-D tensor maps): one launch contracts over every batch, and rows past the end of a
// batch are zero-filled by TMA instead of running into the next batch.
#include "common.cuh"
#include "../../include/arcflow_b200.h"

namespace afb {

namespace {

constexpr int TM = 128;
constexpr int TN = 256;
constexpr int TK = 64;  // tokens per stage
constexpr int TSTAGES = 4;
constexpr int CHUNK_BYTES = TK * 64 * 2;              // one [64 tokens x 64 features] box: 8 KiB
constexpr int TA_STAGE_BYTES = (TM / 64) * CHUNK_BYTES;  // 16 KiB
constexpr int TB_STAGE_BYTES = (TN / 64) * CHUNK_BYTES;  // 32 KiB
constexpr int TN_THREADS = 192;
constexpr size_t TN_SMEM_BYTES = 1024 + size_t(TSTAGES) * (TA_STAGE_BYTES + TB_STAGE_BYTES) + 256;

struct TnParams {
  int M, N, kblocks, kb_per_split, kb_per_batch;
  float* out;
  long long out_ld;
};

__global__ void __launch_bounds__(TN_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + TSTAGES * TA_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + TSTAGES * TB_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TSTAGES;
  uint64_t* acc_full = bars + 2 * TSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], TA_STAGE_BYTES + TB_STAGE_BYTES);
        const int b = kb / p.kb_per_batch;
        const int t0 = (kb - b * p.kb_per_batch) * TK;
        for (int c = 0; c < TM / 64; ++c)
          tma_load_3d(sA + stage * TA_STAGE_BYTES + c * CHUNK_BYTES, &tmA, &full_bar[stage], m0 + c * 64, t0, b);
        for (int c = 0; c < TN / 64; ++c)
          tma_load_3d(sB + stage * TB_STAGE_BYTES + c * CHUNK_BYTES, &tmB, &full_bar[stage], n0 + c * 64, t0, b);
        if (++stage == TSTAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(TM, TN, true, true);  // both operands MN-major
    const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA), CHUNK_BYTES, 1024);
    const uint64_t b_desc0 = make_sw128_desc(smem_u32(sB), CHUNK_BYTES, 1024);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint64_t adesc = a_desc0 + uint64_t((stage * TA_STAGE_BYTES) >> 4);
      const uint64_t bdesc = b_desc0 + uint64_t((stage * TB_STAGE_BYTES) >> 4);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < TK / 16; ++k)  // 16 token rows = 2048 bytes further into every chunk
          umma_ss(tmem_base, adesc + uint64_t((k * 2048) >> 4), bdesc + uint64_t((k * 2048) >> 4), idesc,
                  (kb > kb0 || k > 0) ? 1u : 0u);
        tc_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == TSTAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one_sync()) tc_commit(acc_full);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    if (kb1 > kb0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      float* orow = p.out + (long long)m * p.out_ld;
#pragma unroll 1
      for (int c = 0; c < TN / 32; ++c) {
        const int nb = n0 + c * 32;
        if (nb >= p.N) break;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        if (m < p.M) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            if (nb + i + 3 < p.N) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + nb + i), "f"(__uint_as_float(v[i])),
                           "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                           : "memory");
            } else {
              for (int j = i; j < i + 4; ++j)
                if (nb + j < p.N) atomicAdd(orow + nb + j, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TN);
  }
}

}  // namespace

int gemm_tn_batched_launch(const void* a, int64_t a_ld, int64_t a_bs, const void* b, int64_t b_ld, int64_t b_bs, float* out,
                           int64_t out_ld, int batches, int64_t rows, int m, int n, cudaStream_t stream);
int gemm_tn_launch(const void* a, int64_t a_ld, const void* b, int64_t b_ld, float* out, int64_t out_ld, int64_t tokens,
                   int m, int n, cudaStream_t stream) {
  return gemm_tn_batched_launch(a, a_ld, tokens * a_ld, b, b_ld, tokens * b_ld, out, out_ld, 1, tokens, m, n, stream);
}

int gemm_tn_batched_launch(const void* a, int64_t a_ld, int64_t a_bs, const void* b, int64_t b_ld, int64_t b_bs, float* out,
                           int64_t out_ld, int batches, int64_t rows, int m, int n, cudaStream_t stream) {
  AFB_REQUIRE(a && b && out, "gemm_tn: null pointer");
  AFB_REQUIRE(batches >= 1 && rows >= 1 && m >= 1 && n >= 1, "gemm_tn: empty problem");
  AFB_REQUIRE(m % 8 == 0 && n % 8 == 0 && a_ld % 8 == 0 && b_ld % 8 == 0 && a_bs % 8 == 0 && b_bs % 8 == 0,
              "gemm_tn: M, N, leading dims and batch strides must be multiples of 8");
  AFB_REQUIRE(out_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "gemm_tn: out must be 16-byte aligned, ld %% 4 == 0");
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {uint64_t(m), uint64_t(rows), uint64_t(batches)};
    const uint64_t strides[2] = {uint64_t(a_ld) * 2, uint64_t(batches > 1 ? a_bs : rows * a_ld) * 2};
    const uint32_t box[3] = {64, TK, 1};
    int rc = make_tmap_bf16(&tmA, a, 3, dims, strides, box);
    if (rc != AFB_OK) return rc;
  }
  {
    const uint64_t dims[3] = {uint64_t(n), uint64_t(rows), uint64_t(batches)};
    const uint64_t strides[2] = {uint64_t(b_ld) * 2, uint64_t(batches > 1 ? b_bs : rows * b_ld) * 2};
    const uint32_t box[3] = {64, TK, 1};
    int rc = make_tmap_bf16(&tmB, b, 3, dims, strides, box);
    if (rc != AFB_OK) return rc;
  }
  TnParams p{};
  p.M = m;
  p.N = n;
  p.kb_per_batch = int((rows + TK - 1) / TK);
  p.kblocks = p.kb_per_batch * batches;
  p.out = out;
  p.out_ld = out_ld;
  const int tiles = ((m + TM - 1) / TM) * ((n + TN - 1) / TN);
  int splits = (device_sm_count() + tiles - 1) / tiles;  // one wave; every extra split is another fp32 red pass
  if (splits > p.kblocks) splits = p.kblocks;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kblocks + splits - 1) / splits;
  splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;
  static bool attr_set = false;
  if (!attr_set) {
    AFB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TN_SMEM_BYTES)));
    attr_set = true;
  }
  dim3 grid((m + TM - 1) / TM, (n + TN - 1) / TN, splits);
  gemm_tn_kernel<<<grid, TN_THREADS, TN_SMEM_BYTES, stream>>>(tmA, tmB, p);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

}  // namespace afb
