// arcflow_b200 — attention backward on tcgen05 (adapter-only backward needs the full activation-gradient chain
// through the frozen attention; the reference gets it from torch autograd through F.scaled_dot_product_attention,
// reached at lakonlab/models/architecture/arcflow/arcflux.py:180-230 under gradient checkpointing :181-189).
//
// Two kernels, no atomics and no cross-CTA reduction (7 MMAs per tile pair instead of FlashAttention-2's 5; simple first):
//   dq kernel    CTA = 128 query rows of one (batch, head), loops over KV tiles:
//                  S = Q K^T, dP = dO V^T (SS) -> P = exp2(c S - lse), dS = scale P (dP - delta)  -> dQ += dS K   (TS)
//   dkdv kernel  CTA = 128 KV rows of one (batch, head), loops over Q tiles (everything transposed so the accumulators
//                stay CTA-local): S^T = K Q^T, dP^T = V dO^T (SS) -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q       (TS)
// P / dS are written back as packed bf16 into the TMEM columns of S / dP and consumed as the A operand; the B operands
// of the TS products are the resident [rows, 128] tiles read MN-major (as V in the forward).
// lse is the forward's log2-domain logsumexp, delta[b,h,s] = sum_d dO O (attn_delta_kernel).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/arcflow_b200.h"

namespace afb {

namespace {

constexpr int HD = 128;
constexpr int T = 128;  // tile rows (both q and kv)
constexpr int HALF_BYTES = 128 * 64 * 2;
constexpr int TILE_BYTES = 2 * HALF_BYTES;
constexpr int BWD_THREADS = 256;  // warp 0 TMA, warp 1 MMA, warps 2-3 idle, warps 4-7 math (one row per thread)
constexpr size_t BWD_SMEM_BYTES = 1024 + 6 * size_t(TILE_BYTES) + 1024 + 256;

struct BwdParams {
  int seq, heads, batch;
  float scale, scale_log2;
  const float* lse;    // [batch, heads, seq]
  const float* delta;  // [batch, heads, seq]
  __nv_bfloat16* out0;  // dq kernel: dQ;  dkdv kernel: dK
  __nv_bfloat16* out1;  // dkdv kernel: dV
  long long out_ld, out_bs;
};

__device__ __forceinline__ void math_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Loads one [128 rows x 128 cols] bf16 tile (two SW128 halves) at rows r0.. of (b, h)
__device__ __forceinline__ void load_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int h, int r0, int b) {
  for (int hf = 0; hf < 2; ++hf) tma_load_3d(dst + hf * HALF_BYTES, tm, bar, h * HD + hf * 64, r0, b);
}

// D[tmem] = A[smem, K-major over d] * B[smem, K-major over d]^T : 128 x 128 x 128
__device__ __forceinline__ void mma_ss_128(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc) {
  constexpr uint32_t idesc = make_idesc_bf16(T, T, false, false);
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    const uint32_t off = ((kk >> 2) * HALF_BYTES + (kk & 3) * 32) >> 4;
    umma_ss(d_tmem, a_desc + off, b_desc + off, idesc, kk > 0);
  }
}
// D[tmem] (+)= A[tmem bf16, 128 x 128] * B[smem rows = K, 128 cols MN-major]
__device__ __forceinline__ void mma_ts_128(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc_mn, bool acc) {
  constexpr uint32_t idesc = make_idesc_bf16(T, HD, false, true);
#pragma unroll
  for (int kk = 0; kk < T / 16; ++kk)
    umma_ts(d_tmem, a_tmem + kk * 8, b_desc_mn + uint64_t((kk * 2048) >> 4), idesc, (acc || kk > 0) ? 1u : 0u);
}

// ------------------------------------------------------------------------------------------------
// MODE 1: dK/dV kernel (outer = kv tile, inner = q tiles). MODE 0 is the un-pipelined dQ kernel, kept as the readable
// statement of the algorithm; the launcher uses attention_bwd_dq_kernel below for dQ.
// smem: R0, R1 resident tiles (Q,dO | K,V), ring of 2 x (X, Y) inner tiles (K,V | Q,dO).
// TMEM: [0,128) S or S^T, [128,256) dP or dP^T, [256,384) dQ | dV, [384,512) dK.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sR0 = smem;                    // Q (dq) | K (dkdv)
  uint8_t* sR1 = smem + TILE_BYTES;       // dO (dq) | V (dkdv)
  uint8_t* sX = smem + 2 * TILE_BYTES;    // ring[2]: K (dq) | Q (dkdv)
  uint8_t* sY = smem + 4 * TILE_BYTES;    // ring[2]: V (dq) | dO (dkdv)
  float* sLse = reinterpret_cast<float*>(smem + 6 * TILE_BYTES);  // [128] per-inner-tile lse (dkdv)
  float* sDelta = sLse + 128;                                     // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES + 1024);
  uint64_t* r_full = bars;          // resident tiles landed
  uint64_t* x_full = bars + 1;      // [2]
  uint64_t* x_empty = bars + 3;     // [2]
  uint64_t* s_full = bars + 5;      // S and dP ready (one commit after both MMAs)
  uint64_t* p_full = bars + 6;      // dkdv: P^T written; dq: unused
  uint64_t* ds_full = bars + 7;     // dS written
  uint64_t* acc_done = bars + 8;    // last accumulating MMAs retired
  uint64_t* iter_done = bars + 9;   // TS MMAs of this inner tile retired (S / dP columns reusable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * T;  // outer tile rows
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_inner = (p.seq + T - 1) / T;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmdO);
    mbar_init(r_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(ds_full, 4);
    mbar_init(acc_done, 1);
    mbar_init(iter_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(r_full, 2 * TILE_BYTES);
      load_tile(sR0, MODE == 0 ? &tmQ : &tmK, r_full, h, r0, b);
      load_tile(sR1, MODE == 0 ? &tmdO : &tmV, r_full, h, r0, b);
      for (int j = 0; j < n_inner; ++j) {
        const int st = j & 1;
        mbar_wait(&x_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&x_full[st], 2 * TILE_BYTES);
        load_tile(sX + st * TILE_BYTES, MODE == 0 ? &tmK : &tmQ, &x_full[st], h, j * T, b);
        load_tile(sY + st * TILE_BYTES, MODE == 0 ? &tmV : &tmdO, &x_full[st], h, j * T, b);
      }
    }
  } else if (warp == 1) {
    const uint64_t r0_k = make_sw128_desc(smem_u32(sR0), 16, 1024);   // K-major views (contraction over d)
    const uint64_t r1_k = make_sw128_desc(smem_u32(sR1), 16, 1024);
    const uint64_t x_k = make_sw128_desc(smem_u32(sX), 16, 1024);
    const uint64_t y_k = make_sw128_desc(smem_u32(sY), 16, 1024);
    const uint64_t x_mn = make_sw128_desc(smem_u32(sX), HALF_BYTES, 1024);  // MN-major views (contraction over rows)
    const uint64_t y_mn = make_sw128_desc(smem_u32(sY), HALF_BYTES, 1024);
    const uint32_t tS = tmem_base, tdP = tmem_base + 128, tA0 = tmem_base + 256, tA1 = tmem_base + 384;
    mbar_wait(r_full, 0);
    for (int j = 0; j < n_inner; ++j) {
      const int st = j & 1;
      const uint64_t so = uint64_t((st * TILE_BYTES) >> 4);
      mbar_wait(&x_full[st], (j >> 1) & 1);
      if (j > 0) mbar_wait(iter_done, (j - 1) & 1);  // previous dS / P consumed before S / dP are overwritten
      tc_fence_after();
      if (elect_one_sync()) {
        if (MODE == 0) {
          mma_ss_128(tS, r0_k, x_k + so);    // S    = Q  K_j^T
          mma_ss_128(tdP, r1_k, y_k + so);   // dP   = dO V_j^T
        } else {
          mma_ss_128(tS, r0_k, x_k + so);    // S^T  = K  Q_i^T
          mma_ss_128(tdP, r1_k, y_k + so);   // dP^T = V  dO_i^T
        }
        tc_commit(s_full);
      }
      __syncwarp();
      if (MODE == 1) {
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        if (elect_one_sync()) mma_ts_128(tA0, tS, y_mn + so, j > 0);   // dV += P^T dO_i
        __syncwarp();
      }
      mbar_wait(ds_full, j & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        if (MODE == 0) mma_ts_128(tA0, tS, x_mn + so, j > 0);          // dQ += dS K_j      (dS sits in the S columns)
        else mma_ts_128(tA1, tdP, x_mn + so, j > 0);                   // dK += dS^T Q_i    (dS^T sits in the dP columns)
        tc_commit(&x_empty[st]);
        tc_commit(iter_done);
        if (j == n_inner - 1) tc_commit(acc_done);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const uint32_t tS = tmem_base + lane_base, tdP = tS + 128;
    const int grow = r0 + row;  // global row of this thread in the OUTER tile
    const long long bh = ((long long)b * p.heads + h) * p.seq;
    float my_lse = 0.f, my_delta = 0.f;
    if (MODE == 0 && grow < p.seq) {
      my_lse = p.lse[bh + grow];
      my_delta = p.delta[bh + grow];
    }
    for (int j = 0; j < n_inner; ++j) {
      if (MODE == 1) {  // per-column (query) statistics of this inner tile
        const int q = j * T + row;
        math_bar_sync();  // previous iteration's readers are done with sLse / sDelta
        sLse[row] = q < p.seq ? p.lse[bh + q] : 0.f;
        sDelta[row] = q < p.seq ? p.delta[bh + q] : 0.f;
        math_bar_sync();
      }
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s[T];
#pragma unroll
      for (int i = 0; i < T / 32; ++i) tmem_ld_32x32(tS + i * 32, reinterpret_cast<uint32_t(&)[32]>(s[i * 32]));
      tmem_ld_wait();
      const int valid = p.seq - j * T;  // inner columns beyond the sequence contribute nothing
      uint32_t pk[T / 2];
#pragma unroll
      for (int i = 0; i < T; i += 2) {
        const float l0 = MODE == 0 ? my_lse : sLse[i], l1 = MODE == 0 ? my_lse : sLse[i + 1];
        float p0 = fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, -l0));
        float p1 = fast_exp2(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -l1));
        if (i >= valid) p0 = 0.f;
        if (i + 1 >= valid) p1 = 0.f;
        if (grow >= p.seq) p0 = p1 = 0.f;  // padded outer rows (zero-filled by TMA) must not pollute dK/dV
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      if (MODE == 1) {  // P^T is an MMA operand here: publish it first (dV can start while dS^T is computed)
#pragma unroll
        for (int i = 0; i < T / 32; ++i) tmem_st_32x16(tS + i * 16, reinterpret_cast<const uint32_t(&)[16]>(pk[i * 16]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // dS = scale * P * (dP - delta), 32 columns at a time
#pragma unroll
      for (int cidx = 0; cidx < T / 32; ++cidx) {
        uint32_t d[32];
        tmem_ld_32x32(tdP + cidx * 32, d);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int col = cidx * 32 + i;
          const float d0 = MODE == 0 ? my_delta : sDelta[col], d1 = MODE == 0 ? my_delta : sDelta[col + 1];
          const uint32_t pv = pk[col >> 1];
          o[i >> 1] = pack_bf16x2(p.scale * bf16_lo(pv) * (__uint_as_float(d[i]) - d0),
                                  p.scale * bf16_hi(pv) * (__uint_as_float(d[i + 1]) - d1));
        }
        tmem_st_32x16((MODE == 0 ? tS : tdP) + cidx * 16, o);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    // epilogue: accumulators -> bf16 -> global
    mbar_wait(acc_done, 0);
    tc_fence_after();
    const bool valid_row = grow < p.seq;
    for (int a = 0; a < (MODE == 0 ? 1 : 2); ++a) {
      __nv_bfloat16* dst = (a == 0 ? (MODE == 0 ? p.out0 : p.out1) : p.out0);  // dkdv: accumulator 0 = dV, 1 = dK
      __nv_bfloat16* orow = dst + (long long)b * p.out_bs + (long long)grow * p.out_ld + h * HD;
      const uint32_t tA = tmem_base + lane_base + 256 + a * 128;
#pragma unroll 1
      for (int i = 0; i < HD / 32; ++i) {
        uint32_t o[32];
        tmem_ld_32x32(tA + i * 32, o);
        tmem_ld_wait();
        if (valid_row) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
            w.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
            w.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
            w.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
            *reinterpret_cast<uint4*>(orow + i * 32 + g * 8) = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// dQ kernel, software-pipelined (supersedes attention_bwd_kernel<0>). The math leg of a tile (exp2, dS) is bound by
// SM-wide MUFU / FMA throughput at ~1500 cycles and the three MMAs of a tile take 1536, so the only way to go faster is to
// run them concurrently. TMEM has room for it here: S is double-buffered, S0 S1 dP dQ = 4 x 128 columns. The MMA warp
// issues  S(j), dP(j), dQ += dS(j-1) K(j-1)  per iteration, so S(j) / dP(j) are computed while the math warps are still on
// tile j-1 and dQ(j-1) runs under the exponentials of tile j. K tiles live until their dQ product has retired (3-stage
// ring), V tiles only until dP (2-stage ring); Q and dO are resident.
// ------------------------------------------------------------------------------------------------
constexpr int DQ_K_STAGES = 3, DQ_V_STAGES = 2;
constexpr size_t DQ_SMEM_BYTES = 1024 + size_t(2 + DQ_K_STAGES + DQ_V_STAGES) * TILE_BYTES + 256;

__global__ void __launch_bounds__(BWD_THREADS, 1)
attention_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + TILE_BYTES;
  uint8_t* sK = smem + 2 * TILE_BYTES;
  uint8_t* sV = sK + DQ_K_STAGES * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + DQ_V_STAGES * TILE_BYTES);
  uint64_t* r_full = bars;                       // Q, dO landed
  uint64_t* k_full = r_full + 1;                 // [3]
  uint64_t* k_empty = k_full + DQ_K_STAGES;      // [3]  dQ product of the tile retired
  uint64_t* v_full = k_empty + DQ_K_STAGES;      // [2]
  uint64_t* v_empty = v_full + DQ_V_STAGES;      // [2]  dP product of the tile retired
  uint64_t* s_full = v_empty + DQ_V_STAGES;      // [2]  S(j) ready in buffer j & 1
  uint64_t* dp_full = s_full + 2;                // dP(j) ready
  uint64_t* dp_free = dp_full + 1;               // math warps hold dP(j) in registers
  uint64_t* ds_full = dp_free + 1;               // [2]  dS(j) written over S(j)
  uint64_t* acc_done = ds_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * T;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_inner = (p.seq + T - 1) / T;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmdO);
    mbar_init(r_full, 1);
    for (int s = 0; s < DQ_K_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < DQ_V_STAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&ds_full[s], 4);
    }
    mbar_init(dp_full, 1);
    mbar_init(dp_free, 4);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t cS = 0, cdP = 256, cdQ = 384;  // TMEM columns: S buffers at 0 / 128

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(r_full, 2 * TILE_BYTES);
      load_tile(sQ, &tmQ, r_full, h, r0, b);
      load_tile(sdO, &tmdO, r_full, h, r0, b);
      for (int j = 0; j < n_inner; ++j) {
        const int ks = j % DQ_K_STAGES, vs = j % DQ_V_STAGES;
        mbar_wait(&k_empty[ks], ((j / DQ_K_STAGES) & 1) ^ 1);
        mbar_expect_tx(&k_full[ks], TILE_BYTES);
        load_tile(sK + ks * TILE_BYTES, &tmK, &k_full[ks], h, j * T, b);
        mbar_wait(&v_empty[vs], ((j / DQ_V_STAGES) & 1) ^ 1);
        mbar_expect_tx(&v_full[vs], TILE_BYTES);
        load_tile(sV + vs * TILE_BYTES, &tmV, &v_full[vs], h, j * T, b);
      }
    }
  } else if (warp == 1) {
    const uint64_t q_k = make_sw128_desc(smem_u32(sQ), 16, 1024);
    const uint64_t do_k = make_sw128_desc(smem_u32(sdO), 16, 1024);
    const uint64_t k_k = make_sw128_desc(smem_u32(sK), 16, 1024);
    const uint64_t v_k = make_sw128_desc(smem_u32(sV), 16, 1024);
    const uint64_t k_mn = make_sw128_desc(smem_u32(sK), HALF_BYTES, 1024);
    mbar_wait(r_full, 0);
    for (int j = 0; j <= n_inner; ++j) {
      if (j < n_inner) {
        const int ks = j % DQ_K_STAGES, vs = j % DQ_V_STAGES;
        mbar_wait(&k_full[ks], (j / DQ_K_STAGES) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          mma_ss_128(tmem_base + cS + (j & 1) * 128, q_k, k_k + uint64_t((ks * TILE_BYTES) >> 4));   // S(j) = Q K_j^T
          tc_commit(&s_full[j & 1]);
        }
        __syncwarp();
        mbar_wait(&v_full[vs], (j / DQ_V_STAGES) & 1);
        if (j > 0) mbar_wait(dp_free, (j - 1) & 1);  // dP(j-1) has left TMEM
        tc_fence_after();
        if (elect_one_sync()) {
          mma_ss_128(tmem_base + cdP, do_k, v_k + uint64_t((vs * TILE_BYTES) >> 4));                  // dP(j) = dO V_j^T
          tc_commit(dp_full);
          tc_commit(&v_empty[vs]);
        }
        __syncwarp();
      }
      if (j > 0) {
        const int i = j - 1, ks = i % DQ_K_STAGES;
        mbar_wait(&ds_full[i & 1], (i >> 1) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          mma_ts_128(tmem_base + cdQ, tmem_base + cS + (i & 1) * 128, k_mn + uint64_t((ks * TILE_BYTES) >> 4), i > 0);
          tc_commit(&k_empty[ks]);                                                                     // dQ += dS(i) K_i
          if (i == n_inner - 1) tc_commit(acc_done);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const int grow = r0 + row;
    const long long bh = ((long long)b * p.heads + h) * p.seq;
    const bool valid_row = grow < p.seq;
    const float my_lse = valid_row ? p.lse[bh + grow] : 0.f;
    const float my_delta = valid_row ? p.delta[bh + grow] : 0.f;
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2), nl = make_float2(-my_lse, -my_lse);
    for (int j = 0; j < n_inner; ++j) {
      const uint32_t tS = tmem_base + lane_base + cS + (j & 1) * 128;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[T];
#pragma unroll
      for (int i = 0; i < T / 32; ++i) tmem_ld_32x32(tS + i * 32, reinterpret_cast<uint32_t(&)[32]>(s[i * 32]));
      tmem_ld_wait();
      const int valid = p.seq - j * T;
      uint32_t pk[T / 2];
#pragma unroll
      for (int i = 0; i < T; i += 2) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, nl);
        float2 pv;
        pv.x = fast_exp2(x.x);
        pv.y = fast_exp2(x.y);
        if (i >= valid || !valid_row) pv.x = 0.f;
        if (i + 1 >= valid || !valid_row) pv.y = 0.f;
        pk[i >> 1] = pack_bf16x2(pv.x, pv.y);
      }
      // dP(j): pull it into registers in one go and hand the TMEM columns back (the MMA warp is waiting to issue dP(j+1))
      mbar_wait(dp_full, j & 1);
      tc_fence_after();
      uint32_t d[T];
#pragma unroll
      for (int i = 0; i < T / 32; ++i)
        tmem_ld_32x32(tmem_base + lane_base + cdP + i * 32, reinterpret_cast<uint32_t(&)[32]>(d[i * 32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dp_free);
      // dS = scale * P * (dP - delta), written as packed bf16 over S(j)
#pragma unroll
      for (int cidx = 0; cidx < T / 32; ++cidx) {
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int col = cidx * 32 + i;
          const uint32_t pv = pk[col >> 1];
          o[i >> 1] = pack_bf16x2(p.scale * bf16_lo(pv) * (__uint_as_float(d[col]) - my_delta),
                                  p.scale * bf16_hi(pv) * (__uint_as_float(d[col + 1]) - my_delta));
        }
        tmem_st_32x16(tS + cidx * 16, o);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[j & 1]);
    }
    // epilogue: dQ -> bf16 -> global
    mbar_wait(acc_done, 0);
    tc_fence_after();
    __nv_bfloat16* orow = p.out0 + (long long)b * p.out_bs + (long long)grow * p.out_ld + h * HD;
    const uint32_t tA = tmem_base + lane_base + cdQ;
#pragma unroll 1
    for (int i = 0; i < HD / 32; ++i) {
      uint32_t o[32];
      tmem_ld_32x32(tA + i * 32, o);
      tmem_ld_wait();
      if (valid_row) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
          w.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
          w.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
          w.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
          *reinterpret_cast<uint4*>(orow + i * 32 + g * 8) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// dK/dV kernel with the tensor pipe and the math warps overlapped (supersedes attention_bwd_kernel<1>, which stays
// selectable with AFB_ATTN_BWD_LEGACY=1 for A/B). S^T, dP^T, dV and dK fill the 512 TMEM columns, so nothing can be
// double-buffered — but nothing has to be: a tile's math has two legs that depend on different products (the
// exponentials need S^T, dS^T needs dP^T), and each leg's result feeds a different accumulating product. Issuing
//      dV += P^T(j) dO(j),   S^T(j+1) = K Q(j+1)^T,   dK += dS^T(j) Q(j),   dP^T(j+1) = V dO(j+1)^T
// in THAT order (the tensor pipe executes in issue order, so S^T(j+1) may overwrite the columns P^T(j) was read from, and
// dP^T(j+1) those of dS^T(j)) keeps 1024 cycles of independent tensor work queued behind each leg: the exponentials of
// tile j+1 run under dK(j) and dP^T(j+1), the dS leg under dV(j+1) and S^T(j+2). Q tiles are needed twice, two products
// apart (3-stage ring); dO tiles are released by dV (2-stage ring); K and V are resident.
// The math leg: P is kept in fp32 registers between the legs (no unpacking), a quarter of the exponentials run as a cubic
// on the FMA pipe (the leg is MUFU-bound: 128 x 128 ex2 = 1024 cycles of the 16 lanes an SM has), and the softmax scale is
// applied once to the dK accumulator instead of to every dS element. Columns / rows beyond the sequence need no masking:
// TMA zero-fills them, so the out-of-range rows of Q and dO contribute exact zeros to both accumulators.
// ------------------------------------------------------------------------------------------------
constexpr int KV_Q_STAGES = 3, KV_DO_STAGES = 2;
constexpr size_t DKDV_SMEM_BYTES = size_t(2 + KV_Q_STAGES + KV_DO_STAGES) * TILE_BYTES + 2048 + 256;
static_assert(DKDV_SMEM_BYTES <= 232448, "dK/dV kernel: shared memory over the 227 KiB a CTA can have");

__global__ void __launch_bounds__(BWD_THREADS, 1)
attention_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const BwdParams p) {
  // No alignment slack to spare (224 KiB of tiles): the dynamic shared memory of a kernel without static shared memory
  // starts on a 1 KiB boundary; checked, not assumed.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE_BYTES;
  uint8_t* sQ = smem + 2 * TILE_BYTES;
  uint8_t* sdO = sQ + KV_Q_STAGES * TILE_BYTES;
  float* sNl = reinterpret_cast<float*>(sdO + KV_DO_STAGES * TILE_BYTES);  // [2][128]  -lse of the inner tile's queries
  float* sNd = sNl + 256;                                                  // [2][128]  -delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sNd + 256);
  uint64_t* r_full = bars;                        // K, V landed
  uint64_t* q_full = r_full + 1;                  // [3]
  uint64_t* q_empty = q_full + KV_Q_STAGES;       // [3]  dK product of the tile retired
  uint64_t* do_full = q_empty + KV_Q_STAGES;      // [2]
  uint64_t* do_empty = do_full + KV_DO_STAGES;    // [2]  dV product of the tile retired
  uint64_t* s_full = do_empty + KV_DO_STAGES;     // S^T(j) ready
  uint64_t* p_full = s_full + 1;                  // P^T(j) written over S^T(j)
  uint64_t* dp_full = p_full + 1;                 // dP^T(j) ready
  uint64_t* ds_full = dp_full + 1;                // dS^T(j) written over dP^T(j)
  uint64_t* acc_done = ds_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * T;  // this CTA's KV rows
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_inner = (p.seq + T - 1) / T;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmdO);
    mbar_init(r_full, 1);
    for (int s = 0; s < KV_Q_STAGES; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (int s = 0; s < KV_DO_STAGES; ++s) {
      mbar_init(&do_full[s], 1);
      mbar_init(&do_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(dp_full, 1);
    mbar_init(ds_full, 4);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t cdV = 0, cdK = 128, cS = 256, cdP = 384;  // TMEM columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(r_full, 2 * TILE_BYTES);
      load_tile(sK, &tmK, r_full, h, r0, b);
      load_tile(sV, &tmV, r_full, h, r0, b);
      for (int j = 0; j < n_inner; ++j) {
        const int qs = j % KV_Q_STAGES, ds = j % KV_DO_STAGES;
        mbar_wait(&q_empty[qs], ((j / KV_Q_STAGES) & 1) ^ 1);
        mbar_expect_tx(&q_full[qs], TILE_BYTES);
        load_tile(sQ + qs * TILE_BYTES, &tmQ, &q_full[qs], h, j * T, b);
        mbar_wait(&do_empty[ds], ((j / KV_DO_STAGES) & 1) ^ 1);
        mbar_expect_tx(&do_full[ds], TILE_BYTES);
        load_tile(sdO + ds * TILE_BYTES, &tmdO, &do_full[ds], h, j * T, b);
      }
    }
  } else if (warp == 1) {
    const uint64_t k_k = make_sw128_desc(smem_u32(sK), 16, 1024);            // K-major views (contraction over d)
    const uint64_t v_k = make_sw128_desc(smem_u32(sV), 16, 1024);
    const uint64_t q_k = make_sw128_desc(smem_u32(sQ), 16, 1024);
    const uint64_t do_k = make_sw128_desc(smem_u32(sdO), 16, 1024);
    const uint64_t q_mn = make_sw128_desc(smem_u32(sQ), HALF_BYTES, 1024);   // MN-major views (contraction over rows)
    const uint64_t do_mn = make_sw128_desc(smem_u32(sdO), HALF_BYTES, 1024);
    const uint32_t tS = tmem_base + cS, tdP = tmem_base + cdP;
    auto stage = [](int st) { return uint64_t((st * TILE_BYTES) >> 4); };
    mbar_wait(r_full, 0);
    mbar_wait(&q_full[0], 0);
    tc_fence_after();
    if (elect_one_sync()) {
      mma_ss_128(tS, k_k, q_k);                                              // S^T(0) = K Q_0^T
      tc_commit(s_full);
    }
    __syncwarp();
    mbar_wait(&do_full[0], 0);
    tc_fence_after();
    if (elect_one_sync()) {
      mma_ss_128(tdP, v_k, do_k);                                            // dP^T(0) = V dO_0^T
      tc_commit(dp_full);
    }
    __syncwarp();
    for (int j = 0; j < n_inner; ++j) {
      const int qs = j % KV_Q_STAGES, ds = j % KV_DO_STAGES;
      const bool has_next = j + 1 < n_inner;
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        mma_ts_128(tmem_base + cdV, tS, do_mn + stage(ds), j > 0);           // dV += P^T(j) dO_j
        tc_commit(&do_empty[ds]);
      }
      __syncwarp();
      if (has_next) {
        const int qn = (j + 1) % KV_Q_STAGES;
        mbar_wait(&q_full[qn], ((j + 1) / KV_Q_STAGES) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          mma_ss_128(tS, k_k, q_k + stage(qn));                              // S^T(j+1), over the columns of P^T(j)
          tc_commit(s_full);
        }
        __syncwarp();
      }
      mbar_wait(ds_full, j & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        mma_ts_128(tmem_base + cdK, tdP, q_mn + stage(qs), j > 0);           // dK += dS^T(j) Q_j
        tc_commit(&q_empty[qs]);
        if (!has_next) tc_commit(acc_done);
      }
      __syncwarp();
      if (has_next) {
        const int dn = (j + 1) % KV_DO_STAGES;
        mbar_wait(&do_full[dn], ((j + 1) / KV_DO_STAGES) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          mma_ss_128(tdP, v_k, do_k + stage(dn));                            // dP^T(j+1), over the columns of dS^T(j)
          tc_commit(dp_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + cS, tdP = tmem_base + lane_base + cdP;
    const int grow = r0 + row;
    const long long bh = ((long long)b * p.heads + h) * p.seq;
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2);
    // statistics of inner query j * T + row, negated; finite (0) beyond the sequence so that 0 * P stays 0
    float nl_next = row < p.seq ? -p.lse[bh + row] : 0.f;
    float nd_next = row < p.seq ? -p.delta[bh + row] : 0.f;
    for (int j = 0; j < n_inner; ++j) {
      float* nlj = sNl + (j & 1) * 128;
      float* ndj = sNd + (j & 1) * 128;
      nlj[row] = nl_next;
      ndj[row] = nd_next;
      const int qn = (j + 1) * T + row;   // next tile's statistics: the loads fly under this tile's exponentials
      nl_next = qn < p.seq ? -p.lse[bh + qn] : 0.f;
      nd_next = qn < p.seq ? -p.delta[bh + qn] : 0.f;
      math_bar_sync();  // buffer j & 1 complete; its readers of tile j - 2 passed the barrier of tile j - 1
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s[T];
#pragma unroll
      for (int i = 0; i < T / 32; ++i) tmem_ld_32x32(tS + i * 32, reinterpret_cast<uint32_t(&)[32]>(s[i * 32]));
      tmem_ld_wait();
      // P^T = 2^(c S^T - lse[q]), fp32 in place; published as packed bf16 over the first 64 columns of S^T
#pragma unroll
      for (int cidx = 0; cidx < T / 32; ++cidx) {
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int col = cidx * 32 + i;
          const float4 nl = *reinterpret_cast<const float4*>(nlj + col);
          const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(s[col]), __uint_as_float(s[col + 1])), c2,
                                       make_float2(nl.x, nl.y));
          const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(s[col + 2]), __uint_as_float(s[col + 3])), c2,
                                       make_float2(nl.z, nl.w));
          float2 p0, p1;
          p0.x = fast_exp2(x0.x);
          p0.y = fast_exp2(x0.y);
          if ((i & 4) != 0) {
            p1 = exp2_poly2(x1);
          } else {
            p1.x = fast_exp2(x1.x);
            p1.y = fast_exp2(x1.y);
          }
          s[col] = __float_as_uint(p0.x);
          s[col + 1] = __float_as_uint(p0.y);
          s[col + 2] = __float_as_uint(p1.x);
          s[col + 3] = __float_as_uint(p1.y);
          o[i >> 1] = pack_bf16x2(p0.x, p0.y);
          o[(i >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
        }
        tmem_st_32x16(tS + cidx * 16, o);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // dS^T / scale = P^T (dP^T - delta[q]), 32 columns at a time, packed bf16 over the first 64 columns of dP^T
      mbar_wait(dp_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int cidx = 0; cidx < T / 32; ++cidx) {
        uint32_t d[32];
        tmem_ld_32x32(tdP + cidx * 32, d);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int col = cidx * 32 + i;
          const float4 nd = *reinterpret_cast<const float4*>(ndj + col);
          const float2 t0 = __fadd2_rn(make_float2(__uint_as_float(d[i]), __uint_as_float(d[i + 1])), make_float2(nd.x, nd.y));
          const float2 t1 = __fadd2_rn(make_float2(__uint_as_float(d[i + 2]), __uint_as_float(d[i + 3])), make_float2(nd.z, nd.w));
          const float2 u0 = __fmul2_rn(make_float2(__uint_as_float(s[col]), __uint_as_float(s[col + 1])), t0);
          const float2 u1 = __fmul2_rn(make_float2(__uint_as_float(s[col + 2]), __uint_as_float(s[col + 3])), t1);
          o[i >> 1] = pack_bf16x2(u0.x, u0.y);
          o[(i >> 1) + 1] = pack_bf16x2(u1.x, u1.y);
        }
        tmem_st_32x16(tdP + cidx * 16, o);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    // epilogue: dV, scale * dK -> bf16 -> global
    mbar_wait(acc_done, 0);
    tc_fence_after();
    const bool valid_row = grow < p.seq;
    for (int a = 0; a < 2; ++a) {
      __nv_bfloat16* dst = a == 0 ? p.out1 : p.out0;   // accumulator 0 = dV, 1 = dK
      const float f = a == 0 ? 1.0f : p.scale;
      __nv_bfloat16* orow = dst + (long long)b * p.out_bs + (long long)grow * p.out_ld + h * HD;
      const uint32_t tA = tmem_base + lane_base + (a == 0 ? cdV : cdK);
#pragma unroll 1
      for (int i = 0; i < HD / 32; ++i) {
        uint32_t o[32];
        tmem_ld_32x32(tA + i * 32, o);
        tmem_ld_wait();
        if (valid_row) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 w;
            w.x = pack_bf16x2(f * __uint_as_float(o[g * 8 + 0]), f * __uint_as_float(o[g * 8 + 1]));
            w.y = pack_bf16x2(f * __uint_as_float(o[g * 8 + 2]), f * __uint_as_float(o[g * 8 + 3]));
            w.z = pack_bf16x2(f * __uint_as_float(o[g * 8 + 4]), f * __uint_as_float(o[g * 8 + 5]));
            w.w = pack_bf16x2(f * __uint_as_float(o[g * 8 + 6]), f * __uint_as_float(o[g * 8 + 7]));
            *reinterpret_cast<uint4*>(orow + i * 32 + g * 8) = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int make_view_map(CUtensorMap* tm, const void* ptr, int64_t ld, int64_t bs, int batch, int seq, int heads) {
  const uint64_t dims[3] = {uint64_t(heads) * HD, uint64_t(seq), uint64_t(batch)};
  const uint64_t bstride = batch > 1 ? uint64_t(bs) : uint64_t(seq) * uint64_t(ld);
  const uint64_t strides[2] = {uint64_t(ld) * 2, bstride * 2};
  const uint32_t box[3] = {64, 128, 1};
  return make_tmap_bf16(tm, ptr, 3, dims, strides, box);
}

}  // namespace

int attn_delta_launch(const void* o, int64_t o_ld, int64_t o_bs, const void* d_o, int64_t do_ld, int64_t do_bs, float* delta,
                      int batch, int seq, int heads, cudaStream_t stream);

int attention_backward_launch(const afb_attn_bwd_desc* d, cudaStream_t stream) {
  AFB_REQUIRE(d != nullptr, "attention_backward: null descriptor");
  AFB_REQUIRE(d->q && d->k && d->v && d->o && d->d_o && d->lse && d->delta_ws && d->dq && d->dk && d->dv,
              "attention_backward: null pointer");
  AFB_REQUIRE(d->batch >= 1 && d->seq >= 1 && d->heads >= 1, "attention_backward: empty problem");
  AFB_REQUIRE(d->dqkv_ld % 8 == 0, "attention_backward: gradient leading dim must be a multiple of 8");
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_view_map(&tq, d->q, d->qkv_ld, d->qkv_batch_stride, d->batch, d->seq, d->heads)) != AFB_OK) return rc;
  if ((rc = make_view_map(&tk, d->k, d->qkv_ld, d->qkv_batch_stride, d->batch, d->seq, d->heads)) != AFB_OK) return rc;
  if ((rc = make_view_map(&tv, d->v, d->qkv_ld, d->qkv_batch_stride, d->batch, d->seq, d->heads)) != AFB_OK) return rc;
  if ((rc = make_view_map(&tdo, d->d_o, d->o_ld, d->o_batch_stride, d->batch, d->seq, d->heads)) != AFB_OK) return rc;
  if ((rc = attn_delta_launch(d->o, d->o_ld, d->o_batch_stride, d->d_o, d->o_ld, d->o_batch_stride, d->delta_ws, d->batch,
                              d->seq, d->heads, stream)) != AFB_OK)
    return rc;
  BwdParams p{};
  p.seq = d->seq;
  p.heads = d->heads;
  p.batch = d->batch;
  p.scale = d->scale > 0.f ? d->scale : 1.0f / sqrtf(float(HD));
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.lse = d->lse;
  p.delta = d->delta_ws;
  p.out_ld = d->dqkv_ld;
  p.out_bs = d->dqkv_batch_stride;
  static bool attr_set = false;
  if (!attr_set) {
    AFB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(DQ_SMEM_BYTES)));
    AFB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BWD_SMEM_BYTES)));
    AFB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(DKDV_SMEM_BYTES)));
    attr_set = true;
  }
  dim3 grid((d->seq + T - 1) / T, d->heads, d->batch);
  p.out0 = static_cast<__nv_bfloat16*>(d->dq);
  p.out1 = nullptr;
  attention_bwd_dq_kernel<<<grid, BWD_THREADS, DQ_SMEM_BYTES, stream>>>(tq, tk, tv, tdo, p);
  AFB_CHECK_CUDA(cudaGetLastError());
  p.out0 = static_cast<__nv_bfloat16*>(d->dk);
  p.out1 = static_cast<__nv_bfloat16*>(d->dv);
  static const bool legacy = [] {
    const char* e = getenv("AFB_ATTN_BWD_LEGACY");   // developer A/B: the un-overlapped dK/dV kernel
    return e != nullptr && e[0] == '1';
  }();
  if (legacy)
    attention_bwd_kernel<1><<<grid, BWD_THREADS, BWD_SMEM_BYTES, stream>>>(tq, tk, tv, tdo, p);
  else
    attention_bwd_dkdv_kernel<<<grid, BWD_THREADS, DKDV_SMEM_BYTES, stream>>>(tq, tk, tv, tdo, p);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return AFB_OK;
}

}  // namespace afb
