// arcflow_b200 — persistent warp-specialised tcgen05 GEMM for the MMDiT linears.
//
//   out[b, r, n] = epilogue( sum_k A[b, r, k] * W[n, k] )        bf16 x bf16 -> fp32 (TMEM) -> bf16
//
// Replaces what the reference dispatches to cuBLASLt through torch.nn.Linear inside the diffusers
// blocks it instantiates (reference: lakonlab/models/architecture/arcflow/arcflux.py:60-88 builds the
// Linears; :180-249 calls them) plus the peft LoRA branch, which is folded in
// as a K-extension: A may be given as up to three K-segments (e.g. [y | y·A_lora^T] against
// [W | B_lora]), so "base + LoRA" is ONE accumulation in TMEM and no separate add kernel exists.
//
// Two kernels share one epilogue:
//  * gemm_bf16_2cta_kernel (default): a CTA PAIR (cluster of 2, one TPC) owns a 256 x 256 output tile.
//    tcgen05.mma.cta_group::2 (M256 N256 K16) is issued by the leader CTA only; each CTA stages its own
//    128 rows of A and its own 128-row half of the W tile (6-stage TMA ring, 32 KiB / stage / CTA), so the
//    UMMA shared-memory operand traffic per SM is 8 KiB per 128-cycle MMA (64 B/clk) instead of 12 KiB for
//    a 1-CTA 128 x 256 tile — measured on B200: the SS operand path saturates near 64 B/clk/SM, which
//    capped the 1-CTA kernel at ~70 % of the clock-limited peak. Both CTAs' TMA loads complete on the
//    leader's mbarrier; the leader's tcgen05.commit is multicast to both CTAs' "slot free" and
//    "accumulator full" barriers; the peer's epilogue warps release the accumulator with a remote arrive.
//  * gemm_bf16_kernel (AFB_GEMM_1CTA=1, or tiny problems): one CTA per 128 x 256 tile, 4-stage ring.
// Roles per CTA (CTA-pair kernel): warp 0 TMA producer and warp 1 MMA issuer — both run their loops with the whole warp
// converged and ONE elected lane issuing, so the operands live in uniform registers; each loop turns over once per k-block
// (512 tensor cycles) and is on or near the critical path: the kernel is a template over everything those loops would
// otherwise test at run time (tile width, convolution mode, QK epilogue, transposed W). Warps 2..9 are the epilogue, two per
// TMEM lane quarter, each taking half of a tile's 64-column chunks (tcgen05.ld 32x32b, 32 columns at a time -> bias /
// GELU-tanh / gate*y+residual / RMSNorm+RoPE -> bf16 into a swizzled shared-memory staging tile -> TMA store of whole
// 128-byte rows, 32 rows x 64 columns per warp and bulk group). The 1-CTA kernel keeps 4 epilogue warps, double-buffered.
// Per-row 16-byte global stores (the first version, still selectable with AFB_GEMM_TMA_STORE=0) touch 32 half-written
// sectors per warp instruction; ncu showed 2x the algorithmic DRAM traffic from the resulting sector fills.
// W loads carry an L2 evict_last policy and the output stores evict_first: the streamed output (0.6-0.9 GB per launch)
// must not push the 57-82 MB weight, which every M tile re-reads, out of the 126 MB L2.
// The accumulator is double-buffered in TMEM (2 x 256 columns): the epilogue of tile i overlaps the main
// loop of tile i+1. Tiles are walked N-fastest, or in column groups for wide W (tile_coords): A is read from HBM once per
// group and the group's W panels stay L2-resident.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/arcflow_b200.h"

namespace afb {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 512;
// epilogue staging: per epilogue warp 2 buffers of 32 rows x 64 bf16 columns (128-byte rows, SWIZZLE_128B, 1 KiB-aligned)
constexpr int EPI_COLS = 64;
constexpr int EPI_BUF_BYTES = 32 * EPI_COLS * 2;           // 4 KiB
constexpr int EPI_STAGE_BYTES = 4 * 2 * EPI_BUF_BYTES;     // 32 KiB per CTA
constexpr size_t GEMM_SMEM_BYTES =
    1024 /*align slack*/ + size_t(STAGES) * (A_STAGE_BYTES + B_STAGE_BYTES) + EPI_STAGE_BYTES + 256 /*barriers*/;
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;  // createpolicy.fractional.L2::evict_first, fraction 1.0
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;   // createpolicy.fractional.L2::evict_last, fraction 1.0

struct GemmParams {
  int batches, rows_per_batch, tiles_per_batch;
  int N;
  int nk_end[3];  // cumulative K-block (64) boundaries of the A segments
  int num_m_tiles, num_n_tiles;
  int epi;
  float alpha;   // accumulator scale (1 unless a runtime LoRA scale is set)
  int w_trans;   // W given as [K, N] row-major (dX = dY W): MN-major B operand (2-CTA kernel only)
  int nkb_w0;    // K blocks served by the first W buffer (the rest come from the second one; transposed mode)
  int tma_store; // epilogue through shared memory + TMA store (0: per-row 16-byte global stores)
  int l2_hints;  // W loads evict_last, output stores evict_first
  int gn;        // tile raster: N tiles per column group (0 / >= num_n_tiles: plain N-fastest order)
  int bn;        // N extent of a tile: 256, or 128 for narrow outputs (CTA-pair kernel, K-major W only)
  // implicit-GEMM 3x3 convolution (CTA-pair kernel): A is an NHWC image read through a 4-D map, an M tile is a 16 x 16 pixel
  // patch (each CTA: 8 rows of 16 pixels), K runs over 9 taps x (channels / 64) chunks — tap (dy, dx) is the same box
  // shifted by (dy - 1, dx - 1), zero-filled outside the image by the TMA unit
  int conv, conv_h, conv_w, conv_tw, conv_cpt;
  // AFB_EPI_BIAS_QKNORM_ROPE: per-head RMSNorm (x weight) + rotary embedding on output columns [0, qk_cols) — the q heads
  // then the k heads of a fused QKV projection, 128 columns per head — fused into the epilogue
  const __nv_bfloat16 *norm_q, *norm_k;  // [128] RMSNorm weights
  const float4* rope;                    // afb_rope_pack layout: [positions / 32][2 halves][16][32 lanes] x (cos, sin, cos, sin)
  int rope_row0, qk_cols;
  float norm_eps;
  __nv_bfloat16* out;
  long long out_ld, out_batch_stride;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* gate;
  long long gate_batch_stride;
  const __nv_bfloat16* res;
  long long res_ld, res_batch_stride;
};

// Tile raster. Tiles are handed out round-robin, so tiles [i * G, (i + 1) * G) (G = CTAs or clusters in flight) run
// together. Plain order walks N fastest: a wave spans all N tiles, i.e. streams ALL of W, every wave — for the MLP
// weights (77-96 MB, more than the part of the 126 MB L2 a streamed operand can hold) that re-reads W from DRAM ~20 times
// (ncu: 3.8 GB read for 0.3 GB of operands). With column groups of `gn` N-tiles the order is: group by group, inside a
// group M-major with N fastest — a wave spans gn N-tiles x (G / gn) M-rows, the group's W panels (gn x 256 x K x 2 B) stay
// L2-resident while the sweep goes down M, and A is streamed once per group.
__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& m_tile, int& n_tile) {
  if (p.gn <= 0 || p.gn >= p.num_n_tiles) {
    m_tile = tile / p.num_n_tiles;
    n_tile = tile - m_tile * p.num_n_tiles;
    return;
  }
  const int per_group = p.gn * p.num_m_tiles;
  const int g = tile / per_group;
  const int r = tile - g * per_group;
  const int rest = p.num_n_tiles - g * p.gn;
  const int width = rest < p.gn ? rest : p.gn;  // the last group may be narrower
  m_tile = r / width;
  n_tile = g * p.gn + (r - m_tile * width);
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanh_approx(inner));
}

// One thread = one accumulator row: 256 fp32 columns from TMEM in 8 chunks of 32.
// `bn` (tile width) is a compile-time constant at every call site: it is an argument only so that one body serves both widths.
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const int bn, uint32_t t_base, int n_tile, bool valid,
                                              __nv_bfloat16* out_row, const __nv_bfloat16* res_row,
                                              const __nv_bfloat16* gate_b, int c_begin = 0, int c_end = 1 << 30) {
  if (c_end > bn / 32) c_end = bn / 32;
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const int n0 = n_tile * bn + c * 32;
    if (n0 >= p.N) break;
    uint32_t v[32];
    tmem_ld_32x32(t_base + c * 32, v);
    tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = n0 + g * 8;
        if (n < p.N) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]) * p.alpha;
          if (p.bias) {
            const uint4 bv = *reinterpret_cast<const uint4*>(p.bias + n);
            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              f[2 * i] += bf16_lo(bw[i]);
              f[2 * i + 1] += bf16_hi(bw[i]);
            }
          }
          if (p.epi == AFB_EPI_BIAS_GELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = gelu_tanh_fast(f[i]);
          } else if (p.epi == AFB_EPI_BIAS_RES) {
            const uint4 rv = *reinterpret_cast<const uint4*>(res_row + n);
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              f[2 * i] += bf16_lo(rw[i]);
              f[2 * i + 1] += bf16_hi(rw[i]);
            }
          } else if (p.epi == AFB_EPI_BIAS_GATE_RES) {
            const uint4 gv = *reinterpret_cast<const uint4*>(gate_b + n);
            const uint4 rv = *reinterpret_cast<const uint4*>(res_row + n);
            const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              f[2 * i] = bf16_lo(rw[i]) + bf16_lo(gw[i]) * f[2 * i];
              f[2 * i + 1] = bf16_hi(rw[i]) + bf16_hi(gw[i]) * f[2 * i + 1];
            }
          }
          uint4 o;
          o.x = pack_bf16x2(f[0], f[1]);
          o.y = pack_bf16x2(f[2], f[3]);
          o.z = pack_bf16x2(f[4], f[5]);
          o.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(out_row + n) = o;
        }
      }
    }
  }
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, bool hint) {
  if (hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(
            reinterpret_cast<uint64_t>(map)),
        "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(L2_EVICT_FIRST)
        : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
  }
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One warp = 32 accumulator rows; the tile's 256 columns go out as 4 chunks of 64: tcgen05.ld -> epilogue math -> bf16
// into the warp's staging buffer (row r at r * 128 B, its 16-byte chunk j at ((j ^ (r & 7)) * 16): the SWIZZLE_128B
// pattern, conflict-free for the per-row writes) -> one TMA store per chunk. TMA clips rows >= rows_per_batch and
// columns >= N, so edge tiles need no masks on the store side; loads of bias / gate / residual stay guarded.
// Two bf16-rounded values at once (one F2FP instead of two F2F + shifts): lo / hi come back as fp32.
__device__ __forceinline__ void round_bf16_pair(float& a, float& b) {
  const uint32_t pk = pack_bf16x2(a, b);
  a = bf16_lo(pk);
  b = bf16_hi(pk);
}

// Latency note (ncu source view of the fused QKV launch, profiles/r02_ncu_qkv_fused_stalls.txt): every global load the
// epilogue issued right before its use — bias, gate, norm weights: one address for the whole warp — cost an L2 round trip
// (the L1 is almost entirely carved out as shared memory), 8-16 of them per 64-column chunk, serialised; 60 % of the
// samples of that launch sat on such a first use. Now each per-column vector is fetched ONCE per chunk as one coalesced
// 4-byte load per lane (lane i holds columns 2 i, 2 i + 1) before the TMEM load and handed out by warp shuffles; per-row
// data (residual, rotary pairs) is prefetched for the whole chunk.
template <bool QK>
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, const int bn, const CUtensorMap* tmOut, uint32_t t_base, int n_tile,
                                                  bool valid, int lane, int row0_warp, int b, uint8_t* stage,
                                                  int& buf, const __nv_bfloat16* res_row, const __nv_bfloat16* gate_b,
                                                  int conv_w0 = -1, int c_begin = 0, int c_end = 1 << 30,
                                                  bool single_buf = false) {
  // warp-uniform: anything of this warp's 32 rows inside the batch? (conv: row0_warp is the first of the warp's 2 image rows)
  const bool store_rows = row0_warp < (conv_w0 >= 0 ? p.conv_h : p.rows_per_batch);
  float qk_rstd = 0.f;
  // this row's rotary table entries (QKNORM_ROPE), afb_rope_pack layout: block of 32 positions, lane inside the block
  const long long rope_s = (long long)p.rope_row0 + row0_warp + lane;
  const float4* rope_row = QK ? p.rope + (rope_s >> 5) * (2 * 16 * 32) + (rope_s & 31) : nullptr;
  const bool has_res = p.epi == AFB_EPI_BIAS_RES || p.epi == AFB_EPI_BIAS_GATE_RES;
  if (c_end > bn / EPI_COLS) c_end = bn / EPI_COLS;
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const int n0 = n_tile * bn + c * EPI_COLS;
    if (n0 >= p.N) break;
    uint8_t* sbuf = stage + buf * EPI_BUF_BYTES;
    const bool qk = QK && n0 < p.qk_cols;   // QK instantiation = AFB_EPI_BIAS_QKNORM_ROPE launches only
    const int ncol = n0 + 2 * lane;  // the two columns this lane fetches for the warp (N is a multiple of 8)
    if (qk && (c & 1) == 0) {
      // first 64-column chunk of a head: sum of squares over the head's 128 (bias-added, bf16-rounded) columns. The
      // accumulator is simply read twice from TMEM — the statistics pass keeps nothing but the sum.
      uint2 b128 = make_uint2(0u, 0u);  // bias of columns n0 + 4 lane .. + 3
      if (p.bias) b128 = *reinterpret_cast<const uint2*>(p.bias + n0 + 4 * lane);
      float ss = 0.f;
#pragma unroll 1
      for (int hh = 0; hh < 4; ++hh) {
        uint32_t t[32];
        tmem_ld_32x32(t_base + c * EPI_COLS + hh * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // 4 columns per step: bias words from lane hh * 8 + i
          const uint32_t w0 = __shfl_sync(0xffffffffu, b128.x, hh * 8 + i), w1 = __shfl_sync(0xffffffffu, b128.y, hh * 8 + i);
          float f0 = __uint_as_float(t[4 * i]) * p.alpha + bf16_lo(w0), f1 = __uint_as_float(t[4 * i + 1]) * p.alpha + bf16_hi(w0);
          float f2 = __uint_as_float(t[4 * i + 2]) * p.alpha + bf16_lo(w1), f3 = __uint_as_float(t[4 * i + 3]) * p.alpha + bf16_hi(w1);
          round_bf16_pair(f0, f1);
          round_bf16_pair(f2, f3);
          ss = fmaf(f0, f0, ss);
          ss = fmaf(f1, f1, ss);
          ss = fmaf(f2, f2, ss);
          ss = fmaf(f3, f3, ss);
        }
      }
      qk_rstd = rsqrtf(ss * (1.0f / 128.0f) + p.norm_eps);
    }
    // ---- everything this chunk needs from global memory, issued before the TMEM load -------------------------------
    uint32_t bias_pk = 0u, gate_pk = 0u, nw_pk = 0u;
    if (ncol < p.N) {
      if (p.bias) bias_pk = *reinterpret_cast<const uint32_t*>(p.bias + ncol);
      if (p.epi == AFB_EPI_BIAS_GATE_RES) gate_pk = *reinterpret_cast<const uint32_t*>(gate_b + ncol);
    }
    if (qk) nw_pk = *reinterpret_cast<const uint32_t*>((n0 < (p.qk_cols >> 1) ? p.norm_q : p.norm_k) + (c & 1) * EPI_COLS + 2 * lane);
    float4 pre[16];  // qk: this row's 32 (cos, sin) pairs of the chunk; residual epilogues: pre[0..7] = the row's 64 bf16
    if (qk && valid) {
      const float4* src = rope_row + (c & 1) * (16 * 32);
#pragma unroll
      for (int i = 0; i < 16; ++i) pre[i] = __ldg(src + i * 32);  // 32 lanes x 16 B contiguous per instruction
    } else if (has_res && valid) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (n0 + i * 8 < p.N) pre[i] = *reinterpret_cast<const float4*>(res_row + n0 + i * 8);
    }
    if (lane == 0) {  // the previous store out of this buffer has read its source
      if (single_buf)
        bulk_wait_read<0>();
      else
        bulk_wait_read<1>();
    }
    __syncwarp();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];  // 32 columns at a time: the 320-thread kernel has 168 registers per thread
      tmem_ld_32x32(t_base + c * EPI_COLS + hh * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = n0 + hh * 32 + g * 8;
        const int j = hh * 4 + g;  // 16-byte chunk of this row's 128 bytes = 8 columns = the words of lanes 4 j .. 4 j + 3
        uint32_t bw[4], gw[4], ww[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // shuffles are warp-collective: outside the per-row predicate
          bw[i] = __shfl_sync(0xffffffffu, bias_pk, 4 * j + i);
          if (p.epi == AFB_EPI_BIAS_GATE_RES) gw[i] = __shfl_sync(0xffffffffu, gate_pk, 4 * j + i);
          if (qk) ww[i] = __shfl_sync(0xffffffffu, nw_pk, 4 * j + i);
        }
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (valid && n < p.N) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(v[g * 8 + 2 * i]) * p.alpha + bf16_lo(bw[i]);
            f[2 * i + 1] = __uint_as_float(v[g * 8 + 2 * i + 1]) * p.alpha + bf16_hi(bw[i]);
          }
          if (p.epi == AFB_EPI_BIAS_GELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = gelu_tanh_fast(f[i]);
          } else if (has_res) {
            const uint32_t rw[4] = {__float_as_uint(pre[j].x), __float_as_uint(pre[j].y), __float_as_uint(pre[j].z),
                                    __float_as_uint(pre[j].w)};
            if (p.epi == AFB_EPI_BIAS_RES) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                f[2 * i] += bf16_lo(rw[i]);
                f[2 * i + 1] += bf16_hi(rw[i]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                f[2 * i] = bf16_lo(rw[i]) + bf16_lo(gw[i]) * f[2 * i];
                f[2 * i + 1] = bf16_hi(rw[i]) + bf16_hi(gw[i]) * f[2 * i + 1];
              }
            }
          } else if (qk) {
            // the reference's rounding chain (diffusers RMSNorm + apply_rotary_emb on a bf16 Linear output): Linear ->
            // bf16; x * rstd -> bf16; * weight -> bf16; rotation in fp32 -> bf16. Same order as rmsnorm_rope_kernel.
            const float4 r0 = pre[2 * j], r1 = pre[2 * j + 1];  // this group's 4 (cos, sin) pairs
            const float cs[4] = {r0.x, r0.z, r1.x, r1.z}, sn[4] = {r0.y, r0.w, r1.y, r1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float x0 = f[2 * i], x1 = f[2 * i + 1];
              round_bf16_pair(x0, x1);
              x0 *= qk_rstd;
              x1 *= qk_rstd;
              round_bf16_pair(x0, x1);
              x0 *= bf16_lo(ww[i]);
              x1 *= bf16_hi(ww[i]);
              round_bf16_pair(x0, x1);
              f[2 * i] = x0 * cs[i] - x1 * sn[i];
              f[2 * i + 1] = x1 * cs[i] + x0 * sn[i];
            }
          }
          o.x = pack_bf16x2(f[0], f[1]);
          o.y = pack_bf16x2(f[2], f[3]);
          o.z = pack_bf16x2(f[4], f[5]);
          o.w = pack_bf16x2(f[6], f[7]);
        }
        *reinterpret_cast<uint4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
      }
    }
    fence_proxy_async();  // generic-proxy writes -> visible to the TMA (async proxy)
    __syncwarp();
    if (lane == 0) {
      if (store_rows) {
        if (conv_w0 >= 0)
          tma_store_4d(tmOut, sbuf, n0, conv_w0, row0_warp, b);
        else
          tma_store_3d(tmOut, sbuf, n0, row0_warp, b, p.l2_hints != 0);
      }
      bulk_commit();
    }
    if (!single_buf) buf ^= 1;
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sEpi = sB + STAGES * B_STAGE_BYTES;  // [4 warps][2][4 KiB], 1 KiB-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;        // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int nk = p.nk_end[2];

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_tile, n_tile;
        tile_coords(p, tile, m_tile, n_tile);
        const int b = m_tile / p.tiles_per_batch;
        const int r0 = (m_tile - b * p.tiles_per_batch) * BM;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
          const CUtensorMap* am;
          int kk;
          if (kb < p.nk_end[0]) {
            am = &tmA0;
            kk = kb;
          } else if (kb < p.nk_end[1]) {
            am = &tmA1;
            kk = kb - p.nk_end[0];
          } else {
            am = &tmA2;
            kk = kb - p.nk_end[1];
          }
          tma_load_3d(sA + stage * A_STAGE_BYTES, am, &full_bar[stage], kk * BK, r0, b);
          tma_load_2d(sB + stage * B_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n_tile * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------------
    // Whole warp converged; one elected lane issues (keeps descriptors in uniform registers).
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, false, false);
    const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA), 16, 1024);
    const uint64_t b_desc0 = make_sw128_desc(smem_u32(sB), 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = a_desc0 + uint64_t((stage * A_STAGE_BYTES) >> 4);
        const uint64_t bdesc = b_desc0 + uint64_t((stage * B_STAGE_BYTES) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_ss(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one_sync()) tc_commit(&acc_full[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------- epilogue warps -----------------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint8_t* stage = sEpi + q * 2 * EPI_BUF_BYTES;
    int buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_tile, n_tile;
      tile_coords(p, tile, m_tile, n_tile);
      const int b = m_tile / p.tiles_per_batch;
      const int r0w = (m_tile - b * p.tiles_per_batch) * BM + q * 32;
      const int r = r0w + lane;
      const bool valid = r < p.rows_per_batch;
      __nv_bfloat16* out_row = p.out + (long long)b * p.out_batch_stride + (long long)r * p.out_ld;
      const __nv_bfloat16* res_row =
          p.res ? p.res + (long long)b * p.res_batch_stride + (long long)r * p.res_ld : nullptr;
      const __nv_bfloat16* gate_b = p.gate ? p.gate + (long long)b * p.gate_batch_stride : nullptr;

      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      if (p.tma_store)
        epilogue_tile_tma<false>(p, BN, &tmOut, t_base, n_tile, valid, lane, r0w, b, stage, buf, res_row, gate_b);
      else
        epilogue_tile(p, BN, t_base, n_tile, valid, out_row, res_row, gate_b);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.tma_store && lane == 0) bulk_wait_all();  // the stores' global writes are complete before the grid ends
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ================================================================================================
// CTA-pair kernel (cta_group::2)
// ================================================================================================
constexpr int STAGES2 = 6;
constexpr int HALF_N = BN / 2;                    // W rows staged per CTA
constexpr int B2_STAGE_BYTES = HALF_N * BK * 2;   // 16 KiB
constexpr int STAGE2_BYTES = A_STAGE_BYTES + B2_STAGE_BYTES;
constexpr size_t GEMM2_SMEM_BYTES = 1024 + size_t(STAGES2) * STAGE2_BYTES + EPI_STAGE_BYTES + 256;
// CTA-pair kernel: EIGHT epilogue warps — two per TMEM lane quarter, each taking half of the tile's 64-column chunks (one head
// of a fused-QKV tile). The epilogue is a dependent chain per thread (TMEM load -> convert -> shared memory -> TMA store) with a
// single warp per scheduler; with the fused norm + rotation it no longer fit under the K = 3072 main loop (QKV launches 25 %
// slower). Each warp has ONE 4 KiB staging buffer (8 x 4 KiB = the same 32 KiB).
constexpr int GEMM2_THREADS = 32 * (2 + 8);
static_assert(GEMM2_SMEM_BYTES <= 232448 && GEMM_SMEM_BYTES <= 232448, "shared memory budget (227 KiB per CTA)");

// 10 warps are allocated as 12 (registers go out in groups of 4 warps): 65536 / (12 x 32) = 170 -> 168 registers per thread
// BN_T: tile width (256; 128 for outputs of <= 128 columns). CONV: implicit-GEMM 3x3 convolution producer / row mapping. QK: the
// RMSNorm + RoPE epilogue. WT: W read as the MN-major operand (dX = dY W). The first three were runtime fields of GemmParams in the first conv / fused-epilogue versions; the runtime
// tile width alone cost every launch 5-6 % (same-box A/B of the historical versions, profiles/r02_gemm_version_ab.json).
template <int BN_T, bool CONV, bool QK, bool WT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmOut,
                      const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES2 * A_STAGE_BYTES;
  uint8_t* sEpi = sB + STAGES2 * B2_STAGE_BYTES;  // [4 warps][2][4 KiB], 1 KiB-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES2]  both CTAs' TMA -> leader's MMA (leader copy is used)
  uint64_t* empty_bar = bars + STAGES2;       // [STAGES2]  leader's MMA -> each CTA's producer (multicast)
  uint64_t* acc_full = bars + 2 * STAGES2;    // [2]        leader's MMA -> each CTA's epilogue (multicast)
  uint64_t* acc_empty = acc_full + 2;         // [2]        both CTAs' epilogues -> leader's MMA (leader copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmB1);
    prefetch_tmap(&tmOut);
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's arrive.expect_tx; bytes from both CTAs
      mbar_init(&empty_bar[s], 1);  // one multicast commit per phase
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 16);  // 8 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // peer barriers initialised + both TMEM allocations done
  __syncthreads();     // (ordering is the cluster barrier's; this one is what compute-sanitizer racecheck models)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;  // 256 x 256 tiles
  const int nk = p.nk_end[2];
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------------
    // The whole warp runs the loop converged and ONE elected lane issues (as the MMA warp does): with `if (lane == 0)` around
    // the loop every TMA operand lived in a per-thread register and the compiler moved it to the uniform datapath through an
    // elect + R2UR loop per instruction — ~75 SASS instructions per k-block in one thread, next to the 512 cycles the k-block's
    // MMAs take, which is why a few more instructions in this loop cost whole percents (profiles/r02_gemm_version_ab.json).
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_tile, n_tile;
        tile_coords(p, tile, m_tile, n_tile);
        const int b = m_tile / p.tiles_per_batch;
        const int t_in = m_tile - b * p.tiles_per_batch;
        const int r0 = t_in * (2 * BM) + int(rank) * BM;
        const int half_n = BN_T >> 1;
        const int n0 = n_tile * BN_T + int(rank) * half_n;
        const uint32_t stage_tx = 2u * uint32_t(A_STAGE_BYTES + half_n * BK * 2);
        const int conv_th = CONV ? t_in / p.conv_tw : 0;
        const int conv_w0 = CONV ? (t_in - conv_th * p.conv_tw) * 16 : 0;
        const int conv_h0 = conv_th * 16 + int(rank) * 8;
        if (CONV) {
          for (int kb = 0; kb < nk; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t leader_full = mapa_u32(&full_bar[stage], 0);
            const int tap = kb / p.conv_cpt;
            const int cc = kb - tap * p.conv_cpt;
            const int dy = tap / 3;
            const int dx = tap - dy * 3;
            if (elect_one_sync()) {
              if (leader) mbar_expect_tx(&full_bar[stage], stage_tx);
              tma_load_4d_2cta(sA + stage * A_STAGE_BYTES, &tmA0, leader_full, cc * BK, conv_w0 + dx - 1, conv_h0 + dy - 1, b);
              tma_load_2d_2cta(sB + stage * B2_STAGE_BYTES, &tmB, leader_full, kb * BK, n0);
            }
            __syncwarp();
            if (++stage == STAGES2) {
              stage = 0;
              phase ^= 1;
            }
          }
        } else {
          // one loop per K-segment (x | LoRA | concat): the segment's tensor map is loop-invariant
          int kb = 0;
#pragma unroll 1
          for (int seg = 0; seg < 3; ++seg) {
            const CUtensorMap* am = seg == 0 ? &tmA0 : (seg == 1 ? &tmA1 : &tmA2);
            const int kb_end = p.nk_end[seg];
            for (int kk = 0; kb < kb_end; ++kb, ++kk) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              const uint32_t leader_full = mapa_u32(&full_bar[stage], 0);
              if (elect_one_sync()) {
                if (leader) mbar_expect_tx(&full_bar[stage], stage_tx);
                // (evict_first on the A loads was measured and rejected — profiles/r02_gemm_raster_sweep.json — and is gone)
                tma_load_3d_2cta(sA + stage * A_STAGE_BYTES, am, leader_full, kk * BK, r0, b);
                if (!WT) {
                  if (p.l2_hints)
                    tma_load_2d_2cta_hint(sB + stage * B2_STAGE_BYTES, &tmB, leader_full, kb * BK, n0, L2_EVICT_LAST);
                  else
                    tma_load_2d_2cta(sB + stage * B2_STAGE_BYTES, &tmB, leader_full, kb * BK, n0);
                } else {  // W is [K, N] row-major: two [64 k x 64 n] boxes, N contiguous (MN-major operand)
                  const CUtensorMap* bm = kb < p.nkb_w0 ? &tmB : &tmB1;
                  const int kr = (kb < p.nkb_w0 ? kb : kb - p.nkb_w0) * BK;
                  tma_load_2d_2cta(sB + stage * B2_STAGE_BYTES, bm, leader_full, n0, kr);
                  tma_load_2d_2cta(sB + stage * B2_STAGE_BYTES + HALF_N * BK, bm, leader_full, n0 + 64, kr);
                }
              }
              __syncwarp();
              if (++stage == STAGES2) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) ---------------------------
    if (leader) {
      const uint32_t idesc = WT ? make_idesc_bf16(2 * BM, BN, false, true) : make_idesc_bf16(2 * BM, BN_T, false, false);
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA), 16, 1024);
      // K-major W: 16-element k step = 32 bytes; MN-major W: 64-n chunks 8 KiB apart, k step = 16 rows of 128 bytes
      const uint64_t b_desc0 = WT ? make_sw128_desc(smem_u32(sB), HALF_N * BK, 1024) : make_sw128_desc(smem_u32(sB), 16, 1024);
      const uint64_t b_kstep = WT ? uint64_t(2048 >> 4) : uint64_t(2);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = a_desc0 + uint64_t((stage * A_STAGE_BYTES) >> 4);
          const uint64_t bdesc = b_desc0 + uint64_t((stage * B2_STAGE_BYTES) >> 4);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss_2cta(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k) * b_kstep, idesc,
                           (kb | k) != 0 ? 1u : 0u);
            tc_commit_2cta_mcast(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one_sync()) tc_commit_2cta_mcast(&acc_full[acc]);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------- epilogue warps (both CTAs) -----------------------------
    const int q = warp & 3;               // TMEM lane quarter (warps 2..9: each quarter twice)
    const int chalf = (warp - 2) >> 2;    // which half of the tile's chunks this warp takes
    uint8_t* stage = sEpi + (warp - 2) * EPI_BUF_BYTES;
    int buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_tile, n_tile;
      tile_coords(p, tile, m_tile, n_tile);
      const int b = m_tile / p.tiles_per_batch;
      const int t_in = m_tile - b * p.tiles_per_batch;
      int r0w = t_in * (2 * BM) + int(rank) * BM + q * 32;
      int r = r0w + lane;
      bool valid = r < p.rows_per_batch;
      int conv_w0 = -1;
      if (CONV) {  // this warp's 32 accumulator rows = 2 image rows x 16 pixels of the CTA's 8 x 16 patch
        const int th = t_in / p.conv_tw;
        conv_w0 = (t_in - th * p.conv_tw) * 16;
        r0w = th * 16 + int(rank) * 8 + q * 2;  // first image row of the warp
        const int hh = r0w + (lane >> 4), ww = conv_w0 + (lane & 15);
        valid = hh < p.conv_h && ww < p.conv_w;
        r = hh * p.conv_w + ww;  // pixel index inside the image (row of the NHWC matrix)
      }
      __nv_bfloat16* out_row = p.out + (long long)b * p.out_batch_stride + (long long)r * p.out_ld;
      const __nv_bfloat16* res_row =
          p.res ? p.res + (long long)b * p.res_batch_stride + (long long)r * p.res_ld : nullptr;
      const __nv_bfloat16* gate_b = p.gate ? p.gate + (long long)b * p.gate_batch_stride : nullptr;

      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      if (p.tma_store)
        epilogue_tile_tma<QK>(p, BN_T, &tmOut, t_base, n_tile, valid, lane, r0w, b, stage, buf, res_row, gate_b, conv_w0,
                          chalf * (BN_T / (2 * EPI_COLS)), (chalf + 1) * (BN_T / (2 * EPI_COLS)), true);
      else
        epilogue_tile(p, BN_T, t_base, n_tile, valid, out_row, res_row, gate_b, chalf * (BN_T / 64), (chalf + 1) * (BN_T / 64));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc_empty[acc], 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.tma_store && lane == 0) bulk_wait_all();  // the stores' global writes are complete before the grid ends
  }

  tc_fence_before();
  cluster_sync_all();  // the leader's MMAs read the peer's smem; nobody leaves before everybody is done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

}  // namespace

namespace {
// One launch site per instantiation (each needs its own dynamic-shared-memory attribute, set once).
template <int BN_T, bool CONV, bool QK, bool WT>
int launch_2cta(int clusters, cudaStream_t stream, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& a2,
                const CUtensorMap& b0, const CUtensorMap& b1, const CUtensorMap& out, const GemmParams& p) {
  static bool attr_set = false;
  if (!attr_set) {
    AFB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_2cta_kernel<BN_T, CONV, QK, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        int(GEMM2_SMEM_BYTES)));
    attr_set = true;
  }
  gemm_bf16_2cta_kernel<BN_T, CONV, QK, WT><<<2 * clusters, GEMM2_THREADS, GEMM2_SMEM_BYTES, stream>>>(a0, a1, a2, b0, b1, out, p);
  return AFB_OK;
}
}  // namespace

int gemm_launch(const afb_gemm_desc* d, cudaStream_t stream) {
  AFB_REQUIRE(d != nullptr, "gemm: null descriptor");
  AFB_REQUIRE(d->w && d->out && d->a[0], "gemm: null operand pointer");
  AFB_REQUIRE(d->batches >= 1 && d->rows_per_batch >= 1, "gemm: empty M (batches=%d rows=%d)",
              d->batches, d->rows_per_batch);
  AFB_REQUIRE(d->n >= 8 && d->n % 8 == 0, "gemm: N=%d must be a positive multiple of 8", d->n);
  AFB_REQUIRE(d->epilogue >= AFB_EPI_BIAS && d->epilogue <= AFB_EPI_BIAS_QKNORM_ROPE,
              "gemm: unknown epilogue %d", d->epilogue);
  if (d->epilogue == AFB_EPI_BIAS_QKNORM_ROPE) {
    AFB_REQUIRE(d->norm_q && d->norm_k && d->rope, "gemm: the QK-norm + RoPE epilogue needs norm_q, norm_k and the rope table");
    AFB_REQUIRE(d->qk_cols > 0 && d->qk_cols % 256 == 0 && d->qk_cols <= d->n,
                "gemm: qk_cols=%d must be a positive multiple of 256 (q heads then k heads, 128 columns each) <= N", d->qk_cols);
    AFB_REQUIRE(!d->w_transposed, "gemm: the QK-norm + RoPE epilogue is a forward-projection epilogue");
  }
  if (d->epilogue == AFB_EPI_BIAS_GATE_RES)
    AFB_REQUIRE(d->gate && d->res, "gemm: gate/residual epilogue needs gate and res pointers");
  if (d->epilogue == AFB_EPI_BIAS_RES) AFB_REQUIRE(d->res, "gemm: residual epilogue needs the res pointer");
  AFB_REQUIRE(d->out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
              "gemm: out must be 16-byte aligned with ld %% 8 == 0");

  static int force_1cta = -1;
  if (force_1cta < 0) {
    const char* e = getenv("AFB_GEMM_1CTA");
    force_1cta = (e && atoi(e) != 0) ? 1 : 0;
  }
  const bool two_cta = !force_1cta || d->w_transposed;
  const int tile_m = two_cta ? 2 * BM : BM;
  const int bn = (two_cta && !d->w_transposed && d->n <= 128) ? 128 : BN;  // narrow outputs: half-width tiles

  GemmParams p{};
  CUtensorMap tmA[3];
  int ktot = 0;
  int nseg = 0;
  for (int s = 0; s < 3; ++s) {
    const int ks = d->a_k[s];
    if (ks == 0) {
      AFB_REQUIRE(s > 0, "gemm: first A segment has K=0");
      tmA[s] = tmA[s - 1];
      p.nk_end[s] = p.nk_end[s - 1];
      continue;
    }
    AFB_REQUIRE(nseg == s, "gemm: A segments must be contiguous");
    AFB_REQUIRE(ks % BK == 0, "gemm: segment %d K=%d not a multiple of %d", s, ks, BK);
    AFB_REQUIRE(d->a[s] != nullptr, "gemm: segment %d null", s);
    const uint64_t dims[3] = {uint64_t(ks), uint64_t(d->rows_per_batch), uint64_t(d->batches)};
    const uint64_t bstride = d->batches > 1 ? uint64_t(d->a_batch_stride[s])
                                            : uint64_t(d->rows_per_batch) * uint64_t(d->a_ld[s]);
    const uint64_t strides[2] = {uint64_t(d->a_ld[s]) * 2, bstride * 2};
    const uint32_t box[3] = {BK, BM, 1};
    int rc = make_tmap_bf16(&tmA[s], d->a[s], 3, dims, strides, box);
    if (rc != AFB_OK) return rc;
    ktot += ks;
    p.nk_end[s] = ktot / BK;
    ++nseg;
  }
  CUtensorMap tmB, tmB1;
  p.nkb_w0 = ktot / BK;
  if (d->w_transposed) {
    // W: [K rows, N cols] row-major, optionally continued by a second buffer w2 after w_k rows
    const int k0 = d->w2 ? d->w_k : ktot;
    AFB_REQUIRE(k0 > 0 && k0 % BK == 0 && k0 <= ktot, "gemm: transposed W: bad w_k=%d (K=%d)", k0, ktot);
    AFB_REQUIRE(d->w_ld >= d->n && d->w_ld % 8 == 0, "gemm: transposed W: w_ld=%lld < N=%d", (long long)d->w_ld, d->n);
    const uint32_t box[2] = {64, BK};
    {
      const uint64_t dims[2] = {uint64_t(d->n), uint64_t(k0)};
      const uint64_t strides[1] = {uint64_t(d->w_ld) * 2};
      int rc = make_tmap_bf16(&tmB, d->w, 2, dims, strides, box);
      if (rc != AFB_OK) return rc;
    }
    tmB1 = tmB;
    if (d->w2) {
      AFB_REQUIRE(k0 < ktot && d->w2_ld >= d->n && d->w2_ld % 8 == 0, "gemm: transposed W: bad second W buffer");
      const uint64_t dims[2] = {uint64_t(d->n), uint64_t(ktot - k0)};
      const uint64_t strides[1] = {uint64_t(d->w2_ld) * 2};
      int rc = make_tmap_bf16(&tmB1, d->w2, 2, dims, strides, box);
      if (rc != AFB_OK) return rc;
    }
    p.w_trans = 1;
    p.nkb_w0 = k0 / BK;
  } else {
    AFB_REQUIRE(d->w2 == nullptr, "gemm: a second W buffer needs w_transposed");
    const uint64_t dims[2] = {uint64_t(ktot), uint64_t(d->n)};
    const uint64_t strides[1] = {uint64_t(d->w_ld) * 2};
    const uint32_t box[2] = {BK, uint32_t(two_cta ? bn / 2 : BN)};
    AFB_REQUIRE(d->w_ld >= ktot, "gemm: w_ld=%lld < total K=%d", (long long)d->w_ld, ktot);
    int rc = make_tmap_bf16(&tmB, d->w, 2, dims, strides, box);
    if (rc != AFB_OK) return rc;
    tmB1 = tmB;
  }

  // output map for the TMA-store epilogue: [N, rows, batches], box 64 columns x 32 rows (one epilogue warp's chunk)
  static int tma_store_env = -1, l2_hint_env = -1;
  if (tma_store_env < 0) {
    const char* e = getenv("AFB_GEMM_TMA_STORE");
    tma_store_env = (e && atoi(e) == 0) ? 0 : 1;
    e = getenv("AFB_GEMM_L2_HINTS");
    l2_hint_env = e ? atoi(e) : 1;
  }
  CUtensorMap tmOut = tmA[0];
  const uint64_t out_bs = d->batches > 1 ? uint64_t(d->out_batch_stride) : uint64_t(d->rows_per_batch) * uint64_t(d->out_ld);
  p.tma_store = tma_store_env && (out_bs % 8 == 0);
  if (d->epilogue == AFB_EPI_BIAS_QKNORM_ROPE) {
    AFB_REQUIRE(p.tma_store && two_cta, "gemm: the QK-norm + RoPE epilogue needs the TMA-store CTA-pair kernel");
    p.norm_q = static_cast<const __nv_bfloat16*>(d->norm_q);
    p.norm_k = static_cast<const __nv_bfloat16*>(d->norm_k);
    p.rope = static_cast<const float4*>(d->rope);
    p.rope_row0 = d->rope_row0;
    p.qk_cols = d->qk_cols;
    p.norm_eps = d->norm_eps > 0.f ? d->norm_eps : 1e-6f;
  }
  if (p.tma_store) {
    const uint64_t dims[3] = {uint64_t(d->n), uint64_t(d->rows_per_batch), uint64_t(d->batches)};
    const uint64_t strides[2] = {uint64_t(d->out_ld) * 2, out_bs * 2};
    const uint32_t box[3] = {EPI_COLS, 32, 1};
    int rc = make_tmap_bf16(&tmOut, d->out, 3, dims, strides, box);
    if (rc != AFB_OK) return rc;
  }
  p.l2_hints = l2_hint_env;
  // column-group raster (see tile_coords): keep a group's W panels within ~32 MB; AFB_GEMM_GN overrides (0 = plain order)
  static int gn_env = -2;
  if (gn_env == -2) {
    const char* e = getenv("AFB_GEMM_GN");
    gn_env = e ? atoi(e) : -1;
  }
  {
    const int ntiles = (d->n + bn - 1) / bn;
    const double panel = double(bn) * ktot * 2.0, w_total = double(d->n) * ktot * 2.0;
    int gn = 0;
    if (gn_env >= 0) {
      gn = gn_env;
    } else if (two_cta && !d->w_transposed && w_total > 40e6 && ntiles >= 24) {
      // measured (profiles/r02_gemm_raster_sweep.json): QKV (36 N-tiles, W 57 MB) and MLP-up (48 N-tiles, W 82 MB) drop from
      // 1.1 / 4.0 GB to 0.6 / 0.9 GB of DRAM reads with groups of 12-24 N-tiles; the K-heavy MLP-down / proj_out launches
      // (12 N-tiles, 6-8 MB panels) gain nothing from grouping and keep the plain order
      const int groups = int((w_total + 32e6 - 1) / 32e6);
      gn = (ntiles + groups - 1) / groups;
    }
    (void)panel;
    p.gn = gn >= ntiles ? 0 : gn;
  }

  p.batches = d->batches;
  p.rows_per_batch = d->rows_per_batch;
  p.tiles_per_batch = (d->rows_per_batch + tile_m - 1) / tile_m;
  p.N = d->n;
  p.num_m_tiles = p.tiles_per_batch * d->batches;
  p.num_n_tiles = (d->n + bn - 1) / bn;
  p.bn = bn;
  p.epi = d->epilogue;
  p.alpha = d->alpha != 0.f ? d->alpha : 1.0f;
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_ld = d->out_ld;
  p.out_batch_stride = d->out_batch_stride;
  p.bias = static_cast<const __nv_bfloat16*>(d->bias);
  p.gate = static_cast<const __nv_bfloat16*>(d->gate);
  p.gate_batch_stride = d->gate_batch_stride;
  p.res = static_cast<const __nv_bfloat16*>(d->res);
  p.res_ld = d->res_ld;
  p.res_batch_stride = d->res_batch_stride;

  static bool attr_set = false;
  if (!attr_set) {
    AFB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        int(GEMM_SMEM_BYTES)));
    attr_set = true;
  }
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int sms = device_sm_count();
  if (two_cta) {
    const int max_clusters = sms / 2;
    const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
    int rc;
    if (d->w_transposed)
      rc = launch_2cta<BN, false, false, true>(clusters, stream, tmA[0], tmA[1], tmA[2], tmB, tmB1, tmOut, p);
    else if (d->epilogue == AFB_EPI_BIAS_QKNORM_ROPE)
      rc = launch_2cta<BN, false, true, false>(clusters, stream, tmA[0], tmA[1], tmA[2], tmB, tmB1, tmOut, p);
    else if (bn == BN)
      rc = launch_2cta<BN, false, false, false>(clusters, stream, tmA[0], tmA[1], tmA[2], tmB, tmB1, tmOut, p);
    else
      rc = launch_2cta<BN / 2, false, false, false>(clusters, stream, tmA[0], tmA[1], tmA[2], tmB, tmB1, tmOut, p);
    if (rc != AFB_OK) return rc;
  } else {
    const int grid = num_tiles < sms ? num_tiles : sms;
    gemm_bf16_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(tmA[0], tmA[1], tmA[2], tmB, tmOut, p);
  }
  AFB_CHECK_CUDA(cudaGetLastError());
  return AFB_OK;
}

// 3x3 convolution, stride 1, zero padding 1, NHWC bf16, as an implicit GEMM on the CTA-pair kernel (see GemmParams::conv).
int conv3x3_launch(const afb_conv_desc* d, cudaStream_t stream) {
  AFB_REQUIRE(d != nullptr && d->x && d->w && d->out, "conv3x3: null pointer");
  AFB_REQUIRE(d->n >= 1 && d->h >= 1 && d->w_px >= 1, "conv3x3: empty image (%d x %d x %d)", d->n, d->h, d->w_px);
  AFB_REQUIRE(d->c_in >= 64 && d->c_in % 64 == 0, "conv3x3: c_in=%d must be a positive multiple of 64 (pad the input)", d->c_in);
  AFB_REQUIRE(d->c_out >= 8 && d->c_out % 8 == 0, "conv3x3: c_out=%d must be a positive multiple of 8 (pad the weight)", d->c_out);
  AFB_REQUIRE(d->epilogue == AFB_EPI_BIAS || d->epilogue == AFB_EPI_BIAS_RES, "conv3x3: epilogue %d not supported", d->epilogue);
  if (d->epilogue == AFB_EPI_BIAS_RES) AFB_REQUIRE(d->res, "conv3x3: residual epilogue needs the res pointer");
  const int64_t x_ld = d->x_ld ? d->x_ld : d->c_in, out_ld = d->out_ld ? d->out_ld : d->c_out;
  const int64_t res_ld = d->res_ld ? d->res_ld : d->c_out;
  AFB_REQUIRE(x_ld % 8 == 0 && out_ld % 8 == 0 && res_ld % 8 == 0, "conv3x3: pixel strides must be multiples of 8 elements");
  const int ktot = 9 * d->c_in;
  const int bn = d->c_out <= 128 ? 128 : BN;

  CUtensorMap tmA, tmB, tmOut;
  {
    const uint64_t dims[4] = {uint64_t(d->c_in), uint64_t(d->w_px), uint64_t(d->h), uint64_t(d->n)};
    const uint64_t strides[3] = {uint64_t(x_ld) * 2, uint64_t(d->w_px) * x_ld * 2, uint64_t(d->h) * d->w_px * x_ld * 2};
    const uint32_t box[4] = {BK, 16, 8, 1};
    int rc = make_tmap_bf16(&tmA, d->x, 4, dims, strides, box);
    if (rc != AFB_OK) return rc;
  }
  {
    const uint64_t dims[2] = {uint64_t(ktot), uint64_t(d->c_out)};
    const uint64_t strides[1] = {uint64_t(ktot) * 2};
    const uint32_t box[2] = {BK, uint32_t(bn / 2)};
    int rc = make_tmap_bf16(&tmB, d->w, 2, dims, strides, box);
    if (rc != AFB_OK) return rc;
  }
  {
    const uint64_t dims[4] = {uint64_t(d->c_out), uint64_t(d->w_px), uint64_t(d->h), uint64_t(d->n)};
    const uint64_t strides[3] = {uint64_t(out_ld) * 2, uint64_t(d->w_px) * out_ld * 2, uint64_t(d->h) * d->w_px * out_ld * 2};
    const uint32_t box[4] = {EPI_COLS, 16, 2, 1};
    int rc = make_tmap_bf16(&tmOut, d->out, 4, dims, strides, box);
    if (rc != AFB_OK) return rc;
  }
  GemmParams p{};
  p.conv = 1;
  p.conv_h = d->h;
  p.conv_w = d->w_px;
  p.conv_tw = (d->w_px + 15) / 16;
  p.conv_cpt = d->c_in / BK;
  const int tiles_h = (d->h + 15) / 16;
  p.batches = d->n;
  p.rows_per_batch = d->h * d->w_px;
  p.tiles_per_batch = tiles_h * p.conv_tw;
  p.N = d->c_out;
  p.nk_end[0] = p.nk_end[1] = p.nk_end[2] = ktot / BK;
  p.nkb_w0 = ktot / BK;
  p.num_m_tiles = p.tiles_per_batch * d->n;
  p.num_n_tiles = (d->c_out + bn - 1) / bn;
  p.bn = bn;
  p.epi = d->epilogue;
  p.alpha = 1.0f;
  p.tma_store = 1;
  p.l2_hints = 0;
  p.gn = 0;
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_ld = out_ld;
  p.out_batch_stride = int64_t(d->h) * d->w_px * out_ld;
  p.bias = static_cast<const __nv_bfloat16*>(d->bias);
  p.res = static_cast<const __nv_bfloat16*>(d->res);
  p.res_ld = res_ld;
  p.res_batch_stride = int64_t(d->h) * d->w_px * res_ld;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int max_clusters = device_sm_count() / 2;
  const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
  const int rc = bn == BN ? launch_2cta<BN, true, false, false>(clusters, stream, tmA, tmA, tmA, tmB, tmB, tmOut, p)
                          : launch_2cta<BN / 2, true, false, false>(clusters, stream, tmA, tmA, tmA, tmB, tmB, tmOut, p);
  if (rc != AFB_OK) return rc;
  AFB_CHECK_CUDA(cudaGetLastError());
  return AFB_OK;
}

}  // namespace afb
