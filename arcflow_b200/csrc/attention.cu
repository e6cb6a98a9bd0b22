// arcflow_b200 — joint text+image attention forward (non-causal, head_dim 128) on tcgen05 / TMEM.
//
// Replaces F.scaled_dot_product_attention as reached through the diffusers attention processors the
// reference's blocks use (reference call sites: lakonlab/models/architecture/arcflow/arcflux.py:180-230,
// arcqwen.py:136-155; semantics SURVEY.md Appendix A.1/A.4: scale 1/sqrt(128), no mask, no dropout,
// text tokens first in the joint sequence).
//
// One CTA owns 256 query rows of one (batch, head): two 128-row Q tiles that share every K/V tile.
//   warp 0        TMA producer : Q once, then K_j / V_j into 2-stage rings (128 x 128 bf16, SW128)
//   warp 1        MMA issuer   : S_t = Q_t K_j^T  (SS, M128 N128 K16 x8)  -> TMEM
//                                O_t += P_t V_j   (TS: P read from TMEM, V MN-major from smem)
//   warps 2..3    idle (pad the producer warpgroup so setmaxnreg applies per warpgroup)
//   warps 4..7    softmax WG 0 : one query row per thread; S from TMEM -> exp2 -> P (bf16) back into
//   warps 8..11   softmax WG 1   the same TMEM columns; lazy O rescale; final O / l -> global
// The MMA order  PV_A(j), QK_A(j+1), PV_B(j), QK_B(j+1)  keeps the tensor pipe busy on one tile while
// the other tile's warpgroup is in its softmax (ping-pong).  TMEM: S_A S_B O_A O_B = 4 x 128 columns.
// The exp phase is MUFU-bound (128 ex2 per row per tile = the tensor time of one tile), so
//   * the two warpgroups take turns in it (named-barrier hand-off) instead of halving each other's rate,
//   * one pair in four is evaluated on the FMA pipe (Cody-Waite split + cubic minimax for 2^frac),
//   * scale/subtract, the polynomial and the row sum use packed f32x2 instructions,
//   * registers move from the producer warpgroup to the softmax warpgroups (setmaxnreg).
//
// BOUNDED variant (afb_attn_desc.score_bound > 0): the caller guarantees |scale * q.k| <= score_bound for every pair — the
// MMDiT blocks RMS-normalise q and k per head, so ||q|| ||k|| / sqrt(128) is bounded by the norm weights alone. Then
// P = 2^(c s - B) with the FIXED reference B = score_bound * log2(e) needs no row maximum at all: P <= 1 by construction,
// P >= 2^(-2B) stays a normal bf16 / fp32 number for B <= 60, and O / l and lse = B + log2(l) are the same softmax. The
// row-max leg (350-540 cycles of the serial S -> max -> exp -> P -> PV chain per Q tile), the lazy O rescale and its
// branch leave the kernel; the exp phase starts on the first 64 columns while the second 64 are still being read from TMEM.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/arcflow_b200.h"

namespace afb {
void count_launch(int n);

namespace {

constexpr int HD = 128;
constexpr int QT = 128;
constexpr int NQT = 2;
constexpr int KT = 128;
constexpr int K_STAGES = 3;
constexpr int V_STAGES = 2;
constexpr int HALF_BYTES = 128 * 64 * 2;    // one 64-column SW128 half of a 128 x 128 tile
constexpr int TILE_BYTES = 2 * HALF_BYTES;  // 32 KiB
constexpr int ATT_THREADS = 32 * (4 + 4 * NQT);
constexpr int SOFTMAX_REGS = 208;
constexpr int PRODUCER_REGS = 88;
// setmaxnreg draws from the registers this CTA released: (SOFTMAX - 168) * 256 <= (168 - PRODUCER) * 128
static_assert((SOFTMAX_REGS - 168) * 256 <= (168 - PRODUCER_REGS) * 128, "register hand-off does not balance");
constexpr int BAR_TURN_A = 1;  // named barriers: "warpgroup A may enter its exp phase" / same for B
constexpr int BAR_TURN_B = 2;
constexpr size_t ATT_SMEM_BYTES = 1024 + size_t(NQT + K_STAGES + V_STAGES) * TILE_BYTES + 256;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: skip O rescale while max grows < 2^8

struct AttnParams {
  int seq, heads, batch;
  float scale_log2;  // softmax scale * log2(e)
  float bound_log2;  // BOUNDED: fixed softmax reference point, score_bound * log2(e)
  __nv_bfloat16* o;
  long long o_ld, o_batch_stride;
  float* lse;  // optional [batch, heads, seq]: log2-domain logsumexp of the scaled scores (for the backward)
};

// Developer trace (TRACE instantiations): per-iteration clock stamps of one softmax thread per warpgroup of CTA 0.
__device__ long long g_attn_trace[2][64][8];
__device__ __forceinline__ void trace_stamp(bool on, int t, int j, int slot) {
  if (on && j < 64) g_attn_trace[t][j][slot] = clock64();
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ void named_bar_sync(int id) {
  asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id) {
  asm volatile("bar.arrive %0, 256;" ::"r"(id) : "memory");
}

// TRACE: record clock stamps of the softmax hand-shake (developer builds only; the product instantiation carries no
// instrumentation — the probes' predicates and address arithmetic cost 8 % of the kernel). TURNS: the two warpgroups
// alternate in the exp phase.
template <bool TRACE, bool TURNS, bool BOUNDED>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NQT * TILE_BYTES;
  uint8_t* sV = sK + K_STAGES * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + V_STAGES * TILE_BYTES);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* k_full = q_full + 1;             // [K_STAGES]
  uint64_t* k_empty = k_full + K_STAGES;     // [K_STAGES]
  uint64_t* v_full = k_empty + K_STAGES;     // [V_STAGES]
  uint64_t* v_empty = v_full + V_STAGES;     // [V_STAGES]
  uint64_t* s_full = v_empty + V_STAGES;     // [NQT]  MMA -> softmax : S_t(j) ready
  uint64_t* p_full = s_full + NQT;           // [NQT][2]  softmax -> MMA : columns [64 h, 64 h + 64) of P_t(j) written
  uint64_t* o_done = p_full + 2 * NQT;       // [NQT]  MMA -> softmax : PV_t(j) retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + NQT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (NQT * QT);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.seq + KT - 1) / KT;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < K_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < V_STAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < NQT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[2 * t], 4);  // one arrive per softmax warp
      mbar_init(&p_full[2 * t + 1], 4);
      mbar_init(&o_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------------------
    if (lane == 0) {
      mbar_expect_tx(q_full, NQT * TILE_BYTES);
      for (int t = 0; t < NQT; ++t)
        for (int hf = 0; hf < 2; ++hf)
          tma_load_3d(sQ + t * TILE_BYTES + hf * HALF_BYTES, &tmQ, q_full, h * HD + hf * 64,
                      q0 + t * QT, b);
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j % K_STAGES, vs = j % V_STAGES;
        const uint32_t kph = (j / K_STAGES) & 1, vph = (j / V_STAGES) & 1;
        mbar_wait(&k_empty[ks], kph ^ 1);
        mbar_expect_tx(&k_full[ks], TILE_BYTES);
        for (int hf = 0; hf < 2; ++hf)
          tma_load_3d(sK + ks * TILE_BYTES + hf * HALF_BYTES, &tmK, &k_full[ks], h * HD + hf * 64,
                      j * KT, b);
        mbar_wait(&v_empty[vs], vph ^ 1);
        mbar_expect_tx(&v_full[vs], TILE_BYTES);
        for (int hf = 0; hf < 2; ++hf)
          tma_load_3d(sV + vs * TILE_BYTES + hf * HALF_BYTES, &tmV, &v_full[vs], h * HD + hf * 64,
                      j * KT, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------------
    // The whole warp runs this loop converged; one elected lane issues the tcgen05 instructions.
    {
      constexpr uint32_t idesc_qk = make_idesc_bf16(QT, KT, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(QT, HD, false, true);  // V is MN-major
      const uint64_t q_desc = make_sw128_desc(smem_u32(sQ), 16, 1024);
      const uint64_t k_desc = make_sw128_desc(smem_u32(sK), 16, 1024);
      const uint64_t v_desc = make_sw128_desc(smem_u32(sV), HALF_BYTES, 1024);

      // descriptor start addresses are in 16-byte units, so advancing is an add on the low word
      auto issue_qk = [&](int t, int st) {
        const uint64_t a0 = q_desc + uint64_t((t * TILE_BYTES) >> 4);
        const uint64_t b0 = k_desc + uint64_t((st * TILE_BYTES) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint32_t off = ((kk >> 2) * HALF_BYTES + (kk & 3) * 32) >> 4;
            umma_ss(tmem_base + t * KT, a0 + off, b0 + off, idesc_qk, kk > 0);
          }
        }
        __syncwarp();
      };
      // P_t is published in two 64-column halves, so the first half of the PV product overlaps the second half of
      // the exp phase (the K = 128 contraction is 8 independent K = 16 MMAs anyway).
      auto issue_pv = [&](int t, int st, int hf, bool acc) {
        const uint64_t b0 = v_desc + uint64_t((st * TILE_BYTES) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = hf * (KT / 32); kk < (hf + 1) * (KT / 32); ++kk) {
            // A = P_t (bf16, TMEM, 8 columns per K=16); B = V rows [16 kk, 16 kk + 16) of the stage
            umma_ts(tmem_base + NQT * KT + t * HD, tmem_base + t * KT + kk * 8, b0 + uint64_t((kk * 2048) >> 4),
                    idesc_pv, (acc || kk > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one_sync()) tc_commit(bar);
        __syncwarp();
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < NQT; ++t) {
        issue_qk(t, 0);
        commit(&s_full[t]);
      }
      commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % V_STAGES;
        const uint32_t ph = (j / V_STAGES) & 1;
        const bool has_next = (j + 1) < n_kv;
        const int nst = (j + 1) % K_STAGES;
        const uint32_t nph = ((j + 1) / K_STAGES) & 1;
        mbar_wait(&v_full[st], ph);
        if (has_next) mbar_wait(&k_full[nst], nph);
        for (int t = 0; t < NQT; ++t) {
          mbar_wait(&p_full[2 * t], j & 1);
          tc_fence_after();
          issue_pv(t, st, 0, j > 0);
          mbar_wait(&p_full[2 * t + 1], j & 1);
          tc_fence_after();
          issue_pv(t, st, 1, true);
          commit(&o_done[t]);
          if (has_next) {
            issue_qk(t, nst);
            commit(&s_full[t]);
          }
        }
        commit(&v_empty[st]);
        if (has_next) commit(&k_empty[nst]);
      }
    }
  }
  } else {
    // ------------------------------- softmax warpgroups -------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SOFTMAX_REGS));
    const int t = (warp - 4) >> 2;  // which Q tile
    const int qd = warp & 3;        // TMEM lane quarter
    const int row = qd * 32 + lane;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + t * KT;
    const uint32_t tO = tmem_base + lane_base + NQT * KT + t * HD;
    const float c = p.scale_log2;
    float m = -INFINITY;
    float l = 0.f;

    const bool tr = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && row == 0;
    const int my_turn = t == 0 ? BAR_TURN_A : BAR_TURN_B;
    const int next_turn = t == 0 ? BAR_TURN_B : BAR_TURN_A;
    if (TURNS && t == 1) named_bar_arrive(BAR_TURN_A);  // warpgroup A owns the first turn
    for (int j = 0; j < n_kv; ++j) {
      if (TRACE) trace_stamp(tr, t, j, 0);
      mbar_wait(&s_full[t], j & 1);
      // PV_t(j-1) was issued before QK_t(j) and the tensor pipe retires in order, so this phase has already completed: the
      // wait is one try_wait that observes it (every o_done phase gets its consumer; the O rescale below relies on it)
      if (j > 0) mbar_wait(&o_done[t], (j - 1) & 1);
      tc_fence_after();
      if (TRACE) trace_stamp(tr, t, j, 1);
      uint32_t s[KT];
      if (BOUNDED) {  // first half now, second half lands behind the first half's exponentials
        tmem_ld_32x32(tS, reinterpret_cast<uint32_t(&)[32]>(s[0]));
        tmem_ld_32x32(tS + 32, reinterpret_cast<uint32_t(&)[32]>(s[32]));
        tmem_ld_wait();
        tmem_ld_32x32(tS + 64, reinterpret_cast<uint32_t(&)[32]>(s[64]));
        tmem_ld_32x32(tS + 96, reinterpret_cast<uint32_t(&)[32]>(s[96]));
      } else {
#pragma unroll
        for (int i = 0; i < KT / 32; ++i)
          tmem_ld_32x32(tS + i * 32, reinterpret_cast<uint32_t(&)[32]>(s[i * 32]));
        tmem_ld_wait();
      }
      if (TRACE) trace_stamp(tr, t, j, 6);

      const int valid = p.seq - j * KT;
      if (BOUNDED && valid < KT / 2) {
#pragma unroll
        for (int i = 0; i < KT / 2; ++i)
          if (i >= valid) s[i] = 0xff800000u;
      }
      if (!BOUNDED) {
        if (valid < KT) {
#pragma unroll
          for (int i = 0; i < KT; ++i)
            if (i >= valid) s[i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < KT; i += 4) {
          mx0 = fmax3(mx0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          mx1 = fmax3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        }
        const float mx = fmaxf(mx0, mx1);

        if (j == 0) {
          m = mx;
        } else {
          const bool need = (mx - m) * c > RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, need)) {
            const float f = need ? fast_exp2((m - mx) * c) : 1.0f;
            if (need) m = mx;
            l *= f;
#pragma unroll 1
            for (int i = 0; i < HD / 32; ++i) {
              uint32_t o[32];
              tmem_ld_32x32(tO + i * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
              tmem_st_32x32(tO + i * 32, o);
            }
            tmem_st_wait();
          }
        }
      }

      // ---- exp phase: the two warpgroups alternate (MUFU is the shared, saturated resource) ----
      if (TRACE) trace_stamp(tr, t, j, 2);
      if (TURNS) named_bar_sync(my_turn);
      if (TRACE) trace_stamp(tr, t, j, 3);
      const float2 c2 = make_float2(c, c);
      const float2 nm2 = BOUNDED ? make_float2(-p.bound_log2, -p.bound_log2) : make_float2(-m * c, -m * c);
      float2 lsum = make_float2(0.f, 0.f);
#pragma unroll
      for (int qt = 0; qt < 4; ++qt) {  // quarters of 32 columns
        if (BOUNDED && qt == 2) {
          tmem_ld_wait();  // columns [64, 128) have arrived
          if (valid < KT) {  // last KV tile: columns past the sequence contribute exp2(-inf) = 0
#pragma unroll
            for (int i = KT / 2; i < KT; ++i)
              if (i >= valid) s[i] = 0xff800000u;
          }
        }
#pragma unroll
        for (int i = qt * (KT / 4); i < (qt + 1) * (KT / 4); i += 2) {
          float2 x = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, nm2);
          float2 pv;
          if (((i >> 1) & 3) == 3) {
            pv = exp2_poly2(x);
          } else {
            pv.x = fast_exp2(x.x);
            pv.y = fast_exp2(x.y);
          }
          lsum = __fadd2_rn(lsum, pv);
          s[i >> 1] = pack_bf16x2(pv.x, pv.y);
        }
        if (qt == 1) {  // columns [0, 64) of P are complete: start their store (32 TMEM columns) ...
#pragma unroll
          for (int i = 0; i < KT / 64; ++i)
            tmem_st_32x16(tS + i * 16, reinterpret_cast<const uint32_t(&)[16]>(s[i * 16]));
        }
        if (qt == 2) {  // ... and publish them one quarter later, when the store has drained behind the exponentials
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[2 * t]);
        }
      }
      if (TRACE) trace_stamp(tr, t, j, 4);
      if (TURNS) named_bar_arrive(next_turn);
      l += lsum.x + lsum.y;
#pragma unroll
      for (int i = KT / 64; i < KT / 32; ++i)
        tmem_st_32x16(tS + i * 16, reinterpret_cast<const uint32_t(&)[16]>(s[i * 16]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * t + 1]);
      if (TRACE) trace_stamp(tr, t, j, 5);
    }

    if (TURNS && t == 0) named_bar_sync(BAR_TURN_A);  // consume warpgroup B's last hand-off

    // final: O / l -> bf16 -> global
    mbar_wait(&o_done[t], (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qrow = q0 + t * QT + row;
    const bool valid_row = qrow < p.seq;
    if (p.lse && valid_row)
      p.lse[((long long)b * p.heads + h) * p.seq + qrow] = (BOUNDED ? p.bound_log2 : m * c) + log2f(l);
    __nv_bfloat16* orow =
        p.o + (long long)b * p.o_batch_stride + (long long)qrow * p.o_ld + h * HD;
#pragma unroll 1
    for (int i = 0; i < HD / 32; ++i) {
      uint32_t o[32];
      tmem_ld_32x32(tO + i * 32, o);
      tmem_ld_wait();
      if (valid_row) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
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

// ================================================================================================================
// SPLIT variant of the bounded-score kernel: TWO threads per query row. The fixed-reference softmax has no row maximum, so a
// row's 128 columns need no communication while the tile is processed — each of 4 softmax warpgroups (2 per Q tile) owns a
// 64-column half: tcgen05.ld 64 columns -> 2^(c s - B) -> packed bf16 P into ITS OWN columns of the S region -> one arrive.
// Why: in the one-thread-per-row kernel a tile's softmax is ~780 dependent-ish instructions in ONE warp per scheduler
// (measured: S ready -> P published ~1900 cycles, IPC 0.4), and that latency sits in the serial chain
// QK -> softmax -> PV -> QK of its tile (period ~2830 cycles per KV tile vs 2048 of tensor work). Two warps per scheduler
// working on the same tile halve the chain's softmax leg; the MUFU unit sees the same total work.
//   warps 0..3    as above (TMA, MMA, 2 idle)
//   warps 4..19   softmax: tile t = (warp - 4) / 8, column half hf = ((warp - 4) / 4) & 1, TMEM lane quarter warp & 3
// P layout in TMEM: half hf of tile t lives in columns [t * 128 + hf * 64, + 32) (inside the S columns its own warpgroup
// read), so no warpgroup overwrites S data another one may still be loading.
// ================================================================================================================
constexpr int ATT_SPLIT_THREADS = 32 * (4 + 8 * NQT);
// 640 threads x 96 registers fill the register file; the 64-column softmax fits in 96 (no spills in its loop), so no
// setmaxnreg hand-off is needed here (moving registers away from the MMA-issue warp made IT spill)
constexpr int SPLIT_SOFTMAX_REGS = 96;
constexpr int SPLIT_PRODUCER_REGS = 96;
static_assert((SPLIT_SOFTMAX_REGS - 96) * 512 <= (96 - SPLIT_PRODUCER_REGS) * 128, "register hand-off does not balance");

__global__ void __launch_bounds__(ATT_SPLIT_THREADS, 1)
attention_fwd_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NQT * TILE_BYTES;
  uint8_t* sV = sK + K_STAGES * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + V_STAGES * TILE_BYTES);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* k_full = q_full + 1;             // [K_STAGES]
  uint64_t* k_empty = k_full + K_STAGES;     // [K_STAGES]
  uint64_t* v_full = k_empty + K_STAGES;     // [V_STAGES]
  uint64_t* v_empty = v_full + V_STAGES;     // [V_STAGES]
  uint64_t* s_full = v_empty + V_STAGES;     // [NQT]  MMA -> softmax : S_t(j) ready
  uint64_t* p_full = s_full + NQT;           // [NQT][2]  softmax -> MMA : half hf of P_t(j) written
  uint64_t* o_done = p_full + 2 * NQT;       // [NQT]  MMA -> softmax : PV_t(j) retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + NQT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (NQT * QT);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.seq + KT - 1) / KT;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < K_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < V_STAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < NQT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[2 * t], 4);  // one arrive per warp of the half's warpgroup
      mbar_init(&p_full[2 * t + 1], 4);
      mbar_init(&o_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    if (SPLIT_PRODUCER_REGS < 96) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SPLIT_PRODUCER_REGS));
    if (warp == 0) {
      if (lane == 0) {
        mbar_expect_tx(q_full, NQT * TILE_BYTES);
        for (int t = 0; t < NQT; ++t)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(sQ + t * TILE_BYTES + hf * HALF_BYTES, &tmQ, q_full, h * HD + hf * 64, q0 + t * QT, b);
        for (int j = 0; j < n_kv; ++j) {
          const int ks = j % K_STAGES, vs = j % V_STAGES;
          const uint32_t kph = (j / K_STAGES) & 1, vph = (j / V_STAGES) & 1;
          mbar_wait(&k_empty[ks], kph ^ 1);
          mbar_expect_tx(&k_full[ks], TILE_BYTES);
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(sK + ks * TILE_BYTES + hf * HALF_BYTES, &tmK, &k_full[ks], h * HD + hf * 64, j * KT, b);
          mbar_wait(&v_empty[vs], vph ^ 1);
          mbar_expect_tx(&v_full[vs], TILE_BYTES);
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(sV + vs * TILE_BYTES + hf * HALF_BYTES, &tmV, &v_full[vs], h * HD + hf * 64, j * KT, b);
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(QT, KT, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(QT, HD, false, true);  // V is MN-major
      const uint64_t q_desc = make_sw128_desc(smem_u32(sQ), 16, 1024);
      const uint64_t k_desc = make_sw128_desc(smem_u32(sK), 16, 1024);
      const uint64_t v_desc = make_sw128_desc(smem_u32(sV), HALF_BYTES, 1024);
      auto issue_qk = [&](int t, int st) {
        const uint64_t a0 = q_desc + uint64_t((t * TILE_BYTES) >> 4);
        const uint64_t b0 = k_desc + uint64_t((st * TILE_BYTES) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint32_t off = ((kk >> 2) * HALF_BYTES + (kk & 3) * 32) >> 4;
            umma_ss(tmem_base + t * KT, a0 + off, b0 + off, idesc_qk, kk > 0);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int st, int hf, bool acc) {
        const uint64_t b0 = v_desc + uint64_t((st * TILE_BYTES) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int kl = 0; kl < KT / 32; ++kl) {
            const int kk = hf * (KT / 32) + kl;  // K = 16 step: V rows [16 kk, 16 kk + 16); P half hf starts at column hf * 64
            umma_ts(tmem_base + NQT * KT + t * HD, tmem_base + t * KT + hf * 64 + kl * 8, b0 + uint64_t((kk * 2048) >> 4),
                    idesc_pv, (acc || kk > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one_sync()) tc_commit(bar);
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < NQT; ++t) {
        issue_qk(t, 0);
        commit(&s_full[t]);
      }
      commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % V_STAGES;
        const uint32_t ph = (j / V_STAGES) & 1;
        const bool has_next = (j + 1) < n_kv;
        const int nst = (j + 1) % K_STAGES;
        const uint32_t nph = ((j + 1) / K_STAGES) & 1;
        mbar_wait(&v_full[st], ph);
        if (has_next) mbar_wait(&k_full[nst], nph);
        for (int t = 0; t < NQT; ++t) {
          mbar_wait(&p_full[2 * t], j & 1);
          tc_fence_after();
          issue_pv(t, st, 0, j > 0);
          mbar_wait(&p_full[2 * t + 1], j & 1);
          tc_fence_after();
          issue_pv(t, st, 1, true);
          commit(&o_done[t]);
          if (has_next) {
            issue_qk(t, nst);
            commit(&s_full[t]);
          }
        }
        commit(&v_empty[st]);
        if (has_next) commit(&k_empty[nst]);
      }
    }
  } else {
    if (SPLIT_SOFTMAX_REGS > 96) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SPLIT_SOFTMAX_REGS));
    const int sw = warp - 4;
    const int t = sw >> 3;         // which Q tile
    const int hf = (sw >> 2) & 1;  // which 64-column half of the tile
    const int qd = warp & 3;       // TMEM lane quarter
    const int row = qd * 32 + lane;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + t * KT + hf * 64;  // this thread's 64 S columns; P goes to the first 32
    const uint32_t tO = tmem_base + lane_base + NQT * KT + t * HD + hf * 64;
    const float c = p.scale_log2;
    const float2 c2 = make_float2(c, c);
    const float2 nm2 = make_float2(-p.bound_log2, -p.bound_log2);
    float l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      if (j > 0) mbar_wait(&o_done[t], (j - 1) & 1);  // already complete (in-order tensor pipe): observes the phase
      tc_fence_after();
      uint32_t s[64];
      tmem_ld_32x32(tS, reinterpret_cast<uint32_t(&)[32]>(s[0]));
      tmem_ld_32x32(tS + 32, reinterpret_cast<uint32_t(&)[32]>(s[32]));
      tmem_ld_wait();
      const int valid = p.seq - j * KT - hf * 64;  // columns of this half inside the sequence
      if (valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) s[i] = 0xff800000u;  // -inf -> exp2 = 0
      }
      float2 lsum = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float2 x = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, nm2);
        float2 pv;
        if (((i >> 1) & 3) == 3) {
          pv = exp2_poly2(x);
        } else {
          pv.x = fast_exp2(x.x);
          pv.y = fast_exp2(x.y);
        }
        lsum = __fadd2_rn(lsum, pv);
        s[i >> 1] = pack_bf16x2(pv.x, pv.y);
      }
      l += lsum.x + lsum.y;
      tmem_st_32x32(tS, reinterpret_cast<const uint32_t(&)[32]>(s[0]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * t + hf]);
    }

    // row sum of the two halves through shared memory (the Q tiles are dead once the last QK^T has retired)
    mbar_wait(&o_done[t], (n_kv - 1) & 1);
    tc_fence_after();
    float* sL = reinterpret_cast<float*>(sQ + t * TILE_BYTES);  // [2][128]
    sL[hf * 128 + row] = l;
    asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");   // the 256 threads of this tile
    const float l_tot = sL[row] + sL[128 + row];
    const float inv_l = 1.0f / l_tot;
    const int qrow = q0 + t * QT + row;
    const bool valid_row = qrow < p.seq;
    if (p.lse && valid_row && hf == 0) p.lse[((long long)b * p.heads + h) * p.seq + qrow] = p.bound_log2 + log2f(l_tot);
    __nv_bfloat16* orow = p.o + (long long)b * p.o_batch_stride + (long long)qrow * p.o_ld + h * HD + hf * 64;
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      uint32_t o[32];
      tmem_ld_32x32(tO + i * 32, o);
      tmem_ld_wait();
      if (valid_row) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
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

// ================================================================================================================
// PAIR variant (bounded scores): a CTA PAIR (cluster of 2, tcgen05 cta_group::2) shares every K / V tile.
// Why: in the one-CTA kernels shared memory carries, per KV tile and Q tile, 64 KiB of operand reads for S = Q K^T (SS form:
// Q AND K from shared memory), 32 KiB for O += P V, plus the TMA fill of K and V — 125 B/clk/SM against the 128 B/clk an SM
// has (the same wall the 1-CTA GEMM hit). Here each CTA still owns two 128-row Q tiles (512 query rows per
// pair), but stages only HALF of each K tile (64 of its 128 kv rows) and HALF of each V tile (64 of the 128 head-dim
// columns); the M = 256 MMAs read the two halves from both CTAs: per SM 160 KiB instead of 256 KiB per KV tile.
//   S_t  = [Q_t(cta 0); Q_t(cta 1)] K_j^T   : M 256, N 128 (kv), B operand K-major, N split across the pair
//   O_t += [P_t(cta 0); P_t(cta 1)] V_j     : M 256, N 128 (d),  A from each CTA's TMEM, B MN-major, N split across the pair
// The leader CTA's MMA warp issues for both; TMA loads of both CTAs complete on the leader's barriers; tcgen05.commit is
// multicast to both CTAs' "S ready" / "PV retired" / "slot free" barriers; the peer's softmax warps publish P with remote
// arrives on the leader's barriers. Softmax: fixed reference (no row max), one thread per row, as in the bounded kernel.
// ================================================================================================================
constexpr int PK_STAGES = 4;                      // K half-tiles in flight
constexpr int PV_STAGES = 3;                      // V half-tiles in flight
constexpr int KHALF_BYTES = 64 * 64 * 2;          // one SW128 [64 kv rows x 64 d] block of the CTA's K half
constexpr int KSTAGE_BYTES = 2 * KHALF_BYTES;     // 64 kv rows x 128 d = 16 KiB
constexpr int VSTAGE_BYTES = 128 * 64 * 2;        // 128 kv rows x 64 d = 16 KiB
constexpr size_t ATT_PAIR_SMEM_BYTES = 1024 + size_t(NQT) * TILE_BYTES + size_t(PK_STAGES) * KSTAGE_BYTES +
                                       size_t(PV_STAGES) * VSTAGE_BYTES + 512;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ATT_THREADS, 1)
attention_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NQT * TILE_BYTES;
  uint8_t* sV = sK + PK_STAGES * KSTAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + PV_STAGES * VSTAGE_BYTES);
  uint64_t* q_full = bars;                    // [1]          leader copy used (both CTAs' Q tiles)
  uint64_t* k_full = q_full + 1;              // [PK_STAGES]  leader copy used
  uint64_t* k_empty = k_full + PK_STAGES;     // [PK_STAGES]  per CTA (multicast commit)
  uint64_t* v_full = k_empty + PK_STAGES;     // [PV_STAGES]  leader copy used
  uint64_t* v_empty = v_full + PV_STAGES;     // [PV_STAGES]  per CTA
  uint64_t* s_full = v_empty + PV_STAGES;     // [NQT]        per CTA: S_t(j) ready
  uint64_t* p_full = s_full + NQT;            // [NQT][2]     leader copy used: 8 arrives (4 warps x 2 CTAs)
  uint64_t* o_done = p_full + 2 * NQT;        // [NQT]        per CTA: PV_t(j) retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + NQT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int q0 = blockIdx.x * (NQT * QT);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.seq + KT - 1) / KT;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < PK_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < PV_STAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < NQT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[2 * t], 8);
      mbar_init(&p_full[2 * t + 1], 8);
      mbar_init(&o_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // peer barriers initialised, both TMEM allocations done
  __syncthreads();     // (ordering is the cluster barrier's; this one is what compute-sanitizer racecheck models)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
    if (warp == 0) {
      // ------------------------------- TMA producer (both CTAs) -------------------------------
      if (lane == 0) {
        const uint32_t lq = mapa_u32(q_full, 0);
        if (leader) mbar_expect_tx(q_full, 2 * NQT * TILE_BYTES);
        for (int t = 0; t < NQT; ++t)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d_2cta(sQ + t * TILE_BYTES + hf * HALF_BYTES, &tmQ, lq, h * HD + hf * 64, q0 + t * QT, b);
        for (int j = 0; j < n_kv; ++j) {
          const int ks = j % PK_STAGES, vs = j % PV_STAGES;
          const uint32_t kph = (j / PK_STAGES) & 1, vph = (j / PV_STAGES) & 1;
          mbar_wait(&k_empty[ks], kph ^ 1);
          const uint32_t lk = mapa_u32(&k_full[ks], 0);
          if (leader) mbar_expect_tx(&k_full[ks], 2 * KSTAGE_BYTES);
          for (int hf = 0; hf < 2; ++hf)  // this CTA's 64 kv rows of the tile, two 64-column d blocks
            tma_load_3d_2cta(sK + ks * KSTAGE_BYTES + hf * KHALF_BYTES, &tmK, lk, h * HD + hf * 64, j * KT + int(rank) * 64, b);
          mbar_wait(&v_empty[vs], vph ^ 1);
          const uint32_t lv = mapa_u32(&v_full[vs], 0);
          if (leader) mbar_expect_tx(&v_full[vs], 2 * VSTAGE_BYTES);
          // this CTA's 64 head-dim columns of all 128 kv rows
          tma_load_3d_2cta(sV + vs * VSTAGE_BYTES, &tmV, lv, h * HD + int(rank) * 64, j * KT, b);
        }
      }
    } else if (warp == 1 && leader) {
      // ------------------------------- MMA issuer (leader CTA) --------------------------------
      constexpr uint32_t idesc_qk = make_idesc_bf16(2 * QT, KT, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(2 * QT, HD, false, true);  // V is MN-major
      const uint64_t q_desc = make_sw128_desc(smem_u32(sQ), 16, 1024);
      const uint64_t k_desc = make_sw128_desc(smem_u32(sK), 16, 1024);
      const uint64_t v_desc = make_sw128_desc(smem_u32(sV), VSTAGE_BYTES, 1024);
      auto issue_qk = [&](int t, int st) {
        const uint64_t a0 = q_desc + uint64_t((t * TILE_BYTES) >> 4);
        const uint64_t b0 = k_desc + uint64_t((st * KSTAGE_BYTES) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint32_t aoff = ((kk >> 2) * HALF_BYTES + (kk & 3) * 32) >> 4;
            const uint32_t boff = ((kk >> 2) * KHALF_BYTES + (kk & 3) * 32) >> 4;
            umma_ss_2cta(tmem_base + t * KT, a0 + aoff, b0 + boff, idesc_qk, kk > 0);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int st, int hf, bool acc) {
        const uint64_t b0 = v_desc + uint64_t((st * VSTAGE_BYTES) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = hf * (KT / 32); kk < (hf + 1) * (KT / 32); ++kk)
            umma_ts_2cta(tmem_base + NQT * KT + t * HD, tmem_base + t * KT + kk * 8, b0 + uint64_t((kk * 2048) >> 4), idesc_pv,
                         (acc || kk > 0) ? 1u : 0u);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one_sync()) tc_commit_2cta_mcast(bar);
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < NQT; ++t) {
        issue_qk(t, 0);
        commit(&s_full[t]);
      }
      commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % PV_STAGES;
        const uint32_t ph = (j / PV_STAGES) & 1;
        const bool has_next = (j + 1) < n_kv;
        const int nst = (j + 1) % PK_STAGES;
        const uint32_t nph = ((j + 1) / PK_STAGES) & 1;
        mbar_wait(&v_full[st], ph);
        if (has_next) mbar_wait(&k_full[nst], nph);
        for (int t = 0; t < NQT; ++t) {
          mbar_wait(&p_full[2 * t], j & 1);
          tc_fence_after();
          issue_pv(t, st, 0, j > 0);
          mbar_wait(&p_full[2 * t + 1], j & 1);
          tc_fence_after();
          issue_pv(t, st, 1, true);
          commit(&o_done[t]);
          if (has_next) {
            issue_qk(t, nst);
            commit(&s_full[t]);
          }
        }
        commit(&v_empty[st]);
        if (has_next) commit(&k_empty[nst]);
      }
    }
  } else {
    // ------------------------------- softmax warpgroups (both CTAs) -------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SOFTMAX_REGS));
    const int t = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + t * KT;
    const uint32_t tO = tmem_base + lane_base + NQT * KT + t * HD;
    const float c = p.scale_log2;
    const float2 c2 = make_float2(c, c);
    const float2 nm2 = make_float2(-p.bound_log2, -p.bound_log2);
    const uint32_t p_bar0 = mapa_u32(&p_full[2 * t], 0), p_bar1 = mapa_u32(&p_full[2 * t + 1], 0);
    float l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      if (j > 0) mbar_wait(&o_done[t], (j - 1) & 1);  // already complete (in-order tensor pipe): observes the phase
      tc_fence_after();
      uint32_t s[KT];
      tmem_ld_32x32(tS, reinterpret_cast<uint32_t(&)[32]>(s[0]));
      tmem_ld_32x32(tS + 32, reinterpret_cast<uint32_t(&)[32]>(s[32]));
      tmem_ld_wait();
      tmem_ld_32x32(tS + 64, reinterpret_cast<uint32_t(&)[32]>(s[64]));
      tmem_ld_32x32(tS + 96, reinterpret_cast<uint32_t(&)[32]>(s[96]));
      const int valid = p.seq - j * KT;
      if (valid < KT / 2) {
#pragma unroll
        for (int i = 0; i < KT / 2; ++i)
          if (i >= valid) s[i] = 0xff800000u;
      }
      float2 lsum = make_float2(0.f, 0.f);
#pragma unroll
      for (int qt = 0; qt < 4; ++qt) {
        if (qt == 2) {
          tmem_ld_wait();
          if (valid < KT) {
#pragma unroll
            for (int i = KT / 2; i < KT; ++i)
              if (i >= valid) s[i] = 0xff800000u;
          }
        }
#pragma unroll
        for (int i = qt * (KT / 4); i < (qt + 1) * (KT / 4); i += 2) {
          float2 x = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, nm2);
          float2 pv;
          if (((i >> 1) & 3) == 3) {
            pv = exp2_poly2(x);
          } else {
            pv.x = fast_exp2(x.x);
            pv.y = fast_exp2(x.y);
          }
          lsum = __fadd2_rn(lsum, pv);
          s[i >> 1] = pack_bf16x2(pv.x, pv.y);
        }
        if (qt == 1) {
#pragma unroll
          for (int i = 0; i < KT / 64; ++i)
            tmem_st_32x16(tS + i * 16, reinterpret_cast<const uint32_t(&)[16]>(s[i * 16]));
        }
        if (qt == 2) {
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(p_bar0);
        }
      }
      l += lsum.x + lsum.y;
#pragma unroll
      for (int i = KT / 64; i < KT / 32; ++i)
        tmem_st_32x16(tS + i * 16, reinterpret_cast<const uint32_t(&)[16]>(s[i * 16]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(p_bar1);
    }

    mbar_wait(&o_done[t], (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qrow = q0 + t * QT + row;
    const bool valid_row = qrow < p.seq;
    if (p.lse && valid_row) p.lse[((long long)b * p.heads + h) * p.seq + qrow] = p.bound_log2 + log2f(l);
    __nv_bfloat16* orow = p.o + (long long)b * p.o_batch_stride + (long long)qrow * p.o_ld + h * HD;
#pragma unroll 1
    for (int i = 0; i < HD / 32; ++i) {
      uint32_t o[32];
      tmem_ld_32x32(tO + i * 32, o);
      tmem_ld_wait();
      if (valid_row) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + i * 32 + g * 8) = w;
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory and TMEM: nobody leaves before everybody is done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

int attention_read_trace(long long* out, int n) {
  long long host[2 * 64 * 8];
  AFB_CHECK_CUDA(cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(host)));
  for (int i = 0; i < n && i < 2 * 64 * 8; ++i) out[i] = host[i];
  return AFB_OK;
}

int attention_launch(const afb_attn_desc* d, cudaStream_t stream) {
  AFB_REQUIRE(d != nullptr, "attention: null descriptor");
  AFB_REQUIRE(d->q && d->k && d->v && d->o, "attention: null operand pointer");
  AFB_REQUIRE(d->batch >= 1 && d->seq >= 1 && d->heads >= 1, "attention: empty problem");
  AFB_REQUIRE(d->o_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d->o) & 15) == 0,
              "attention: o must be 16-byte aligned with ld %% 8 == 0");
  // Developer variants (env AFB_ATTN_DEBUG_MODE): see the selection below. 13 = the CTA-pair kernel needs its own K / V boxes.
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("AFB_ATTN_DEBUG_MODE");
    dbg = e ? atoi(e) : 0;
  }
  const float bound_log2 = d->score_bound > 0.f ? d->score_bound * 1.4426950408889634f : 0.f;
  const bool bounded = dbg != 10 && bound_log2 > 0.f && bound_log2 <= 60.0f;
  const bool pair = bounded && dbg == 13;
  CUtensorMap tm[3];
  const void* ptr[3] = {d->q, d->k, d->v};
  const int64_t ld[3] = {d->q_ld, d->k_ld, d->v_ld};
  const int64_t bs[3] = {d->q_batch_stride, d->k_batch_stride, d->v_batch_stride};
  for (int i = 0; i < 3; ++i) {
    const uint64_t dims[3] = {uint64_t(d->heads) * HD, uint64_t(d->seq), uint64_t(d->batch)};
    const uint64_t bstride = d->batch > 1 ? uint64_t(bs[i]) : uint64_t(d->seq) * uint64_t(ld[i]);
    const uint64_t strides[2] = {uint64_t(ld[i]) * 2, bstride * 2};
    // pair kernel: K box = 64 kv rows (this CTA's half of the tile), V box = 128 kv rows x 64 d (this CTA's d half)
    const uint32_t box[3] = {64, uint32_t(pair && i == 1 ? 64 : 128), 1};
    int rc = make_tmap_bf16(&tm[i], ptr[i], 3, dims, strides, box);
    if (rc != AFB_OK) return rc;
  }
  AttnParams p{};
  p.seq = d->seq;
  p.heads = d->heads;
  p.batch = d->batch;
  const float scale = d->scale > 0.f ? d->scale : 1.0f / sqrtf(float(HD));
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = static_cast<__nv_bfloat16*>(d->o);
  p.o_ld = d->o_ld;
  p.o_batch_stride = d->o_batch_stride;
  p.lse = d->lse;
  // score_bound > 0: the caller guarantees |scale * q.k| <= score_bound -> fixed-reference softmax (no row max, no O
  // rescale). Only used while 2 B stays far inside the fp32 / bf16 exponent range; otherwise the running-max kernel runs.
  // Developer variants (env AFB_ATTN_DEBUG_MODE): 7 = clock-stamp trace, 8 = trace without warpgroup turn-taking,
  // 9 = no turn-taking, no trace, 10 = ignore score_bound (always the running-max kernel), 11 / 12 = bounded scores with ONE
  // thread per row, without / with turn-taking, 13 = bounded scores on a CTA pair sharing K / V, 14 = two threads per row
  // on one CTA. Anything else is the product path, which carries no instrumentation.
  p.bound_log2 = bounded ? bound_log2 : 0.f;
  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, AttnParams);
  KernelFn fn;
  int threads = ATT_THREADS;
  size_t smem_bytes = ATT_SMEM_BYTES;
  dim3 grid((d->seq + NQT * QT - 1) / (NQT * QT), d->heads, d->batch);
  if (pair) {
    fn = attention_fwd_pair_kernel;
    smem_bytes = ATT_PAIR_SMEM_BYTES;
    grid.x = (grid.x + 1) & ~1u;  // whole CTA pairs; a CTA past the sequence still stages its K / V halves for its partner
  } else if (bounded && dbg != 11 && dbg != 12) {  // two threads per row
    fn = attention_fwd_split_kernel;
    threads = ATT_SPLIT_THREADS;
  } else if (bounded) {
    fn = dbg == 11 ? attention_fwd_kernel<false, false, true> : attention_fwd_kernel<false, true, true>;
  } else {
    fn = dbg == 7   ? attention_fwd_kernel<true, true, false>
         : dbg == 8 ? attention_fwd_kernel<true, false, false>
         : dbg == 9 ? attention_fwd_kernel<false, false, false>
                    : attention_fwd_kernel<false, true, false>;
  }
  static KernelFn attr_set_for[3] = {nullptr, nullptr, nullptr};
  const int slot = pair ? 2 : (bounded ? 1 : 0);
  if (attr_set_for[slot] != fn) {
    AFB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
    attr_set_for[slot] = fn;
  }
  fn<<<grid, threads, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
  AFB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return AFB_OK;
}

}  // namespace afb
