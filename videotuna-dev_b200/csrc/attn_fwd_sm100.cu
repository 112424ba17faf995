// attn_fwd_sm100.cu — dense (non-causal) flash-attention forward for sm_100a.
//
// Replaces the attention core of
//   lvdm CrossAttention.forward          (videotuna/models/lvdm/modules/attention.py:126-149)
//   hunyuan attention(mode="flash"|"torch") (videotuna/models/hunyuan/hyvideo_t2v/modules/attenion.py:101-120)
//   wan flash_attention                  (videotuna/models/wan/wan/modules/attention.py:96-127)
//
// One CTA owns NQ (1 or 2) query tiles of 128 rows for one (problem, head) and streams 128-key tiles:
//   warp 8*NQ     : TMA producer (Q once; K/V ring, 128B-swizzled boxes of 128 rows x 64 elems)
//   warp 8*NQ + 1 : tcgen05.mma issuer (single thread).  S_t = Q_t K^T (SS), O_t += P_t V (A = P from TMEM)
//   warps [0,8*NQ): softmax, TWO threads per query row (two warpgroups per tile, each owning 64 of the 128 key
//                   columns): tcgen05.ld S -> row max exchanged through shared memory -> online softmax (exp2 on packed
//                   f32x2 math, part of it on the FMA pipe, lazy rescale) -> bf16 P written back over S -> mbarrier.
//                   Halving the per-thread work halves the softmax latency, which — not MUFU or MMA throughput —
//                   bounded the kernel: the per-tile chain softmax -> PV -> QK -> softmax was 1850 + 1024 + ~200 cycles.
// With NQ == 2 the two tiles ping-pong: while one tile's warpgroups do softmax the tensor core runs the other tile's
// P V and Q K^T, so MMA and MUFU work overlap.
// TMEM columns: S_t at 128*t (P_t aliases its first 64 columns), O_t at 128*NQ + D*t.
#include <cstdlib>
#include <cuda_bf16.h>
#include <math_constants.h>

#include "attn_common.h"
#include "capi_util.h"
#include "sm100_ptx.cuh"

#ifndef VT_FWD_EMU
#define VT_FWD_EMU 2
#endif

namespace vt {
namespace {

template <int D, int NQ>
struct FwdCfg {
  static_assert(D == 64 || D == 128, "head dim must be 64 or 128");
  static constexpr int KCH = D / 64;               // 128-byte chunks along the head dim
  static constexpr int CHUNK = 128 * 128;          // bytes: 128 rows x 128 B (one swizzled box)
  static constexpr int TILE = CHUNK * KCH;         // bytes of a 128 x D bf16 tile
  static constexpr int KS = 2, VS = 2;             // K / V ring depth
  static constexpr int EMU = VT_FWD_EMU;           // of every 8 exponential pairs, how many run on the FMA pipe
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NQ * TILE;
  static constexpr int OFF_V = OFF_K + KS * TILE;
  static constexpr int OFF_MX = OFF_V + VS * TILE;    // row-max / row-sum exchange: float [2 parity][NQ][2 halves][128]
  static constexpr int OFF_BAR = OFF_MX + 2 * NQ * 2 * 128 * 4;
  static constexpr int NBAR = 1 + 2 * KS + 2 * VS + 3 * NQ;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int BYTES = OFF_TMEM + 16 + 1024;  // + alignment slack
  static constexpr int TMEM_USED = NQ * (128 + D);
  static constexpr int TMEM_COLS = TMEM_USED <= 256 ? 256 : 512;
  // 8 softmax warps per tile + producer + issuer. 576 threads (NQ == 2) launch with 112 registers each, which is what
  // the softmax threads need, so no setmaxnreg re-balancing is involved.
  static constexpr int THREADS = (NQ * 8 + 2) * 32;
};

enum : uint32_t {
  TAG_Q_FULL = 0x100, TAG_K_FULL, TAG_K_EMPTY, TAG_V_FULL, TAG_V_EMPTY, TAG_S_FULL, TAG_P_FULL, TAG_O_FULL
};

template <int D, int NQ>
__global__ void __launch_bounds__(FwdCfg<D, NQ>::THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  using C = FwdCfg<D, NQ>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  // ---- which problem / rows ------------------------------------------------------------------
  const int prob = blockIdx.z, h = blockIdx.y;
  int q_base = 0, q_len = p.seq.Lq, bq = prob;
  int k_base = 0, k_len = p.seq.Lk, bk = prob;
  if (p.seq.cu_q != nullptr) {
    q_base = p.seq.cu_q[prob];
    q_len = p.seq.cu_q[prob + 1] - q_base;
    bq = 0;
  }
  if (p.seq.cu_k != nullptr) {
    k_base = p.seq.cu_k[prob];
    k_len = p.seq.cu_k[prob + 1] - k_base;
    bk = 0;
  } else if (p.seq.seqlens_k != nullptr) {
    k_len = min(max(p.seq.seqlens_k[prob], 0), p.seq.Lk);
  }
  const int q0 = blockIdx.x * (NQ * 128);
  if (q0 >= q_len) return;  // CTA-uniform
  const int n_kv = (k_len + 127) >> 7;

  if (n_kv == 0) {  // no keys: softmax over the empty set -> zeros, lse = -inf (matches flash-attn)
    for (int r = threadIdx.x; r < NQ * 128; r += blockDim.x) {
      const int row = q0 + r;
      if (row < q_len) {
        __nv_bfloat16* optr = p.o + bq * p.o_sb + static_cast<int64_t>(q_base + row) * p.o_sl + h * p.o_sh;
        for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(optr + c) = make_uint4(0, 0, 0, 0);
        p.lse[bq * p.lse_sb + h * p.lse_sh + q_base + row] = -CUDART_INF_F;
      }
    }
    return;
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + C::KS;
  uint64_t* v_full = k_empty + C::KS;
  uint64_t* v_empty = v_full + C::VS;
  uint64_t* s_full = v_empty + C::VS;
  uint64_t* p_full = s_full + NQ;
  uint64_t* o_full = p_full + NQ;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);

  constexpr int PROD_WARP = NQ * 8, MMA_WARP = NQ * 8 + 1;

  if (warp == PROD_WARP && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == MMA_WARP && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < C::KS; ++i) { mbar_init(k_full + i, 1); mbar_init(k_empty + i, 1); }
    for (int i = 0; i < C::VS; ++i) { mbar_init(v_full + i, 1); mbar_init(v_empty + i, 1); }
    for (int i = 0; i < NQ; ++i) { mbar_init(s_full + i, 1); mbar_init(p_full + i, 256); mbar_init(o_full + i, 1); }
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform: stays in a uniform register

  // NQ == 2: 384 threads x 168 regs at launch; the service warpgroup (warps 8..11) gives registers to the two
  // softmax warpgroups (setmaxnreg is warpgroup-wide, hence the idle warps 10,11 take the first branch too).
  if (warp >= PROD_WARP) {
    if (warp == PROD_WARP && lane == 0) {
    // ================================ TMA producer ============================================
    {
      mbar_arrive_expect_tx(q_full, NQ * C::TILE);
#pragma unroll
      for (int t = 0; t < NQ; ++t)
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
          tma_load_4d(smem + C::OFF_Q + t * C::TILE + c * C::CHUNK, &tm_q, q_full, c * 64, q_base + q0 + t * 128, h,
                      bq);
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j % C::KS, vs = j % C::VS;
        mbar_wait(k_empty + ks, ((j / C::KS) & 1) ^ 1, TAG_K_EMPTY);
        mbar_arrive_expect_tx(k_full + ks, C::TILE);
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
          tma_load_4d(smem + C::OFF_K + ks * C::TILE + c * C::CHUNK, &tm_k, k_full + ks, c * 64, k_base + j * 128, h,
                      bk);
        mbar_wait(v_empty + vs, ((j / C::VS) & 1) ^ 1, TAG_V_EMPTY);
        mbar_arrive_expect_tx(v_full + vs, C::TILE);
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
          tma_load_4d(smem + C::OFF_V + vs * C::TILE + c * C::CHUNK, &tm_v, v_full + vs, c * 64, k_base + j * 128, h,
                      bk);
      }
    }
    } else if (warp == MMA_WARP && elect_one()) {
    // ================================ MMA issuer ==============================================
    {
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 128, 0, 0);  // A=Q K-major, B=K K-major
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, D, 0, 1);    // A=P (TMEM), B=V MN-major
      // shared-memory addresses in 16-byte units (the descriptor's address field)
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t q_smem = sb16 + (C::OFF_Q >> 4), k_smem = sb16 + (C::OFF_K >> 4), v_smem = sb16 + (C::OFF_V >> 4);
      constexpr uint32_t CH16 = C::CHUNK >> 4, TILE16 = C::TILE >> 4;

      auto issue_qk = [&](int t, int ks) {
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t a = umma_desc_sw128_a16(q_smem + t * TILE16 + c * CH16 + kk * 2, 16, 1024);
            const uint64_t b = umma_desc_sw128_a16(k_smem + ks * TILE16 + c * CH16 + kk * 2, 16, 1024);
            umma_ss(tmem_base + t * 128, a, b, IDESC_QK, (c | kk) != 0);
          }
      };
      auto issue_pv = [&](int t, int vs, bool acc) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 keys per MMA
          // V tile: [128 keys][64 d] x KCH boxes. MN-major B: LBO = next 64-d box, SBO = next 8 keys.
          const uint64_t b = umma_desc_sw128_a16(v_smem + vs * TILE16 + kk * 128, C::CHUNK, 1024);
          umma_ts(tmem_base + NQ * 128 + t * D, tmem_base + t * 128 + kk * 8, b, IDESC_PV, acc || kk != 0);
        }
      };

      mbar_wait(q_full, 0, TAG_Q_FULL);
      mbar_wait(k_full + 0, 0, TAG_K_FULL);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < NQ; ++t) {
        issue_qk(t, 0);
        tc_commit(s_full + t);
      }
      tc_commit(k_empty + 0);

      for (int j = 0; j < n_kv; ++j) {
        const int vs = j % C::VS;
        mbar_wait(v_full + vs, (j / C::VS) & 1, TAG_V_FULL);
        const bool has_next = (j + 1 < n_kv);
        const int ksn = (j + 1) % C::KS;
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
          trace_mark(p.trace, 1, j, t * 3);
          mbar_wait(p_full + t, j & 1, TAG_P_FULL);
          tc_fence_after();
          trace_mark(p.trace, 1, j, t * 3 + 1);
          issue_pv(t, vs, j > 0);
          if (t == NQ - 1) tc_commit(v_empty + vs);
          if (has_next) {
            if (t == 0) {
              mbar_wait(k_full + ksn, ((j + 1) / C::KS) & 1, TAG_K_FULL);
              tc_fence_after();
            }
            issue_qk(t, ksn);
            tc_commit(s_full + t);
            if (t == NQ - 1) tc_commit(k_empty + ksn);
          } else {
            tc_commit(o_full + t);
          }
          trace_mark(p.trace, 1, j, t * 3 + 2);
        }
      }
    }
    }
  } else {
    // ================================ softmax warpgroups ======================================
    const int t = warp >> 3;         // query tile
    const int hf = (warp >> 2) & 1;  // which 64 of the 128 key columns (and which half of the O columns)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + t * 128 + hf * 64;   // this thread's 64 S columns
    const uint32_t p_addr = tmem_base + lane_addr + t * 128 + hf * 32;   // its 32 packed-bf16 P columns
    constexpr int OH = D / 2;                                            // O columns per thread
    const uint32_t o_addr = tmem_base + lane_addr + NQ * 128 + t * D + hf * OH;
    const float sl2 = p.scale_log2;
    float* mx = reinterpret_cast<float*>(smem + C::OFF_MX);  // [parity][t][half][row]
    const uint32_t pair_bar = 1 + t;                         // named barrier of this tile's 256 threads

    const bool tr = row == 0 && hf == 0;
    const int trole = t == 0 ? 0 : 2;
    float m = -CUDART_INF_F, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      if (tr) trace_mark(p.trace, trole, j, 0);
      mbar_wait(s_full + t, j & 1, TAG_S_FULL);
      tc_fence_after();
      if (tr) trace_mark(p.trace, trole, j, 1);
      uint32_t su[64];
      tmem_ld_x32(s_addr + 0, su + 0);
      tmem_ld_x32(s_addr + 32, su + 32);
      tc_wait_ld();
      float* s = reinterpret_cast<float*>(su);
      if (j == n_kv - 1) {
        const int valid = k_len - j * 128 - hf * 64;
        if (valid < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= valid) s[c] = -CUDART_INF_F;
        }
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int c = 4; c < 64; c += 4) {
        mx0 = fmaxf(mx0, s[c]);
        mx1 = fmaxf(mx1, s[c + 1]);
        mx2 = fmaxf(mx2, s[c + 2]);
        mx3 = fmaxf(mx3, s[c + 3]);
      }
      // exchange the half-row maxima. The barrier also orders every S read of this tile (tcgen05.wait::ld above)
      // before any P write below: the partner's P columns overlay S columns this thread has just read.
      float* slot = mx + (((j & 1) * NQ + t) * 2) * 128;
      slot[hf * 128 + row] = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      named_bar_sync(pair_bar, 256);
      const float m_new = fmaxf(m, fmaxf(slot[row], slot[128 + row]));
      if (tr) trace_mark(p.trace, trole, j, 2);
      if (j == 0) {
        m = m_new;
      } else {
        // Lazy rescale: keep the stale max while it is within 2^8 of the new one. Both threads of a row see the same
        // m and m_new, so they take the same decision and each rescales its half of the O columns.
        const bool need = (m_new - m) * sl2 > 8.f;
        if (__any_sync(0xffffffffu, need)) {
          const float f = need ? ex2_approx((m - m_new) * sl2) : 1.f;
          if (need) m = m_new;
          l *= f;
          // rare path: 8 columns at a time keeps its register footprint (and with it the hot path's) small
#pragma unroll 1
          for (int c0 = 0; c0 < OH; c0 += 8) {
            uint32_t ou[8];
            tmem_ld_x8(o_addr + c0, ou);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 8; ++c) ou[c] = __float_as_uint(__uint_as_float(ou[c]) * f);
            tmem_st_x8(o_addr + c0, ou);
          }
        }
      }
      if (tr) trace_mark(p.trace, trole, j, 3);
      // p = 2^(s * scale_log2 - m * scale_log2) on packed pairs: one FFMA2 for the argument, one FADD2 for the row sum.
      // EMU of every 8 pairs take the FMA-pipe polynomial instead of MUFU.EX2 (see ex2_poly2).
      const float msc = m * sl2;
      const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-msc, -msc);
      float2 lacc = make_float2(0.f, 0.f);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float2 x = __ffma2_rn(make_float2(s[c0 + c], s[c0 + c + 1]), sc2, nm2);
          float2 pv;
          if ((((c0 + c) >> 1) & 7) < C::EMU) {
            pv = ex2_poly2(x);
          } else {
            pv.x = ex2_approx(x.x);
            pv.y = ex2_approx(x.y);
          }
          lacc = __fadd2_rn(lacc, pv);
          pk[c >> 1] = pack_bf16x2(pv.x, pv.y);
        }
        tmem_st_x16(p_addr + (c0 >> 1), pk);
      }
      l += lacc.x + lacc.y;
      if (tr) trace_mark(p.trace, trole, j, 4);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full + t);
      if (tr) trace_mark(p.trace, trole, j, 5);
    }

    // ---- epilogue: total row sum, O / l -> bf16 -> global; lse ---------------------------------
    {
      float* slot = mx + (((n_kv & 1) * NQ + t) * 2) * 128;  // the parity the last iteration did not use
      slot[hf * 128 + row] = l;
      named_bar_sync(pair_bar, 256);
      l = slot[row] + slot[128 + row];
    }
    mbar_wait(o_full + t, 0, TAG_O_FULL);
    tc_fence_after();
    const float inv = 1.f / l;
    const int row_g = q0 + t * 128 + row;
    const bool valid_row = row_g < q_len;
    __nv_bfloat16* optr = p.o + bq * p.o_sb + static_cast<int64_t>(q_base + row_g) * p.o_sl + h * p.o_sh + hf * OH;
#pragma unroll
    for (int c0 = 0; c0 < OH; c0 += 32) {
      uint32_t ou[32];
      tmem_ld_x32(o_addr + c0, ou);
      tc_wait_ld();
      if (valid_row) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(ou[c + 0]) * inv, __uint_as_float(ou[c + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(ou[c + 2]) * inv, __uint_as_float(ou[c + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(ou[c + 4]) * inv, __uint_as_float(ou[c + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(ou[c + 6]) * inv, __uint_as_float(ou[c + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + c0 + c) = w;
        }
      }
    }
    if (valid_row && hf == 0) p.lse[bq * p.lse_sb + h * p.lse_sh + q_base + row_g] = m * p.scale + __logf(l);
  }

  // ---- teardown ------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// =====================================================================================================================
// attn_fwd_db_kernel<D>: ONE 128-row query tile per CTA with S double-buffered in TMEM.
//
// In the ping-pong kernel above bf16 P overlays its own S tile, so S_t(j+1) = Q_t K(j+1)^T cannot be issued before
// O_t += P_t(j) V(j) has consumed P: each tile runs the serial chain softmax -> PV -> QK -> softmax (measured
// ~1850 + 1024 + ~200 cycles per 128 keys for 2 x 1024 cycles of MMA, i.e. 66 % tensor-active). Here TMEM holds
// S[0] | S[1] | O (384 of 512 columns): QK(j+2) is issued right behind PV(j) into the buffer PV(j) frees, so S(j+1) is
// always ready when softmax(j) ends and the softmax warps never wait for the tensor pipe; the kernel is bound by the
// softmax throughput of one tile (two threads per row, packed f32x2 math, part of exp2 on the FMA pipe).
// Warps 0-7 softmax (half = key columns [64*hf, 64*hf+64)), warp 8 TMA producer (K ring of 3, V ring of 2), warp 9
// issuer. O is rescaled by the softmax threads only when the row max grew by > 2^8 (lazy), after waiting for PV(j-1).
enum : uint32_t { TAG_PV_DONE = 0x110, TAG_FWD_ALIGN };

#ifdef VT_EXPERIMENTS  // one-tile kernel with S double-buffered in TMEM: an earlier variant kept for A/B builds
template <int D>
struct FwdDbCfg {
  static constexpr int KCH = D / 64;
  static constexpr int CHUNK = 128 * 128;
  static constexpr int TILE = CHUNK * KCH;
  static constexpr int KS = 3, VS = 2;
  static constexpr int EMU = VT_FWD_EMU;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + TILE;
  static constexpr int OFF_V = OFF_K + KS * TILE;
  static constexpr int OFF_MX = OFF_V + VS * TILE;  // float [2 parity][2 halves][128]
  static constexpr int OFF_BAR = OFF_MX + 2 * 2 * 128 * 4;
  static constexpr int NBAR = 1 + 2 * KS + 2 * VS + 2 + 3;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int BYTES = OFF_TMEM + 16;
  static constexpr int THREADS = 10 * 32;
  static constexpr uint32_t T_S = 0, T_O = 256;
};


template <int D>
__global__ void __launch_bounds__(FwdDbCfg<D>::THREADS, 1)
attn_fwd_db_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  using C = FwdDbCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) watchdog_trap(TAG_FWD_ALIGN);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  const int prob = blockIdx.z, h = blockIdx.y;
  int q_base = 0, q_len = p.seq.Lq, bq = prob;
  int k_base = 0, k_len = p.seq.Lk, bk = prob;
  if (p.seq.cu_q != nullptr) {
    q_base = p.seq.cu_q[prob];
    q_len = p.seq.cu_q[prob + 1] - q_base;
    bq = 0;
  }
  if (p.seq.cu_k != nullptr) {
    k_base = p.seq.cu_k[prob];
    k_len = p.seq.cu_k[prob + 1] - k_base;
    bk = 0;
  } else if (p.seq.seqlens_k != nullptr) {
    k_len = min(max(p.seq.seqlens_k[prob], 0), p.seq.Lk);
  }
  const int q0 = blockIdx.x * 128;
  if (q0 >= q_len) return;  // CTA-uniform
  const int n_kv = (k_len + 127) >> 7;

  if (n_kv == 0) {  // no keys: softmax over the empty set -> zeros, lse = -inf (matches flash-attn)
    for (int r = threadIdx.x; r < 128; r += blockDim.x) {
      const int row = q0 + r;
      if (row < q_len) {
        __nv_bfloat16* optr = p.o + bq * p.o_sb + static_cast<int64_t>(q_base + row) * p.o_sl + h * p.o_sh;
        for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(optr + c) = make_uint4(0, 0, 0, 0);
        p.lse[bq * p.lse_sb + h * p.lse_sh + q_base + row] = -CUDART_INF_F;
      }
    }
    return;
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + C::KS;
  uint64_t* v_full = k_empty + C::KS;
  uint64_t* v_empty = v_full + C::VS;
  uint64_t* s_full = v_empty + C::VS;  // [2]
  uint64_t* p_full = s_full + 2;
  uint64_t* pv_done = p_full + 1;
  uint64_t* o_full = pv_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);
  constexpr int PROD_WARP = 8, MMA_WARP = 9;

  if (warp == PROD_WARP && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == MMA_WARP && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < C::KS; ++i) { mbar_init(k_full + i, 1); mbar_init(k_empty + i, 1); }
    for (int i = 0; i < C::VS; ++i) { mbar_init(v_full + i, 1); mbar_init(v_empty + i, 1); }
    mbar_init(s_full + 0, 1);
    mbar_init(s_full + 1, 1);
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == PROD_WARP) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::TILE);
#pragma unroll
      for (int c = 0; c < C::KCH; ++c)
        tma_load_4d(smem + C::OFF_Q + c * C::CHUNK, &tm_q, q_full, c * 64, q_base + q0, h, bq);
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j % C::KS, vs = j % C::VS;
        mbar_wait(k_empty + ks, ((j / C::KS) & 1) ^ 1, TAG_K_EMPTY);
        mbar_arrive_expect_tx(k_full + ks, C::TILE);
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
          tma_load_4d(smem + C::OFF_K + ks * C::TILE + c * C::CHUNK, &tm_k, k_full + ks, c * 64, k_base + j * 128, h, bk);
        mbar_wait(v_empty + vs, ((j / C::VS) & 1) ^ 1, TAG_V_EMPTY);
        mbar_arrive_expect_tx(v_full + vs, C::TILE);
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
          tma_load_4d(smem + C::OFF_V + vs * C::TILE + c * C::CHUNK, &tm_v, v_full + vs, c * 64, k_base + j * 128, h, bk);
      }
    }
  } else if (warp == MMA_WARP) {
    if (elect_one()) {
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, D, 0, 1);
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t q_smem = sb16 + (C::OFF_Q >> 4), k_smem = sb16 + (C::OFF_K >> 4), v_smem = sb16 + (C::OFF_V >> 4);
      constexpr uint32_t CH16 = C::CHUNK >> 4, TILE16 = C::TILE >> 4;
      auto issue_qk = [&](int jj) {  // S[jj & 1] = Q K(jj)^T
        const int ks = jj % C::KS;
        mbar_wait(k_full + ks, (jj / C::KS) & 1, TAG_K_FULL);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem_base + C::T_S + (jj & 1) * 128, umma_desc_sw128_a16(q_smem + c * CH16 + kk * 2, 16, 1024),
                    umma_desc_sw128_a16(k_smem + ks * TILE16 + c * CH16 + kk * 2, 16, 1024), IDESC_QK, (c | kk) != 0);
        tc_commit(s_full + (jj & 1));
        tc_commit(k_empty + ks);
      };
      mbar_wait(q_full, 0, TAG_Q_FULL);
      issue_qk(0);
      if (n_kv > 1) issue_qk(1);
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j % C::VS;
        mbar_wait(v_full + vs, (j / C::VS) & 1, TAG_V_FULL);
        trace_mark(p.trace, 1, j, 0);
        mbar_wait(p_full, j & 1, TAG_P_FULL);
        tc_fence_after();
        trace_mark(p.trace, 1, j, 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // 16 keys per MMA; P(j) sits in the first 64 columns of S[j & 1]
          umma_ts(tmem_base + C::T_O, tmem_base + C::T_S + (j & 1) * 128 + kk * 8,
                  umma_desc_sw128_a16(v_smem + vs * TILE16 + kk * 128, C::CHUNK, 1024), IDESC_PV, (j > 0) || kk != 0);
        tc_commit(v_empty + vs);
        tc_commit(pv_done);
        if (j + 2 < n_kv) issue_qk(j + 2);
        if (j == n_kv - 1) tc_commit(o_full);
        trace_mark(p.trace, 1, j, 2);
      }
    }
  } else {
    // ================================ softmax: two threads per query row ==========================
    const int hf = warp >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    constexpr int OH = D / 2;
    const uint32_t o_addr = tmem_base + lane_addr + C::T_O + hf * OH;
    const float sl2 = p.scale_log2;
    float* mx = reinterpret_cast<float*>(smem + C::OFF_MX);
    const bool tr = row == 0 && hf == 0;
    float m = -CUDART_INF_F, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t s_addr = tmem_base + lane_addr + C::T_S + (j & 1) * 128 + hf * 64;
      const uint32_t p_addr = tmem_base + lane_addr + C::T_S + (j & 1) * 128 + hf * 32;
      if (tr) trace_mark(p.trace, 0, j, 0);
      mbar_wait(s_full + (j & 1), (j >> 1) & 1, TAG_S_FULL);
      tc_fence_after();
      if (tr) trace_mark(p.trace, 0, j, 1);
      uint32_t su[64];
      tmem_ld_x32(s_addr + 0, su + 0);
      tmem_ld_x32(s_addr + 32, su + 32);
      tc_wait_ld();
      float* s = reinterpret_cast<float*>(su);
      if (j == n_kv - 1) {
        const int valid = k_len - j * 128 - hf * 64;
        if (valid < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= valid) s[c] = -CUDART_INF_F;
        }
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int c = 4; c < 64; c += 4) {
        mx0 = fmaxf(mx0, s[c]);
        mx1 = fmaxf(mx1, s[c + 1]);
        mx2 = fmaxf(mx2, s[c + 2]);
        mx3 = fmaxf(mx3, s[c + 3]);
      }
      // exchange the half-row maxima; the barrier also orders this tile's S reads before the partner's P writes
      float* slot = mx + (j & 1) * 256;
      slot[hf * 128 + row] = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      named_bar_sync(1, 256);
      const float m_new = fmaxf(m, fmaxf(slot[row], slot[128 + row]));
      if (tr) trace_mark(p.trace, 0, j, 2);
      if (j == 0) {
        m = m_new;
      } else {
        const bool need = (m_new - m) * sl2 > 8.f;  // lazy rescale, same decision in both threads of a row
        if (__any_sync(0xffffffffu, need)) {
          const float f = need ? ex2_approx((m - m_new) * sl2) : 1.f;
          if (need) m = m_new;
          l *= f;
          mbar_wait(pv_done, (j - 1) & 1, TAG_PV_DONE);  // O += P(j-1) V(j-1) has landed
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < OH; c0 += 8) {
            uint32_t ou[8];
            tmem_ld_x8(o_addr + c0, ou);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 8; ++c) ou[c] = __float_as_uint(__uint_as_float(ou[c]) * f);
            tmem_st_x8(o_addr + c0, ou);
          }
        }
      }
      if (tr) trace_mark(p.trace, 0, j, 3);
      const float msc = m * sl2;
      const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-msc, -msc);
      float2 lacc = make_float2(0.f, 0.f);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float2 x = __ffma2_rn(make_float2(s[c0 + c], s[c0 + c + 1]), sc2, nm2);
          float2 pv;
          if ((((c0 + c) >> 1) & 7) < C::EMU) {
            pv = ex2_poly2(x);
          } else {
            pv.x = ex2_approx(x.x);
            pv.y = ex2_approx(x.y);
          }
          lacc = __fadd2_rn(lacc, pv);
          pk[c >> 1] = pack_bf16x2(pv.x, pv.y);
        }
        tmem_st_x16(p_addr + (c0 >> 1), pk);
      }
      l += lacc.x + lacc.y;
      if (tr) trace_mark(p.trace, 0, j, 4);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
      if (tr) trace_mark(p.trace, 0, j, 5);
    }
    {
      float* slot = mx + (n_kv & 1) * 256;
      slot[hf * 128 + row] = l;
      named_bar_sync(1, 256);
      l = slot[row] + slot[128 + row];
    }
    mbar_wait(o_full, 0, TAG_O_FULL);
    tc_fence_after();
    const float inv = 1.f / l;
    const int row_g = q0 + row;
    const bool valid_row = row_g < q_len;
    __nv_bfloat16* optr = p.o + bq * p.o_sb + static_cast<int64_t>(q_base + row_g) * p.o_sl + h * p.o_sh + hf * OH;
#pragma unroll
    for (int c0 = 0; c0 < OH; c0 += 32) {
      uint32_t ou[32];
      tmem_ld_x32(o_addr + c0, ou);
      tc_wait_ld();
      if (valid_row) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(ou[c + 0]) * inv, __uint_as_float(ou[c + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(ou[c + 2]) * inv, __uint_as_float(ou[c + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(ou[c + 4]) * inv, __uint_as_float(ou[c + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(ou[c + 6]) * inv, __uint_as_float(ou[c + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + c0 + c) = w;
        }
      }
    }
    if (valid_row && hf == 0) p.lse[bq * p.lse_sb + h * p.lse_sh + q_base + row_g] = m * p.scale + __logf(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// attn_fwd_rot_kernel<D = 64>: TWO 128-row query tiles per CTA sharing THREE rotating S buffers in TMEM.
//
// At head dim 64 a tile's MMAs (Q K^T 256 + P V 256 cycles per 128 keys) are much shorter than its softmax (~1500
// cycles), so the ping-pong kernel's per-tile chain softmax -> PV -> QK -> softmax leaves the tensor pipe 34 % busy and
// both tiles' softmax warps idle while they wait for their next S. Head dim 64 leaves TMEM room for a third S buffer
// (3 x 128 + 2 x 64 = 512 columns): job n = 2 j + t (key tile j, query tile t) uses buffer n % 3, and Q K^T of job
// n + 3 is issued right behind P V of job n into the buffer that P V frees. S of a tile's next key block is therefore
// complete long before its softmax warps finish the current one: they never wait, and the kernel runs at the SM's
// softmax throughput (MUFU + FMA-pipe exponentials of both tiles overlapped) instead of at the chain latency.
// Warps 0-7 / 8-15: softmax of tile 0 / 1 (two threads per row), warp 16 TMA producer, warp 17 MMA issuer.
#endif  // VT_EXPERIMENTS

template <int D>
struct FwdRotCfg {
  static_assert(D == 64, "three S buffers + two O tiles fit TMEM only at head dim 64");
  static constexpr int CHUNK = 128 * 128;
  static constexpr int TILE = CHUNK;  // 128 x 64 bf16
  static constexpr int KS = 3, VS = 3;  // ring depths = jobs-per-group / 2, so ring slots are compile-time in the issuer
#ifdef VT_FWD_EMU64
  static constexpr int EMU = VT_FWD_EMU64;
#else
  static constexpr int EMU = 2;  // K3 fwd: 0/8 .., 1/8 908, 2/8 905, 3/8 886 TFLOP/s — flat: the loop is issue-bound, not MUFU-bound
#endif
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + 2 * TILE;
  static constexpr int OFF_V = OFF_K + KS * TILE;
  static constexpr int OFF_MX = OFF_V + VS * TILE;  // float [2 parity][2 tiles][2 halves][128]
  static constexpr int OFF_BAR = OFF_MX + 2 * 2 * 2 * 128 * 4;
  static constexpr int NBAR = 1 + 2 * KS + 2 * VS + 3 + 2 + 2 + 2;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int BYTES = OFF_TMEM + 16;
  static constexpr int THREADS = 18 * 32;
  static constexpr uint32_t T_S = 0, T_O = 384;
};

template <int D>
__global__ void __launch_bounds__(FwdRotCfg<D>::THREADS, 1)
attn_fwd_rot_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  using C = FwdRotCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) watchdog_trap(TAG_FWD_ALIGN);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  const int prob = blockIdx.z, h = blockIdx.y;
  int q_base = 0, q_len = p.seq.Lq, bq = prob;
  int k_base = 0, k_len = p.seq.Lk, bk = prob;
  if (p.seq.cu_q != nullptr) {
    q_base = p.seq.cu_q[prob];
    q_len = p.seq.cu_q[prob + 1] - q_base;
    bq = 0;
  }
  if (p.seq.cu_k != nullptr) {
    k_base = p.seq.cu_k[prob];
    k_len = p.seq.cu_k[prob + 1] - k_base;
    bk = 0;
  } else if (p.seq.seqlens_k != nullptr) {
    k_len = min(max(p.seq.seqlens_k[prob], 0), p.seq.Lk);
  }
  const int q0 = blockIdx.x * 256;
  if (q0 >= q_len) return;  // CTA-uniform
  const int n_kv = (k_len + 127) >> 7;

  if (n_kv == 0) {  // no keys: softmax over the empty set -> zeros, lse = -inf (matches flash-attn)
    for (int r = threadIdx.x; r < 256; r += blockDim.x) {
      const int row = q0 + r;
      if (row < q_len) {
        __nv_bfloat16* optr = p.o + bq * p.o_sb + static_cast<int64_t>(q_base + row) * p.o_sl + h * p.o_sh;
        for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(optr + c) = make_uint4(0, 0, 0, 0);
        p.lse[bq * p.lse_sb + h * p.lse_sh + q_base + row] = -CUDART_INF_F;
      }
    }
    return;
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + C::KS;
  uint64_t* v_full = k_empty + C::KS;
  uint64_t* v_empty = v_full + C::VS;
  uint64_t* s_full = v_empty + C::VS;  // [3] one per S buffer
  uint64_t* p_full = s_full + 3;       // [2] one per query tile
  uint64_t* pv_done = p_full + 2;      // [2]
  uint64_t* o_full = pv_done + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);
  constexpr int PROD_WARP = 16, MMA_WARP = 17;

  if (warp == PROD_WARP && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == MMA_WARP && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < C::KS; ++i) { mbar_init(k_full + i, 1); mbar_init(k_empty + i, 1); }
    for (int i = 0; i < C::VS; ++i) { mbar_init(v_full + i, 1); mbar_init(v_empty + i, 1); }
    for (int i = 0; i < 3; ++i) mbar_init(s_full + i, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(p_full + i, 256); mbar_init(pv_done + i, 1); mbar_init(o_full + i, 1); }
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int njobs = 2 * n_kv;

  if (warp == PROD_WARP) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * C::TILE);
      tma_load_4d(smem + C::OFF_Q, &tm_q, q_full, 0, q_base + q0, h, bq);
      tma_load_4d(smem + C::OFF_Q + C::TILE, &tm_q, q_full, 0, q_base + q0 + 128, h, bq);
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j % C::KS, vs = j % C::VS;
        mbar_wait(k_empty + ks, ((j / C::KS) & 1) ^ 1, TAG_K_EMPTY);
        mbar_arrive_expect_tx(k_full + ks, C::TILE);
        tma_load_4d(smem + C::OFF_K + ks * C::TILE, &tm_k, k_full + ks, 0, k_base + j * 128, h, bk);
        mbar_wait(v_empty + vs, ((j / C::VS) & 1) ^ 1, TAG_V_EMPTY);
        mbar_arrive_expect_tx(v_full + vs, C::TILE);
        tma_load_4d(smem + C::OFF_V + vs * C::TILE, &tm_v, v_full + vs, 0, k_base + j * 128, h, bk);
      }
    }
  } else if (warp == MMA_WARP) {
    if (elect_one()) {
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, D, 0, 1);
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t q_smem = sb16 + (C::OFF_Q >> 4), k_smem = sb16 + (C::OFF_K >> 4), v_smem = sb16 + (C::OFF_V >> 4);
      constexpr uint32_t TILE16 = C::TILE >> 4;
      // Jobs n = 2 j + t run in groups of six (three key tiles x two query tiles): inside the unrolled group every
      // S buffer (n % 3), ring slot (j % 3) and query tile is a compile-time constant and only the phase parities depend
      // on the group index g. The issuer is one thread on a scheduler it shares with four busy softmax warps, so its
      // instruction count per MMA is what limits the kernel at head dim 64.
      // S[u % 3] = Q_t K_j^T for the job at position u (0..8: positions 6..8 are the next group's first jobs)
      auto issue_qk = [&](int u, uint32_t g1) {
        const int t = u & 1, jj = u >> 1, ks = jj % 3, buf = u % 3;
        const uint32_t par = (jj >= 3 ? g1 ^ 1u : g1);  // (j / 3) & 1
        if (t == 0) {
          mbar_wait(k_full + ks, par, TAG_K_FULL);
          tc_fence_after();
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem_base + C::T_S + buf * 128, umma_desc_sw128_a16(q_smem + t * TILE16 + kk * 2, 16, 1024),
                  umma_desc_sw128_a16(k_smem + ks * TILE16 + kk * 2, 16, 1024), IDESC_QK, kk != 0);
        tc_commit(s_full + buf);
        if (t == 1) tc_commit(k_empty + ks);
      };
      mbar_wait(q_full, 0, TAG_Q_FULL);
      issue_qk(0, 0);
      issue_qk(1, 0);
      if (njobs > 2) issue_qk(2, 0);
      for (int n0 = 0; n0 < njobs; n0 += 6) {
        const uint32_t g1 = static_cast<uint32_t>(n0 / 6) & 1u;
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          const int n = n0 + u;
          if (n < njobs) {
            const int t = u & 1, jj = u >> 1, buf = u % 3;
            const uint32_t jpar = g1 ^ static_cast<uint32_t>(jj & 1);  // j & 1 with j = 3 g + jj
            if (t == 0) mbar_wait(v_full + jj, g1, TAG_V_FULL);
            trace_mark(p.trace, 1, n >> 1, t * 3);
            mbar_wait(p_full + t, jpar, TAG_P_FULL);
            tc_fence_after();
            trace_mark(p.trace, 1, n >> 1, t * 3 + 1);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // 16 keys per MMA; P sits in the first 64 columns of its S buffer
              umma_ts(tmem_base + C::T_O + t * D, tmem_base + C::T_S + buf * 128 + kk * 8,
                      umma_desc_sw128_a16(v_smem + jj * TILE16 + kk * 128, C::CHUNK, 1024), IDESC_PV, (n > 1) || kk != 0);
            tc_commit(pv_done + t);
            if (t == 1) tc_commit(v_empty + jj);
            if (n + 3 < njobs) issue_qk(u + 3, g1);  // into the buffer this P V frees
            if (n + 2 >= njobs) tc_commit(o_full + t);
            trace_mark(p.trace, 1, n >> 1, t * 3 + 2);
          }
        }
      }
    }
  } else {
    // ================================ softmax: two threads per query row ==========================
    const int t = warp >> 3;
    const int hf = (warp >> 2) & 1;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    constexpr int OH = D / 2;
    const uint32_t o_addr = tmem_base + lane_addr + C::T_O + t * D + hf * OH;
    const float sl2 = p.scale_log2;
    float* mx = reinterpret_cast<float*>(smem + C::OFF_MX);
    const uint32_t pair_bar = 1 + t;
    const bool tr = row == 0 && hf == 0;
    const int trole = t == 0 ? 0 : 2;
    float m = -CUDART_INF_F, l = 0.f;
    int buf = t;          // (2 j + t) % 3
    uint32_t use = 0;     // (2 j + t) / 3
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t s_addr = tmem_base + lane_addr + C::T_S + buf * 128 + hf * 64;
      const uint32_t p_addr = tmem_base + lane_addr + C::T_S + buf * 128 + hf * 32;
      if (tr) trace_mark(p.trace, trole, j, 0);
      mbar_wait(s_full + buf, use & 1, TAG_S_FULL);
      tc_fence_after();
      if (tr) trace_mark(p.trace, trole, j, 1);
      uint32_t su[64];
      tmem_ld_x32(s_addr + 0, su + 0);
      tmem_ld_x32(s_addr + 32, su + 32);
      tc_wait_ld();
      float* s = reinterpret_cast<float*>(su);
      if (j == n_kv - 1) {
        const int valid = k_len - j * 128 - hf * 64;
        if (valid < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= valid) s[c] = -CUDART_INF_F;
        }
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int c = 4; c < 64; c += 4) {
        mx0 = fmaxf(mx0, s[c]);
        mx1 = fmaxf(mx1, s[c + 1]);
        mx2 = fmaxf(mx2, s[c + 2]);
        mx3 = fmaxf(mx3, s[c + 3]);
      }
      // exchange the half-row maxima; the barrier also orders this tile's S reads before the partner's P writes
      float* slot = mx + (((j & 1) * 2 + t) * 2) * 128;
      slot[hf * 128 + row] = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      named_bar_sync(pair_bar, 256);
      const float m_new = fmaxf(m, fmaxf(slot[row], slot[128 + row]));
      if (tr) trace_mark(p.trace, trole, j, 2);
      if (j == 0) {
        m = m_new;
      } else {
        const bool need = (m_new - m) * sl2 > 8.f;  // lazy rescale, same decision in both threads of a row
        if (__any_sync(0xffffffffu, need)) {
          const float f = need ? ex2_approx((m - m_new) * sl2) : 1.f;
          if (need) m = m_new;
          l *= f;
          // S of this job was issued behind P V of job n - 3, not n - 2: wait for this tile's previous P V explicitly
          mbar_wait(pv_done + t, (j - 1) & 1, TAG_PV_DONE);
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < OH; c0 += 8) {
            uint32_t ou[8];
            tmem_ld_x8(o_addr + c0, ou);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 8; ++c) ou[c] = __float_as_uint(__uint_as_float(ou[c]) * f);
            tmem_st_x8(o_addr + c0, ou);
          }
        }
      }
      if (tr) trace_mark(p.trace, trole, j, 3);
      const float msc = m * sl2;
      const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-msc, -msc);
      float2 lacc = make_float2(0.f, 0.f);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float2 x = __ffma2_rn(make_float2(s[c0 + c], s[c0 + c + 1]), sc2, nm2);
          float2 pv;
          if ((((c0 + c) >> 1) & 7) < C::EMU) {
            pv = ex2_poly2(x);
          } else {
            pv.x = ex2_approx(x.x);
            pv.y = ex2_approx(x.y);
          }
          lacc = __fadd2_rn(lacc, pv);
          pk[c >> 1] = pack_bf16x2(pv.x, pv.y);
        }
        tmem_st_x16(p_addr + (c0 >> 1), pk);
      }
      l += lacc.x + lacc.y;
      if (tr) trace_mark(p.trace, trole, j, 4);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full + t);
      if (tr) trace_mark(p.trace, trole, j, 5);
      // next job of this tile: n + 2
      buf += 2;
      if (buf >= 3) { buf -= 3; ++use; }
    }
    {
      float* slot = mx + (((n_kv & 1) * 2 + t) * 2) * 128;
      slot[hf * 128 + row] = l;
      named_bar_sync(pair_bar, 256);
      l = slot[row] + slot[128 + row];
    }
    mbar_wait(o_full + t, 0, TAG_O_FULL);
    tc_fence_after();
    const float inv = 1.f / l;
    const int row_g = q0 + t * 128 + row;
    const bool valid_row = row_g < q_len;
    __nv_bfloat16* optr = p.o + bq * p.o_sb + static_cast<int64_t>(q_base + row_g) * p.o_sl + h * p.o_sh + hf * OH;
    {
      uint32_t ou[OH];
      tmem_ld_x32(o_addr, ou);
      tc_wait_ld();
      if (valid_row) {
#pragma unroll
        for (int c = 0; c < OH; c += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(ou[c + 0]) * inv, __uint_as_float(ou[c + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(ou[c + 2]) * inv, __uint_as_float(ou[c + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(ou[c + 4]) * inv, __uint_as_float(ou[c + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(ou[c + 6]) * inv, __uint_as_float(ou[c + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + c) = w;
        }
      }
    }
    if (valid_row && hf == 0) p.lse[bq * p.lse_sb + h * p.lse_sh + q_base + row_g] = m * p.scale + __logf(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int D>
cudaError_t launch_rot(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v, const AttnFwdParams& p,
                       cudaStream_t stream) {
  using C = FwdRotCfg<D>;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_rot_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::BYTES);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((p.seq.Lq + 255) / 256, p.seq.H, p.seq.nprob);
  attn_fwd_rot_kernel<D><<<grid, C::THREADS, C::BYTES, stream>>>(tm_q, tm_k, tm_v, p);
  return cudaGetLastError();
}

#ifdef VT_EXPERIMENTS
template <int D>
cudaError_t launch_db(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v, const AttnFwdParams& p,
                      cudaStream_t stream) {
  using C = FwdDbCfg<D>;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_db_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::BYTES);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((p.seq.Lq + 127) / 128, p.seq.H, p.seq.nprob);
  attn_fwd_db_kernel<D><<<grid, C::THREADS, C::BYTES, stream>>>(tm_q, tm_k, tm_v, p);
  return cudaGetLastError();
}

#endif  // VT_EXPERIMENTS

template <int D, int NQ>
cudaError_t launch_one(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                       const AttnFwdParams& p, cudaStream_t stream) {
  using C = FwdCfg<D, NQ>;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<D, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::BYTES);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((p.seq.Lq + NQ * 128 - 1) / (NQ * 128), p.seq.H, p.seq.nprob);
  attn_fwd_kernel<D, NQ><<<grid, C::THREADS, C::BYTES, stream>>>(tm_q, tm_k, tm_v, p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attn_fwd(int D, const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                            const AttnFwdParams& p, int q_tiles_hint, cudaStream_t stream) {
#ifdef VT_EXPERIMENTS
  // A/B builds only (tools/build_variant.sh -DVT_EXPERIMENTS): VT_FWD_KERNEL=db selects the one-tile kernel with S
  // double-buffered in TMEM, VT_FWD_KERNEL=pp the two-tile ping-pong kernel. Neither has the fused exchange epilogue, so
  // a scatter launch (p.sc_n > 0) never takes them.
  static const char fwd_sel = [] { const char* e = getenv("VT_FWD_KERNEL"); return e != nullptr ? e[0] : '\0'; }();
  if (p.sc_n == 0) {
    const bool single_x = (q_tiles_hint == 1) || (q_tiles_hint == 0 && p.seq.Lq <= 128);
    if (fwd_sel == 'd' && q_tiles_hint == 0) {
      if (D == 128) return launch_db<128>(tm_q, tm_k, tm_v, p, stream);
      if (D == 64) return launch_db<64>(tm_q, tm_k, tm_v, p, stream);
    }
    if (D == 128 && (fwd_sel == 'p' || q_tiles_hint != 0))
      return single_x ? launch_one<128, 1>(tm_q, tm_k, tm_v, p, stream) : launch_one<128, 2>(tm_q, tm_k, tm_v, p, stream);
    if (D == 64 && fwd_sel == 'p' && !single_x) return launch_one<64, 2>(tm_q, tm_k, tm_v, p, stream);
  }
#endif
  (void)q_tiles_hint;
  // Head dim 128: one tile per CTA, three rotating S buffers, two alternating softmax sets (attn_fwd_alt_sm100.cu) — the
  // only kernel with the fused Ulysses exchange epilogue.
  if (D == 128) return launch_attn_fwd_alt(tm_q, tm_k, tm_v, p, stream);
  if (D == 64) {
    if (p.sc_n > 0) return cudaErrorNotSupported;  // capi.cu rejects this earlier (VT_ERR_UNSUPPORTED)
    // One query tile per CTA when the (max) query length fits a single 128-row tile; otherwise two tiles sharing three
    // rotating S buffers (K3: see DESIGN.md §4.1).
    if (p.seq.Lq <= 128) return launch_one<64, 1>(tm_q, tm_k, tm_v, p, stream);
    return launch_rot<64>(tm_q, tm_k, tm_v, p, stream);
  }
  return cudaErrorInvalidValue;
}

cudaError_t attn_fwd_set_debug_ptr(unsigned int* ptr) { return cudaMemcpyToSymbol(g_vt_dbg, &ptr, sizeof(ptr)); }

}  // namespace vt
