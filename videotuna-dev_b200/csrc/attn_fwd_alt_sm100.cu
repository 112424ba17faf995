// attn_fwd_alt_sm100.cu — flash-attention forward at head dim 128: ONE 128-row query tile per CTA, THREE rotating S
// buffers in TMEM and TWO softmax warp sets that take alternate key blocks.
//
// Why: at head dim 128 TMEM (512 columns) holds either two query tiles with one S buffer each (the ping-pong kernel in
// attn_fwd_sm100.cu) or one tile with three S buffers (3 x 128 + O 128). With one S buffer per tile, bf16 P overlays its
// own S, so every tile runs the serial chain softmax -> P V -> Q K^T -> softmax (measured ~1530 + 320 + 1024 + 150
// cycles per 128 keys) and two chains keep the tensor pipe 68 % busy. Here Q K^T of block j + 3 is issued right behind
// P V of block j into the buffer that P V frees, so S of the next two blocks is always waiting in TMEM, and the long
// exponential phases of consecutive blocks overlap because they run on different warp sets:
//   set A: key blocks 0, 2, 4, ...      set B: key blocks 1, 3, 5, ...
// The online softmax stays sequential only through the running row maximum m: block j reads m(j-1) that the other set
// published right after ITS max step (before its exponentials), decides m(j) (lazy: raised only when the block maximum
// exceeds it by 2^8), publishes it, and then computes P(j) = 2^(s - m(j)) at leisure. Each set keeps its own partial
// row sum and rescales it whenever m moved since its previous block; the set that raises m also rescales O (after
// waiting for P V of the previous block). The epilogue aligns and adds the two partial sums.
//   warps 0 / 3: tcgen05.mma issuers for even / odd key blocks (one thread each)
//   warp 1: TMA producer for Q and the K ring (2 slots), TMEM allocation
//   warp 2: TMA producer for the V ring (3 slots)   warps 4-11 / 12-19: softmax sets A / B
// K and V are loaded by different threads because their slots are released at different times (K right after
// Q K^T, three blocks ahead of its next use; V only after P V): one in-order producer would hold every K load behind
// the preceding V load and expose the L2 latency of K on the critical path (measured: 2900 instead of 1024 cycles per
// block).
// TMEM columns: S[b] at 128 b (P overlays its first 64 columns), O at 384.
#include <cstdlib>
#include <cuda_bf16.h>
#include <math_constants.h>

#include "attn_common.h"
#include "capi_util.h"
#include "sm100_ptx.cuh"

// Of every 8 exponential pairs, how many run on the FMA pipe (ex2_poly2). With both softmax sets busy the loop is
// bound by instruction issue, not MUFU: K1 forward 1/8: 1255, 2/8: 1194, 3/8: 1174 TFLOP/s.
#ifndef VT_FWD_EMU
#define VT_FWD_EMU 1
#endif

namespace vt {
namespace {

struct AltCfg {
  static constexpr int D = 128;
  static constexpr int CHUNK = 128 * 128;  // bytes: 128 rows x 128 B (one swizzled box)
  static constexpr int TILE = 2 * CHUNK;   // 128 x 128 bf16
  static constexpr int KS = 2, VS = 3;
  static constexpr int EMU = VT_FWD_EMU;   // of every 8 exponential pairs, how many run on the FMA pipe
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + TILE;
  static constexpr int OFF_V = OFF_K + KS * TILE;
  static constexpr int OFF_MX = OFF_V + VS * TILE;       // half-row max exchange: float [2 sets][2 parity][2 halves][128]
  static constexpr int OFF_MRUN = OFF_MX + 2 * 2 * 2 * 128 * 4;  // published running max: float [2 parity][128]
  static constexpr int OFF_LSUM = OFF_MRUN + 2 * 128 * 4;        // partial row sums for the epilogue: float [4][128]
  static constexpr int OFF_BAR = OFF_LSUM + 4 * 128 * 4;
  static constexpr int NBAR = 1 + 2 * KS + 2 * VS + 3 + 2 + 2 + 1 + 1 + 2;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int BYTES = OFF_TMEM + 16;
  static constexpr int THREADS = 20 * 32;
  static constexpr uint32_t T_S = 0, T_O = 384;
};

enum : uint32_t {
  AT_Q_FULL = 0x300, AT_K_FULL, AT_K_EMPTY, AT_V_FULL, AT_V_EMPTY, AT_S_FULL, AT_P_FULL, AT_PV_DONE, AT_O_FULL, AT_MPUB,
  AT_ALIGN, AT_TOK
};

__global__ void __launch_bounds__(AltCfg::THREADS, 1)
attn_fwd_alt_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  using C = AltCfg;
  constexpr int D = C::D;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) watchdog_trap(AT_ALIGN);
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
  uint64_t* s_full = v_empty + C::VS;  // [3] one per S buffer
  uint64_t* p_full = s_full + 3;       // [2] one per softmax set
  uint64_t* mpub = p_full + 2;         // [2] running max of block j published (index j & 1)
  uint64_t* pv_done = mpub + 2;        // P V of block j complete (phase j)
  uint64_t* o_full = pv_done + 1;
  uint64_t* tok = o_full + 1;          // [2] tok[i]: the other issuer has queued its P V, issuer i may queue the next
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);
  // Service warps come FIRST: the warp scheduler favours older (lower-numbered) warps, and the single-thread MMA issuer
  // must never queue behind the four softmax warps of its scheduler — as warp 17 its scalar bookkeeping ran at ~12
  // cycles per instruction and the issue loop (1420 cycles per block) was slower than the MMAs it feeds (1024).
  constexpr int MMA_WARP = 0, PROD_WARP = 1, VPROD_WARP = 2, MMA_WARP1 = 3, SOFTMAX_WARP0 = 4;

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
    for (int i = 0; i < 2; ++i) { mbar_init(p_full + i, 256); mbar_init(mpub + i, 256); }
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    mbar_init(tok + 0, 1);
    mbar_init(tok + 1, 1);
    fence_mbar_init();
  }
  if (warp == PROD_WARP) {
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
      for (int c = 0; c < 2; ++c) tma_load_4d(smem + C::OFF_Q + c * C::CHUNK, &tm_q, q_full, c * 64, q_base + q0, h, bq);
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j % C::KS;
        mbar_wait(k_empty + ks, ((j / C::KS) & 1) ^ 1, AT_K_EMPTY);
        mbar_arrive_expect_tx(k_full + ks, C::TILE);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_4d(smem + C::OFF_K + ks * C::TILE + c * C::CHUNK, &tm_k, k_full + ks, c * 64, k_base + j * 128, h, bk);
      }
    }
  } else if (warp == VPROD_WARP) {
    if (lane == 0) {
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j % C::VS;
        mbar_wait(v_empty + vs, ((j / C::VS) & 1) ^ 1, AT_V_EMPTY);
        mbar_arrive_expect_tx(v_full + vs, C::TILE);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_4d(smem + C::OFF_V + vs * C::TILE + c * C::CHUNK, &tm_v, v_full + vs, c * 64, k_base + j * 128, h, bk);
      }
    }
  } else if (warp == MMA_WARP || warp == MMA_WARP1) {
    if (elect_one()) {
      // TWO issuer threads take alternate key blocks. A tcgen05.mma / commit sequence holds its issuing thread until
      // the tensor pipe has nearly drained it, and the scalar tail of the loop (barrier polls, counters, the back-edge:
      // ~400 cycles measured for a lone thread next to four busy warps) would then leave the pipe idle before the next
      // block. With two issuers one thread's tail overlaps the other's MMAs. P V of consecutive blocks must still
      // execute in block order (pv_done counts one phase per block; P V of block 0 initialises O). A plain arrive does
      // not order two threads' MMAs — it can overtake the arriving thread's own tcgen05.mma — so the hand-off is a
      // tcgen05.commit: issuer i may queue P V of block j + 1 once P V of block j has completed. By then Q K^T of block
      // j + 3 (queued behind P V(j) by the same thread, 512 cycles long) is executing, so the pipe does not drain.
      const int me = (warp == MMA_WARP) ? 0 : 1;
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 128, 0, 0);  // A = Q K-major, B = K K-major
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, D, 0, 1);    // A = P (TMEM), B = V MN-major
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t q_smem = sb16 + (C::OFF_Q >> 4), k_smem = sb16 + (C::OFF_K >> 4), v_smem = sb16 + (C::OFF_V >> 4);
      constexpr uint32_t CH16 = C::CHUNK >> 4, TILE16 = C::TILE >> 4;
      // S[buf] = Q K^T for the key tile in K slot ks; kpar = (block / 2) & 1. Ring slots and S buffers are runtime values:
      // their descriptors are then computed per block on the uniform datapath. (Unrolling a six-block group makes every
      // descriptor a loop invariant, which the compiler hoists into ~200 vector registers and feeds back through R2UR per
      // MMA: 3x slower issue, measured.)
      auto issue_qk = [&](int buf, int ks, uint32_t kpar) {
        mbar_wait(k_full + ks, kpar, AT_K_FULL);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem_base + C::T_S + buf * 128, umma_desc_sw128_a16(q_smem + c * CH16 + kk * 2, 16, 1024),
                    umma_desc_sw128_a16(k_smem + ks * TILE16 + c * CH16 + kk * 2, 16, 1024), IDESC_QK, (c | kk) != 0);
        tc_commit(s_full + buf);
        tc_commit(k_empty + ks);
      };
      if (me == 0) {
        mbar_wait(q_full, 0, AT_Q_FULL);
        issue_qk(0, 0, 0);
        if (n_kv > 1) issue_qk(1, 1, 0);
        if (n_kv > 2) issue_qk(2, 0, 1);
      }
      // Ring depths: K(j + 3) goes into the slot of K(j + 1), free since Q K^T of block j + 1 (issued behind P V of block
      // j - 2) completed; V(j + 1) goes into the slot of V(j - 2), free since P V of block j - 2 completed — both more
      // than a block ahead of their use, which covers the L2 -> shared-memory latency.
      for (int j = me; j < n_kv; j += 2) {
        const int buf = j % 3;                                    // S buffer and V slot
        const uint32_t half = static_cast<uint32_t>(j >> 1);      // this issuer's block counter
        mbar_wait(v_full + buf, static_cast<uint32_t>(j / 3) & 1u, AT_V_FULL);
        trace_mark(p.trace, 1 + 2 * me, j, 0);
        mbar_wait(p_full + me, half & 1u, AT_P_FULL);
        if (j > 0) mbar_wait(tok + me, (me == 1 ? half : half + 1u) & 1u, AT_TOK);  // P V of block j - 1 has completed
        tc_fence_after();
        trace_mark(p.trace, 1 + 2 * me, j, 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // 16 keys per MMA; V tile: [128 keys][64 d] x 2 boxes, MN-major B
          umma_ts(tmem_base + C::T_O, tmem_base + C::T_S + buf * 128 + kk * 8,
                  umma_desc_sw128_a16(v_smem + buf * TILE16 + kk * 128, C::CHUNK, 1024), IDESC_PV, (j > 0) || kk != 0);
        tc_commit(tok + (me ^ 1));  // arrives when this P V has COMPLETED (see above)
        tc_commit(pv_done);
        tc_commit(v_empty + buf);
        trace_mark(p.trace, 1 + 2 * me, j, 2);
        if (j + 3 < n_kv) issue_qk(buf, (j + 1) & 1, static_cast<uint32_t>((j + 3) >> 1) & 1u);  // into the buffer P V frees
        if (j == n_kv - 1) tc_commit(o_full);
        trace_mark(p.trace, 1 + 2 * me, j, 3);
      }
    }
  } else if (warp >= SOFTMAX_WARP0) {
    // ================================ softmax sets: two threads per query row ======================
    const int sw = warp - SOFTMAX_WARP0;
    const int x = sw >> 3;         // set: key blocks j with (j & 1) == x
    const int hf = (sw >> 2) & 1;  // which 64 of the block's 128 key columns
    const int quarter = warp & 3;  // TMEM lane quarter a warp may access = warp id % 4
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t o_half = tmem_base + lane_addr + C::T_O + hf * 64;  // the O columns this thread rescales
    const float sl2 = p.scale_log2;
    float* mx = reinterpret_cast<float*>(smem + C::OFF_MX) + x * (2 * 2 * 128);
    float* mrun = reinterpret_cast<float*>(smem + C::OFF_MRUN);
    const uint32_t pair_bar = 1 + x * 4 + quarter;  // the two warps (column halves) that share this set's row quarter
    const bool tr = row == 0 && hf == 0;
    const int trole = x == 0 ? 0 : 2;

    float m_own = -CUDART_INF_F;  // running max as of this set's previous block
    float l = 0.f;                // this set's partial row sum, relative to m_own
    int buf = x;                  // j % 3
    uint32_t use = 0;             // j / 3
    for (int j = x; j < n_kv; j += 2) {
      const uint32_t s_addr = tmem_base + lane_addr + C::T_S + buf * 128 + hf * 64;
      const uint32_t p_addr = tmem_base + lane_addr + C::T_S + buf * 128 + hf * 32;
      if (tr) trace_mark(p.trace, trole, j, 0);
      mbar_wait(s_full + buf, use & 1, AT_S_FULL);
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
      // exchange the half-row maxima between the two warps that share these 32 rows (a 64-thread barrier: no waiting for
      // the slowest of the set's eight warps); it also orders this block's S reads before the partner's P writes (P
      // overlays S columns the partner has just read)
      float* slot = mx + ((j >> 1) & 1) * 256;
      slot[hf * 128 + row] = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      named_bar_sync(pair_bar, 64);
      const float bm = fmaxf(slot[row], slot[128 + row]);
      // running max after the previous block, published by the other set
      float m_prev = -CUDART_INF_F;
      if (j > 0) {
        mbar_wait(mpub + ((j - 1) & 1), static_cast<uint32_t>((j - 1) >> 1) & 1u, AT_MPUB);
        m_prev = mrun[((j - 1) & 1) * 128 + row];
      }
      // lazy: keep the stale max while the block maximum is within 2^8 of it (both threads of a row decide alike)
      const bool raise = (j == 0) || (bm - m_prev) * sl2 > 8.f;
      const float m_j = raise ? fmaxf(bm, m_prev) : m_prev;
      // every thread of the set arrives — after its own read of m_prev — so that the other set, which overwrites that
      // slot two blocks later, cannot race with a slower reader of this set
      if (hf == 0) mrun[(j & 1) * 128 + row] = m_j;
      mbar_arrive(mpub + (j & 1));
      if (tr) trace_mark(p.trace, trole, j, 2);
      if (m_j != m_own) {  // m moved since this set's previous block (by either set): re-base the partial row sum
        l = (l == 0.f) ? 0.f : l * ex2_approx((m_own - m_j) * sl2);
        m_own = m_j;
      }
      const bool need = raise && j > 0;
      if (__any_sync(0xffffffffu, need)) {
        // O holds blocks <= j - 1 at scale m_prev once P V of block j - 1 has landed. S(j) being full only implies
        // that P V of block j - 3 is complete, so the one-phase-per-block barrier may still be two phases back: wait
        // for block j - 2 first, otherwise the parity test for block j - 1 passes spuriously.
        if (j > 1) mbar_wait(pv_done, static_cast<uint32_t>(j - 2) & 1u, AT_PV_DONE);
        mbar_wait(pv_done, static_cast<uint32_t>(j - 1) & 1u, AT_PV_DONE);
        tc_fence_after();
        const float f = need ? ex2_approx((m_prev - m_j) * sl2) : 1.f;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 8) {
          uint32_t ou[8];
          tmem_ld_x8(o_half + c0, ou);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c) ou[c] = __float_as_uint(__uint_as_float(ou[c]) * f);
          tmem_st_x8(o_half + c0, ou);
        }
      }
      if (tr) trace_mark(p.trace, trole, j, 3);
      // p = 2^(s * scale_log2 - m * scale_log2) on packed pairs; EMU of every 8 pairs on the FMA pipe (ex2_poly2)
      const float msc = m_j * sl2;
      const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-msc, -msc);
      float2 lacc = make_float2(0.f, 0.f);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float2 xx = __ffma2_rn(make_float2(s[c0 + c], s[c0 + c + 1]), sc2, nm2);
          float2 pv;
          if ((((c0 + c) >> 1) & 7) < C::EMU) {
            pv = ex2_poly2(xx);
          } else {
            pv.x = ex2_approx(xx.x);
            pv.y = ex2_approx(xx.y);
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
      mbar_arrive(p_full + x);
      if (tr) trace_mark(p.trace, trole, j, 5);
      buf += 2;
      if (buf >= 3) { buf -= 3; ++use; }
    }

    // ---- epilogue: align the partial sums to the final max, add them, O / l -> bf16 -> global; lse --------------
    const int last = n_kv - 1;
    float m_fin = m_own;
    if ((last & 1) != x) {
      mbar_wait(mpub + (last & 1), static_cast<uint32_t>(last >> 1) & 1u, AT_MPUB);
      m_fin = mrun[(last & 1) * 128 + row];
    }
    if (m_own != m_fin) l = (l == 0.f) ? 0.f : l * ex2_approx((m_own - m_fin) * sl2);
    float* lsum = reinterpret_cast<float*>(smem + C::OFF_LSUM);
    lsum[(x * 2 + hf) * 128 + row] = l;
    named_bar_sync(9, 512);
    l = (lsum[row] + lsum[128 + row]) + (lsum[256 + row] + lsum[384 + row]);
    mbar_wait(o_full, 0, AT_O_FULL);
    tc_fence_after();
    const float inv = 1.f / l;
    const int row_g = q0 + row;
    const bool valid_row = row_g < q_len;
    const int part = x * 2 + hf;  // which 32 of the 128 output columns this thread writes
    __nv_bfloat16* optr = p.o + bq * p.o_sb + static_cast<int64_t>(q_base + row_g) * p.o_sl + h * p.o_sh + part * 32;
    {
      uint32_t ou[32];
      uint4 wv[4];
      tmem_ld_x32(tmem_base + lane_addr + C::T_O + part * 32, ou);
      tc_wait_ld();
      if (valid_row) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(ou[c + 0]) * inv, __uint_as_float(ou[c + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(ou[c + 2]) * inv, __uint_as_float(ou[c + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(ou[c + 4]) * inv, __uint_as_float(ou[c + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(ou[c + 6]) * inv, __uint_as_float(ou[c + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + c) = w;
          wv[c >> 3] = w;
        }
        if (p.sc_n > 0 && row_g < p.sc_n * p.sc_rpr) {
          // fused exchange: the same 64 bytes go to the rank that owns this query row (peer-mapped pointer)
          const int dst = row_g / p.sc_rpr;
          __nv_bfloat16* pptr = p.sc_base[dst] + static_cast<int64_t>(row_g - dst * p.sc_rpr) * p.sc_sl + h * p.sc_sh + part * 32;
#pragma unroll
          for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(pptr + c * 8) = wv[c];
        }
      }
    }
    if (valid_row && part == 0) p.lse[bq * p.lse_sb + h * p.lse_sh + q_base + row_g] = m_fin * p.scale + __logf(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PROD_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace

cudaError_t launch_attn_fwd_alt(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                                const AttnFwdParams& p, cudaStream_t stream) {
  using C = AltCfg;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_alt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::BYTES);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((p.seq.Lq + 127) / 128, p.seq.H, p.seq.nprob);
  attn_fwd_alt_kernel<<<grid, C::THREADS, C::BYTES, stream>>>(tm_q, tm_k, tm_v, p);
  return cudaGetLastError();
}

cudaError_t attn_fwd_alt_set_debug_ptr(unsigned int* ptr) { return cudaMemcpyToSymbol(g_vt_dbg, &ptr, sizeof(ptr)); }

}  // namespace vt
