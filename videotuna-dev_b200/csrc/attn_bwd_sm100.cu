// attn_bwd_sm100.cu — flash-attention backward for sm_100a (dense, non-causal), 5 GEMMs per (Q tile, KV tile).
//
// Replaces autograd through the reference torch attention path (same call sites as attn_fwd_sm100.cu).
//
// One CTA owns one 128-key tile of one (problem, head): K and V stay in shared memory, dK and dV accumulate in TMEM
// over a loop on 128-row query tiles; dQ is produced per (Q tile, KV tile), staged through shared memory in 32-column
// fp32 chunks and added into an fp32 accumulator in global memory with TMA reductions
// (cp.reduce.async.bulk.tensor .add): one bulk L2 reduction per 16 KB instead of 4096 per-thread red.global ops.
//
// Transposed formulation (keys on TMEM lanes) so that P^T and dS^T are directly MMA A-operands:
//   MMA1  S^T  = K  Q_i^T         SS   (A = K tile K-major, B = Q_i K-major)               -> TMEM S
//   MMA2  dP^T = V  dO_i^T        SS                                                        -> TMEM dP
//   SM1   P^T  = exp2(S^T * c - lse_i)  (compute warpgroups; bf16 P^T written back into the S columns)
//   MMA3  dV  += P^T dO_i         TS   (A = P^T from TMEM, B = dO_i MN-major)
//   SM2   dS^T = P^T o (dP^T - delta_i) -> bf16 -> shared memory tile [key][q] (128B-swizzled)
//   MMA4  dQ_i = dS  K            SS   (A = dS as MN-major view of the same tile, B = K MN-major) -> TMEM dQ
//   MMA5  dK  += dS^T Q_i         SS   (A = dS^T K-major, B = Q_i MN-major)   (runs while dQ_i is drained)
//   DR    dQ_i: TMEM -> registers -> swizzled smem chunk -> TMA reduce-add into dq_acc (drain warpgroup)
// Warps: 0-7 compute (thread = key row x half of the q columns), 8-11 dQ drain (thread = q row), 13 MMA issuer,
// 12/14/15 loaders (K/V by TMA once; Q and dO tiles by cp.async; lse/delta staging).
// TMEM columns: S [0,128) (P^T bf16 at [32,96)), dP [128,256), dV [256,256+D), dK [256+D,256+2D),
//   dQ: D=128 aliases dP (dP is dead once dS is in shared memory); D=64 uses [384,448).
#include <cstdlib>
#include <cuda_bf16.h>
#include <math_constants.h>

#include "attn_common.h"
#include "capi_util.h"
#include "sm100_ptx.cuh"

namespace vt {
namespace {

template <int D>
struct BwdCfg {
  static_assert(D == 64 || D == 128, "head dim must be 64 or 128");
  static constexpr int KCH = D / 64;
  static constexpr int CHUNK = 128 * 128;   // bytes of one [128 rows][64 bf16] swizzled box
  static constexpr int TILE = CHUNK * KCH;  // 128 x D bf16
  // Ring depths. D = 128 fills shared memory with Q x2 / dO x1. D = 64 has room for Q x3 / dO x2, which it needs: a Q
  // slot is released by dK of iteration i - QS + 1 and its TMA load then queues behind that iteration's dQ reductions
  // (~2000 cycles measured), so with two slots S^T of the next tile waited ~1400 cycles for Q in every iteration.
  static constexpr int QS = (D == 64) ? 3 : 2;
  static constexpr int DOS = (D == 64) ? 2 : 1;
  static constexpr bool TWO_ISSUERS = (D == 64);  // see the MMA issuer section of the kernel
#ifdef VT_BWD_EMU
  static constexpr int EMU = VT_BWD_EMU;    // of every 8 exponential pairs, how many run on the FMA pipe (ex2_poly2)
#else
  static constexpr int EMU = (D == 64) ? 3 : 2;
#endif
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE;
  static constexpr int OFF_Q = OFF_V + TILE;
  static constexpr int OFF_DO = OFF_Q + QS * TILE;
  static constexpr int OFF_DS = OFF_DO + DOS * TILE;    // 128 x 128 bf16 = 2 boxes
  static constexpr int DQ_CHUNK = 128 * 128;            // bytes: 128 rows x 32 fp32 columns (one swizzled box)
  static constexpr int OFF_DQS = OFF_DS + 2 * CHUNK;    // 2 staging buffers for the dQ TMA reduction
  static constexpr int OFF_STAT = OFF_DQS + 2 * DQ_CHUNK;  // QS x {lse_log2[128], delta[128]} fp32
  static constexpr int OFF_BAR = OFF_STAT + QS * 1024;
  static constexpr int NBAR = 1 + 2 * QS + QS + 2 * DOS + 9;  // dkv_full[2], tok_s unused slot kept for layout
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  // The dynamic shared memory base is 1024-byte aligned (checked at kernel entry), so no alignment slack is spent:
  // at D = 128 the layout uses 231 608 of the 232 448 bytes a CTA can have.
  static constexpr int BYTES = OFF_TMEM + 16;
  static_assert(BYTES <= 232448, "shared memory budget exceeded");
  static constexpr int THREADS = 512;
  static constexpr uint32_t T_S = 0, T_P = 32, T_DP = 128, T_DV = 256, T_DK = 256 + D;
  static constexpr uint32_t T_DQ = (D == 128) ? 128 : 384;
  static constexpr bool DQ_ALIASES_DP = (D == 128);
};

enum : uint32_t {
  BT_KV_FULL = 0x200, BT_Q_FULL, BT_Q_EMPTY, BT_STAT_FULL, BT_DO_FULL, BT_DO_EMPTY, BT_S_FULL, BT_P_READY, BT_DP_FULL,
  BT_DS_READY, BT_DQ_FULL, BT_DQ_DRAINED, BT_DKV_FULL, BT_DS_FREE, BT_ALIGN, BT_TOK_S
};


// Q and dO tiles arrive by TMA. A cp.async (LSU) producer was tried to leave the per-SM TMA engine to the dQ reduction
// and measured 40 % slower on K1 (659 vs 1076 TFLOP/s): LSU writes into shared memory starve behind the UMMA operand
// fetch, which saturates the 128 B/clk shared-memory port, whereas TMA writes do not (history: commit 782d835).
// DIRECT: single key tile per (sample, head) — the dQ tile is written straight to dq (AttnBwdParams::dq_direct). A compile-
// time variant: as a run-time branch in the drain loop it cost the general kernel 2 % at K1 (same-box A/B).
template <int D, bool DIRECT>
__global__ void __launch_bounds__(BwdCfg<D>::THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, float* __restrict__ dq_acc, const int Lq_total,
                const AttnBwdParams p) {
  using C = BwdCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) watchdog_trap(BT_ALIGN);  // swizzled tiles need a 1024-byte aligned base
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  const int prob = blockIdx.z, h = blockIdx.y;
  int q_base = 0, q_len = p.seq.Lq, bq = prob;
  int k_base = 0, k_len = p.seq.Lk, bk = prob;
  int k_rows_total = p.seq.Lk;  // rows that exist in the k/v tensors for this problem (fixed mode: padded length)
  if (p.seq.cu_q != nullptr) {
    q_base = p.seq.cu_q[prob];
    q_len = p.seq.cu_q[prob + 1] - q_base;
    bq = 0;
  }
  if (p.seq.cu_k != nullptr) {
    k_base = p.seq.cu_k[prob];
    k_len = p.seq.cu_k[prob + 1] - k_base;
    k_rows_total = k_len;
    bk = 0;
  } else if (p.seq.seqlens_k != nullptr) {
    k_len = min(max(p.seq.seqlens_k[prob], 0), p.seq.Lk);
  }
  const int kv0 = blockIdx.x * 128;
  if (kv0 >= k_rows_total) return;
  const int n_q = (q_len + 127) >> 7;
  const int kv_valid = min(k_len - kv0, 128);  // may be <= 0: tile holds only masked (padding) keys

  if (kv_valid <= 0 || n_q == 0) {  // nothing attends to these keys: dK = dV = 0
    for (int r = threadIdx.x; r < 128; r += blockDim.x) {
      if (kv0 + r < k_rows_total) {
        __nv_bfloat16* dkp = p.dk + bk * p.dk_sb + static_cast<int64_t>(k_base + kv0 + r) * p.dk_sl + h * p.dk_sh;
        __nv_bfloat16* dvp = p.dv + bk * p.dv_sb + static_cast<int64_t>(k_base + kv0 + r) * p.dv_sl + h * p.dv_sh;
        for (int c = 0; c < D; c += 8) {
          *reinterpret_cast<uint4*>(dkp + c) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(dvp + c) = make_uint4(0, 0, 0, 0);
        }
      }
    }
    if (DIRECT) {  // no valid key for this (sample, head): its dQ rows are zero and nobody else writes them
      for (int r = threadIdx.x; r < q_len; r += blockDim.x) {
        __nv_bfloat16* dqp = p.dq_direct + bq * p.dq_sb + static_cast<int64_t>(q_base + r) * p.dq_sl + h * p.dq_sh;
        for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(dqp + c) = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* kv_full = bars;
  uint64_t* q_full = kv_full + 1;
  uint64_t* q_empty = q_full + C::QS;
  uint64_t* stat_full = q_empty + C::QS;
  uint64_t* do_full = stat_full + C::QS;
  uint64_t* do_empty = do_full + C::DOS;
  uint64_t* s_full = do_empty + C::DOS;
  uint64_t* p_ready = s_full + 1;
  uint64_t* dp_full = p_ready + 1;
  uint64_t* ds_ready = dp_full + 1;
  uint64_t* dq_full = ds_ready + 1;
  uint64_t* dq_drained = dq_full + 1;
  uint64_t* dkv_full = dq_drained + 1;
  uint64_t* ds_free = dkv_full + 2;    // dkv_full[0]: dK complete (issuer 0), dkv_full[1]: dV complete (issuer 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);

  constexpr int PROD_WARP = 12, MMA_WARP = 13, MMA_WARP1 = 14;

  if (warp == PROD_WARP && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
  }
  if (warp == MMA_WARP && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < C::QS; ++i) {
      mbar_init(q_full + i, 1);
      mbar_init(q_empty + i, 1);
      mbar_init(stat_full + i, 32);
    }
    for (int i = 0; i < C::DOS; ++i) {
      mbar_init(do_full + i, 1);
      mbar_init(do_empty + i, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(dp_full, 1);
    mbar_init(ds_ready, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_drained, 128);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_full + 1, 1);
    mbar_init(ds_free, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform: stays in a uniform register

  if (warp >= PROD_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == PROD_WARP) {
      // ================================ producer: K/V once, Q / dO rings, lse / delta staging =======
      if (lane == 0) {
        mbar_arrive_expect_tx(kv_full, 2 * C::TILE);
#pragma unroll
        for (int c = 0; c < C::KCH; ++c) {
          tma_load_4d(smem + C::OFF_K + c * C::CHUNK, &tm_k, kv_full, c * 64, k_base + kv0, h, bk);
          tma_load_4d(smem + C::OFF_V + c * C::CHUNK, &tm_v, kv_full, c * 64, k_base + kv0, h, bk);
        }
      }
      const float* lse_row = p.lse + bq * p.lse_sb + h * p.lse_sh + q_base;
      const float* dl_row = p.delta + bq * p.lse_sb + h * p.lse_sh + q_base;
      for (int i = 0; i < n_q; ++i) {
        const int s = i % C::QS, ds = i % C::DOS;
        mbar_wait(q_empty + s, ((i / C::QS) & 1) ^ 1, BT_Q_EMPTY);
        if (lane == 0) {
          trace_mark(p.trace, 3, i, 0);
          mbar_arrive_expect_tx(q_full + s, C::TILE);
#pragma unroll
          for (int c = 0; c < C::KCH; ++c)
            tma_load_4d(smem + C::OFF_Q + s * C::TILE + c * C::CHUNK, &tm_q, q_full + s, c * 64, q_base + i * 128, h, bq);
        }
        // -lse * log2(e) and -delta for the 128 rows of this Q tile (negated so that the compute warps use them as
        // FFMA2 / FADD2 addends); rows past q_len get -inf so that P == 0 there.
        float* st = reinterpret_cast<float*>(smem + C::OFF_STAT + s * 1024);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = r * 32 + lane;
          const int qrow = i * 128 + row;
          float l = -CUDART_INF_F, d = 0.f;
          if (qrow < q_len) {
            l = lse_row[qrow] * -1.4426950408889634f;
            d = -dl_row[qrow];
          }
          st[row] = l;
          st[128 + row] = d;
        }
        mbar_arrive(stat_full + s);
        if (lane == 0) {
          mbar_wait(do_empty + ds, ((i / C::DOS) & 1) ^ 1, BT_DO_EMPTY);
          trace_mark(p.trace, 3, i, 1);
          mbar_arrive_expect_tx(do_full + ds, C::TILE);
#pragma unroll
          for (int c = 0; c < C::KCH; ++c)
            tma_load_4d(smem + C::OFF_DO + ds * C::TILE + c * C::CHUNK, &tm_do, do_full + ds, c * 64, q_base + i * 128, h, bq);
        }
        __syncwarp();
      }
    } else if ((warp == MMA_WARP || warp == MMA_WARP1) && elect_one()) {
      // ================================ MMA issuers ==============================================
      // TWO issuer threads. A tcgen05.mma / commit sequence holds its issuing thread until the tensor pipe has nearly
      // drained it, and every barrier poll plus the scalar code around it then runs with the pipe idle (a lone thread
      // executes ~10 cycles per instruction next to four busy warps). With two threads one's polling overlaps the
      // other's MMAs:
      //   issuer 0 (warp 13): dQ = dS K and dK += dS^T Q          — wait for dS from the compute warps
      //   issuer 1 (warp 14): dV += P^T dO, then S^T = K Q^T and dP^T = V dO^T of the NEXT Q tile
      // S^T(i+1) overwrites the TMEM columns P^T(i) lives in, so it must enter the pipe behind dV(i): both are issued
      // by the same thread, in program order (a hand-off flag between two threads does not order their MMAs: the arrive
      // can overtake the issuing thread's own tcgen05.mma — measured as wrong dV). Cross-thread hazards are all covered
      // by barriers that already exist: dP^T(i+1) overwrites dP^T(i) / dQ(i) only after dq_drained(i) (D = 128) or
      // ds_ready(i) (D = 64); the Q tile is released by issuer 0 after dK(i), which S^T(i) precedes by data dependence.
      constexpr uint32_t IDESC_ST = umma_idesc_bf16(128, 128, 0, 0);  // S^T, dP^T: both operands K-major
      constexpr uint32_t IDESC_KD = umma_idesc_bf16(128, D, 0, 1);    // dV, dK: A K-major (TMEM / smem), B MN-major
      constexpr uint32_t IDESC_DQ = umma_idesc_bf16(128, D, 1, 1);    // dQ: A = dS MN-major, B = K MN-major
      // shared-memory addresses in 16-byte units (the descriptor's address field); all tile offsets are constants
      const uint32_t sb16 = smem_u32(smem) >> 4;
      const uint32_t k_s = sb16 + (C::OFF_K >> 4), v_s = sb16 + (C::OFF_V >> 4);
      const uint32_t q_s = sb16 + (C::OFF_Q >> 4), do_s = sb16 + (C::OFF_DO >> 4);
      const uint32_t ds_s = sb16 + (C::OFF_DS >> 4);
      constexpr uint32_t CH16 = C::CHUNK >> 4, TILE16 = C::TILE >> 4;

      auto mma_kmajor_pair = [&](uint32_t d_tmem, uint32_t a_base, uint32_t b_base) {  // D = A B^T over the head dim
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(d_tmem, umma_desc_sw128_a16(a_base + c * CH16 + kk * 2, 16, 1024),
                    umma_desc_sw128_a16(b_base + c * CH16 + kk * 2, 16, 1024), IDESC_ST, (c | kk) != 0);
      };

      mbar_wait(kv_full, 0, BT_KV_FULL);
      if (!C::TWO_ISSUERS) {
        // ---- one issuer (head dim 128: the kernel is bound by the dQ reduction, a second polling thread only costs power) ----
        if (warp == MMA_WARP) {
          mbar_wait(q_full + 0, 0, BT_Q_FULL);
          tc_fence_after();
          mma_kmajor_pair(tmem + C::T_S, k_s, q_s);
          tc_commit(s_full);
          mbar_wait(do_full + 0, 0, BT_DO_FULL);
          tc_fence_after();
          mma_kmajor_pair(tmem + C::T_DP, v_s, do_s);
          tc_commit(dp_full);

          for (int i = 0; i < n_q; ++i) {
            const int s = i % C::QS, ds = i % C::DOS;
            const bool has_next = i + 1 < n_q;
            // ---- dV += P^T dO_i ----
            trace_mark(p.trace, 1, i, 0);
            mbar_wait(p_ready, i & 1, BT_P_READY);
            tc_fence_after();
            trace_mark(p.trace, 1, i, 1);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts(tmem + C::T_DV, tmem + C::T_P + kk * 8,
                      umma_desc_sw128_a16(do_s + ds * TILE16 + kk * 128, C::CHUNK, 1024), IDESC_KD, (i > 0) || kk != 0);
            tc_commit(do_empty + ds);
            // ---- S^T for the next Q tile ----
            if (has_next) {
              const int sn = (i + 1) % C::QS;
              mbar_wait(q_full + sn, ((i + 1) / C::QS) & 1, BT_Q_FULL);
              tc_fence_after();
              trace_mark(p.trace, 1, i, 2);
              mma_kmajor_pair(tmem + C::T_S, k_s, q_s + sn * TILE16);
              tc_commit(s_full);
            }
            // ---- dQ_i = dS K (drained while the next MMA runs) ; dK += dS^T Q_i ----
            mbar_wait(ds_ready, i & 1, BT_DS_READY);
            tc_fence_after();
            trace_mark(p.trace, 1, i, 3);
            if (!C::DQ_ALIASES_DP && i > 0) {
              mbar_wait(dq_drained, (i - 1) & 1, BT_DQ_DRAINED);
              tc_fence_after();
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // k = 16 keys per step
              umma_ss(tmem + C::T_DQ, umma_desc_sw128_a16(ds_s + kk * 128, C::CHUNK, 1024),
                      umma_desc_sw128_a16(k_s + kk * 128, C::CHUNK, 1024), IDESC_DQ, kk != 0);
            tc_commit(dq_full);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // k = 16 query rows per step; dS^T tile is [key][q] in two 64-q boxes
              umma_ss(tmem + C::T_DK, umma_desc_sw128_a16(ds_s + (kk >> 2) * CH16 + (kk & 3) * 2, 16, 1024),
                      umma_desc_sw128_a16(q_s + s * TILE16 + kk * 128, C::CHUNK, 1024), IDESC_KD, (i > 0) || kk != 0);
            tc_commit(q_empty + s);
            tc_commit(ds_free);
            // ---- dP^T for the next Q tile ----
            if (has_next) {
              if (C::DQ_ALIASES_DP) {
                mbar_wait(dq_drained, i & 1, BT_DQ_DRAINED);
                tc_fence_after();
              }
              trace_mark(p.trace, 1, i, 4);
              const int dn = (i + 1) % C::DOS;
              mbar_wait(do_full + dn, ((i + 1) / C::DOS) & 1, BT_DO_FULL);
              tc_fence_after();
              trace_mark(p.trace, 1, i, 5);
              mma_kmajor_pair(tmem + C::T_DP, v_s, do_s + dn * TILE16);
              tc_commit(dp_full);
            }
          }
          tc_commit(dkv_full);
          tc_commit(dkv_full + 1);
        }
      } else if (warp == MMA_WARP1) {
        // prologue: S^T(0), dP^T(0)
        mbar_wait(q_full + 0, 0, BT_Q_FULL);
        tc_fence_after();
        mma_kmajor_pair(tmem + C::T_S, k_s, q_s);
        tc_commit(s_full);
        mbar_wait(do_full + 0, 0, BT_DO_FULL);
        tc_fence_after();
        mma_kmajor_pair(tmem + C::T_DP, v_s, do_s);
        tc_commit(dp_full);
        for (int i = 0; i < n_q; ++i) {
          const int ds = i % C::DOS;
          const bool has_next = i + 1 < n_q;
          // ---- dV += P^T dO_i ----
          trace_mark(p.trace, 3, i, 2);
          mbar_wait(p_ready, i & 1, BT_P_READY);
          tc_fence_after();
          trace_mark(p.trace, 3, i, 3);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ts(tmem + C::T_DV, tmem + C::T_P + kk * 8,
                    umma_desc_sw128_a16(do_s + ds * TILE16 + kk * 128, C::CHUNK, 1024), IDESC_KD, (i > 0) || kk != 0);
          tc_commit(do_empty + ds);  // dP^T(i) was issued by this thread before dV(i): one commit covers both readers
          if (has_next) {
            // ---- S^T(i+1) = K Q_{i+1}^T ----
            const int sn = (i + 1) % C::QS, dn = (i + 1) % C::DOS;
            mbar_wait(q_full + sn, ((i + 1) / C::QS) & 1, BT_Q_FULL);
            tc_fence_after();
            mma_kmajor_pair(tmem + C::T_S, k_s, q_s + sn * TILE16);
            tc_commit(s_full);
            // ---- dP^T(i+1) = V dO_{i+1}^T ----
            mbar_wait(do_full + dn, ((i + 1) / C::DOS) & 1, BT_DO_FULL);
            if (C::DQ_ALIASES_DP) mbar_wait(dq_drained, i & 1, BT_DQ_DRAINED);  // dQ(i) has left these columns
            else mbar_wait(ds_ready, i & 1, BT_DS_READY);                        // dP^T(i) has been read
            tc_fence_after();
            trace_mark(p.trace, 3, i, 4);
            mma_kmajor_pair(tmem + C::T_DP, v_s, do_s + dn * TILE16);
            tc_commit(dp_full);
          }
        }
        tc_commit(dkv_full + 1);
      } else {
        for (int i = 0; i < n_q; ++i) {
          const int s = i % C::QS;
          // ---- dQ_i = dS K (drained while the next MMAs run) ; dK += dS^T Q_i ----
          trace_mark(p.trace, 1, i, 0);
          mbar_wait(ds_ready, i & 1, BT_DS_READY);
          if (!C::DQ_ALIASES_DP && i > 0) mbar_wait(dq_drained, (i - 1) & 1, BT_DQ_DRAINED);
          tc_fence_after();
          trace_mark(p.trace, 1, i, 1);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // k = 16 keys per step
            umma_ss(tmem + C::T_DQ, umma_desc_sw128_a16(ds_s + kk * 128, C::CHUNK, 1024),
                    umma_desc_sw128_a16(k_s + kk * 128, C::CHUNK, 1024), IDESC_DQ, kk != 0);
          tc_commit(dq_full);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // k = 16 query rows per step; dS^T tile is [key][q] in two 64-q boxes
            umma_ss(tmem + C::T_DK, umma_desc_sw128_a16(ds_s + (kk >> 2) * CH16 + (kk & 3) * 2, 16, 1024),
                    umma_desc_sw128_a16(q_s + s * TILE16 + kk * 128, C::CHUNK, 1024), IDESC_KD, (i > 0) || kk != 0);
          tc_commit(q_empty + s);
          tc_commit(ds_free);
          trace_mark(p.trace, 1, i, 2);
        }
        tc_commit(dkv_full);
      }
    }
  } else if (warp >= 8) {
    // ================================ dQ drain warpgroup ===========================================
    // The whole 128 x D fp32 tile is pulled into registers first, so the TMEM columns (which dP of the next iteration
    // reuses at D = 128) are released ~100 cycles after the MMA completes; the registers are then fed to the TMA
    // reduction through two 16 KB staging buffers at the engine's own pace (~640 cycles per box, measured).
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool leader = threadIdx.x == 256;
    const uint32_t dq_addr = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + C::T_DQ;
    uint8_t* stage = smem + C::OFF_DQS;
    const int sw = row & 7;
    // The TMA engine sustains ~25.6 B/clk of fp32 reduction per SM (measured, tools/tma_reduce_rate.py). Sending one of
    // the chunks through the LSU instead (red.global.add.v4.f32 from registers) was measured 25 % slower.
    constexpr int NCH = D / 32;
    uint32_t g = 0;  // running chunk counter: staging buffer = g & 1
    for (int i = 0; i < n_q; ++i) {
      if (leader) trace_mark(p.trace, 2, i, 0);
      mbar_wait(dq_full, i & 1, BT_DQ_FULL);
      tc_fence_after();
      if (leader) trace_mark(p.trace, 2, i, 1);
      uint32_t r[D];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld_x32(dq_addr + c * 32, r + c * 32);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(dq_drained);  // every TMEM read of this tile is complete: the issuer may overwrite the columns
      if (leader) trace_mark(p.trace, 2, i, 2);
      if (DIRECT) {  // single key tile: this tile IS dQ (see AttnBwdParams::dq_direct)
        const int qrow = i * 128 + row;
        if (qrow < q_len) {
          __nv_bfloat16* dqp = p.dq_direct + bq * p.dq_sb + static_cast<int64_t>(q_base + qrow) * p.dq_sl + h * p.dq_sh;
#pragma unroll
          for (int c = 0; c < D; c += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(r[c + 0]) * p.scale, __uint_as_float(r[c + 1]) * p.scale);
            w.y = pack_bf16x2(__uint_as_float(r[c + 2]) * p.scale, __uint_as_float(r[c + 3]) * p.scale);
            w.z = pack_bf16x2(__uint_as_float(r[c + 4]) * p.scale, __uint_as_float(r[c + 5]) * p.scale);
            w.w = pack_bf16x2(__uint_as_float(r[c + 6]) * p.scale, __uint_as_float(r[c + 7]) * p.scale);
            *reinterpret_cast<uint4*>(dqp + c) = w;
          }
        }
        continue;
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c, ++g) {
        // the reduction issued two chunks ago has finished reading this staging buffer
        if (leader) tma_wait_group_read<1>();
        named_bar_sync(1, 128);
        uint8_t* buf = stage + (g & 1) * C::DQ_CHUNK;
        uint8_t* rowp = buf + row * 128;  // 128B swizzle: 16-byte chunk ch of row r lives at ch ^ (r & 7)
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<uint4*>(rowp + ((ch ^ sw) << 4)) =
              make_uint4(r[c * 32 + 4 * ch], r[c * 32 + 4 * ch + 1], r[c * 32 + 4 * ch + 2], r[c * 32 + 4 * ch + 3]);
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (leader) {
          // rows past the end of the tensor are clipped by TMA; rows of another varlen segment receive exact zeros
          tma_reduce_add_4d(&tm_dq, buf, c * 32, q_base + i * 128, h, bq);
          tma_commit_group();
        }
      }
      if (leader) trace_mark(p.trace, 2, i, 3);
    }
    if (leader) tma_wait_group<0>();
  } else {
    // ================================ compute warpgroups ===========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int half = warp >> 2;  // which 64 query columns
    const int quarter = warp & 3;
    const int krow = quarter * 32 + lane;  // key row of this thread
    const bool key_valid = krow < kv_valid;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_addr = tmem + lane_addr + C::T_S + half * 64;
    const uint32_t p_addr = tmem + lane_addr + C::T_P + half * 32;
    const uint32_t dp_addr = tmem + lane_addr + C::T_DP + half * 64;
    const float sl2 = p.scale_log2;
    // this thread's 128-byte row inside the dS^T box of its column half (128B swizzle: 16-byte chunk c -> c ^ (row & 7))
    uint8_t* ds_row = smem + C::OFF_DS + half * C::CHUNK + krow * 128;
    const int sw = krow & 7;

    for (int i = 0; i < n_q; ++i) {
      const int s = i % C::QS;
      const float* st = reinterpret_cast<const float*>(smem + C::OFF_STAT + s * 1024) + half * 64;
      const bool tr = threadIdx.x == 0;
      if (tr) trace_mark(p.trace, 0, i, 0);
      mbar_wait(stat_full + s, (i / C::QS) & 1, BT_STAT_FULL);
      mbar_wait(s_full, i & 1, BT_S_FULL);
      tc_fence_after();
      if (tr) trace_mark(p.trace, 0, i, 1);
      float pr[64];
      {
        uint32_t su[64];
        tmem_ld_x32(s_addr, su);
        tmem_ld_x32(s_addr + 32, su + 32);
        tc_wait_ld();
        // P^T = 2^(S^T * scale_log2 - lse_log2) on packed pairs; EMU of every 8 pairs take the FMA-pipe polynomial
        const float2 sc2 = make_float2(sl2, sl2);
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(st + c);  // -lse * log2(e)
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const float2 nl = e == 0 ? make_float2(l4.x, l4.y) : make_float2(l4.z, l4.w);
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(su[c + e]), __uint_as_float(su[c + e + 1])), sc2, nl);
            float2 pv;
            if ((((c + e) >> 1) & 7) < C::EMU) {
              pv = ex2_poly2(x);
            } else {
              pv.x = ex2_approx(x.x);
              pv.y = ex2_approx(x.y);
            }
            pr[c + e] = pv.x;
            pr[c + e + 1] = pv.y;
          }
        }
      }
      if (!key_valid) {
#pragma unroll
        for (int c = 0; c < 64; ++c) pr[c] = 0.f;
      }
      if (i == n_q - 1) {
        // query rows past q_len (lse = -inf): MUFU gives an exact 0, the polynomial 2^-125 — force zeros so that rows
        // of another varlen segment receive exact zeros in dQ
        const int qv = q_len - i * 128 - half * 64;
        if (qv < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= qv) pr[c] = 0.f;
        }
      }
      {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = pack_bf16x2(pr[2 * c], pr[2 * c + 1]);
        tmem_st_x32(p_addr, pk);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_ready);
      if (tr) trace_mark(p.trace, 0, i, 2);

      // ---- dS^T = P^T o (dP^T - delta) ----
      mbar_wait(dp_full, i & 1, BT_DP_FULL);
      tc_fence_after();
      if (tr) trace_mark(p.trace, 0, i, 3);
      uint32_t dsp[32];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t du[32];
        tmem_ld_x32(dp_addr + c0, du);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 d4 = *reinterpret_cast<const float4*>(st + 128 + c0 + c);  // -delta
          const float2 a01 = __fmul2_rn(make_float2(pr[c0 + c + 0], pr[c0 + c + 1]),
                                        __fadd2_rn(make_float2(__uint_as_float(du[c + 0]), __uint_as_float(du[c + 1])),
                                                   make_float2(d4.x, d4.y)));
          const float2 a23 = __fmul2_rn(make_float2(pr[c0 + c + 2], pr[c0 + c + 3]),
                                        __fadd2_rn(make_float2(__uint_as_float(du[c + 2]), __uint_as_float(du[c + 3])),
                                                   make_float2(d4.z, d4.w)));
          dsp[(c0 + c) >> 1] = pack_bf16x2(a01.x, a01.y);
          dsp[((c0 + c) >> 1) + 1] = pack_bf16x2(a23.x, a23.y);
        }
      }
      // the dS tile of the previous iteration must have been consumed by MMA4/MMA5
      if (tr) trace_mark(p.trace, 0, i, 4);
      if (i > 0) mbar_wait(ds_free, (i - 1) & 1, BT_DS_FREE);
      if (tr) trace_mark(p.trace, 0, i, 5);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<uint4*>(ds_row + ((ch ^ sw) << 4)) =
            make_uint4(dsp[4 * ch], dsp[4 * ch + 1], dsp[4 * ch + 2], dsp[4 * ch + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
      if (tr) trace_mark(p.trace, 0, i, 6);
    }

    // ---- epilogue: dV, dK (this warpgroup's half of the head dim) ---------------------------------
    mbar_wait(dkv_full, 0, BT_DKV_FULL);
    mbar_wait(dkv_full + 1, 0, BT_DKV_FULL);
    tc_fence_after();
    constexpr int HC = D / 2;
    const bool store_row = kv0 + krow < k_rows_total;
    __nv_bfloat16* dvp = p.dv + bk * p.dv_sb + static_cast<int64_t>(k_base + kv0 + krow) * p.dv_sl + h * p.dv_sh + half * HC;
    __nv_bfloat16* dkp = p.dk + bk * p.dk_sb + static_cast<int64_t>(k_base + kv0 + krow) * p.dk_sl + h * p.dk_sh + half * HC;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t addr = tmem + lane_addr + (which == 0 ? C::T_DV : C::T_DK) + half * HC;
      const float mul = which == 0 ? 1.f : p.scale;
      __nv_bfloat16* dst = which == 0 ? dvp : dkp;
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(addr + c0, r);
        tc_wait_ld();
        if (store_row) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint4 w;
            // rows of masked (padding) keys hold exact zeros because P == 0 there
            w.x = pack_bf16x2(__uint_as_float(r[c + 0]) * mul, __uint_as_float(r[c + 1]) * mul);
            w.y = pack_bf16x2(__uint_as_float(r[c + 2]) * mul, __uint_as_float(r[c + 3]) * mul);
            w.z = pack_bf16x2(__uint_as_float(r[c + 4]) * mul, __uint_as_float(r[c + 5]) * mul);
            w.w = pack_bf16x2(__uint_as_float(r[c + 6]) * mul, __uint_as_float(r[c + 7]) * mul);
            *reinterpret_cast<uint4*>(dst + c0 + c) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// delta[b,h,row] = sum_d dO[row,d] * O[row,d]   (one warp per (row, head); 4 B/elem of traffic)
template <int D>
__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ o,
                                      float* __restrict__ delta, int64_t do_sb, int64_t do_sl, int64_t do_sh,
                                      int64_t o_sb, int64_t o_sl, int64_t o_sh, int B, int L, int H) {
  const int64_t gw = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total = static_cast<int64_t>(B) * L * H;
  if (gw >= total) return;
  const int h = static_cast<int>(gw % H);
  const int64_t bl = gw / H;
  const int l = static_cast<int>(bl % L), b = static_cast<int>(bl / L);
  constexpr int PER = D / 32;  // 2 or 4 elements per lane
  const __nv_bfloat16* a = dout + b * do_sb + static_cast<int64_t>(l) * do_sl + h * do_sh + lane * PER;
  const __nv_bfloat16* c = o + b * o_sb + static_cast<int64_t>(l) * o_sl + h * o_sh + lane * PER;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < PER; i += 2) {
    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + i));
    const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(c + i));
    acc += x.x * y.x + x.y * y.y;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) delta[(static_cast<int64_t>(b) * H + h) * L + l] = acc;
}

// dq = bf16(dq_acc * scale)
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int64_t sb,
                                           int64_t sl, int64_t sh, int B, int L, int H, int D, float scale) {
  const int64_t idx = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t total = static_cast<int64_t>(B) * L * H * D;
  if (idx >= total) return;
  const int d = static_cast<int>(idx % D);
  const int64_t r = idx / D;
  const int h = static_cast<int>(r % H);
  const int64_t bl = r / H;
  const int l = static_cast<int>(bl % L), b = static_cast<int>(bl / L);
  const float4 x = *reinterpret_cast<const float4*>(acc + idx);
  const float4 y = *reinterpret_cast<const float4*>(acc + idx + 4);
  uint4 w;
  w.x = pack_bf16x2(x.x * scale, x.y * scale);
  w.y = pack_bf16x2(x.z * scale, x.w * scale);
  w.z = pack_bf16x2(y.x * scale, y.y * scale);
  w.w = pack_bf16x2(y.z * scale, y.w * scale);
  *reinterpret_cast<uint4*>(dq + b * sb + static_cast<int64_t>(l) * sl + h * sh + d) = w;
}

template <int D, bool DIRECT>
cudaError_t launch_bwd_one(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                           const CUtensorMap& tm_do, const CUtensorMap& tm_dq, float* dq_acc, int Lq_total,
                           const AttnBwdParams& p, cudaStream_t stream) {
  using C = BwdCfg<D>;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<D, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::BYTES);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((p.seq.Lk + 127) / 128, p.seq.H, p.seq.nprob);
  attn_bwd_kernel<D, DIRECT><<<grid, C::THREADS, C::BYTES, stream>>>(tm_q, tm_k, tm_v, tm_do, tm_dq, dq_acc, Lq_total, p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attn_bwd(int D, const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                            const CUtensorMap& tm_do, const CUtensorMap& tm_dq, float* dq_acc, int Lq_total,
                            const AttnBwdParams& p, cudaStream_t stream) {
  const bool direct = p.dq_direct != nullptr;
  if (D == 128 && direct) return launch_bwd_one<128, true>(tm_q, tm_k, tm_v, tm_do, tm_dq, dq_acc, Lq_total, p, stream);
  if (D == 128) return launch_bwd_one<128, false>(tm_q, tm_k, tm_v, tm_do, tm_dq, dq_acc, Lq_total, p, stream);
  if (D == 64 && direct) return launch_bwd_one<64, true>(tm_q, tm_k, tm_v, tm_do, tm_dq, dq_acc, Lq_total, p, stream);
  if (D == 64) return launch_bwd_one<64, false>(tm_q, tm_k, tm_v, tm_do, tm_dq, dq_acc, Lq_total, p, stream);
  return cudaErrorInvalidValue;
}

cudaError_t launch_attn_bwd_delta(int D, const void* dout, const void* o, float* delta, const int64_t* do_strides,
                                  const int64_t* o_strides, int B, int L, int H, cudaStream_t stream) {
  const int64_t warps = static_cast<int64_t>(B) * L * H;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((warps * 32 + threads - 1) / threads);
  auto a = static_cast<const __nv_bfloat16*>(dout);
  auto c = static_cast<const __nv_bfloat16*>(o);
  if (D == 128)
    attn_bwd_delta_kernel<128><<<blocks, threads, 0, stream>>>(a, c, delta, do_strides[0], do_strides[1], do_strides[2],
                                                               o_strides[0], o_strides[1], o_strides[2], B, L, H);
  else
    attn_bwd_delta_kernel<64><<<blocks, threads, 0, stream>>>(a, c, delta, do_strides[0], do_strides[1], do_strides[2],
                                                              o_strides[0], o_strides[1], o_strides[2], B, L, H);
  return cudaGetLastError();
}

cudaError_t launch_attn_bwd_dq_convert(const float* acc, void* dq, const int64_t* dq_strides, int B, int L, int H, int D,
                                       float scale, cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(B) * L * H * D;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total / 8 + threads - 1) / threads);
  attn_bwd_dq_convert_kernel<<<blocks, threads, 0, stream>>>(acc, static_cast<__nv_bfloat16*>(dq), dq_strides[0],
                                                             dq_strides[1], dq_strides[2], B, L, H, D, scale);
  return cudaGetLastError();
}

cudaError_t attn_bwd_set_debug_ptr(unsigned int* ptr) { return cudaMemcpyToSymbol(g_vt_dbg, &ptr, sizeof(ptr)); }

}  // namespace vt
