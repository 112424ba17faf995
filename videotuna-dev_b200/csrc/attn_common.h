// attn_common.h — host/device parameter blocks shared by the attention kernels and the C-ABI.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace vt {

// Logical tensors are (B, L, H, D) with arbitrary element strides for B, L, H and unit stride for D.
// Two sequence modes:
//   fixed : problem p = batch p; q rows [0,Lq), k rows [0, seqlens_k ? seqlens_k[p] : Lk)
//   varlen: B == 1, tensors are packed (total, H, D); problem p = segment p with rows
//           [cu_q[p], cu_q[p+1]) and keys [cu_k[p], cu_k[p+1]).  (reference: flash_attn_varlen_func as called
//           from hunyuan attenion.py:108 and wan attention.py:113)
struct AttnSeq {
  const int32_t* cu_q;       // device, nullable
  const int32_t* cu_k;       // device, nullable
  const int32_t* seqlens_k;  // device, nullable (fixed mode only)
  int Lq, Lk;                // fixed: per-batch lengths; varlen: max_seqlen_q / max_seqlen_k
  int H;
  int nprob;                 // batches (fixed) or segments (varlen)
};

// Optional timeline trace (debug): CTA (0,0,0) stores clock64() at protocol points of its first kTraceIters loop
// iterations into trace[(role * kTraceIters + iter) * kTraceSlots + slot]. nullptr (the default) disables it.
constexpr int kTraceIters = 64, kTraceSlots = 8, kTraceRoles = 4;

struct AttnFwdParams {
  AttnSeq seq;
  long long* trace;
  __nv_bfloat16* o;
  float* lse;                // (B, H, Lq_total) fp32, natural-log units
  int64_t o_sb, o_sl, o_sh;  // element strides of o
  int64_t lse_sb, lse_sh;    // element strides of lse (row stride 1)
  float scale;               // softmax scale
  float scale_log2;          // scale * log2(e)
  // Optional fused "head -> sequence" exchange of Ulysses sequence parallelism (sc_n > 0, fixed mode, B == 1): besides
  // the local o, query row l < sc_n * sc_rpr of head h is also stored to rank l / sc_rpr at
  //   sc_base[l / sc_rpr] + (l % sc_rpr) * sc_sl + h * sc_sh
  // through its peer-mapped pointer (NVLink), so the output all-to-all needs no separate kernel or pack copy.
  __nv_bfloat16* sc_base[8];
  int sc_n, sc_rpr;
  int64_t sc_sl, sc_sh;
};

struct AttnBwdParams {
  AttnSeq seq;
  long long* trace;
  const __nv_bfloat16* q;     // for the cp.async (LSU) tile loads of the backward producer
  const __nv_bfloat16* dout;
  int64_t q_sb, q_sl, q_sh;
  int64_t do_sb, do_sl, do_sh;
  const float* lse;          // from forward
  const float* delta;        // rowsum(dO * O), same layout as lse
  int64_t lse_sb, lse_sh;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  int64_t dk_sb, dk_sl, dk_sh;
  int64_t dv_sb, dv_sl, dv_sh;
  float scale;
  float scale_log2;
  // Single key tile (fixed mode, Lk <= 128: lvdm cross-attention on 77 text tokens, the 40-token level): a CTA's dQ tile is
  // the whole gradient, so the drain warps write it straight to dq as scaled bf16 — no fp32 accumulator memset, no TMA
  // reduce-add, no convert pass. NULL = accumulate (the general path).
  __nv_bfloat16* dq_direct;
  int64_t dq_sb, dq_sl, dq_sh;
};

// Host-side launchers implemented in the kernel translation units. Return cudaError_t of the launch.
cudaError_t launch_attn_fwd(int D, const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                            const AttnFwdParams& p, int q_tiles_hint, cudaStream_t stream);
cudaError_t launch_attn_fwd_alt(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                                const AttnFwdParams& p, cudaStream_t stream);  // head dim 128 only
cudaError_t launch_attn_bwd(int D, const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v,
                            const CUtensorMap& tm_do, const CUtensorMap& tm_dq, float* dq_acc, int Lq_total,
                            const AttnBwdParams& p, cudaStream_t stream);
cudaError_t launch_attn_bwd_delta(int D, const void* dout, const void* o, float* delta, const int64_t* do_strides,
                                  const int64_t* o_strides, int B, int L, int H, cudaStream_t stream);
cudaError_t launch_attn_bwd_dq_convert(const float* acc, void* dq, const int64_t* dq_strides, int B, int L, int H, int D,
                                       float scale, cudaStream_t stream);
long long* debug_trace_ptr();  // capi.cu: device buffer set through vt_debug_set_trace, or nullptr
cudaError_t attn_fwd_set_debug_ptr(unsigned int* p);
cudaError_t attn_fwd_alt_set_debug_ptr(unsigned int* p);
cudaError_t attn_bwd_set_debug_ptr(unsigned int* p);

}  // namespace vt
