/* b200vt.h — C ABI of libb200vt.so: the B200 (sm_100a) attention hot path for VideoTuna-style video denoisers.
 *
 * VideoTuna has no C FFI for this path; its "plugin interface" is a set of Python callables.
 * Every entry point below names the reference callable whose arithmetic it replaces. The Python host layer
 * (videotuna-dev_b200/ops.py) binds these with ctypes and registers them as torch.library ops.
 *
 * Conventions
 *  - All functions return 0 on success or a negative VT_ERR_* code; they never throw and never exit.
 *    vt_last_error() returns a thread-local human-readable message for the last failure on this thread.
 *  - The caller owns every buffer. Pointers are raw device pointers (16-byte aligned) unless stated otherwise.
 *  - `stream` is a cudaStream_t passed as void*. Nothing synchronises the host.
 *  - bf16 = __nv_bfloat16 bit pattern. Strides are in ELEMENTS. "(b,l,h)" stride arrays have 3 entries:
 *    batch, sequence position, head; the head-dim stride is always 1.
 */
#ifndef B200VT_H_
#define B200VT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VT_OK 0
#define VT_ERR_SHAPE (-1)       /* unsupported or inconsistent shape */
#define VT_ERR_DTYPE (-2)       /* unsupported dtype / head dim */
#define VT_ERR_ALIGN (-3)       /* pointer or stride not 16-byte aligned */
#define VT_ERR_CUDA (-4)        /* CUDA runtime/driver error; see vt_last_error */
#define VT_ERR_NULL (-5)        /* required pointer is NULL */
#define VT_ERR_UNSUPPORTED (-6) /* feature not available (e.g. device is not sm_100) */

int vt_version(void);
/* Copies the calling thread's last error message (NUL-terminated) into buf; returns its length. */
int vt_last_error(char* buf, size_t n);
/* Optional: create per-device state eagerly. Returns VT_ERR_UNSUPPORTED if the device is not compute 10.x. */
int vt_init(int device);
/* Reads the watchdog words written by a kernel that trapped on a stuck barrier: out[0..3] =
 * {call-site tag, blockIdx.x, blockIdx.y | blockIdx.z<<16, threadIdx.x}; all zero if none fired. */
int vt_debug_watchdog(uint32_t out[4]);

/* Strided host<->device copy on `stream` (cudaMemcpy2DAsync): `rows` rows of `width_bytes`, pitches in bytes. Used by
 * the host-buffer attention entry point to move one head group of a pinned (B, L, H, D) tensor per DMA. The host
 * buffer must be pinned for the copy to be asynchronous. to_device: 1 = host->device, 0 = device->host. */
int vt_memcpy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t rows,
                      int to_device, void* stream);

/* Debug: when device_buf is non-NULL, CTA (0,0,0) of the attention kernels stores clock64() timestamps of its first 64
 * loop iterations into it (int64[4 roles][64 iterations][8 slots]; tools/trace_timeline.py decodes them). NULL = off. */
int vt_debug_set_trace(long long* device_buf);

/* Optional per-kernel device timing (bench.py's roofline leg). While enabled, the attention entry points bracket each
 * of their kernel launches with CUDA events recorded on the caller's stream. vt_profile_read() waits for the recorded
 * events, adds their elapsed times into per-kernel totals and returns the totals accumulated since the last
 * vt_profile_enable(1). Kernel ids: */
#define VT_K_ATTN_FWD 0
#define VT_K_ATTN_BWD 1        /* the 5-GEMM backward kernel */
#define VT_K_ATTN_BWD_DELTA 2  /* rowsum(dO * O) */
#define VT_K_ATTN_BWD_DQ 3     /* fp32 dQ accumulator -> bf16 */
#define VT_K_COUNT 4
int vt_profile_enable(int on);
int vt_profile_read(int kernel_id, double* total_ms, int64_t* launches);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense non-causal attention forward:  O = softmax(scale * Q K^T) V,  LSE = logsumexp(scale * Q K^T) (natural log)
 * Replaces: lvdm CrossAttention.forward einsum/softmax/einsum   videotuna/models/lvdm/modules/attention.py:128-144
 *           hunyuan attention() core                            videotuna/models/hunyuan/hyvideo_t2v/modules/attenion.py:101-120
 *           wan flash_attention() core                          videotuna/models/wan/wan/modules/attention.py:96-127
 * q (B,Lq,H,D), k/v (B,Lk,H,D), o (B,Lq,H,D) bf16; lse (B,H,Lq) fp32 contiguous. D in {64,128}.
 * Sequence modes:
 *   fixed  (num_segments == 0): keys of batch b are [0, seqlens_k ? seqlens_k[b] : Lk).  (wan k_lens, attention.py:62-71)
 *   varlen (num_segments  > 0): B must be 1 and the tensors are packed (total,H,D); segment s covers rows
 *          [cu_seqlens_q[s], cu_seqlens_q[s+1]) and keys [cu_seqlens_k[s], cu_seqlens_k[s+1]); Lq/Lk are the packed
 *          totals and max_seqlen_q/k bound the longest segment.               (hunyuan get_cu_seqlens, attenion.py:34-57)
 * cu_seqlens_*, seqlens_k are DEVICE int32 pointers (nullable as described).
 * ------------------------------------------------------------------------------------------------------------- */
int vt_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                const int64_t* q_strides, const int64_t* k_strides, const int64_t* v_strides,
                const int64_t* o_strides, int B, int H, int Lq, int Lk, int D,
                const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int num_segments, int max_seqlen_q,
                int max_seqlen_k, const int32_t* seqlens_k, float softmax_scale, void* stream);

/* vt_attn_fwd (fixed mode, B == 1, D == 128) with the Ulysses "head -> sequence" exchange fused into the epilogue:
 * besides the local o / lse, query row l < n_peers * rows_per_peer of head h is also stored through NVLink to
 *   (bf16*)peer_bases[l / rows_per_peer] + (l % rows_per_peer) * peer_strides[0] + h * peer_strides[1]
 * peer_bases: HOST array of n_peers (<= 8) peer-mapped device pointers (symmetric memory; entry r is rank r's
 * sequence-sharded output buffer, already offset to this rank's head slot). Rows past n_peers * rows_per_peer
 * (replicated text tokens) stay local. Replaces xfuser's output all-to-all as called from hunyuan parallel_attention
 * (attenion.py:169-180) and wan usp_attn_forward (xdit_context_parallel.py:179-186). The caller synchronises the ranks
 * before the launch (destination buffers free) and after it (all stores visible). */
int vt_attn_fwd_scatter(const void* q, const void* k, const void* v, void* o, float* lse, const int64_t* q_strides,
                        const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides, int H, int Lq, int Lk,
                        int D, const int32_t* seqlens_k, float softmax_scale, void* const* peer_bases, int n_peers,
                        int rows_per_peer, const int64_t* peer_strides, void* stream);

/* Bytes of device workspace vt_attn_bwd needs (fp32 dQ accumulator + delta = rowsum(dO*O)). */
int64_t vt_attn_bwd_workspace_bytes(int B, int H, int Lq, int D);

/* Attention backward: given dO, recomputes P from (q,k,lse) and produces dq, dk, dv (bf16; layouts given by their own
 * stride arrays). Replaces autograd through the reference torch attention path (same call sites as vt_attn_fwd).
 * In varlen mode, rows of dk/dv/dq outside every segment are zero for dq and untouched for dk/dv; in fixed mode with
 * seqlens_k, dk/dv rows of masked keys are written as zeros. `workspace` must hold vt_attn_bwd_workspace_bytes(...)
 * bytes; it is overwritten. */
int vt_attn_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o, const float* lse,
                void* dq, void* dk, void* dv, const int64_t* do_strides, const int64_t* q_strides,
                const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides,
                const int64_t* dq_strides, const int64_t* dk_strides, const int64_t* dv_strides, int B, int H, int Lq,
                int Lk, int D, const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int num_segments,
                int max_seqlen_q, int max_seqlen_k, const int32_t* seqlens_k, float softmax_scale, void* workspace,
                int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Micro-attention over very short sequences (N <= 32), one warp per (sequence, head) pair; HBM-bound.
 * Replaces the attention core of lvdm TemporalTransformer blocks: CrossAttention.forward over t = 16 frames for
 * b*h*w sequences (videotuna/models/lvdm/modules/attention.py:126-149 as called from :475-519).
 * q, k, v, o: (B, N, H, D) bf16 with (b,n,h) element strides, D in {64,128}. mask: optional (N, N) fp32 shared by all
 * pairs, > 0.5 = keep, others filled with -FLT_MAX before the softmax (attention.py:136-140). Backward recomputes the
 * probabilities; dq, dk, dv are written as contiguous (B, N, H, D) bf16.
 * ------------------------------------------------------------------------------------------------------------- */
int vt_temporal_attn_fwd(const void* q, const void* k, const void* v, void* o, const float* mask,
                         const int64_t* q_strides, const int64_t* k_strides, const int64_t* v_strides,
                         const int64_t* o_strides, int B, int N, int H, int D, float softmax_scale, void* stream);
int vt_temporal_attn_bwd(const void* dout, const void* q, const void* k, const void* v, void* dq, void* dk, void* dv,
                         const float* mask, const int64_t* do_strides, const int64_t* q_strides,
                         const int64_t* k_strides, const int64_t* v_strides, int B, int N, int H, int D,
                         float softmax_scale, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused QK-RMSNorm + 3-D RoPE (interleaved pairs) in one pass over q or k, in place of
 *   hunyuan RMSNorm (norm_layers.py:5-59) + apply_rotary_emb (posemb_layers.py:140-188)      -> norm_mode 1
 *   wan WanRMSNorm over dim = H*D (model.py:70-86) + rope_apply (model.py:40-67)              -> norm_mode 2
 *   rope only (CogVideoX-5B style, or q/k already normalised)                                 -> norm_mode 0
 * x: (B,L,H,D) bf16 with (b,l,h) strides -> y likewise. D in {64,128}.
 * norm_mode 1: RMS over each head's D values, w is (D) fp32;  2: RMS over the token's H*D values, w is (H*D) fp32.
 * cos/sin: (L_rope, D) fp32 tables in the reference's repeat_interleave(2) form, or both NULL (no rotation):
 *   y[2i] = n[2i]*cos[2i] - n[2i+1]*sin[2i];  y[2i+1] = n[2i+1]*cos[2i+1] + n[2i]*sin[2i+1]
 * Tokens l >= L_rope are not rotated (text tokens). rstd_out (nullable): (B,L,H) for mode 1, (B,L) for mode 2.
 * Backward accumulates dw into dw_accum (fp32, caller zeroes; nullable).
 * ------------------------------------------------------------------------------------------------------------- */
int vt_qk_rmsnorm_rope_fwd(const void* x, void* y, float* rstd_out, const float* w, const float* cos, const float* sin,
                           const int64_t* x_strides, const int64_t* y_strides, int B, int L, int H, int D, int L_rope,
                           int norm_mode, float eps, void* stream);
int vt_qk_rmsnorm_rope_bwd(const void* dy, const void* x, const float* rstd, void* dx, float* dw_accum, const float* w,
                           const float* cos, const float* sin, const int64_t* dy_strides, const int64_t* x_strides,
                           const int64_t* dx_strides, int B, int L, int H, int D, int L_rope, int norm_mode,
                           void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * LayerNorm (optional affine) fused with adaLN modulate:  y = LN(x) * (1 + scale[b]) + shift[b]
 * Replaces hunyuan  modulate(norm(x), shift, scale)   (modulate_layers.py:31-49 with nn.LayerNorm, models.py:161-164)
 *          wan      norm1(x).float() * (1 + e1) + e0  (wan/modules/model.py:294-296)
 *          lvdm     nn.LayerNorm(dim) (affine, no modulate: scale == shift == NULL)  (attention.py:299-310)
 * x: (B,L,C) contiguous rows, bf16 (x_dtype 0) or fp32 (x_dtype 1: Wan keeps its residual stream in fp32 and feeds the
 * modulated activation to bf16 Linears, model.py:294-296); y: (B,L,C) bf16; gamma/beta: (C) fp32 or NULL;
 * scale/shift: (B,C) fp32 or NULL. mean/rstd: (B*L) fp32 saved for backward.
 * ------------------------------------------------------------------------------------------------------------- */
int vt_ln_modulate_fwd(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta,
                       const float* scale, const float* shift, int B, int L, int C, float eps, int x_dtype,
                       void* stream);
/* dy bf16 -> dx in x's dtype; and fp32 accumulators (atomically added, caller zeroes): dgamma,dbeta (C);
 * dscale,dshift (B,C). Nullable. */
int vt_ln_modulate_bwd(const void* dy, const void* x, const float* mean, const float* rstd, void* dx,
                       const float* gamma, const float* beta, const float* scale, float* dgamma, float* dbeta,
                       float* dscale, float* dshift, int B, int L, int C, int x_dtype, void* stream);

/* Gated residual: y = x + branch * gate[b]   (hunyuan apply_gate, modulate_layers.py:52-68 & models.py:231;
 * wan x + y * e2, model.py:298). x, y: (B,L,C) bf16 (x_dtype 0) or fp32 (x_dtype 1, Wan's fp32 residual stream);
 * branch: (B,L,C) bf16; gate: (B,C) fp32 or NULL (plain add). */
int vt_gate_residual_fwd(const void* x, const void* branch, void* y, const float* gate, int B, int L, int C,
                         int x_dtype, void* stream);
/* dbranch (bf16) = dy * gate, dy in x's dtype; dgate (B,C) fp32 accumulated atomically (caller zeroes); dx = dy is the
 * caller's alias. */
int vt_gate_residual_bwd(const void* dy, const void* branch, void* dbranch, const float* gate, float* dgate, int B,
                         int L, int C, int x_dtype, void* stream);

/* GroupNorm(G) [+ SiLU] on NCHW-like tensors x: (N, C, S) bf16 or fp32 (S = product of spatial dims), fp32 statistics.
 * Replaces lvdm normalization()/GroupNormSpecific + nn.SiLU (lvdm/modules/utils.py:192-203, openaimodel3d.py:229-255)
 * and SpatialTransformer/TemporalTransformer.norm (attention.py:376-392,475-519). dtype: 0 = bf16, 1 = fp32. */
int vt_groupnorm_silu_fwd(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta,
                          int N, int C, int S, int G, float eps, int apply_silu, int dtype, void* stream);
int vt_groupnorm_silu_bwd(const void* dy, const void* x, const float* mean, const float* rstd, void* dx,
                          const float* gamma, const float* beta, float* dgamma, float* dbeta, int N, int C, int S, int G,
                          int apply_silu, int dtype, void* stream);

/* The same GroupNorm(G) [+ SiLU] for CHANNELS-LAST activations: x, y, dy, dx are (N, S, C) in memory (torch.channels_last /
 * channels_last_3d views of (N, C, *spatial) tensors), the layout the tensor-core convolutions run in and in which lvdm's
 * `b c h w -> b (h w) c` (attention.py:381) is a free view. Same reference callables as vt_groupnorm_silu_fwd/bwd.
 * C % 8 == 0 (bf16; C <= 4096) or C % 4 == 0 (fp32; C <= 2048). workspace: vt_groupnorm_nhwc_workspace_bytes(N, G) bytes,
 * overwritten. dgamma / dbeta accumulate atomically (caller zeroes; nullable).
 * addend (nullable): fp32 per-channel term added to x before the normalisation, GroupNorm(x + e): e[n * addend_stride + c]
 * with addend_stride = C (per sample: ResBlock's `h + emb_out`, openaimodel3d.py:247-255, plus the preceding convolution's
 * bias) or 0 (one vector for every sample: a convolution bias, TemporalConvBlock openaimodel3d.py:258-310). It is folded
 * into the per-channel constants / the statistics' epilogue: the separate broadcast-add passes over the activation vanish.
 * dx is the gradient of x (= that of x + e); the addend itself gets no gradient here. */
int64_t vt_groupnorm_nhwc_workspace_bytes(int N, int G);
int vt_groupnorm_silu_nhwc_fwd(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta,
                               const float* addend, int addend_stride, void* workspace, int N, int C, int S, int G,
                               float eps, int apply_silu, int dtype, void* stream);
int vt_groupnorm_silu_nhwc_bwd(const void* dy, const void* x, const float* mean, const float* rstd, void* dx,
                               const float* gamma, const float* beta, const float* addend, int addend_stride,
                               float* dgamma, float* dbeta, void* workspace, int N, int C, int S, int G, int apply_silu,
                               int dtype, void* stream);

/* Gated GELU of lvdm's feed-forward in one pass: y[m, f] = xin[m, f] * gelu(xin[m, F + f]) (exact erf GELU), xin (M, 2F),
 * y (M, F) bf16 contiguous, F % 8 == 0. Replaces GEGLU.forward after its Linear (lvdm/modules/attention.py:522-529).
 * Backward: dxin (M, 2F) = [dy * gelu(gate) | dy * x * gelu'(gate)]. */
int vt_geglu_fwd(const void* xin, void* y, int64_t M, int F, void* stream);
int vt_geglu_bwd(const void* dy, const void* xin, void* dxin, int64_t M, int F, void* stream);

#ifdef VT_EXPERIMENTS
/* ---------------------------------------------------------------------------------------------------------------
 * Not part of the product library: the hooks below exist only in builds made with -DVT_EXPERIMENTS
 * (tools/build_variant.sh <out.so> -DVT_EXPERIMENTS; point the Python side at it with B200VT_LIB=<out.so>), together
 * with the earlier kernel variants those builds keep selectable for A/B measurements (VT_FWD_KERNEL=pp|db,
 * VT_TEMPORAL_SIMT=1, VT_GN_REG=1).
 * ------------------------------------------------------------------------------------------------------------- */
/* Self-test hook used by tests/: one 128x128x128 bf16 GEMM tile through TMA + tcgen05 with selectable operand
 * sources (see csrc/umma_probe.cu). Not part of the product path. */
int vt_umma_probe(const void* a, const void* b, float* d, int a_mode, int b_mode, int n, uint32_t a_lbo, uint32_t a_sbo,
                  uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep, void* stream);

/* Microbenchmark hook (tools/umma_rate.py): cycles for `iters` groups of eight 128 x n x 16 bf16 tcgen05.mma on `blocks`
 * CTAs, per operand sourcing (see csrc/umma_rate.cu), with `noise_warps` warps streaming shared-memory stores.
 * cycles_out: device int64[blocks]. Not part of the product path. */
int vt_umma_rate(int mode, int n, int iters, int noise_warps, int blocks, long long* cycles_out, void* stream);
/* Same, for the fp32 TMA reduce-add path (16 KB boxes into acc, float[n_tiles*128][128]); depth = bulk groups in flight,
 * spread = 0: all CTAs hit the same rows at once, 1: each CTA starts at its own tile. */
int vt_tma_reduce_rate(float* acc, int n_tiles, int iters, int depth, int spread, int blocks, long long* cycles_out,
                       void* stream);

/* Same reduction stream with a second thread streaming 16 KB TMA loads (load_mode != 0) from src_bf16
 * (bf16[n_tiles*128][128]); loads_out counts the boxes loaded meanwhile. */
int vt_tma_mixed_rate(float* acc, const void* src_bf16, int n_tiles, int iters, int load_mode, int blocks,
                      long long* cycles_out, long long* loads_out, void* stream);

#endif /* VT_EXPERIMENTS */

#ifdef __cplusplus
}
#endif
#endif /* B200VT_H_ */
