// temporal_attn.cu — micro-attention over very short sequences (N <= 32 tokens), forward and backward.
//
// Replaces the attention core of lvdm's TemporalTransformer blocks (videotuna/models/lvdm/modules/attention.py:475-519
// reshapes (b, c, t, h, w) to (b*h*w, t, c); CrossAttention.forward :126-149 then runs softmax(q k^T * scale [mask]) v
// over t = 16 frames for tens of thousands of (position, head) pairs). The work per pair is ~64 KFLOP against 8 KB of
// q/k/v/o, so the kernel is HBM-bound: one warp owns one (sequence, head) pair, stages q, k, v (and dO) in shared
// memory with coalesced 128-bit loads, does the N x N arithmetic in fp32 registers and writes each output row as one
// coalesced store. No tensor cores: a 16 x 16 problem cannot fill a 128-row MMA tile, and the bytes dominate anyway.
//
// mask (optional): (N, N) fp32, > 0.5 = keep, shared by every pair — the causal temporal mask of attention.py:487-489;
// masked scores are filled with -FLT_MAX before the softmax exactly like masked_fill(~mask, -finfo.max) (:136-140).
#include <cfloat>
#include <cstdlib>
#include <cuda_bf16.h>

#include "capi_util.h"

namespace vt {
namespace {

constexpr int kMaxN = 32;

#ifdef VT_EXPERIMENTS  // first-generation SIMT kernel (VT_TEMPORAL_SIMT=1 in A/B builds); the product path is temporal_attn_mma.cu
constexpr int kWarps = 8;       // (sequence, head) pairs in flight per CTA (fewer when N = 32, D = 128 would not fit)

// Per-warp shared memory, sized by the ACTUAL sequence length N (not the 32-token maximum): for N = 16, D = 64 that is
// 7.4 KB (forward) / 10.6 KB (backward) instead of 25 KB, which is what bounds the number of resident warps — with the
// fixed 32-row layout only 8 warps fitted per SM and the kernel ran at 7 % of the HBM roofline, slower than the three
// torch kernels it replaces.
template <int D>
struct TemporalSmem {
  static constexpr int PITCH = D + 2;  // bf16 elements; +2 keeps row-per-lane reads conflict-free (pitch/2 odd)
  __nv_bfloat16 *q, *k, *v, *g;        // [N][PITCH]; g = dO (backward only)
  float *p, *ds;                       // [N][N + 1] probabilities [query][key]; dS (backward only)
  int pp;                              // row pitch of p / ds = N + 1
  __host__ __device__ static int bytes(int N, bool bwd) {
    const int mat = ((N * PITCH * 2 + 15) / 16) * 16, sq = ((N * (N + 1) * 4 + 15) / 16) * 16;
    return (bwd ? 4 : 3) * mat + (bwd ? 2 : 1) * sq;
  }
  __device__ TemporalSmem(uint8_t* base, int N, bool bwd) {
    const int mat = ((N * PITCH * 2 + 15) / 16) * 16, sq = ((N * (N + 1) * 4 + 15) / 16) * 16;
    q = reinterpret_cast<__nv_bfloat16*>(base);
    k = reinterpret_cast<__nv_bfloat16*>(base + mat);
    v = reinterpret_cast<__nv_bfloat16*>(base + 2 * mat);
    g = reinterpret_cast<__nv_bfloat16*>(base + 3 * mat);
    uint8_t* f = base + (bwd ? 4 : 3) * mat;
    p = reinterpret_cast<float*>(f);
    ds = reinterpret_cast<float*>(f + sq);
    pp = N + 1;
  }
};

struct TemporalArgs {
  const __nv_bfloat16 *q, *k, *v, *dout;
  __nv_bfloat16 *o, *dq, *dk, *dv;
  const float* mask;
  int64_t q_s[3], k_s[3], v_s[3], o_s[3], g_s[3];  // (b, n, h) element strides; gradients are contiguous (B, N, H, D)
  int B, N, H;
  float scale;
};

// Stage one (N x D) bf16 matrix with row stride `sn` into padded shared memory; 128-bit global loads.
template <int D>
__device__ __forceinline__ void stage(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t sn, int N, int lane) {
  constexpr int PITCH = TemporalSmem<D>::PITCH;
  constexpr int VPR = D / 8;  // 16-byte vectors per row
  for (int idx = lane; idx < N * VPR; idx += 32) {
    const int r = idx / VPR, c = (idx % VPR) * 8;
    const uint4 w = *reinterpret_cast<const uint4*>(src + r * sn + c);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + r * PITCH + c);  // PITCH is even: 4-byte aligned
    d32[0] = w.x; d32[1] = w.y; d32[2] = w.z; d32[3] = w.w;
  }
}

__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

__device__ __forceinline__ float group_max(float x, int width) {
  for (int o = width >> 1; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ float group_sum(float x, int width) {
  for (int o = width >> 1; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// Scores and probabilities for the staged q, k. N <= 16: the two half-warps take alternate query rows (lane % 16 owns
// key lane % 16, its K row held in registers); N > 16: lane j owns key j. Writes P[i][j] to sm.p.
template <int D>
__device__ __forceinline__ void softmax_rows(TemporalSmem<D>& sm, const TemporalArgs& a, int lane) {
  constexpr int PITCH = TemporalSmem<D>::PITCH;
  const int N = a.N;
  const bool split = N <= 16;
  const int width = split ? 16 : 32;
  const int key = split ? (lane & 15) : lane, half = split ? (lane >> 4) : 0, step = split ? 2 : 1;
  const bool key_ok = key < N;
  // D = 64: the lane's K row lives in registers (64 floats) for all query rows; D = 128 would need 128 and re-reads it
  constexpr bool KREG = (D == 64);
  const __nv_bfloat16* krow = sm.k + (key_ok ? key : 0) * PITCH;
  float2 kreg[KREG ? D / 2 : 1];
  if constexpr (KREG) {
#pragma unroll
    for (int d = 0; d < D; d += 2) kreg[d / 2] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(krow + d));
  }
  for (int i0 = 0; i0 < N; i0 += step) {
    const int i = i0 + half;
    const bool row_ok = i < N;
    const __nv_bfloat16* qrow = sm.q + (row_ok ? i : 0) * PITCH;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int d = 0; d < D; d += 2) {
      const float2 qq = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qrow + d));
      float2 kk;
      if constexpr (KREG) kk = kreg[d / 2];
      else kk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(krow + d));
      acc0 = fmaf(qq.x, kk.x, acc0);
      acc1 = fmaf(qq.y, kk.y, acc1);
    }
    float s = (acc0 + acc1) * a.scale;
    if (a.mask != nullptr && key_ok && row_ok && !(a.mask[i * N + key] > 0.5f)) s = -FLT_MAX;
    if (!key_ok) s = -INFINITY;
    const float m = group_max(s, width);
    const float e = key_ok ? __expf(s - m) : 0.f;
    const float l = group_sum(e, width);
    if (key_ok && row_ok) sm.p[i * sm.pp + key] = e / l;
  }
  __syncwarp();
}

// out[r][d] = sum_c coef(r, c) * mat[c][d] for the lane's D/32 consecutive dims; coef read as broadcast from smem.
template <int D, bool TRANSPOSED>
__device__ __forceinline__ void mix_rows(const float* coef, int pp, const __nv_bfloat16* mat, int N, int lane, int r,
                                         float (&out)[D / 32]) {
  constexpr int PITCH = TemporalSmem<D>::PITCH;
  constexpr int PER = D / 32;
#pragma unroll
  for (int e = 0; e < PER; ++e) out[e] = 0.f;
  for (int c = 0; c < N; ++c) {
    const float w = TRANSPOSED ? coef[c * pp + r] : coef[r * pp + c];
    const __nv_bfloat16* mrow = mat + c * PITCH + lane * PER;
#pragma unroll
    for (int e = 0; e < PER; e += 2) {
      const float2 mm = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(mrow + e));
      out[e] = fmaf(w, mm.x, out[e]);
      out[e + 1] = fmaf(w, mm.y, out[e + 1]);
    }
  }
}

template <int D>
__device__ __forceinline__ void store_row(__nv_bfloat16* dst, const float (&val)[D / 32], int lane, float mul) {
  constexpr int PER = D / 32;
  if constexpr (PER == 2) {
    *reinterpret_cast<__nv_bfloat162*>(dst + lane * 2) = __floats2bfloat162_rn(val[0] * mul, val[1] * mul);
  } else {
    __nv_bfloat162 lo = __floats2bfloat162_rn(val[0] * mul, val[1] * mul);
    __nv_bfloat162 hi = __floats2bfloat162_rn(val[2] * mul, val[3] * mul);
    uint2 w;
    w.x = *reinterpret_cast<uint32_t*>(&lo);
    w.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + lane * 4) = w;
  }
}

template <int D, bool BWD>
__global__ void __launch_bounds__(kWarps * 32)
temporal_attn_kernel(const TemporalArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TemporalSmem<D> sm(smem_raw + warp * TemporalSmem<D>::bytes(a.N, BWD), a.N, BWD);
  const int64_t pairs = static_cast<int64_t>(a.B) * a.H;
  const int N = a.N;
  const int nwarps = blockDim.x >> 5;
  for (int64_t pair = static_cast<int64_t>(blockIdx.x) * nwarps + warp; pair < pairs;
       pair += static_cast<int64_t>(gridDim.x) * nwarps) {
    const int b = static_cast<int>(pair / a.H), h = static_cast<int>(pair % a.H);
    stage<D>(sm.q, a.q + b * a.q_s[0] + h * a.q_s[2], a.q_s[1], N, lane);
    stage<D>(sm.k, a.k + b * a.k_s[0] + h * a.k_s[2], a.k_s[1], N, lane);
    stage<D>(sm.v, a.v + b * a.v_s[0] + h * a.v_s[2], a.v_s[1], N, lane);
    if (BWD) stage<D>(sm.g, a.dout + b * a.g_s[0] + h * a.g_s[2], a.g_s[1], N, lane);
    __syncwarp();
    softmax_rows<D>(sm, a, lane);
    if (!BWD) {
      for (int i = 0; i < N; ++i) {
        float o[D / 32];
        mix_rows<D, false>(sm.p, sm.pp, sm.v, N, lane, i, o);
        store_row<D>(a.o + b * a.o_s[0] + i * a.o_s[1] + h * a.o_s[2], o, lane, 1.f);
      }
    } else {
      constexpr int PITCH = TemporalSmem<D>::PITCH;
      // dP[i][j] = dO_i . v_j (lane j); delta_i = sum_j P dP; dS = P (dP - delta)
      const bool key_ok = lane < N;
      const __nv_bfloat16* vrow = sm.v + (key_ok ? lane : 0) * PITCH;
      for (int i = 0; i < N; ++i) {
        const __nv_bfloat16* grow = sm.g + i * PITCH;
        float acc = 0.f;
#pragma unroll 8
        for (int d = 0; d < D; d += 2) {
          const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vrow + d));
          const float2 gg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(grow + d));
          acc = fmaf(gg.x, vv.x, acc);
          acc = fmaf(gg.y, vv.y, acc);
        }
        const float pij = key_ok ? sm.p[i * sm.pp + lane] : 0.f;
        const float delta = warp_sum(pij * acc);
        if (key_ok) sm.ds[i * sm.pp + lane] = pij * (acc - delta);
      }
      __syncwarp();
      const int64_t gbase = (static_cast<int64_t>(b) * N * a.H + h) * D;  // contiguous (B, N, H, D) gradients
      const int64_t gn = static_cast<int64_t>(a.H) * D;
      for (int r = 0; r < N; ++r) {
        float t[D / 32];
        mix_rows<D, true>(sm.p, sm.pp, sm.g, N, lane, r, t);    // dV_r = sum_i P[i][r] dO_i
        store_row<D>(a.dv + gbase + r * gn, t, lane, 1.f);
        mix_rows<D, true>(sm.ds, sm.pp, sm.q, N, lane, r, t);   // dK_r = scale * sum_i dS[i][r] q_i
        store_row<D>(a.dk + gbase + r * gn, t, lane, a.scale);
        mix_rows<D, false>(sm.ds, sm.pp, sm.k, N, lane, r, t);  // dQ_r = scale * sum_j dS[r][j] k_j
        store_row<D>(a.dq + gbase + r * gn, t, lane, a.scale);
      }
    }
    __syncwarp();
  }
}

template <int D, bool BWD>
cudaError_t launch(const TemporalArgs& a, cudaStream_t st) {
  const int per_warp = TemporalSmem<D>::bytes(a.N, BWD);
  int warps = kWarps;
  while (warps > 1 && warps * per_warp > 200 * 1024) warps >>= 1;
  const int bytes = warps * per_warp;
  cudaError_t e = cudaFuncSetAttribute(temporal_attn_kernel<D, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  const int64_t pairs = static_cast<int64_t>(a.B) * a.H;
  // persistent-style grid: a multiple of the 148 SMs, several CTAs per SM to cover the load latency
  int64_t blocks = (pairs + warps - 1) / warps;
  const int64_t cap = 148 * 6;
  if (blocks > cap) blocks = cap;
  temporal_attn_kernel<D, BWD><<<static_cast<unsigned>(blocks), warps * 32, bytes, st>>>(a);
  return cudaGetLastError();
}

#endif  // VT_EXPERIMENTS

int check(const void* q, const void* k, const void* v, int B, int N, int H, int D, const int64_t* const* strides,
          int nstr) {
  VT_REQUIRE(q && k && v, VT_ERR_NULL, "temporal attention: NULL tensor");
  VT_REQUIRE(D == 64 || D == 128, VT_ERR_DTYPE, "head dim %d unsupported (64 or 128)", D);
  VT_REQUIRE(B > 0 && H > 0 && N > 0 && N <= kMaxN, VT_ERR_SHAPE, "temporal attention needs 0 < N <= %d (N=%d)", kMaxN, N);
  for (int t = 0; t < nstr; ++t) {
    VT_REQUIRE(strides[t] != nullptr, VT_ERR_NULL, "temporal attention: NULL stride array");
    for (int i = 0; i < 3; ++i)
      VT_REQUIRE(strides[t][i] % 8 == 0, VT_ERR_ALIGN, "strides must be multiples of 8 elements");
  }
  return 0;
}

#ifdef VT_EXPERIMENTS
void fill(TemporalArgs& a, const int64_t* qs, const int64_t* ks, const int64_t* vs) {
  for (int i = 0; i < 3; ++i) {
    a.q_s[i] = qs[i];
    a.k_s[i] = ks[i];
    a.v_s[i] = vs[i];
  }
}
#endif

}  // namespace
}  // namespace vt

namespace vt {
cudaError_t temporal_attn_fwd_mma(const void* q, const void* k, const void* v, void* o, const float* mask,
                                  const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os, int B, int N,
                                  int H, int D, float scale, cudaStream_t st);  // temporal_attn_mma.cu
cudaError_t temporal_attn_bwd_mma(const void* dout, const void* q, const void* k, const void* v, void* dq, void* dk, void* dv,
                                  const float* mask, const int64_t* gs, const int64_t* qs, const int64_t* ks, const int64_t* vs,
                                  int B, int N, int H, int D, float scale, cudaStream_t st);
}

using namespace vt;

extern "C" int vt_temporal_attn_fwd(const void* q, const void* k, const void* v, void* o, const float* mask,
                                    const int64_t* q_strides, const int64_t* k_strides, const int64_t* v_strides,
                                    const int64_t* o_strides, int B, int N, int H, int D, float softmax_scale,
                                    void* stream) {
  const int64_t* strides[4] = {q_strides, k_strides, v_strides, o_strides};
  if (int rc = check(q, k, v, B, N, H, D, strides, 4)) return rc;
  VT_REQUIRE(o != nullptr, VT_ERR_NULL, "o is NULL");
  VT_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), VT_ERR_ALIGN, "tensors must be 16-byte aligned");
#ifdef VT_EXPERIMENTS
  if (getenv("VT_TEMPORAL_SIMT") == nullptr)
#endif
  {  // the tensor-core (mma.sync) forward
    VT_CHECK_CUDA(temporal_attn_fwd_mma(q, k, v, o, mask, q_strides, k_strides, v_strides, o_strides, B, N, H, D,
                                        softmax_scale, static_cast<cudaStream_t>(stream)));
    return 0;
  }
#ifdef VT_EXPERIMENTS
  TemporalArgs a{};
  a.q = static_cast<const __nv_bfloat16*>(q);
  a.k = static_cast<const __nv_bfloat16*>(k);
  a.v = static_cast<const __nv_bfloat16*>(v);
  a.o = static_cast<__nv_bfloat16*>(o);
  a.mask = mask;
  fill(a, q_strides, k_strides, v_strides);
  for (int i = 0; i < 3; ++i) a.o_s[i] = o_strides[i];
  a.B = B; a.N = N; a.H = H;
  a.scale = softmax_scale;
  auto st = static_cast<cudaStream_t>(stream);
  VT_CHECK_CUDA(D == 64 ? (launch<64, false>(a, st)) : (launch<128, false>(a, st)));
  return 0;
#endif
}

extern "C" int vt_temporal_attn_bwd(const void* dout, const void* q, const void* k, const void* v, void* dq, void* dk,
                                    void* dv, const float* mask, const int64_t* do_strides, const int64_t* q_strides,
                                    const int64_t* k_strides, const int64_t* v_strides, int B, int N, int H, int D,
                                    float softmax_scale, void* stream) {
  const int64_t* strides[4] = {q_strides, k_strides, v_strides, do_strides};
  if (int rc = check(q, k, v, B, N, H, D, strides, 4)) return rc;
  VT_REQUIRE(dout && dq && dk && dv, VT_ERR_NULL, "temporal attention backward: NULL tensor");
  VT_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(dout) && aligned16(dq) && aligned16(dk) && aligned16(dv),
             VT_ERR_ALIGN, "tensors must be 16-byte aligned");
#ifdef VT_EXPERIMENTS
  if (getenv("VT_TEMPORAL_SIMT") == nullptr)
#endif
  {  // the tensor-core (mma.sync) backward
    VT_CHECK_CUDA(temporal_attn_bwd_mma(dout, q, k, v, dq, dk, dv, mask, do_strides, q_strides, k_strides, v_strides, B, N, H,
                                        D, softmax_scale, static_cast<cudaStream_t>(stream)));
    return 0;
  }
#ifdef VT_EXPERIMENTS
  TemporalArgs a{};
  a.q = static_cast<const __nv_bfloat16*>(q);
  a.k = static_cast<const __nv_bfloat16*>(k);
  a.v = static_cast<const __nv_bfloat16*>(v);
  a.dout = static_cast<const __nv_bfloat16*>(dout);
  a.dq = static_cast<__nv_bfloat16*>(dq);
  a.dk = static_cast<__nv_bfloat16*>(dk);
  a.dv = static_cast<__nv_bfloat16*>(dv);
  a.mask = mask;
  fill(a, q_strides, k_strides, v_strides);
  for (int i = 0; i < 3; ++i) a.g_s[i] = do_strides[i];
  a.B = B; a.N = N; a.H = H;
  a.scale = softmax_scale;
  auto st = static_cast<cudaStream_t>(stream);
  VT_CHECK_CUDA(D == 64 ? (launch<64, true>(a, st)) : (launch<128, true>(a, st)));
  return 0;
#endif
}
