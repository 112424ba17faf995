// geglu.cu — gated GELU of lvdm's feed-forward, one pass:  y = x * gelu(gate)  with (x, gate) the two halves of the
// projection's output row.
//
// Replaces GEGLU.forward after its Linear (videotuna/models/lvdm/modules/attention.py:522-529: `x, gate =
// proj(x).chunk(2, -1); return x * F.gelu(gate)`), inside BasicTransformerBlock's `x + ff(norm3(x))` (:299-310). As
// separate torch kernels (gelu, mul, and their backwards) this is 22 % of the graph-captured VideoCrafter2 LoRA step
// (profiles/r2_s17_vc2_profile_channels_last.txt: GeluCUDAKernel, BinaryFunctor mul, GeluBackward): the projection output
// is the widest tensor of the UNet, (b*t*h*w, 8*C), and the unfused form reads / writes it ~10 times. Here: forward reads
// 2F and writes F elements per row; backward reads 3F and writes 2F. HBM-bound (bf16): 6 B per output element forward, 10 B
// backward. Exact (erf) GELU as F.gelu's default: erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7 in exact arithmetic, 5.3e-7 in
// float32: oracle/ref_ops.py erf_abramowitz_stegun_f32, tests/test_oracle_golden.py), one MUFU.EX2 + one
// MUFU.RCP per element.
#include <cuda_bf16.h>

#include "capi_util.h"

namespace vt {
namespace {

__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.f - p * t * __expf(-ax * ax);
  return copysignf(e, x);
}
__device__ __forceinline__ float gelu_f(float g) { return 0.5f * g * (1.f + erf_as(g * 0.70710678118654752f)); }
// d/dg gelu(g) = Phi(g) + g * phi(g)
__device__ __forceinline__ float dgelu_f(float g) {
  const float cdf = 0.5f * (1.f + erf_as(g * 0.70710678118654752f));
  return fmaf(g * 0.3989422804014327f, __expf(-0.5f * g * g), cdf);
}

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// xin: (M, 2F) rows [x | gate]; y: (M, F). One thread = one 8-element vector of the output; grid-stride.
__global__ void __launch_bounds__(256) geglu_fwd_kernel(const __nv_bfloat16* __restrict__ xin, __nv_bfloat16* __restrict__ y,
                                                        int64_t M, int F) {
  const int vpr = F >> 3;
  const int64_t total = M * vpr;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vpr;
    const int v = static_cast<int>(i - row * vpr);
    const __nv_bfloat16* src = xin + row * (2 * static_cast<int64_t>(F)) + v * 8;
    float a[8], g[8];
    ld8(src, a);
    ld8(src + F, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= gelu_f(g[k]);
    st8(y + row * F + v * 8, a);
  }
}

// dxin[:, :F] = dy * gelu(gate);  dxin[:, F:] = dy * x * gelu'(gate)
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ xin,
                                                        __nv_bfloat16* __restrict__ dxin, int64_t M, int F) {
  const int vpr = F >> 3;
  const int64_t total = M * vpr;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vpr;
    const int v = static_cast<int>(i - row * vpr);
    const int64_t off = row * (2 * static_cast<int64_t>(F)) + v * 8;
    float a[8], g[8], d[8], da[8], dg[8];
    ld8(xin + off, a);
    ld8(xin + off + F, g);
    ld8(dy + row * F + v * 8, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      da[k] = d[k] * gelu_f(g[k]);
      dg[k] = d[k] * a[k] * dgelu_f(g[k]);
    }
    st8(dxin + off, da);
    st8(dxin + off + F, dg);
  }
}

unsigned grid_for(int64_t vectors) {
  const int64_t want = (vectors + 255) / 256;
  const int64_t cap = 148 * 16;
  return static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
}  // namespace vt

using namespace vt;

extern "C" {

int vt_geglu_fwd(const void* xin, void* y, int64_t M, int F, void* stream) {
  VT_REQUIRE(xin && y, VT_ERR_NULL, "vt_geglu_fwd: NULL argument");
  VT_REQUIRE(M > 0 && F > 0 && F % 8 == 0, VT_ERR_SHAPE, "vt_geglu_fwd: M=%lld F=%d (F must be a multiple of 8)", (long long)M, F);
  VT_REQUIRE(aligned16(xin) && aligned16(y), VT_ERR_ALIGN, "tensors must be 16-byte aligned");
  geglu_fwd_kernel<<<grid_for(M * (F / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(xin), static_cast<__nv_bfloat16*>(y), M, F);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vt_geglu_bwd(const void* dy, const void* xin, void* dxin, int64_t M, int F, void* stream) {
  VT_REQUIRE(dy && xin && dxin, VT_ERR_NULL, "vt_geglu_bwd: NULL argument");
  VT_REQUIRE(M > 0 && F > 0 && F % 8 == 0, VT_ERR_SHAPE, "vt_geglu_bwd: M=%lld F=%d (F must be a multiple of 8)", (long long)M, F);
  VT_REQUIRE(aligned16(dy) && aligned16(xin) && aligned16(dxin), VT_ERR_ALIGN, "tensors must be 16-byte aligned");
  geglu_bwd_kernel<<<grid_for(M * (F / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(xin), static_cast<__nv_bfloat16*>(dxin), M, F);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
