// groupnorm_nhwc.cu — GroupNorm(G) [+ SiLU] on CHANNELS-LAST tensors: x is (N, S, C) in memory (S = pixels: h*w, or
// t*h*w for the 5-D temporal input), the layout cuDNN's tensor-core convolutions run in and the layout in which
// `b c h w -> b (h w) c` (lvdm SpatialTransformer, attention.py:381) is a free view.
//
// Replaces lvdm normalization() / GroupNormSpecific + nn.SiLU (lvdm/modules/utils.py:192-203, openaimodel3d.py:229-255,
// 258-310) and SpatialTransformer / TemporalTransformer.norm (attention.py:376-392, 475-519) when the activation arrives
// channels-last; statistics in fp32. Why a second GroupNorm: with NCHW activations every 3x3 convolution of the
// VideoCrafter2 UNet is bracketed by cuDNN nchwToNhwc / nhwcToNchw conversion kernels (10.8 % of the LoRA step,
// profiles/r2_s17_vc2_profile_nchw.txt) and the transformers' layout changes are copies; channels-last end to end removes
// both, but then a group is no longer a contiguous slab — it is `cpg` adjacent channels of every pixel row — so the
// bulk-copy / cluster kernels of groupnorm.cu do not apply.
//
// Two streaming sweeps per direction (the second one hits L2: the tensors are 26-105 MB against 126 MB of L2):
//   forward   stats: per-thread per-channel sum / sum of squares over the CTA's rows -> shared-memory per-channel bins ->
//             per-CTA per-group partials in a (N, chunks, G, 2) workspace (plain stores: no memset launch, no global
//             atomics, deterministic), summed over the chunks by every CTA of the second sweep;  apply: y = silu(x * a[c] + d[c]) with a = rstd*gamma,
//             d = beta - mean*a precomputed per CTA in shared memory.
//   backward  sums: gz = dy * silu'(z) recomputed from x; per-channel sum gz, sum gz*xhat -> dbeta / dgamma atomics and
//             per-group gamma-weighted sums;  apply: dx = gz * (rstd*gamma[c]) + x * c2[g] + c0[g] (gz recomputed).
// A thread owns one 16-byte vector column (8 bf16 / 4 fp32 channels) and walks down the rows, so its accumulators live in
// registers; blockDim is the largest multiple of C/VEC that fits 512 threads.
#include <cuda_bf16.h>

#include "capi_util.h"

namespace vt {
namespace {

template <typename T>
struct V;
template <>
struct V<__nv_bfloat16> {
  static constexpr int VEC = 8;
  __device__ static void load(const __nv_bfloat16* p, float* f) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct V<float> {
  static constexpr int VEC = 4;
  __device__ static void load(const float* p, float* f) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static void store(float* p, const float* f) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
};

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// silu(z) = h + h tanh(h), h = z/2; silu'(z) = s (1 + z (1 - s)), s = (1 + tanh(h)) / 2   (one MUFU op per element)
__device__ __forceinline__ float silu_fast(float z) {
  const float h = 0.5f * z;
  return fmaf(h, tanh_fast(h), h);
}
__device__ __forceinline__ float dsilu_fast(float z) {
  const float s = fmaf(0.5f, tanh_fast(0.5f * z), 0.5f);
  return s * fmaf(z, 1.f - s, 1.f);
}

struct RowSplit {
  int vpr, rpp, row0, row1, v, r;  // vectors per row, rows per pass, this CTA's row range, this thread's column / row offset
  bool active;
};
template <int VEC>
__device__ __forceinline__ RowSplit row_split(int C, int S, int rows_per_cta) {
  RowSplit rs;
  rs.vpr = C / VEC;
  rs.rpp = blockDim.x / rs.vpr;
  rs.row0 = blockIdx.x * rows_per_cta;
  rs.row1 = min(S, rs.row0 + rows_per_cta);
  rs.v = threadIdx.x % rs.vpr;
  rs.r = threadIdx.x / rs.vpr;
  rs.active = rs.r < rs.rpp;
  return rs;
}

// ---- forward ---------------------------------------------------------------------------------------------------------
// ws: float[N][chunks][G][2] per-CTA partial (sum, sum of squares): plain stores, reduced by the apply kernel's prologue —
// no memset launch, no global atomics, deterministic.
template <typename T>
__global__ void gn_nhwc_stats_kernel(const T* __restrict__ x, float* __restrict__ ws, int C, int S, int G, int rows_per_cta) {
  constexpr int VEC = V<T>::VEC;
  extern __shared__ float sm[];  // [2][C]
  const int n = blockIdx.y;
  const RowSplit rs = row_split<VEC>(C, S, rows_per_cta);
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  if (rs.active) {
    float s1[VEC], s2[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) s1[k] = s2[k] = 0.f;
    const T* base = x + (static_cast<size_t>(n) * S) * C + rs.v * VEC;
    for (int row = rs.row0 + rs.r; row < rs.row1; row += rs.rpp) {
      float f[VEC];
      V<T>::load(base + static_cast<size_t>(row) * C, f);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        s1[k] += f[k];
        s2[k] = fmaf(f[k], f[k], s2[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      atomicAdd(&sm[rs.v * VEC + k], s1[k]);
      atomicAdd(&sm[C + rs.v * VEC + k], s2[k]);
    }
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += sm[c];
      b += sm[C + c];
    }
    float* dst = ws + ((static_cast<size_t>(n) * gridDim.x + blockIdx.x) * G + g) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

// sum of the per-CTA partials of (n, g): every thread that needs a group total walks the (few dozen) chunks
__device__ __forceinline__ float2 group_total(const float* __restrict__ ws, int n, int g, int G, int chunks) {
  float a = 0.f, b = 0.f;
  const float* p = ws + (static_cast<size_t>(n) * chunks * G + g) * 2;
  for (int c = 0; c < chunks; ++c, p += 2 * G) {
    a += p[0];
    b += p[1];
  }
  return make_float2(a, b);
}

template <typename T>
__global__ void gn_nhwc_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ ws,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, int C, int S, int G,
                                     float eps, int apply_silu, int rows_per_cta) {
  constexpr int VEC = V<T>::VEC;
  extern __shared__ float sm[];  // a[C], d[C]
  const int n = blockIdx.y;
  const RowSplit rs = row_split<VEC>(C, S, rows_per_cta);
  const int cpg = C / G;
  const float inv = 1.f / (static_cast<float>(cpg) * S);
  float* gs = sm + 2 * C;  // [G][2]: mean, rstd
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float2 t = group_total(ws, n, g, G, gridDim.x);
    const float m = t.x * inv;
    gs[2 * g] = m;
    gs[2 * g + 1] = rsqrtf(fmaxf(t.y * inv - m * m, 0.f) + eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float m = gs[2 * g], r = gs[2 * g + 1];
    const float a = r * (gamma ? gamma[c] : 1.f);
    sm[c] = a;
    sm[C + c] = fmaf(-m, a, beta ? beta[c] : 0.f);
    if (blockIdx.x == 0 && c == g * cpg) {
      mean_out[n * G + g] = m;
      rstd_out[n * G + g] = r;
    }
  }
  __syncthreads();
  if (!rs.active) return;
  float a[VEC], d[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    a[k] = sm[rs.v * VEC + k];
    d[k] = sm[C + rs.v * VEC + k];
  }
  const size_t off = (static_cast<size_t>(n) * S) * C + rs.v * VEC;
  for (int row = rs.row0 + rs.r; row < rs.row1; row += rs.rpp) {
    float f[VEC];
    V<T>::load(x + off + static_cast<size_t>(row) * C, f);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float z = fmaf(f[k], a[k], d[k]);
      f[k] = apply_silu ? silu_fast(z) : z;
    }
    V<T>::store(y + off + static_cast<size_t>(row) * C, f);
  }
}

// ---- backward --------------------------------------------------------------------------------------------------------
// ws: float[N][chunks][G][2] per-CTA partial (sum gz*gamma, sum gz*gamma*xhat), plain stores; dgamma / dbeta accumulate
// atomically (caller zeroes)
template <typename T>
__global__ void gn_nhwc_bwd_sums_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float* __restrict__ ws, float* __restrict__ dgamma,
                                        float* __restrict__ dbeta, int C, int S, int G, int apply_silu, int rows_per_cta) {
  constexpr int VEC = V<T>::VEC;
  extern __shared__ float sm[];  // [2][C]: sum gz, sum gz*xhat per channel
  const int n = blockIdx.y;
  const RowSplit rs = row_split<VEC>(C, S, rows_per_cta);
  const int cpg = C / G;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  if (rs.active) {
    float ga[VEC], be[VEC], s1[VEC], s2[VEC], mu[VEC], rs_[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = rs.v * VEC + k, g = c / cpg;
      ga[k] = gamma ? gamma[c] : 1.f;
      be[k] = beta ? beta[c] : 0.f;
      mu[k] = mean[n * G + g];
      rs_[k] = rstd[n * G + g];
      s1[k] = s2[k] = 0.f;
    }
    const size_t off = (static_cast<size_t>(n) * S) * C + rs.v * VEC;
    for (int row = rs.row0 + rs.r; row < rs.row1; row += rs.rpp) {
      float fx[VEC], fg[VEC];
      V<T>::load(x + off + static_cast<size_t>(row) * C, fx);
      V<T>::load(dy + off + static_cast<size_t>(row) * C, fg);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float xh = (fx[k] - mu[k]) * rs_[k];
        float gz = fg[k];
        if (apply_silu) gz *= dsilu_fast(fmaf(xh, ga[k], be[k]));
        s1[k] += gz;
        s2[k] = fmaf(gz, xh, s2[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      atomicAdd(&sm[rs.v * VEC + k], s1[k]);
      atomicAdd(&sm[C + rs.v * VEC + k], s2[k]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (dbeta) atomicAdd(&dbeta[c], sm[c]);
    if (dgamma) atomicAdd(&dgamma[c], sm[C + c]);
  }
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float gm = gamma ? gamma[c] : 1.f;
      a = fmaf(sm[c], gm, a);
      b = fmaf(sm[C + c], gm, b);
    }
    float* dst = ws + ((static_cast<size_t>(n) * gridDim.x + blockIdx.x) * G + g) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

template <typename T>
__global__ void gn_nhwc_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, const float* __restrict__ ws, T* __restrict__ dx,
                                         int C, int S, int G, int apply_silu, int rows_per_cta) {
  constexpr int VEC = V<T>::VEC;
  extern __shared__ float sm[];  // [G][2]: group totals
  const int n = blockIdx.y;
  const RowSplit rs = row_split<VEC>(C, S, rows_per_cta);
  const int cpg = C / G;
  const float inv = 1.f / (static_cast<float>(cpg) * S);
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float2 t = group_total(ws, n, g, G, gridDim.x);
    sm[2 * g] = t.x;
    sm[2 * g + 1] = t.y;
  }
  __syncthreads();
  if (!rs.active) return;
  float ga[VEC], be[VEC], mu[VEC], rr[VEC], rg[VEC], c2[VEC], c0[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const int c = rs.v * VEC + k, g = c / cpg;
    ga[k] = gamma ? gamma[c] : 1.f;
    be[k] = beta ? beta[c] : 0.f;
    mu[k] = mean[n * G + g];
    rr[k] = rstd[n * G + g];
    const float m1 = sm[2 * g] * inv, m2 = sm[2 * g + 1] * inv;
    rg[k] = rr[k] * ga[k];
    // dx = rstd * (gz*gamma - m1 - xhat*m2) = gz*rg - rstd*m1 - (x - mu)*rstd^2*m2
    c2[k] = -rr[k] * rr[k] * m2;
    c0[k] = -rr[k] * m1 - mu[k] * c2[k];
  }
  const size_t off = (static_cast<size_t>(n) * S) * C + rs.v * VEC;
  for (int row = rs.row0 + rs.r; row < rs.row1; row += rs.rpp) {
    float fx[VEC], fg[VEC];
    V<T>::load(x + off + static_cast<size_t>(row) * C, fx);
    V<T>::load(dy + off + static_cast<size_t>(row) * C, fg);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float gz = fg[k];
      if (apply_silu) gz *= dsilu_fast(fmaf((fx[k] - mu[k]) * rr[k], ga[k], be[k]));
      fx[k] = fmaf(gz, rg[k], fmaf(fx[k], c2[k], c0[k]));
    }
    V<T>::store(dx + off + static_cast<size_t>(row) * C, fx);
  }
}

struct Plan {
  int threads, rows_per_cta, chunks;
};
template <int VEC>
bool make_plan(int N, int C, int S, Plan* p) {
  if (C % VEC != 0) return false;
  const int vpr = C / VEC;
  if (vpr > 512) return false;
  p->threads = (512 / vpr) * vpr;
  const int rpp = p->threads / vpr;
  // ~6 CTAs per SM over the whole grid, at least one pass of rows per CTA
  int chunks = (148 * 6 + N - 1) / N;
  int rows = (S + chunks - 1) / chunks;
  rows = ((rows + rpp - 1) / rpp) * rpp;
  if (rows < 4 * rpp) rows = 4 * rpp;
  p->rows_per_cta = rows;
  p->chunks = (S + rows - 1) / rows;
  return true;
}

}  // namespace
}  // namespace vt

using namespace vt;

extern "C" {

// Workspace: per-CTA partial group sums, float[N][chunks][G][2]; the bound below covers every plan make_plan() can produce.
int64_t vt_groupnorm_nhwc_workspace_bytes(int N, int G) {
  return static_cast<int64_t>(N) * ((148 * 6 + N - 1) / N + 1) * G * 2 * sizeof(float);
}

int vt_groupnorm_silu_nhwc_fwd(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta,
                               void* workspace, int N, int C, int S, int G, float eps, int apply_silu, int dtype,
                               void* stream) {
  VT_REQUIRE(x && y && mean && rstd && workspace, VT_ERR_NULL, "vt_groupnorm_silu_nhwc_fwd: NULL argument");
  VT_REQUIRE(N > 0 && C > 0 && S > 0 && G > 0 && C % G == 0 && N <= 65535, VT_ERR_SHAPE, "bad shape N=%d C=%d S=%d G=%d", N, C, S, G);
  VT_REQUIRE(dtype == 0 || dtype == 1, VT_ERR_DTYPE, "dtype %d (0=bf16, 1=fp32)", dtype);
  VT_REQUIRE(aligned16(x) && aligned16(y), VT_ERR_ALIGN, "tensors must be 16-byte aligned");
  auto st = static_cast<cudaStream_t>(stream);
  Plan p;
  const bool ok = dtype == 0 ? make_plan<8>(N, C, S, &p) : make_plan<4>(N, C, S, &p);
  VT_REQUIRE(ok, VT_ERR_UNSUPPORTED, "channels-last GroupNorm needs C %% %d == 0 and C <= %d (C=%d)", dtype == 0 ? 8 : 4,
             dtype == 0 ? 4096 : 2048, C);
  dim3 grid(p.chunks, N);
  const size_t smem = 2 * static_cast<size_t>(C) * sizeof(float);
  const size_t smem_apply = smem + 2 * static_cast<size_t>(G) * sizeof(float);
  float* ws = static_cast<float*>(workspace);
  if (dtype == 0) {
    gn_nhwc_stats_kernel<__nv_bfloat16><<<grid, p.threads, smem, st>>>(static_cast<const __nv_bfloat16*>(x), ws, C, S, G, p.rows_per_cta);
    gn_nhwc_apply_kernel<__nv_bfloat16><<<grid, p.threads, smem_apply, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y),
                                                                      ws, mean, rstd, gamma, beta, C, S, G, eps, apply_silu, p.rows_per_cta);
  } else {
    gn_nhwc_stats_kernel<float><<<grid, p.threads, smem, st>>>(static_cast<const float*>(x), ws, C, S, G, p.rows_per_cta);
    gn_nhwc_apply_kernel<float><<<grid, p.threads, smem_apply, st>>>(static_cast<const float*>(x), static_cast<float*>(y), ws, mean, rstd,
                                                              gamma, beta, C, S, G, eps, apply_silu, p.rows_per_cta);
  }
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vt_groupnorm_silu_nhwc_bwd(const void* dy, const void* x, const float* mean, const float* rstd, void* dx,
                               const float* gamma, const float* beta, float* dgamma, float* dbeta, void* workspace, int N,
                               int C, int S, int G, int apply_silu, int dtype, void* stream) {
  VT_REQUIRE(dy && x && mean && rstd && dx && workspace, VT_ERR_NULL, "vt_groupnorm_silu_nhwc_bwd: NULL argument");
  VT_REQUIRE(N > 0 && C > 0 && S > 0 && G > 0 && C % G == 0 && N <= 65535, VT_ERR_SHAPE, "bad shape N=%d C=%d S=%d G=%d", N, C, S, G);
  VT_REQUIRE(dtype == 0 || dtype == 1, VT_ERR_DTYPE, "dtype %d (0=bf16, 1=fp32)", dtype);
  VT_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx), VT_ERR_ALIGN, "tensors must be 16-byte aligned");
  auto st = static_cast<cudaStream_t>(stream);
  Plan p;
  const bool ok = dtype == 0 ? make_plan<8>(N, C, S, &p) : make_plan<4>(N, C, S, &p);
  VT_REQUIRE(ok, VT_ERR_UNSUPPORTED, "channels-last GroupNorm needs C %% %d == 0 and C <= %d (C=%d)", dtype == 0 ? 8 : 4,
             dtype == 0 ? 4096 : 2048, C);
  dim3 grid(p.chunks, N);
  const size_t smem = 2 * static_cast<size_t>(C) * sizeof(float);
  const size_t smem_g = 2 * static_cast<size_t>(G) * sizeof(float);
  float* ws = static_cast<float*>(workspace);
  if (dtype == 0) {
    auto a = static_cast<const __nv_bfloat16*>(dy);
    auto b = static_cast<const __nv_bfloat16*>(x);
    gn_nhwc_bwd_sums_kernel<__nv_bfloat16><<<grid, p.threads, smem, st>>>(a, b, mean, rstd, gamma, beta, ws, dgamma, dbeta, C, S, G,
                                                                         apply_silu, p.rows_per_cta);
    gn_nhwc_bwd_apply_kernel<__nv_bfloat16><<<grid, p.threads, smem_g, st>>>(a, b, mean, rstd, gamma, beta, ws, static_cast<__nv_bfloat16*>(dx),
                                                                       C, S, G, apply_silu, p.rows_per_cta);
  } else {
    auto a = static_cast<const float*>(dy);
    auto b = static_cast<const float*>(x);
    gn_nhwc_bwd_sums_kernel<float><<<grid, p.threads, smem, st>>>(a, b, mean, rstd, gamma, beta, ws, dgamma, dbeta, C, S, G, apply_silu,
                                                                 p.rows_per_cta);
    gn_nhwc_bwd_apply_kernel<float><<<grid, p.threads, smem_g, st>>>(a, b, mean, rstd, gamma, beta, ws, static_cast<float*>(dx), C, S, G,
                                                               apply_silu, p.rows_per_cta);
  }
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
