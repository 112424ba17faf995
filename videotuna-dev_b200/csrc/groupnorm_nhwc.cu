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
//   forward   stats: per-thread per-channel sum / sum of squares over the CTA's rows (registers) -> tree over the CTA's row
//             groups in shared memory (plain stores) -> per-CTA per-group partials in a (N, chunks, G, 2) workspace (plain
//             stores: no memset launch, no global atomics, deterministic). The partials of a sample are joined cooperatively
//             in the second sweep's prologue (<= 32 chunks) or by the finalize kernel (more: the few-sample 5-D inputs).
//             apply: y = silu(x * a[c] + d[c]) with a = rstd*gamma, d = beta - mean*a precomputed per CTA in shared memory.
//   backward  sums: gz = dy * silu'(z), z = x*A + D recomputed from x; per-channel sum gz, sum gz*x (centred to sum gz*xhat
//             once per channel in the epilogue) -> dbeta / dgamma atomics (skipped for frozen parameters) and per-group
//             gamma-weighted partials;  apply: dx = gz * A[c] + x * c2[c] + c0[c] (gz recomputed).
// A thread owns one 16-byte vector column (8 bf16 / 4 fp32 channels) and walks down the rows with 2-4 rows of raw vectors
// in flight; blockDim is the largest multiple of C/VEC that fits 512 threads. An optional fp32 addend e[n][c] (a
// convolution bias, ResBlock's timestep embedding) is normalised with x — GroupNorm(x + e) — at no in-loop cost: it enters
// the per-channel constants, and the statistics' per-CTA epilogue (sum += rows*e, sum of squares += 2 e sum + rows e^2).
#include <cstdlib>
#include <cuda_bf16.h>

#include "capi_util.h"

namespace vt {
namespace {

template <typename T>
struct V;
template <>
struct V<__nv_bfloat16> {
  static constexpr int VEC = 8;
  using Raw = uint4;
  __device__ static Raw load_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static void unpack(const Raw& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
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
  using Raw = float4;
  __device__ static Raw load_raw(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ static void unpack(const Raw& v, float* f) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
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
// Reduce per-thread per-channel partials over the CTA's `rpp` row groups: every thread stores its VEC values into
// sm[r][channel] (plain stores), one barrier, then each thread sums one channel over the rpp rows. (Shared-memory float
// atomics — the first version — are CAS loops: with 12 row groups on one address they cost more than the sweep itself.)
template <int VEC, int NACC>
__device__ __forceinline__ void cta_channel_sums(float* sm, const RowSplit& rs, int C, const float (*acc)[VEC], float* out) {
  // sm: [rpp][NACC][C]; out: [NACC][C] (may alias the first row group's slab only after the barrier: use a separate slab)
  if (rs.active) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int k = 0; k < VEC; ++k) sm[(rs.r * NACC + a) * C + rs.v * VEC + k] = acc[a][k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NACC * C; i += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < rs.rpp; ++r) t += sm[r * NACC * C + i];
    out[i] = t;
  }
  __syncthreads();
}

template <typename T>
__global__ void gn_nhwc_stats_kernel(const T* __restrict__ x, const float* __restrict__ addend, int addend_stride,
                                     float* __restrict__ ws, int C, int S, int G, int rows_per_cta) {
  constexpr int VEC = V<T>::VEC;
  extern __shared__ float sm[];  // [rpp][2][C] partials, then [2][C] totals
  const int n = blockIdx.y;
  const RowSplit rs = row_split<VEC>(C, S, rows_per_cta);
  float acc[2][VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[0][k] = acc[1][k] = 0.f;
  if (rs.active) {
    const T* base = x + (static_cast<size_t>(n) * S) * C + rs.v * VEC;
    int row = rs.row0 + rs.r;
    for (; row + 3 * rs.rpp < rs.row1; row += 4 * rs.rpp) {  // four independent 16-byte loads in flight per thread
      float f[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) V<T>::load(base + static_cast<size_t>(row + u * rs.rpp) * C, f[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          acc[0][k] += f[u][k];
          acc[1][k] = fmaf(f[u][k], f[u][k], acc[1][k]);
        }
    }
    for (; row < rs.row1; row += rs.rpp) {
      float f[VEC];
      V<T>::load(base + static_cast<size_t>(row) * C, f);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        acc[0][k] += f[k];
        acc[1][k] = fmaf(f[k], f[k], acc[1][k]);
      }
    }
  }
  float* tot = sm + static_cast<size_t>(rs.rpp) * 2 * C;
  cta_channel_sums<VEC, 2>(sm, rs, C, acc, tot);
  const int cpg = C / G;
  if (addend != nullptr) {  // statistics of x + e[n][c]:  sum += rows*e,  sum of squares += 2 e sum + rows e^2
    const float rows = static_cast<float>(max(rs.row1 - rs.row0, 0));
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float e = addend[static_cast<size_t>(n) * addend_stride + c];
      tot[C + c] += e * fmaf(rows, e, 2.f * tot[c]);
      tot[c] = fmaf(rows, e, tot[c]);
    }
    __syncthreads();
  }
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += tot[c];
      b += tot[C + c];
    }
    float* dst = ws + ((static_cast<size_t>(n) * gridDim.x + blockIdx.x) * G + g) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

// Sum one sample's per-CTA partials, ws_n[chunks][w] (w = 2G), into out[w] (shared memory): thread t owns column t % w and
// every (blockDim / w)-th chunk — coalesced w-float rows, independent loads — and a shared-memory pass joins the slices.
// scratch: blockDim floats. (One thread per group walking the chunks serially cost 427 dependent L2 loads on the 5-D input:
// 104 us per CTA of the second sweep.)
__device__ __forceinline__ void cta_reduce_partials(const float* __restrict__ ws_n, int w, int chunks, float* scratch, float* out) {
  if (chunks == 1) {
    for (int i = threadIdx.x; i < w; i += blockDim.x) out[i] = ws_n[i];
    __syncthreads();
    return;
  }
  const int slices = blockDim.x / w;
  const int col = threadIdx.x % w, sl = threadIdx.x / w;
  if (sl < slices) {
    float acc = 0.f;
#pragma unroll 4
    for (int c = sl; c < chunks; c += slices) acc += ws_n[static_cast<size_t>(c) * w + col];
    scratch[sl * w + col] = acc;
  }
  __syncthreads();
  if (threadIdx.x < w) {
    float t = 0.f;
    for (int i = 0; i < slices; ++i) t += scratch[i * w + threadIdx.x];
    out[threadIdx.x] = t;
  }
  __syncthreads();
}

// Samples cut into many chunks (the 5-D temporal input: N = 2, 296 chunks each) get their partials summed ONCE by this
// kernel (one CTA per sample) -> totals[n][G][2]; with <= kInlineChunks chunks the second sweep's prologue does it itself
// and the launch disappears.
constexpr int kInlineChunks = 32;
__global__ void gn_nhwc_finalize_kernel(const float* __restrict__ ws, float* __restrict__ totals, int G, int chunks) {
  extern __shared__ float sm[];  // [blockDim] scratch + [2G]
  const int n = blockIdx.x, w = 2 * G;
  float* out = sm + blockDim.x;
  cta_reduce_partials(ws + static_cast<size_t>(n) * chunks * w, w, chunks, sm, out);
  if (threadIdx.x < w) totals[static_cast<size_t>(n) * w + threadIdx.x] = out[threadIdx.x];
}

// parts: float[N][nparts][G][2] — the first sweep's per-CTA partials (nparts = its chunk count, summed here) or the
// finalize kernel's totals (nparts = 1).
template <typename T>
__global__ void gn_nhwc_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ parts, int nparts,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ addend, int addend_stride, int C, int S, int G,
                                     float eps, int apply_silu, int rows_per_cta) {
  constexpr int VEC = V<T>::VEC;
  extern __shared__ float sm[];  // a[C], d[C], totals[2G], scratch[blockDim]
  const int n = blockIdx.y;
  const RowSplit rs = row_split<VEC>(C, S, rows_per_cta);
  const int cpg = C / G;
  const float inv = 1.f / (static_cast<float>(cpg) * S);
  float* gs = sm + 2 * C;  // [G][2]: sum, sum of squares -> mean, rstd
  cta_reduce_partials(parts + static_cast<size_t>(n) * nparts * 2 * G, 2 * G, nparts, gs + 2 * G, gs);
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float m = gs[2 * g] * inv;
    const float r = rsqrtf(fmaxf(gs[2 * g + 1] * inv - m * m, 0.f) + eps);
    gs[2 * g] = m;
    gs[2 * g + 1] = r;
    if (blockIdx.x == 0) {
      mean_out[n * G + g] = m;
      rstd_out[n * G + g] = r;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float a = gs[2 * g + 1] * (gamma ? gamma[c] : 1.f);
    const float e = addend ? addend[static_cast<size_t>(n) * addend_stride + c] : 0.f;
    sm[c] = a;
    sm[C + c] = fmaf(e - gs[2 * g], a, beta ? beta[c] : 0.f);  // (x + e - mean) * a + beta
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
  auto one = [&](const typename V<T>::Raw& raw, int row) {
    float f[VEC];
    V<T>::unpack(raw, f);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float z = fmaf(f[k], a[k], d[k]);
      f[k] = apply_silu ? silu_fast(z) : z;
    }
    V<T>::store(y + off + static_cast<size_t>(row) * C, f);
  };
  int row = rs.row0 + rs.r;
  for (; row + 3 * rs.rpp < rs.row1; row += 4 * rs.rpp) {  // four 16-byte loads in flight per thread
    typename V<T>::Raw raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = V<T>::load_raw(x + off + static_cast<size_t>(row + u * rs.rpp) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) one(raw[u], row + u * rs.rpp);
  }
  for (; row < rs.row1; row += rs.rpp) one(V<T>::load_raw(x + off + static_cast<size_t>(row) * C), row);
}

// ---- backward --------------------------------------------------------------------------------------------------------
// z = x*A + D with A = rstd*gamma, D = beta - mean*A;  gz = dy * silu'(z);  xhat = (x - mean) * rstd.
// sums: per channel  s0 = sum gz,  s1 = sum gz*x  (raw x: two per-channel constants instead of four in registers; the
//       centring  sum gz*xhat = rstd * (s1 - mean*s0)  happens once per channel in the epilogue)
//       -> dbeta / dgamma atomics (caller zeroes) and per-CTA per-group gamma-weighted partials in ws (plain stores).
// apply: dx = rstd * (gz*gamma - m1 - xhat*m2) = gz*A + x*c2 + c0,  c2 = -rstd^2*m2,  c0 = -rstd*m1 - mean*c2.
template <typename T>
__global__ void __launch_bounds__(512, 2) gn_nhwc_bwd_sums_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const float* __restrict__ addend, int addend_stride,
                                        float* __restrict__ ws, float* __restrict__ dgamma,
                                        float* __restrict__ dbeta, int C, int S, int G, int apply_silu, int rows_per_cta) {
  constexpr int VEC = V<T>::VEC;
  extern __shared__ float sm[];  // [rpp][2][C] partials, then [2][C]: sum gz, sum gz*x per channel
  const int n = blockIdx.y;
  const RowSplit rs = row_split<VEC>(C, S, rows_per_cta);
  const int cpg = C / G;
  float acc[2][VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[0][k] = acc[1][k] = 0.f;
  if (rs.active) {
    float A[VEC], D[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = rs.v * VEC + k, g = c / cpg;
      const float e = addend ? addend[static_cast<size_t>(n) * addend_stride + c] : 0.f;
      A[k] = rstd[n * G + g] * (gamma ? gamma[c] : 1.f);
      D[k] = fmaf(e - mean[n * G + g], A[k], beta ? beta[c] : 0.f);
    }
    const size_t off = (static_cast<size_t>(n) * S) * C + rs.v * VEC;
    auto one = [&](const typename V<T>::Raw& rx, const typename V<T>::Raw& rg) {
      float fx[VEC], fg[VEC];
      V<T>::unpack(rx, fx);
      V<T>::unpack(rg, fg);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float gz = fg[k];
        if (apply_silu) gz *= dsilu_fast(fmaf(fx[k], A[k], D[k]));
        acc[0][k] += gz;
        acc[1][k] = fmaf(gz, fx[k], acc[1][k]);
      }
    };
    int row = rs.row0 + rs.r;
    for (; row + rs.rpp < rs.row1; row += 2 * rs.rpp) {  // two rows of x and dy in flight per thread (raw: 16 registers)
      const size_t o0 = off + static_cast<size_t>(row) * C, o1 = o0 + static_cast<size_t>(rs.rpp) * C;
      const typename V<T>::Raw x0 = V<T>::load_raw(x + o0), g0 = V<T>::load_raw(dy + o0);
      const typename V<T>::Raw x1 = V<T>::load_raw(x + o1), g1 = V<T>::load_raw(dy + o1);
      one(x0, g0);
      one(x1, g1);
    }
    if (row < rs.row1) {
      const size_t o0 = off + static_cast<size_t>(row) * C;
      one(V<T>::load_raw(x + o0), V<T>::load_raw(dy + o0));
    }
  }
  float* tot = sm + static_cast<size_t>(rs.rpp) * 2 * C;
  cta_channel_sums<VEC, 2>(sm, rs, C, acc, tot);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {  // centre: sum gz*xhat = rstd * (sum gz*x - mean * sum gz)
    const int g = c / cpg;
    const float e = addend ? addend[static_cast<size_t>(n) * addend_stride + c] : 0.f;  // sum gz*(x + e) = sum gz*x + e sum gz
    tot[C + c] = rstd[n * G + g] * fmaf(e - mean[n * G + g], tot[c], tot[C + c]);
    if (dbeta) atomicAdd(&dbeta[c], tot[c]);
    if (dgamma) atomicAdd(&dgamma[c], tot[C + c]);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float gm = gamma ? gamma[c] : 1.f;
      a = fmaf(tot[c], gm, a);
      b = fmaf(tot[C + c], gm, b);
    }
    float* dst = ws + ((static_cast<size_t>(n) * gridDim.x + blockIdx.x) * G + g) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

template <typename T>
__global__ void __launch_bounds__(512, 2) gn_nhwc_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, const float* __restrict__ addend, int addend_stride,
                                         const float* __restrict__ parts, int nparts,
                                         T* __restrict__ dx, int C, int S, int G, int apply_silu, int rows_per_cta) {
  constexpr int VEC = V<T>::VEC;
  extern __shared__ float sm[];  // totals[2G], scratch[blockDim]
  const int n = blockIdx.y;
  const RowSplit rs = row_split<VEC>(C, S, rows_per_cta);
  const int cpg = C / G;
  const float inv = 1.f / (static_cast<float>(cpg) * S);
  cta_reduce_partials(parts + static_cast<size_t>(n) * nparts * 2 * G, 2 * G, nparts, sm + 2 * G, sm);
  if (!rs.active) return;
  float A[VEC], D[VEC], c2[VEC], c0[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const int c = rs.v * VEC + k, g = c / cpg;
    const float mu = mean[n * G + g], rr = rstd[n * G + g];
    const float m1 = sm[2 * g] * inv, m2 = sm[2 * g + 1] * inv;
    const float e = addend ? addend[static_cast<size_t>(n) * addend_stride + c] : 0.f;
    A[k] = rr * (gamma ? gamma[c] : 1.f);
    D[k] = fmaf(e - mu, A[k], beta ? beta[c] : 0.f);
    c2[k] = -rr * rr * m2;
    c0[k] = -rr * m1 - (mu - e) * c2[k];  // (x + e - mu) * c2
  }
  const size_t off = (static_cast<size_t>(n) * S) * C + rs.v * VEC;
  auto one = [&](const typename V<T>::Raw& rx, const typename V<T>::Raw& rg, size_t o) {
    float fx[VEC], fg[VEC];
    V<T>::unpack(rx, fx);
    V<T>::unpack(rg, fg);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float gz = fg[k];
      if (apply_silu) gz *= dsilu_fast(fmaf(fx[k], A[k], D[k]));
      fx[k] = fmaf(gz, A[k], fmaf(fx[k], c2[k], c0[k]));
    }
    V<T>::store(dx + o, fx);
  };
  int row = rs.row0 + rs.r;
  for (; row + rs.rpp < rs.row1; row += 2 * rs.rpp) {
    const size_t o0 = off + static_cast<size_t>(row) * C, o1 = o0 + static_cast<size_t>(rs.rpp) * C;
    const typename V<T>::Raw x0 = V<T>::load_raw(x + o0), g0 = V<T>::load_raw(dy + o0);
    const typename V<T>::Raw x1 = V<T>::load_raw(x + o1), g1 = V<T>::load_raw(dy + o1);
    one(x0, g0, o0);
    one(x1, g1, o1);
  }
  if (row < rs.row1) {
    const size_t o0 = off + static_cast<size_t>(row) * C;
    one(V<T>::load_raw(x + o0), V<T>::load_raw(dy + o0), o0);
  }
}

struct Plan {
  int threads, rows_per_cta, chunks, rpp;
};
template <int VEC>
bool make_plan(int N, int C, int S, Plan* p) {
  if (C % VEC != 0) return false;
  const int vpr = C / VEC;
  if (vpr > 512) return false;
  p->threads = (512 / vpr) * vpr;
  const int rpp = p->threads / vpr;
  p->rpp = rpp;
  // CTAs per SM over the whole grid: 4 (2 for the few-sample 5-D inputs, whose partials need the finalize launch); at
  // least eight passes of rows per CTA so the per-CTA prologue / reduction epilogue stays amortised on the small levels
  // (measured: profiles/r2_s29_gncl.txt). VT_GNCL_CTAS=1..8 overrides.
  static const int forced = [] {
    const char* e = getenv("VT_GNCL_CTAS");
    const int v = e != nullptr ? atoi(e) : 0;
    return v < 0 ? 0 : (v > 8 ? 8 : v);
  }();
  const int per_sm = forced != 0 ? forced : (N < 8 ? 2 : 4);
  int chunks = (148 * per_sm + N - 1) / N;
  int rows = (S + chunks - 1) / chunks;
  rows = ((rows + rpp - 1) / rpp) * rpp;
  if (rows < 8 * rpp) rows = 8 * rpp;
  p->rows_per_cta = rows;
  p->chunks = (S + rows - 1) / rows;
  return true;
}

// finalize launch: a multiple of 2G threads (whole slices), at most 1024
int finalize_threads(int G) {
  const int w = 2 * G;
  int slices = 1024 / w;
  if (slices > 8) slices = 8;
  if (slices < 1) slices = 1;
  return slices * w;
}

}  // namespace
}  // namespace vt

using namespace vt;

extern "C" {

// Workspace: per-CTA partial group sums, float[N][chunks][G][2]; the bound below covers every plan make_plan() can produce.
int64_t vt_groupnorm_nhwc_workspace_bytes(int N, int G) {
  return static_cast<int64_t>(N) * ((148 * 8 + N - 1) / N + 2) * G * 2 * sizeof(float);  // partials (<= 8 CTAs/SM) + totals
}

int vt_groupnorm_silu_nhwc_fwd(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta,
                               const float* addend, int addend_stride, void* workspace, int N, int C, int S, int G, float eps, int apply_silu, int dtype,
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
  const size_t smem = static_cast<size_t>(p.rpp + 1) * 2 * C * sizeof(float);  // row-group partials + totals
  const size_t smem_apply = (2 * static_cast<size_t>(C) + 2 * static_cast<size_t>(G) + p.threads) * sizeof(float);
  float* ws = static_cast<float*>(workspace);
  if (smem > 48 * 1024 && first_on_device(dtype == 0 ? reinterpret_cast<const void*>(gn_nhwc_stats_kernel<__nv_bfloat16>)
                                                      : reinterpret_cast<const void*>(gn_nhwc_stats_kernel<float>))) {
    VT_CHECK_CUDA(cudaFuncSetAttribute(gn_nhwc_stats_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VT_CHECK_CUDA(cudaFuncSetAttribute(gn_nhwc_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  VT_REQUIRE(2 * G <= p.threads, VT_ERR_SHAPE, "G=%d too large", G);
  const bool fin = p.chunks > kInlineChunks;  // many chunks per sample: sum the partials once, in their own launch
  const float* parts = fin ? ws + static_cast<size_t>(N) * p.chunks * G * 2 : ws;
  const int nparts = fin ? 1 : p.chunks;
  const int fin_threads = finalize_threads(G);
  const size_t fin_smem = static_cast<size_t>(fin_threads + 2 * G) * sizeof(float);
  if (dtype == 0) {
    gn_nhwc_stats_kernel<__nv_bfloat16><<<grid, p.threads, smem, st>>>(static_cast<const __nv_bfloat16*>(x), addend, addend_stride, ws, C, S, G,
                                                                      p.rows_per_cta);
    if (fin) gn_nhwc_finalize_kernel<<<N, fin_threads, fin_smem, st>>>(ws, const_cast<float*>(parts), G, p.chunks);
    gn_nhwc_apply_kernel<__nv_bfloat16><<<grid, p.threads, smem_apply, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y),
                                                                      parts, nparts, mean, rstd, gamma, beta, addend, addend_stride, C, S, G, eps,
                                                                      apply_silu, p.rows_per_cta);
  } else {
    gn_nhwc_stats_kernel<float><<<grid, p.threads, smem, st>>>(static_cast<const float*>(x), addend, addend_stride, ws, C, S, G, p.rows_per_cta);
    if (fin) gn_nhwc_finalize_kernel<<<N, fin_threads, fin_smem, st>>>(ws, const_cast<float*>(parts), G, p.chunks);
    gn_nhwc_apply_kernel<float><<<grid, p.threads, smem_apply, st>>>(static_cast<const float*>(x), static_cast<float*>(y), parts, nparts, mean,
                                                              rstd, gamma, beta, addend, addend_stride, C, S, G, eps, apply_silu,
                                                              p.rows_per_cta);
  }
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vt_groupnorm_silu_nhwc_bwd(const void* dy, const void* x, const float* mean, const float* rstd, void* dx,
                               const float* gamma, const float* beta, const float* addend, int addend_stride, float* dgamma,
                               float* dbeta, void* workspace, int N,
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
  const size_t smem = static_cast<size_t>(p.rpp + 1) * 2 * C * sizeof(float);  // row-group partials + totals
  const size_t smem_g = (2 * static_cast<size_t>(G) + p.threads) * sizeof(float);
  float* ws = static_cast<float*>(workspace);
  if (smem > 48 * 1024 && first_on_device(dtype == 0 ? reinterpret_cast<const void*>(gn_nhwc_bwd_sums_kernel<__nv_bfloat16>)
                                                      : reinterpret_cast<const void*>(gn_nhwc_bwd_sums_kernel<float>))) {
    VT_CHECK_CUDA(cudaFuncSetAttribute(gn_nhwc_bwd_sums_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VT_CHECK_CUDA(cudaFuncSetAttribute(gn_nhwc_bwd_sums_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  VT_REQUIRE(2 * G <= p.threads, VT_ERR_SHAPE, "G=%d too large", G);
  const bool fin = p.chunks > kInlineChunks;
  const float* parts = fin ? ws + static_cast<size_t>(N) * p.chunks * G * 2 : ws;
  const int nparts = fin ? 1 : p.chunks;
  const int fin_threads = finalize_threads(G);
  const size_t fin_smem = static_cast<size_t>(fin_threads + 2 * G) * sizeof(float);
  if (dtype == 0) {
    auto a = static_cast<const __nv_bfloat16*>(dy);
    auto b = static_cast<const __nv_bfloat16*>(x);
    gn_nhwc_bwd_sums_kernel<__nv_bfloat16><<<grid, p.threads, smem, st>>>(a, b, mean, rstd, gamma, beta, addend, addend_stride, ws, dgamma,
                                                                         dbeta, C, S, G, apply_silu, p.rows_per_cta);
    if (fin) gn_nhwc_finalize_kernel<<<N, fin_threads, fin_smem, st>>>(ws, const_cast<float*>(parts), G, p.chunks);
    gn_nhwc_bwd_apply_kernel<__nv_bfloat16><<<grid, p.threads, smem_g, st>>>(a, b, mean, rstd, gamma, beta, addend, addend_stride, parts,
                                                                       nparts, static_cast<__nv_bfloat16*>(dx), C, S, G, apply_silu,
                                                                       p.rows_per_cta);
  } else {
    auto a = static_cast<const float*>(dy);
    auto b = static_cast<const float*>(x);
    gn_nhwc_bwd_sums_kernel<float><<<grid, p.threads, smem, st>>>(a, b, mean, rstd, gamma, beta, addend, addend_stride, ws, dgamma, dbeta, C,
                                                                 S, G, apply_silu, p.rows_per_cta);
    if (fin) gn_nhwc_finalize_kernel<<<N, fin_threads, fin_smem, st>>>(ws, const_cast<float*>(parts), G, p.chunks);
    gn_nhwc_bwd_apply_kernel<float><<<grid, p.threads, smem_g, st>>>(a, b, mean, rstd, gamma, beta, addend, addend_stride, parts, nparts,
                                                               static_cast<float*>(dx), C, S, G, apply_silu, p.rows_per_cta);
  }
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
