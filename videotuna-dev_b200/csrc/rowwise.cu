// rowwise.cu — HBM-bound row kernels around the attention core:
//   LayerNorm(+affine)+adaLN-modulate, gated residual, fused QK-RMSNorm + RoPE; forward and backward.
// Common shape: a CTA walks rows of one batch; thread t owns the 8 contiguous columns [8t, 8t+8) of every row, so
//   - every global access is a 16-byte vector, coalesced across the CTA,
//   - column reductions (dscale/dshift/dgate/dgamma/dw) accumulate in registers over the CTA's rows and are flushed
//     with one atomicAdd per column per CTA,
//   - row statistics use one shuffle+smem block reduction per group of RPI rows (RPI loads in flight per thread).
// Algorithmic traffic: 4 B/elem forward (read x, write y), 6 B/elem backward (read dy, x; write dx).
#include <cstdlib>
#include <cuda_bf16.h>

#include "capi_util.h"

namespace vt {
namespace {

// Kernels are templated on <MAXT, RPI>: MAXT bounds the CTA size (register budget = 64K / MAXT) and RPI is the number
// of rows processed per iteration (independent 16-byte loads in flight per thread). Wide rows trade RPI for threads.

struct alignas(16) Vec8 {
  uint4 u;
};
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// 8 consecutive elements of a row in either storage type (bf16: one 16-byte access; fp32: two).
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { uint4 a; };
template <> struct Raw8<float> { uint4 a, b; };
__device__ __forceinline__ void zero8(Raw8<__nv_bfloat16>& r) { r.a = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ void zero8(Raw8<float>& r) { r.a = r.b = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ void ldg8(const __nv_bfloat16* p, Raw8<__nv_bfloat16>& r) { r.a = ldg_stream(p); }
__device__ __forceinline__ void ldg8(const float* p, Raw8<float>& r) { r.a = ldg_stream(p); r.b = ldg_stream(p + 4); }
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float* f) { unpack8(r.a, f); }
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float* f) {
  f[0] = __uint_as_float(r.a.x); f[1] = __uint_as_float(r.a.y); f[2] = __uint_as_float(r.a.z); f[3] = __uint_as_float(r.a.w);
  f[4] = __uint_as_float(r.b.x); f[5] = __uint_as_float(r.b.y); f[6] = __uint_as_float(r.b.z); f[7] = __uint_as_float(r.b.w);
}
__device__ __forceinline__ void stg8(__nv_bfloat16* p, const float* f) { stg_stream(p, pack8(f)); }
__device__ __forceinline__ void stg8(float* p, const float* f) {
  stg_stream(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
  stg_stream(p + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
}

// Sum NV values per thread across the CTA. `red` holds NV * 32 floats. All threads get the totals.
template <int NV>
__device__ __forceinline__ void block_sum(float* v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  if (nwarp == 1) return;
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float t = lane < nwarp ? red[i * 32 + lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    v[i] = t;
  }
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm + modulate
// ------------------------------------------------------------------------------------------------------------
template <typename XT, int MAXT, int RPI>
__global__ void __launch_bounds__(MAXT) ln_modulate_fwd_kernel(const XT* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                       float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ scale, const float* __restrict__ shift, int L, int C,
                                       float eps) {
  __shared__ float red[2 * RPI * 32];
  const int b = blockIdx.y;
  const int col = threadIdx.x * 8;
  const bool active = col < C;
  float g[8], be[8], sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { g[i] = 1.f; be[i] = 0.f; sc[i] = 1.f; sh[i] = 0.f; }
  if (active) {
    if (gamma) load8f(gamma + col, g);
    if (beta) load8f(beta + col, be);
    if (scale) {
      load8f(scale + static_cast<size_t>(b) * C + col, sc);
#pragma unroll
      for (int i = 0; i < 8; ++i) sc[i] += 1.f;
    }
    if (shift) load8f(shift + static_cast<size_t>(b) * C + col, sh);
  }
  const float invC = 1.f / C;
  for (int l0 = blockIdx.x * RPI; l0 < L; l0 += gridDim.x * RPI) {
    Raw8<XT> raw[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      zero8(raw[r]);
      if (active && l0 + r < L) ldg8(x + (static_cast<size_t>(b) * L + l0 + r) * C + col, raw[r]);
    }
    float s[RPI];
    float f[RPI][8];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      unpack8(raw[r], f[r]);
      s[r] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s[r] += f[r][i];
    }
    block_sum<RPI>(s, red);
    float q[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      s[r] *= invC;  // mean
      q[r] = 0.f;
      if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = f[r][i] - s[r];
          q[r] += d * d;
        }
      }
    }
    block_sum<RPI>(q, red);
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      if (l0 + r >= L) continue;
      const float rstd = rsqrtf(q[r] * invC + eps);
      const size_t row = static_cast<size_t>(b) * L + l0 + r;
      if (threadIdx.x == 0) {
        if (mean_out) mean_out[row] = s[r];
        if (rstd_out) rstd_out[row] = rstd;
      }
      if (active) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ((f[r][i] - s[r]) * rstd * g[i] + be[i]) * sc[i] + sh[i];
        stg_stream(y + row * C + col, pack8(o));
      }
    }
  }
}

// dx = rstd * (gh - mean(gh) - xh * mean(gh * xh)),  gh = dy * (1+scale) * gamma,  xh = (x - mean) * rstd
// dshift += dy; dscale += dy * (xh*gamma + beta); dbeta += dy*(1+scale); dgamma += dy*(1+scale)*xh
template <typename XT, int MAXT, int RPI>
__global__ void __launch_bounds__(MAXT) ln_modulate_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const XT* __restrict__ x,
                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                       XT* __restrict__ dx, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ scale,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ dscale, float* __restrict__ dshift, int L, int C) {
  __shared__ float red[2 * RPI * 32];
  const int b = blockIdx.y;
  const int col = threadIdx.x * 8;
  const bool active = col < C;
  float g[8], be[8], sc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { g[i] = 1.f; be[i] = 0.f; sc[i] = 1.f; }
  if (active) {
    if (gamma) load8f(gamma + col, g);
    if (beta) load8f(beta + col, be);
    if (scale) {
      load8f(scale + static_cast<size_t>(b) * C + col, sc);
#pragma unroll
      for (int i = 0; i < 8; ++i) sc[i] += 1.f;
    }
  }
  float a_dgamma[8] = {0}, a_dbeta[8] = {0}, a_dscale[8] = {0}, a_dshift[8] = {0};
  const float invC = 1.f / C;
  for (int l0 = blockIdx.x * RPI; l0 < L; l0 += gridDim.x * RPI) {
    uint4 rdy[RPI];
    Raw8<XT> rx[RPI];
    float mu[RPI], rs[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      rdy[r] = make_uint4(0, 0, 0, 0);
      zero8(rx[r]);
      mu[r] = 0.f;
      rs[r] = 0.f;
      if (l0 + r < L) {
        const size_t row = static_cast<size_t>(b) * L + l0 + r;
        mu[r] = mean[row];
        rs[r] = rstd[row];
        if (active) {
          rdy[r] = ldg_stream(dy + row * C + col);
          ldg8(x + row * C + col, rx[r]);
        }
      }
    }
    float gh[RPI][8], xh[RPI][8];
    float sums[2 * RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      float fdy[8], fx[8];
      unpack8(rdy[r], fdy);
      unpack8(rx[r], fx);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[r][i] = active ? (fx[i] - mu[r]) * rs[r] : 0.f;
        const float gm = fdy[i] * sc[i];  // grad wrt (xh*gamma+beta)
        gh[r][i] = gm * g[i];
        s1 += gh[r][i];
        s2 += gh[r][i] * xh[r][i];
        a_dshift[i] += fdy[i];
        a_dscale[i] += fdy[i] * (xh[r][i] * g[i] + be[i]);
        a_dbeta[i] += gm;
        a_dgamma[i] += gm * xh[r][i];
      }
      sums[2 * r] = s1;
      sums[2 * r + 1] = s2;
    }
    block_sum<2 * RPI>(sums, red);
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      if (!active || l0 + r >= L) continue;
      const float m1 = sums[2 * r] * invC, m2 = sums[2 * r + 1] * invC;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = rs[r] * (gh[r][i] - m1 - xh[r][i] * m2);
      stg8(dx + (static_cast<size_t>(b) * L + l0 + r) * C + col, o);
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (dgamma) atomicAdd(dgamma + col + i, a_dgamma[i]);
      if (dbeta) atomicAdd(dbeta + col + i, a_dbeta[i]);
      if (dscale) atomicAdd(dscale + static_cast<size_t>(b) * C + col + i, a_dscale[i]);
      if (dshift) atomicAdd(dshift + static_cast<size_t>(b) * C + col + i, a_dshift[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// gated residual
// ------------------------------------------------------------------------------------------------------------
template <typename XT, int MAXT, int RPI>
__global__ void __launch_bounds__(MAXT) gate_residual_fwd_kernel(const XT* __restrict__ x, const __nv_bfloat16* __restrict__ br,
                                         XT* __restrict__ y, const float* __restrict__ gate, int L, int C) {
  const int b = blockIdx.y;
  const int col = threadIdx.x * 8;
  if (col >= C) return;
  float g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = 1.f;
  if (gate) load8f(gate + static_cast<size_t>(b) * C + col, g);
  for (int l0 = blockIdx.x * RPI; l0 < L; l0 += gridDim.x * RPI) {
    Raw8<XT> rx[RPI];
    uint4 rb[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      if (l0 + r < L) {
        const size_t off = (static_cast<size_t>(b) * L + l0 + r) * C + col;
        ldg8(x + off, rx[r]);
        rb[r] = ldg_stream(br + off);
      }
    }
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      if (l0 + r >= L) continue;
      float fx[8], fb[8], o[8];
      unpack8(rx[r], fx);
      unpack8(rb[r], fb);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(fb[i], g[i], fx[i]);
      stg8(y + (static_cast<size_t>(b) * L + l0 + r) * C + col, o);
    }
  }
}

template <typename XT, int MAXT, int RPI>
__global__ void __launch_bounds__(MAXT) gate_residual_bwd_kernel(const XT* __restrict__ dy, const __nv_bfloat16* __restrict__ br,
                                         __nv_bfloat16* __restrict__ dbr, const float* __restrict__ gate,
                                         float* __restrict__ dgate, int L, int C) {
  const int b = blockIdx.y;
  const int col = threadIdx.x * 8;
  if (col >= C) return;
  float g[8], acc[8] = {0};
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = 1.f;
  if (gate) load8f(gate + static_cast<size_t>(b) * C + col, g);
  for (int l0 = blockIdx.x * RPI; l0 < L; l0 += gridDim.x * RPI) {
    Raw8<XT> rd[RPI];
    uint4 rb[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      if (l0 + r < L) {
        const size_t off = (static_cast<size_t>(b) * L + l0 + r) * C + col;
        ldg8(dy + off, rd[r]);
        if (dgate) rb[r] = ldg_stream(br + off);
      }
    }
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      if (l0 + r >= L) continue;
      float fd[8], fb[8], o[8];
      unpack8(rd[r], fd);
      if (dgate) {
        unpack8(rb[r], fb);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += fd[i] * fb[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fd[i] * g[i];
      stg_stream(dbr + (static_cast<size_t>(b) * L + l0 + r) * C + col, pack8(o));
    }
  }
  if (dgate) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(dgate + static_cast<size_t>(b) * C + col + i, acc[i]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// fused RMSNorm (per head or full row) + interleaved RoPE.  Row = one token (b,l) with H heads of D elements.
//   norm_mode 0: none;  1: per head, weight (D);  2: whole row, weight (H*D)
//   y[2i]   = n[2i]*cos[2i]   - n[2i+1]*sin[2i]          (n = normalised x)
//   y[2i+1] = n[2i+1]*cos[2i+1] + n[2i]*sin[2i+1]       tokens l >= L_rope are not rotated
// ------------------------------------------------------------------------------------------------------------
template <int NORM, int MAXT, int RPI>
__global__ void __launch_bounds__(MAXT) rmsnorm_rope_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                        float* __restrict__ rstd_out, const float* __restrict__ w,
                                        const float* __restrict__ cosT, const float* __restrict__ sinT, int64_t x_sb,
                                        int64_t x_sl, int64_t x_sh, int64_t y_sb, int64_t y_sl, int64_t y_sh, int L,
                                        int H, int D, int L_rope, float eps) {
  __shared__ float red[RPI * 32];
  const int b = blockIdx.y;
  const int col = threadIdx.x * 8;
  const int C = H * D;
  const bool active = col < C;
  const int h = active ? col / D : 0, d0 = active ? col % D : 0;
  const int tpg = D / 8;  // threads per head
  float wv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = 1.f;
  if (active && w != nullptr && NORM != 0) load8f(w + (NORM == 1 ? d0 : col), wv);

  for (int l0 = blockIdx.x * RPI; l0 < L; l0 += gridDim.x * RPI) {
    uint4 raw[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      raw[r] = make_uint4(0, 0, 0, 0);
      if (active && l0 + r < L) raw[r] = ldg_stream(x + b * x_sb + static_cast<int64_t>(l0 + r) * x_sl + h * x_sh + d0);
    }
    float f[RPI][8], ss[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      unpack8(raw[r], f[r]);
      ss[r] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss[r] += f[r][i] * f[r][i];
    }
    if (NORM == 1) {  // reduce over the D/8 consecutive threads of one head (D/8 is 8 or 16: within a warp)
#pragma unroll
      for (int r = 0; r < RPI; ++r)
        for (int o = tpg >> 1; o > 0; o >>= 1) ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);
    } else if (NORM == 2) {
      block_sum<RPI>(ss, red);
    }
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      const int l = l0 + r;
      if (l >= L) continue;
      float rstd = 1.f;
      if (NORM != 0) {
        rstd = rsqrtf(ss[r] / (NORM == 1 ? D : C) + eps);
        if (rstd_out != nullptr) {
          if (NORM == 1) {
            if (active && (threadIdx.x % tpg) == 0) rstd_out[(static_cast<size_t>(b) * L + l) * H + h] = rstd;
          } else if (threadIdx.x == 0) {
            rstd_out[static_cast<size_t>(b) * L + l] = rstd;
          }
        }
      }
      if (!active) continue;
      float n[8], o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) n[i] = f[r][i] * rstd * wv[i];
      if (cosT != nullptr && l < L_rope) {
        float cs[8], sn[8];
        load8f(cosT + static_cast<size_t>(l) * D + d0, cs);
        load8f(sinT + static_cast<size_t>(l) * D + d0, sn);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          o[i] = n[i] * cs[i] - n[i + 1] * sn[i];
          o[i + 1] = n[i + 1] * cs[i + 1] + n[i] * sn[i + 1];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = n[i];
      }
      stg_stream(y + b * y_sb + static_cast<int64_t>(l) * y_sl + h * y_sh + d0, pack8(o));
    }
  }
}

// Backward: g = R^T dy (inverse rotation), then RMSNorm backward:
//   xh = x*rstd; gw = g*w; dx = rstd * (gw - xh * mean(gw*xh)); dw += g*xh
template <int NORM, int MAXT, int RPI>
__global__ void __launch_bounds__(MAXT) rmsnorm_rope_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                        const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx,
                                        float* __restrict__ dw, const float* __restrict__ w,
                                        const float* __restrict__ cosT, const float* __restrict__ sinT, int64_t g_sb,
                                        int64_t g_sl, int64_t g_sh, int64_t x_sb, int64_t x_sl, int64_t x_sh,
                                        int64_t o_sb, int64_t o_sl, int64_t o_sh, int L, int H, int D, int L_rope) {
  __shared__ float red[RPI * 32];
  const int b = blockIdx.y;
  const int col = threadIdx.x * 8;
  const int C = H * D;
  const bool active = col < C;
  const int h = active ? col / D : 0, d0 = active ? col % D : 0;
  const int tpg = D / 8;
  float wv[8], acc[8] = {0};
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = 1.f;
  if (active && w != nullptr && NORM != 0) load8f(w + (NORM == 1 ? d0 : col), wv);

  for (int l0 = blockIdx.x * RPI; l0 < L; l0 += gridDim.x * RPI) {
    uint4 rg[RPI], rx[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      rg[r] = make_uint4(0, 0, 0, 0);
      rx[r] = make_uint4(0, 0, 0, 0);
      if (active && l0 + r < L) {
        rg[r] = ldg_stream(dy + b * g_sb + static_cast<int64_t>(l0 + r) * g_sl + h * g_sh + d0);
        if (NORM != 0) rx[r] = ldg_stream(x + b * x_sb + static_cast<int64_t>(l0 + r) * x_sl + h * x_sh + d0);
      }
    }
    float g[RPI][8], xh[RPI][8], dot[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      const int l = l0 + r;
      float fd[8];
      unpack8(rg[r], fd);
      if (cosT != nullptr && l < L_rope && l < L && active) {
        float cs[8], sn[8];
        load8f(cosT + static_cast<size_t>(l) * D + d0, cs);
        load8f(sinT + static_cast<size_t>(l) * D + d0, sn);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          g[r][i] = fd[i] * cs[i] + fd[i + 1] * sn[i + 1];
          g[r][i + 1] = fd[i + 1] * cs[i + 1] - fd[i] * sn[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[r][i] = fd[i];
      }
      dot[r] = 0.f;
      if (NORM != 0) {
        float rs = 0.f;
        if (l < L) rs = NORM == 1 ? rstd_in[(static_cast<size_t>(b) * L + l) * H + h] : rstd_in[static_cast<size_t>(b) * L + l];
        float fx[8];
        unpack8(rx[r], fx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[r][i] = fx[i] * rs;
          dot[r] += g[r][i] * wv[i] * xh[r][i];
          acc[i] += g[r][i] * xh[r][i];
        }
      }
    }
    if (NORM == 1) {
#pragma unroll
      for (int r = 0; r < RPI; ++r)
        for (int o = tpg >> 1; o > 0; o >>= 1) dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], o);
    } else if (NORM == 2) {
      block_sum<RPI>(dot, red);
    }
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      const int l = l0 + r;
      if (!active || l >= L) continue;
      float o[8];
      if (NORM != 0) {
        const float rs = NORM == 1 ? rstd_in[(static_cast<size_t>(b) * L + l) * H + h] : rstd_in[static_cast<size_t>(b) * L + l];
        const float m = dot[r] / (NORM == 1 ? D : C);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rs * (g[r][i] * wv[i] - xh[r][i] * m);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = g[r][i];
      }
      stg_stream(dx + b * o_sb + static_cast<int64_t>(l) * o_sl + h * o_sh + d0, pack8(o));
    }
  }
  if (active && dw != nullptr && NORM != 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(dw + (NORM == 1 ? d0 : col) + i, acc[i]);
  }
}

// CTA size = ceil(C/8) rounded to a warp; (MAXT, RPI) chosen so the register budget 64K/MAXT holds RPI rows in flight.
struct RowCfg { int maxt, rpi; };
inline RowCfg row_cfg(int C) {
  const int threads = ((C / 8) + 31) / 32 * 32;
  if (threads <= 256) return {256, 4};
  if (threads <= 512) return {512, 4};
  if (threads <= 768) return {768, 2};
  return {1024, 1};
}
int row_launch_dims(int L, int C, int B, int RPI, dim3* grid, dim3* block) {
  VT_REQUIRE(C % 8 == 0 && C >= 8, VT_ERR_SHAPE, "row width %d must be a multiple of 8", C);
  const int threads = ((C / 8) + 31) / 32 * 32;
  VT_REQUIRE(threads <= 1024, VT_ERR_SHAPE, "row width %d exceeds 8192", C);
  VT_REQUIRE(B >= 1 && B <= 65535 && L >= 1, VT_ERR_SHAPE, "bad B=%d L=%d", B, L);
  // enough CTAs to fill 148 SMs at full residency, but several row groups per CTA so register accumulators pay off
  const int resident = 2048 / threads;
  long long want = 148LL * (resident > 0 ? resident : 1);
  long long per_batch = (want + B - 1) / B;
  long long max_blocks = (L + RPI - 1) / RPI;
  if (per_batch > max_blocks) per_batch = max_blocks;
  if (per_batch < 1) per_batch = 1;
  *grid = dim3(static_cast<unsigned>(per_batch), B);
  *block = dim3(threads);
  return 0;
}

// Expands KERNEL<..., MAXT, RPI><<<grid, block, 0, st>>>(ARGS) for the configuration row_cfg() selects.
#define VT_ROW_DISPATCH(C_, B_, L_, ST_, KERNEL, ...)                                   \
  do {                                                                                    \
    const RowCfg cfg_ = row_cfg(C_);                                                      \
    dim3 grid_, block_;                                                                   \
    if (int rc_ = row_launch_dims(L_, C_, B_, cfg_.rpi, &grid_, &block_)) return rc_;     \
    if (cfg_.maxt == 256) KERNEL(256, 4)<<<grid_, block_, 0, ST_>>>(__VA_ARGS__);        \
    else if (cfg_.maxt == 512) KERNEL(512, 4)<<<grid_, block_, 0, ST_>>>(__VA_ARGS__);   \
    else if (cfg_.maxt == 768) KERNEL(768, 2)<<<grid_, block_, 0, ST_>>>(__VA_ARGS__);   \
    else KERNEL(1024, 1)<<<grid_, block_, 0, ST_>>>(__VA_ARGS__);                        \
    VT_CHECK_CUDA(cudaGetLastError());                                                    \
  } while (0)

}  // namespace
}  // namespace vt

namespace vt {
// layernorm.cu: specialised LayerNorm(+modulate) kernels for the common row widths; return 1 = "no configuration".
int ln_fwd_fast(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta, const float* scale,
                const float* shift, int B, int L, int C, float eps, int x_dtype, cudaStream_t st);
int ln_bwd_fast(const void* dy, const void* x, const float* mean, const float* rstd, void* dx, const float* gamma,
                const float* beta, const float* scale, float* dgamma, float* dbeta, float* dscale, float* dshift, int B, int L,
                int C, int x_dtype, cudaStream_t st);
int rope_bwd_fast(const void* dy, const void* x, const float* rstd, void* dx, float* dw, const float* w, const float* c,
                  const float* s, const int64_t* gs, const int64_t* xs, const int64_t* os, int B, int L, int H, int D,
                  int L_rope, int norm_mode, cudaStream_t st);
int rope_fwd_fast(const void* x, void* y, float* rstd, const float* w, const float* c, const float* s, const int64_t* xs,
                  const int64_t* ys, int B, int L, int H, int D, int L_rope, int norm_mode, float eps, cudaStream_t st);
}  // namespace vt

using namespace vt;
using bf16 = __nv_bfloat16;

extern "C" {

int vt_ln_modulate_fwd(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta,
                       const float* scale, const float* shift, int B, int L, int C, float eps, int x_dtype, void* stream) {
  VT_REQUIRE(x && y, VT_ERR_NULL, "vt_ln_modulate_fwd: NULL argument");
  VT_REQUIRE(aligned16(x) && aligned16(y), VT_ERR_ALIGN, "x/y must be 16-byte aligned");
  VT_REQUIRE(x_dtype == 0 || x_dtype == 1, VT_ERR_DTYPE, "x_dtype %d (0 = bf16, 1 = fp32)", x_dtype);
  VT_REQUIRE(B >= 1 && L >= 1 && C >= 8 && C % 8 == 0, VT_ERR_SHAPE, "bad B=%d L=%d C=%d", B, L, C);
  if (getenv("VT_LN_GENERIC") == nullptr) {
    const int rc = ln_fwd_fast(x, y, mean, rstd, gamma, beta, scale, shift, B, L, C, eps, x_dtype, static_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
#define K16_(M, R) ln_modulate_fwd_kernel<bf16, M, R>
#define K32_(M, R) ln_modulate_fwd_kernel<float, M, R>
  if (x_dtype == 0)
    VT_ROW_DISPATCH(C, B, L, static_cast<cudaStream_t>(stream), K16_,
        static_cast<const bf16*>(x), static_cast<bf16*>(y), mean, rstd, gamma, beta, scale, shift, L, C, eps);
  else
    VT_ROW_DISPATCH(C, B, L, static_cast<cudaStream_t>(stream), K32_,
        static_cast<const float*>(x), static_cast<bf16*>(y), mean, rstd, gamma, beta, scale, shift, L, C, eps);
#undef K16_
#undef K32_
  return 0;
}

int vt_ln_modulate_bwd(const void* dy, const void* x, const float* mean, const float* rstd, void* dx,
                       const float* gamma, const float* beta, const float* scale, float* dgamma, float* dbeta,
                       float* dscale, float* dshift, int B, int L, int C, int x_dtype, void* stream) {
  VT_REQUIRE(dy && x && mean && rstd && dx, VT_ERR_NULL, "vt_ln_modulate_bwd: NULL argument");
  VT_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(dx), VT_ERR_ALIGN, "dy/x/dx must be 16-byte aligned");
  VT_REQUIRE(x_dtype == 0 || x_dtype == 1, VT_ERR_DTYPE, "x_dtype %d (0 = bf16, 1 = fp32)", x_dtype);
  VT_REQUIRE(B >= 1 && L >= 1 && C >= 8 && C % 8 == 0, VT_ERR_SHAPE, "bad B=%d L=%d C=%d", B, L, C);
  if (getenv("VT_LN_GENERIC") == nullptr) {
    const int rc = ln_bwd_fast(dy, x, mean, rstd, dx, gamma, beta, scale, dgamma, dbeta, dscale, dshift, B, L, C, x_dtype,
                               static_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
#define K16_(M, R) ln_modulate_bwd_kernel<bf16, M, R>
#define K32_(M, R) ln_modulate_bwd_kernel<float, M, R>
  if (x_dtype == 0)
    VT_ROW_DISPATCH(C, B, L, static_cast<cudaStream_t>(stream), K16_,
        static_cast<const bf16*>(dy), static_cast<const bf16*>(x), mean, rstd, static_cast<bf16*>(dx), gamma, beta,
        scale, dgamma, dbeta, dscale, dshift, L, C);
  else
    VT_ROW_DISPATCH(C, B, L, static_cast<cudaStream_t>(stream), K32_,
        static_cast<const bf16*>(dy), static_cast<const float*>(x), mean, rstd, static_cast<float*>(dx), gamma, beta,
        scale, dgamma, dbeta, dscale, dshift, L, C);
#undef K16_
#undef K32_
  return 0;
}

int vt_gate_residual_fwd(const void* x, const void* branch, void* y, const float* gate, int B, int L, int C,
                         int x_dtype, void* stream) {
  VT_REQUIRE(x && branch && y, VT_ERR_NULL, "vt_gate_residual_fwd: NULL argument");
  VT_REQUIRE(aligned16(x) && aligned16(branch) && aligned16(y), VT_ERR_ALIGN, "x/branch/y must be 16-byte aligned");
  VT_REQUIRE(x_dtype == 0 || x_dtype == 1, VT_ERR_DTYPE, "x_dtype %d (0 = bf16, 1 = fp32)", x_dtype);
#define K16_(M, R) gate_residual_fwd_kernel<bf16, M, R>
#define K32_(M, R) gate_residual_fwd_kernel<float, M, R>
  if (x_dtype == 0)
    VT_ROW_DISPATCH(C, B, L, static_cast<cudaStream_t>(stream), K16_,
        static_cast<const bf16*>(x), static_cast<const bf16*>(branch), static_cast<bf16*>(y), gate, L, C);
  else
    VT_ROW_DISPATCH(C, B, L, static_cast<cudaStream_t>(stream), K32_,
        static_cast<const float*>(x), static_cast<const bf16*>(branch), static_cast<float*>(y), gate, L, C);
#undef K16_
#undef K32_
  return 0;
}

int vt_gate_residual_bwd(const void* dy, const void* branch, void* dbranch, const float* gate, float* dgate, int B,
                         int L, int C, int x_dtype, void* stream) {
  VT_REQUIRE(dy && dbranch, VT_ERR_NULL, "vt_gate_residual_bwd: NULL argument");
  VT_REQUIRE(dgate == nullptr || branch != nullptr, VT_ERR_NULL, "dgate needs branch");
  VT_REQUIRE(aligned16(dy) && aligned16(dbranch) && aligned16(branch), VT_ERR_ALIGN, "buffers must be 16-byte aligned");
  VT_REQUIRE(x_dtype == 0 || x_dtype == 1, VT_ERR_DTYPE, "x_dtype %d (0 = bf16, 1 = fp32)", x_dtype);
#define K16_(M, R) gate_residual_bwd_kernel<bf16, M, R>
#define K32_(M, R) gate_residual_bwd_kernel<float, M, R>
  if (x_dtype == 0)
    VT_ROW_DISPATCH(C, B, L, static_cast<cudaStream_t>(stream), K16_,
        static_cast<const bf16*>(dy), static_cast<const bf16*>(branch), static_cast<bf16*>(dbranch), gate, dgate, L, C);
  else
    VT_ROW_DISPATCH(C, B, L, static_cast<cudaStream_t>(stream), K32_,
        static_cast<const float*>(dy), static_cast<const bf16*>(branch), static_cast<bf16*>(dbranch), gate, dgate, L, C);
#undef K16_
#undef K32_
  return 0;
}

static int check_rope_args(const int64_t* s, int H, int D, const char* name) {
  VT_REQUIRE(s != nullptr, VT_ERR_NULL, "%s strides are NULL", name);
  VT_REQUIRE(s[0] % 8 == 0 && s[1] % 8 == 0 && s[2] % 8 == 0, VT_ERR_ALIGN, "%s strides must be multiples of 8 elements", name);
  VT_REQUIRE(D % 8 == 0 && (D == 64 || D == 128), VT_ERR_DTYPE, "head dim %d unsupported", D);
  (void)H;
  return 0;
}

int vt_qk_rmsnorm_rope_fwd(const void* x, void* y, float* rstd_out, const float* w, const float* cos, const float* sin,
                           const int64_t* x_strides, const int64_t* y_strides, int B, int L, int H, int D, int L_rope,
                           int norm_mode, float eps, void* stream) {
  VT_REQUIRE(x && y, VT_ERR_NULL, "vt_qk_rmsnorm_rope_fwd: NULL argument");
  VT_REQUIRE(aligned16(x) && aligned16(y), VT_ERR_ALIGN, "x/y must be 16-byte aligned");
  VT_REQUIRE((cos == nullptr) == (sin == nullptr), VT_ERR_NULL, "cos and sin must both be given or both NULL");
  VT_REQUIRE(norm_mode >= 0 && norm_mode <= 2, VT_ERR_SHAPE, "norm_mode %d", norm_mode);
  if (int rc = check_rope_args(x_strides, H, D, "x")) return rc;
  if (int rc = check_rope_args(y_strides, H, D, "y")) return rc;
  auto st = static_cast<cudaStream_t>(stream);
  if (getenv("VT_LN_GENERIC") == nullptr) {
    const int rc = rope_fwd_fast(x, y, rstd_out, w, cos, sin, x_strides, y_strides, B, L, H, D, L_rope, norm_mode, eps, st);
    if (rc <= 0) return rc;
  }
#define VT_RR_ARGS static_cast<const bf16*>(x), static_cast<bf16*>(y), rstd_out, w, cos, sin, x_strides[0], x_strides[1], \
                   x_strides[2], y_strides[0], y_strides[1], y_strides[2], L, H, D, L_rope, eps
#define K0_(M, R) rmsnorm_rope_fwd_kernel<0, M, R>
#define K1_(M, R) rmsnorm_rope_fwd_kernel<1, M, R>
#define K2_(M, R) rmsnorm_rope_fwd_kernel<2, M, R>
  if (norm_mode == 0) VT_ROW_DISPATCH(H * D, B, L, st, K0_, VT_RR_ARGS);
  else if (norm_mode == 1) VT_ROW_DISPATCH(H * D, B, L, st, K1_, VT_RR_ARGS);
  else VT_ROW_DISPATCH(H * D, B, L, st, K2_, VT_RR_ARGS);
#undef K0_
#undef K1_
#undef K2_
#undef VT_RR_ARGS
  return 0;
}

int vt_qk_rmsnorm_rope_bwd(const void* dy, const void* x, const float* rstd, void* dx, float* dw_accum, const float* w,
                           const float* cos, const float* sin, const int64_t* dy_strides, const int64_t* x_strides,
                           const int64_t* dx_strides, int B, int L, int H, int D, int L_rope, int norm_mode,
                           void* stream) {
  VT_REQUIRE(dy && dx, VT_ERR_NULL, "vt_qk_rmsnorm_rope_bwd: NULL argument");
  VT_REQUIRE(norm_mode == 0 || (x && rstd), VT_ERR_NULL, "norm backward needs x and rstd");
  VT_REQUIRE(aligned16(dy) && aligned16(dx) && aligned16(x), VT_ERR_ALIGN, "buffers must be 16-byte aligned");
  VT_REQUIRE((cos == nullptr) == (sin == nullptr), VT_ERR_NULL, "cos and sin must both be given or both NULL");
  VT_REQUIRE(norm_mode >= 0 && norm_mode <= 2, VT_ERR_SHAPE, "norm_mode %d", norm_mode);
  if (int rc = check_rope_args(dy_strides, H, D, "dy")) return rc;
  if (int rc = check_rope_args(dx_strides, H, D, "dx")) return rc;
  const int64_t zero3[3] = {0, 0, 0};
  const int64_t* xs = x_strides ? x_strides : zero3;
  auto st = static_cast<cudaStream_t>(stream);
  if (getenv("VT_LN_GENERIC") == nullptr) {
    const int rc = rope_bwd_fast(dy, x, rstd, dx, dw_accum, w, cos, sin, dy_strides, xs, dx_strides, B, L, H, D, L_rope,
                                 norm_mode, st);
    if (rc <= 0) return rc;
  }
#define VT_RB_ARGS static_cast<const bf16*>(dy), static_cast<const bf16*>(x), rstd, static_cast<bf16*>(dx), dw_accum, w, cos, \
                   sin, dy_strides[0], dy_strides[1], dy_strides[2], xs[0], xs[1], xs[2], dx_strides[0], dx_strides[1],        \
                   dx_strides[2], L, H, D, L_rope
#define K0_(M, R) rmsnorm_rope_bwd_kernel<0, M, R>
#define K1_(M, R) rmsnorm_rope_bwd_kernel<1, M, R>
#define K2_(M, R) rmsnorm_rope_bwd_kernel<2, M, R>
  if (norm_mode == 0) VT_ROW_DISPATCH(H * D, B, L, st, K0_, VT_RB_ARGS);
  else if (norm_mode == 1) VT_ROW_DISPATCH(H * D, B, L, st, K1_, VT_RB_ARGS);
  else VT_ROW_DISPATCH(H * D, B, L, st, K2_, VT_RB_ARGS);
#undef K0_
#undef K1_
#undef K2_
#undef VT_RB_ARGS
  return 0;
}

}  // extern "C"
