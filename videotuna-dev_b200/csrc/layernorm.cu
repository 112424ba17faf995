// layernorm.cu — LayerNorm (+affine) + adaLN modulate, forward and backward, for the row widths the denoisers use.
//
// Second-generation layout of the kernels in rowwise.cu (which stay as the fallback for other widths). The first layout
// gave every thread 8 columns and every CTA (C/8 = 384..640 threads) RPI rows per iteration: 95-128 registers x 384
// threads left ONE CTA per SM and 24 KB of loads in flight — 32 % of the measured HBM bandwidth on the HunyuanVideo K1
// activation (2.1 of 6.45 TB/s). Here a small CTA (64 or 128 threads) owns one row at a time, every thread holds VPT
// 16-byte vectors of it (columns (j * T + t) * 8: each j is one coalesced sweep), the next row's loads are issued before
// the current row is reduced (double buffering), statistics take two cheap 4-warp block sums per row (mean, then the
// centred second moment), and the per-column factors (gamma, beta, 1 + scale, shift) live pre-combined in shared memory:
//   y = xhat * mul + add,   mul = gamma * (1 + scale),   add = beta * (1 + scale) + shift.
// 6-8 CTAs per SM with two rows each in flight keep ~100 KB of loads outstanding per SM.
// Replaces: hunyuan modulate(LayerNorm(x)) (modulate_layers.py:31-49, models.py:161-164), wan norm(x).float()*(1+e)+e
// (wan/modules/model.py:294-296, fp32 residual stream), lvdm nn.LayerNorm (lvdm/modules/attention.py:299-310).
#include <cstdlib>
#include <cuda_bf16.h>

#include "capi_util.h"
#include "sm100_ptx.cuh"

namespace vt {
namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ uint4 ldg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// 8 consecutive row elements in their storage type
template <typename T> struct Vec8;
template <> struct Vec8<bf16> { uint4 a; };
template <> struct Vec8<float> { uint4 a, b; };
__device__ __forceinline__ void load8(const bf16* p, Vec8<bf16>& r) { r.a = ldg16(p); }
__device__ __forceinline__ void load8(const float* p, Vec8<float>& r) { r.a = ldg16(p); r.b = ldg16(p + 4); }
__device__ __forceinline__ void zero8(Vec8<bf16>& r) { r.a = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ void zero8(Vec8<float>& r) { r.a = r.b = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ void unpack(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void unpack(const Vec8<bf16>& r, float* f) { unpack(r.a, f); }
__device__ __forceinline__ void unpack(const Vec8<float>& r, float* f) {
  f[0] = __uint_as_float(r.a.x); f[1] = __uint_as_float(r.a.y); f[2] = __uint_as_float(r.a.z); f[3] = __uint_as_float(r.a.w);
  f[4] = __uint_as_float(r.b.x); f[5] = __uint_as_float(r.b.y); f[6] = __uint_as_float(r.b.z); f[7] = __uint_as_float(r.b.w);
}
__device__ __forceinline__ uint4 pack(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ void store8(bf16* p, const float* f) { stg16(p, pack(f)); }
__device__ __forceinline__ void store8(float* p, const float* f) {
  stg16(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
  stg16(p + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
}
__device__ __forceinline__ void lds8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <int T, int NV>
__device__ __forceinline__ void block_sum(float* v, float* red /* [2][T/32][NV] */, int parity) {
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  constexpr int NW = T / 32;
  if (NW == 1) return;
  float* r = red + parity * NW * NV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) r[warp * NV + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float t = r[i];
#pragma unroll
    for (int w = 1; w < NW; ++w) t += r[w * NV + i];
    v[i] = t;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <typename XT, int T, int VPT>
__global__ void __launch_bounds__(T) ln_fwd_kernel(const XT* __restrict__ x, bf16* __restrict__ y, float* __restrict__ mean_out,
                                                   float* __restrict__ rstd_out, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, const float* __restrict__ scale,
                                                   const float* __restrict__ shift, int L, int C, float eps) {
  extern __shared__ __align__(16) float sm[];  // mul[CP], add[CP], red[2 uses][2 parities][T/32]   (CP = T * VPT * 8 >= C)
  constexpr int CP = T * VPT * 8;
  float* mul = sm;
  float* add = sm + CP;
  float* red = sm + 2 * CP;
  const int b = blockIdx.y, t = threadIdx.x;
  for (int c = t; c < CP; c += T) {
    float g = 1.f, be = 0.f, s1 = 1.f, sh = 0.f;
    if (c < C) {
      if (gamma) g = gamma[c];
      if (beta) be = beta[c];
      if (scale) s1 = 1.f + scale[static_cast<size_t>(b) * C + c];
      if (shift) sh = shift[static_cast<size_t>(b) * C + c];
    }
    mul[c] = g * s1;
    add[c] = be * s1 + sh;
  }
  __syncthreads();
  bool act[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) act[j] = (j * T + t) * 8 < C;
  const size_t base = static_cast<size_t>(b) * L;
  Vec8<XT> cur[VPT], nxt[VPT];
  int l = blockIdx.x;
  if (l < L) {
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      zero8(cur[j]);
      if (act[j]) load8(x + (base + l) * C + (j * T + t) * 8, cur[j]);
    }
  }
  int parity = 0;
  for (; l < L; l += gridDim.x, parity ^= 1) {
    const int ln = l + gridDim.x;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      zero8(nxt[j]);
      if (ln < L && act[j]) load8(x + (base + ln) * C + (j * T + t) * 8, nxt[j]);
    }
    // two plain block sums (mean, then centred second moment): 2 x (5 shuffles + one 4-warp barrier) per row — cheaper
    // than a single Chan merge of (n, mean, M2) triples, whose unequal-count form costs a division per butterfly stage
    float f[VPT][8];
    float s[1] = {0.f};
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      unpack(cur[j], f[j]);  // inactive vectors are zeros
#pragma unroll
      for (int i = 0; i < 8; ++i) s[0] += f[j][i];
    }
    block_sum<T, 1>(s, red, parity);
    const float mean = s[0] / C;
    float q[1] = {0.f};
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      if (act[j]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = f[j][i] - mean;
          q[0] += d * d;
        }
      }
    }
    block_sum<T, 1>(q, red + 2 * (T / 32), parity);
    const float rstd = rsqrtf(q[0] / C + eps);
    if (t == 0) {
      if (mean_out) mean_out[base + l] = mean;
      if (rstd_out) rstd_out[base + l] = rstd;
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      if (!act[j]) continue;
      const int col = (j * T + t) * 8;
      float m8[8], a8[8], o[8];
      lds8(mul + col, m8);
      lds8(add + col, a8);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf((f[j][i] - mean) * rstd, m8[i], a8[i]);
      stg16(y + (base + l) * C + col, pack(o));
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) cur[j] = nxt[j];
  }
}

// Two rows per iteration (bf16 rows: 2 x VPT vectors current + 2 x VPT prefetched per thread). ncu on the one-row kernel at
// K1: shared-memory data pipe 67 % busy, 53 % of the stall samples on short-scoreboard (LDS) waits — every 16 bytes of x
// cost 64 bytes of mul/add reads. Here each mul/add vector read serves two rows and each block-sum round (shuffles +
// barrier) reduces two rows' values at once, halving both.
template <typename XT, int T, int VPT>
__global__ void __launch_bounds__(T, 5) ln_fwd2_kernel(const XT* __restrict__ x, bf16* __restrict__ y, float* __restrict__ mean_out,
                                                    float* __restrict__ rstd_out, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, int L, int C, float eps) {
  extern __shared__ __align__(16) float sm[];  // mul[CP], add[CP], red[2 uses][2 parities][T/32][2 rows]
  constexpr int CP = T * VPT * 8;
  float* mul = sm;
  float* add = sm + CP;
  float* red = sm + 2 * CP;
  const int b = blockIdx.y, t = threadIdx.x;
  for (int c = t; c < CP; c += T) {
    float g = 1.f, be = 0.f, s1 = 1.f, sh = 0.f;
    if (c < C) {
      if (gamma) g = gamma[c];
      if (beta) be = beta[c];
      if (scale) s1 = 1.f + scale[static_cast<size_t>(b) * C + c];
      if (shift) sh = shift[static_cast<size_t>(b) * C + c];
    }
    mul[c] = g * s1;
    add[c] = be * s1 + sh;
  }
  __syncthreads();
  bool act[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) act[j] = (j * T + t) * 8 < C;
  const size_t base = static_cast<size_t>(b) * L;
  Vec8<XT> cur[2][VPT], nxt[2][VPT];
  int l = 2 * blockIdx.x;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      zero8(cur[r][j]);
      if (l + r < L && act[j]) load8(x + (base + l + r) * C + (j * T + t) * 8, cur[r][j]);
    }
  int parity = 0;
  for (; l < L; l += 2 * gridDim.x, parity ^= 1) {
    const int ln = l + 2 * gridDim.x;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        zero8(nxt[r][j]);
        if (ln + r < L && act[j]) load8(x + (base + ln + r) * C + (j * T + t) * 8, nxt[r][j]);
      }
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        float f[8];
        unpack(cur[r][j], f);  // inactive vectors and a missing second row are zeros
#pragma unroll
        for (int i = 0; i < 8; ++i) s[r] += f[i];
      }
    block_sum<T, 2>(s, red, parity);
    const float mean[2] = {s[0] / C, s[1] / C};
    float q[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        if (act[j]) {
          float f[8];
          unpack(cur[r][j], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float d = f[i] - mean[r];
            q[r] = fmaf(d, d, q[r]);
          }
        }
      }
    block_sum<T, 2>(q, red + 2 * 2 * (T / 32), parity);
    const float rstd[2] = {rsqrtf(q[0] / C + eps), rsqrtf(q[1] / C + eps)};
    const bool two = l + 1 < L;
    if (t < 2 && (t == 0 || two)) {
      if (mean_out) mean_out[base + l + t] = mean[t];
      if (rstd_out) rstd_out[base + l + t] = rstd[t];
    }
    // xhat * mul + add = x * (rstd * mul) + (add - mean * rstd * mul)
    const float nb[2] = {-mean[0] * rstd[0], -mean[1] * rstd[1]};
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      if (!act[j]) continue;
      const int col = (j * T + t) * 8;
      float m8[8], a8[8];
      lds8(mul + col, m8);
      lds8(add + col, a8);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 1 && !two) break;
        float f[8], o[8];
        unpack(cur[r][j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(fmaf(f[i], rstd[r], nb[r]), m8[i], a8[i]);
        stg16(y + (base + l + r) * C + col, pack(o));
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < VPT; ++j) cur[r][j] = nxt[r][j];
  }
}

// fp32 rows (Wan's residual stream, 20 KB per row): two rows in registers do not fit, and the one-row kernel sits at 72 %
// (this one: 79 %; two 100 KB CTAs per SM, grid = one resident wave — 3 per SM of grid drops to 60 %).
// Here the rows arrive in a shared-memory ring by 1-D bulk asynchronous copies two rows ahead (like the RoPE kernels), so
// no register is tied up by data in flight; 256 threads per CTA, the factor table (2 x C fp32) once per long-lived CTA.
__device__ __forceinline__ void ln_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int T, int VPT>
__global__ void __launch_bounds__(T) ln_fwd_ring_f32_kernel(const float* __restrict__ x, bf16* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            int L, int C, float eps) {
  constexpr int S = 3;
  extern __shared__ __align__(128) float ring[];  // S x row[C] | mul[C] | add[C]
  __shared__ __align__(8) uint64_t full[S];
  __shared__ float red[4 * (T / 32)];
  float* mul = ring + S * C;
  float* add = mul + C;
  const int b = blockIdx.y, t = threadIdx.x;
  for (int c = t; c < C; c += T) {
    const float g = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const float s1 = scale ? 1.f + scale[static_cast<size_t>(b) * C + c] : 1.f;
    const float sh = shift ? shift[static_cast<size_t>(b) * C + c] : 0.f;
    mul[c] = g * s1;
    add[c] = be * s1 + sh;
  }
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  bool act[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) act[j] = (j * T + t) * 8 < C;
  const size_t base = static_cast<size_t>(b) * L;
  const int n_row = blockIdx.x < static_cast<unsigned>(L) ? (L - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  auto issue = [&](int k) {  // thread 0 only
    const int l = blockIdx.x + k * gridDim.x;
    mbar_arrive_expect_tx(&full[k % S], C * 4);
    ln_bulk_g2s(ring + (k % S) * C, x + (base + l) * C, C * 4, &full[k % S]);
  };
  if (t == 0)
    for (int k = 0; k < S - 1 && k < n_row; ++k) issue(k);
  for (int k = 0; k < n_row; ++k) {
    const int l = blockIdx.x + k * gridDim.x;
    __syncthreads();  // every thread is done with stage (k - 1) % S
    if (t == 0 && k + S - 1 < n_row) issue(k + S - 1);
    mbar_wait(&full[k % S], (k / S) & 1, 0x7e03);
    const float* row = ring + (k % S) * C;
    float f[VPT][8];
    float s[1] = {0.f};
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[j][i] = 0.f;
      if (act[j]) lds8(row + (j * T + t) * 8, f[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[0] += f[j][i];
    }
    block_sum<T, 1>(s, red, k & 1);
    const float mean = s[0] / C;
    float q[1] = {0.f};
#pragma unroll
    for (int j = 0; j < VPT; ++j)
      if (act[j]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = f[j][i] - mean;
          q[0] = fmaf(d, d, q[0]);
        }
      }
    block_sum<T, 1>(q, red + 2 * (T / 32), k & 1);
    const float rstd = rsqrtf(q[0] / C + eps);
    if (t == 0) {
      if (mean_out) mean_out[base + l] = mean;
      if (rstd_out) rstd_out[base + l] = rstd;
    }
    const float nb = -mean * rstd;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      if (!act[j]) continue;
      const int col = (j * T + t) * 8;
      float m8[8], a8[8], o[8];
      lds8(mul + col, m8);
      lds8(add + col, a8);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(fmaf(f[j][i], rstd, nb), m8[i], a8[i]);
      stg16(y + (base + l) * C + col, pack(o));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward.  xh = (x - mean) * rstd;  y = xh * mul + add  with  mul = gamma * s1, add = beta * s1 + shift, s1 = 1 + scale
//   gh = dy * mul;  dx = rstd * (gh - mean_C(gh) - xh * mean_C(gh * xh))
//   dshift += dy;  dscale += dy * (xh * gamma + beta);  dbeta += dy * s1;  dgamma += dy * s1 * xh
// Column sums accumulate in registers over the CTA's rows (compile-time MODE keeps only the needed ones) and are
// flushed with one atomicAdd per column per CTA.   MODE bit 0: dscale/dshift, bit 1: dgamma/dbeta.
// ------------------------------------------------------------------------------------------------------------------
template <typename XT, int T, int VPT, int MODE>
__global__ void __launch_bounds__(T) ln_bwd_kernel(const bf16* __restrict__ dy, const XT* __restrict__ x,
                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                   XT* __restrict__ dx, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, const float* __restrict__ scale,
                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                   float* __restrict__ dscale, float* __restrict__ dshift, int L, int C) {
  extern __shared__ __align__(16) float sm[];  // mul[CP], (MODE & 1) gam[CP], bet[CP], (MODE & 2) s1[CP], red[2][T/32][2]
  constexpr int CP = T * VPT * 8;
  constexpr bool MOD = (MODE & 1) != 0, AFF = (MODE & 2) != 0;
  float* mul = sm;
  float* gam = mul + CP;                       // gamma          (MOD only)
  float* bet = gam + (MOD ? CP : 0);           // beta           (MOD only)
  float* s1v = bet + (MOD ? CP : 0);           // 1 + scale      (AFF only)
  float* red = s1v + (AFF ? CP : 0);
  const int b = blockIdx.y, t = threadIdx.x;
  for (int c = t; c < CP; c += T) {
    float g = 1.f, be = 0.f, s1 = 1.f;
    if (c < C) {
      if (gamma) g = gamma[c];
      if (beta) be = beta[c];
      if (scale) s1 = 1.f + scale[static_cast<size_t>(b) * C + c];
    }
    mul[c] = g * s1;
    if constexpr (MOD) { gam[c] = g; bet[c] = be; }
    if constexpr (AFF) s1v[c] = s1;
  }
  __syncthreads();
  bool act[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) act[j] = (j * T + t) * 8 < C;
  float a_dshift[MOD ? VPT : 1][8], a_dscale[MOD ? VPT : 1][8], a_dbeta[AFF ? VPT : 1][8], a_dgamma[AFF ? VPT : 1][8];
#pragma unroll
  for (int j = 0; j < VPT; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (MOD) { a_dshift[j][i] = 0.f; a_dscale[j][i] = 0.f; }
      if constexpr (AFF) { a_dbeta[j][i] = 0.f; a_dgamma[j][i] = 0.f; }
    }
  const size_t base = static_cast<size_t>(b) * L;
  const float invC = 1.f / C;
  uint4 cdy[VPT], ndy[VPT];
  Vec8<XT> cx[VPT], nx[VPT];
  int l = blockIdx.x;
  if (l < L) {
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      cdy[j] = make_uint4(0, 0, 0, 0);
      zero8(cx[j]);
      if (act[j]) {
        cdy[j] = ldg16(dy + (base + l) * C + (j * T + t) * 8);
        load8(x + (base + l) * C + (j * T + t) * 8, cx[j]);
      }
    }
  }
  int parity = 0;
  for (; l < L; l += gridDim.x, parity ^= 1) {
    const int ln = l + gridDim.x;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      ndy[j] = make_uint4(0, 0, 0, 0);
      zero8(nx[j]);
      if (ln < L && act[j]) {
        ndy[j] = ldg16(dy + (base + ln) * C + (j * T + t) * 8);
        load8(x + (base + ln) * C + (j * T + t) * 8, nx[j]);
      }
    }
    const float mu = mean[base + l], rs = rstd[base + l];
    float gh[VPT][8], xh[VPT][8];
    float sums[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      float fdy[8], fx[8], m8[8];
      unpack(cdy[j], fdy);
      unpack(cx[j], fx);
      const int col = (j * T + t) * 8;
      lds8(mul + col, m8);
      float g8[8], b8[8], s8[8];
      if constexpr (MOD) { lds8(gam + col, g8); lds8(bet + col, b8); }
      if constexpr (AFF) lds8(s1v + col, s8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[j][i] = act[j] ? (fx[i] - mu) * rs : 0.f;
        gh[j][i] = fdy[i] * m8[i];
        sums[0] += gh[j][i];
        sums[1] += gh[j][i] * xh[j][i];
        if constexpr (MOD) {
          a_dshift[j][i] += fdy[i];
          a_dscale[j][i] += fdy[i] * fmaf(xh[j][i], g8[i], b8[i]);
        }
        if constexpr (AFF) {
          const float gm = fdy[i] * s8[i];
          a_dbeta[j][i] += gm;
          a_dgamma[j][i] += gm * xh[j][i];
        }
      }
    }
    block_sum<T, 2>(sums, red, parity);
    const float m1 = sums[0] * invC, m2 = sums[1] * invC;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      if (!act[j]) continue;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = rs * (gh[j][i] - m1 - xh[j][i] * m2);
      store8(dx + (base + l) * C + (j * T + t) * 8, o);
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) { cdy[j] = ndy[j]; cx[j] = nx[j]; }
  }
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    if (!act[j]) continue;
    const int col = (j * T + t) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (MOD) {
        atomicAdd(dshift + static_cast<size_t>(b) * C + col + i, a_dshift[j][i]);
        atomicAdd(dscale + static_cast<size_t>(b) * C + col + i, a_dscale[j][i]);
      }
      if constexpr (AFF) {
        atomicAdd(dbeta + col + i, a_dbeta[j][i]);
        atomicAdd(dgamma + col + i, a_dgamma[j][i]);
      }
    }
  }
}

struct LnCfg { int T, VPT; };
inline bool ln_cfg(int C, LnCfg* c) {
  if (C % 8 != 0 || C < 8) return false;
  if (C <= 512) { *c = {64, 1}; return true; }
  if (C <= 1024) { *c = {128, 1}; return true; }
  if (C <= 2048) { *c = {128, 2}; return true; }
  if (C <= 3072) { *c = {128, 3}; return true; }
  if (C > 4096 && C <= 5120) { *c = {128, 5}; return true; }
  return false;  // other widths: the generic kernels in rowwise.cu
}
inline unsigned ln_grid_x(int L, int B, int ctas_per_sm) {
  long long want = (148LL * ctas_per_sm + B - 1) / B;
  if (want > L) want = L;
  if (want < 1) want = 1;
  return static_cast<unsigned>(want);
}

template <typename XT, int T, int VPT>
int launch_fwd(const XT* x, bf16* y, float* mean, float* rstd, const float* gamma, const float* beta, const float* scale,
               const float* shift, int B, int L, int C, float eps, cudaStream_t st) {
  constexpr int CP = T * VPT * 8;
  // bf16 rows: two rows per iteration (ln_fwd2_kernel); fp32 rows (Wan residual stream) would need 4 x VPT x 8 data registers
  // per thread for that and stay on the one-row kernel. VT_LN_ROWS=1 selects the one-row kernel for A/B measurements.
  static const bool one_row = sizeof(XT) == 4 || (getenv("VT_LN_ROWS") != nullptr && atoi(getenv("VT_LN_ROWS")) == 1);
  const int smem = (2 * CP + (one_row ? 4 : 8) * (T / 32)) * 4;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    VT_CHECK_CUDA(cudaFuncSetAttribute(ln_fwd_kernel<XT, T, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    VT_CHECK_CUDA(cudaFuncSetAttribute(ln_fwd2_kernel<XT, T, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  if (one_row) {
    static const int want1 = getenv("VT_LN1_CTAS") != nullptr ? atoi(getenv("VT_LN1_CTAS")) : 8;
    dim3 grid(ln_grid_x(L, B, want1), B);
    ln_fwd_kernel<XT, T, VPT><<<grid, T, smem, st>>>(x, y, mean, rstd, gamma, beta, scale, shift, L, C, eps);
  } else {
    // CTAs per SM of grid (not of residency): the row -> CTA assignment is static, and many short-lived CTAs let the hardware
    // scheduler even out SM-to-SM differences (measured at K1: 5 per SM 90 %, 32 per SM 97 % of the copy bandwidth)
    static const int want = getenv("VT_LN_CTAS") != nullptr ? atoi(getenv("VT_LN_CTAS")) : 32;
    dim3 grid(ln_grid_x((L + 1) / 2, B, want), B);
    ln_fwd2_kernel<XT, T, VPT><<<grid, T, smem, st>>>(x, y, mean, rstd, gamma, beta, scale, shift, L, C, eps);
  }
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename XT, int T, int VPT, int MODE>
int launch_bwd(const bf16* dy, const XT* x, const float* mean, const float* rstd, XT* dx, const float* gamma,
               const float* beta, const float* scale, float* dgamma, float* dbeta, float* dscale, float* dshift, int B, int L,
               int C, cudaStream_t st) {
  constexpr int CP = T * VPT * 8;
  const int smem = (CP * (1 + ((MODE & 1) ? 2 : 0) + ((MODE & 2) ? 1 : 0)) + 2 * (T / 32) * 2) * 4;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    VT_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<XT, T, VPT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  // grid CTAs per SM (each flushes its column sums with atomics at the end): measured best 64 for bf16 rows (K1: 89 -> 93.5 %
  // of the copy bandwidth) and 4 for the fp32 rows of the Wan stream (98.7 %; 64 there: 87 %)
  static const int wantb = getenv("VT_LNB_CTAS") != nullptr ? atoi(getenv("VT_LNB_CTAS")) : (sizeof(XT) == 2 ? 64 : 4);
  dim3 grid(ln_grid_x(L, B, wantb), B);
  ln_bwd_kernel<XT, T, VPT, MODE><<<grid, T, smem, st>>>(dy, x, mean, rstd, dx, gamma, beta, scale, dgamma, dbeta, dscale,
                                                        dshift, L, C);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename XT, int T, int VPT>
int launch_bwd_mode(int mode, const bf16* dy, const XT* x, const float* mean, const float* rstd, XT* dx, const float* gamma,
                    const float* beta, const float* scale, float* dgamma, float* dbeta, float* dscale, float* dshift, int B,
                    int L, int C, cudaStream_t st) {
  switch (mode) {
    case 0: return launch_bwd<XT, T, VPT, 0>(dy, x, mean, rstd, dx, gamma, beta, scale, dgamma, dbeta, dscale, dshift, B, L, C, st);
    case 1: return launch_bwd<XT, T, VPT, 1>(dy, x, mean, rstd, dx, gamma, beta, scale, dgamma, dbeta, dscale, dshift, B, L, C, st);
    case 2: return launch_bwd<XT, T, VPT, 2>(dy, x, mean, rstd, dx, gamma, beta, scale, dgamma, dbeta, dscale, dshift, B, L, C, st);
    default: return launch_bwd<XT, T, VPT, 3>(dy, x, mean, rstd, dx, gamma, beta, scale, dgamma, dbeta, dscale, dshift, B, L, C, st);
  }
}

#define VT_LN_CFGS(X) X(64, 1) X(128, 1) X(128, 2) X(128, 3) X(128, 5)

}  // namespace

// Returns 1 when no specialised configuration exists for this width (caller falls back to rowwise.cu), 0 on success,
// < 0 on error.
int ln_fwd_fast(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta, const float* scale,
                const float* shift, int B, int L, int C, float eps, int x_dtype, cudaStream_t st) {
  LnCfg c;
  if (!ln_cfg(C, &c) || B > 65535) return 1;
  // fp32 rows wider than 4096 (Wan: 5120): bulk-copy ring kernel; VT_LN_RING=0 keeps the one-row register kernel (A/B)
  static const bool ring_ok = !(getenv("VT_LN_RING") != nullptr && atoi(getenv("VT_LN_RING")) == 0);
  if (ring_ok && x_dtype == 1 && C > 4096 && C <= 6144 && C % 8 == 0 && aligned16(x)) {
    constexpr int T = 256, VPT = 3;  // 512 x 2 measured slower (68 vs 79 %)
    const int smem = (3 * C + 2 * C) * 4;
    static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
    if (first_on_device(&cfg_site)) {
      VT_CHECK_CUDA(cudaFuncSetAttribute(ln_fwd_ring_f32_kernel<T, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         5 * 6144 * 4));
    }
    static const int want = getenv("VT_LNR_CTAS") != nullptr ? atoi(getenv("VT_LNR_CTAS")) : 2;
    dim3 grid(ln_grid_x(L, B, want), B);
    ln_fwd_ring_f32_kernel<T, VPT><<<grid, T, smem, st>>>(static_cast<const float*>(x), static_cast<bf16*>(y), mean, rstd, gamma,
                                                         beta, scale, shift, L, C, eps);
    VT_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
#define X(T_, V_)                                                                                                        \
  if (c.T == T_ && c.VPT == V_)                                                                                          \
    return x_dtype == 0 ? launch_fwd<bf16, T_, V_>(static_cast<const bf16*>(x), static_cast<bf16*>(y), mean, rstd, gamma, beta, \
                                                  scale, shift, B, L, C, eps, st)                                       \
                        : launch_fwd<float, T_, V_>(static_cast<const float*>(x), static_cast<bf16*>(y), mean, rstd, gamma,    \
                                                   beta, scale, shift, B, L, C, eps, st);
  VT_LN_CFGS(X)
#undef X
  return 1;
}

int ln_bwd_fast(const void* dy, const void* x, const float* mean, const float* rstd, void* dx, const float* gamma,
                const float* beta, const float* scale, float* dgamma, float* dbeta, float* dscale, float* dshift, int B, int L,
                int C, int x_dtype, cudaStream_t st) {
  LnCfg c;
  if (!ln_cfg(C, &c) || B > 65535) return 1;
  const bool mod = dscale != nullptr || dshift != nullptr, aff = dgamma != nullptr || dbeta != nullptr;
  if ((mod && !(dscale && dshift)) || (aff && !(dgamma && dbeta))) return 1;  // half-requested pairs: generic kernel
  const int mode = (mod ? 1 : 0) | (aff ? 2 : 0);
#define X(T_, V_)                                                                                                        \
  if (c.T == T_ && c.VPT == V_)                                                                                          \
    return x_dtype == 0 ? launch_bwd_mode<bf16, T_, V_>(mode, static_cast<const bf16*>(dy), static_cast<const bf16*>(x), mean, \
                                                       rstd, static_cast<bf16*>(dx), gamma, beta, scale, dgamma, dbeta, dscale, \
                                                       dshift, B, L, C, st)                                              \
                        : launch_bwd_mode<float, T_, V_>(mode, static_cast<const bf16*>(dy), static_cast<const float*>(x),     \
                                                        mean, rstd, static_cast<float*>(dx), gamma, beta, scale, dgamma, dbeta, \
                                                        dscale, dshift, B, L, C, st);
  VT_LN_CFGS(X)
#undef X
  return 1;
}

}  // namespace vt

// =====================================================================================================================
// Fused QK-RMSNorm + RoPE forward in the same layout (second generation of rmsnorm_rope_fwd_kernel in rowwise.cu, which
// ran at 34 % of the HBM roofline on the K1 q tensor for the same occupancy reason). One 128-thread CTA per token; thread
// t holds VPT vectors at columns (j * 128 + t) * 8, i.e. head j * (1024 / D) + t / (D / 8), always the same 8 head-dim
// positions d0 = (t % (D / 8)) * 8 — so its cos / sin slice is loaded once per token and reused for all its heads, the
// per-head sum of squares is a shuffle over D / 8 lanes, and the next token's loads are issued before this one is
// normalised. NORM: 0 none, 1 per head (weight (D)), 2 whole token (weight (H * D), one 4-warp block sum).
// =====================================================================================================================
namespace vt {
namespace {

template <int NORM, int VPT, int D>
__global__ void __launch_bounds__(128, (VPT <= 3 ? 5 : 3)) rope_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, float* __restrict__ rstd_out,
                                                       const float* __restrict__ w, const float* __restrict__ cosT,
                                                       const float* __restrict__ sinT, int64_t x_sb, int64_t x_sl,
                                                       int64_t x_sh, int64_t y_sb, int64_t y_sl, int64_t y_sh, int L, int H,
                                                       int L_rope, float eps) {
  constexpr int T = 128, LPH = D / 8, HPS = 1024 / D;  // lanes per head, heads per 128-thread sweep
  __shared__ float red[2 * (T / 32)];
  const int b = blockIdx.y, t = threadIdx.x;
  const int d0 = (t % LPH) * 8, hsub = t / LPH;
  const int C = H * D;
  float wv[NORM == 2 ? VPT : 1][8];
#pragma unroll
  for (int j = 0; j < (NORM == 2 ? VPT : 1); ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) wv[j][i] = 1.f;
  if (w != nullptr) {
    if (NORM == 1) lds8(w + d0, wv[0]);  // plain global loads through the same helper (16-byte aligned)
    if (NORM == 2) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) lds8(w + (j * T + t) * 8, wv[NORM == 2 ? j : 0]);
    }
  }
  Vec8<bf16> cur[VPT], nxt[VPT];
  int l = blockIdx.x;
  if (l < L) {
#pragma unroll
    for (int j = 0; j < VPT; ++j) load8(x + b * x_sb + static_cast<int64_t>(l) * x_sl + (j * HPS + hsub) * x_sh + d0, cur[j]);
  }
  // the token's cos / sin slice travels with its row: fetched one iteration ahead, like the row itself (fetched at the point
  // of use it exposed a full L2 / HBM latency per token)
  float cs[8], sn[8], csn[8], snn[8];
  if (l < L && cosT != nullptr && l < L_rope) {
    lds8(cosT + static_cast<size_t>(l) * D + d0, cs);
    lds8(sinT + static_cast<size_t>(l) * D + d0, sn);
  }
  int parity = 0;
  for (; l < L; l += gridDim.x, parity ^= 1) {
    const int ln = l + gridDim.x;
    if (ln < L) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) load8(x + b * x_sb + static_cast<int64_t>(ln) * x_sl + (j * HPS + hsub) * x_sh + d0, nxt[j]);
      if (cosT != nullptr && ln < L_rope) {
        lds8(cosT + static_cast<size_t>(ln) * D + d0, csn);
        lds8(sinT + static_cast<size_t>(ln) * D + d0, snn);
      }
    }
    const bool rot = cosT != nullptr && l < L_rope;
    float f[VPT][8], ss[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      unpack(cur[j], f[j]);
      ss[j] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss[j] += f[j][i] * f[j][i];
    }
    float rstd[VPT];
    if (NORM == 1) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
#pragma unroll
        for (int o = LPH >> 1; o > 0; o >>= 1) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], o);
        rstd[j] = rsqrtf(ss[j] / D + eps);
        if (rstd_out != nullptr && (t % LPH) == 0)
          rstd_out[(static_cast<size_t>(b) * L + l) * H + j * HPS + hsub] = rstd[j];
      }
    } else if (NORM == 2) {
      float tot[1] = {0.f};
#pragma unroll
      for (int j = 0; j < VPT; ++j) tot[0] += ss[j];
      block_sum<T, 1>(tot, red, parity);
      const float r = rsqrtf(tot[0] / C + eps);
#pragma unroll
      for (int j = 0; j < VPT; ++j) rstd[j] = r;
      if (rstd_out != nullptr && t == 0) rstd_out[static_cast<size_t>(b) * L + l] = r;
    } else {
#pragma unroll
      for (int j = 0; j < VPT; ++j) rstd[j] = 1.f;
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      float n[8], o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) n[i] = f[j][i] * rstd[j] * wv[NORM == 2 ? j : 0][i];
      if (rot) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          o[i] = n[i] * cs[i] - n[i + 1] * sn[i];
          o[i + 1] = n[i + 1] * cs[i + 1] + n[i] * sn[i + 1];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = n[i];
      }
      stg16(y + b * y_sb + static_cast<int64_t>(l) * y_sl + (j * HPS + hsub) * y_sh + d0, pack(o));
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) cur[j] = nxt[j];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      cs[i] = csn[i];
      sn[i] = snn[i];
    }
  }
}

// Third generation of the same kernel: the token's row (contiguous H * D bf16 = 6 KB at K1) and its cos / sin slices arrive
// in a shared-memory ring by 1-D BULK ASYNCHRONOUS COPIES (cp.async.bulk + mbarrier complete_tx) issued by one thread
// STAGES - 1 tokens ahead. ncu on the register-prefetch kernel: 50 % of the stall samples are the first use of the
// prefetched row (one token of look-ahead, 5 CTAs per SM by its 96 registers: ~30 KB in flight per SM). Here nothing in
// flight holds a register, the look-ahead is two tokens with eight CTAs per SM, and the thread -> column mapping, the
// statistics and the rotation are unchanged. Needs heads contiguous inside a token (x_sh == D) and 16-byte aligned rows.
constexpr int ROPE_STAGES = 4;

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NORM, int VPT, int D>
__global__ void __launch_bounds__(128) rope_fwd_bulk_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                            float* __restrict__ rstd_out, const float* __restrict__ w,
                                                            const float* __restrict__ cosT, const float* __restrict__ sinT,
                                                            int64_t x_sb, int64_t x_sl, int64_t y_sb, int64_t y_sl,
                                                            int64_t y_sh, int L, int H, int L_rope, float eps) {
  constexpr int T = 128, LPH = D / 8, HPS = 1024 / D, S = ROPE_STAGES;
  extern __shared__ __align__(128) uint8_t rb_smem[];  // S x { row[C] bf16, cos[D] fp32, sin[D] fp32 }
  __shared__ __align__(8) uint64_t full[S];
  __shared__ float red[2 * (T / 32)];
  const int b = blockIdx.y, t = threadIdx.x;
  const int d0 = (t % LPH) * 8, hsub = t / LPH;
  const int C = H * D;
  const int stage_bytes = C * 2 + 2 * D * 4;
  float wv[NORM == 2 ? VPT : 1][8];
#pragma unroll
  for (int j = 0; j < (NORM == 2 ? VPT : 1); ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) wv[j][i] = 1.f;
  if (w != nullptr) {
    if (NORM == 1) lds8(w + d0, wv[0]);
    if (NORM == 2) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) lds8(w + (j * T + t) * 8, wv[NORM == 2 ? j : 0]);
    }
  }
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int n_tok = blockIdx.x < static_cast<unsigned>(L) ? (L - 1 - blockIdx.x) / gridDim.x + 1 : 0;  // tokens of this CTA
  auto issue = [&](int k) {  // thread 0 only: token k of this CTA into stage k % S
    const int l = blockIdx.x + k * gridDim.x;
    uint8_t* st = rb_smem + (k % S) * stage_bytes;
    const bool rot = cosT != nullptr && l < L_rope;
    mbar_arrive_expect_tx(&full[k % S], C * 2 + (rot ? 2 * D * 4 : 0));
    bulk_g2s(st, x + b * x_sb + static_cast<int64_t>(l) * x_sl, C * 2, &full[k % S]);
    if (rot) {
      bulk_g2s(st + C * 2, cosT + static_cast<size_t>(l) * D, D * 4, &full[k % S]);
      bulk_g2s(st + C * 2 + D * 4, sinT + static_cast<size_t>(l) * D, D * 4, &full[k % S]);
    }
  };
  if (t == 0)
    for (int k = 0; k < S - 1 && k < n_tok; ++k) issue(k);
  for (int k = 0; k < n_tok; ++k) {
    const int l = blockIdx.x + k * gridDim.x;
    // every thread has finished reading stage (k - 1) % S (its values are in registers / stored): refill it
    __syncthreads();
    if (t == 0 && k + S - 1 < n_tok) issue(k + S - 1);
    mbar_wait(&full[k % S], (k / S) & 1, 0x7e01);
    const uint8_t* st = rb_smem + (k % S) * stage_bytes;
    const bf16* row = reinterpret_cast<const bf16*>(st);
    const bool rot = cosT != nullptr && l < L_rope;
    float f[VPT][8], ss[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + (j * T + t) * 8);
      unpack(u, f[j]);
      ss[j] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss[j] = fmaf(f[j][i], f[j][i], ss[j]);
    }
    float cs[8], sn[8];
    if (rot) {
      lds8(reinterpret_cast<const float*>(st + C * 2) + d0, cs);
      lds8(reinterpret_cast<const float*>(st + C * 2 + D * 4) + d0, sn);
    }
    float rstd[VPT];
    if (NORM == 1) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
#pragma unroll
        for (int o = LPH >> 1; o > 0; o >>= 1) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], o);
        rstd[j] = rsqrtf(ss[j] / D + eps);
        if (rstd_out != nullptr && (t % LPH) == 0)
          rstd_out[(static_cast<size_t>(b) * L + l) * H + j * HPS + hsub] = rstd[j];
      }
    } else if (NORM == 2) {
      float tot[1] = {0.f};
#pragma unroll
      for (int j = 0; j < VPT; ++j) tot[0] += ss[j];
      block_sum<T, 1>(tot, red, k & 1);
      const float r = rsqrtf(tot[0] / C + eps);
#pragma unroll
      for (int j = 0; j < VPT; ++j) rstd[j] = r;
      if (rstd_out != nullptr && t == 0) rstd_out[static_cast<size_t>(b) * L + l] = r;
    } else {
#pragma unroll
      for (int j = 0; j < VPT; ++j) rstd[j] = 1.f;
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      float n[8], o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) n[i] = f[j][i] * rstd[j] * wv[NORM == 2 ? j : 0][i];
      if (rot) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          o[i] = n[i] * cs[i] - n[i + 1] * sn[i];
          o[i + 1] = n[i + 1] * cs[i + 1] + n[i] * sn[i + 1];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = n[i];
      }
      stg16(y + b * y_sb + static_cast<int64_t>(l) * y_sl + (j * HPS + hsub) * y_sh + d0, pack(o));
    }
  }
}

template <int NORM, int VPT, int D>
int launch_rope_bulk(const bf16* x, bf16* y, float* rstd, const float* w, const float* c, const float* s, const int64_t* xs,
                     const int64_t* ys, int B, int L, int H, int L_rope, float eps, cudaStream_t st) {
  const int smem = ROPE_STAGES * (H * D * 2 + 2 * D * 4);
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    VT_CHECK_CUDA(cudaFuncSetAttribute(rope_fwd_bulk_kernel<NORM, VPT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       ROPE_STAGES * (5120 * 2 + 2 * 128 * 4)));
  }
  // Grid: many more CTAs than are resident (96 per SM, ~8 tokens each at K1). The token -> CTA assignment is static; with
  // one wave of exactly-resident CTAs the kernel ran at 83 % of the copy bandwidth, with 8 / 24 / 96 per SM at 89 / 96 /
  // 100 %: short-lived CTAs let the hardware scheduler even out SM-to-SM and DRAM-channel differences.
  static const int want_f = getenv("VT_ROPE_CTAS_F") != nullptr ? atoi(getenv("VT_ROPE_CTAS_F")) : 96;
  dim3 grid(ln_grid_x(L, B, want_f), B);
  rope_fwd_bulk_kernel<NORM, VPT, D><<<grid, 128, smem, st>>>(x, y, rstd, w, c, s, xs[0], xs[1], ys[0], ys[1], ys[2], L, H,
                                                               L_rope, eps);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int NORM, int VPT, int D>
int launch_rope(const bf16* x, bf16* y, float* rstd, const float* w, const float* c, const float* s, const int64_t* xs,
                const int64_t* ys, int B, int L, int H, int L_rope, float eps, cudaStream_t st) {
  dim3 grid(ln_grid_x(L, B, 6), B);
  rope_fwd_kernel<NORM, VPT, D><<<grid, 128, 0, st>>>(x, y, rstd, w, c, s, xs[0], xs[1], xs[2], ys[0], ys[1], ys[2], L, H,
                                                       L_rope, eps);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Backward of the fused RMSNorm + RoPE in the same bulk-copy ring layout (the generic kernel in rowwise.cu is the first
// layout: one thread per 8 columns, C / 8 threads per CTA): per token the ring stage holds the dy row, the x row (norm
// modes), the cos / sin slices and, for the per-head norm, the token's H saved rstd values.
//   g = R^T dy (inverse rotation);  xh = x * rstd;  dx = rstd * (g w - xh * mean(g w xh));  dw += g xh
template <int NORM, int VPT, int D>
__global__ void __launch_bounds__(128) rope_bwd_bulk_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                            const float* __restrict__ rstd_in, bf16* __restrict__ dx,
                                                            float* __restrict__ dw, const float* __restrict__ w,
                                                            const float* __restrict__ cosT, const float* __restrict__ sinT,
                                                            int64_t g_sb, int64_t g_sl, int64_t x_sb, int64_t x_sl,
                                                            int64_t o_sb, int64_t o_sl, int64_t o_sh, int L, int H, int L_rope) {
  constexpr int T = 128, LPH = D / 8, HPS = 1024 / D, S = ROPE_STAGES;
  extern __shared__ __align__(128) uint8_t rb_smem[];  // S x { dy[C], x[C] bf16 (norm), cos[D], sin[D], rstd[H] fp32 (NORM 1) }
  __shared__ __align__(8) uint64_t full[S];
  __shared__ float red[2 * (T / 32)];
  __shared__ float sdw[NORM == 1 ? D : 1];
  const int b = blockIdx.y, t = threadIdx.x;
  const int d0 = (t % LPH) * 8, hsub = t / LPH;
  const int C = H * D;
  const int off_x = C * 2, off_cs = off_x + (NORM != 0 ? C * 2 : 0), off_rs = off_cs + 2 * D * 4;
  const int stage_bytes = off_rs + (NORM == 1 ? ((H * 4 + 15) & ~15) : 0);
  float wv[NORM == 2 ? VPT : 1][8], acc[NORM == 2 ? VPT : 1][8];
#pragma unroll
  for (int j = 0; j < (NORM == 2 ? VPT : 1); ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      wv[j][i] = 1.f;
      acc[j][i] = 0.f;
    }
  if (w != nullptr) {
    if (NORM == 1) lds8(w + d0, wv[0]);
    if (NORM == 2) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) lds8(w + (j * T + t) * 8, wv[NORM == 2 ? j : 0]);
    }
  }
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  if (NORM == 1)
    for (int i = t; i < D; i += T) sdw[i] = 0.f;
  __syncthreads();
  const int n_tok = blockIdx.x < static_cast<unsigned>(L) ? (L - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  auto issue = [&](int k) {  // thread 0 only
    const int l = blockIdx.x + k * gridDim.x;
    uint8_t* st = rb_smem + (k % S) * stage_bytes;
    const bool rot = cosT != nullptr && l < L_rope;
    mbar_arrive_expect_tx(&full[k % S], C * 2 + (NORM != 0 ? C * 2 : 0) + (rot ? 2 * D * 4 : 0) + (NORM == 1 ? H * 4 : 0));
    bulk_g2s(st, dy + b * g_sb + static_cast<int64_t>(l) * g_sl, C * 2, &full[k % S]);
    if (NORM != 0) bulk_g2s(st + off_x, x + b * x_sb + static_cast<int64_t>(l) * x_sl, C * 2, &full[k % S]);
    if (rot) {
      bulk_g2s(st + off_cs, cosT + static_cast<size_t>(l) * D, D * 4, &full[k % S]);
      bulk_g2s(st + off_cs + D * 4, sinT + static_cast<size_t>(l) * D, D * 4, &full[k % S]);
    }
    if (NORM == 1) bulk_g2s(st + off_rs, rstd_in + (static_cast<size_t>(b) * L + l) * H, H * 4, &full[k % S]);
  };
  if (t == 0)
    for (int k = 0; k < S - 1 && k < n_tok; ++k) issue(k);
  float rs_tok = 0.f;  // NORM 2: one rstd per token, fetched one token ahead
  if (NORM == 2 && n_tok > 0) rs_tok = rstd_in[static_cast<size_t>(b) * L + blockIdx.x];
  for (int k = 0; k < n_tok; ++k) {
    const int l = blockIdx.x + k * gridDim.x;
    __syncthreads();  // every thread is done with stage (k - 1) % S
    if (t == 0 && k + S - 1 < n_tok) issue(k + S - 1);
    float rs_next = 0.f;
    if (NORM == 2 && k + 1 < n_tok) rs_next = rstd_in[static_cast<size_t>(b) * L + l + gridDim.x];
    mbar_wait(&full[k % S], (k / S) & 1, 0x7e02);
    const uint8_t* st = rb_smem + (k % S) * stage_bytes;
    const bf16* grow = reinterpret_cast<const bf16*>(st);
    const bf16* xrow = reinterpret_cast<const bf16*>(st + off_x);
    const float* srs = reinterpret_cast<const float*>(st + off_rs);
    const bool rot = cosT != nullptr && l < L_rope;
    float cs[8], sn[8];
    if (rot) {
      lds8(reinterpret_cast<const float*>(st + off_cs) + d0, cs);
      lds8(reinterpret_cast<const float*>(st + off_cs + D * 4) + d0, sn);
    }
    float g[VPT][8], xh[NORM == 2 ? VPT : 1][8], dot2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      float fd[8];
      unpack(*reinterpret_cast<const uint4*>(grow + (j * T + t) * 8), fd);
      if (rot) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          g[j][i] = fd[i] * cs[i] + fd[i + 1] * sn[i + 1];
          g[j][i + 1] = fd[i + 1] * cs[i + 1] - fd[i] * sn[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[j][i] = fd[i];
      }
      bf16* dst = dx + b * o_sb + static_cast<int64_t>(l) * o_sl + (j * HPS + hsub) * o_sh + d0;
      if (NORM == 0) {
        stg16(dst, pack(g[j]));
      } else if (NORM == 1) {
        // the reduction is over one head = LPH adjacent lanes: finish this vector before the next
        float fx[8], xv[8], o[8], dot = 0.f;
        unpack(*reinterpret_cast<const uint4*>(xrow + (j * T + t) * 8), fx);
        const float rs = srs[j * HPS + hsub];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xv[i] = fx[i] * rs;
          dot = fmaf(g[j][i] * wv[0][i], xv[i], dot);
          acc[0][i] = fmaf(g[j][i], xv[i], acc[0][i]);
        }
#pragma unroll
        for (int o2 = LPH >> 1; o2 > 0; o2 >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o2);
        const float m = dot / D;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rs * (g[j][i] * wv[0][i] - xv[i] * m);
        stg16(dst, pack(o));
      } else {
        float fx[8];
        unpack(*reinterpret_cast<const uint4*>(xrow + (j * T + t) * 8), fx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[NORM == 2 ? j : 0][i] = fx[i] * rs_tok;
          dot2 = fmaf(g[j][i] * wv[NORM == 2 ? j : 0][i], xh[NORM == 2 ? j : 0][i], dot2);
          acc[NORM == 2 ? j : 0][i] = fmaf(g[j][i], xh[NORM == 2 ? j : 0][i], acc[NORM == 2 ? j : 0][i]);
        }
      }
    }
    if (NORM == 2) {
      float tot[1] = {dot2};
      block_sum<T, 1>(tot, red, k & 1);
      const float m = tot[0] / C;
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rs_tok * (g[j][i] * wv[NORM == 2 ? j : 0][i] - xh[NORM == 2 ? j : 0][i] * m);
        stg16(dx + b * o_sb + static_cast<int64_t>(l) * o_sl + (j * HPS + hsub) * o_sh + d0, pack(o));
      }
      rs_tok = rs_next;
    }
  }
  if (dw != nullptr && NORM == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&sdw[d0 + i], acc[0][i]);
    __syncthreads();
    for (int i = t; i < D; i += T) atomicAdd(dw + i, sdw[i]);
  }
  if (dw != nullptr && NORM == 2) {
#pragma unroll
    for (int j = 0; j < VPT; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(dw + (j * T + t) * 8 + i, acc[NORM == 2 ? j : 0][i]);
  }
}

template <int NORM, int VPT, int D>
int launch_rope_bwd_bulk(const bf16* dy, const bf16* x, const float* rstd, bf16* dx, float* dw, const float* w, const float* c,
                         const float* s, const int64_t* gs, const int64_t* xs, const int64_t* os, int B, int L, int H,
                         int L_rope, cudaStream_t st) {
  const int C = H * D;
  const int stage = C * 2 + (NORM != 0 ? C * 2 : 0) + 2 * D * 4 + (NORM == 1 ? ((H * 4 + 15) & ~15) : 0);
  const int smem = ROPE_STAGES * stage;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    VT_CHECK_CUDA(cudaFuncSetAttribute(rope_bwd_bulk_kernel<NORM, VPT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       ROPE_STAGES * (5120 * 4 + 2 * 128 * 4 + 512)));
  }
  // 48 CTAs per SM of grid (see launch_rope_bulk); the full-row norm flushes C weight-gradient atomics per CTA: fewer there
  static const int want_b = getenv("VT_ROPE_CTAS_B") != nullptr ? atoi(getenv("VT_ROPE_CTAS_B")) : (NORM == 2 ? 8 : 48);
  dim3 grid(ln_grid_x(L, B, want_b), B);
  rope_bwd_bulk_kernel<NORM, VPT, D><<<grid, 128, smem, st>>>(dy, x, rstd, dx, dw, w, c, s, gs[0], gs[1], xs[0], xs[1], os[0],
                                                               os[1], os[2], L, H, L_rope);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// 1 = no specialised configuration (caller falls back to rowwise.cu)
int rope_fwd_fast(const void* x, void* y, float* rstd, const float* w, const float* c, const float* s, const int64_t* xs,
                  const int64_t* ys, int B, int L, int H, int D, int L_rope, int norm_mode, float eps, cudaStream_t st) {
  const int C = H * D;
  if (C % 1024 != 0 || B > 65535 || (D != 64 && D != 128)) return 1;
  const int vpt = C / 1024;
  const bf16* xp = static_cast<const bf16*>(x);
  bf16* yp = static_cast<bf16*>(y);
  // rows that are contiguous and 16-byte aligned (heads packed inside a token, as every caller's views are) take the
  // bulk-copy kernel; VT_ROPE_REG=1 keeps the register-prefetch kernel (A/B measurements)
  static const bool reg_only = getenv("VT_ROPE_REG") != nullptr && atoi(getenv("VT_ROPE_REG")) != 0;
  const bool bulk_ok = !reg_only && xs[2] == D && (xs[1] % 8) == 0 && (xs[0] % 8) == 0 && aligned16(x) &&
                       (c == nullptr || (aligned16(c) && aligned16(s)));
#define VT_ROPE(N_, V_)                                                                                                  \
  if (norm_mode == N_ && vpt == V_) {                                                                                   \
    if (bulk_ok)                                                                                                        \
      return D == 128 ? launch_rope_bulk<N_, V_, 128>(xp, yp, rstd, w, c, s, xs, ys, B, L, H, L_rope, eps, st)          \
                      : launch_rope_bulk<N_, V_, 64>(xp, yp, rstd, w, c, s, xs, ys, B, L, H, L_rope, eps, st);          \
    return D == 128 ? launch_rope<N_, V_, 128>(xp, yp, rstd, w, c, s, xs, ys, B, L, H, L_rope, eps, st)                 \
                    : launch_rope<N_, V_, 64>(xp, yp, rstd, w, c, s, xs, ys, B, L, H, L_rope, eps, st);                 \
  }
  VT_ROPE(0, 3) VT_ROPE(1, 3) VT_ROPE(2, 3) VT_ROPE(0, 5) VT_ROPE(1, 5) VT_ROPE(2, 5) VT_ROPE(0, 1) VT_ROPE(1, 1) VT_ROPE(2, 1)
  VT_ROPE(0, 2) VT_ROPE(1, 2) VT_ROPE(2, 2)
#undef VT_ROPE
  return 1;
}

// 1 = no specialised configuration (caller falls back to the generic kernel in rowwise.cu)
int rope_bwd_fast(const void* dy, const void* x, const float* rstd, void* dx, float* dw, const float* w, const float* c,
                  const float* s, const int64_t* gs, const int64_t* xs, const int64_t* os, int B, int L, int H, int D,
                  int L_rope, int norm_mode, cudaStream_t st) {
  const int C = H * D;
  if (C % 1024 != 0 || B > 65535 || (D != 64 && D != 128)) return 1;
  static const bool generic = getenv("VT_ROPE_REG") != nullptr && atoi(getenv("VT_ROPE_REG")) != 0;
  if (generic) return 1;
  // contiguous, 16-byte aligned token rows for dy and x; per-head rstd rows must be 16-byte multiples
  if (gs[2] != D || (gs[1] % 8) != 0 || (gs[0] % 8) != 0 || !aligned16(dy)) return 1;
  if (norm_mode != 0 && (xs[2] != D || (xs[1] % 8) != 0 || (xs[0] % 8) != 0 || !aligned16(x))) return 1;
  if (norm_mode == 1 && ((H % 4) != 0 || !aligned16(rstd))) return 1;
  if (c != nullptr && (!aligned16(c) || !aligned16(s))) return 1;
  const int vpt = C / 1024;
  const bf16* gp = static_cast<const bf16*>(dy);
  const bf16* xp = static_cast<const bf16*>(x);
  bf16* op = static_cast<bf16*>(dx);
#define VT_ROPE_B(N_, V_)                                                                                                \
  if (norm_mode == N_ && vpt == V_)                                                                                     \
    return D == 128 ? launch_rope_bwd_bulk<N_, V_, 128>(gp, xp, rstd, op, dw, w, c, s, gs, xs, os, B, L, H, L_rope, st) \
                    : launch_rope_bwd_bulk<N_, V_, 64>(gp, xp, rstd, op, dw, w, c, s, gs, xs, os, B, L, H, L_rope, st);
  VT_ROPE_B(0, 3) VT_ROPE_B(1, 3) VT_ROPE_B(2, 3) VT_ROPE_B(0, 5) VT_ROPE_B(1, 5) VT_ROPE_B(2, 5) VT_ROPE_B(0, 1) VT_ROPE_B(1, 1)
  VT_ROPE_B(2, 1) VT_ROPE_B(0, 2) VT_ROPE_B(1, 2) VT_ROPE_B(2, 2)
#undef VT_ROPE_B
  return 1;
}

}  // namespace vt
