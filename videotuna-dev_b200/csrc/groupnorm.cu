// groupnorm.cu — GroupNorm(G) [+ SiLU] over (N, C, S) tensors with fp32 statistics, forward and backward.
// Replaces lvdm GroupNormSpecific + nn.SiLU (videotuna/models/lvdm/modules/utils.py:192-203,
// networks/openaimodel3d.py:229-255) and the `norm` of Spatial/TemporalTransformer (attention.py:376-392,475-519).
//
// One CTA per (sample, group): the group's (C/G)*S elements are contiguous in memory. Pass 1 reads the slab and
// reduces sum / sum of squares (Welford-free: shifted by the first element for stability); pass 2 re-reads it (an
// L2 hit: slabs are 50-400 KB, far below the 126 MB L2), normalises, applies SiLU and writes. DRAM traffic is therefore
// 4 B/elem (bf16) although the kernel is two-pass.
//
// Second generation (groupnorm_{fwd,bwd}_bulk_kernel, the default whenever the slab is 16-byte sliceable): the slab of
// one (sample, group) is cut into CL equal chunks, one per CTA of a thread-block CLUSTER of CL CTAs (CL = 1..16). One
// thread fetches the CTA's chunk into shared memory with four 1-D bulk asynchronous copies (cp.async.bulk + mbarrier
// complete_tx), so the whole chunk is in flight at once with no registers tied up; the statistics are reduced over the
// chunk as the sub-copies land, combined across the cluster through distributed shared memory, and the normalise +
// SiLU pass reads the chunk back from shared memory: HBM sees every element exactly once in each direction. The cluster
// dimension is what makes the TemporalTransformer's 5-D GroupNorm parallel: there the slab is (C/G) * t*h*w elements
// (819 KB at VideoCrafter2 level 0) and only N*G = 64 slabs exist, i.e. 64 CTAs for 148 SMs in the first generation.
#include <cstdlib>
#include <mutex>
#include <set>
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "capi_util.h"
#include "sm100_ptx.cuh"

namespace cg = cooperative_groups;

namespace vt {
namespace {

constexpr int GN_THREADS = 512;

template <typename T>
struct Io;
template <>
struct Io<float> {
  static constexpr int VEC = 4;
  __device__ static void load(const float* p, float* f) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static void store(float* p, const float* f) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
  __device__ static float ld1(const float* p) { return *p; }
  __device__ static void st1(float* p, float v) { *p = v; }
};
template <>
struct Io<__nv_bfloat16> {
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
  __device__ static float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
};

template <int NV, int THREADS = GN_THREADS>
__device__ __forceinline__ void cta_sum(float* v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = THREADS / 32;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * NW + warp] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float t = lane < NW ? red[i * NW + lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    v[i] = t;
  }
}

__device__ __forceinline__ float silu_f(float z) { return z / (1.f + __expf(-z)); }
// One MUFU op per element instead of two (ex2 + rcp): sigmoid(z) = 0.5 + 0.5 tanh(z / 2), tanh.approx.f32 (max relative
// error 2^-11, below the 2^-9 rounding of a bf16 result — used by the bf16 instantiations only). At one element per lane
// and clock the two-MUFU form costs 11.6 us on the VideoCrafter2 level-0 tensor, as much as its HBM time.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T> __device__ __forceinline__ float silu_t(float z) { return silu_f(z); }
template <> __device__ __forceinline__ float silu_t<__nv_bfloat16>(float z) {
  const float h = 0.5f * z;
  return fmaf(h, tanh_approx(h), h);
}
__device__ __forceinline__ float dsilu_f(float z) {
  const float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
}
template <typename T> __device__ __forceinline__ float dsilu_t(float z) { return dsilu_f(z); }
template <> __device__ __forceinline__ float dsilu_t<__nv_bfloat16>(float z) {
  const float s = fmaf(0.5f, tanh_approx(0.5f * z), 0.5f);
  return s * fmaf(z, 1.f - s, 1.f);
}

// `vec_ok`: S is a multiple of the vector width and the base is 16-byte aligned -> vector path; else scalar path.
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, const float* __restrict__ gamma, const float* __restrict__ beta,
                     int C, int S, int G, float eps, int apply_silu, int vec_ok) {
  __shared__ float red[2 * (GN_THREADS / 32)];
  constexpr int VEC = Io<T>::VEC;
  const int n = blockIdx.y, g = blockIdx.x;
  const int cpg = C / G;
  const size_t base = (static_cast<size_t>(n) * C + static_cast<size_t>(g) * cpg) * S;
  const int count = cpg * S;
  const T* xs = x + base;
  T* ys = y + base;
  const float shiftv = Io<T>::ld1(xs);  // shifted sums: avoids cancellation when |mean| >> std

  float acc[2] = {0.f, 0.f};
  if (vec_ok) {
    for (int i = threadIdx.x * VEC; i < count; i += GN_THREADS * VEC) {
      float f[VEC];
      Io<T>::load(xs + i, f);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float d = f[k] - shiftv;
        acc[0] += d;
        acc[1] += d * d;
      }
    }
  } else {
    for (int i = threadIdx.x; i < count; i += GN_THREADS) {
      const float d = Io<T>::ld1(xs + i) - shiftv;
      acc[0] += d;
      acc[1] += d * d;
    }
  }
  cta_sum<2>(acc, red);
  const float inv = 1.f / count;
  const float md = acc[0] * inv;
  const float var = fmaxf(acc[1] * inv - md * md, 0.f);
  const float mean = md + shiftv;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[n * G + g] = mean;
    if (rstd_out) rstd_out[n * G + g] = rstd;
  }
  if (vec_ok) {
    for (int i = threadIdx.x * VEC; i < count; i += GN_THREADS * VEC) {
      const int c = g * cpg + i / S;  // a vector never straddles channels because S % VEC == 0
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float a = rstd * ga, bsh = be - mean * a;
      float f[VEC];
      Io<T>::load(xs + i, f);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float z = fmaf(f[k], a, bsh);
        f[k] = apply_silu ? silu_f(z) : z;
      }
      Io<T>::store(ys + i, f);
    }
  } else {
    for (int i = threadIdx.x; i < count; i += GN_THREADS) {
      const int c = g * cpg + i / S;
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float z = (Io<T>::ld1(xs + i) - mean) * rstd * ga + be;
      Io<T>::st1(ys + i, apply_silu ? silu_f(z) : z);
    }
  }
}

#ifdef VT_EXPERIMENTS
// EXPERIMENT (off unless VT_GN_REG is set; measured SLOWER: 116 vs 79 us on the VC2 level-0 tensor, because ~100 registers x
// 512 threads leave one CTA per SM while the two-pass kernel keeps four). Register-resident single pass for slabs of at most THREADS * NVEC vectors (every VideoCrafter2 level at its own
// channel count: 6 400 .. 25 600 elements): the slab is read from HBM once, kept in registers across the statistics, and
// written once — no L2 re-read and half the dependent latency of the two-pass kernel, which stays for larger slabs.
template <typename T, int THREADS, int NVEC>
__global__ void __launch_bounds__(THREADS)
groupnorm_fwd_reg_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ mean_out,
                         float* __restrict__ rstd_out, const float* __restrict__ gamma, const float* __restrict__ beta,
                         int C, int S, int G, float eps, int apply_silu) {
  __shared__ float red[2 * (THREADS / 32)];
  constexpr int VEC = Io<T>::VEC;
  const int n = blockIdx.y, g = blockIdx.x;
  const int cpg = C / G;
  const size_t base = (static_cast<size_t>(n) * C + static_cast<size_t>(g) * cpg) * S;
  const int count = cpg * S;
  const T* xs = x + base;
  T* ys = y + base;
  const float shiftv = Io<T>::ld1(xs);
  float f[NVEC][VEC];
  float acc[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NVEC; ++j) {
    const int i = (j * THREADS + threadIdx.x) * VEC;
    if (i < count) {
      Io<T>::load(xs + i, f[j]);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float d = f[j][k] - shiftv;
        acc[0] += d;
        acc[1] += d * d;
      }
    }
  }
  cta_sum<2, THREADS>(acc, red);
  const float inv = 1.f / count;
  const float md = acc[0] * inv;
  const float var = fmaxf(acc[1] * inv - md * md, 0.f);
  const float mean = md + shiftv;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[n * G + g] = mean;
    if (rstd_out) rstd_out[n * G + g] = rstd;
  }
#pragma unroll
  for (int j = 0; j < NVEC; ++j) {
    const int i = (j * THREADS + threadIdx.x) * VEC;
    if (i < count) {
      const int c = g * cpg + i / S;  // a vector never straddles channels because S % VEC == 0
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float a = rstd * ga, bsh = be - mean * a;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float z = fmaf(f[j][k], a, bsh);
        f[j][k] = apply_silu ? silu_f(z) : z;
      }
      Io<T>::store(ys + i, f[j]);
    }
  }
}

template <typename T>
bool launch_gn_fwd_reg(const T* x, T* y, float* mean, float* rstd, const float* gamma, const float* beta, int N, int C, int S,
                       int G, float eps, int apply_silu, cudaStream_t st) {
  constexpr int VEC = Io<T>::VEC;
  const long long vectors = static_cast<long long>(C / G) * S / VEC;
  dim3 grid(G, N);
  if (vectors <= 128 * 8)
    groupnorm_fwd_reg_kernel<T, 128, 8><<<grid, 128, 0, st>>>(x, y, mean, rstd, gamma, beta, C, S, G, eps, apply_silu);
  else if (vectors <= 256 * 8)
    groupnorm_fwd_reg_kernel<T, 256, 8><<<grid, 256, 0, st>>>(x, y, mean, rstd, gamma, beta, C, S, G, eps, apply_silu);
  else if (vectors <= 512 * 8)
    groupnorm_fwd_reg_kernel<T, 512, 8><<<grid, 512, 0, st>>>(x, y, mean, rstd, gamma, beta, C, S, G, eps, apply_silu);
  else
    return false;
  return true;
}
#endif  // VT_EXPERIMENTS

// z = xh*gamma + beta; y = silu(z) or z; gz = dy * silu'(z); gh = gz * gamma
// dx = rstd * (gh - mean_g(gh) - xh * mean_g(gh * xh)); dgamma[c] += sum gz*xh; dbeta[c] += sum gz
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, T* __restrict__ dx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ dgamma, float* __restrict__ dbeta, int C,
                     int S, int G, int apply_silu, int vec_ok) {
  __shared__ float red[2 * (GN_THREADS / 32)];
  const int n = blockIdx.y, g = blockIdx.x;
  const int cpg = C / G;
  const size_t base = (static_cast<size_t>(n) * C + static_cast<size_t>(g) * cpg) * S;
  const float mean = mean_in[n * G + g], rstd = rstd_in[n * G + g];
  float tot[2] = {0.f, 0.f};
  // pass 1: per-channel sums (also the group sums)
  for (int cc = 0; cc < cpg; ++cc) {
    const int c = g * cpg + cc;
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const T* xs = x + base + static_cast<size_t>(cc) * S;
    const T* ds = dy + base + static_cast<size_t>(cc) * S;
    float a[2] = {0.f, 0.f};
    if (vec_ok) {
      constexpr int VEC = Io<T>::VEC;
      for (int i = threadIdx.x * VEC; i < S; i += GN_THREADS * VEC) {
        float fx[VEC], fg[VEC];
        Io<T>::load(xs + i, fx);
        Io<T>::load(ds + i, fg);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float xh = (fx[k] - mean) * rstd;
          float gz = fg[k];
          if (apply_silu) gz *= dsilu_f(xh * ga + be);
          a[0] += gz;
          a[1] += gz * xh;
        }
      }
    } else {
      for (int i = threadIdx.x; i < S; i += GN_THREADS) {
        const float xh = (Io<T>::ld1(xs + i) - mean) * rstd;
        float gz = Io<T>::ld1(ds + i);
        if (apply_silu) gz *= dsilu_f(xh * ga + be);
        a[0] += gz;
        a[1] += gz * xh;
      }
    }
    cta_sum<2>(a, red);
    if (threadIdx.x == 0) {
      if (dbeta) atomicAdd(dbeta + c, a[0]);
      if (dgamma) atomicAdd(dgamma + c, a[1]);
    }
    tot[0] += a[0] * ga;
    tot[1] += a[1] * ga;
  }
  const float inv = 1.f / (cpg * S);
  const float m1 = tot[0] * inv, m2 = tot[1] * inv;
  // pass 2
  for (int cc = 0; cc < cpg; ++cc) {
    const int c = g * cpg + cc;
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const T* xs = x + base + static_cast<size_t>(cc) * S;
    const T* ds = dy + base + static_cast<size_t>(cc) * S;
    T* os = dx + base + static_cast<size_t>(cc) * S;
    if (vec_ok) {
      constexpr int VEC = Io<T>::VEC;
      for (int i = threadIdx.x * VEC; i < S; i += GN_THREADS * VEC) {
        float fx[VEC], fg[VEC];
        Io<T>::load(xs + i, fx);
        Io<T>::load(ds + i, fg);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float xh = (fx[k] - mean) * rstd;
          float gz = fg[k];
          if (apply_silu) gz *= dsilu_f(xh * ga + be);
          fx[k] = rstd * (gz * ga - m1 - xh * m2);
        }
        Io<T>::store(os + i, fx);
      }
    } else {
      for (int i = threadIdx.x; i < S; i += GN_THREADS) {
        const float xh = (Io<T>::ld1(xs + i) - mean) * rstd;
        float gz = Io<T>::ld1(ds + i);
        if (apply_silu) gz *= dsilu_f(xh * ga + be);
        Io<T>::st1(os + i, rstd * (gz * ga - m1 - xh * m2));
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// bulk-copy + cluster kernels
// ------------------------------------------------------------------------------------------------------------------
constexpr int GB_THREADS = 256;
constexpr int GB_NSUB = 4;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Channel index of the vector a thread visits next, advanced without a division: a thread's vectors are `stride` elements
// apart, so (channel, offset inside the channel) move by (stride / S, stride % S) with one carry.
struct ChanIter {
  int cc, rem, dq, dr, S;
  __device__ __forceinline__ ChanIter(int pos, int stride, int S_) : S(S_) {
    cc = pos / S_;
    rem = pos - cc * S_;
    dq = stride / S_;
    dr = stride - dq * S_;
  }
  __device__ __forceinline__ void next() {
    cc += dq;
    rem += dr;
    if (rem >= S) {
      rem -= S;
      ++cc;
    }
  }
};

// Sum `NV` per-CTA values over the CTAs of the cluster: every CTA publishes its partials in its own shared memory, the
// cluster synchronises, and every CTA reads all peers' partials through DSMEM. The caller must keep its shared memory
// alive until the peers have read it (cluster.sync() before exit).
template <int NV>
__device__ __forceinline__ void cluster_sum(float* v, float* partial, int CL) {
  if (CL == 1) return;
  cg::cluster_group cluster = cg::this_cluster();
  if (threadIdx.x == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) partial[i] = v[i];
  cluster.sync();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = 0.f;
  for (int r = 0; r < CL; ++r) {
    const float* peer = cluster.map_shared_rank(partial, r);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += peer[i];
  }
}

// grid (CL, G, N), cluster (CL, 1, 1); `chunk` = elements per CTA (a multiple of the vector width), dynamic shared memory
// = chunk * sizeof(T).
template <typename T>
__global__ void __launch_bounds__(GB_THREADS)
groupnorm_fwd_bulk_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ mean_out,
                          float* __restrict__ rstd_out, const float* __restrict__ gamma, const float* __restrict__ beta,
                          int C, int S, int G, float eps, int apply_silu, int chunk, int CL) {
  extern __shared__ __align__(128) unsigned char gb_smem[];
  T* buf = reinterpret_cast<T*>(gb_smem);
  float* sgb = reinterpret_cast<float*>(buf + chunk);  // [2][cpg]: gamma, beta of this group's channels
  __shared__ __align__(8) uint64_t bars[GB_NSUB];
  __shared__ float red[2 * (GB_THREADS / 32)];
  __shared__ float partial[2];
  constexpr int VEC = Io<T>::VEC;
  const int rank = blockIdx.x, g = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
  const int cpg = C / G;
  const size_t base = (static_cast<size_t>(n) * C + static_cast<size_t>(g) * cpg) * S;
  const int count = cpg * S;
  const int start = rank * chunk;
  const int len = max(0, min(chunk, count - start));
  // sub-copies of whole CTA sweeps (GB_THREADS vectors), so a thread's m-th vector is the same in every pass
  const int sub = ((chunk / VEC + GB_NSUB * GB_THREADS - 1) / (GB_NSUB * GB_THREADS)) * (GB_THREADS * VEC);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GB_NSUB; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < cpg; i += GB_THREADS) {
    sgb[i] = gamma ? gamma[g * cpg + i] : 1.f;
    sgb[cpg + i] = beta ? beta[g * cpg + i] : 0.f;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GB_NSUB; ++s) {
      const int off = s * sub, l = min(sub, len - off);
      if (l > 0) {
        mbar_arrive_expect_tx(&bars[s], l * sizeof(T));
        bulk_load_1d(buf + off, x + base + start + off, l * sizeof(T), &bars[s]);
      }
    }
  }
  const float shiftv = Io<T>::ld1(x + base);  // the same shift in every CTA of the cluster
  float acc[2] = {0.f, 0.f};
#pragma unroll 1
  for (int s = 0; s < GB_NSUB; ++s) {
    const int off = s * sub, l = min(sub, len - off);
    if (l <= 0) break;
    mbar_wait(&bars[s], 0, 0x6e01);
    for (int i = off + tid * VEC; i < off + l; i += GB_THREADS * VEC) {
      float f[VEC];
      Io<T>::load(buf + i, f);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float d = f[k] - shiftv;
        acc[0] += d;
        acc[1] = fmaf(d, d, acc[1]);
      }
    }
  }
  cta_sum<2, GB_THREADS>(acc, red);
  cluster_sum<2>(acc, partial, CL);
  const float inv = 1.f / count;
  const float md = acc[0] * inv;
  const float var = fmaxf(acc[1] * inv - md * md, 0.f);
  const float mean = md + shiftv;
  const float rstd = rsqrtf(var + eps);
  if (tid == 0 && rank == 0) {
    if (mean_out) mean_out[n * G + g] = mean;
    if (rstd_out) rstd_out[n * G + g] = rstd;
  }
  T* ys = y + base + start;
  ChanIter ch(start + tid * VEC, GB_THREADS * VEC, S);  // a vector never straddles channels because S % VEC == 0
  if (apply_silu && sizeof(T) == 2) {
    // silu(z) = h + h tanh(h), h = z / 2 = x * (rstd gamma / 2) + (beta - mean rstd gamma) / 2: 2 FMA + 1 MUFU per element
    for (int i = tid * VEC; i < len; i += GB_THREADS * VEC, ch.next()) {
      const float a = 0.5f * rstd * sgb[ch.cc], bsh = fmaf(-mean, a, 0.5f * sgb[cpg + ch.cc]);
      float f[VEC];
      Io<T>::load(buf + i, f);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float h = fmaf(f[k], a, bsh);
        f[k] = fmaf(h, tanh_approx(h), h);
      }
      Io<T>::store(ys + i, f);
    }
  } else {
    for (int i = tid * VEC; i < len; i += GB_THREADS * VEC, ch.next()) {
      const float a = rstd * sgb[ch.cc], bsh = fmaf(-mean, a, sgb[cpg + ch.cc]);
      float f[VEC];
      Io<T>::load(buf + i, f);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float z = fmaf(f[k], a, bsh);
        f[k] = apply_silu ? silu_t<T>(z) : z;
      }
      Io<T>::store(ys + i, f);
    }
  }
  if (CL > 1) cg::this_cluster().sync();  // peers may still be reading `partial`
}

// Backward in the same layout: x and dy chunks both staged in shared memory (2 * chunk * sizeof(T)); per-channel
// dgamma / dbeta partials go through shared-memory bins (one warp-reduced atomic per warp and vector sweep) and leave
// the CTA as one global atomic per channel; the two group sums are combined across the cluster like the statistics.
template <typename T>
__global__ void __launch_bounds__(GB_THREADS)
groupnorm_bwd_bulk_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, T* __restrict__ dx, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ dgamma, float* __restrict__ dbeta, int C,
                          int S, int G, int apply_silu, int chunk, int CL) {
  extern __shared__ __align__(128) unsigned char gb_smem[];
  const int cpg = C / G;
  T* bx = reinterpret_cast<T*>(gb_smem);
  T* bg = bx + chunk;
  float* sgb = reinterpret_cast<float*>(bg + chunk);   // [2][cpg]: gamma, beta of this group's channels
  float* bins = sgb + 2 * cpg;  // [warps][2][cpg]: sum gz, sum gz * xh per channel, one private copy per warp (no atomics)
  __shared__ __align__(8) uint64_t bars[GB_NSUB];
  __shared__ float red[2 * (GB_THREADS / 32)];
  __shared__ float partial[2];
  constexpr int VEC = Io<T>::VEC;
  const int rank = blockIdx.x, g = blockIdx.y, n = blockIdx.z, tid = threadIdx.x, lane = threadIdx.x & 31;
  constexpr int NW = GB_THREADS / 32;
  const size_t base = (static_cast<size_t>(n) * C + static_cast<size_t>(g) * cpg) * S;
  const int count = cpg * S;
  const int start = rank * chunk;
  const int len = max(0, min(chunk, count - start));
  // sub-copies of whole CTA sweeps (GB_THREADS vectors), so a thread's m-th vector is the same in every pass
  const int sub = ((chunk / VEC + GB_NSUB * GB_THREADS - 1) / (GB_NSUB * GB_THREADS)) * (GB_THREADS * VEC);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GB_NSUB; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < NW * 2 * cpg; i += GB_THREADS) bins[i] = 0.f;
  for (int i = tid; i < cpg; i += GB_THREADS) {
    sgb[i] = gamma ? gamma[g * cpg + i] : 1.f;
    sgb[cpg + i] = beta ? beta[g * cpg + i] : 0.f;
  }
  float* wbins = bins + (tid >> 5) * 2 * cpg;
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GB_NSUB; ++s) {
      const int off = s * sub, l = min(sub, len - off);
      if (l > 0) {
        mbar_arrive_expect_tx(&bars[s], 2 * l * sizeof(T));
        bulk_load_1d(bx + off, x + base + start + off, l * sizeof(T), &bars[s]);
        bulk_load_1d(bg + off, dy + base + start + off, l * sizeof(T), &bars[s]);
      }
    }
  }
  const float mean = mean_in[n * G + g], rstd = rstd_in[n * G + g];
  const float nmr = -mean * rstd;
  float tot[2] = {0.f, 0.f};
  ChanIter ch(start + tid * VEC, GB_THREADS * VEC, S);
  if constexpr (sizeof(T) == 2) {
    if (apply_silu) {
      // ---- bf16 + SiLU fast path: packed f32x2 arithmetic and DEFERRED bin updates ----------------------------------
      // The first version spent 54 executed instructions per element (issue slots 70 % busy, profiles/r1_s22): scalar FMAs
      // and a ten-shuffle warp reduction per 8-element vector. Here every arithmetic step handles two elements
      // (fma.rn.f32x2), and while a warp's 32 vectors stay inside one channel — the common case: a channel is S elements
      // long — the per-thread sums simply keep accumulating; the warp reduction and the bin update happen once per channel
      // change instead of once per vector.
      int cur_c = -1;
      float2 A0 = make_float2(0.f, 0.f), A1 = make_float2(0.f, 0.f);
      auto flush = [&]() {
        if (cur_c >= 0) {
          float v0 = A0.x + A0.y, v1 = A1.x + A1.y;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
          }
          if (lane == 0) {
            const float ga = sgb[cur_c];
            wbins[cur_c] += v0;
            wbins[cpg + cur_c] += v1;
            tot[0] = fmaf(v0, ga, tot[0]);
            tot[1] = fmaf(v1, ga, tot[1]);
          }
          __syncwarp();
        }
        A0 = make_float2(0.f, 0.f);
        A1 = make_float2(0.f, 0.f);
      };
      const float2 one2 = make_float2(1.f, 1.f), mone2 = make_float2(-1.f, -1.f), half2 = make_float2(0.5f, 0.5f);
      const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(nmr, nmr);
#pragma unroll 1
      for (int s = 0; s < GB_NSUB; ++s) {
        const int off = s * sub, l = min(sub, len - off);
        if (l <= 0) break;
        mbar_wait(&bars[s], 0, 0x6e02);
        for (int i0 = off + (tid - lane) * VEC; i0 < off + l; i0 += GB_THREADS * VEC, ch.next()) {
          const int i = i0 + lane * VEC;
          const bool ok = i < off + l;
          const int cc = ok ? ch.cc : -1;
          const int c0 = __shfl_sync(0xffffffffu, cc, 0);
          const bool uniform = __all_sync(0xffffffffu, cc == c0 || !ok);
          if (uniform && c0 != cur_c) {
            flush();
            cur_c = c0;
          }
          float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
          if (ok) {
            const float ga = sgb[cc];
            // silu'(z) = s (1 + z (1 - s)) with s = (1 + t) / 2, t = tanh(h), h = z / 2:  = (1/2 + t/2) (1 + h (1 - t))
            const float ha = 0.5f * rstd * ga, hb = fmaf(mean, -ha, 0.5f * sgb[cpg + cc]);
            const float2 ha2 = make_float2(ha, ha), hb2 = make_float2(hb, hb);
            const uint4 ux = *reinterpret_cast<const uint4*>(bx + i);
            uint4 ug = *reinterpret_cast<const uint4*>(bg + i);
            const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&ux);
            __nv_bfloat162* hg = reinterpret_cast<__nv_bfloat162*>(&ug);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 x2 = __bfloat1622float2(hx[k]);
              const float2 g2 = __bfloat1622float2(hg[k]);
              const float2 h = __ffma2_rn(x2, ha2, hb2);
              const float2 t = make_float2(tanh_approx(h.x), tanh_approx(h.y));
              const float2 u = __ffma2_rn(h, __ffma2_rn(t, mone2, one2), one2);   // 1 + h (1 - t)
              const float2 sg = __ffma2_rn(t, half2, half2);                       // sigmoid(z)
              const float2 gz = __fmul2_rn(g2, __fmul2_rn(sg, u));
              const float2 xh = __ffma2_rn(x2, rstd2, nmr2);
              a0 = __fadd2_rn(a0, gz);
              a1 = __ffma2_rn(gz, xh, a1);
              hg[k] = __floats2bfloat162_rn(gz.x, gz.y);
            }
            *reinterpret_cast<uint4*>(bg + i) = ug;  // the second pass reads gz = dy * silu'(z) instead of recomputing it
          }
          if (uniform) {
            A0 = __fadd2_rn(A0, a0);
            A1 = __fadd2_rn(A1, a1);
          } else {
            // a channel boundary runs through this warp's 256 elements: settle what is pending, then one reduction round
            // per distinct channel
            flush();
            cur_c = -1;
            float s0 = a0.x + a0.y, s1 = a1.x + a1.y;
            if (ok) {
              const float ga = sgb[cc];
              tot[0] = fmaf(s0, ga, tot[0]);
              tot[1] = fmaf(s1, ga, tot[1]);
            }
            unsigned remaining = __ballot_sync(0xffffffffu, ok);
            while (remaining != 0) {
              const int leader = __ffs(remaining) - 1;
              const int c = __shfl_sync(0xffffffffu, cc, leader);
              const bool mine = ok && cc == c;
              float v0 = mine ? s0 : 0.f, v1 = mine ? s1 : 0.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
              }
              if (lane == leader) {
                wbins[c] += v0;
                wbins[cpg + c] += v1;
              }
              __syncwarp();
              remaining &= ~__ballot_sync(0xffffffffu, mine);
            }
          }
        }
      }
      flush();
    }
  }
  if (!(sizeof(T) == 2 && apply_silu)) {
#pragma unroll 1
  for (int s = 0; s < GB_NSUB; ++s) {
    const int off = s * sub, l = min(sub, len - off);
    if (l <= 0) break;
    mbar_wait(&bars[s], 0, 0x6e02);
    // `sub` is a whole number of CTA sweeps: the trip count is warp-uniform and the shuffles below are convergent
    for (int i0 = off + (tid - lane) * VEC; i0 < off + l; i0 += GB_THREADS * VEC, ch.next()) {
      const int i = i0 + lane * VEC;
      const bool ok = i < off + l;
      const int cc = ok ? ch.cc : -1;
      float a0 = 0.f, a1 = 0.f;
      if (ok) {
        const float ga = sgb[cc];
        float fx[VEC], fg[VEC];
        Io<T>::load(bx + i, fx);
        Io<T>::load(bg + i, fg);
        if (apply_silu && sizeof(T) == 2) {
          // silu'(z) = s (1 + z (1 - s)) with s = (1 + t) / 2, t = tanh(h), h = z / 2:  = (1/2 + t/2) (1 + h - h t)
          const float ha = 0.5f * rstd * ga, hb = fmaf(mean, -ha, 0.5f * sgb[cpg + cc]);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float h = fmaf(fx[k], ha, hb);
            const float t = tanh_approx(h);
            const float u = fmaf(-h, t, h + 1.f);
            const float gz = fg[k] * (fmaf(0.5f, t, 0.5f) * u);
            const float xh = fmaf(fx[k], rstd, nmr);
            fg[k] = gz;
            a0 += gz;
            a1 = fmaf(gz, xh, a1);
          }
          Io<T>::store(bg + i, fg);  // the second pass reads gz = dy * silu'(z) instead of recomputing it
        } else {
          const float be = sgb[cpg + cc];
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float xh = fmaf(fx[k], rstd, nmr);
            float gz = fg[k];
            if (apply_silu) gz *= dsilu_t<T>(fmaf(xh, ga, be));
            fg[k] = gz;
            a0 += gz;
            a1 = fmaf(gz, xh, a1);
          }
          if (apply_silu) Io<T>::store(bg + i, fg);
        }
        tot[0] = fmaf(a0, ga, tot[0]);
        tot[1] = fmaf(a1, ga, tot[1]);
      }
      // per-channel bins, private to the warp (shared-memory float atomics are CAS loops; with eight warps on one address
      // they were a quarter of the stall samples). Fast path: the warp's 32 vectors lie in one channel — one shuffle
      // reduction, the first lane adds the pair to the warp's bin with plain loads / stores. Across a channel boundary:
      // one such round per distinct channel.
      const int c0 = __shfl_sync(0xffffffffu, cc, 0);
      if (__all_sync(0xffffffffu, cc == c0 || !ok)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a0 += __shfl_xor_sync(0xffffffffu, a0, o);
          a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        }
        if (lane == 0 && c0 >= 0) {
          wbins[c0] += a0;
          wbins[cpg + c0] += a1;
        }
      } else {
        unsigned remaining = __ballot_sync(0xffffffffu, ok);
        while (remaining != 0) {
          const int leader = __ffs(remaining) - 1;
          const int c = __shfl_sync(0xffffffffu, cc, leader);
          const bool mine = ok && cc == c;
          float v0 = mine ? a0 : 0.f, v1 = mine ? a1 : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
          }
          if (lane == leader) {
            wbins[c] += v0;
            wbins[cpg + c] += v1;
          }
          __syncwarp();  // the next leader may be another lane updating the same bin
          remaining &= ~__ballot_sync(0xffffffffu, mine);
        }
      }
      __syncwarp();
    }
  }
  }
  cta_sum<2, GB_THREADS>(tot, red);  // (its barriers also order the bin updates before the flush below)
  for (int i = tid; i < 2 * cpg; i += GB_THREADS) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) v += bins[w * 2 * cpg + i];
    float* dst = i < cpg ? dbeta : dgamma;
    if (dst) atomicAdd(dst + g * cpg + (i < cpg ? i : i - cpg), v);
  }
  cluster_sum<2>(tot, partial, CL);
  const float inv = 1.f / count;
  const float m1 = tot[0] * inv, m2 = tot[1] * inv;
  T* os = dx + base + start;
  ChanIter ch2(start + tid * VEC, GB_THREADS * VEC, S);
  const float rm1 = rstd * m1, rm2 = rstd * m2;
  const float c2 = -rm2 * rstd, c0 = -fmaf(rm2, nmr, rm1);  // dx = gz * (rstd gamma) + x * c2 + c0
  for (int i = tid * VEC; i < len; i += GB_THREADS * VEC, ch2.next()) {
    const float rg = rstd * sgb[ch2.cc];
    if constexpr (sizeof(T) == 2) {  // packed f32x2: two elements per instruction
      const uint4 ux = *reinterpret_cast<const uint4*>(bx + i);
      const uint4 ug = *reinterpret_cast<const uint4*>(bg + i);  // gz (pass 1 rewrote it; cta_sum's barriers order those writes)
      const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&ux);
      const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&ug);
      const float2 rg2 = make_float2(rg, rg), c22 = make_float2(c2, c2), c02 = make_float2(c0, c0);
      uint4 uo;
      __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&uo);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 r = __ffma2_rn(__bfloat1622float2(hg[k]), rg2, __ffma2_rn(__bfloat1622float2(hx[k]), c22, c02));
        ho[k] = __floats2bfloat162_rn(r.x, r.y);
      }
      *reinterpret_cast<uint4*>(os + i) = uo;
    } else {
      float fx[VEC], fg[VEC];
      Io<T>::load(bx + i, fx);
      Io<T>::load(bg + i, fg);  // gz (pass 1 rewrote it; cta_sum's barriers order those writes)
#pragma unroll
      for (int k = 0; k < VEC; ++k) fx[k] = fmaf(fg[k], rg, fmaf(fx[k], c2, c0));
      Io<T>::store(os + i, fx);
    }
  }
  if (CL > 1) cg::this_cluster().sync();
}

// Chunking policy. `streams` = staged tensors (1 forward, 2 backward). Prefer the smallest cluster whose chunk lets four
// (forward) CTAs share an SM, but never fewer than ~4 CTAs per SM over the whole grid when the slab can still be split.
constexpr size_t hard_smem() { return 112 * 1024; }
struct BulkPlan {
  int CL = 0, chunk = 0;
  size_t smem = 0;
};
template <typename T>
BulkPlan plan_bulk(int N, int C, int S, int G, int streams, const void* p0, const void* p1, const void* p2) {
  BulkPlan plan;
  constexpr int VEC = Io<T>::VEC;
  const long long count = static_cast<long long>(C / G) * S;
  static const bool two_pass = getenv("VT_GN_TWOPASS") != nullptr;
  if (two_pass) return plan;
  if (S % VEC != 0 || !aligned16(p0) || !aligned16(p1) || (p2 != nullptr && !aligned16(p2))) return plan;
  if (N > 65535 || G > 65535) return plan;
  const long long vectors = count / VEC;
  const size_t extra = (2 + (streams == 2 ? 2 * (GB_THREADS / 32) : 0)) * static_cast<size_t>(C / G) * sizeof(float);
  // soft: four CTAs per SM by shared memory (eight with VT_GN_SOFT_KB=27: an A/B knob); hard: two
  static const size_t soft_kb = getenv("VT_GN_SOFT_KB") != nullptr ? static_cast<size_t>(atoi(getenv("VT_GN_SOFT_KB"))) : 54;
  const size_t soft = soft_kb * 1024, hard = 110 * 1024;
  for (int cl = 1; cl <= 16; cl *= 2) {
    const long long chunk = (vectors + cl - 1) / cl * VEC;
    const size_t bytes = static_cast<size_t>(chunk) * sizeof(T) * streams + extra;
    const long long ctas = static_cast<long long>(N) * G * cl;
    const bool enough_ctas = ctas >= 4 * 148 || chunk * static_cast<long long>(sizeof(T)) <= 8 * 1024;
    if ((bytes <= soft && enough_ctas) || (cl == 16 && bytes <= hard)) {
      plan.CL = cl;
      plan.chunk = static_cast<int>(chunk);
      plan.smem = bytes;
      return plan;
    }
  }
  return plan;
}

template <typename K, typename... Args>
cudaError_t launch_cluster(K kernel, const BulkPlan& plan, int N, int G, cudaStream_t st, Args... args) {
  // function attributes once per kernel (keyed by its address: several instantiations share this template)
  if (first_on_device(reinterpret_cast<const void*>(kernel))) {  // function attributes are per device
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hard_smem()));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.CL, G, N);
  cfg.blockDim = dim3(GB_THREADS);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan.CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace
}  // namespace vt

using namespace vt;

extern "C" {

int vt_groupnorm_silu_fwd(const void* x, void* y, float* mean, float* rstd, const float* gamma, const float* beta,
                          int N, int C, int S, int G, float eps, int apply_silu, int dtype, void* stream) {
  VT_REQUIRE(x && y, VT_ERR_NULL, "vt_groupnorm_silu_fwd: NULL argument");
  VT_REQUIRE(N > 0 && C > 0 && S > 0 && G > 0 && C % G == 0, VT_ERR_SHAPE, "bad shape N=%d C=%d S=%d G=%d", N, C, S, G);
  VT_REQUIRE(N <= 65535, VT_ERR_SHAPE, "N=%d exceeds grid.y", N);
  VT_REQUIRE(static_cast<long long>(C / G) * S < (1LL << 31), VT_ERR_SHAPE, "group too large");
  VT_REQUIRE(dtype == 0 || dtype == 1, VT_ERR_DTYPE, "dtype %d (0=bf16, 1=fp32)", dtype);
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == 0) {
    const BulkPlan plan = plan_bulk<__nv_bfloat16>(N, C, S, G, 1, x, y, nullptr);
    if (plan.CL > 0 &&
        launch_cluster(groupnorm_fwd_bulk_kernel<__nv_bfloat16>, plan, N, G, st, static_cast<const __nv_bfloat16*>(x),
                       static_cast<__nv_bfloat16*>(y), mean, rstd, gamma, beta, C, S, G, eps, apply_silu, plan.chunk,
                       plan.CL) == cudaSuccess)
      return 0;
  } else {
    const BulkPlan plan = plan_bulk<float>(N, C, S, G, 1, x, y, nullptr);
    if (plan.CL > 0 && launch_cluster(groupnorm_fwd_bulk_kernel<float>, plan, N, G, st, static_cast<const float*>(x),
                                      static_cast<float*>(y), mean, rstd, gamma, beta, C, S, G, eps, apply_silu,
                                      plan.chunk, plan.CL) == cudaSuccess)
      return 0;
  }
  (void)cudaGetLastError();  // a refused cluster launch falls through to the two-pass kernel
  dim3 grid(G, N);
  if (dtype == 0) {
    const int vec_ok = (S % 8 == 0) && aligned16(x) && aligned16(y);
#ifdef VT_EXPERIMENTS
    if (vec_ok && getenv("VT_GN_REG") != nullptr && launch_gn_fwd_reg<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), mean,
                                                   rstd, gamma, beta, N, C, S, G, eps, apply_silu, st)) {
      VT_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
#endif
    groupnorm_fwd_kernel<__nv_bfloat16><<<grid, GN_THREADS, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), mean, rstd, gamma, beta, C, S, G, eps,
        apply_silu, vec_ok);
  } else {
    const int vec_ok = (S % 4 == 0) && aligned16(x) && aligned16(y);
#ifdef VT_EXPERIMENTS
    if (vec_ok && getenv("VT_GN_REG") != nullptr && launch_gn_fwd_reg<float>(static_cast<const float*>(x), static_cast<float*>(y), mean, rstd, gamma, beta, N, C, S,
                                           G, eps, apply_silu, st)) {
      VT_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
#endif
    groupnorm_fwd_kernel<float><<<grid, GN_THREADS, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(y), mean,
                                                             rstd, gamma, beta, C, S, G, eps, apply_silu, vec_ok);
  }
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vt_groupnorm_silu_bwd(const void* dy, const void* x, const float* mean, const float* rstd, void* dx,
                          const float* gamma, const float* beta, float* dgamma, float* dbeta, int N, int C, int S, int G,
                          int apply_silu, int dtype, void* stream) {
  VT_REQUIRE(dy && x && mean && rstd && dx, VT_ERR_NULL, "vt_groupnorm_silu_bwd: NULL argument");
  VT_REQUIRE(N > 0 && C > 0 && S > 0 && G > 0 && C % G == 0, VT_ERR_SHAPE, "bad shape N=%d C=%d S=%d G=%d", N, C, S, G);
  VT_REQUIRE(N <= 65535, VT_ERR_SHAPE, "N=%d exceeds grid.y", N);
  VT_REQUIRE(dtype == 0 || dtype == 1, VT_ERR_DTYPE, "dtype %d (0=bf16, 1=fp32)", dtype);
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == 0) {
    const BulkPlan plan = plan_bulk<__nv_bfloat16>(N, C, S, G, 2, x, dy, dx);
    if (plan.CL > 0 &&
        launch_cluster(groupnorm_bwd_bulk_kernel<__nv_bfloat16>, plan, N, G, st, static_cast<const __nv_bfloat16*>(dy),
                       static_cast<const __nv_bfloat16*>(x), mean, rstd, static_cast<__nv_bfloat16*>(dx), gamma, beta,
                       dgamma, dbeta, C, S, G, apply_silu, plan.chunk, plan.CL) == cudaSuccess)
      return 0;
  } else {
    const BulkPlan plan = plan_bulk<float>(N, C, S, G, 2, x, dy, dx);
    if (plan.CL > 0 &&
        launch_cluster(groupnorm_bwd_bulk_kernel<float>, plan, N, G, st, static_cast<const float*>(dy),
                       static_cast<const float*>(x), mean, rstd, static_cast<float*>(dx), gamma, beta, dgamma, dbeta, C, S,
                       G, apply_silu, plan.chunk, plan.CL) == cudaSuccess)
      return 0;
  }
  (void)cudaGetLastError();
  dim3 grid(G, N);
  if (dtype == 0)
    groupnorm_bwd_kernel<__nv_bfloat16><<<grid, GN_THREADS, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), mean, rstd,
        static_cast<__nv_bfloat16*>(dx), gamma, beta, dgamma, dbeta, C, S, G, apply_silu,
        (S % 8 == 0) && aligned16(x) && aligned16(dy) && aligned16(dx));
  else
    groupnorm_bwd_kernel<float><<<grid, GN_THREADS, 0, st>>>(static_cast<const float*>(dy), static_cast<const float*>(x),
                                                             mean, rstd, static_cast<float*>(dx), gamma, beta, dgamma,
                                                             dbeta, C, S, G, apply_silu,
                                                             (S % 4 == 0) && aligned16(x) && aligned16(dy) && aligned16(dx));
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
