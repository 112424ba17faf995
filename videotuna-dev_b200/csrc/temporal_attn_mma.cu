// temporal_attn_mma.cu — forward of the N <= 32 micro-attention on warp-level tensor-core MMAs.
//
// The SIMT kernel in temporal_attn.cu spends ~1500 instructions per (sequence, head) pair on scalar FMAs and shared-memory
// reads for 8 KB of q/k/v/o and reaches 15 % of the HBM roofline. The whole problem of one pair is a 16 x 16 (or 32 x 32)
// score tile: here a warp stages q, k, v with cp.async, computes S = Q K^T with mma.sync.m16n8k16 (bf16 in, fp32 out),
// does the softmax on the accumulator fragments (row max / sum by two quad shuffles), feeds P back as the A operand
// (the accumulator layout of two adjacent n-tiles IS the A fragment of the next k-step) and computes O = P V with V read
// through ldmatrix.trans — ~150 instructions per pair. tcgen05 is not an option at this size: its smallest tile is 128
// rows (64 with half-rate), one pair fills 16.
// Replaces the forward of lvdm CrossAttention.forward (videotuna/models/lvdm/modules/attention.py:126-149) as called by
// TemporalTransformer over t = 16 frames (:475-519); mask semantics as temporal_attn.cu.
#include <cfloat>
#include <cstdlib>
#include <cuda_bf16.h>

#include "capi_util.h"

namespace vt {
namespace {

using bf16 = __nv_bfloat16;

struct MmaArgs {
  const bf16 *q, *k, *v;
  bf16* o;
  const float* mask;
  int64_t q_s[3], k_s[3], v_s[3], o_s[3];  // (b, n, h) element strides
  int B, N, H;
  float scale_log2;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MT: query / key tiles of 16 (1: N <= 16, 2: N <= 32). One warp per (sequence, head) pair; `WARPS` pairs per CTA pass.
template <int D, int MT>
__global__ void __launch_bounds__(128) temporal_mma_fwd_kernel(const MmaArgs a) {
  constexpr int R = 16 * MT;      // padded rows
  constexpr int PITCH = D + 8;    // bf16 elements: 16-byte row skew keeps ldmatrix conflict-free
  constexpr int KT = D / 16;      // k-steps of Q K^T
  constexpr int NTS = 2 * MT;     // 8-key n-tiles of S
  constexpr int NTO = D / 8;      // 8-dim n-tiles of O
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* sq = reinterpret_cast<bf16*>(smem_raw) + warp * 3 * R * PITCH;
  bf16* sk = sq + R * PITCH;
  bf16* sv = sk + R * PITCH;
  const int N = a.N;
  const int g = lane >> 2, t = lane & 3;
  const int64_t pairs = static_cast<int64_t>(a.B) * a.H;
  // rows N..R-1 are never loaded: zero them once (their scores are masked, but 0 * garbage must not be NaN in P V)
  for (int idx = lane; idx < (R - N) * (D / 8); idx += 32) {
    const int r = N + idx / (D / 8), c = (idx % (D / 8)) * 8;
    *reinterpret_cast<uint4*>(sq + r * PITCH + c) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(sk + r * PITCH + c) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(sv + r * PITCH + c) = make_uint4(0, 0, 0, 0);
  }
  const int nwarps = blockDim.x >> 5;
  for (int64_t pair = static_cast<int64_t>(blockIdx.x) * nwarps + warp; pair < pairs;
       pair += static_cast<int64_t>(gridDim.x) * nwarps) {
    const int b = static_cast<int>(pair / a.H), h = static_cast<int>(pair % a.H);
    const bf16* gq = a.q + b * a.q_s[0] + h * a.q_s[2];
    const bf16* gk = a.k + b * a.k_s[0] + h * a.k_s[2];
    const bf16* gv = a.v + b * a.v_s[0] + h * a.v_s[2];
    for (int idx = lane; idx < N * (D / 8); idx += 32) {
      const int r = idx / (D / 8), c = (idx % (D / 8)) * 8;
      cp_async16(sq + r * PITCH + c, gq + r * a.q_s[1] + c);
      cp_async16(sk + r * PITCH + c, gk + r * a.k_s[1] + c);
      cp_async16(sv + r * PITCH + c, gv + r * a.v_s[1] + c);
    }
    cp_async_wait_all();
    __syncwarp();

    // ---- S = Q K^T : acc[mt][nt][4], rows mt*16 + g (+8), keys nt*8 + 2t (+1) ----
    float s[MT][NTS][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[mt][nt][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      uint32_t qa[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        ldmatrix_x4(qa[mt], sq + (mt * 16 + (lane & 15)) * PITCH + kt * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int np = 0; np < MT; ++np) {  // pairs of key n-tiles (16 keys)
        uint32_t kb[4];  // {b0, b1} of n-tile 2np, {b0, b1} of n-tile 2np + 1
        ldmatrix_x4(kb, sk + (np * 16 + (lane & 7) + (lane >> 4) * 8) * PITCH + kt * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma16816(s[mt][2 * np], qa[mt], kb[0], kb[1]);
          mma16816(s[mt][2 * np + 1], qa[mt], kb[2], kb[3]);
        }
      }
    }
    // ---- softmax over the keys of each query row (exp2 domain) ----
    uint32_t pa[MT][MT][4];  // P as A fragments: [query tile][key k-step of 16]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = mt * 16 + g + (i >> 1) * 8, key = nt * 8 + 2 * t + (i & 1);
          float x = s[mt][nt][i] * a.scale_log2;
          if (a.mask != nullptr && row < N && key < N && !(a.mask[row * N + key] > 0.5f)) x = -FLT_MAX;
          if (key >= N) x = -INFINITY;
          s[mt][nt][i] = x;
          mx[i >> 1] = fmaxf(mx[i >> 1], x);
        }
      float sum[2] = {0.f, 0.f};
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        mx[hrow] = fmaxf(mx[hrow], __shfl_xor_sync(0xffffffffu, mx[hrow], 1));
        mx[hrow] = fmaxf(mx[hrow], __shfl_xor_sync(0xffffffffu, mx[hrow], 2));
      }
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float e = ex2f(s[mt][nt][i] - mx[i >> 1]);  // masked (-FLT_MAX) and padded (-inf) keys give exact 0
          s[mt][nt][i] = e;
          sum[i >> 1] += e;
        }
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        sum[hrow] += __shfl_xor_sync(0xffffffffu, sum[hrow], 1);
        sum[hrow] += __shfl_xor_sync(0xffffffffu, sum[hrow], 2);
        sum[hrow] = 1.f / sum[hrow];
      }
#pragma unroll
      for (int ks = 0; ks < MT; ++ks) {  // k-step of 16 keys = n-tiles 2ks, 2ks + 1
        pa[mt][ks][0] = pack2(s[mt][2 * ks][0] * sum[0], s[mt][2 * ks][1] * sum[0]);
        pa[mt][ks][1] = pack2(s[mt][2 * ks][2] * sum[1], s[mt][2 * ks][3] * sum[1]);
        pa[mt][ks][2] = pack2(s[mt][2 * ks + 1][0] * sum[0], s[mt][2 * ks + 1][1] * sum[0]);
        pa[mt][ks][3] = pack2(s[mt][2 * ks + 1][2] * sum[1], s[mt][2 * ks + 1][3] * sum[1]);
      }
    }
    // ---- O = P V : V read transposed (B operand wants keys contiguous) ----
    __syncwarp();  // every lane is done reading Q: its buffer becomes the O staging tile
#pragma unroll
    for (int np = 0; np < NTO / 2; ++np) {  // pairs of 8-dim n-tiles
      float o[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[mt][j][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < MT; ++ks) {
        uint32_t vb[4];  // {b0, b1} of dims np*16 .. +7, {b0, b1} of dims np*16 + 8 .. +15, for keys ks*16 .. +15
        ldmatrix_x4_trans(vb, sv + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + np * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma16816(o[mt][0], pa[mt][ks], vb[0], vb[1]);
          mma16816(o[mt][1], pa[mt][ks], vb[2], vb[3]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = np * 16 + j * 8 + 2 * t;
          *reinterpret_cast<uint32_t*>(sq + (mt * 16 + g) * PITCH + col) = pack2(o[mt][j][0], o[mt][j][1]);
          *reinterpret_cast<uint32_t*>(sq + (mt * 16 + g + 8) * PITCH + col) = pack2(o[mt][j][2], o[mt][j][3]);
        }
    }
    __syncwarp();
    bf16* go = a.o + b * a.o_s[0] + h * a.o_s[2];
    for (int idx = lane; idx < N * (D / 8); idx += 32) {
      const int r = idx / (D / 8), c = (idx % (D / 8)) * 8;
      *reinterpret_cast<uint4*>(go + r * a.o_s[1] + c) = *reinterpret_cast<const uint4*>(sq + r * PITCH + c);
    }
    __syncwarp();
    // rows N..R-1 of the Q buffer were overwritten by O of padded rows: restore the zeros the next pair relies on
    for (int idx = lane; idx < (R - N) * (D / 8); idx += 32) {
      const int r = N + idx / (D / 8), c = (idx % (D / 8)) * 8;
      *reinterpret_cast<uint4*>(sq + r * PITCH + c) = make_uint4(0, 0, 0, 0);
    }
  }
}

template <int D, int MT>
cudaError_t launch_mma(const MmaArgs& a, cudaStream_t st) {
  constexpr int WARPS = 4;
  constexpr int bytes = WARPS * 3 * (16 * MT) * (D + 8) * 2;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    cudaError_t e = cudaFuncSetAttribute(temporal_mma_fwd_kernel<D, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
  }
  const int64_t pairs = static_cast<int64_t>(a.B) * a.H;
  int64_t blocks = (pairs + WARPS - 1) / WARPS;
  static const int64_t per_sm = getenv("VT_TEMPORAL_CTAS") != nullptr ? atoi(getenv("VT_TEMPORAL_CTAS")) : 8;
  const int64_t cap = 148 * per_sm;
  if (blocks > cap) blocks = cap;
  temporal_mma_fwd_kernel<D, MT><<<static_cast<unsigned>(blocks), WARPS * 32, bytes, st>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: S, P as in the forward; dP = dO V^T; dS = P o (dP - rowsum(P o dP)); dQ = scale dS K; dV = P^T dO;
// dK = scale dS^T Q. P and dS go through a small bf16 tile in shared memory so that their transposes can be read back
// as A fragments with ldmatrix.trans; dQ takes dS straight from the accumulator registers.
// ---------------------------------------------------------------------------------------------------------------------
struct MmaBwdArgs {
  const bf16 *q, *k, *v, *g;
  bf16 *dq, *dk, *dv;
  const float* mask;
  int64_t q_s[3], k_s[3], v_s[3], g_s[3];
  int B, N, H;
  float scale, scale_log2;
};

template <int D, int MT>
__global__ void __launch_bounds__(128) temporal_mma_bwd_kernel(const MmaBwdArgs a) {
  constexpr int R = 16 * MT, PITCH = D + 8, KT = D / 16, NTS = 2 * MT, NTO = D / 8, PP = R + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int PER_WARP = 4 * R * PITCH + 2 * R * PP;  // bf16 elements
  bf16* sq = reinterpret_cast<bf16*>(smem_raw) + warp * PER_WARP;
  bf16* sk = sq + R * PITCH;
  bf16* sv = sk + R * PITCH;
  bf16* sg = sv + R * PITCH;   // dO
  bf16* sp = sg + R * PITCH;   // P   [query][key]
  bf16* sd = sp + R * PP;      // dS  [query][key]
  const int N = a.N;
  const int g = lane >> 2, t = lane & 3;
  const int64_t pairs = static_cast<int64_t>(a.B) * a.H;
  auto zero_pad = [&](bf16* m) {
    for (int idx = lane; idx < (R - N) * (D / 8); idx += 32) {
      const int r = N + idx / (D / 8), c = (idx % (D / 8)) * 8;
      *reinterpret_cast<uint4*>(m + r * PITCH + c) = make_uint4(0, 0, 0, 0);
    }
  };
  zero_pad(sq); zero_pad(sk); zero_pad(sv); zero_pad(sg);
  const int nwarps = blockDim.x >> 5;
  for (int64_t pair = static_cast<int64_t>(blockIdx.x) * nwarps + warp; pair < pairs;
       pair += static_cast<int64_t>(gridDim.x) * nwarps) {
    const int b = static_cast<int>(pair / a.H), h = static_cast<int>(pair % a.H);
    const bf16* gq = a.q + b * a.q_s[0] + h * a.q_s[2];
    const bf16* gk = a.k + b * a.k_s[0] + h * a.k_s[2];
    const bf16* gv = a.v + b * a.v_s[0] + h * a.v_s[2];
    const bf16* gg = a.g + b * a.g_s[0] + h * a.g_s[2];
    for (int idx = lane; idx < N * (D / 8); idx += 32) {
      const int r = idx / (D / 8), c = (idx % (D / 8)) * 8;
      cp_async16(sq + r * PITCH + c, gq + r * a.q_s[1] + c);
      cp_async16(sk + r * PITCH + c, gk + r * a.k_s[1] + c);
      cp_async16(sv + r * PITCH + c, gv + r * a.v_s[1] + c);
      cp_async16(sg + r * PITCH + c, gg + r * a.g_s[1] + c);
    }
    cp_async_wait_all();
    __syncwarp();

    // ---- S = Q K^T and dP = dO V^T (same fragment layout: rows = queries, cols = keys) ----
    float s[MT][NTS][4], dp[MT][NTS][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[mt][nt][i] = 0.f; dp[mt][nt][i] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      uint32_t qa[MT][4], ga[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        ldmatrix_x4(qa[mt], sq + (mt * 16 + (lane & 15)) * PITCH + kt * 16 + (lane >> 4) * 8);
        ldmatrix_x4(ga[mt], sg + (mt * 16 + (lane & 15)) * PITCH + kt * 16 + (lane >> 4) * 8);
      }
#pragma unroll
      for (int np = 0; np < MT; ++np) {
        uint32_t kb[4], vb[4];
        const int off = (np * 16 + (lane & 7) + (lane >> 4) * 8) * PITCH + kt * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(kb, sk + off);
        ldmatrix_x4(vb, sv + off);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma16816(s[mt][2 * np], qa[mt], kb[0], kb[1]);
          mma16816(s[mt][2 * np + 1], qa[mt], kb[2], kb[3]);
          mma16816(dp[mt][2 * np], ga[mt], vb[0], vb[1]);
          mma16816(dp[mt][2 * np + 1], ga[mt], vb[2], vb[3]);
        }
      }
    }
    // ---- P (normalised), dS = P o (dP - delta); both to shared memory as bf16; dS also kept as A fragments ----
    uint32_t da[MT][MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = mt * 16 + g + (i >> 1) * 8, key = nt * 8 + 2 * t + (i & 1);
          float x = s[mt][nt][i] * a.scale_log2;
          if (a.mask != nullptr && row < N && key < N && !(a.mask[row * N + key] > 0.5f)) x = -FLT_MAX;
          if (key >= N) x = -INFINITY;
          s[mt][nt][i] = x;
          mx[i >> 1] = fmaxf(mx[i >> 1], x);
        }
      float sum[2] = {0.f, 0.f};
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        mx[hrow] = fmaxf(mx[hrow], __shfl_xor_sync(0xffffffffu, mx[hrow], 1));
        mx[hrow] = fmaxf(mx[hrow], __shfl_xor_sync(0xffffffffu, mx[hrow], 2));
      }
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float e = ex2f(s[mt][nt][i] - mx[i >> 1]);
          s[mt][nt][i] = e;
          sum[i >> 1] += e;
        }
      float delta[2] = {0.f, 0.f};
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        sum[hrow] += __shfl_xor_sync(0xffffffffu, sum[hrow], 1);
        sum[hrow] += __shfl_xor_sync(0xffffffffu, sum[hrow], 2);
        sum[hrow] = 1.f / sum[hrow];
      }
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s[mt][nt][i] *= sum[i >> 1];  // P
          delta[i >> 1] += s[mt][nt][i] * dp[mt][nt][i];
        }
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        delta[hrow] += __shfl_xor_sync(0xffffffffu, delta[hrow], 1);
        delta[hrow] += __shfl_xor_sync(0xffffffffu, delta[hrow], 2);
      }
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dp[mt][nt][i] = s[mt][nt][i] * (dp[mt][nt][i] - delta[i >> 1]);  // dS
        const int col = nt * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(sp + (mt * 16 + g) * PP + col) = pack2(s[mt][nt][0], s[mt][nt][1]);
        *reinterpret_cast<uint32_t*>(sp + (mt * 16 + g + 8) * PP + col) = pack2(s[mt][nt][2], s[mt][nt][3]);
        *reinterpret_cast<uint32_t*>(sd + (mt * 16 + g) * PP + col) = pack2(dp[mt][nt][0], dp[mt][nt][1]);
        *reinterpret_cast<uint32_t*>(sd + (mt * 16 + g + 8) * PP + col) = pack2(dp[mt][nt][2], dp[mt][nt][3]);
      }
#pragma unroll
      for (int ks = 0; ks < MT; ++ks) {
        da[mt][ks][0] = pack2(dp[mt][2 * ks][0], dp[mt][2 * ks][1]);
        da[mt][ks][1] = pack2(dp[mt][2 * ks][2], dp[mt][2 * ks][3]);
        da[mt][ks][2] = pack2(dp[mt][2 * ks + 1][0], dp[mt][2 * ks + 1][1]);
        da[mt][ks][3] = pack2(dp[mt][2 * ks + 1][2], dp[mt][2 * ks + 1][3]);
      }
    }
    __syncwarp();
    // A fragments of P^T and dS^T: rows = keys, k = queries, read transposed from the [query][key] tiles
    uint32_t pt[MT][MT][4], dt[MT][MT][4];  // [key tile][query k-step]
#pragma unroll
    for (int mk = 0; mk < MT; ++mk)
#pragma unroll
      for (int kq = 0; kq < MT; ++kq) {
        // matrices: (keys 0-7, q 0-7), (keys 8-15, q 0-7), (keys 0-7, q 8-15), (keys 8-15, q 8-15) of this 16 x 16 block;
        // stored as [q][key], so each 8 x 8 source tile is rows q, cols key and .trans delivers (row = key, k = q)
        const int off = (kq * 16 + (lane & 7) + (lane >> 4) * 8) * PP + mk * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(pt[mk][kq], sp + off);
        ldmatrix_x4_trans(dt[mk][kq], sd + off);
      }
    const int64_t gbase = (static_cast<int64_t>(b) * N * a.H + h) * D;  // contiguous (B, N, H, D) gradients
    const int64_t gn = static_cast<int64_t>(a.H) * D;
    auto stage_out = [&](bf16* buf, float (&o)[MT][2][4], int np, float mul) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = np * 16 + j * 8 + 2 * t;
          *reinterpret_cast<uint32_t*>(buf + (mt * 16 + g) * PITCH + col) = pack2(o[mt][j][0] * mul, o[mt][j][1] * mul);
          *reinterpret_cast<uint32_t*>(buf + (mt * 16 + g + 8) * PITCH + col) = pack2(o[mt][j][2] * mul, o[mt][j][3] * mul);
        }
    };
    auto write_out = [&](bf16* dst, const bf16* buf) {
      for (int idx = lane; idx < N * (D / 8); idx += 32) {
        const int r = idx / (D / 8), c = (idx % (D / 8)) * 8;
        *reinterpret_cast<uint4*>(dst + gbase + r * gn + c) = *reinterpret_cast<const uint4*>(buf + r * PITCH + c);
      }
    };
    // ---- dV = P^T dO (B = dO read transposed), staged into the V buffer (V is dead) ----
#pragma unroll
    for (int np = 0; np < NTO / 2; ++np) {
      float o[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[mt][j][i] = 0.f;
#pragma unroll
      for (int kq = 0; kq < MT; ++kq) {
        uint32_t gb[4];
        ldmatrix_x4_trans(gb, sg + (kq * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + np * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int mk = 0; mk < MT; ++mk) {
          mma16816(o[mk][0], pt[mk][kq], gb[0], gb[1]);
          mma16816(o[mk][1], pt[mk][kq], gb[2], gb[3]);
        }
      }
      stage_out(sv, o, np, 1.f);
    }
    // ---- dQ = scale dS K (B = K read transposed), staged into the dO buffer once dV no longer needs it ----
    __syncwarp();
#pragma unroll
    for (int np = 0; np < NTO / 2; ++np) {
      float o[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[mt][j][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < MT; ++ks) {
        uint32_t kb[4];
        ldmatrix_x4_trans(kb, sk + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + np * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma16816(o[mt][0], da[mt][ks], kb[0], kb[1]);
          mma16816(o[mt][1], da[mt][ks], kb[2], kb[3]);
        }
      }
      stage_out(sg, o, np, a.scale);
    }
    // ---- dK = scale dS^T Q (B = Q read transposed), staged into the K buffer once dQ no longer needs it ----
    __syncwarp();
#pragma unroll
    for (int np = 0; np < NTO / 2; ++np) {
      float o[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[mt][j][i] = 0.f;
#pragma unroll
      for (int kq = 0; kq < MT; ++kq) {
        uint32_t qb[4];
        ldmatrix_x4_trans(qb, sq + (kq * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + np * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int mk = 0; mk < MT; ++mk) {
          mma16816(o[mk][0], dt[mk][kq], qb[0], qb[1]);
          mma16816(o[mk][1], dt[mk][kq], qb[2], qb[3]);
        }
      }
      stage_out(sk, o, np, a.scale);
    }
    __syncwarp();
    write_out(a.dv, sv);
    write_out(a.dq, sg);
    write_out(a.dk, sk);
    __syncwarp();
    zero_pad(sk); zero_pad(sv); zero_pad(sg);  // padded rows were overwritten by staged outputs
  }
}

template <int D, int MT>
cudaError_t launch_mma_bwd(const MmaBwdArgs& a, cudaStream_t st) {
  constexpr int WARPS = 4;
  constexpr int R = 16 * MT;
  constexpr int bytes = WARPS * (4 * R * (D + 8) + 2 * R * (R + 8)) * 2;
  static char cfg_site;  // per call site; the attribute is per DEVICE (first_on_device)
  if (first_on_device(&cfg_site)) {
    cudaError_t e = cudaFuncSetAttribute(temporal_mma_bwd_kernel<D, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
  }
  const int64_t pairs = static_cast<int64_t>(a.B) * a.H;
  int64_t blocks = (pairs + WARPS - 1) / WARPS;
  static const int64_t per_sm = getenv("VT_TEMPORAL_CTAS") != nullptr ? atoi(getenv("VT_TEMPORAL_CTAS")) : 8;
  const int64_t cap = 148 * per_sm;
  if (blocks > cap) blocks = cap;
  temporal_mma_bwd_kernel<D, MT><<<static_cast<unsigned>(blocks), WARPS * 32, bytes, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

// Forward through the tensor-core kernel. Returns cudaErrorNotSupported for shapes it does not take.
cudaError_t temporal_attn_fwd_mma(const void* q, const void* k, const void* v, void* o, const float* mask,
                                  const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os, int B, int N,
                                  int H, int D, float scale, cudaStream_t st) {
  if (N < 1 || N > 32 || (D != 64 && D != 128)) return cudaErrorNotSupported;
  MmaArgs a{};
  a.q = static_cast<const bf16*>(q);
  a.k = static_cast<const bf16*>(k);
  a.v = static_cast<const bf16*>(v);
  a.o = static_cast<bf16*>(o);
  a.mask = mask;
  for (int i = 0; i < 3; ++i) {
    a.q_s[i] = qs[i];
    a.k_s[i] = ks[i];
    a.v_s[i] = vs[i];
    a.o_s[i] = os[i];
  }
  a.B = B;
  a.N = N;
  a.H = H;
  a.scale_log2 = scale * 1.4426950408889634f;
  if (D == 64) return N <= 16 ? launch_mma<64, 1>(a, st) : launch_mma<64, 2>(a, st);
  return N <= 16 ? launch_mma<128, 1>(a, st) : launch_mma<128, 2>(a, st);
}

}  // namespace vt

namespace vt {
cudaError_t temporal_attn_bwd_mma(const void* dout, const void* q, const void* k, const void* v, void* dq, void* dk, void* dv,
                                  const float* mask, const int64_t* gs, const int64_t* qs, const int64_t* ks, const int64_t* vs,
                                  int B, int N, int H, int D, float scale, cudaStream_t st) {
  if (N < 1 || N > 32 || (D != 64 && D != 128)) return cudaErrorNotSupported;
  MmaBwdArgs a{};
  a.q = static_cast<const bf16*>(q);
  a.k = static_cast<const bf16*>(k);
  a.v = static_cast<const bf16*>(v);
  a.g = static_cast<const bf16*>(dout);
  a.dq = static_cast<bf16*>(dq);
  a.dk = static_cast<bf16*>(dk);
  a.dv = static_cast<bf16*>(dv);
  a.mask = mask;
  for (int i = 0; i < 3; ++i) {
    a.q_s[i] = qs[i];
    a.k_s[i] = ks[i];
    a.v_s[i] = vs[i];
    a.g_s[i] = gs[i];
  }
  a.B = B;
  a.N = N;
  a.H = H;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  if (D == 64) return N <= 16 ? launch_mma_bwd<64, 1>(a, st) : launch_mma_bwd<64, 2>(a, st);
  return N <= 16 ? launch_mma_bwd<128, 1>(a, st) : launch_mma_bwd<128, 2>(a, st);
}
}  // namespace vt
