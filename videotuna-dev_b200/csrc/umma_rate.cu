// Built only with -DVT_EXPERIMENTS (tools/build_variant.sh): not part of the product library.
#ifdef VT_EXPERIMENTS
// umma_rate.cu — tcgen05.mma issue-rate microbenchmark (tools/umma_rate.py). Measures cycles per 128 x N x 16 bf16 MMA
// for the operand sourcings the attention kernels use, optionally with other warps streaming shared-memory stores, so
// the kernels' MMA-floor arithmetic in DESIGN.md rests on measured numbers. Test infrastructure only.
#include "capi_util.h"
#include "sm100_ptx.cuh"

namespace vt {
int lib_init();
namespace {

// mode 0: SS, A K-major, B K-major   (S = Q K^T)          mode 1: TS, A in TMEM, B MN-major (O += P V)
// mode 2: SS, A MN-major, B MN-major (dQ = dS K)          mode 3: SS, A K-major, B MN-major (dK += dS^T Q)
template <int MODE, int N>
__global__ void __launch_bounds__(256, 1)
umma_rate_kernel(int iters, int noise_warps, int dep, long long* cycles_out) {
  constexpr int mode = MODE, n = N;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int CHUNK = 128 * 128;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 2 * CHUNK;
  uint8_t* noise = smem + 4 * CHUNK;  // 32 KB scratch for the store stream
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * CHUNK);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(bars + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 6 * CHUNK / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    fence_mbar_init();
    *stop = 0;
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, n, mode == 2 ? 1 : 0, mode == 0 ? 0 : 1);
    const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        // dep == 1: all eight k-steps accumulate into one tile (as in the kernels); dep == 0: eight independent tiles
        const uint32_t d = tmem + (dep ? (it & 1) * 128 : (kk & 1) * 128);
        if (mode == 0) {
          umma_ss(d, umma_desc_sw128(a_addr + (kk >> 2) * CHUNK + (kk & 3) * 32, 16, 1024),
                  umma_desc_sw128(b_addr + (kk >> 2) * CHUNK + (kk & 3) * 32, 16, 1024), idesc, kk != 0);
        } else if (mode == 1) {
          umma_ts(d, tmem + 256 + kk * 8, umma_desc_sw128(b_addr + kk * 2048, CHUNK, 1024), idesc, kk != 0);
        } else if (mode == 2) {
          umma_ss(d, umma_desc_sw128(a_addr + kk * 2048, CHUNK, 1024), umma_desc_sw128(b_addr + kk * 2048, CHUNK, 1024),
                  idesc, kk != 0);
        } else {
          umma_ss(d, umma_desc_sw128(a_addr + (kk >> 2) * CHUNK + (kk & 3) * 32, 16, 1024),
                  umma_desc_sw128(b_addr + kk * 2048, CHUNK, 1024), idesc, kk != 0);
        }
      }
    }
    tc_commit(bars);
    mbar_wait(bars, 0, 0x910);
    const long long t1 = clock64();
    *stop = 1;
    cycles_out[blockIdx.x] = t1 - t0;
  } else if (warp >= 1 && warp <= noise_warps) {
    // stream 16-byte stores over the scratch region until the issuer is done
    uint4* dst = reinterpret_cast<uint4*>(noise) + (warp - 1) * 256 + lane;
    uint4 v = make_uint4(lane, warp, 0, 0);
    while (*stop == 0) {
#pragma unroll
      for (int r = 0; r < 8; ++r) dst[r * 32] = v;
      v.z++;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace vt

using namespace vt;

template <int MODE, int N>
static cudaError_t launch_rate(int iters, int noise_warps, int dep, int blocks, long long* out, cudaStream_t st) {
  const int bytes = 6 * 128 * 128 + 64;
  cudaError_t e = cudaFuncSetAttribute(umma_rate_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  umma_rate_kernel<MODE, N><<<blocks, 256, bytes, st>>>(iters, noise_warps, dep, out);
  return cudaGetLastError();
}

extern "C" int vt_umma_rate(int mode, int n, int iters, int noise_warps, int blocks, long long* cycles_out, void* stream) {
  VT_REQUIRE(cycles_out != nullptr, VT_ERR_NULL, "vt_umma_rate: NULL output");
  // mode bit 4 (16) selects independent accumulators instead of one accumulate chain
  const int dep = (mode & 16) ? 0 : 1;
  mode &= 15;
  VT_REQUIRE(mode >= 0 && mode <= 3 && (n == 64 || n == 128) && iters > 0 && blocks > 0 && noise_warps >= 0 &&
                 noise_warps <= 7, VT_ERR_SHAPE, "vt_umma_rate: bad argument");
  if (int rc = lib_init()) return rc;
  auto st = static_cast<cudaStream_t>(stream);
#define VT_RATE_CASE(M, NN) \
  if (mode == M && n == NN) VT_CHECK_CUDA((launch_rate<M, NN>(iters, noise_warps, dep, blocks, cycles_out, st)))
  VT_RATE_CASE(0, 128); VT_RATE_CASE(0, 64); VT_RATE_CASE(1, 128); VT_RATE_CASE(1, 64);
  VT_RATE_CASE(2, 128); VT_RATE_CASE(2, 64); VT_RATE_CASE(3, 128); VT_RATE_CASE(3, 64);
#undef VT_RATE_CASE
  return 0;
}

// ---- TMA reduce-add (fp32) throughput: the dQ accumulation path of the backward kernel -----------------------------
namespace vt {
int make_tmap_4d(CUtensorMap* out, const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t D, int64_t L,
                 int64_t H, int64_t B, const int64_t* strides, int box_d, int box_rows);
namespace {
// Each CTA issues `iters` reductions of one 128-row x 32-column fp32 box (16 KB) with `depth` bulk groups in flight.
// spread == 0: every CTA targets the same rows at the same time (what unstaggered backward CTAs do);
// spread == 1: CTA b starts at row tile b and walks from there.
// Concurrent TMA loads (what the backward producer does while the drain reduces): thread 32 streams 16 KB boxes of a
// bf16 tensor into a second smem region, 4 in flight, until the reducing thread is done. load_mode 0: no loads.
__global__ void __launch_bounds__(128, 1)
tma_mixed_rate_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_ld, int iters,
                      int n_tiles, int load_mode, long long* cycles_out, long long* loads_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8 * 16384);
  volatile int* stop = reinterpret_cast<volatile int*>(bars + 4);
  for (int i = threadIdx.x; i < 4 * 16384 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
    *stop = 0;
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int tile = (it / 4 + blockIdx.x) % n_tiles;
      tma_reduce_add_4d(&tm, smem + (it & 1) * 16384, (it & 3) * 32, tile * 128, 0, 0);
      tma_commit_group();
      tma_wait_group_read<1>();
    }
    tma_wait_group<0>();
    cycles_out[blockIdx.x] = clock64() - t0;
    *stop = 1;
  } else if (threadIdx.x == 32 && load_mode != 0) {
    long long n = 0;
    int it = 0;
    while (*stop == 0) {
      const int s = it & 3;
      if (it >= 4) mbar_wait(bars + s, ((it >> 2) - 1) & 1, 0x920);
      mbar_arrive_expect_tx(bars + s, 16384);
      tma_load_4d(smem + (4 + s) * 16384, &tm_ld, bars + s, (it & 1) * 64, ((it >> 1) % (n_tiles * 2)) * 64 % (n_tiles * 128 - 128), 0, 0);
      ++it;
      ++n;
    }
    for (int k = 0; k < 4 && k < it; ++k) {
      const int j = it - 1 - k;
      mbar_wait(bars + (j & 3), (j >> 2) & 1, 0x921);
    }
    loads_out[blockIdx.x] = n;
  }
}

__global__ void __launch_bounds__(128, 1)
tma_reduce_rate_kernel(const __grid_constant__ CUtensorMap tm, float* linear, int iters, int depth, int spread,
                       int n_tiles, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 4 * 16384 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int tile = (it / 4 + (spread ? blockIdx.x : 0)) % n_tiles;
      if (linear != nullptr) {
        // 1-D bulk reduction of 16 KB of contiguous floats (no tensor map): rows of the accumulator are contiguous
        float* dst = linear + (static_cast<size_t>(tile) * 4 + (it & 3)) * 4096;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                     "r"(smem_u32(smem + (it & 3) * 16384)), "r"(16384)
                     : "memory");
      } else {
        tma_reduce_add_4d(&tm, smem + (it & 3) * 16384, (it & 3) * 32, tile * 128, 0, 0);
      }
      tma_commit_group();
      if (depth == 1) tma_wait_group_read<0>();
      else if (depth == 2) tma_wait_group_read<1>();
      else if (depth == 3) tma_wait_group_read<2>();
      else tma_wait_group_read<3>();
    }
    tma_wait_group<0>();
    cycles_out[blockIdx.x] = clock64() - t0;
  }
}
}  // namespace
}  // namespace vt

extern "C" int vt_tma_mixed_rate(float* acc, const void* src_bf16, int n_tiles, int iters, int load_mode, int blocks,
                                 long long* cycles_out, long long* loads_out, void* stream) {
  VT_REQUIRE(acc && src_bf16 && cycles_out && loads_out, VT_ERR_NULL, "vt_tma_mixed_rate: NULL argument");
  if (int rc = lib_init()) return rc;
  CUtensorMap tm, tm_ld;
  const int64_t st[3] = {static_cast<int64_t>(n_tiles) * 128 * 128, 128, 128};
  if (int rc = make_tmap_4d(&tm, acc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 128, static_cast<int64_t>(n_tiles) * 128, 1, 1, st, 32, 128))
    return rc;
  if (int rc = make_tmap_4d(&tm_ld, src_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 128, static_cast<int64_t>(n_tiles) * 128, 1, 1, st, 64, 128))
    return rc;
  const int bytes = 8 * 16384 + 64;
  VT_CHECK_CUDA(cudaFuncSetAttribute(tma_mixed_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  tma_mixed_rate_kernel<<<blocks, 128, bytes, static_cast<cudaStream_t>(stream)>>>(tm, tm_ld, iters, n_tiles, load_mode,
                                                                                  cycles_out, loads_out);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vt_tma_reduce_rate(float* acc, int n_tiles, int iters, int depth, int spread, int blocks,
                                  long long* cycles_out, void* stream) {
  VT_REQUIRE(acc != nullptr && cycles_out != nullptr, VT_ERR_NULL, "vt_tma_reduce_rate: NULL argument");
  VT_REQUIRE(n_tiles > 0 && iters > 0 && ((depth >= 1 && depth <= 4) || (depth >= 17 && depth <= 20)) && blocks > 0,
             VT_ERR_SHAPE, "bad argument");
  if (int rc = lib_init()) return rc;
  CUtensorMap tm;
  const int64_t st[3] = {static_cast<int64_t>(n_tiles) * 128 * 128, 128, 128};
  if (int rc = make_tmap_4d(&tm, acc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 128, static_cast<int64_t>(n_tiles) * 128, 1, 1, st, 32, 128))
    return rc;
  VT_CHECK_CUDA(cudaFuncSetAttribute(tma_reduce_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384));
  // depth >= 16 selects the 1-D (non-tensor) bulk reduction with depth - 16 groups in flight
  float* linear = depth >= 16 ? acc : nullptr;
  if (depth >= 16) depth -= 16;
  tma_reduce_rate_kernel<<<blocks, 128, 4 * 16384, static_cast<cudaStream_t>(stream)>>>(tm, linear, iters, depth, spread,
                                                                                       n_tiles, cycles_out);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#endif  // VT_EXPERIMENTS
