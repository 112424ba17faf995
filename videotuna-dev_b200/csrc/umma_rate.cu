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
__global__ void __launch_bounds__(256, 1)
umma_rate_kernel(int mode, int n, int iters, int noise_warps, long long* cycles_out) {
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
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n, mode == 2 ? 1 : 0, mode == 0 ? 0 : 1);
    const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t d = tmem + (it & 1) * 128;
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

extern "C" int vt_umma_rate(int mode, int n, int iters, int noise_warps, int blocks, long long* cycles_out, void* stream) {
  VT_REQUIRE(cycles_out != nullptr, VT_ERR_NULL, "vt_umma_rate: NULL output");
  VT_REQUIRE(mode >= 0 && mode <= 3 && n >= 16 && n <= 128 && n % 16 == 0 && iters > 0 && blocks > 0 && noise_warps >= 0 &&
                 noise_warps <= 7, VT_ERR_SHAPE, "vt_umma_rate: bad argument");
  if (int rc = lib_init()) return rc;
  const int bytes = 6 * 128 * 128 + 64;
  VT_CHECK_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  umma_rate_kernel<<<blocks, 256, bytes, static_cast<cudaStream_t>(stream)>>>(mode, n, iters, noise_warps, cycles_out);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}
