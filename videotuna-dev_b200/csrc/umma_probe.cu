// Built only with -DVT_EXPERIMENTS (tools/build_variant.sh): not part of the product library.
#ifdef VT_EXPERIMENTS
// umma_probe.cu — single-tile self-test of the TMA + tcgen05 building blocks the attention kernels rely on.
// D[128 x n] = A[128 x 128] * B, bf16 in / fp32 out, with selectable operand sources:
//   a_mode 0: A (M x K, K contiguous) from smem, K-major       (Q, K, V, dO as "row" operands)
//   a_mode 1: A^T given as (K x M, M contiguous) from smem, MN-major   (dS as the A operand of dQ = dS K)
//   a_mode 2: A (M x K) staged into TMEM as packed bf16 (tcgen05.st) (P as the A operand of O += P V)
//   b_mode 0: B given as (N x K, K contiguous), K-major        (K in S = Q K^T)
//   b_mode 1: B given as (K x N, N contiguous), MN-major       (V in O = P V)
// Descriptor offsets (leading / stride byte offsets, per-k-step start advance) are runtime arguments so a test can
// confirm the encodings the kernels hard-code. Test infrastructure only; nothing on the product path calls it.
#include "capi_util.h"
#include "sm100_ptx.cuh"

namespace vt {
int make_tmap_4d(CUtensorMap* out, const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t D, int64_t L,
                 int64_t H, int64_t B, const int64_t* strides, int box_d, int box_rows);
int lib_init();

namespace {

struct ProbeArgs {
  const __nv_bfloat16* a_gmem;  // for a_mode 2
  float* d;
  int a_mode, b_mode, n;
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                  const ProbeArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int CHUNK = 128 * 128;
  uint8_t* a_s = smem;               // 2 boxes of [128 rows][64 elems]
  uint8_t* b_s = smem + 2 * CHUNK;   // 2 boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * CHUNK);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (threadIdx.x == 0) {
    uint32_t bytes = 2 * CHUNK;
    if (args.a_mode != 2) bytes += 2 * CHUNK;
    mbar_arrive_expect_tx(bars + 0, bytes);
    if (args.a_mode != 2) {
      tma_load_4d(a_s, &tm_a, bars + 0, 0, 0, 0, 0);
      tma_load_4d(a_s + CHUNK, &tm_a, bars + 0, 64, 0, 0, 0);
    }
    tma_load_4d(b_s, &tm_b, bars + 0, 0, 0, 0, 0);
    tma_load_4d(b_s + CHUNK, &tm_b, bars + 0, 64, 0, 0, 0);
  }
  if (args.a_mode == 2) {
    // stage A into TMEM columns [128, 192): thread r owns row r; column c holds (A[r][2c], A[r][2c+1])
    const __nv_bfloat16* arow = args.a_gmem + static_cast<size_t>(threadIdx.x) * 128;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16) + 128;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t w[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) w[c] = *reinterpret_cast<const uint32_t*>(arow + 2 * (c0 + c));
      tmem_st_x16(taddr + c0, w);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    mbar_wait(bars + 0, 0, 0x900);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, args.n, args.a_mode == 1 ? 1 : 0, args.b_mode == 1 ? 1 : 0);
    const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
    for (int kk = 0; kk < 8; ++kk) {  // K = 128 in steps of 16
      // K-major operands: k-step kk lives in box kk/4 at byte offset (kk%4)*kstep; MN-major: linear in kk
      const uint32_t a_off = args.a_mode == 0 ? (kk >> 2) * CHUNK + (kk & 3) * args.a_kstep : kk * args.a_kstep;
      const uint32_t b_off = args.b_mode == 0 ? (kk >> 2) * CHUNK + (kk & 3) * args.b_kstep : kk * args.b_kstep;
      const uint64_t bd = umma_desc_sw128(b_addr + b_off, args.b_lbo, args.b_sbo);
      if (args.a_mode == 2) {
        umma_ts(tmem, tmem + 128 + kk * 8, bd, idesc, kk != 0);
      } else {
        const uint64_t ad = umma_desc_sw128(a_addr + a_off, args.a_lbo, args.a_sbo);
        umma_ss(tmem, ad, bd, idesc, kk != 0);
      }
    }
    tc_commit(bars + 1);
  }
  mbar_wait(bars + 1, 0, 0x901);
  tc_fence_after();
  {
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    float* drow = args.d + static_cast<size_t>(threadIdx.x) * args.n;
    for (int c0 = 0; c0 < args.n; c0 += 16) {
      uint32_t r[16];
      tmem_ld_x16(taddr + c0, r);
      tc_wait_ld();
#pragma unroll
      for (int c = 0; c < 16; ++c) drow[c0 + c] = __uint_as_float(r[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace
}  // namespace vt

using namespace vt;

extern "C" int vt_umma_probe(const void* a, const void* b, float* d, int a_mode, int b_mode, int n, uint32_t a_lbo,
                             uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                             void* stream) {
  VT_REQUIRE(a && b && d, VT_ERR_NULL, "vt_umma_probe: NULL argument");
  VT_REQUIRE(n >= 16 && n <= 128 && n % 16 == 0, VT_ERR_SHAPE, "n must be a multiple of 16 in [16,128]");
  VT_REQUIRE(a_mode >= 0 && a_mode <= 2 && b_mode >= 0 && b_mode <= 1, VT_ERR_SHAPE, "bad mode");
  if (int rc = lib_init()) return rc;
  static bool once = false;
  if (!once) {
    unsigned int* dummy = nullptr;
    (void)dummy;
    VT_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 128 * 128 + 2048));
    once = true;
  }
  // Both operands are 128 x 128 bf16 row-major buffers; what the rows mean depends on the mode (see header comment).
  // K-major operand: rows = M (or N) index, cols = K. MN-major: rows = K, cols = M (or N).
  CUtensorMap tm_a, tm_b;
  const int64_t st[3] = {128 * 128, 128, 128 * 128};
  if (int rc = make_tmap_4d(&tm_a, a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 128, 128, 1, 1, st, 64, 128)) return rc;
  const int64_t b_cols = 128, b_rows = (b_mode == 0) ? n : 128;
  (void)b_cols;
  if (int rc = make_tmap_4d(&tm_b, b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 128, b_rows, 1, 1, st, 64, 128)) return rc;
  ProbeArgs args{static_cast<const __nv_bfloat16*>(a), d, a_mode, b_mode, n, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep};
  umma_probe_kernel<<<1, 128, 4 * 128 * 128 + 2048, static_cast<cudaStream_t>(stream)>>>(tm_a, tm_b, args);
  VT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#endif  // VT_EXPERIMENTS
