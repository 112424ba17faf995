// capi.cu — library state, TMA tensor-map construction and the attention entry points of the C ABI.
#include <cstring>
#include <mutex>
#include <vector>

#include "attn_common.h"
#include "capi_util.h"

namespace vt {

char* last_error_buf() {
  static thread_local char buf[kErrBuf] = {0};
  return buf;
}

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct LibState {
  std::mutex mu;
  bool ready = false;
  int init_rc = 0;
  EncodeTiledFn encode = nullptr;
  unsigned int* dbg_host = nullptr;  // pinned, mapped
  unsigned int* dbg_dev = nullptr;
};
LibState g_state;

// The tensor-map encoder is a DRIVER entry point and needs a context current on the calling thread. PyTorch's autograd
// worker threads select their device lazily (no cudaSetDevice until a runtime call needs it), so a backward entry point
// reached without the dispatcher's device guard — the eager path of ops.py — can arrive with no context bound
// (CUDA_ERROR_INVALID_CONTEXT from cuTensorMapEncodeTiled). One runtime call per thread binds the primary context of the
// thread's current device.
int bind_context_once() {
  thread_local int bound_dev = -1;  // the device whose primary context this thread last bound
  int dev = 0;
  VT_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev != bound_dev) {
    VT_CHECK_CUDA(cudaFree(nullptr));
    bound_dev = dev;
  }
  return 0;
}

// Per-device part of the initialisation: the compute-capability check and the watchdog pointer, which lives in a
// __device__ symbol of every attention translation unit (one copy per device).
int ensure_device_init() {
  static char site;
  if (!first_on_device(&site)) return 0;
  int dev = 0;
  VT_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  VT_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  VT_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) return fail(VT_ERR_UNSUPPORTED, "device %d is sm_%d%d; libb200vt needs sm_100", dev, major, minor);
  VT_CHECK_CUDA(attn_fwd_set_debug_ptr(g_state.dbg_dev));
  VT_CHECK_CUDA(attn_bwd_set_debug_ptr(g_state.dbg_dev));
  VT_CHECK_CUDA(attn_fwd_alt_set_debug_ptr(g_state.dbg_dev));
  return 0;
}

int ensure_init() {
  if (int rc = bind_context_once()) return rc;
  std::lock_guard<std::mutex> lock(g_state.mu);
  if (g_state.ready) return g_state.init_rc != 0 ? g_state.init_rc : ensure_device_init();
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  VT_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr)
    return fail(VT_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  g_state.encode = reinterpret_cast<EncodeTiledFn>(fn);
  VT_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_state.dbg_host), 64, cudaHostAllocMapped | cudaHostAllocPortable));
  std::memset(g_state.dbg_host, 0, 64);
  VT_CHECK_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_state.dbg_dev), g_state.dbg_host, 0));
  g_state.ready = true;
  g_state.init_rc = 0;
  return ensure_device_init();
}

}  // namespace

// 4-D tensor map over a (B, L, H, D)-logical tensor with element strides (sb, sl, sh, 1):
// TMA dims innermost-first = (D, L, H, B); box = (box_d, box_rows, 1, 1); 128-byte swizzle; OOB reads as zero.
int make_tmap_4d(CUtensorMap* out, const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t D, int64_t L,
                 int64_t H, int64_t B, const int64_t* strides /* b,l,h */, int box_d, int box_rows) {
  VT_REQUIRE(aligned16(ptr), VT_ERR_ALIGN, "tensor base %p is not 16-byte aligned", ptr);
  const int64_t sb = strides[0], sl = strides[1], sh = strides[2];
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(sl * elem_bytes), static_cast<cuuint64_t>(sh * elem_bytes),
                        static_cast<cuuint64_t>(sb * elem_bytes)};
  // size-1 dims may come with arbitrary (even zero) strides from PyTorch; give them a harmless legal value
  if (H == 1) gstr[1] = static_cast<cuuint64_t>(D * elem_bytes);
  if (B == 1) gstr[2] = static_cast<cuuint64_t>(D * elem_bytes);
  if (L == 1) gstr[0] = static_cast<cuuint64_t>(D * elem_bytes);
  for (int i = 0; i < 3; ++i)
    VT_REQUIRE(gstr[i] % 16 == 0 && gstr[i] > 0, VT_ERR_ALIGN, "stride %d (%llu bytes) must be a positive multiple of 16",
               i, static_cast<unsigned long long>(gstr[i]));
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_d), static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_state.encode(out, dt, 4, const_cast<void*>(ptr), dims, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VT_REQUIRE(r == CUDA_SUCCESS, VT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (dims %lld,%lld,%lld,%lld)",
             static_cast<int>(r), (long long)D, (long long)L, (long long)H, (long long)B);
  return 0;
}

int lib_init() { return ensure_init(); }

static long long* g_trace = nullptr;
long long* debug_trace_ptr() { return g_trace; }
void debug_set_trace(long long* p) { g_trace = p; }

// ---- optional per-kernel event timing (vt_profile_*) ---------------------------------------------------------
namespace {
struct ProfSpan {
  int id;
  cudaEvent_t e0, e1;
};
struct Profiler {
  std::mutex mu;
  bool on = false;
  std::vector<ProfSpan> pending;
  std::vector<cudaEvent_t> pool;
  double total_ms[VT_K_COUNT] = {0};
  int64_t launches[VT_K_COUNT] = {0};
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
Profiler g_prof;
}  // namespace

// RAII span: records an event before and after whatever is launched on `st` during its lifetime.
struct ProfScope {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t st;
  int id;
  ProfScope(int kernel_id, cudaStream_t stream) : st(stream), id(kernel_id) {
    std::lock_guard<std::mutex> lock(g_prof.mu);
    if (!g_prof.on) return;
    e0 = g_prof.get();
    e1 = g_prof.get();
    if (e0 != nullptr) cudaEventRecord(e0, st);
  }
  ~ProfScope() {
    if (e0 == nullptr || e1 == nullptr) return;
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lock(g_prof.mu);
    g_prof.pending.push_back({id, e0, e1});
  }
};

}  // namespace vt

using namespace vt;

extern "C" {

int vt_version(void) { return 100; }

int vt_last_error(char* buf, size_t n) {
  const char* s = last_error_buf();
  size_t len = std::strlen(s);
  if (buf != nullptr && n > 0) {
    size_t c = len < n - 1 ? len : n - 1;
    std::memcpy(buf, s, c);
    buf[c] = 0;
  }
  return static_cast<int>(len);
}

int vt_init(int device) {
  VT_CHECK_CUDA(cudaSetDevice(device));
  return ensure_init();
}

int vt_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  for (auto& sp : g_prof.pending) {
    g_prof.pool.push_back(sp.e0);
    g_prof.pool.push_back(sp.e1);
  }
  g_prof.pending.clear();
  for (int i = 0; i < VT_K_COUNT; ++i) {
    g_prof.total_ms[i] = 0;
    g_prof.launches[i] = 0;
  }
  g_prof.on = on != 0;
  return 0;
}

int vt_profile_read(int kernel_id, double* total_ms, int64_t* launches) {
  VT_REQUIRE(kernel_id >= 0 && kernel_id < VT_K_COUNT, VT_ERR_SHAPE, "kernel id %d out of range", kernel_id);
  VT_REQUIRE(total_ms != nullptr && launches != nullptr, VT_ERR_NULL, "vt_profile_read: NULL output");
  std::lock_guard<std::mutex> lock(g_prof.mu);
  for (auto& sp : g_prof.pending) {
    VT_CHECK_CUDA(cudaEventSynchronize(sp.e1));
    float ms = 0.f;
    VT_CHECK_CUDA(cudaEventElapsedTime(&ms, sp.e0, sp.e1));
    g_prof.total_ms[sp.id] += ms;
    g_prof.launches[sp.id] += 1;
    g_prof.pool.push_back(sp.e0);
    g_prof.pool.push_back(sp.e1);
  }
  g_prof.pending.clear();
  *total_ms = g_prof.total_ms[kernel_id];
  *launches = g_prof.launches[kernel_id];
  return 0;
}

int vt_memcpy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t rows,
                      int to_device, void* stream) {
  VT_REQUIRE(dst != nullptr && src != nullptr, VT_ERR_NULL, "vt_memcpy2d_async: NULL pointer");
  VT_REQUIRE(width_bytes <= dpitch && width_bytes <= spitch, VT_ERR_SHAPE, "vt_memcpy2d_async: width exceeds pitch");
  VT_CHECK_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, rows,
                                  to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                  static_cast<cudaStream_t>(stream)));
  return 0;
}

int vt_debug_set_trace(long long* device_buf) {
  debug_set_trace(device_buf);
  return 0;
}

int vt_debug_watchdog(uint32_t out[4]) {
  VT_REQUIRE(out != nullptr, VT_ERR_NULL, "out is NULL");
  for (int i = 0; i < 4; ++i) out[i] = g_state.dbg_host ? g_state.dbg_host[i] : 0u;
  return 0;
}

static int check_attn_common(int B, int H, int Lq, int Lk, int D, int num_segments, const int32_t* cu_q,
                             const int32_t* cu_k, int max_q, int max_k) {
  VT_REQUIRE(D == 64 || D == 128, VT_ERR_DTYPE, "head dim %d unsupported (64 or 128)", D);
  VT_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk >= 0, VT_ERR_SHAPE, "bad shape B=%d H=%d Lq=%d Lk=%d", B, H, Lq, Lk);
  VT_REQUIRE(H <= 65535, VT_ERR_SHAPE, "H=%d exceeds grid.y", H);
  if (num_segments > 0) {
    VT_REQUIRE(B == 1, VT_ERR_SHAPE, "varlen mode needs packed tensors with B == 1 (got %d)", B);
    VT_REQUIRE(cu_q != nullptr && cu_k != nullptr, VT_ERR_NULL, "varlen mode needs cu_seqlens_q and cu_seqlens_k");
    VT_REQUIRE(max_q > 0 && max_k >= 0, VT_ERR_SHAPE, "varlen mode needs max_seqlen_q/k");
    VT_REQUIRE(num_segments <= 65535, VT_ERR_SHAPE, "too many segments");
  } else {
    VT_REQUIRE(B <= 65535, VT_ERR_SHAPE, "B=%d exceeds grid.z", B);
  }
  return 0;
}

struct FwdScatter {
  void* const* bases;
  int n, rows_per_peer;
  const int64_t* strides;  // (row, head) element strides at the destination
};

static int attn_fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, const int64_t* q_strides,
                         const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides, int B, int H, int Lq,
                         int Lk, int D, const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int num_segments,
                         int max_seqlen_q, int max_seqlen_k, const int32_t* seqlens_k, float softmax_scale,
                         const FwdScatter* sc, void* stream) {
  VT_REQUIRE(q && k && v && o && lse && q_strides && k_strides && v_strides && o_strides, VT_ERR_NULL,
             "vt_attn_fwd: NULL argument");
  if (int rc = check_attn_common(B, H, Lq, Lk, D, num_segments, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k)) return rc;
  if (int rc = ensure_init()) return rc;
  VT_REQUIRE(aligned16(o), VT_ERR_ALIGN, "o is not 16-byte aligned");
  VT_REQUIRE(o_strides[0] % 8 == 0 && o_strides[1] % 8 == 0 && o_strides[2] % 8 == 0, VT_ERR_ALIGN,
             "o strides must be multiples of 8 elements");
  VT_REQUIRE(softmax_scale > 0.f, VT_ERR_SHAPE, "softmax_scale must be positive");

  CUtensorMap tm_q, tm_k, tm_v;
  const int64_t Lk_map = Lk > 0 ? Lk : 1;
  if (int rc = make_tmap_4d(&tm_q, q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, D, Lq, H, B, q_strides, 64, 128)) return rc;
  if (int rc = make_tmap_4d(&tm_k, k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, D, Lk_map, H, B, k_strides, 64, 128)) return rc;
  if (int rc = make_tmap_4d(&tm_v, v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, D, Lk_map, H, B, v_strides, 64, 128)) return rc;

  AttnFwdParams p{};
  p.trace = debug_trace_ptr();
  p.seq.cu_q = num_segments > 0 ? cu_seqlens_q : nullptr;
  p.seq.cu_k = num_segments > 0 ? cu_seqlens_k : nullptr;
  p.seq.seqlens_k = num_segments > 0 ? nullptr : seqlens_k;
  p.seq.Lq = num_segments > 0 ? max_seqlen_q : Lq;
  p.seq.Lk = num_segments > 0 ? max_seqlen_k : Lk;
  p.seq.H = H;
  p.seq.nprob = num_segments > 0 ? num_segments : B;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.o_sb = o_strides[0];
  p.o_sl = o_strides[1];
  p.o_sh = o_strides[2];
  p.lse_sb = static_cast<int64_t>(H) * Lq;
  p.lse_sh = Lq;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  if (sc != nullptr) {
    VT_REQUIRE(D == 128 && num_segments == 0 && B == 1, VT_ERR_UNSUPPORTED,
               "the fused exchange epilogue needs head dim 128, fixed mode and batch 1");
    VT_REQUIRE(sc->bases && sc->strides && sc->n >= 1 && sc->n <= 8 && sc->rows_per_peer >= 1 &&
                   static_cast<int64_t>(sc->n) * sc->rows_per_peer <= Lq, VT_ERR_SHAPE, "bad peer layout");
    VT_REQUIRE(sc->strides[0] % 8 == 0 && sc->strides[1] % 8 == 0, VT_ERR_ALIGN, "peer strides must be multiples of 8");
    for (int i = 0; i < sc->n; ++i) {
      VT_REQUIRE(sc->bases[i] != nullptr && aligned16(sc->bases[i]), VT_ERR_ALIGN, "peer base %d is NULL or misaligned", i);
      p.sc_base[i] = static_cast<__nv_bfloat16*>(sc->bases[i]);
    }
    p.sc_n = sc->n;
    p.sc_rpr = sc->rows_per_peer;
    p.sc_sl = sc->strides[0];
    p.sc_sh = sc->strides[1];
  }
  {
    ProfScope span(VT_K_ATTN_FWD, static_cast<cudaStream_t>(stream));
    VT_CHECK_CUDA(launch_attn_fwd(D, tm_q, tm_k, tm_v, p, 0, static_cast<cudaStream_t>(stream)));
  }
  return 0;
}

int vt_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const int64_t* q_strides,
                const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides, int B, int H, int Lq,
                int Lk, int D, const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int num_segments,
                int max_seqlen_q, int max_seqlen_k, const int32_t* seqlens_k, float softmax_scale, void* stream) {
  return attn_fwd_impl(q, k, v, o, lse, q_strides, k_strides, v_strides, o_strides, B, H, Lq, Lk, D, cu_seqlens_q,
                       cu_seqlens_k, num_segments, max_seqlen_q, max_seqlen_k, seqlens_k, softmax_scale, nullptr, stream);
}

int vt_attn_fwd_scatter(const void* q, const void* k, const void* v, void* o, float* lse, const int64_t* q_strides,
                        const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides, int H, int Lq, int Lk,
                        int D, const int32_t* seqlens_k, float softmax_scale, void* const* peer_bases, int n_peers,
                        int rows_per_peer, const int64_t* peer_strides, void* stream) {
  const FwdScatter sc{peer_bases, n_peers, rows_per_peer, peer_strides};
  return attn_fwd_impl(q, k, v, o, lse, q_strides, k_strides, v_strides, o_strides, 1, H, Lq, Lk, D, nullptr, nullptr, 0,
                       0, 0, seqlens_k, softmax_scale, &sc, stream);
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

int64_t vt_attn_bwd_workspace_bytes(int B, int H, int Lq, int D) {
  if (B <= 0 || H <= 0 || Lq <= 0 || D <= 0) return 0;
  const int64_t rows = static_cast<int64_t>(B) * Lq * H;
  return align_up(rows * D * 4, 256) + align_up(rows * 4, 256);
}

int vt_attn_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o, const float* lse,
                void* dq, void* dk, void* dv, const int64_t* do_strides, const int64_t* q_strides,
                const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides,
                const int64_t* dq_strides, const int64_t* dk_strides, const int64_t* dv_strides, int B, int H, int Lq,
                int Lk, int D, const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int num_segments,
                int max_seqlen_q, int max_seqlen_k, const int32_t* seqlens_k, float softmax_scale, void* workspace,
                int64_t workspace_bytes, void* stream) {
  VT_REQUIRE(dout && q && k && v && o && lse && dq && dk && dv && workspace, VT_ERR_NULL, "vt_attn_bwd: NULL argument");
  VT_REQUIRE(do_strides && q_strides && k_strides && v_strides && o_strides && dq_strides && dk_strides && dv_strides,
             VT_ERR_NULL, "vt_attn_bwd: NULL stride array");
  if (int rc = check_attn_common(B, H, Lq, Lk, D, num_segments, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k)) return rc;
  if (int rc = ensure_init()) return rc;
  VT_REQUIRE(workspace_bytes >= vt_attn_bwd_workspace_bytes(B, H, Lq, D), VT_ERR_SHAPE, "workspace too small");
  VT_REQUIRE(aligned16(workspace) && aligned16(dq) && aligned16(dk) && aligned16(dv) && aligned16(o) && aligned16(dout),
             VT_ERR_ALIGN, "buffers must be 16-byte aligned");
  for (const int64_t* s : {dq_strides, dk_strides, dv_strides, o_strides, do_strides})
    VT_REQUIRE(s[0] % 8 == 0 && s[1] % 8 == 0 && s[2] % 8 == 0, VT_ERR_ALIGN, "strides must be multiples of 8 elements");
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t rows = static_cast<int64_t>(B) * Lq * H;
  float* dq_acc = static_cast<float*>(workspace);
  float* delta = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up(rows * D * 4, 256));

  if (Lk == 0) {  // no keys at all: every gradient is zero
    VT_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(rows) * D * 4, st));
    VT_CHECK_CUDA(launch_attn_bwd_dq_convert(dq_acc, dq, dq_strides, B, Lq, H, D, 0.f, st));
    return 0;
  }
  CUtensorMap tm_q, tm_k, tm_v, tm_do, tm_dq;
  {  // fp32 dQ accumulator (B, Lq, H, D) contiguous; reduced into through 128-row x 32-column swizzled boxes
    const int64_t acc_strides[3] = {static_cast<int64_t>(Lq) * H * D, static_cast<int64_t>(H) * D, D};
    if (int rc = make_tmap_4d(&tm_dq, dq_acc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, D, Lq, H, B, acc_strides, 32, 128)) return rc;
  }
  if (int rc = make_tmap_4d(&tm_q, q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, D, Lq, H, B, q_strides, 64, 128)) return rc;
  if (int rc = make_tmap_4d(&tm_k, k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, D, Lk, H, B, k_strides, 64, 128)) return rc;
  if (int rc = make_tmap_4d(&tm_v, v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, D, Lk, H, B, v_strides, 64, 128)) return rc;
  if (int rc = make_tmap_4d(&tm_do, dout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, D, Lq, H, B, do_strides, 64, 128)) return rc;

  // one key tile per (sample, head) in fixed mode: dQ is written directly by the kernel (AttnBwdParams::dq_direct)
  const bool dq_direct = num_segments <= 0 && Lk <= 128;
  if (!dq_direct) VT_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(rows) * D * 4, st));
  {
    ProfScope span(VT_K_ATTN_BWD_DELTA, st);
    VT_CHECK_CUDA(launch_attn_bwd_delta(D, dout, o, delta, do_strides, o_strides, B, Lq, H, st));
  }

  AttnBwdParams p{};
  p.trace = debug_trace_ptr();
  p.seq.cu_q = num_segments > 0 ? cu_seqlens_q : nullptr;
  p.seq.cu_k = num_segments > 0 ? cu_seqlens_k : nullptr;
  p.seq.seqlens_k = num_segments > 0 ? nullptr : seqlens_k;
  p.seq.Lq = num_segments > 0 ? max_seqlen_q : Lq;
  p.seq.Lk = num_segments > 0 ? max_seqlen_k : Lk;
  p.seq.H = H;
  p.seq.nprob = num_segments > 0 ? num_segments : B;
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.q_sb = q_strides[0]; p.q_sl = q_strides[1]; p.q_sh = q_strides[2];
  p.do_sb = do_strides[0]; p.do_sl = do_strides[1]; p.do_sh = do_strides[2];
  p.lse = lse;
  p.delta = delta;
  p.lse_sb = static_cast<int64_t>(H) * Lq;
  p.lse_sh = Lq;
  p.dk = static_cast<__nv_bfloat16*>(dk);
  p.dv = static_cast<__nv_bfloat16*>(dv);
  p.dk_sb = dk_strides[0]; p.dk_sl = dk_strides[1]; p.dk_sh = dk_strides[2];
  p.dv_sb = dv_strides[0]; p.dv_sl = dv_strides[1]; p.dv_sh = dv_strides[2];
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.dq_direct = dq_direct ? static_cast<__nv_bfloat16*>(dq) : nullptr;
  p.dq_sb = dq_strides[0]; p.dq_sl = dq_strides[1]; p.dq_sh = dq_strides[2];
  {
    ProfScope span(VT_K_ATTN_BWD, st);
    VT_CHECK_CUDA(launch_attn_bwd(D, tm_q, tm_k, tm_v, tm_do, tm_dq, dq_acc, Lq, p, st));
  }
  if (!dq_direct) {
    ProfScope span(VT_K_ATTN_BWD_DQ, st);
    VT_CHECK_CUDA(launch_attn_bwd_dq_convert(dq_acc, dq, dq_strides, B, Lq, H, D, softmax_scale, st));
  }
  return 0;
}

}  // extern "C"
