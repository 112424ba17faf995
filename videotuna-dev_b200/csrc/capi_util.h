// capi_util.h — error plumbing shared by the extern "C" entry points.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <mutex>
#include <set>
#include <utility>

#include "../../include/b200vt.h"

namespace vt {

// Thread-local last-error message (the op may be called from PyTorch's autograd worker threads).
char* last_error_buf();
constexpr int kErrBuf = 512;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), kErrBuf, fmt, ap);
  va_end(ap);
  return code;
}
inline int cuda_fail(cudaError_t e, const char* what) {
  return fail(VT_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define VT_CHECK_CUDA(expr)                                 \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return ::vt::cuda_fail(_e, #expr); \
  } while (0)

#define VT_REQUIRE(cond, code, ...)                     \
  do {                                                  \
    if (!(cond)) return ::vt::fail(code, __VA_ARGS__);  \
  } while (0)

// True the first time `site` is seen on the calling thread's current device. Function attributes
// (cudaFuncAttributeMaxDynamicSharedMemorySize, cluster opt-in) and __device__ symbols are PER DEVICE: a process that
// drives several GPUs must configure every kernel once on each of them.
inline bool first_on_device(const void* site) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> seen;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  return seen.insert({site, dev}).second;
}

inline unsigned cdiv(long long a, long long b) { return static_cast<unsigned>((a + b - 1) / b); }

}  // namespace vt
