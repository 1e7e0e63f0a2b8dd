// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/commu_b200.h"

namespace cb_host {

char* error_buffer();  // thread-local, 512 bytes (defined in api.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CB_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return cb_host::fail(COMMU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                      \
  } while (0)

#define CB_REQUIRE(cond, ...)                                                                \
  do {                                                                                       \
    if (!(cond)) return cb_host::fail(COMMU_ERR_INVALID, __VA_ARGS__);                       \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Optional per-kernel-class device timers (used by bench.py for the roofline line): when a class
// is armed, launches of that class are bracketed with cudaEvents on the launching stream.
enum ProfClass { PROF_GEMM = 0, PROF_ATTN_FWD = 1, PROF_ATTN_BWD = 2, PROF_DECODE_ATTN = 3, PROF_NUM = 4 };
struct ProfScope {
  ProfScope(int cls, cudaStream_t s);
  ~ProfScope();
  int cls_;
  cudaStream_t s_;
  int slot_;
};
void count_launch(int n = 1);

}  // namespace cb_host
