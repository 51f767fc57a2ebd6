// Shared helpers for libpclseg (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>

#include "../../include/pclseg.h"

namespace pcls {

void set_error(const char* fmt, ...);

#define PCLS_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      pcls::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                      __LINE__);                                                           \
      return PCLS_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define PCLS_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      pcls::set_error(__VA_ARGS__);        \
      return PCLS_ERR_INVALID;             \
    }                                      \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return PCLS_ERR_CUDA;
  }
  return PCLS_OK;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pcls
