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

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// Every kernel of the forward is launched with the programmatic-stream-serialization attribute: its CTAs may become
// resident while the kernel in front of it is still draining, run their prologue (barrier init, TMEM allocation, weight
// loads - nothing that depends on the predecessor) and then block in pdl_wait() until the predecessor has completed and
// its writes are visible.  Rules the kernels follow: (1) every thread executes pdl_wait() before its first access -
// read OR write, the arena is reused by liveness - to memory another kernel of the forward touches; (2) pdl_trigger()
// at the start: the dependent grid is released once every CTA of this grid has started, i.e. when nothing of this grid
// is left to schedule.  Captured into the CUDA graph the attribute becomes a programmatic dependency edge.
// Measured (B200, SqueezeSegV2 64x2048, forward ms without / with the early trigger): batch 1 0.351 / 0.331, batch 2
// 0.424 / 0.396, batch 4 0.601 / 0.570, batch 8 0.934 / 0.919, batch 16 1.647 / 1.680, batch 32 3.04 / 3.16 - the
// prologues hide behind the previous layer (~30 us per forward), but resident dependents cost the running grid ~3.5 % of
// its time (forcing one CTA per SM with a padded shared-memory request changes nothing; without an explicit trigger PDL
// is neutral).  So the trigger is a per-pass decision: Net::run_pass sets pdl_early_now for passes of at most
// pdl_early_px pixels and the launchers hand it to the kernels as a parameter.
extern int pdl_mode;   // 1 = on (default), 0 = plain stream order (PCLS_PDL=0 in the environment, A/B measurements)
extern long long pdl_early_px;                 // PCLS_PDL_EARLY_PX, default 2^20 pixels (8 frames of 64x2048)
extern thread_local int pdl_early_now;

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef PCLS_PDL_TRIGGER
#define PCLS_PDL_TRIGGER 1   // 0: no explicit trigger, the dependent grid is released when this grid has completed
#endif
__device__ __forceinline__ void pdl_trigger(int early) {
  if (PCLS_PDL_TRIGGER && early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_mode ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

}  // namespace pcls
