// Confusion-matrix accumulation for mIoU + the one multi-GPU exchange step (sm_100a).
//
// Replaces tf.keras.metrics.MeanIoU.update_state (pcl_segmentation/eval.py:41,48;
// nets/SegmentationNetwork.py:52,113,129): cm[label, pred] += 1, rows = label, cols = prediction.
// HBM-bound: 8 B read per pixel.  Per-warp privatised shared-memory histograms (NC*NC <= 1024 bins of u32)
// with __match_any_sync aggregation, so the heavily skewed (None, None) bin costs one shared atomic per
// warp-instruction instead of 32 serialised ones; int64 global atomics once per block at the end.
#include "common.cuh"
#include <dlfcn.h>

namespace pcls {

constexpr int CM_WARPS = 8;
constexpr int CM_MAX_NC = 32;

__global__ void __launch_bounds__(CM_WARPS * 32)
confusion_kernel(const int32_t* __restrict__ label, const int32_t* __restrict__ pred, int64_t n, int nc,
                 unsigned long long* __restrict__ cm, unsigned long long* __restrict__ dropped) {
  extern __shared__ unsigned int hist[];  // [CM_WARPS][nbins]
  const int nbins = nc * nc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < CM_WARPS * nbins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned int* h = hist + warp * nbins;
  unsigned int my_dropped = 0;

  auto count = [&](int l, int p, bool active) {
    const bool valid = active && (unsigned)l < (unsigned)nc && (unsigned)p < (unsigned)nc;
    if (active && !valid) ++my_dropped;
    const int bin = valid ? l * nc + p : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (valid && lane == (__ffs(peers) - 1)) atomicAdd(h + bin, (unsigned)__popc(peers));
  };

  // 128-bit loads: 4 pixels per lane per iteration; whole-warp trip count so the warp collectives stay converged
  const int64_t n4 = n / 4;
  const int64_t gw = (int64_t)blockIdx.x * CM_WARPS + warp, nw = (int64_t)gridDim.x * CM_WARPS;
  const int4* l4 = reinterpret_cast<const int4*>(label);
  const int4* p4 = reinterpret_cast<const int4*>(pred);
  for (int64_t base = gw * 32; base < n4; base += nw * 32) {
    const int64_t i = base + lane;
    const bool act = i < n4;
    int4 a = make_int4(0, 0, 0, 0), b = make_int4(0, 0, 0, 0);
    if (act) { a = __ldg(l4 + i); b = __ldg(p4 + i); }
    count(a.x, b.x, act); count(a.y, b.y, act); count(a.z, b.z, act); count(a.w, b.w, act);
  }
  if (gw == 0) {  // tail (< 4 elements)
    const int64_t i = n4 * 4 + lane;
    const bool act = i < n;
    count(act ? label[i] : 0, act ? pred[i] : 0, act);
  }
  __syncthreads();
  for (int bin = threadIdx.x; bin < nbins; bin += blockDim.x) {
    unsigned long long s = 0;
#pragma unroll
    for (int w = 0; w < CM_WARPS; ++w) s += hist[w * nbins + bin];
    if (s) atomicAdd(cm + bin, s);
  }
  if (dropped != nullptr && my_dropped) atomicAdd(dropped, (unsigned long long)my_dropped);
}

int launch_confusion(const int32_t* label, const int32_t* pred, int64_t n, int nc, int64_t* cm, int64_t* dropped,
                     cudaStream_t s) {
  if (n == 0) return PCLS_OK;
  const size_t smem = (size_t)CM_WARPS * nc * nc * sizeof(unsigned int);
  int64_t blocks = ceil_div(ceil_div(n, 4), CM_WARPS * 32 * 8);
  int64_t cap = (int64_t)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  confusion_kernel<<<(int)blocks, CM_WARPS * 32, smem, s>>>(label, pred, n, nc,
                                                           reinterpret_cast<unsigned long long*>(cm),
                                                           reinterpret_cast<unsigned long long*>(dropped));
  return check_launch("confusion_kernel");
}

// ---- NCCL through dlopen: no link-time dependency; uses the libnccl the process (torch) already loaded ----
typedef int (*nccl_get_unique_id_t)(void*);
typedef int (*nccl_comm_init_rank_t)(void**, int, struct NcclId, int);
struct NcclId { char internal[128]; };
typedef int (*nccl_comm_destroy_t)(void*);
typedef int (*nccl_all_reduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_get_error_string_t)(int);

struct NcclApi {
  void* handle = nullptr;
  nccl_get_unique_id_t get_unique_id = nullptr;
  nccl_comm_init_rank_t comm_init_rank = nullptr;
  nccl_comm_destroy_t comm_destroy = nullptr;
  nccl_all_reduce_t all_reduce = nullptr;
  nccl_get_error_string_t err_str = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (api.handle) {
      api.get_unique_id = (nccl_get_unique_id_t)dlsym(api.handle, "ncclGetUniqueId");
      api.comm_init_rank = (nccl_comm_init_rank_t)dlsym(api.handle, "ncclCommInitRank");
      api.comm_destroy = (nccl_comm_destroy_t)dlsym(api.handle, "ncclCommDestroy");
      api.all_reduce = (nccl_all_reduce_t)dlsym(api.handle, "ncclAllReduce");
      api.err_str = (nccl_get_error_string_t)dlsym(api.handle, "ncclGetErrorString");
    }
  }
  if (!api.handle || !api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_reduce) return nullptr;
  return &api;
}

#define PCLS_NCCL(api, expr, what)                                                            \
  do {                                                                                        \
    int _r = (expr);                                                                          \
    if (_r != 0) {                                                                            \
      set_error("%s failed: %s", what, (api)->err_str ? (api)->err_str(_r) : "nccl error");   \
      return PCLS_ERR_NCCL;                                                                   \
    }                                                                                         \
  } while (0)

}  // namespace pcls

using namespace pcls;

extern "C" int pcls_confusion_update(const int32_t* label, const int32_t* pred, int64_t n, int num_classes,
                                     int64_t* cm, int64_t* dropped, pcls_stream stream) {
  PCLS_REQUIRE(num_classes >= 1 && num_classes <= CM_MAX_NC, "pcls_confusion_update: num_classes %d not in [1,%d]",
               num_classes, CM_MAX_NC);
  PCLS_REQUIRE(n >= 0 && cm != nullptr, "pcls_confusion_update: bad arguments");
  PCLS_REQUIRE(n == 0 || (label != nullptr && pred != nullptr), "pcls_confusion_update: label/pred must not be NULL");
  PCLS_REQUIRE(((uintptr_t)label % 16 == 0) && ((uintptr_t)pred % 16 == 0), "pcls_confusion_update: label/pred must be 16-byte aligned");
  return launch_confusion(label, pred, n, num_classes, cm, dropped, (cudaStream_t)stream);
}

extern "C" int pcls_comm_unique_id(char* h_id) {
  NcclApi* api = nccl_api();
  if (!api) { set_error("NCCL (libnccl.so.2) could not be loaded"); return PCLS_ERR_NCCL; }
  PCLS_NCCL(api, api->get_unique_id(h_id), "ncclGetUniqueId");
  return PCLS_OK;
}

extern "C" int pcls_comm_init(void** comm, int nranks, const char* h_id, int rank) {
  NcclApi* api = nccl_api();
  if (!api) { set_error("NCCL (libnccl.so.2) could not be loaded"); return PCLS_ERR_NCCL; }
  NcclId id;
  memcpy(id.internal, h_id, sizeof(id.internal));
  PCLS_NCCL(api, api->comm_init_rank(comm, nranks, id, rank), "ncclCommInitRank");
  return PCLS_OK;
}

extern "C" int pcls_comm_destroy(void* comm) {
  NcclApi* api = nccl_api();
  if (!api) { set_error("NCCL (libnccl.so.2) could not be loaded"); return PCLS_ERR_NCCL; }
  PCLS_NCCL(api, api->comm_destroy(comm), "ncclCommDestroy");
  return PCLS_OK;
}

extern "C" int pcls_confusion_allreduce(int64_t* cm, int num_classes, void* comm, pcls_stream stream) {
  NcclApi* api = nccl_api();
  if (!api) { set_error("NCCL (libnccl.so.2) could not be loaded"); return PCLS_ERR_NCCL; }
  PCLS_REQUIRE(cm != nullptr && comm != nullptr && num_classes > 0, "pcls_confusion_allreduce: bad arguments");
  // ncclInt64 = 4, ncclSum = 0 (nccl.h ncclDataType_t / ncclRedOp_t)
  PCLS_NCCL(api, api->all_reduce(cm, cm, (size_t)num_classes * num_classes, 4, 0, comm, (cudaStream_t)stream),
            "ncclAllReduce");
  return PCLS_OK;
}
