// test_step on the device (sm_100a): loss sums and the class-weighted confusion matrix in ONE pass over the forward's
// outputs.
//
// Replaces, forward only, PCLSegmentationNetwork.test_step (pcl_segmentation/nets/SegmentationNetwork.py:118-131):
//   focal loss (:71-91)       sum((1 - p)^gamma * onehot(label) * -log(p) * weight * mask) / sum(mask) * coef,
//                             p = probabilities + DENOM_EPSILON (float32, like the TF graph)
//   sparse CE (:49, :125)     Keras SparseCategoricalCrossentropy on probabilities with sample weights: clip every class
//                             to [1e-7, 1 - 1e-7], -(log p_c[label] - log sum_c p_c) * weight, mean over ALL elements
//   weighted MeanIoU (:129)   tf.math.confusion_matrix(label, pred, weights=weight): cm[label, pred] += weight
// The kernel only accumulates numerators / denominators (float64); the divisions and the running mean stay on the host.
//
// HBM-bound: NC * 4 (probabilities) + 4 (label) + 4 (prediction) + 4 (weight) + 1 (mask) bytes per pixel, one read each.
// A warp owns 32 consecutive pixels: their probability rows are one contiguous run of 32 * NC floats, staged through
// shared memory with coalesced loads (lane = pixel afterwards).  Per-block float64 histogram in shared memory, one float64
// global atomic per non-empty bin and block.
#include "common.cuh"

namespace pcls {

constexpr int VS_WARPS = 8;
constexpr int VS_MAX_NC = 32;

__global__ void __launch_bounds__(VS_WARPS * 32)
validation_kernel(const float* __restrict__ probs, const int32_t* __restrict__ label, const int32_t* __restrict__ pred,
                  const uint8_t* __restrict__ mask, const float* __restrict__ weight, int64_t n, int nc, int loss_kind,
                  float eps, float gamma, double* __restrict__ loss_acc, double* __restrict__ cm_w,
                  unsigned long long* __restrict__ dropped) {
  extern __shared__ double vs_smem[];          // [nc * nc] weighted histogram | [VS_WARPS][32 * nc + 32] f32 rows
  double* const hist = vs_smem;
  const int nbins = cm_w ? nc * nc : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pitch = nc | 1;                    // odd pitch: lane = pixel reads are conflict-free
  float* const rows = reinterpret_cast<float*>(vs_smem + nbins) + warp * (32 * pitch);
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0.0;
  __syncthreads();

  double num = 0.0, den = 0.0;
  unsigned int my_dropped = 0;
  const int64_t n_groups = (n + 31) / 32;
  for (int64_t g = (int64_t)blockIdx.x * VS_WARPS + warp; g < n_groups; g += (int64_t)gridDim.x * VS_WARPS) {
    const int64_t p0 = g * 32;
    const int np = (int)min((int64_t)32, n - p0);
    if (loss_kind != 0) {                      // both losses read probabilities
      const float* src = probs + p0 * nc;
      for (int e = lane; e < np * nc; e += 32) rows[(e / nc) * pitch + (e % nc)] = __ldg(src + e);
      __syncwarp();
    }
    if (lane < np) {
      const int64_t i = p0 + lane;
      const int y = __ldg(label + i);
      const float w = weight ? __ldg(weight + i) : 1.0f;
      const bool y_ok = (unsigned)y < (unsigned)nc;
      if (loss_kind == 1) {                    // focal (tf.one_hot of an out-of-range label is a zero row)
        const float m = mask ? (mask[i] ? 1.0f : 0.0f) : 1.0f;
        den += (double)m;
        if (y_ok && m != 0.0f) {
          const float p = rows[lane * pitch + y] + eps;
          num += (double)(powf(1.0f - p, gamma) * -logf(p)) * (double)w;
        }
      } else if (loss_kind == 2) {             // Keras sparse categorical cross-entropy on probabilities
        float s = 0.0f;
        for (int c = 0; c < nc; ++c) s += fminf(fmaxf(rows[lane * pitch + c], 1e-7f), 1.0f - 1e-7f);
        den += 1.0;
        if (y_ok) {
          const float py = fminf(fmaxf(rows[lane * pitch + y], 1e-7f), 1.0f - 1e-7f);
          num += ((double)logf(s) - (double)logf(py)) * (double)w;
        }
      }
      if (cm_w) {
        const int q = __ldg(pred + i);
        if (y_ok && (unsigned)q < (unsigned)nc) atomicAdd(hist + y * nc + q, (double)w);
        else ++my_dropped;
      }
    }
    __syncwarp();
  }
  // loss sums: warp shuffle, then one pair of global atomics per warp
  if (loss_kind != 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      num += __shfl_down_sync(0xffffffffu, num, o);
      den += __shfl_down_sync(0xffffffffu, den, o);
    }
    if (lane == 0 && (num != 0.0 || den != 0.0)) { atomicAdd(loss_acc, num); atomicAdd(loss_acc + 1, den); }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += blockDim.x)
    if (hist[b] != 0.0) atomicAdd(cm_w + b, hist[b]);
  if (dropped != nullptr && my_dropped) atomicAdd(dropped, (unsigned long long)my_dropped);
}

}  // namespace pcls

using namespace pcls;

extern "C" int pcls_validation_update(const float* probs, const int32_t* label, const int32_t* pred, const uint8_t* mask,
                                      const float* weight, int64_t n, int num_classes, int loss_kind, double eps,
                                      double gamma, double* loss_acc, double* cm_w, int64_t* dropped, pcls_stream stream) {
  PCLS_REQUIRE(n >= 0 && num_classes >= 1 && num_classes <= VS_MAX_NC, "pcls_validation_update: bad sizes n=%lld NC=%d (NC <= %d)",
               (long long)n, num_classes, VS_MAX_NC);
  PCLS_REQUIRE(loss_kind >= 0 && loss_kind <= 2, "pcls_validation_update: loss_kind must be 0 (none), 1 (focal) or 2 (sparse CE)");
  PCLS_REQUIRE(loss_kind == 0 || n == 0 || (probs != nullptr && loss_acc != nullptr), "pcls_validation_update: probs / loss_acc is NULL");
  PCLS_REQUIRE(cm_w == nullptr || n == 0 || pred != nullptr, "pcls_validation_update: pred is NULL");
  PCLS_REQUIRE(n == 0 || label != nullptr, "pcls_validation_update: label is NULL");
  if (n == 0 || (loss_kind == 0 && cm_w == nullptr)) return PCLS_OK;
  const int nc = num_classes;
  const size_t smem = (size_t)(cm_w ? nc * nc : 0) * sizeof(double) + (size_t)VS_WARPS * 32 * (nc | 1) * sizeof(float);
  int64_t blocks = ceil_div(ceil_div(n, 32), VS_WARPS * 4);
  const int64_t cap = (int64_t)sm_count() * 6;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  validation_kernel<<<(int)blocks, VS_WARPS * 32, smem, (cudaStream_t)stream>>>(
      probs, label, pred, mask, weight, n, nc, loss_kind, (float)eps, (float)gamma, loss_acc, cm_w,
      reinterpret_cast<unsigned long long*>(dropped));
  return check_launch("validation_kernel");
}
