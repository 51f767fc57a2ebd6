// Segmentation head: softmax -> argmax -> depth-zero mask (sm_100a).
//
// Replaces PCLSegmentationNetwork.segmentation_head (pcl_segmentation/nets/SegmentationNetwork.py:58-69).
// HBM-bound: NC*4 B read + 4 B written per pixel (+ NC*4 B when probabilities are materialised).
// Each warp stages 32 pixels x NC logits through shared memory so that global loads and the
// probability stores are fully coalesced 128-byte transactions; one lane then owns one pixel.
#include "nn_kernels.cuh"

namespace pcls {

constexpr int HEAD_MAX_NC = 32;
constexpr int HEAD_WARPS = 8;

__global__ void __launch_bounds__(HEAD_WARPS * 32)
head_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask, int64_t n_pixels, int nc,
            int none_index, float* __restrict__ probs, int32_t* __restrict__ preds) {
  __shared__ float tile[HEAD_WARPS][32 * (HEAD_MAX_NC + 1)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* t = tile[warp];
  const int pitch = nc + 1;  // odd pitch for nc even: conflict-free per-pixel rows
  const int64_t n_groups = (n_pixels + 31) / 32;
  for (int64_t g = (int64_t)blockIdx.x * HEAD_WARPS + warp; g < n_groups; g += (int64_t)gridDim.x * HEAD_WARPS) {
    const int64_t p0 = g * 32;
    const int np = (int)min((int64_t)32, n_pixels - p0);
    const float* src = logits + p0 * nc;
    const int n_el = np * nc;
    for (int e = lane; e < n_el; e += 32) t[(e / nc) * pitch + (e % nc)] = __ldg(src + e);
    __syncwarp();
    if (lane < np) {
      float v[HEAD_MAX_NC];
#pragma unroll
      for (int c = 0; c < HEAD_MAX_NC; ++c) v[c] = (c < nc) ? t[lane * pitch + c] : 0.0f;
      int best = softmax_argmax<HEAD_MAX_NC>(v, nc);
      if (mask != nullptr && mask[p0 + lane] == 0) best = none_index;
      preds[p0 + lane] = best;
      if (probs != nullptr) {
#pragma unroll
        for (int c = 0; c < HEAD_MAX_NC; ++c) if (c < nc) t[lane * pitch + c] = v[c];
      }
    }
    __syncwarp();
    if (probs != nullptr) {
      float* dst = probs + p0 * nc;
      for (int e = lane; e < n_el; e += 32) dst[e] = t[(e / nc) * pitch + (e % nc)];
    }
    __syncwarp();
  }
}

int launch_head(const float* logits, const uint8_t* mask, int64_t n_pixels, int nc, int none_index,
                float* probs, int32_t* preds, cudaStream_t s) {
  if (n_pixels == 0) return PCLS_OK;
  int64_t groups = (n_pixels + 31) / 32;
  int64_t blocks = ceil_div(groups, HEAD_WARPS);
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  head_kernel<<<(int)blocks, HEAD_WARPS * 32, 0, s>>>(logits, mask, n_pixels, nc, none_index, probs, preds);
  return check_launch("head_kernel");
}

}  // namespace pcls

extern "C" int pcls_head(const float* logits, const uint8_t* mask, int64_t n_pixels, int num_classes,
                         int none_index, float* probs, int32_t* preds, pcls_stream stream) {
  PCLS_REQUIRE(num_classes >= 1 && num_classes <= pcls::HEAD_MAX_NC, "pcls_head: num_classes %d not in [1,%d]",
               num_classes, pcls::HEAD_MAX_NC);
  PCLS_REQUIRE(n_pixels >= 0, "pcls_head: negative n_pixels");
  PCLS_REQUIRE(n_pixels == 0 || (logits != nullptr && preds != nullptr), "pcls_head: logits/preds must not be NULL");
  return pcls::launch_head(logits, mask, n_pixels, num_classes, none_index, probs, preds, (cudaStream_t)stream);
}
