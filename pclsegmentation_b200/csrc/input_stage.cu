// Input stage: mask / normalise / zero-fill / append-mask (sm_100a).
// Replaces inference.py:47-72 == DataLoader.parse_sample (pcl_segmentation/data_loader/data_loader.py:153-187).
// (The network forward fuses the same arithmetic into its first load; this standalone entry exists for the
// eval.py flow, which also needs the label fix-up, and for parity tests of the stage itself.)
#include "common.cuh"

namespace pcls {

struct Norm5 { double mean[5]; double inv_unused; double std[5]; };

__global__ void __launch_bounds__(256)
input_stage_kernel(const float* __restrict__ sample, int channels, int64_t n_pixels, Norm5 nrm, int none_index,
                   float* __restrict__ lidar, uint8_t* __restrict__ mask, int32_t* __restrict__ label) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pixels; p += stride) {
    const float* s = sample + p * channels;
    float v[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) v[c] = __ldg(s + c);
    const bool m = v[4] > 0.0f;  // mask = depth > 0 (inference.py:53)
    if (lidar) {
      float* o = lidar + p * 6;
#pragma unroll
      for (int c = 0; c < 5; ++c)  // float64 like numpy (mean/std are float64 arrays), then float32
        o[c] = m ? (float)(((double)v[c] - nrm.mean[c]) / nrm.std[c]) : 0.0f;
      o[5] = m ? 1.0f : 0.0f;
    }
    if (mask) mask[p] = m ? 1 : 0;
    if (label) label[p] = m ? (int32_t)__ldg(s + 5) : none_index;  // label[~mask] = None (inference.py:65-68)
  }
}

}  // namespace pcls

extern "C" int pcls_input_stage(const float* sample, int channels, int64_t n_pixels, const double* h_mean5,
                                const double* h_std5, int none_index, float* lidar, uint8_t* mask,
                                int32_t* label, pcls_stream stream) {
  using namespace pcls;
  PCLS_REQUIRE(channels == 5 || channels == 6, "pcls_input_stage: channels must be 5 or 6, got %d", channels);
  PCLS_REQUIRE(label == nullptr || channels == 6, "pcls_input_stage: label output needs a 6-channel sample");
  PCLS_REQUIRE(h_mean5 != nullptr && h_std5 != nullptr, "pcls_input_stage: mean/std must not be NULL");
  PCLS_REQUIRE(n_pixels >= 0, "pcls_input_stage: negative n_pixels");
  if (n_pixels == 0) return PCLS_OK;
  PCLS_REQUIRE(sample != nullptr, "pcls_input_stage: sample is NULL");
  Norm5 nrm;
  for (int c = 0; c < 5; ++c) { nrm.mean[c] = h_mean5[c]; nrm.std[c] = h_std5[c]; }
  int64_t blocks = ceil_div(n_pixels, 256);
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  input_stage_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(sample, channels, n_pixels, nrm, none_index,
                                                                   lidar, mask, label);
  return check_launch("input_stage_kernel");
}
