// Input stage: mask / normalise / zero-fill / append-mask (sm_100a).
// Replaces inference.py:47-72 == DataLoader.parse_sample (pcl_segmentation/data_loader/data_loader.py:153-187).
// (The network forward fuses the same arithmetic into its first load; this standalone entry exists for the
// eval.py flow, which also needs the label fix-up, and for parity tests of the stage itself.)
#include "common.cuh"

namespace pcls {

struct Norm5 { double mean[5]; double inv_unused; double std[5]; };
constexpr int IS_MAX_NC = 32;
struct ClsWeight { float w[IS_MAX_NC]; int nc; };

__global__ void __launch_bounds__(256)
input_stage_kernel(const float* __restrict__ sample, int channels, int64_t n_pixels, Norm5 nrm, int none_index,
                   float* __restrict__ lidar, uint8_t* __restrict__ mask, int32_t* __restrict__ label, ClsWeight cw,
                   float* __restrict__ weight) {
  __shared__ float cw_s[IS_MAX_NC];
  if (threadIdx.x < IS_MAX_NC) cw_s[threadIdx.x] = cw.w[threadIdx.x];
  __syncthreads();
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
    if (label || weight) {
      const float lf = m ? __ldg(s + 5) : (float)none_index;       // label[~mask] = None (inference.py:65-68)
      if (label) label[p] = (int32_t)lf;
      if (weight) {   // weight = zeros; weight[label == l] = CLS_LOSS_WEIGHT[l] for l < NUM_CLASS (data_loader.py:181-185)
        const int li = (int)lf;
        weight[p] = (lf == (float)li && li >= 0 && li < cw.nc) ? cw_s[li] : 0.0f;
      }
    }
  }
}

// float64 -> float32 (round to nearest even, what numpy's astype / TF's cast do): the range-image files of the reference
// are float64 on disk (dataset_convert/semantic_kitti.py:173); uploading them as they are and narrowing on the device
// takes the conversion pass off the host.  16 bytes in, 8 bytes out per thread and step.
__global__ void __launch_bounds__(256)
cast_f64_f32_kernel(const double2* __restrict__ in, float2* __restrict__ out, int64_t n_pairs, const double* __restrict__ in1,
                    float* __restrict__ out1, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
    const double2 v = in[i];
    out[i] = make_float2((float)v.x, (float)v.y);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out1[n - 1] = (float)in1[n - 1];
}

// nuScenes LiDAR records (x, y, z, intensity, ring index: five float32 per point, laserscan_nuscenes.py:27-28, :139-145)
// -> the [n,4] (x,y,z,remission) point array the projection reads + the int32 ring index (astype(np.int32): truncation).
__global__ void __launch_bounds__(256)
unpack_xyzir_kernel(const float* __restrict__ rec5, int64_t n, float4* __restrict__ points4, int32_t* __restrict__ ring) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float* r = rec5 + i * 5;
    points4[i] = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
    if (ring) ring[i] = (int32_t)__ldg(r + 4);
  }
}

}  // namespace pcls

extern "C" int pcls_unpack_xyzir(const float* rec5, int64_t n, float* points4, int32_t* ring, pcls_stream stream) {
  using namespace pcls;
  PCLS_REQUIRE(n >= 0, "pcls_unpack_xyzir: negative n");
  if (n == 0) return PCLS_OK;
  PCLS_REQUIRE(rec5 != nullptr && points4 != nullptr, "pcls_unpack_xyzir: NULL buffer");
  PCLS_REQUIRE(((uintptr_t)points4 & 15) == 0, "pcls_unpack_xyzir: points4 must be 16-byte aligned");
  int64_t blocks = ceil_div(n, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  unpack_xyzir_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(rec5, n, reinterpret_cast<float4*>(points4), ring);
  return check_launch("unpack_xyzir_kernel");
}

extern "C" int pcls_cast_f64_f32(const double* in, float* out, int64_t n, pcls_stream stream) {
  using namespace pcls;
  PCLS_REQUIRE(n >= 0, "pcls_cast_f64_f32: negative n");
  if (n == 0) return PCLS_OK;
  PCLS_REQUIRE(in != nullptr && out != nullptr, "pcls_cast_f64_f32: NULL buffer");
  PCLS_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 7) == 0, "pcls_cast_f64_f32: buffers must be 16 / 8 byte aligned");
  int64_t blocks = ceil_div(n / 2 + 1, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cast_f64_f32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double2*>(in),
                                                                    reinterpret_cast<float2*>(out), n / 2, in, out, n);
  return check_launch("cast_f64_f32_kernel");
}

extern "C" int pcls_input_stage(const float* sample, int channels, int64_t n_pixels, const double* h_mean5,
                                const double* h_std5, int none_index, float* lidar, uint8_t* mask,
                                int32_t* label, const double* h_cls_loss_weight, int num_classes, float* weight,
                                pcls_stream stream) {
  using namespace pcls;
  PCLS_REQUIRE(channels == 5 || channels == 6, "pcls_input_stage: channels must be 5 or 6, got %d", channels);
  PCLS_REQUIRE((label == nullptr && weight == nullptr) || channels == 6,
               "pcls_input_stage: label / weight outputs need a 6-channel sample");
  PCLS_REQUIRE(weight == nullptr || (h_cls_loss_weight != nullptr && num_classes >= 1 && num_classes <= IS_MAX_NC),
               "pcls_input_stage: the weight output needs h_cls_loss_weight[num_classes], 1 <= num_classes <= %d", IS_MAX_NC);
  PCLS_REQUIRE(h_mean5 != nullptr && h_std5 != nullptr, "pcls_input_stage: mean/std must not be NULL");
  PCLS_REQUIRE(n_pixels >= 0, "pcls_input_stage: negative n_pixels");
  if (n_pixels == 0) return PCLS_OK;
  PCLS_REQUIRE(sample != nullptr, "pcls_input_stage: sample is NULL");
  Norm5 nrm;
  for (int c = 0; c < 5; ++c) { nrm.mean[c] = h_mean5[c]; nrm.std[c] = h_std5[c]; }
  int64_t blocks = ceil_div(n_pixels, 256);
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ClsWeight cw;
  cw.nc = weight ? num_classes : 0;
  for (int c = 0; c < IS_MAX_NC; ++c) cw.w[c] = (weight && c < num_classes) ? (float)h_cls_loss_weight[c] : 0.0f;
  input_stage_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(sample, channels, n_pixels, nrm, none_index,
                                                                   lidar, mask, label, cw, weight);
  return check_launch("input_stage_kernel");
}
