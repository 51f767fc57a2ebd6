// Spherical projection of raw LiDAR scans into range images (sm_100a).
//
// Replaces LaserScan.do_range_projection (dataset_convert/laserscan_semantic_kitti.py:106-166),
// do_range_projection_ring (dataset_convert/laserscan_nuscenes.py:191-223),
// SemLaserScan.do_label_projection (laserscan_semantic_kitti.py:269-279) and the converter assembly
// (dataset_convert/semantic_kitti.py:162-173).
//
// The reference sorts the points far-to-near (argsort) and relies on numpy's last-write-wins scatter.
// Here each point does ONE 64-bit atomicMin on its pixel's key (depth bits << 32 | index): the nearest
// point wins, ties go to the lowest index, no sort, no permutation gathers.  A second pass turns the
// keys into the [H,W,6] image with one 16-byte gather per occupied pixel.
//
// HBM-bound integer/byte work: point reads are 128-bit coalesced; the per-pixel pass writes 24 B + 4 B
// per pixel fully coalesced.  Algorithmic bytes per scan: 16 N read + H W (24 + 4) written.
//
// Bit-exactness: every float32 operation is spelled with an _rn intrinsic so that nvcc cannot contract
// mul+add into FMA - numpy evaluates each ufunc with its own rounding.
#include "common.cuh"

#include <type_traits>

namespace pcls {

struct ProjConsts {
  float abs_fov_down;  // f32(|fov_down|)
  float fov;           // f32(|fov_down| + |fov_up|)
  float pi;            // f32(pi)
  float Wf, Hf;
  float thr_x, thr_y;  // distance from an integer below which the float64 trig path decides the pixel
  int W, H;
};

__device__ __forceinline__ float point_depth(float x, float y, float z) {
  // np.linalg.norm(points, 2, axis=1) in float32 == sqrt((x*x + y*y) + z*z), separate roundings (:118)
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

__device__ __forceinline__ int scan_of_point(const int64_t* __restrict__ offsets, int B, int64_t i) {
  int lo = 0, hi = B;  // find b with offsets[b] <= i < offsets[b+1]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(offsets + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
project_scatter_kernel(const float4* __restrict__ points, const int32_t* __restrict__ ring,
                       const int64_t* __restrict__ offsets, int B, int64_t total, ProjConsts c,
                       unsigned long long* __restrict__ keys, int32_t* __restrict__ proj_x,
                       int32_t* __restrict__ proj_y, float* __restrict__ unproj_range) {
  // Every CTA walks ONE contiguous chunk of the point list, so a thread's scan index only ever moves forward: one binary
  // search at the start, then `while (i >= end of scan) ++b` (ncu on the grid-stride version: the kernel was issue-bound,
  // 386 instructions per point, 40 of them the per-point search over the scan offsets).
  const int64_t chunk = (total + gridDim.x - 1) / gridDim.x;
  const int64_t i_lo = (int64_t)blockIdx.x * chunk, i_hi = min(total, i_lo + chunk);
  if (i_lo >= i_hi) return;
  int b = scan_of_point(offsets, B, i_lo + threadIdx.x < i_hi ? i_lo + threadIdx.x : i_lo);
  int64_t off_b = __ldg(offsets + b), off_e = __ldg(offsets + b + 1);
  for (int64_t i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
    const float4 p = __ldg(points + i);
    while (i >= off_e) { ++b; off_b = off_e; off_e = __ldg(offsets + b + 1); }   // (empty scans: several steps)
    const uint32_t local = (uint32_t)(i - off_b);

    const float depth = point_depth(p.x, p.y, p.z);
    // Column.  Reference: yaw = -arctan2(y, x); proj_x = 0.5 * (yaw / pi + 1.0); proj_x *= W; floor; clamp (:126-140).
    // The contract is the CORRECTLY ROUNDED float32 arctan2 (float64 evaluation rounded once).  Only floor(proj_x)
    // leaves the kernel, so the 2-ulp float32 atan2f decides it unless proj_x lands within `thr` of an integer;
    // only those few points (< 1 %) pay for the float64 evaluation.  thr >> the propagated 2-ulp error (< 1e-3 px).
    float yaw = -atan2f(p.y, p.x);
    float fx = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(__fdiv_rn(yaw, c.pi), 1.0f)), c.Wf);
    float fl = floorf(fx);
    if (!(fx - fl >= c.thr_x && fx - fl <= 1.0f - c.thr_x)) {  // near an integer, or not finite
      yaw = -(float)atan2((double)p.y, (double)p.x);
      fx = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(__fdiv_rn(yaw, c.pi), 1.0f)), c.Wf);
      fl = floorf(fx);
    }
    bool ok = isfinite(fl);
    int col = (int)fmaxf(0.0f, fminf((float)(c.W - 1), fl));  // :138-140
    int row;
    if (ring == nullptr) {
      // pitch = arcsin(z / depth) (:127); proj_y = (1 - (pitch + |fov_down|) / fov) * H (:131,:135); same scheme
      const float q = __fdiv_rn(p.z, depth);
      float pitch = asinf(q);
      float fy = __fmul_rn(__fsub_rn(1.0f, __fdiv_rn(__fadd_rn(pitch, c.abs_fov_down), c.fov)), c.Hf);
      float fly = floorf(fy);
      if (!(fy - fly >= c.thr_y && fy - fly <= 1.0f - c.thr_y)) {
        pitch = (float)asin((double)q);
        fy = __fmul_rn(__fsub_rn(1.0f, __fdiv_rn(__fadd_rn(pitch, c.abs_fov_down), c.fov)), c.Hf);
        fly = floorf(fy);
      }
      ok = ok && isfinite(fly);
      row = (int)fmaxf(0.0f, fminf((float)(c.H - 1), fly));  // :143-145
    } else {
      row = (c.H - 1) - __ldg(ring + i);  // laserscan_nuscenes.py:215
      ok = ok && row >= 0 && row < c.H;
    }
    if (!ok) { row = -1; col = -1; }
    if (proj_x) proj_x[i] = col;
    if (proj_y) proj_y[i] = row;
    if (unproj_range) unproj_range[i] = depth;
    if (ok) {
      // depth >= 0, so its IEEE bits order like the value.  Depth variant: min depth, then min index.
      // Ring variant: in-order scatter == highest index wins == min of ~index.
      const unsigned long long hi = (ring == nullptr) ? (unsigned long long)__float_as_uint(depth)
                                                      : (unsigned long long)(~local);
      const unsigned long long key = (hi << 32) | (unsigned long long)local;
      atomicMin(keys + ((int64_t)b * c.H + row) * c.W + col, key);
    }
  }
}

__global__ void __launch_bounds__(256)
project_resolve_kernel(const float4* __restrict__ points, const uint32_t* __restrict__ labels,
                       const int64_t* __restrict__ offsets, int64_t n_pixels, int HW,
                       const unsigned long long* __restrict__ keys, const int32_t* __restrict__ lut,
                       int lut_len, float empty_fill, float* __restrict__ image,
                       int32_t* __restrict__ proj_idx, int32_t* __restrict__ proj_sem) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < n_pixels; pix += stride) {
    const unsigned long long key = __ldg(keys + pix);
    float x = empty_fill, y = empty_fill, z = empty_fill, r = empty_fill, d = empty_fill;
    int idx = -1, sem = 0;
    if (key != ~0ull) {
      idx = (int)(uint32_t)(key & 0xffffffffull);
      const int b = (int)(pix / HW);
      const int64_t g = __ldg(offsets + b) + idx;
      const float4 p = __ldg(points + g);
      x = p.x; y = p.y; z = p.z; r = p.w;
      d = point_depth(p.x, p.y, p.z);
      if (labels) sem = (int)(__ldg(labels + g) & 0xFFFFu);  // set_label :246
      if (!(d > 0.0f)) {  // converter: mask = proj_range > 0 (semantic_kitti.py:162-165)
        if (empty_fill == 0.0f) { x = y = z = r = d = 0.0f; }
      }
    }
    if (image) {
      int mapped = sem;
      if (lut) mapped = (sem >= 0 && sem < lut_len) ? __ldg(lut + sem) : 0;
      float2* o = reinterpret_cast<float2*>(image + pix * 6);
      o[0] = make_float2(x, y);
      o[1] = make_float2(z, r);
      o[2] = make_float2(d, (float)mapped);
    }
    if (proj_idx) proj_idx[pix] = idx;
    if (proj_sem) proj_sem[pix] = sem;
  }
}

// Winner keys -> the NETWORK INPUT itself (fused scan -> labels pipeline): per pixel the winning point's (x, y, z,
// remission, range) goes through the input stage of inference.py:50-62 (mask = range > 0, float64 normalise, zero where
// empty, mask as channel 5) and is stored as the 16-byte [B,H,W,8] 16-bit pixel conv1 reads, plus the mask byte the head
// reads: no float32 [B,H,W,6] image round trip and no separate input kernel.
struct NormD { double mean[5], std[5]; };
template <typename T>
__global__ void __launch_bounds__(256)
project_resolve_net_input_kernel(const float4* __restrict__ points, const int64_t* __restrict__ offsets, int64_t n_pixels,
                                 int HW, const unsigned long long* __restrict__ keys, NormD nrm, int4* __restrict__ input8,
                                 uint8_t* __restrict__ mask, int32_t* __restrict__ proj_idx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < n_pixels; pix += stride) {
    const unsigned long long key = __ldg(keys + pix);
    float v[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    int idx = -1;
    if (key != ~0ull) {
      idx = (int)(uint32_t)(key & 0xffffffffull);
      const float4 p = __ldg(points + __ldg(offsets + (int)(pix / HW)) + idx);
      v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w;
      v[4] = point_depth(p.x, p.y, p.z);
    }
    const bool m = v[4] > 0.0f;
    T h[8];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const float f = m ? (float)(((double)v[c] - nrm.mean[c]) / nrm.std[c]) : 0.0f;
      h[c] = sizeof(T) == 2 && std::is_same<T, __half>::value ? (T)__float2half_rn(f) : (T)__float2bfloat16_rn(f);
    }
    const float one = m ? 1.0f : 0.0f;
    h[5] = std::is_same<T, __half>::value ? (T)__float2half_rn(one) : (T)__float2bfloat16_rn(one);
    h[6] = h[7] = std::is_same<T, __half>::value ? (T)__float2half_rn(0.0f) : (T)__float2bfloat16_rn(0.0f);
    input8[pix] = *reinterpret_cast<const int4*>(h);
    mask[pix] = m ? 1 : 0;
    if (proj_idx) proj_idx[pix] = idx;
  }
}

static int grid_for(int64_t n, int threads) {
  int64_t blocks = ceil_div(n, threads);
  int64_t cap = (int64_t)sm_count() * 16;  // 16 resident 256-thread CTAs... grid-stride beyond that
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace pcls

using namespace pcls;

extern "C" int pcls_project_scatter(const float* points, const int32_t* ring, const int64_t* offsets, int B,
                                    int64_t total_points, int H, int W, double fov_up_deg,
                                    double fov_down_deg, uint64_t* keys, int32_t* proj_x, int32_t* proj_y,
                                    float* unproj_range, pcls_stream stream) {
  PCLS_REQUIRE(B >= 0 && H > 0 && W > 0 && total_points >= 0, "pcls_project_scatter: bad sizes B=%d H=%d W=%d total=%lld",
               B, H, W, (long long)total_points);
  PCLS_REQUIRE(keys != nullptr && offsets != nullptr, "pcls_project_scatter: keys/offsets must not be NULL");
  PCLS_REQUIRE(total_points == 0 || points != nullptr, "pcls_project_scatter: points is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_pixels = (int64_t)B * H * W;
  if (n_pixels > 0) PCLS_CHECK_CUDA(cudaMemsetAsync(keys, 0xFF, n_pixels * sizeof(uint64_t), s));
  if (total_points == 0 || B == 0) return PCLS_OK;
  // laser parameters exactly as the reference computes them in Python doubles (:113-115)
  const double pi = 3.141592653589793;
  const double fov_up = fov_up_deg / 180.0 * pi;
  const double fov_down = fov_down_deg / 180.0 * pi;
  const double fov = fabs(fov_down) + fabs(fov_up);
  ProjConsts c;
  c.abs_fov_down = (float)fabs(fov_down);
  c.fov = (float)fov;
  c.pi = (float)pi;
  c.W = W; c.H = H; c.Wf = (float)W; c.Hf = (float)H;
  // How far the float32 fast path can be from the correctly rounded one, in pixels:
  //   column: atan2f <= 2 ulp of |yaw| <= pi (4.8e-7 rad) -> W / (2 pi) * 4.8e-7 = 7.6e-8 W, plus the roundings of yaw / pi,
  //           + 1 and * W (<= 1.0e-7 W together): < 1.8e-7 W (3.7e-4 px at W = 2048);
  //   row:    z / depth (0.5 ulp) and asinf (2 ulp of |pitch| < 0.6): < 1e-7 rad -> 1e-7 H / fov, plus three roundings of
  //           values <= H: < 2.5e-7 H / fov + 2e-7 H (3e-5 px at H = 64, fov 28 deg).
  // The float64 evaluation decides every point within thr of a pixel boundary, thr = 4x (column) / 6x (row) those bounds.
  // (Round 1 used 8e-3 / 4e-3 px: 41 % / 23 % of the warps had a lane on the slow path and the whole warp paid for it.)
  c.thr_x = fmaxf(1.5e-3f, 7.5e-7f * (float)W);
  c.thr_y = fmaxf(2e-4f, 1.5e-6f * (float)H / (float)fov + 1.2e-6f * (float)H);
  project_scatter_kernel<<<grid_for(total_points, 256), 256, 0, s>>>(
      reinterpret_cast<const float4*>(points), ring, offsets, B, total_points, c,
      reinterpret_cast<unsigned long long*>(keys), proj_x, proj_y, unproj_range);
  return check_launch("project_scatter_kernel");
}

extern "C" int pcls_project_resolve(const float* points, const uint32_t* labels, const int64_t* offsets, int B,
                                    int H, int W, const uint64_t* keys, const int32_t* label_lut, int lut_len,
                                    float empty_fill, float* image, int32_t* proj_idx, int32_t* proj_sem_label,
                                    pcls_stream stream) {
  PCLS_REQUIRE(B >= 0 && H > 0 && W > 0, "pcls_project_resolve: bad sizes");
  PCLS_REQUIRE(keys != nullptr && offsets != nullptr, "pcls_project_resolve: keys/offsets must not be NULL");
  const int64_t n_pixels = (int64_t)B * H * W;
  if (n_pixels == 0) return PCLS_OK;
  project_resolve_kernel<<<grid_for(n_pixels, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(points), labels, offsets, n_pixels, H * W,
      reinterpret_cast<const unsigned long long*>(keys), label_lut, lut_len, empty_fill, image, proj_idx,
      proj_sem_label);
  return check_launch("project_resolve_kernel");
}

extern "C" int pcls_project_resolve_net_input(const float* points, const int64_t* offsets, int B, int H, int W,
                                              const uint64_t* keys, const double* h_mean5, const double* h_std5,
                                              int precision, void* input8, uint8_t* mask, int32_t* proj_idx,
                                              pcls_stream stream) {
  PCLS_REQUIRE(B >= 0 && H > 0 && W > 0, "pcls_project_resolve_net_input: bad sizes");
  PCLS_REQUIRE(keys != nullptr && offsets != nullptr && input8 != nullptr && mask != nullptr && h_mean5 != nullptr && h_std5 != nullptr,
               "pcls_project_resolve_net_input: NULL argument");
  PCLS_REQUIRE(precision == PCLS_F16 || precision == PCLS_BF16, "pcls_project_resolve_net_input: bad precision %d", precision);
  const int64_t n_pixels = (int64_t)B * H * W;
  if (n_pixels == 0) return PCLS_OK;
  NormD nrm;
  for (int c = 0; c < 5; ++c) { nrm.mean[c] = h_mean5[c]; nrm.std[c] = h_std5[c]; }
  const int grid = grid_for(n_pixels, 256);
  if (precision == PCLS_F16)
    project_resolve_net_input_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(points), offsets, n_pixels, H * W, reinterpret_cast<const unsigned long long*>(keys), nrm,
        reinterpret_cast<int4*>(input8), mask, proj_idx);
  else
    project_resolve_net_input_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(points), offsets, n_pixels, H * W, reinterpret_cast<const unsigned long long*>(keys), nrm,
        reinterpret_cast<int4*>(input8), mask, proj_idx);
  return check_launch("project_resolve_net_input_kernel");
}
