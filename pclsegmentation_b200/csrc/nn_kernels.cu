// CUDA-core network kernels (sm_100a): generic direct convolution (every conv flavour of the two nets),
// network input conversion with the fused input stage, 3x3/s2 max-pool, the fused CAM gate.
//
// The direct convolution is the always-available CUDA path and the on-GPU cross-check for the tcgen05
// implicit-GEMM kernel (conv_tc.cu), which takes over every layer whose channel counts fit UMMA tiles.
#include "nn_kernels.cuh"

namespace pcls {

// --------------------------------------------------------------------------------------------------
// tap geometry shared by all conv implementations
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tap_source(const ConvParams& p, int tap, int h, int wo, int& hi, int& wi) {
  switch (p.mode) {
    case MODE_1x1:
      hi = h; wi = wo; return true;
    case MODE_3x3_S1:
      hi = h + tap / 3 - 1; wi = wo + tap % 3 - 1; break;
    case MODE_3x3_S2:  // TF SAME: total pad 1 for even Win -> 0 left / 1 right
      hi = h + tap / 3 - 1; wi = 2 * wo + tap % 3 - p.pad_left; break;
    default: {         // MODE_DECONV: out[m] += in[j] * w[k] with m = 2j + k - 1
      const int t = wo + 1 - tap;
      if (t < 0 || (t & 1)) return false;
      hi = h; wi = t >> 1; break;
    }
  }
  return hi >= 0 && hi < p.H && wi >= 0 && wi < p.Win;
}

// --------------------------------------------------------------------------------------------------
// direct convolution: 64 output pixels x 64 output channels per CTA, fp32 accumulate
// --------------------------------------------------------------------------------------------------
constexpr int DC_PX = 64, DC_CO = 64, DC_K = 16, DC_LD = 68;

template <typename T>
__global__ void __launch_bounds__(128)
conv_direct_kernel(const ConvParams p, const int64_t n_out) {
  __shared__ __align__(16) float As[DC_K][DC_LD];
  __shared__ __align__(16) float Ws[DC_K][DC_LD];
  const int tid = threadIdx.x;
  const int pxg = tid & 15, cg = tid >> 4;
  const int64_t pix0 = (int64_t)blockIdx.x * DC_PX;
  const int co0 = blockIdx.y * DC_CO;

  // the pixel / weight row this thread stages
  const int lpx = tid >> 1, lj = tid & 1;
  const int64_t lpix = pix0 + lpx;
  const bool lvalid = lpix < n_out;
  int lb = 0, lh = 0, lwo = 0;
  if (lvalid) {
    lwo = (int)(lpix % p.Wout);
    const int64_t r = lpix / p.Wout;
    lh = (int)(r % p.H);
    lb = (int)(r / p.H);
  }
  const int lco = co0 + lpx;  // weight row staged by this thread
  const T* in = reinterpret_cast<const T*>(p.in);
  const T* wt = reinterpret_cast<const T*>(p.w);

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  for (int tap = 0; tap < p.ntaps; ++tap) {
    int hi = 0, wi = 0;
    const bool src_ok = lvalid && tap_source(p, tap, lh, lwo, hi, wi);
    const T* src = in + (((int64_t)lb * p.H + hi) * p.Win + wi) * p.in_channels;
    const T* wrow = wt + ((int64_t)tap * p.cout_pad + lco) * p.cin_pad;
    for (int ck = 0; ck < p.cin_pad; ck += DC_K) {
      const int c = ck + 8 * lj;
      float f[8];
      if (src_ok && c < p.in_channels) {
        unpack8<T>(__ldg(reinterpret_cast<const int4*>(src + c)), f);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) As[8 * lj + i][lpx] = f[i];
      if (lco < p.cout_pad) {
        unpack8<T>(__ldg(reinterpret_cast<const int4*>(wrow + c)), f);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) Ws[8 * lj + i][lpx] = f[i];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < DC_K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][pxg * 4]);
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][cg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][cg * 8 + 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // epilogue: bias (BN folded), activation, residual adds (after the activation), store
  const T* res0 = reinterpret_cast<const T*>(p.res0);
  const T* res1 = reinterpret_cast<const T*>(p.res1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pix = pix0 + pxg * 4 + i;
    if (pix >= n_out) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = co0 + cg * 8 + j;
      if (co >= p.cout) continue;
      float v = apply_act(acc[i][j] + __ldg(p.bias + co), p.act);
      if (res0) v += to_f32<T>(res0[pix * p.res0_channels + p.out_coff + co]);
      if (res1) v += to_f32<T>(res1[pix * p.res1_channels + p.out_coff + co]);
      if (p.out_f32) reinterpret_cast<float*>(p.out)[pix * p.out_channels + p.out_coff + co] = v;
      else reinterpret_cast<T*>(p.out)[pix * p.out_channels + p.out_coff + co] = from_f32<T>(v);
    }
  }
}

template <typename T>
int launch_conv_direct(const ConvParams& p, int B, cudaStream_t s) {
  const int64_t n_out = (int64_t)B * p.H * p.Wout;
  if (n_out == 0) return PCLS_OK;
  dim3 grid((unsigned)ceil_div(n_out, DC_PX), (unsigned)ceil_div(p.cout, DC_CO));
  conv_direct_kernel<T><<<grid, 128, 0, s>>>(p, n_out);
  return check_launch("conv_direct_kernel");
}
template int launch_conv_direct<__half>(const ConvParams&, int, cudaStream_t);
template int launch_conv_direct<__nv_bfloat16>(const ConvParams&, int, cudaStream_t);

// --------------------------------------------------------------------------------------------------
// network input: float32 [n,ch] -> 16-bit [n,8] (6 channels + 2 zero pad) + u8 mask, optional fused input stage
// --------------------------------------------------------------------------------------------------
struct Norm5f { double mean[5]; double std[5]; };

template <typename T>
__global__ void __launch_bounds__(256)
net_input_kernel(const float* __restrict__ lidar, int channels, const uint8_t* __restrict__ mask_in, int raw, Norm5f nrm,
                 int64_t n_pixels, T* __restrict__ out8, uint8_t* __restrict__ mask_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pixels; p += stride) {
    const float* s = lidar + p * channels;
    float f[8];
#pragma unroll
    for (int c = 0; c < 5; ++c) f[c] = __ldg(s + c);
    bool m;
    if (raw) {  // inference.py:50-62 fused: mask = depth > 0, float64 normalise, zero where empty, append mask
      m = f[4] > 0.0f;
#pragma unroll
      for (int c = 0; c < 5; ++c) f[c] = m ? (float)(((double)f[c] - nrm.mean[c]) / nrm.std[c]) : 0.0f;
      f[5] = m ? 1.0f : 0.0f;
      if (mask_in) m = mask_in[p] != 0;
    } else {    // already normalised 6-channel reference input; channel 5 is the mask
      f[5] = __ldg(s + 5);
      m = mask_in ? (mask_in[p] != 0) : (f[5] != 0.0f);
    }
    f[6] = 0.0f; f[7] = 0.0f;
    reinterpret_cast<int4*>(out8)[p] = pack8<T>(f);
    mask_out[p] = m ? 1 : 0;
  }
}

template <typename T>
int launch_net_input(const float* lidar, int channels, const uint8_t* mask_in, bool raw, const double* mean5,
                     const double* std5, int64_t n_pixels, T* out8, uint8_t* mask_out, cudaStream_t s) {
  if (n_pixels == 0) return PCLS_OK;
  Norm5f nrm;
  for (int c = 0; c < 5; ++c) { nrm.mean[c] = raw ? mean5[c] : 0.0; nrm.std[c] = raw ? std5[c] : 1.0; }
  int64_t blocks = ceil_div(n_pixels, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  net_input_kernel<T><<<(int)blocks, 256, 0, s>>>(lidar, channels, mask_in, raw ? 1 : 0, nrm, n_pixels, out8, mask_out);
  return check_launch("net_input_kernel");
}
template int launch_net_input<__half>(const float*, int, const uint8_t*, bool, const double*, const double*, int64_t,
                                      __half*, uint8_t*, cudaStream_t);
template int launch_net_input<__nv_bfloat16>(const float*, int, const uint8_t*, bool, const double*, const double*,
                                             int64_t, __nv_bfloat16*, uint8_t*, cudaStream_t);

// --------------------------------------------------------------------------------------------------
// tf.nn.max_pool2d(ksize=3, strides=[1,2], padding='SAME'): rows h-1..h+1, cols 2wo-pl .. 2wo-pl+2
// one thread = one output pixel x 8 channels (128-bit loads/stores); padding never wins the max
// --------------------------------------------------------------------------------------------------
template <typename T> struct Vec2;
template <> struct Vec2<__half> { using type = __half2; };
template <> struct Vec2<__nv_bfloat16> { using type = __nv_bfloat162; };

template <typename T>
__device__ __forceinline__ int4 max8(const int4& a, const int4& b) {
  using V = typename Vec2<T>::type;
  int4 r;
  const V* x = reinterpret_cast<const V*>(&a);
  const V* y = reinterpret_cast<const V*>(&b);
  V* o = reinterpret_cast<V*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = __hmax2(x[i], y[i]);
  return r;
}

template <typename T> __device__ __forceinline__ int4 neg_inf8();
template <> __device__ __forceinline__ int4 neg_inf8<__half>() { return make_int4(0xFC00FC00, 0xFC00FC00, 0xFC00FC00, 0xFC00FC00); }
template <> __device__ __forceinline__ int4 neg_inf8<__nv_bfloat16>() { return make_int4(0xFF80FF80, 0xFF80FF80, 0xFF80FF80, 0xFF80FF80); }

// Column-streaming form: a thread owns one output column x 8 channels of one frame and walks down the H rows.  Per
// input row it takes the horizontal 3-max (three 128-bit loads, two of them shared with the neighbour thread through
// L1) and keeps the last two row maxima in registers, so every input vector is requested 1.5x instead of 4.5x.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3_s2_kernel(const int4* __restrict__ in, int4* __restrict__ out, int64_t n_cols, int H, int Win, int Wout,
                     int CV, int pad_left) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (frame, wo, cv)
  if (i >= n_cols) return;
  const int cv = (int)(i % CV);
  const int wo = (int)((i / CV) % Wout);
  const int64_t b = i / ((int64_t)CV * Wout);
  const int4* img = in + b * H * Win * CV + cv;
  int4* oimg = out + (b * H * Wout + wo) * CV + cv;
  const int w_lo = 2 * wo - pad_left;
  const int4 NEG = neg_inf8<T>();
  auto hrow = [&](int h) -> int4 {
    if (h < 0 || h >= H) return NEG;
    const int4* row = img + (int64_t)h * Win * CV;
    int4 m = NEG;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int wi = w_lo + dx;
      if (wi >= 0 && wi < Win) m = max8<T>(m, __ldg(row + (int64_t)wi * CV));
    }
    return m;
  };
  int4 prev = NEG, cur = hrow(0);
  for (int h = 0; h < H; ++h) {
    const int4 nxt = hrow(h + 1);
    oimg[(int64_t)h * Wout * CV] = max8<T>(max8<T>(prev, cur), nxt);
    prev = cur; cur = nxt;
  }
}

template <typename T>
int launch_maxpool3x3_s2(const T* in, T* out, int B, int H, int Win, int Wout, int C, int pad_left, cudaStream_t s) {
  const int CV = C / 8;
  const int64_t n_cols = (int64_t)B * Wout * CV;
  if (n_cols == 0) return PCLS_OK;
  maxpool3x3_s2_kernel<T><<<(unsigned)ceil_div(n_cols, 256), 256, 0, s>>>(
      reinterpret_cast<const int4*>(in), reinterpret_cast<int4*>(out), n_cols, H, Win, Wout, CV, pad_left);
  return check_launch("maxpool3x3_s2_kernel");
}
template int launch_maxpool3x3_s2<__half>(const __half*, __half*, int, int, int, int, int, int, cudaStream_t);
template int launch_maxpool3x3_s2<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, int, int, int, int, int, int,
                                                 cudaStream_t);

// --------------------------------------------------------------------------------------------------
// CAM (nets/SqueezeSegV2.py:66-70), one row-streaming kernel:
//   pool = maxpool7x7_SAME(x); s = relu(W1^T pool + b1); e = sigmoid(W2^T s + b2); out = x * e
// A CTA owns a strip of TW columns of one frame and walks down the H rows.  Each input row is staged ONCE in
// shared memory (TW + 6 columns, double buffered, next row prefetched into registers while the current one is
// processed); a thread owns 8 channels of one column, takes the horizontal 7-max from the staged row and keeps the
// last seven horizontal maxima plus the last four inputs in registers, so the vertical 7-max and the gate of row
// h - 3 need no further memory traffic.  HBM traffic: (TW + 6) / TW reads + 1 write of the tensor.
// The two 1x1 convolutions are tiny: partial dot products per 8-channel lane, xor-shuffle reduction over the
// C/8 lanes of a pixel.
// --------------------------------------------------------------------------------------------------
template <typename T, int C>
__global__ void __launch_bounds__(256)
cam_kernel(const int4* __restrict__ in, int4* __restrict__ out, CamParams p, int H, int W) {
  constexpr int CV = C / 8;          // 16-byte vectors per pixel (8 or 16)
  constexpr int R = C / 16;          // reduced channels (4 or 8)
  constexpr int TW = 256 / CV;       // columns per CTA (32 or 16)
  constexpr int ROWV = (TW + 6) * CV;  // vectors per staged row
  constexpr int LPT = (ROWV + 255) / 256;
  constexpr int NB = 5, DIST = 3;  // cp.async ring: rows r+1..r+3 in flight while row r is processed
  __shared__ int4 rowbuf[NB][ROWV];
  // weights in shared memory, laid out so that a thread fetches its 8 x R squeeze weights / R x 8 excitation weights
  // with 128-bit loads: per-cv blocks of 8*R floats, padded by 4 floats so that the quarter-warp's 16-byte accesses hit
  // distinct banks (lanes of another pixel with the same cv broadcast)
  //   w1s[cv * WS + k * R + j] = W1[cv * 8 + k][j]      w2s[cv * WS + j * 8 + k] = W2[j][cv * 8 + k]
  constexpr int WS = 8 * R + 4;
  __shared__ __align__(16) float w1s[CV * WS], w2s[CV * WS], b1s[R], b2s[C];
  for (int i = threadIdx.x; i < C * R; i += 256) {
    { const int ch = i / R, j = i % R; w1s[(ch / 8) * WS + (ch % 8) * R + j] = p.w1[i]; }
    { const int j = i / C, ch = i % C; w2s[(ch / 8) * WS + j * 8 + (ch % 8)] = p.w2[i]; }
  }
  for (int i = threadIdx.x; i < R; i += 256) b1s[i] = p.b1[i];
  for (int i = threadIdx.x; i < C; i += 256) b2s[i] = p.b2[i];

  const int w0 = blockIdx.x * TW;
  const int64_t b = blockIdx.y;
  const int4* img = in + b * H * W * CV;
  int4* oimg = out + b * H * W * CV;
  const int cv = threadIdx.x % CV, c = threadIdx.x / CV;  // this thread's column (0..TW-1) and channel vector
  const int4 NEG = neg_inf8<T>();

  // stage input row r into ring slot r % NB: in-image vectors travel global -> shared with cp.async (no registers,
  // DIST rows in flight), padding (outside the image) is written as -inf so that it never wins the max
  auto stage_row = [&](int r) {
    int4* dst = rowbuf[((r % NB) + NB) % NB];
#pragma unroll
    for (int k = 0; k < LPT; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < ROWV) {
        const int col = w0 - 3 + i / CV;
        if (r >= 0 && r < H && col >= 0 && col < W) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i);
          const int4* src = img + ((int64_t)r * W + col) * CV + (i % CV);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
        } else {
          dst[i] = NEG;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int4 win[7];   // horizontal maxima of the last seven input rows (oldest first)
  int4 xr[4];    // inputs of the last four rows at this thread's pixel (oldest first)
#pragma unroll
  for (int k = 0; k < 7; ++k) win[k] = NEG;
#pragma unroll
  for (int k = 0; k < 4; ++k) xr[k] = NEG;

  for (int r = 0; r < DIST; ++r) stage_row(r);  // rows -3..-1 are padding and already -inf in the window
  const bool col_ok = (w0 + c) < W;
  for (int r = 0; r < H + 3; ++r) {          // r = newest input row in the window; output row o = r - 3
    const int buf = r % NB;
    stage_row(r + DIST);                     // rows >= H are written as -inf
    asm volatile("cp.async.wait_group %0;" ::"n"(DIST) : "memory");   // row r has landed (this thread's copies)
    __syncthreads();                         // ... and everybody else's; also fences the slot reused by stage_row(r + DIST + 1)
    // horizontal 7-max of row r at this column: staged columns c .. c+6 (c+3 is the centre)
    int4 hm = rowbuf[buf][(c + 0) * CV + cv];
#pragma unroll
    for (int d = 1; d < 7; ++d) hm = max8<T>(hm, rowbuf[buf][(c + d) * CV + cv]);
    const int4 centre = rowbuf[buf][(c + 3) * CV + cv];
#pragma unroll
    for (int k = 0; k < 6; ++k) win[k] = win[k + 1];
    win[6] = hm;
#pragma unroll
    for (int k = 0; k < 3; ++k) xr[k] = xr[k + 1];
    xr[3] = centre;
    const int o = r - 3;
    if (o >= 0) {  // uniform across the CTA
      int4 vm = win[0];
#pragma unroll
      for (int k = 1; k < 7; ++k) vm = max8<T>(vm, win[k]);
      float pooled[8];
      unpack8<T>(vm, pooled);
      if (!col_ok) {
#pragma unroll
        for (int k = 0; k < 8; ++k) pooled[k] = 0.0f;
      }
      float sq[R];
#pragma unroll
      for (int j = 0; j < R; ++j) sq[j] = 0.0f;
      const float4* w1v = reinterpret_cast<const float4*>(w1s + cv * WS);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int jj = 0; jj < R / 4; ++jj) {
          const float4 wv = w1v[k * (R / 4) + jj];
          sq[jj * 4 + 0] = fmaf(pooled[k], wv.x, sq[jj * 4 + 0]);
          sq[jj * 4 + 1] = fmaf(pooled[k], wv.y, sq[jj * 4 + 1]);
          sq[jj * 4 + 2] = fmaf(pooled[k], wv.z, sq[jj * 4 + 2]);
          sq[jj * 4 + 3] = fmaf(pooled[k], wv.w, sq[jj * 4 + 3]);
        }
      }
#pragma unroll
      for (int off = CV / 2; off >= 1; off >>= 1)
#pragma unroll
        for (int j = 0; j < R; ++j) sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], off);
#pragma unroll
      for (int j = 0; j < R; ++j) sq[j] = fmaxf(sq[j] + b1s[j], 0.0f);
      if (col_ok) {
        float x[8];
        unpack8<T>(xr[0], x);
        const float4* w2v = reinterpret_cast<const float4*>(w2s + cv * WS);
        const float4 bb0 = *reinterpret_cast<const float4*>(b2s + cv * 8), bb1 = *reinterpret_cast<const float4*>(b2s + cv * 8 + 4);
        float e[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const float4 wa = w2v[j * 2], wb = w2v[j * 2 + 1];
          e[0] = fmaf(sq[j], wa.x, e[0]); e[1] = fmaf(sq[j], wa.y, e[1]); e[2] = fmaf(sq[j], wa.z, e[2]); e[3] = fmaf(sq[j], wa.w, e[3]);
          e[4] = fmaf(sq[j], wb.x, e[4]); e[5] = fmaf(sq[j], wb.y, e[5]); e[6] = fmaf(sq[j], wb.z, e[6]); e[7] = fmaf(sq[j], wb.w, e[7]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = __fdividef(x[k], 1.0f + __expf(-e[k]));  // sigmoid gate (MUFU ex2 + rcp)
        oimg[((int64_t)o * W + (w0 + c)) * CV + cv] = pack8<T>(x);
      }
    }
  }
}

template <typename T>
int launch_cam(const T* in, T* out, const CamParams& p, int B, int H, int W, cudaStream_t s) {
  if (B == 0) return PCLS_OK;
  PCLS_REQUIRE(p.C == 64 || p.C == 128, "CAM: channels must be 64 or 128, got %d", p.C);
  PCLS_REQUIRE(p.R == p.C / 16, "CAM: reduced channels must be C/16");
  const int TW = 256 / (p.C / 8);
  dim3 grid((unsigned)ceil_div(W, TW), (unsigned)B);
  if (p.C == 64)
    cam_kernel<T, 64><<<grid, 256, 0, s>>>(reinterpret_cast<const int4*>(in), reinterpret_cast<int4*>(out), p, H, W);
  else
    cam_kernel<T, 128><<<grid, 256, 0, s>>>(reinterpret_cast<const int4*>(in), reinterpret_cast<int4*>(out), p, H, W);
  return check_launch("cam_kernel");
}
template int launch_cam<__half>(const __half*, __half*, const CamParams&, int, int, int, cudaStream_t);
template int launch_cam<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, const CamParams&, int, int, int, cudaStream_t);

// --------------------------------------------------------------------------------------------------
template <typename T>
__global__ void tensor_to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = to_f32<T>(in[i]);
}
template <typename T>
int launch_tensor_to_f32(const T* in, float* out, int64_t n, cudaStream_t s) {
  if (n == 0) return PCLS_OK;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 65535) blocks = 65535;
  tensor_to_f32_kernel<T><<<(int)blocks, 256, 0, s>>>(in, out, n);
  return check_launch("tensor_to_f32_kernel");
}
template int launch_tensor_to_f32<__half>(const __half*, float*, int64_t, cudaStream_t);
template int launch_tensor_to_f32<__nv_bfloat16>(const __nv_bfloat16*, float*, int64_t, cudaStream_t);

}  // namespace pcls
