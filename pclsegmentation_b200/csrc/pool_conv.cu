// Max-pool 3x3 / stride [1,2] fused with the 1x1 convolution that consumes it (sm_100a).
//
// SqueezeSegV2 (nets/SqueezeSegV2.py:295-306): every max-pool output has exactly one consumer, the squeeze 1x1 conv of
// the next Fire module (pool1 -> fire2/squeeze 64->16, pool3 -> fire4/squeeze 128->32, pool5 -> fire6/squeeze 256->48).
// As separate ops the pooled tensor makes a round trip through HBM (write + read) and costs a launch; here the pooled
// pixel never leaves the registers:
//
//   * a warp owns 16 consecutive output pixels x one 64-channel slab; lane (g, t) holds channels [16 t, 16 t + 16) of
//     pixels 2 g and 2 g + 1 (two 16-byte vectors each).  Walking down the rows of its segment it takes, per input row, the
//     horizontal 3-max (columns 2 wo - pad .. + 2: six 128-bit loads per pixel) and keeps the last two row maxima in
//     registers; the vertical 3-max is the pooled pixel (-inf padding never wins, TF SAME semantics).
//   * those registers ARE the A fragments of mma.sync.m16n8k16 (the K order is a free permutation of the channels, chosen
//     so that a lane's 16 channels are its K slots over four K steps); the folded squeeze weights live in registers as B
//     fragments.  C / 64 slabs (warps) of a pixel group leave partial sums in shared memory, the first one adds them,
//     applies bias + activation and stores the 16-bit squeeze output.
//
// HBM traffic: the un-pooled tensor once (the 3x3 windows overlap inside L1 / L2) + the squeeze output.
#include "nn_kernels.cuh"

namespace pcls {

template <typename T> __device__ __forceinline__ void mma16816_pc(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]);
template <> __device__ __forceinline__ void mma16816_pc<__half>(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <> __device__ __forceinline__ void mma16816_pc<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <typename T> struct Vec2pc;
template <> struct Vec2pc<__half> { using type = __half2; };
template <> struct Vec2pc<__nv_bfloat16> { using type = __nv_bfloat162; };
template <typename T>
__device__ __forceinline__ int4 max8pc(const int4& a, const int4& b) {
  using V = typename Vec2pc<T>::type;
  int4 r;
  const V* x = reinterpret_cast<const V*>(&a);
  const V* y = reinterpret_cast<const V*>(&b);
  V* o = reinterpret_cast<V*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = __hmax2(x[i], y[i]);
  return r;
}
template <typename T> __device__ __forceinline__ int4 neg_inf8pc();
template <> __device__ __forceinline__ int4 neg_inf8pc<__half>() { return make_int4(0xFC00FC00, 0xFC00FC00, 0xFC00FC00, 0xFC00FC00); }
template <> __device__ __forceinline__ int4 neg_inf8pc<__nv_bfloat16>() { return make_int4(0xFF80FF80, 0xFF80FF80, 0xFF80FF80, 0xFF80FF80); }


template <typename T, int C, int S>
__global__ void __launch_bounds__(256, (S <= 48 ? 2 : 1))
pool_conv1x1_kernel(const PoolConvParams p) {
  constexpr int SLABS = C / 64;          // warps that share a pixel group (each contracts 64 channels)
  constexpr int PG = 8 / SLABS;          // pixel groups (16 output pixels each) per CTA
  constexpr int NT = S / 8;              // n tiles of the MMA
  constexpr int CV = C / 8;              // 16-byte vectors per input pixel
  extern __shared__ float pc_smem[];     // [2 (row parity)][PG][SLABS][16][S] partial sums (SLABS > 1)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int slab = warp % SLABS, pg = warp / SLABS;
  const int4 NEG = neg_inf8pc<T>();

  // ---- B fragments: K slot pair (2t, 2t+1) / (2t+8, 2t+9) of K step s <-> channels c0 .. c0+3, c0 = 64 slab + 16 t + 4 s.
  // One slab (C = 64): in registers.  More slabs: in shared memory in fragment order [slab][s][j][lane] (conflict-free
  // 8-byte reads) - 48-64 registers less per thread, i.e. two resident CTAs per SM instead of one.
  constexpr bool W_SMEM = SLABS > 1;
  uint32_t wf[W_SMEM ? 1 : 4][W_SMEM ? 1 : NT][2];
  uint2* const wsm = reinterpret_cast<uint2*>(pc_smem + (W_SMEM ? 2 * PG * SLABS * 16 * S : 0));
  {
    const T* w = reinterpret_cast<const T*>(p.w);
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if constexpr (W_SMEM) {
          if (pg == 0)   // the warps of pixel group 0 cover every slab once
            wsm[((slab * 4 + s) * NT + j) * 32 + lane] =
                *reinterpret_cast<const uint2*>(w + (size_t)(8 * j + g) * p.w_stride + 64 * slab + 16 * t + 4 * s);
        } else {
          const uint2 v = *reinterpret_cast<const uint2*>(w + (size_t)(8 * j + g) * p.w_stride + 64 * slab + 16 * t + 4 * s);
          wf[s][j][0] = v.x; wf[s][j][1] = v.y;
        }
      }
    if constexpr (W_SMEM) __syncthreads();
  }
  pdl_trigger(p.pdl_early);   // PDL (common.cuh): the weight fragments are loaded, the activations need the producing layer
  pdl_wait();
  const float lo = p.act == PCLS_ACT_RELU ? 0.0f : -INFINITY;
  const bool leaky = p.act == PCLS_ACT_LEAKY;
  T* const outp = reinterpret_cast<T*>(p.out);

  // Work = (strip, row) pairs, strip = (frame, tile of 16 PG output columns); the grid splits the strip-major list of rows
  // EVENLY (one wave, every CTA within one row of the same load): a CTA walks rows [r_lo, r_hi) and restarts the vertical
  // window where it enters a new strip.
  const long long total_rows = (long long)p.n_strips * p.H;
  const long long r_lo = total_rows * blockIdx.x / gridDim.x, r_hi = total_rows * (blockIdx.x + 1) / gridDim.x;
  int4 prev[2][2], cur[2][2], nxt[2][2];
  int strip = -1, wo0 = 0;
  const int4* img = nullptr;
  int64_t bframe = 0;
  // horizontal 3-max of input row h for this lane's two pixels (two 16-byte vectors each): m[pixel][vector]
  // All twelve 128-bit loads of a row are UNCONDITIONAL (clamped addresses, the out-of-image ones replaced by -inf
  // afterwards): inside `if (inside the image)` blocks the compiler kept every load next to its max and a row cost six
  // exposed memory latencies (ncu: 2.3 TB/s at 24 % issue / 24 % LSU utilisation).
  auto hrow = [&](int h, int4 (&m)[2][2]) {
    const bool hv = h >= 0 && h < p.H;
    const int4* row = img + (int64_t)min(max(h, 0), p.H - 1) * p.Win * CV;
    // the lane's two pixels are ADJACENT (wo, wo + 1): their windows share the middle one of five input columns - ten
    // 128-bit loads per row instead of twelve (ncu: the kernel's busiest unit is L1 / LSU at 70 %)
    const int wi0 = 2 * min(wo0 + 2 * g, p.Wout - 1) - p.pad_left;   // (columns right of the image: computed, never stored)
    int4 v[5][2];
    bool ok[5];
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
      const int wi = wi0 + dx;
      ok[dx] = hv && wi >= 0 && wi < p.Win;
      const int4* src = row + (int64_t)min(max(wi, 0), p.Win - 1) * CV;
      v[dx][0] = __ldg(src);
      v[dx][1] = __ldg(src + 1);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int4 mid = ok[2] ? v[2][k] : NEG;
      m[0][k] = max8pc<T>(max8pc<T>(ok[0] ? v[0][k] : NEG, ok[1] ? v[1][k] : NEG), mid);
      m[1][k] = max8pc<T>(max8pc<T>(ok[3] ? v[3][k] : NEG, ok[4] ? v[4][k] : NEG), mid);
    }
  };
  for (long long r = r_lo; r < r_hi; ++r) {
    const int st = (int)(r / p.H), h = (int)(r % p.H);
    if (st != strip) {
      strip = st;
      const int tile = st % p.tiles_per_row;
      bframe = st / p.tiles_per_row;
      wo0 = (tile * PG + pg) * 16;                      // first output pixel of this warp's group
      img = reinterpret_cast<const int4*>(p.in) + bframe * p.H * p.Win * CV + slab * 8 + t * 2;
      hrow(h - 1, prev);
      hrow(h, cur);
    }
    hrow(h + 1, nxt);
    // pooled pixel = vertical 3-max; its eight 32-bit words per pixel are the A fragments of the four K steps
    uint32_t aw[2][8];
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int4 m = max8pc<T>(max8pc<T>(prev[px][v], cur[px][v]), nxt[px][v]);
        aw[px][4 * v] = (uint32_t)m.x; aw[px][4 * v + 1] = (uint32_t)m.y; aw[px][4 * v + 2] = (uint32_t)m.z; aw[px][4 * v + 3] = (uint32_t)m.w;
      }
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[j][0] = 0.0f; acc[j][1] = 0.0f; acc[j][2] = 0.0f; acc[j][3] = 0.0f; }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const uint32_t a[4] = {aw[0][2 * s], aw[1][2 * s], aw[0][2 * s + 1], aw[1][2 * s + 1]};
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if constexpr (W_SMEM) {
          const uint2 bw = wsm[((slab * 4 + s) * NT + j) * 32 + lane];
          const uint32_t bfr[2] = {bw.x, bw.y};
          mma16816_pc<T>(acc[j], a, bfr);
        } else {
          mma16816_pc<T>(acc[j], a, wf[s][j]);
        }
      }
    }
    // (pixels right of the image hold -inf: their rows are never stored; -inf * 0 = NaN stays inside those rows)
    // epilogue of n tile j by warp j % SLABS of the pixel group: sum of the slabs' partial tiles, bias, activation, store
    auto finish = [&](int j, float (&a4)[4]) {
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const int wo = wo0 + 2 * g + px;
        if (wo < p.Wout) {
          float v0 = a4[2 * px] + __ldg(p.bias + 8 * j + 2 * t), v1 = a4[2 * px + 1] + __ldg(p.bias + 8 * j + 2 * t + 1);
          v0 = leaky ? fmaxf(v0, 0.1f * v0) : fmaxf(v0, lo);
          v1 = leaky ? fmaxf(v1, 0.1f * v1) : fmaxf(v1, lo);
          T pr[2] = {from_f32<T>(v0), from_f32<T>(v1)};
          *reinterpret_cast<uint32_t*>(outp + ((bframe * p.H + h) * p.Wout + wo) * p.out_channels + 8 * j + 2 * t) =
              *reinterpret_cast<uint32_t*>(pr);
        }
      }
    };
    if (p.zero_to > S && slab == 0) {   // padded output tensor (48 -> 64 channels): the pad channels are written as zeros
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const int wo = wo0 + 2 * g + px;
        for (int c = S + 2 * t; c < p.zero_to && wo < p.Wout; c += 8)
          *reinterpret_cast<uint32_t*>(outp + ((bframe * p.H + h) * p.Wout + wo) * p.out_channels + c) = 0u;
      }
    }
    if (SLABS == 1) {
#pragma unroll
      for (int j = 0; j < NT; ++j) finish(j, acc[j]);
    } else {
      float* red = pc_smem + (size_t)((r & 1) * PG + pg) * SLABS * 16 * S;
      float* dst = red + (size_t)slab * 16 * S;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (j % SLABS == slab) continue;               // (own tiles stay in registers)
        *reinterpret_cast<float2*>(dst + g * S + 8 * j + 2 * t) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2*>(dst + (g + 8) * S + 8 * j + 2 * t) = make_float2(acc[j][2], acc[j][3]);
      }
      // only the SLABS warps of this pixel group exchange partial sums: a named barrier per group lets the PG groups of the
      // CTA drift apart (their load latencies overlap instead of lining up behind one CTA-wide barrier per row)
      asm volatile("bar.sync %0, %1;" ::"r"(1 + pg), "r"(SLABS * 32) : "memory");
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (j % SLABS != slab) continue;
#pragma unroll
        for (int k = 0; k < SLABS; ++k) {
          if (k == slab) continue;
          const float* src = red + (size_t)k * 16 * S;
          const float2 u0 = *reinterpret_cast<const float2*>(src + g * S + 8 * j + 2 * t);
          const float2 u1 = *reinterpret_cast<const float2*>(src + (g + 8) * S + 8 * j + 2 * t);
          acc[j][0] += u0.x; acc[j][1] += u0.y; acc[j][2] += u1.x; acc[j][3] += u1.y;
        }
        finish(j, acc[j]);
      }
    }
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int v = 0; v < 2; ++v) { prev[px][v] = cur[px][v]; cur[px][v] = nxt[px][v]; }
  }
}

template <typename T, int C, int S>
static int launch_pc(const PoolConvParams& p, int B, cudaStream_t s) {
  constexpr int SLABS = C / 64, PG = 8 / SLABS;
  PoolConvParams q = p;
  q.pdl_early = pdl_early_now;
  q.tiles_per_row = (int)ceil_div(p.Wout, 16 * PG);
  q.n_strips = q.tiles_per_row * B;
  const size_t smem = SLABS > 1 ? (size_t)2 * PG * SLABS * 16 * S * sizeof(float) + (size_t)SLABS * 4 * (S / 8) * 32 * 8 : 0;
  static int ctas_per_sm = 0;
  auto kern = pool_conv1x1_kernel<T, C, S>;
  if (ctas_per_sm == 0) {
    if (smem > 48 * 1024) PCLS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PCLS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, 256, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  // one wave of CTAs, the rows of all strips split evenly between them (>= 6 rows per CTA: two halo rows per restart)
  long long grid = (long long)sm_count() * ctas_per_sm;
  const long long total_rows = (long long)q.n_strips * p.H;
  if (grid > total_rows / 6) grid = total_rows / 6;
  if (grid < 1) grid = 1;
  PCLS_CHECK_CUDA(launch_pdl(pool_conv1x1_kernel<T, C, S>, dim3((unsigned)grid), dim3(256), (size_t)smem, s, q));
  return check_launch("pool_conv1x1_kernel");
}

bool pool_conv1x1_supported(int C, int S) { return (C == 64 || C == 128 || C == 256) && (S == 16 || S == 32 || S == 48 || S == 64); }

template <typename T>
int launch_pool_conv1x1(const PoolConvParams& p, int C, int S, int B, cudaStream_t s) {
  if (B == 0) return PCLS_OK;
#define PC_CASE(c, n) if (C == c && S == n) return launch_pc<T, c, n>(p, B, s)
  PC_CASE(64, 16); PC_CASE(64, 32); PC_CASE(64, 64);
  PC_CASE(128, 16); PC_CASE(128, 32); PC_CASE(128, 64);
  PC_CASE(256, 16); PC_CASE(256, 32); PC_CASE(256, 48); PC_CASE(256, 64);
#undef PC_CASE
  set_error("pool_conv1x1: unsupported shape C=%d S=%d", C, S);
  return PCLS_ERR_INVALID;
}
template int launch_pool_conv1x1<__half>(const PoolConvParams&, int, int, int, cudaStream_t);
template int launch_pool_conv1x1<__nv_bfloat16>(const PoolConvParams&, int, int, int, cudaStream_t);

}  // namespace pcls
