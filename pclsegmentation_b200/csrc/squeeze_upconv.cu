// FireDeconv front end: squeeze 1x1 conv (+BN, ReLU) fused with the transposed [1,4] / stride [1,2] convolution behind
// it (sm_100a).
//
// SqueezeSegV2 FIREUP (nets/SqueezeSegV2.py:191-199): squeeze = relu(bn(conv1x1(x))); upconv = relu(conv2d_transpose(
// squeeze, [1,4], strides [1,2], SAME)); the squeeze tensor has one reader.  As two ops it makes a round trip through HBM
// and both ops are short launches over narrow tensors (16 channels = 32 bytes per pixel: the tcgen05 kernel ran them at
// 0.37-0.46 of the HBM roofline).  Here the squeeze pixel goes from the accumulator registers through a per-warp shared
// memory tile straight into the transposed conv:
//
//   * a WARP is an independent pipeline over tiles of 16 consecutive input pixels of one image row (no block-level
//     synchronisation): it loads 16 px x C channels (next tile's loads in flight while the current one is computed),
//   * squeeze: mma.sync.m16n8k16, A = the loaded registers (the K order is a free permutation of the channels, lane (g, t)
//     holds channels [64 slab + 16 t, +16) of pixels g and g + 8), B = folded weights as fragments in shared memory;
//     bias + activation, rounded to 16 bit exactly like the stand-alone op, written to the warp's [16][S] tile
//     (pixels outside the image are zeros: the transposed conv's SAME padding),
//   * transposed conv: out[2m] = sq[m] K1 + sq[m-1] K3, out[2m+1] = sq[m] K2 + sq[m+1] K0 (nets: conv2d_transpose with
//     pad_left 1): four MMAs per (K step, n tile) whose A fragments are ldmatrix reads of the tile at row offsets -1/0/+1.
//     The tile's outer rows only serve as neighbours: a tile yields 14 pixels (28 output pixels), tiles overlap by two,
//   * the 28 x S outputs are contiguous in memory: staged in shared memory and stored with 16-byte coalesced writes.
//
// HBM traffic: the squeeze input once (the two overlap pixels per tile hit L1 / L2) + the up-sampled output.
#include "nn_kernels.cuh"

namespace pcls {

template <typename T> __device__ __forceinline__ void mma16816_su(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]);
template <> __device__ __forceinline__ void mma16816_su<__half>(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <> __device__ __forceinline__ void mma16816_su<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
template <typename T> __device__ __forceinline__ uint32_t pack2_su(float lo, float hi) {
  T h[2] = {from_f32<T>(lo), from_f32<T>(hi)};
  return *reinterpret_cast<uint32_t*>(h);
}
__device__ __forceinline__ float act_su(float v, int act) {
  return act == PCLS_ACT_RELU ? fmaxf(v, 0.0f) : (act == PCLS_ACT_LEAKY ? fmaxf(v, 0.1f * v) : v);
}

constexpr int SU_WARPS = 8;     // warps per CTA
constexpr int SU_TP = 14;       // pixels a tile yields (16 computed, the outer two are neighbours only)

template <int C, int S> struct SuGeom {
  static constexpr int SLABS = C / 64, NT = S / 8, KS2 = S / 16;
  static constexpr int NLD = 4 * SLABS;                         // 16-byte loads per thread and tile
  static constexpr int W1_FRAGS = SLABS * 4 * NT;               // uint2 per lane
  static constexpr int W2_FRAGS = 4 * KS2 * NT;
  static constexpr int QP = S * 2 + 16;                         // squeeze tile row pitch (bytes): ldmatrix rows on disjoint banks
  static constexpr int OP = S * 4 + 16;                         // output staging pitch: one input pixel = 2 output pixels x S
  static constexpr int WARP_BYTES = 16 * QP + 16 * OP;
  static constexpr int SMEM = (W1_FRAGS + W2_FRAGS) * 32 * 8 + SU_WARPS * WARP_BYTES;
};

template <typename T, int C, int S>
// resident CTAs: every warp keeps two tiles of loads in registers (C / 4 registers each): 3 CTAs (<= 85 registers) for C = 64,
// 2 for C = 128, 1 for C = 256 - about 48-64 KB of loads in flight per SM in every case
__global__ void __launch_bounds__(SU_WARPS * 32, (C <= 64 ? 3 : (C <= 128 ? 2 : 1)))
squeeze_upconv_kernel(const SqueezeUpconvParams p) {
  using G = SuGeom<C, S>;
  constexpr int SLABS = G::SLABS, NT = G::NT, KS2 = G::KS2, NLD = G::NLD, QP = G::QP, OP = G::OP;
  constexpr int CV = C / 8;                                      // 16-byte vectors per input pixel
  extern __shared__ int4 su_smem[];
  uint2* const w1s = reinterpret_cast<uint2*>(su_smem);          // [slab][k step][n tile][lane]
  uint2* const w2s = w1s + G::W1_FRAGS * 32;                     // [tap][k step][n tile][lane]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  unsigned char* const wbase = reinterpret_cast<unsigned char*>(w2s + G::W2_FRAGS * 32) + warp * G::WARP_BYTES;
  unsigned char* const sq = wbase;                               // [16][QP]
  unsigned char* const st = wbase + 16 * QP;                     // [16][OP]

  // ---- weights -> B fragments in shared memory (once per CTA) ----
  {
    const T* w1 = reinterpret_cast<const T*>(p.w1);
    for (int i = threadIdx.x; i < G::W1_FRAGS * 32; i += SU_WARPS * 32) {
      const int ln = i & 31, f = i >> 5, j = f % NT, s = (f / NT) % 4, sl = f / (NT * 4);
      // K slots (2t, 2t+1 | 2t+8, 2t+9) of K step s <-> channels c0 .. c0 + 3, c0 = 64 slab + 16 t + 4 s
      w1s[i] = *reinterpret_cast<const uint2*>(w1 + (size_t)(8 * j + (ln >> 2)) * p.w1_stride + 64 * sl + 16 * (ln & 3) + 4 * s);
    }
    const T* w2 = reinterpret_cast<const T*>(p.w2);
    for (int i = threadIdx.x; i < G::W2_FRAGS * 32; i += SU_WARPS * 32) {
      const int ln = i & 31, f = i >> 5, j = f % NT, s2 = (f / NT) % KS2, k = f / (NT * KS2);
      const T* row = w2 + ((size_t)k * p.w2_cout_pad + 8 * j + (ln >> 2)) * p.w2_cin_pad + 16 * s2 + 2 * (ln & 3);
      w2s[i] = make_uint2(*reinterpret_cast<const uint32_t*>(row), *reinterpret_cast<const uint32_t*>(row + 8));
    }
    __syncthreads();
  }
  float b1r[NT][2], b2r[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    b1r[j][0] = p.b1[8 * j + 2 * t]; b1r[j][1] = p.b1[8 * j + 2 * t + 1];
    b2r[j][0] = p.b2[8 * j + 2 * t]; b2r[j][1] = p.b2[8 * j + 2 * t + 1];
  }
  // ldmatrix lane addresses of the three A operands (tile rows r - 1, r, r + 1; the tile's outer rows are clamped, their
  // results are never stored): row = (lane & 7) + 8 ((lane >> 3) & 1), K offset 8 (lane >> 4) elements
  uint32_t a_addr[3];
  {
    const int r = (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int rr = min(max(r + d - 1, 0), 15);
      a_addr[d] = (uint32_t)__cvta_generic_to_shared(sq) + rr * QP + 16 * (lane >> 4);
    }
  }

  pdl_trigger(p.pdl_early);   // PDL (common.cuh): weights are staged, the activations need the producing layer
  pdl_wait();
  // ---- tiles: (image row, tile of SU_TP pixels), consecutive warps take consecutive tiles ----
  const long long total = (long long)p.rows * p.tiles_per_row;
  const long long wstep = (long long)gridDim.x * SU_WARPS;
  const int4* const in = reinterpret_cast<const int4*>(p.in);
  T* const out = reinterpret_cast<T*>(p.out);
  int4 nx[NLD];
  auto load = [&](long long wt, int4 (&v)[NLD]) {
    const long long row = wt / p.tiles_per_row;
    const int p0 = (int)(wt % p.tiles_per_row) * SU_TP - 1;      // pixel of tile row 0
    const int4* base = in + row * (long long)p.W * CV + 2 * t;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int4* src = base + (long long)min(max(p0 + g + 8 * px, 0), p.W - 1) * CV;   // (clamped: replaced by zeros below)
#pragma unroll
      for (int sl = 0; sl < SLABS; ++sl) {
        v[(px * SLABS + sl) * 2] = __ldg(src + sl * 8);
        v[(px * SLABS + sl) * 2 + 1] = __ldg(src + sl * 8 + 1);
      }
    }
  };
  long long wt = (long long)blockIdx.x * SU_WARPS + warp;
  if (wt < total) load(wt, nx);
  for (; wt < total; wt += wstep) {
    int4 cur[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) cur[i] = nx[i];
    if (wt + wstep < total) load(wt + wstep, nx);
    const long long row = wt / p.tiles_per_row;
    const int q = (int)(wt % p.tiles_per_row), p0 = q * SU_TP - 1;

    // ---- squeeze: [16 px][C] x [C][S] ----
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[j][0] = 0.0f; acc[j][1] = 0.0f; acc[j][2] = 0.0f; acc[j][3] = 0.0f; }
#pragma unroll
    for (int sl = 0; sl < SLABS; ++sl) {
      const int4 &a0 = cur[sl * 2], &a1 = cur[sl * 2 + 1], &c0 = cur[(SLABS + sl) * 2], &c1 = cur[(SLABS + sl) * 2 + 1];
      const uint32_t aw[2][8] = {{(uint32_t)a0.x, (uint32_t)a0.y, (uint32_t)a0.z, (uint32_t)a0.w, (uint32_t)a1.x, (uint32_t)a1.y, (uint32_t)a1.z, (uint32_t)a1.w},
                                 {(uint32_t)c0.x, (uint32_t)c0.y, (uint32_t)c0.z, (uint32_t)c0.w, (uint32_t)c1.x, (uint32_t)c1.y, (uint32_t)c1.z, (uint32_t)c1.w}};
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint32_t a[4] = {aw[0][2 * s], aw[1][2 * s], aw[0][2 * s + 1], aw[1][2 * s + 1]};
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const uint2 bw = w1s[((sl * 4 + s) * NT + j) * 32 + lane];
          const uint32_t bfr[2] = {bw.x, bw.y};
          mma16816_su<T>(acc[j], a, bfr);
        }
      }
    }
    __syncwarp();                                               // the previous tile's ldmatrix / staging reads are done
    {
      const bool v0 = p0 + g >= 0 && p0 + g < p.W, v1 = p0 + g + 8 >= 0 && p0 + g + 8 < p.W;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const uint32_t lo = v0 ? pack2_su<T>(act_su(acc[j][0] + b1r[j][0], p.act1), act_su(acc[j][1] + b1r[j][1], p.act1)) : 0u;
        const uint32_t hi = v1 ? pack2_su<T>(act_su(acc[j][2] + b1r[j][0], p.act1), act_su(acc[j][3] + b1r[j][1], p.act1)) : 0u;
        *reinterpret_cast<uint32_t*>(sq + g * QP + (8 * j + 2 * t) * 2) = lo;
        *reinterpret_cast<uint32_t*>(sq + (g + 8) * QP + (8 * j + 2 * t) * 2) = hi;
      }
    }
    __syncwarp();

    // ---- transposed conv: even outputs E = sq[m] K1 + sq[m-1] K3, odd outputs O = sq[m] K2 + sq[m+1] K0 ----
    float ev[NT][4], od[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) { ev[j][i] = 0.0f; od[j][i] = 0.0f; }
#pragma unroll
    for (int s2 = 0; s2 < KS2; ++s2) {
      uint32_t al[4], ac[4], ar[4];
      ldmatrix_x4(al, a_addr[0] + 32 * s2);
      ldmatrix_x4(ac, a_addr[1] + 32 * s2);
      ldmatrix_x4(ar, a_addr[2] + 32 * s2);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        auto frag = [&](int k, uint32_t (&b)[2]) { const uint2 w = w2s[((k * KS2 + s2) * NT + j) * 32 + lane]; b[0] = w.x; b[1] = w.y; };
        uint32_t b[2];
        frag(1, b); mma16816_su<T>(ev[j], ac, b);
        frag(3, b); mma16816_su<T>(ev[j], al, b);
        frag(2, b); mma16816_su<T>(od[j], ac, b);
        frag(0, b); mma16816_su<T>(od[j], ar, b);
      }
    }
    // staging: tile row r holds [E(r) | O(r)] = output pixels 2m, 2m + 1 of input pixel m = p0 + r
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      unsigned char* d0 = st + g * OP + (8 * j + 2 * t) * 2;
      unsigned char* d1 = st + (g + 8) * OP + (8 * j + 2 * t) * 2;
      *reinterpret_cast<uint32_t*>(d0) = pack2_su<T>(act_su(ev[j][0] + b2r[j][0], p.act2), act_su(ev[j][1] + b2r[j][1], p.act2));
      *reinterpret_cast<uint32_t*>(d1) = pack2_su<T>(act_su(ev[j][2] + b2r[j][0], p.act2), act_su(ev[j][3] + b2r[j][1], p.act2));
      *reinterpret_cast<uint32_t*>(d0 + S * 2) = pack2_su<T>(act_su(od[j][0] + b2r[j][0], p.act2), act_su(od[j][1] + b2r[j][1], p.act2));
      *reinterpret_cast<uint32_t*>(d1 + S * 2) = pack2_su<T>(act_su(od[j][2] + b2r[j][0], p.act2), act_su(od[j][3] + b2r[j][1], p.act2));
    }
    __syncwarp();
    // rows 1 .. n_out of the staging tile = n_out * 4 S contiguous bytes of the output row
    {
      constexpr int CPR = S * 4 / 16;                           // 16-byte chunks per input pixel
      const int n_out = min(SU_TP, p.W - q * SU_TP);
      int4* dst = reinterpret_cast<int4*>(out + ((row * p.W + (long long)q * SU_TP) * 2) * S);
#pragma unroll
      for (int c = lane; c < SU_TP * CPR; c += 32)
        if (c < n_out * CPR)
          dst[c] = *reinterpret_cast<const int4*>(st + (1 + c / CPR) * OP + (c % CPR) * 16);
    }
  }
}

template <typename T, int C, int S>
static int launch_su(const SqueezeUpconvParams& p, int B, cudaStream_t s) {
  using G = SuGeom<C, S>;
  SqueezeUpconvParams q = p;
  q.pdl_early = pdl_early_now;
  q.rows = B * p.H;
  q.tiles_per_row = (int)ceil_div(p.W, SU_TP);
  auto kern = squeeze_upconv_kernel<T, C, S>;
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    if (G::SMEM > 48 * 1024) PCLS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
    PCLS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, SU_WARPS * 32, G::SMEM));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  const long long total = (long long)q.rows * q.tiles_per_row;
  long long grid = (long long)sm_count() * ctas_per_sm;
  if (grid > ceil_div(total, SU_WARPS)) grid = ceil_div(total, SU_WARPS);
  if (grid < 1) grid = 1;
  PCLS_CHECK_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(SU_WARPS * 32), (size_t)G::SMEM, s, q));
  return check_launch("squeeze_upconv_kernel");
}

bool squeeze_upconv_supported(int C, int S) {
  return (C == 64 && S == 16) || (C == 128 && S == 16) || (C == 256 && S == 32);
}

template <typename T>
int launch_squeeze_upconv(const SqueezeUpconvParams& p, int C, int S, int B, cudaStream_t s) {
  if (B == 0) return PCLS_OK;
  if (C == 64 && S == 16) return launch_su<T, 64, 16>(p, B, s);
  if (C == 128 && S == 16) return launch_su<T, 128, 16>(p, B, s);
  if (C == 256 && S == 32) return launch_su<T, 256, 32>(p, B, s);
  set_error("squeeze_upconv: unsupported shape C=%d S=%d", C, S);
  return PCLS_ERR_INVALID;
}
template int launch_squeeze_upconv<__half>(const SqueezeUpconvParams&, int, int, int, cudaStream_t);
template int launch_squeeze_upconv<__nv_bfloat16>(const SqueezeUpconvParams&, int, int, int, cudaStream_t);

}  // namespace pcls
