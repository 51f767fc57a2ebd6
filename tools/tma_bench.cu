// Microbenchmark: how fast can one SM / the whole chip pull activations into shared memory with TMA tiled loads, as a
// function of the box shape and of the pixel stride in global memory?  (conv_tc.cu loads its A operand as boxes of
// {KC channels, 128 pixels}: 128 rows of KC*2 bytes, each row one pixel.)  Persistent CTAs, one producer thread, a ring of
// stages, a consumer thread that releases every stage as soon as it has landed.  Streams a tensor much larger than L2.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/tma_bench tools/tma_bench.cu -lcuda && tools/_build/tma_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\nW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\nD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// mode 0: tensor-map tiled load, box {box_c, box_rows}; tile t covers rows [t*box_rows, ..) at channel chunk (t % kchunks)
// mode 1: cp.async.bulk 1-D copies of box_c*2*box_rows contiguous bytes
__global__ void __launch_bounds__(64, 1)
tma_bench(const __grid_constant__ CUtensorMap map, const char* base, int mode, int box_c, int box_rows, int kchunks,
          long long n_tiles, int stages, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[32];
  const uint32_t tile_bytes = (uint32_t)box_c * 2u * (uint32_t)box_rows;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[16 + s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long t0 = clock64();
  if (threadIdx.x == 0) {            // producer
    int stage = 0; uint32_t phase = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(smem_u32(&bars[16 + stage]), phase ^ 1u);
      const uint32_t fb = smem_u32(&bars[stage]);
      mbar_expect_tx(fb, tile_bytes);
      const uint32_t dst = sbase + (uint32_t)stage * ((tile_bytes + 1023u) & ~1023u);
      if (mode == 0) tma_load_2d(dst, &map, fb, (int)(t % kchunks) * box_c, (int)(t / kchunks) * box_rows);
      else bulk_load_1d(dst, base + (size_t)t * tile_bytes, tile_bytes, fb);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (threadIdx.x == 32) {    // consumer: release as soon as landed
    int stage = 0; uint32_t phase = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(smem_u32(&bars[stage]), phase);
      mbar_arrive(smem_u32(&bars[16 + stage]));
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  const size_t bytes = (size_t)3 << 30;   // 3 GiB >> 126 MB L2
  char* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
  unsigned long long* d_out; cudaMalloc(&d_out, 148 * 8);
  cudaFuncSetAttribute(tma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  struct Case { const char* name; int mode, C, box_c, box_rows, swz, stages, grid; };
  const Case cases[] = {
    {"tiled  C=64   box 64x128 swz128  (16 KB, rows contiguous)", 0, 64, 64, 128, 128, 12, 148},
    {"tiled  C=64   box 64x256 swz128  (32 KB)", 0, 64, 64, 256, 128, 6, 148},
    {"tiled  C=512  box 64x128 swz128  (16 KB, row stride 1 KB)", 0, 512, 64, 128, 128, 12, 148},
    {"tiled  C=128  box 64x128 swz128  (row stride 256 B)", 0, 128, 64, 128, 128, 12, 148},
    {"tiled  C=32   box 32x128 swz64   (8 KB)", 0, 32, 32, 128, 64, 12, 148},
    {"tiled  C=32   box 32x256 swz64   (16 KB)", 0, 32, 32, 256, 64, 12, 148},
    {"tiled  C=16   box 16x128 swz32   (4 KB)", 0, 16, 16, 128, 32, 12, 148},
    {"tiled  C=16   box 16x256 swz32   (8 KB)", 0, 16, 16, 256, 32, 12, 148},
    {"tiled  C=64   box 64x128 no swizzle", 0, 64, 64, 128, 0, 12, 148},
    {"tiled  C=256  box 256x32 no swizzle (512-byte rows, 16 KB)", 0, 256, 256, 32, 0, 12, 148},
    {"tiled  C=64   box 64x128 swz128, 4 stages", 0, 64, 64, 128, 128, 4, 148},
    {"tiled  C=64   box 64x128 swz128, 1 CTA", 0, 64, 64, 128, 128, 12, 1},
    {"bulk 1-D 16 KB contiguous", 1, 64, 64, 128, 0, 12, 148},
    {"bulk 1-D 16 KB contiguous, 1 CTA", 1, 64, 64, 128, 0, 12, 1},
    {"bulk 1-D 4 KB contiguous", 1, 16, 16, 128, 0, 12, 148},
  };
  printf("%-62s %10s %12s %12s\n", "case", "GB/s", "B/clk/SM", "cyc/row");
  for (const Case& c : cases) {
    const size_t rows = bytes / ((size_t)c.C * 2);
    CUtensorMap map;
    cuuint64_t gd[2] = {(cuuint64_t)c.C, (cuuint64_t)rows};
    cuuint64_t gs[1] = {(cuuint64_t)c.C * 2};
    cuuint32_t bx[2] = {(cuuint32_t)c.box_c, (cuuint32_t)c.box_rows}, es[2] = {1, 1};
    CUtensorMapSwizzle sw = c.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : c.swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : c.swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-62s encode failed %d\n", c.name, (int)r); continue; }
    const int kchunks = c.C / c.box_c;
    long long n_tiles = (long long)(rows / c.box_rows) * kchunks;
    if (c.grid == 1) n_tiles /= 148;
    const size_t tile_bytes = (size_t)c.box_c * 2 * c.box_rows;
    const size_t smem = (size_t)c.stages * ((tile_bytes + 1023) & ~1023) + 1024;
    float best = 1e30f; double cyc = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      tma_bench<<<c.grid, 64, smem>>>(map, buf, c.mode, c.box_c, c.box_rows, kchunks, n_tiles, c.stages, d_out);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%-62s error %s\n", c.name, cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) {
        best = ms;
        std::vector<unsigned long long> h(c.grid);
        cudaMemcpy(h.data(), d_out, c.grid * 8, cudaMemcpyDeviceToHost);
        cyc = 0; for (auto v : h) cyc += (double)v / c.grid;
      }
    }
    const double total = (double)n_tiles * tile_bytes;
    printf("%-62s %10.0f %12.2f %12.2f\n", c.name, total / best / 1e6, total / c.grid / cyc,
           cyc / ((double)n_tiles / c.grid * c.box_rows));
  }
  return 0;
}
