// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16, SS mode) as a function of N, of the
// swizzle mode of the operands (row width 32 / 64 / 128 bytes = K chunk of 16 / 32 / 64 channels) and of how the
// accumulator is used (one accumulator vs a rotation).  Answers "what does a small-N MMA cost" for conv_tc.cu's logits
// layer (N = 32, 36 MMAs per tile) and for the Cin = 48 layers (32-byte rows).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/umma_bench tools/umma_bench.cu && tools/_build/umma_bench
//
// Operands are whatever the (zero-initialised) shared memory holds: timing does not depend on the values.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma_elect(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint32_t bar) {
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
               "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\nW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\nD:\n\t}"
               ::"r"(bar), "r"(parity) : "memory");
}

struct Cfg { int N, KC, n_mma, n_acc, a_mode; };   // a_mode 0: A addresses walk like conv_tc (K steps inside one tile, 3 sub-rows); 1: same A every time

__global__ void __launch_bounds__(64, 1) umma_bench(const Cfg* cfgs, int n_cfg, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    uint32_t parity = 0;
    for (int c = 0; c < n_cfg; ++c) {
      const Cfg cf = cfgs[c];
      const uint32_t swz = (uint32_t)cf.KC * 2u;
      const uint32_t layout = swz == 128 ? 2u : swz == 64 ? 4u : 6u;
      const uint32_t desc_hi = ((8u * swz) >> 4) | (1u << 14) | (layout << 29);
      const uint32_t idesc = (1u << 4) | ((uint32_t)(cf.N >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_tile = (130u * swz + 1023u) / 1024u * 1024u;          // one A stage (130 rows)
      const uint32_t b_base = base + 4u * a_tile;
      const uint32_t b_tile = (uint32_t)cf.N * swz;
      const int ksteps = cf.KC / 16;
      int nb = (int)((196u * 1024u - 4u * a_tile) / b_tile);                 // distinct B tiles that fit
      if (nb > 6) nb = 6;
      if (nb < 1) nb = 1;
      // descriptors of 12 consecutive MMAs (one "K chunk row": 3 sub-rows x up to 4 K steps) precomputed: the timed loop
      // is nothing but UTCHMMA issue, fully unrolled
      uint64_t ad[12], bd[12];
#pragma unroll
      for (int u = 0; u < 12; ++u) {
        const int sub = (u / ksteps) % 3, j = u % ksteps, stage = (u / (3 * ksteps)) & 3;
        const uint32_t a_addr = base + (uint32_t)(cf.a_mode ? 0 : stage) * a_tile + (cf.a_mode ? 0u : (uint32_t)sub * swz);
        const uint32_t b_addr = b_base + (uint32_t)((u / ksteps) % nb) * b_tile;
        ad[u] = (((uint64_t)desc_hi << 32) | (uint64_t)(((a_addr >> 4) & 0x3FFFu) | 0x10000u)) + (uint64_t)(2 * j);
        bd[u] = (((uint64_t)desc_hi << 32) | (uint64_t)(((b_addr >> 4) & 0x3FFFu) | 0x10000u)) + (uint64_t)(2 * j);
      }
      for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms up
        __syncwarp();
        const unsigned long long t0 = clock64();
        int acc = 0;
        for (int n = 0; n < cf.n_mma; n += 12) {
          const uint32_t d = tmem + (uint32_t)acc * (uint32_t)cf.N;
#pragma unroll
          for (int u = 0; u < 12; ++u) umma_elect(d, ad[u], bd[u], idesc, 1u);
          if (++acc == cf.n_acc) acc = 0;
        }
        commit_elect(smem_u32(&bar));
        const unsigned long long t1 = clock64();
        mbar_wait(smem_u32(&bar), parity);
        parity ^= 1u;
        const unsigned long long t2 = clock64();
        if (rep == 1 && lane == 0) {
          out[(size_t)(blockIdx.x * n_cfg + c) * 2 + 0] = t1 - t0;   // issue time
          out[(size_t)(blockIdx.x * n_cfg + c) * 2 + 1] = t2 - t0;   // until the last MMA completed
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  std::vector<Cfg> cfgs;
  const int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
  const int KCs[] = {64, 32, 16};
  for (int kc : KCs)
    for (int n : Ns) {
      const int n_acc = 512 / n > 4 ? 4 : 512 / n;
      cfgs.push_back({n, kc, 1152, 1, 0});
      cfgs.push_back({n, kc, 1152, n_acc, 0});
    }
  for (int n : Ns) cfgs.push_back({n, 64, 1152, 1, 1});
  Cfg* d_cfg; unsigned long long* d_out;
  const int grid = 148;
  cudaMalloc(&d_cfg, cfgs.size() * sizeof(Cfg));
  cudaMalloc(&d_out, (size_t)grid * cfgs.size() * 16);
  cudaMemcpy(d_cfg, cfgs.data(), cfgs.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(umma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  for (int g : {1, grid}) {
    umma_bench<<<g, 64, 201 * 1024>>>(d_cfg, (int)cfgs.size(), d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<unsigned long long> h((size_t)g * cfgs.size() * 2);
    cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost);
    printf("# grid %d   (cycles per MMA: issue-side | until complete), M=128 K=16 SS\n", g);
    printf("%5s %4s %6s %6s %9s %9s\n", "N", "KC", "n_acc", "a_mode", "issue", "complete");
    for (size_t c = 0; c < cfgs.size(); ++c) {
      double a = 0, b = 0;
      for (int blk = 0; blk < g; ++blk) { a += h[((size_t)blk * cfgs.size() + c) * 2]; b += h[((size_t)blk * cfgs.size() + c) * 2 + 1]; }
      printf("%5d %4d %6d %6d %9.1f %9.1f\n", cfgs[c].N, cfgs[c].KC, cfgs[c].n_acc, cfgs[c].a_mode, a / g / cfgs[c].n_mma,
             b / g / cfgs[c].n_mma);
    }
  }
  return 0;
}
