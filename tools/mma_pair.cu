// Microbenchmark: SUSTAINED tcgen05.mma throughput under the power cap, one CTA per SM (cta_group::1, M = 128) against
// CTA pairs (cta_group::2, M = 256, each CTA holding half of B), both N = 256, K = 16, kind::f16, A in tensor memory
// (TS mode), random operands (zeros would not toggle the datapath), ~1.5 s per run.  Prints TFLOP/s and the average SM
// clock.  Question it answers: how much tensor throughput does halving the per-SM shared-memory operand traffic buy when
// the chip is power-limited?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_pair tools/mma_pair.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// two fp16 values uniform in (-2^-8, 2^-8)
__device__ __forceinline__ uint32_t rand_h2(uint32_t seed) {
  const uint32_t h = hash32(seed);
  const float a = ((int)(h & 0xffff) - 32768) * (1.f / 32768.f / 256.f), b = ((int)(h >> 16) - 32768) * (1.f / 32768.f / 256.f);
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int CG>
__global__ void __launch_bounds__(160, 1) pair_kernel(long long n_mma, long long* cycles, int b_bytes_per_cta) {
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bar;
  extern __shared__ __align__(1024) uint8_t ops[];   // B operand: 4 K-steps x (N or N/2) rows x 32 B, swizzled layout (contents random)
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < b_bytes_per_cta / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(ops)[i] = rand_h2(blockIdx.x * 65536u + i);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  // A operand: 128 packed columns at tmem + 256 (the four warps write their lane quarters), accumulators zeroed
  if (threadIdx.x < 128) {
    const uint32_t lane_base = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
    for (int c = 0; c < 128; c += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) r[j] = rand_h2(0x1234567u + threadIdx.x * 977u + (c + j) * 131u + blockIdx.x * 7919u);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tmem + lane_base + 256u + c),
                   "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
    }
    for (int c = 0; c < 256; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tmem + lane_base + c), "r"(0));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 128 && rank == 0) {
    const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | (((128u * CG) >> 4) << 24);
    const uint32_t sb = smem_u32(ops);
    const long long t0 = clock64();
    for (long long i = 0; i < n_mma; ++i) {
      const uint64_t bd = make_desc(sb + (uint32_t)(i & 3) * 32u);
      const uint32_t a = tmem + 256u + (uint32_t)(i & 15) * 8u;
      if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(tmem), "r"(a), "l"(bd), "r"(idesc), "r"(1));
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(tmem), "r"(a), "l"(bd), "r"(idesc), "r"(1));
    }
    if (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3));
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0));
    cycles[blockIdx.x] = clock64() - t0;
  }
  if (CG == 2 && rank == 1 && threadIdx.x == 128)   // the peer waits for the pair's work too (its TMEM is in use until then)
    asm volatile("{\n\t.reg .pred P1;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(&bar)), "r"(0));
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

template <int CG>
void run(const char* name, long long n_mma) {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaMemset(cyc, 0, 148 * 8);
  auto k = pair_kernel<CG>;
  const int b_bytes = 32768 / CG;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(160); cfg.dynamicSmemBytes = 32768;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, k, (long long)4096, cyc, b_bytes);          // warm-up
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, k, n_mma, cyc, b_bytes);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double issuers = 148.0 / CG;
  const double flop = 2.0 * (128.0 * CG) * 256.0 * 16.0 * (double)n_mma * issuers;
  printf("%-34s %s  %.1f cycles/MMA  %.0f TFLOP/s sustained over %.2f s  avg SM clock %.0f MHz\n", name, cudaGetErrorString(e),
         (double)h[0] / n_mma, flop / (ms * 1e-3) / 1e12, ms * 1e-3, (double)h[0] / (ms * 1e3));
  cudaFree(cyc);
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 16000000LL;
  run<1>("cta_group::1  M=128 N=256 TS", n);
  run<2>("cta_group::2  M=256 N=256 TS", n);
  run<1>("cta_group::1  M=128 N=256 TS", n);
  run<2>("cta_group::2  M=256 N=256 TS", n);
  return 0;
}
