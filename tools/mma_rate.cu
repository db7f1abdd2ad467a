// Microbenchmark: issue rate of tcgen05.mma kind::f16 (M=128, K=16) for N in {64,128,256}, A from shared
// memory (SS) or tensor memory (TS).  One CTA per SM, one issuing thread, operands zero.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n_mma, long long* cycles) {
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bar;
  extern __shared__ __align__(1024) uint8_t ops[];   // 16 KB A + 32 KB B
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ops)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 32) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint64_t bd = make_desc(smem_u32(ops) + 16384 + (i & 3) * 32);
      if (TS) {
        const uint32_t a = tmem + 256 + (i & 15) * 16;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(tmem + (uint32_t)((i & 1) * (N == 256 ? 0 : 128))), "r"(a), "l"(bd), "r"(idesc), "r"(1));
      } else {
        const uint64_t ad = make_desc(smem_u32(ops) + (i & 3) * 32);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem + (uint32_t)((i & 1) * (N == 256 ? 0 : 128))), "l"(ad), "l"(bd), "r"(idesc), "r"(1));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}"
                 ::"r"(smem_u32(&bar)), "r"(0));
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, bool TS>
void run(const char* name) {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  auto k = rate_kernel<N, TS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  const int n = 8192;
  k<<<148, 128, 49152>>>(n, cyc);
  k<<<148, 128, 49152>>>(n, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-14s %s  %.1f cycles/MMA  -> %.0f MAC/cycle/SM\n", name, cudaGetErrorString(e), (double)h[0] / n,
         128.0 * N * 16 * n / (double)h[0]);
  cudaFree(cyc);
}

int main() {
  run<256, false>("N=256 SS");
  run<256, true>("N=256 TS");
  run<128, false>("N=128 SS");
  run<128, true>("N=128 TS");
  run<64, false>("N=64  SS");
  run<64, true>("N=64  TS");
  return 0;
}
