// Microbenchmark: does mixing tcgen05.mma shapes (N = 256 / N = 128) or accumulator column offsets inside one
// dependent stream cost anything?  TS mode (A in tensor memory), one CTA per SM, one issuing thread; a "layer" is a
// pattern of MMAs of 2048 ideal cycles; the accumulator region ping-pongs per layer like mlp_tc_kernel's.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_mix tools/mma_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
struct Pattern { int n; int nn[48]; int dcol[48]; };

__global__ void __launch_bounds__(128, 1) mix_kernel(const __grid_constant__ Pattern pat, int layers, long long* cycles) {
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bar;
  extern __shared__ __align__(1024) uint8_t ops[];   // 32 KB B
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ops)[i] = 0;
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
    const long long t0 = clock64();
    for (int l = 0; l < layers; ++l) {
      const uint32_t dreg = tmem + (uint32_t)(l & 1) * 256u, areg = tmem + (uint32_t)((l & 1) ^ 1) * 256u;
      for (int i = 0; i < pat.n; ++i) {
        const int N = pat.nn[i];
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t bd = make_desc(smem_u32(ops) + (pat.dcol[i] ? 16384 : 0) + (i & 3) * 32);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(dreg + (uint32_t)pat.dcol[i]), "r"(areg + (uint32_t)((i & 15) * 16)), "l"(bd), "r"(idesc), "r"(1));
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

void run(const char* name, const Pattern& p) {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  const int layers = 512;
  mix_kernel<<<148, 128, 32768>>>(p, layers, cyc);
  mix_kernel<<<148, 128, 32768>>>(p, layers, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double ideal = 0;
  for (int i = 0; i < p.n; ++i) ideal += p.nn[i] / 2;
  printf("%-44s %s  %.0f cycles/layer (ideal %.0f)\n", name, cudaGetErrorString(e), (double)h[0] / layers, ideal);
  cudaFree(cyc);
}

static void add(Pattern& p, int count, int n, int dcol) { for (int i = 0; i < count; ++i) { p.nn[p.n] = n; p.dcol[p.n] = dcol; ++p.n; } }

int main() {
  { Pattern p{}; add(p, 16, 256, 0); run("16 x N256", p); }
  { Pattern p{}; add(p, 8, 256, 0); add(p, 8, 128, 0); add(p, 8, 128, 128); run("8 x N256, 8 x N128 h0, 8 x N128 h1", p); }
  { Pattern p{}; add(p, 12, 256, 0); add(p, 4, 128, 0); add(p, 4, 128, 128); run("12 x N256, 4 x N128 h0, 4 x N128 h1", p); }
  { Pattern p{}; add(p, 16, 128, 0); add(p, 16, 128, 128); run("16 x N128 h0, 16 x N128 h1", p); }
  { Pattern p{}; for (int i = 0; i < 16; ++i) { add(p, 1, 128, 0); add(p, 1, 128, 128); } run("16 x (N128 h0, N128 h1) alternating", p); }
  { Pattern p{}; for (int i = 0; i < 2; ++i) { add(p, 4, 128, 0); add(p, 4, 128, 128); } add(p, 8, 128, 0); add(p, 8, 128, 128);
    run("2 x (4 h0, 4 h1), 8 h0, 8 h1, all N128", p); }
  { Pattern p{}; add(p, 8, 256, 0); add(p, 8, 192, 0); add(p, 8, 64, 192); run("8 x N256, 8 x N192, 8 x N64 @192", p); }
  return 0;
}
