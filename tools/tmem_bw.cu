// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM for several shapes / warp counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_bw tools/tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>
__device__ __forceinline__ uint32_t do_ld(uint32_t taddr) {
  uint32_t acc = 0;
  if (SHAPE == 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  } else if (SHAPE == 32) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
  } else if (SHAPE == 256) {   // 16x256b.x4: 16 lanes x (4 x 8) columns -> 16 regs per thread; two calls cover 32 lanes
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  }
  return acc;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// PIPE: 0 = loads only, 1 = loads + a concurrent stream of 128x256x16 MMAs into columns 0..255 (loads read 256..511)
template <int SHAPE, int PIPE>
__global__ void __launch_bounds__(544, 1) bw_kernel(int iters, int nwarps, long long* cycles, uint32_t* sink, int n_mma, int mma_mode, int with_st, int d_col = 0, int d_alt = 128, int ld_base = -1) {
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t mma_bar;
  extern __shared__ __align__(1024) uint8_t ops[];   // 16 KB A + 32 KB B (contents irrelevant, zeroed)
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (49152 / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(ops)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  uint32_t acc = 0;
  long long t0 = clock64();
  long long t_mma = 0;
  if (PIPE == 1 && warp == 16) {
    if ((threadIdx.x & 31) == 0) {
      if (mma_mode == 0) {
        const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        for (int i = 0; i < n_mma; ++i) {
          const uint64_t ad = make_desc(smem_u32(ops) + (i & 3) * 32), bd = make_desc(smem_u32(ops) + 16384 + (i & 3) * 32);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1));
        }
      } else {
        // TS mode: N = 128 (mode 1) or 256 (mode 2); A = packed columns 256..383, D = columns 0..255
        const uint32_t n = mma_mode == 1 ? 128u : 256u;
        const uint32_t idesc = (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
        for (int i = 0; i < n_mma; ++i) {
          const uint64_t bd = make_desc(smem_u32(ops) + 16384 + (i & 3) * 32);
          const uint32_t a = tmem + 256 + (i & 7) * 16;
          const uint32_t d = tmem + (uint32_t)d_col + (mma_mode == 1 ? ((i >> 3) & 1) * (uint32_t)d_alt : 0);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(1));
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar)));
      asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}"
                   ::"r"(smem_u32(&mma_bar)), "r"(0));
      t_mma = clock64() - t0;
      cycles[148 + blockIdx.x] = t_mma;
    }
  }
  const uint32_t col_base = ld_base >= 0 ? (uint32_t)ld_base : (PIPE == 1 ? (mma_mode ? 384u : 256u) : 0u);
  const uint32_t col_mask = PIPE == 1 ? (mma_mode ? 127u : 255u) : 511u;
  if (warp < nwarps) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int cols_per = SHAPE == 32 ? 32 : (SHAPE == 16 ? 16 : 32);
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = col_base + ((uint32_t)(((it * (nwarps / 4 > 0 ? nwarps / 4 : 1) + (warp >> 2)) * cols_per)) & col_mask & (uint32_t)(512 - cols_per));
      if (SHAPE == 256) {
        acc ^= do_ld<256>(tmem + lane_base + col);
        acc ^= do_ld<256>(tmem + lane_base + (16u << 16) + col);
      } else {
        const uint32_t v = do_ld<SHAPE>(tmem + lane_base + col);
        acc ^= v;
        if (with_st) {   // store 8 packed columns back in place, like the MLP epilogue
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tmem + lane_base + col), "r"(v));
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int SHAPE, int PIPE>
void run(int nwarps, const char* name, int mma_mode = 0, int with_st = 0, int iters = 4096, int d_col = 0, int d_alt = 128, int ld_base = -1) {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 296 * 8); cudaMalloc(&sink, 4);
  cudaMemset(cyc, 0, 296 * 8);
  auto k = bw_kernel<SHAPE, PIPE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  const double bytes = (double)iters * nwarps * 32 * (SHAPE == 16 ? 16 : 32) * 4;
  // size the MMA stream to last about as long as the loads would at ~60 B/cycle (128 cycles per MMA)
  const int n_mma = PIPE ? (int)(bytes / 200.0 / (mma_mode == 1 ? 64.0 : 128.0)) : 0;
  k<<<148, 544, 49152>>>(iters, nwarps, cyc, sink, n_mma, mma_mode, with_st, d_col, d_alt, ld_base);
  k<<<148, 544, 49152>>>(iters, nwarps, cyc, sink, n_mma, mma_mode, with_st, d_col, d_alt, ld_base);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-12s %s mode=%d st=%d warps=%2d D=%d/alt%d ld@%d  %s  ld cycles=%lld -> %.1f B/cycle/SM", name, PIPE ? "+MMA" : "    ", mma_mode, with_st, nwarps, d_col, d_alt, ld_base, cudaGetErrorString(e), h[0],
         bytes / (double)h[0]);
  if (PIPE) printf("   | %d MMAs in %lld cycles = %.1f cycles/MMA", n_mma, h[148], (double)h[148] / n_mma);
  printf("\n");
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  run<16, 1>(16, "32x32b.x16", 0, 0);   // SS N=256 + loads
  run<16, 1>(16, "32x32b.x16", 2, 0);   // TS N=256 + loads
  run<16, 1>(16, "32x32b.x16", 2, 1);   // TS N=256 + loads + stores
  run<16, 1>(16, "32x32b.x16", 1, 0);   // TS N=128 + loads
  run<16, 1>(16, "32x32b.x16", 1, 1);   // TS N=128 + loads + stores
  run<16, 1>(4, "32x32b.x16", 1, 1);
  run<16, 1>(1, "32x32b.x16", 1, 1);
  // where the concurrent ld/st stream sits relative to the MMAs' accumulator columns (A is always read from 256..383)
  run<16, 1>(16, "32x32b.x16", 1, 1, 4096, 128, 0, 0);     // N=128 D=128..255, ld/st 0..127   (same 256-column region)
  run<16, 1>(16, "32x32b.x16", 1, 1, 4096, 128, 0, 384);   // N=128 D=128..255, ld/st 384..511
  run<16, 1>(16, "32x32b.x16", 1, 1, 4096, 0, 0, 384);     // N=128 D=0..127,   ld/st 384..511
  run<16, 1>(16, "32x32b.x16", 1, 1, 4096, 0, 0, 128);     // N=128 D=0..127,   ld/st 128..255
  run<16, 1>(16, "32x32b.x16", 2, 1, 4096, 0, 0, 0);       // N=256 D=0..255,   ld/st 0..127 (same columns)
  run<16, 1>(16, "32x32b.x16", 1, 0, 4096, 128, 0, 0);     // N=128 D=128..255, ld only 0..127
  return 0;
}
