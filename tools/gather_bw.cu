// Microbenchmark: achievable rate of random 4-byte gathers (half2 table entries) from an L2-resident table of
// the hash grid's size (24.5 MB) -- the denominator for the hash-grid encoding's "fraction of achievable
// memory bandwidth".  Patterns: fully random entries, and random PAIRS of adjacent entries (the x / x+1 corners
// of the trilinear stencil share a 32-byte sector).  8 independent loads in flight per thread, like the kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_bw tools/gather_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <bool PAIRS>
__global__ void gather_kernel(const uint32_t* __restrict__ table, uint32_t n_entries, int iters, uint32_t* __restrict__ sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0, s = mix(tid * 2654435761u + 12345u);
  for (int it = 0; it < iters; ++it) {
    uint32_t idx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (PAIRS) {
        if ((c & 1) == 0) { s = mix(s + 0x9e3779b9u); idx[c] = s % (n_entries - 1); }
        else idx[c] = idx[c - 1] + 1;
      } else {
        s = mix(s + 0x9e3779b9u);
        idx[c] = s % n_entries;
      }
    }
    uint32_t v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = __ldg(table + idx[c]);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc ^= v[c];
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <bool PAIRS>
void run(const char* name, const uint32_t* table, uint32_t n, uint32_t* sink) {
  const int iters = 16, threads = 1 << 23;   // 16 levels x 8 corners per "sample"
  gather_kernel<PAIRS><<<threads / 128, 128>>>(table, n, iters, sink);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) gather_kernel<PAIRS><<<threads / 128, 128>>>(table, n, iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  const double loads = (double)threads * iters * 8;
  printf("%-28s %.1f G gathers/s = %.0f GB/s of 4-byte entries (%.0f Msamples/s at 128 gathers per sample)\n", name,
         loads / ms / 1e6, loads * 4 / ms / 1e6, loads / 128 / ms / 1e3);
}

int main() {
  const uint32_t n = 6119864;   // entries of the 16-level, 2^19 hash grid (evpp_b200.ngp.level_table)
  uint32_t *table, *sink;
  cudaMalloc(&table, (size_t)n * 4);
  cudaMalloc(&sink, 4);
  cudaMemset(table, 1, (size_t)n * 4);
  run<false>("random entries", table, n, sink);
  run<true>("random adjacent pairs", table, n, sink);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
