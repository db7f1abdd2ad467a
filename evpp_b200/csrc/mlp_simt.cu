// fp32 (exact-mode) fused NeRF + uncertainty MLP: positional encoding + 8x256 trunk with skip +
// feature / alpha / view branch / rgb / uncertainty heads, one kernel, activations never leave
// shared memory.  Replaces Embedder.embed (run_nerf_helpers.py:22-55), run_network
// (run_nerf.py:84-113) and NeRF.forward (run_nerf_helpers.py:108-134) of the reference.
//
// This is the FFMA path used for the fp32 top-K re-score and as the numerical anchor of the
// tcgen05 path (mlp_tc.cu); it is FP32-pipe bound (roofline: fp32 FFMA peak, not tensor).
//
// Tile = 64 points per CTA pass, 256 threads, 8x8 (or 8x4) register micro-tiles.
// Weights stream through a 2-stage cp.async ring of 8 KB chunks ([kc][N] fp32, K-major rows
// pre-transposed and zero-padded on the host so every segment is a multiple of the chunk).
#include "common.cuh"

namespace evpp {

namespace {

constexpr int TM = 64;
constexpr int NT = 256;
constexpr int EMB_LD = 68;   // 63 emb + 1 zero pad + 4 skew
constexpr int H_LD = 260;    // 256 + 4 skew (conflict-free float4 row reads)
constexpr int DE_LD = 36;    // 27 dir emb + 5 zero pad + 4 skew

struct Smem {
  float emb[TM][EMB_LD];
  float h[TM][H_LD];
  float demb[TM][DE_LD];
  float wbuf[2][kSimtChunkFloats];
  float alpha[TM];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void prefetch_chunk(const float* stream, int g, float* wbuf) {
  const float* src = stream + (size_t)g * kSimtChunkFloats;
  float* dst = wbuf + (g & 1) * kSimtChunkFloats;
  const int t = threadIdx.x;
  cp_async16(dst + t * 4, src + t * 4);
  cp_async16(dst + (t + NT) * 4, src + (t + NT) * 4);
  cp_async_commit();
}

// One dense layer over the 64-row tile.  Input = concat of two shared-memory segments
// (k1 and k2 columns, each a multiple of the chunk depth), output written in place after
// a barrier.  `g` is the running chunk index in the weight stream.
template <int NOUT, bool RELU>
__device__ __forceinline__ void dense_layer(const float* in1, int ld1, int k1, const float* in2, int ld2,
                                            int k2, float* out, int ldo, const float* __restrict__ bias,
                                            const float* __restrict__ stream, int& g, float* wbuf) {
  constexpr int KC = kSimtChunkFloats / NOUT;   // 8 for N=256, 16 for N=128
  constexpr int CPT = NOUT / 32;                // cols per thread: 8 or 4
  const int tx = threadIdx.x & 31;
  const int ty = threadIdx.x >> 5;
  float acc[8][CPT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

#pragma unroll 1
  for (int seg = 0; seg < 2; ++seg) {
    const float* in = seg == 0 ? in1 : in2;
    const int ld = seg == 0 ? ld1 : ld2;
    const int kseg = seg == 0 ? k1 : k2;
#pragma unroll 1
    for (int k0 = 0; k0 < kseg; k0 += KC) {
      int gn = g + 1;
      if (gn == kSimtChunksTotal) gn = 0;
      prefetch_chunk(stream, gn, wbuf);
      cp_async_wait<1>();
      __syncthreads();
      const float* wb = wbuf + (g & 1) * kSimtChunkFloats;
#pragma unroll
      for (int k4 = 0; k4 < KC / 4; ++k4) {
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          a[i] = *reinterpret_cast<const float4*>(in + (ty * 8 + i) * ld + k0 + k4 * 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float* wr = wb + (k4 * 4 + kk) * NOUT;
          float4 w0 = *reinterpret_cast<const float4*>(wr + tx * 4);
          float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (CPT == 8) w1 = *reinterpret_cast<const float4*>(wr + 128 + tx * 4);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
            acc[i][0] = fmaf(av, w0.x, acc[i][0]);
            acc[i][1] = fmaf(av, w0.y, acc[i][1]);
            acc[i][2] = fmaf(av, w0.z, acc[i][2]);
            acc[i][3] = fmaf(av, w0.w, acc[i][3]);
            if (CPT == 8) {
              acc[i][4] = fmaf(av, w1.x, acc[i][4]);
              acc[i][5] = fmaf(av, w1.y, acc[i][5]);
              acc[i][6] = fmaf(av, w1.z, acc[i][6]);
              acc[i][7] = fmaf(av, w1.w, acc[i][7]);
            }
          }
        }
      }
      __syncthreads();   // everyone done with wbuf[g&1] (and, at layer end, with the inputs)
      g = gn;
    }
  }
  // epilogue: bias (+ReLU), in-place store
  const float4 b0 = *reinterpret_cast<const float4*>(bias + tx * 4);
  float4 b1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (CPT == 8) b1 = *reinterpret_cast<const float4*>(bias + 128 + tx * 4);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 o0 = make_float4(acc[i][0] + b0.x, acc[i][1] + b0.y, acc[i][2] + b0.z, acc[i][3] + b0.w);
    if (RELU) { o0.x = fmaxf(o0.x, 0.f); o0.y = fmaxf(o0.y, 0.f); o0.z = fmaxf(o0.z, 0.f); o0.w = fmaxf(o0.w, 0.f); }
    *reinterpret_cast<float4*>(out + (ty * 8 + i) * ldo + tx * 4) = o0;
    if (CPT == 8) {
      float4 o1 = make_float4(acc[i][4] + b1.x, acc[i][5] + b1.y, acc[i][6] + b1.z, acc[i][7] + b1.w);
      if (RELU) { o1.x = fmaxf(o1.x, 0.f); o1.y = fmaxf(o1.y, 0.f); o1.z = fmaxf(o1.z, 0.f); o1.w = fmaxf(o1.w, 0.f); }
      *reinterpret_cast<float4*>(out + (ty * 8 + i) * ldo + 128 + tx * 4) = o1;
    }
  }
  __syncthreads();
}

template <bool RAYS>
__global__ void __launch_bounds__(NT, 2)
mlp_simt_kernel(SimtWeights w, RayBatch rb, const float* __restrict__ pts, const float* __restrict__ dirs,
                long long M, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const long long ntiles = (M + TM - 1) / TM;
  int g = 0;
  if ((long long)blockIdx.x < ntiles) prefetch_chunk(w.stream, 0, &s.wbuf[0][0]);

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // ---- positional encoding of the tile (run_nerf_helpers.py:46-55) -------------------------
    {
      const int row = tid >> 2, q = tid & 3;
      const long long m = tile * TM + row;
      float p[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 1.f};
      if (m < M) {
        if (RAYS) {
          const long long ray = m / rb.S;
          const float z = rb.z[m];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            // pts = rays_o + rays_d * z (run_nerf.py:466), mul and add rounded separately
            p[c] = __fadd_rn(rb.rays_o[ray * 3 + c], __fmul_rn(rb.rays_d[ray * 3 + c], z));
            dv[c] = rb.viewdirs[ray * 3 + c];
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) { p[c] = pts[m * 3 + c]; dv[c] = dirs[m * 3 + c]; }
        }
      }
      if (q == 0) {
        s.emb[row][0] = p[0]; s.emb[row][1] = p[1]; s.emb[row][2] = p[2]; s.emb[row][63] = 0.f;
        s.demb[row][0] = dv[0]; s.demb[row][1] = dv[1]; s.demb[row][2] = dv[2];
#pragma unroll
        for (int c = 27; c < 32; ++c) s.demb[row][c] = 0.f;
      }
      for (int j = q; j < 3 * kLpos; j += 4) {
        const int c = j % 3, l = j / 3;
        float sn, cs;
        sincosf(p[c] * (float)(1 << l), &sn, &cs);
        s.emb[row][3 + l * 6 + c] = sn;
        s.emb[row][3 + l * 6 + 3 + c] = cs;
      }
      for (int j = q; j < 3 * kLdir; j += 4) {
        const int c = j % 3, l = j / 3;
        float sn, cs;
        sincosf(dv[c] * (float)(1 << l), &sn, &cs);
        s.demb[row][3 + l * 6 + c] = sn;
        s.demb[row][3 + l * 6 + 3 + c] = cs;
      }
    }
    __syncthreads();

    // ---- trunk (run_nerf_helpers.py:111-116) ---------------------------------------------------
    float* H = &s.h[0][0];
    const float* E = &s.emb[0][0];
    dense_layer<256, true>(E, EMB_LD, 64, nullptr, 0, 0, H, H_LD, w.bias + 0 * 256, w.stream, g, &s.wbuf[0][0]);
#pragma unroll 1
    for (int l = 1; l <= 4; ++l)
      dense_layer<256, true>(H, H_LD, 256, nullptr, 0, 0, H, H_LD, w.bias + l * 256, w.stream, g, &s.wbuf[0][0]);
    // skip: layer 5 consumes cat[input_pts, h]
    dense_layer<256, true>(E, EMB_LD, 64, H, H_LD, 256, H, H_LD, w.bias + 5 * 256, w.stream, g, &s.wbuf[0][0]);
    dense_layer<256, true>(H, H_LD, 256, nullptr, 0, 0, H, H_LD, w.bias + 6 * 256, w.stream, g, &s.wbuf[0][0]);
    dense_layer<256, true>(H, H_LD, 256, nullptr, 0, 0, H, H_LD, w.bias + 7 * 256, w.stream, g, &s.wbuf[0][0]);

    // ---- alpha head on h7 (:118) -----------------------------------------------------------------
    {
      const int row = tid >> 2, q = tid & 3;
      float a = 0.f;
      for (int k = q; k < 256; k += 4) a = fmaf(s.h[row][k], __ldg(w.head_w + k), a);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      if (q == 0) s.alpha[row] = a + __ldg(w.head_b + 0);
    }
    __syncthreads();
    // ---- feature (no ReLU, :119) and view branch (:120-124) ------------------------------------
    dense_layer<256, false>(H, H_LD, 256, nullptr, 0, 0, H, H_LD, w.bias + 8 * 256, w.stream, g, &s.wbuf[0][0]);
    dense_layer<128, true>(H, H_LD, 256, &s.demb[0][0], DE_LD, 32, H, H_LD, w.bias + 9 * 256, w.stream, g,
                           &s.wbuf[0][0]);
    // ---- rgb / uncertainty heads (:126-130) ------------------------------------------------------
    {
      const int row = tid >> 2, q = tid & 3;
      const float* hw = w.head_w + 256 + q * 128;   // q<3: rgb rows, q==3: uncertainty row
      float a = 0.f;
#pragma unroll 4
      for (int k = 0; k < 128; ++k) a = fmaf(s.h[row][k], __ldg(hw + k), a);
      a += __ldg(w.head_b + 1 + q);
      const long long m = tile * TM + row;
      if (m < M) {
        if (q < 3) {
          out[m * 5 + q] = a;
        } else {
          out[m * 5 + 3] = s.alpha[row];
          out[m * 5 + 4] = expf(a);   // uncertainty = exp(uncertainty_linear(h)) (:128-129)
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

int grid_for(long long M) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long ntiles = (M + TM - 1) / TM;
  long long g = 2LL * sms;
  return (int)(ntiles < g ? (ntiles > 0 ? ntiles : 1) : g);
}

}  // namespace

int launch_mlp_simt(const SimtWeights& w, const RayBatch& rb, float* raw, cudaStream_t st) {
  const long long M = (long long)rb.n * rb.S;
  if (M <= 0) return 0;
  // per launch, not once per process: the attribute is per device, and one process may drive several (nerf.py:269-273)
  cudaFuncSetAttribute(mlp_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  mlp_simt_kernel<true><<<grid_for(M), NT, sizeof(Smem), st>>>(w, rb, nullptr, nullptr, M, raw);
  return 1;
}

int launch_mlp_simt_points(const SimtWeights& w, const float* pts, const float* dirs, int64_t M, float* out,
                           cudaStream_t st) {
  if (M <= 0) return 0;
  cudaFuncSetAttribute(mlp_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  RayBatch rb{};
  mlp_simt_kernel<false><<<grid_for(M), NT, sizeof(Smem), st>>>(w, rb, pts, dirs, (long long)M, out);
  return 1;
}

}  // namespace evpp
