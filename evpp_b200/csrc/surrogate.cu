// Planner surrogate g_phi: the small MLP the planners fit to the (location -> information gain) samples of one
// planning step and then query inside weighted A* (SURVEY.md section 8, row f3).
//   reference: class MLP                 plannerServer_Object/core/vpp.py:26-67   (3 -> 64, seven 64 -> 64 + ReLU, 64 -> 1;
//                                                                                  layer2_2 is declared but never used)
//              Viewpath_planner.train_only                  vpp.py:619-654        (300 full-batch Adam steps, lr 1e-3, MSE)
//              queries mlp(x)                                vpp.py:233,242,487; localPlanner.py:126
//
// The whole fit is ONE kernel launch: 300 epochs x (8 forward + 7 backward stages) of latency-bound 100 x 64 x 64 products,
// far too small to spread over the GPU (a grid-wide barrier per stage would cost more than the stage).  The default
// runs on one cluster of 8 CTAs that split every layer by columns and exchange slices through L2 with one cluster
// barrier per stage (surrogate_fit_cluster_kernel); the single-CTA twin (surrogate_fit_kernel) does the same arithmetic
// on one SM and is the fallback.  Parameters, Adam moments and activations live in global memory (L2 resident,
// < 0.7 MB); every stage stages its weight rows / columns (row stride 65: conflict-free for both the forward and the
// transposed backward product) and one activation tile in shared memory.  Gradients are never written to global
// memory: they are applied by the Adam update of the thread (or CTA) that produced them.
#include "surrogate.cuh"

namespace evpp {

namespace {

constexpr int kSgThreads = 1024;
constexpr int kH = 64;                 // hidden width
constexpr int kHid = 7;                // hidden 64 -> 64 layers actually used by forward()
constexpr int kWs = 65;                // padded row stride of the staged weight matrix
constexpr int kDs = 68;                // row stride of the delta tile: 16-byte aligned rows for float4 reads along c
// offsets into the packed state_dict (layer1.w [64,3], layer1.b, layer2_1..8 (w [64,64], b), layer3.w [1,64], layer3.b)
__host__ __device__ constexpr int off_l1w() { return 0; }
__host__ __device__ constexpr int off_l1b() { return 192; }
__host__ __device__ constexpr int off_l2w(int j) { return 256 + j * 4160; }          // j = 0..7 <-> layer2_{j+1}
__host__ __device__ constexpr int off_l2b(int j) { return 256 + j * 4160 + 4096; }
__host__ __device__ constexpr int off_l3w() { return 256 + 8 * 4160; }
__host__ __device__ constexpr int off_l3b() { return 256 + 8 * 4160 + 64; }
static_assert(off_l3b() + 1 == kSurrogateParams, "packed layout");
// forward() uses layer2_1, 2_3, 2_4, 2_5, 2_6, 2_7, 2_8 (vpp.py:43-62)
__device__ __forceinline__ int used_layer(int i) { return i == 0 ? 0 : i + 1; }

struct AdamStep {
  float lr_over_bc1, inv_sqrt_bc2, eps, beta1, beta2;
};

// torch.optim.Adam (single-tensor form, defaults betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad)
__device__ __forceinline__ void adam_update(float* p, float* m, float* v, int idx, float g, const AdamStep& a) {
  float mi = m[idx], vi = v[idx];
  mi = mi + (g - mi) * (1.f - a.beta1);                       // exp_avg.lerp_(grad, 1 - beta1)
  vi = vi * a.beta2 + (1.f - a.beta2) * g * g;                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  m[idx] = mi;
  v[idx] = vi;
  const float denom = sqrtf(vi) * a.inv_sqrt_bc2 + a.eps;     // sqrt(v) / sqrt(bias_correction2) + eps
  p[idx] = p[idx] - a.lr_over_bc1 * (mi / denom);             // param.addcdiv_(exp_avg, denom, value=-lr / bias_correction1)
}

struct __align__(16) Smem {
  float W[kH * kWs];          // staged weight matrix, W[c][k] at c*65 + k
  float A[kSurrogateMaxN * kH];    // input activations of the stage, A[r][k]
  float D[kSurrogateMaxN * kDs];   // deltas of the stage (backward), D[r][c] at r*68 + c
  float b[kH];
  float y[kSurrogateMaxN];
  float red[32];
};

__device__ __forceinline__ void stage_weights(Smem& s, const float* w, const float* b) {
  for (int i = threadIdx.x; i < kH * kH; i += kSgThreads) s.W[(i >> 6) * kWs + (i & 63)] = w[i];
  if (threadIdx.x < kH) s.b[threadIdx.x] = b[threadIdx.x];
}

// out[r][c] = relu(b[c] + sum_k A[r][k] W[c][k]) for r < n; thread = (c, row group of 16)
__device__ __forceinline__ void hidden_forward(const Smem& s, int n, float* out) {
  const int c = threadIdx.x & 63, rg = threadIdx.x >> 6;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = s.b[c];
  const float* wrow = s.W + c * kWs;
  // four k per step: the activation reads are 16-byte broadcasts (all lanes of a warp share the row), so a step
  // costs 8 + 4 shared-memory loads for 32 FMAs; the products are added in k order
#pragma unroll 2
  for (int k = 0; k < kH; k += 4) {
    const float w0 = wrow[k], w1 = wrow[k + 1], w2 = wrow[k + 2], w3 = wrow[k + 3];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(s.A + (rg + 16 * j) * kH + k);   // rows beyond n: stale tile data, never stored
      acc[j] = fmaf(a.x, w0, acc[j]);
      acc[j] = fmaf(a.y, w1, acc[j]);
      acc[j] = fmaf(a.z, w2, acc[j]);
      acc[j] = fmaf(a.w, w3, acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = rg + 16 * j;
    if (r < n) out[r * kH + c] = fmaxf(acc[j], 0.f);
  }
}

// first layer, 3 -> 64
__device__ __forceinline__ void first_forward(const float* P, const float* x, int n, float* out) {
  const float* w = P + off_l1w();
  const float* b = P + off_l1b();
  for (int i = threadIdx.x; i < n * kH; i += kSgThreads) {
    const int r = i >> 6, c = i & 63;
    float a = b[c];
    a = fmaf(x[r * 3 + 0], w[c * 3 + 0], a);
    a = fmaf(x[r * 3 + 1], w[c * 3 + 1], a);
    a = fmaf(x[r * 3 + 2], w[c * 3 + 2], a);
    out[i] = fmaxf(a, 0.f);
  }
}

// y[r] = b3 + sum_k A[r][k] w3[k]; one warp per row
__device__ __forceinline__ void last_forward(const float* P, const float* act, int n, float* y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* w = P + off_l3w();
  for (int r = warp; r < n; r += kSgThreads / 32) {
    float a = act[r * kH + lane] * w[lane] + act[r * kH + 32 + lane] * w[32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) y[r] = a + P[off_l3b()];
  }
}

__device__ __forceinline__ void load_tile(float* dst, const float* src, int count) {
  for (int i = threadIdx.x; i < count; i += kSgThreads) dst[i] = src[i];
}

// Full forward pass; act[l] (l = 0..7, [n][64]) are kept for the backward pass.
// (no __restrict__ / read-only loads on P and act: the fit kernel rewrites both between passes)
__device__ void forward_all(Smem& s, const float* P, const float* x, int n, float* act, float* y, int act_stride) {
  first_forward(P, x, n, act);
  __syncthreads();
  for (int i = 0; i < kHid; ++i) {
    const int j = used_layer(i);
    stage_weights(s, P + off_l2w(j), P + off_l2b(j));
    load_tile(s.A, act + (size_t)i * act_stride, n * kH);
    __syncthreads();
    hidden_forward(s, n, act + (size_t)(i + 1) * act_stride);
    __syncthreads();
  }
  last_forward(P, act + (size_t)kHid * act_stride, n, y);
  __syncthreads();
}

__global__ void __launch_bounds__(kSgThreads, 1)
surrogate_fit_kernel(float* P, float* M, float* V, const float* __restrict__ x, const float* __restrict__ target, int n,
                     int epochs, float lr, float* act, float* yws, float* __restrict__ loss_out) {
  extern __shared__ __align__(16) uint8_t sg_smem[];
  Smem& s = *reinterpret_cast<Smem*>(sg_smem);
  const int act_stride = n * kH;
  const float beta1 = 0.9f, beta2 = 0.999f;
  double b1t = 1.0, b2t = 1.0;
  for (int ep = 0; ep < epochs; ++ep) {
    forward_all(s, P, x, n, act, yws, act_stride);
    // ---- loss = mean((y - t)^2) (F.mse_loss), dy = 2 (y - t) / n
    {
      float e = 0.f;
      if (threadIdx.x < n) {
        const float d = yws[threadIdx.x] - target[threadIdx.x];
        s.y[threadIdx.x] = 2.f * d / (float)n;
        e = d * d;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = e;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (n + 31) / 32; ++w) t += s.red[w];
        if (loss_out) loss_out[ep] = t / (float)n;
      }
    }
    b1t *= (double)beta1;
    b2t *= (double)beta2;
    AdamStep as;
    as.beta1 = beta1; as.beta2 = beta2; as.eps = 1e-8f;
    as.lr_over_bc1 = (float)((double)lr / (1.0 - b1t));
    as.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - b2t));
    // ---- last layer: dW3[k] = sum_r dy[r] A7[r][k], db3 = sum_r dy[r]; delta7[r][k] = dy[r] w3[k] (A7[r][k] > 0)
    const float* a7 = act + (size_t)kHid * act_stride;
    load_tile(s.A, a7, n * kH);
    if (threadIdx.x < kH) s.b[threadIdx.x] = P[off_l3w() + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.x; i < n * kH; i += kSgThreads) {
      const int r = i >> 6, k = i & 63;
      s.D[r * kDs + k] = s.A[i] > 0.f ? s.y[r] * s.b[k] : 0.f;
    }
    if (threadIdx.x < kH) {
      float g = 0.f;
      for (int r = 0; r < n; ++r) g = fmaf(s.y[r], s.A[r * kH + threadIdx.x], g);
      adam_update(P, M, V, off_l3w() + threadIdx.x, g, as);
    } else if (threadIdx.x == kH) {
      float g = 0.f;
      for (int r = 0; r < n; ++r) g += s.y[r];
      adam_update(P, M, V, off_l3b(), g, as);
    }
    __syncthreads();
    // ---- hidden layers, last to first.  s.D holds the delta at the layer's output.
    for (int i = kHid - 1; i >= 0; --i) {
      const int j = used_layer(i);
      stage_weights(s, P + off_l2w(j), P + off_l2b(j));        // the weights BEFORE this step's update
      load_tile(s.A, act + (size_t)i * act_stride, n * kH);    // the layer's input activations
      __syncthreads();
      // dW[c][k] = sum_r D[r][c] A[r][k] (+ Adam); thread = (k, four consecutive output rows c = 4 cg .. 4 cg + 3)
      {
        const int k = threadIdx.x & 63, cg = threadIdx.x >> 6;
        float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int r = 0; r < n; ++r) {
          const float a = s.A[r * kH + k];
          const float4 d = *reinterpret_cast<const float4*>(s.D + r * kDs + 4 * cg);
          g[0] = fmaf(d.x, a, g[0]); g[1] = fmaf(d.y, a, g[1]); g[2] = fmaf(d.z, a, g[2]); g[3] = fmaf(d.w, a, g[3]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) adam_update(P, M, V, off_l2w(j) + (4 * cg + q) * kH + k, g[q], as);
        if (threadIdx.x < kH) {                                 // db[c] = sum_r D[r][c]
          float gb = 0.f;
          for (int r = 0; r < n; ++r) gb += s.D[r * kDs + threadIdx.x];
          adam_update(P, M, V, off_l2b(j) + threadIdx.x, gb, as);
        }
      }
      // delta at the layer's input: dA[r][k] = sum_c D[r][c] W[c][k], masked by the input's ReLU; thread = (k, row group)
      float dn[8];
      {
        const int k = threadIdx.x & 63, rg = threadIdx.x >> 6;
#pragma unroll
        for (int q = 0; q < 8; ++q) dn[q] = 0.f;
#pragma unroll 2
        for (int c = 0; c < kH; c += 4) {
          const float w0 = s.W[c * kWs + k], w1 = s.W[(c + 1) * kWs + k], w2 = s.W[(c + 2) * kWs + k], w3 = s.W[(c + 3) * kWs + k];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 d = *reinterpret_cast<const float4*>(s.D + (rg + 16 * q) * kDs + c);
            dn[q] = fmaf(d.x, w0, dn[q]);
            dn[q] = fmaf(d.y, w1, dn[q]);
            dn[q] = fmaf(d.z, w2, dn[q]);
            dn[q] = fmaf(d.w, w3, dn[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = rg + 16 * q;
          dn[q] = (r < n && s.A[r * kH + k] > 0.f) ? dn[q] : 0.f;
        }
      }
      __syncthreads();                                          // everyone is done reading D
      {
        const int k = threadIdx.x & 63, rg = threadIdx.x >> 6;
#pragma unroll
        for (int q = 0; q < 8; ++q) s.D[(rg + 16 * q) * kDs + k] = dn[q];
      }
      __syncthreads();
    }
    // ---- first layer: dW1[c][d] = sum_r D[r][c] x[r][d], db1[c] = sum_r D[r][c]
    if (threadIdx.x < kH * 4) {
      const int c = threadIdx.x >> 2, d = threadIdx.x & 3;
      float g = 0.f;
      if (d < 3) {
        for (int r = 0; r < n; ++r) g = fmaf(s.D[r * kDs + c], x[r * 3 + d], g);
        adam_update(P, M, V, off_l1w() + c * 3 + d, g, as);
      } else {
        for (int r = 0; r < n; ++r) g += s.D[r * kDs + c];
        adam_update(P, M, V, off_l1b() + c, g, as);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The fit on a CLUSTER of 8 CTAs.  CTA j owns columns 8j .. 8j+7 of every layer: it computes that slice of every
// forward activation, of every input delta, the gradients of its 8 weight rows, and applies their Adam updates at the
// end of the epoch (so every weight read of an epoch sees the previous epoch's values).  Slices are exchanged through
// global memory (L2) with one cluster barrier per stage -- barrier.cluster release / acquire orders the global writes,
// and the readers bypass L1 (ld.global.cg), whose lines may be stale from the previous epoch.  The 64 -> 1 output
// layer, the loss and the output delta are computed redundantly (identically) by all CTAs, which saves three barriers.
// 15 barriers and ~1.5 us of compute per stage instead of ~9 us on one SM.
constexpr int kClu = 8;
constexpr int kCluThreads = 256;
constexpr int kSl = kH / kClu;          // columns per CTA

struct __align__(16) CluSmem {
  float A[kSurrogateMaxN * kH];         // full input activations of the stage
  float D[kSurrogateMaxN * kDs];        // full deltas at the stage's output
  float Wf[kSl * kWs];                  // own weight rows W[cs + c][k]
  float Wc[kH * kSl];                   // weight columns W[c][ks + kk], stored [c][kk]
  float bown[kSl];
  float dW[kHid][kSl * kH];             // gradients of the own rows, applied after the backward pass
  float db[kHid][kSl];
  float dW1[kSl * 4];                   // first layer: [c][0..2] weight, [c][3] bias
  float w3[kH + 1], m3[kH + 1], v3[kH + 1], g3[kH + 1];   // output layer (+ bias), a private identical copy per CTA
  float y[kSurrogateMaxN];
  float red[8];
};

__device__ __forceinline__ uint32_t clu_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void clu_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// [n][64] tile L2 -> shared memory (row stride `ds` floats), bypassing L1; all of a thread's 16-byte loads are issued
// before the first store so that the copy costs one L2 latency, not one per element
__device__ __forceinline__ void clu_load(float* dst, const float* src, int n, int ds) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  const int n4 = n * (kH / 4);
  float4 v[kSurrogateMaxN * (kH / 4) / kCluThreads];            // 8
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int idx = threadIdx.x + u * kCluThreads;
    if (idx < n4) v[u] = __ldcg(s4 + idx);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int idx = threadIdx.x + u * kCluThreads;
    if (idx < n4) *reinterpret_cast<float4*>(dst + (idx >> 4) * ds + (idx & 15) * 4) = v[u];
  }
}

__global__ void __cluster_dims__(kClu, 1, 1) __launch_bounds__(kCluThreads, 1)
surrogate_fit_cluster_kernel(float* P, float* M, float* V, const float* __restrict__ x, const float* __restrict__ target,
                             int n, int epochs, float lr, float* act, float* dbuf, float* __restrict__ loss_out) {
  extern __shared__ __align__(16) uint8_t sg_smem[];
  CluSmem& s = *reinterpret_cast<CluSmem*>(sg_smem);
  const int tid = threadIdx.x;
  const int cs = (int)clu_rank() * kSl;          // first column of this CTA's slice
  const int act_stride = n * kH;
  const float beta1 = 0.9f, beta2 = 0.999f;
  double b1t = 1.0, b2t = 1.0;
  if (tid <= kH) { s.w3[tid] = P[off_l3w() + tid]; s.m3[tid] = 0.f; s.v3[tid] = 0.f; }   // off_l3b() == off_l3w() + 64
  __syncthreads();
  for (int ep = 0; ep < epochs; ++ep) {
    // ================= forward =================
    for (int i = tid; i < n * kSl; i += kCluThreads) {          // layer 1 (3 -> 64), own columns
      const int r = i >> 3, c = cs + (i & 7);
      float a = P[off_l1b() + c];
      a = fmaf(x[r * 3 + 0], P[off_l1w() + c * 3 + 0], a);
      a = fmaf(x[r * 3 + 1], P[off_l1w() + c * 3 + 1], a);
      a = fmaf(x[r * 3 + 2], P[off_l1w() + c * 3 + 2], a);
      act[r * kH + c] = fmaxf(a, 0.f);
    }
    clu_sync();
    for (int i = 0; i < kHid; ++i) {
      const int j = used_layer(i);
      clu_load(s.A, act + (size_t)i * act_stride, n, kH);
      for (int q = tid; q < kSl * kH; q += kCluThreads) s.Wf[(q >> 6) * kWs + (q & 63)] = P[off_l2w(j) + (cs + (q >> 6)) * kH + (q & 63)];
      if (tid < kSl) s.bown[tid] = P[off_l2b(j) + cs + tid];
      __syncthreads();
      {
        const int c = tid & 7, rg = tid >> 3;
        float acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = s.bown[c];
        const float* wrow = s.Wf + c * kWs;
#pragma unroll 4
        for (int k = 0; k < kH; k += 4) {
          const float w0 = wrow[k], w1 = wrow[k + 1], w2 = wrow[k + 2], w3 = wrow[k + 3];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(s.A + (rg + 32 * q) * kH + k);
            acc[q] = fmaf(a.x, w0, acc[q]);
            acc[q] = fmaf(a.y, w1, acc[q]);
            acc[q] = fmaf(a.z, w2, acc[q]);
            acc[q] = fmaf(a.w, w3, acc[q]);
          }
        }
        float* out = act + (size_t)(i + 1) * act_stride;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = rg + 32 * q;
          if (r < n) out[r * kH + cs + c] = fmaxf(acc[q], 0.f);
        }
      }
      clu_sync();
    }
    // output layer, loss and output delta: every CTA computes all of it from the same data
    clu_load(s.A, act + (size_t)kHid * act_stride, n, kH);
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int r = warp; r < n; r += kCluThreads / 32) {
        float a = s.A[r * kH + lane] * s.w3[lane] + s.A[r * kH + 32 + lane] * s.w3[32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) s.y[r] = a + s.w3[kH];
      }
    }
    __syncthreads();
    {
      float e = 0.f;
      if (tid < n) {
        const float d = s.y[tid] - target[tid];
        s.y[tid] = 2.f * d / (float)n;                          // dy
        e = d * d;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      if ((tid & 31) == 0) s.red[tid >> 5] = e;
      __syncthreads();
      if (tid == 0 && cs == 0 && loss_out) {
        float t = 0.f;
        for (int w = 0; w < (n + 31) / 32; ++w) t += s.red[w];
        loss_out[ep] = t / (float)n;
      }
    }
    b1t *= (double)beta1;
    b2t *= (double)beta2;
    AdamStep as;
    as.beta1 = beta1; as.beta2 = beta2; as.eps = 1e-8f;
    as.lr_over_bc1 = (float)((double)lr / (1.0 - b1t));
    as.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - b2t));
    // ================= backward =================
    for (int i = tid; i < n * kH; i += kCluThreads) {
      const int r = i >> 6, k = i & 63;
      s.D[r * kDs + k] = s.A[i] > 0.f ? s.y[r] * s.w3[k] : 0.f;
    }
    if (tid < kH) {
      float g = 0.f;
      for (int r = 0; r < n; ++r) g = fmaf(s.y[r], s.A[r * kH + tid], g);
      s.g3[tid] = g;
    } else if (tid == kH) {
      float g = 0.f;
      for (int r = 0; r < n; ++r) g += s.y[r];
      s.g3[kH] = g;
    }
    __syncthreads();
    for (int i = kHid - 1; i >= 0; --i) {
      const int j = used_layer(i);
      clu_load(s.A, act + (size_t)i * act_stride, n, kH);       // the layer's input activations
      for (int q = tid; q < kH * kSl; q += kCluThreads) s.Wc[q] = __ldcg(P + off_l2w(j) + (q >> 3) * kH + cs + (q & 7));
      __syncthreads();
      {   // gradients of the own rows: dW[c][k] = sum_r D[r][cs + c] A[r][k], two rows per thread
        const int k = tid & 63, cq = tid >> 6;
        float g0 = 0.f, g1 = 0.f;
#pragma unroll 4
        for (int r = 0; r < n; ++r) {
          const float a = s.A[r * kH + k];
          const float2 d = *reinterpret_cast<const float2*>(s.D + r * kDs + cs + 2 * cq);
          g0 = fmaf(d.x, a, g0);
          g1 = fmaf(d.y, a, g1);
        }
        s.dW[i][(2 * cq) * kH + k] = g0;
        s.dW[i][(2 * cq + 1) * kH + k] = g1;
        {   // bias gradients: warp w sums column cs + w over the rows
          const int w = tid >> 5, lane = tid & 31;
          float gb = 0.f;
          for (int r = lane; r < n; r += 32) gb += s.D[r * kDs + cs + w];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
          if (lane == 0) s.db[i][w] = gb;
        }
      }
      {   // own columns of the input delta: dA[r][ks + kk] = sum_c D[r][c] W[c][ks + kk], masked by the input's ReLU
        const int kk = tid & 7, rg = tid >> 3;
        float dn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int c = 0; c < kH; c += 4) {
          const float w0 = s.Wc[c * kSl + kk], w1 = s.Wc[(c + 1) * kSl + kk], w2 = s.Wc[(c + 2) * kSl + kk], w3 = s.Wc[(c + 3) * kSl + kk];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 d = *reinterpret_cast<const float4*>(s.D + (rg + 32 * q) * kDs + c);
            dn[q] = fmaf(d.x, w0, dn[q]);
            dn[q] = fmaf(d.y, w1, dn[q]);
            dn[q] = fmaf(d.z, w2, dn[q]);
            dn[q] = fmaf(d.w, w3, dn[q]);
          }
        }
        float* dst = dbuf + (size_t)(i & 1) * act_stride;       // ping-pong: the other buffer may still be read by a slower CTA
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = rg + 32 * q;
          if (r < n) dst[r * kH + cs + kk] = s.A[r * kH + cs + kk] > 0.f ? dn[q] : 0.f;
        }
      }
      clu_sync();
      clu_load(s.D, dbuf + (size_t)(i & 1) * act_stride, n, kDs);
      __syncthreads();
    }
    if (tid < kSl * 4) {   // first layer, own rows
      const int c = tid >> 2, d = tid & 3;
      float g = 0.f;
      if (d < 3) { for (int r = 0; r < n; ++r) g = fmaf(s.D[r * kDs + cs + c], x[r * 3 + d], g); }
      else { for (int r = 0; r < n; ++r) g += s.D[r * kDs + cs + c]; }
      s.dW1[tid] = g;
    }
    __syncthreads();
    // ================= Adam, own parameters =================
    for (int i = 0; i < kHid; ++i) {
      const int j = used_layer(i);
      for (int q = tid; q < kSl * kH; q += kCluThreads)
        adam_update(P, M, V, off_l2w(j) + (cs + (q >> 6)) * kH + (q & 63), s.dW[i][q], as);
      if (tid < kSl) adam_update(P, M, V, off_l2b(j) + cs + tid, s.db[i][tid], as);
    }
    if (tid < kSl * 4) {
      const int c = cs + (tid >> 2), d = tid & 3;
      adam_update(P, M, V, d < 3 ? off_l1w() + c * 3 + d : off_l1b() + c, s.dW1[tid], as);
    }
    if (tid <= kH) {       // output layer: identical private update in every CTA; CTA 0 publishes it
      adam_update(s.w3, s.m3, s.v3, tid, s.g3[tid], as);
      if (cs == 0) P[off_l3w() + tid] = s.w3[tid];
    }
    __syncthreads();
  }
  clu_sync();
}

// Batched queries g_phi(x): one CTA per 128 points, activations ping-pong in shared memory.
__global__ void __launch_bounds__(kSgThreads, 1)
surrogate_query_kernel(const float* __restrict__ P, const float* __restrict__ x, int M, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t sg_smem[];
  Smem& s = *reinterpret_cast<Smem*>(sg_smem);
  const int base = blockIdx.x * kSurrogateMaxN;
  const int n = min(kSurrogateMaxN, M - base);
  float* buf0 = s.A;
  float* buf1 = s.D;           // D is [128][68] >= [128][64]
  first_forward(P, x + (size_t)base * 3, n, buf1);
  __syncthreads();
  for (int i = 0; i < kHid; ++i) {
    const int j = used_layer(i);
    stage_weights(s, P + off_l2w(j), P + off_l2b(j));
    load_tile(buf0, buf1, n * kH);     // hidden_forward reads s.A
    __syncthreads();
    hidden_forward(s, n, buf1);
    __syncthreads();
  }
  last_forward(P, buf1, n, out + base);
}

}  // namespace

size_t surrogate_workspace_floats(int n) { return (size_t)(kHid + 1) * n * kH + kSurrogateMaxN + (size_t)2 * n * kH; }

int launch_surrogate_fit(float* params, float* m, float* v, const float* x, const float* target, int n, int epochs, float lr,
                         float* workspace, float* loss_out, cudaStream_t st) {
  if (n <= 0 || epochs <= 0) return 0;
  // kernel attributes are per device and one process may drive several: set them per launch
  if (cudaFuncSetAttribute(surrogate_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess) return -1;
  float* act = workspace;
  float* yws = workspace + (size_t)(kHid + 1) * n * kH;
  float* dbuf = yws + kSurrogateMaxN;
  static int use_cluster = -1;       // process-wide preference; the attribute itself is set per launch
  if (use_cluster < 0) {
    const char* e = getenv("EVPP_SG_CLUSTER");
    use_cluster = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (use_cluster && cudaFuncSetAttribute(surrogate_fit_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(CluSmem)) != cudaSuccess) {
    cudaGetLastError();
    use_cluster = 0;
  }
  if (use_cluster) {
    surrogate_fit_cluster_kernel<<<kClu, kCluThreads, sizeof(CluSmem), st>>>(params, m, v, x, target, n, epochs, lr, act, dbuf, loss_out);
    if (cudaPeekAtLastError() == cudaSuccess) return 1;
    cudaGetLastError();          // no 8-CTA cluster available (e.g. a partitioned device): fall back to one CTA
    use_cluster = 0;
  }
  surrogate_fit_kernel<<<1, kSgThreads, sizeof(Smem), st>>>(params, m, v, x, target, n, epochs, lr, act, yws, loss_out);
  return 1;
}

int launch_surrogate_query(const float* params, const float* x, int M, float* out, cudaStream_t st) {
  if (M <= 0) return 0;
  if (cudaFuncSetAttribute(surrogate_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess) return -1;
  surrogate_query_kernel<<<(M + kSurrogateMaxN - 1) / kSurrogateMaxN, kSgThreads, sizeof(Smem), st>>>(params, x, M, out);
  return 1;
}

}  // namespace evpp
