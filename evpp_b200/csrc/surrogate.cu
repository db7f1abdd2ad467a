// Planner surrogate g_phi: the small MLP the planners fit to the (location -> information gain) samples of one
// planning step and then query inside weighted A* (SURVEY.md section 8, row f3).
//   reference: class MLP                 plannerServer_Object/core/vpp.py:26-67   (3 -> 64, seven 64 -> 64 + ReLU, 64 -> 1;
//                                                                                  layer2_2 is declared but never used)
//              Viewpath_planner.train_only                  vpp.py:619-654        (300 full-batch Adam steps, lr 1e-3, MSE)
//              queries mlp(x)                                vpp.py:233,242,487; localPlanner.py:126
//
// The whole fit is ONE kernel launch by ONE CTA: 300 epochs x (9 forward + 9 backward stages) are latency-bound
// 100 x 64 x 64 products, far too small to spread over SMs (a grid-wide barrier per stage would cost more than the
// stage).  Parameters, Adam moments and activations live in global memory (L2 / L1 resident, < 0.7 MB); every stage
// stages one weight matrix (row stride 65: conflict-free for both the forward and the transposed backward product)
// and one activation tile in shared memory.  Gradients are never materialised: the thread that finishes dW[c][k]
// applies the Adam update of that parameter at once.
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

size_t surrogate_workspace_floats(int n) { return (size_t)(kHid + 1) * n * kH + kSurrogateMaxN; }

int launch_surrogate_fit(float* params, float* m, float* v, const float* x, const float* target, int n, int epochs, float lr,
                         float* workspace, float* loss_out, cudaStream_t st) {
  if (n <= 0 || epochs <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(surrogate_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess ||
        cudaFuncSetAttribute(surrogate_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess)
      return -1;
    configured = true;
  }
  float* act = workspace;
  float* yws = workspace + (size_t)(kHid + 1) * n * kH;
  surrogate_fit_kernel<<<1, kSgThreads, sizeof(Smem), st>>>(params, m, v, x, target, n, epochs, lr, act, yws, loss_out);
  return 1;
}

int launch_surrogate_query(const float* params, const float* x, int M, float* out, cudaStream_t st) {
  if (M <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(surrogate_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess)
      return -1;
    configured = true;
  }
  surrogate_query_kernel<<<(M + kSurrogateMaxN - 1) / kSurrogateMaxN, kSgThreads, sizeof(Smem), st>>>(params, x, M, out);
  return 1;
}

}  // namespace evpp
