// Planner surrogate g_phi: the small MLP the planners fit to the (location -> information gain) samples of one
// planning step and then query inside weighted A* (SURVEY.md section 8, row f3).
//   reference: class MLP                 plannerServer_Object/core/vpp.py:26-67   (3 -> 64, seven 64 -> 64 + ReLU, 64 -> 1;
//                                                                                  layer2_2 is declared but never used)
//              Viewpath_planner.train_only                  vpp.py:619-654        (300 full-batch Adam steps, lr 1e-3, MSE)
//              queries mlp(x)                                vpp.py:233,242,487; localPlanner.py:126
//
// The whole fit is ONE kernel launch: 300 epochs x (8 forward + 7 backward stages) of latency-bound 100 x 64 x 64 products,
// far too small to spread over the GPU (a grid-wide barrier per stage would cost more than the stage).  The default
// runs on one cluster of 16 (or 8) CTAs that split every layer by columns, keep parameters, Adam moments and tiles resident in
// shared memory and exchange slices through distributed shared memory with one cluster barrier per stage
// (surrogate_fit_cluster_kernel).  The single-CTA twin (surrogate_fit_kernel) does the same arithmetic on one SM with
// parameters, moments and activations in global memory (L2 resident, < 0.7 MB) and is the fallback; the two produce
// bit-identical parameters (tools/surrogate_ab.py).  Weight rows are staged with stride 65: conflict-free for both the
// forward and the transposed backward product.  Gradients are never written to global memory: they are applied by the
// Adam update of the thread (or CTA) that produced them.
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
// Every product / sum is spelled out: left to the compiler, `v * beta2 + (1 - beta2) * g * g` contracts into an FMA
// around either product depending on the surrounding code, and the two fit kernels would differ in the last bit.
__device__ __forceinline__ float adam_step(float p, float& mi, float& vi, float g, const AdamStep& a) {
  mi = fmaf(__fsub_rn(g, mi), 1.f - a.beta1, mi);                                     // exp_avg.lerp_(grad, 1 - beta1)
  vi = fmaf(__fmul_rn(1.f - a.beta2, g), g, __fmul_rn(vi, a.beta2));                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = fmaf(__fsqrt_rn(vi), a.inv_sqrt_bc2, a.eps);                    // sqrt(v) / sqrt(bias_correction2) + eps
  return fmaf(-a.lr_over_bc1, __fdiv_rn(mi, denom), p);                               // param.addcdiv_(exp_avg, denom, value=-lr / bias_correction1)
}
__device__ __forceinline__ void adam_update(float* p, float* m, float* v, int idx, float g, const AdamStep& a) {
  float mi = m[idx], vi = v[idx];
  p[idx] = adam_step(p[idx], mi, vi, g, a);
  m[idx] = mi;
  v[idx] = vi;
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
    float a = fmaf(act[r * kH + 32 + lane], w[32 + lane], __fmul_rn(act[r * kH + lane], w[lane]));   // explicit: same bits in every kernel
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
// The fit on a CLUSTER of CLU = 16 or 8 CTAs (described for 8).  CTA j owns columns 8j .. 8j+7 of every layer: it
// computes that slice of every forward activation, of every input delta, the gradients of its 8 weight rows, and
// applies their Adam updates at the end of the epoch (so every weight read of an epoch sees the previous epoch's values).
// Everything the epoch loop touches lives in shared memory: the CTA's weight rows with their Adam moments, a copy of
// the weight COLUMNS it needs for the transposed backward product, ping-pong activation and delta tiles.  Slices are
// exchanged through DISTRIBUTED SHARED MEMORY: the producer stores its 8 columns of every row straight into the tile
// of all eight CTAs (st.shared::cluster, 32-byte segments), one barrier.cluster (release / acquire) per stage makes
// them visible, and the consumer's operands are local when it wakes up -- a stage costs its FMAs, ~0.6 us of remote
// stores and one cluster barrier instead of two L2 round trips (the first version exchanged through global memory:
// 83 us per epoch).  Updated weight rows are pushed the same way into the column copies of their readers during the
// Adam step.  Global memory only receives the forward activations (fire-and-forget; the backward pass prefetches each
// layer's input tile with cp.async one stage ahead) and the loss curve; parameters are written back once at the end.
// The 3 -> 64 input layer, the 64 -> 1 output layer, the loss and the output delta are computed redundantly
// (identically) by all CTAs, which saves four exchanges.  Same arithmetic, operation by operation, as
// surrogate_fit_kernel: the two produce bit-identical parameters.
// CLU = 8 CTAs (portable cluster size) x 512 threads, or 16 CTAs (non-portable, preferred when the device can place
// one) x 256 threads: a CTA is a (columns of the slice) x (64 row groups) grid of threads either way.
constexpr int kAs = 68;                 // row stride of the activation tiles: the four rows a warp reads as 16-byte broadcasts
                                        // sit in different banks (stride 64 would be a 4-way conflict)
constexpr int kRg = 64;                 // row groups
constexpr int kRq = kSurrogateMaxN / kRg;                   // rows per thread in the (column, row group) products
__host__ __device__ constexpr int clu_threads(int clu) { return kRg * (kH / clu); }

template <int CLU>
struct __align__(16) CluSmem {
  static constexpr int kSl = kH / CLU;  // columns per CTA
  float A[2][kSurrogateMaxN * kAs];     // forward: input / output activations of the stage; backward: layer inputs
  float D[2][kSurrogateMaxN * kDs];     // deltas at the stage's output / input
  float Wf[kHid][kSl * kWs];            // own weight rows W[cs + c][k] of every hidden layer (resident, updated in place)
  float Mo[kHid][kSl * kH];             // their Adam moments
  float Vo[kHid][kSl * kH];
  float Wc[kHid][kH * kSl];             // weight columns W[c][cs + kk] of every hidden layer, stored [c][kk]; kept current by
                                        // the owners of the rows (remote stores in their Adam step)
  float dW[kHid][kSl * kH];             // gradients of the own rows, applied after the backward pass
  float bown[kHid][kSl], mb[kHid][kSl], vb[kHid][kSl], db[kHid][kSl];
  float w1[kH * 4];                     // input layer, all 64 units: [c][0..2] weight, [c][3] bias (owners push updates)
  float m1[kSl * 4], v1[kSl * 4], dW1[kSl * 4];
  float w3[kH + 1], m3[kH + 1], v3[kH + 1], g3[kH + 1];   // output layer (+ bias), a private identical copy per CTA
  float x[kSurrogateMaxN * 3], t[kSurrogateMaxN];
  float y[kSurrogateMaxN];
  float red[8];
};
static_assert(sizeof(CluSmem<8>) <= 227 * 1024 && sizeof(CluSmem<16>) <= 227 * 1024, "cluster fit: shared memory");

__device__ __forceinline__ uint32_t clu_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void clu_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t clu_map(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void clu_store(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// v -> the same location in all CTAs of the cluster (this one included)
template <int CLU>
__device__ __forceinline__ void clu_store_all(const float* p, float v) {
#pragma unroll
  for (uint32_t j = 0; j < CLU; ++j) clu_store(clu_map(p, j), v);
}
// [n][64] tile global (L2) -> shared memory, asynchronously (the caller waits with clu_prefetch_wait + a barrier)
__device__ __forceinline__ void clu_prefetch(float* dst, const float* src, int n) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  for (int idx = threadIdx.x; idx < n * (kH / 4); idx += blockDim.x)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + ((idx >> 4) * kAs + (idx & 15) * 4) * 4), "l"(src + idx * 4) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void clu_prefetch_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int CLU>
__global__ void __launch_bounds__(clu_threads(CLU), 1)
surrogate_fit_cluster_kernel(float* P, const float* __restrict__ x, const float* __restrict__ target, int n, int epochs,
                             float lr, float* act, float* __restrict__ loss_out) {
  extern __shared__ __align__(16) uint8_t sg_smem[];
  CluSmem<CLU>& s = *reinterpret_cast<CluSmem<CLU>*>(sg_smem);
  constexpr int kSl = kH / CLU, kCluThreads = clu_threads(CLU);
  static_assert(kCluThreads >= 4 * kH && kCluThreads % kH == 0 && kSl == kCluThreads / kH, "cluster fit: thread mapping");
  const int tid = threadIdx.x;
  const int cs = (int)clu_rank() * kSl;          // first column of this CTA's slice
  const int act_stride = n * kH;
  const float beta1 = 0.9f, beta2 = 0.999f;
  double b1t = 1.0, b2t = 1.0;
  // ---- one-time staging: parameters, zero moments (a fresh optimizer per fit, like train_only), inputs
  for (int i = 0; i < kHid; ++i) {
    const int j = used_layer(i);
    for (int q = tid; q < kSl * kH; q += kCluThreads) {
      s.Wf[i][(q >> 6) * kWs + (q & 63)] = P[off_l2w(j) + (cs + (q >> 6)) * kH + (q & 63)];
      s.Mo[i][q] = 0.f;
      s.Vo[i][q] = 0.f;
      s.Wc[i][q] = P[off_l2w(j) + (q / kSl) * kH + cs + (q % kSl)];
    }
    if (tid < kSl) { s.bown[i][tid] = P[off_l2b(j) + cs + tid]; s.mb[i][tid] = 0.f; s.vb[i][tid] = 0.f; }
  }
  if (tid < kH * 4) s.w1[tid] = (tid & 3) < 3 ? P[off_l1w() + (tid >> 2) * 3 + (tid & 3)] : P[off_l1b() + (tid >> 2)];
  if (tid < kSl * 4) { s.m1[tid] = 0.f; s.v1[tid] = 0.f; }
  if (tid <= kH) { s.w3[tid] = P[off_l3w() + tid]; s.m3[tid] = 0.f; s.v3[tid] = 0.f; }   // off_l3b() == off_l3w() + 64
  for (int i = tid; i < n * 3; i += kCluThreads) s.x[i] = x[i];
  for (int i = tid; i < n; i += kCluThreads) s.t[i] = target[i];
  clu_sync();                                    // every CTA of the cluster is resident before the first remote store
  for (int ep = 0; ep < epochs; ++ep) {
    // ================= forward =================
    for (int i = tid; i < n * kH; i += kCluThreads) {           // layer 1 (3 -> 64): all columns, every CTA the same
      const int r = i >> 6, c = i & 63;
      float a = s.w1[c * 4 + 3];
      a = fmaf(s.x[r * 3 + 0], s.w1[c * 4 + 0], a);
      a = fmaf(s.x[r * 3 + 1], s.w1[c * 4 + 1], a);
      a = fmaf(s.x[r * 3 + 2], s.w1[c * 4 + 2], a);
      a = fmaxf(a, 0.f);
      s.A[0][r * kAs + c] = a;
      if ((c / kSl) * kSl == cs) act[i] = a;                    // kept for the backward pass (own columns)
    }
    __syncthreads();
    for (int i = 0; i < kHid; ++i) {
      const float* Ain = s.A[i & 1];
      float* Aout = s.A[(i + 1) & 1];
      {
        const int c = tid % kSl, rg = tid / kSl;
        float acc[kRq];
#pragma unroll
        for (int q = 0; q < kRq; ++q) acc[q] = s.bown[i][c];
        const float* wrow = s.Wf[i] + c * kWs;
#pragma unroll 4
        for (int k = 0; k < kH; k += 4) {
          const float w0 = wrow[k], w1 = wrow[k + 1], w2 = wrow[k + 2], w3 = wrow[k + 3];
#pragma unroll
          for (int q = 0; q < kRq; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(Ain + (rg + kRg * q) * kAs + k);
            acc[q] = fmaf(a.x, w0, acc[q]);
            acc[q] = fmaf(a.y, w1, acc[q]);
            acc[q] = fmaf(a.z, w2, acc[q]);
            acc[q] = fmaf(a.w, w3, acc[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < kRq; ++q) {
          const int r = rg + kRg * q;
          acc[q] = fmaxf(acc[q], 0.f);
          if (r < n) clu_store_all<CLU>(Aout + r * kAs + cs + c, acc[q]);
        }
        clu_sync();
        // act[1..5] are re-read by the backward pass (6 and 7 stay in A).  Stored AFTER the barrier: its release would
        // otherwise wait for these global stores to be acknowledged by L2 in every stage
        if (i + 1 < kHid - 1) {
          float* out = act + (size_t)(i + 1) * act_stride;
#pragma unroll
          for (int q = 0; q < kRq; ++q) {
            const int r = rg + kRg * q;
            if (r < n) out[r * kH + cs + c] = acc[q];
          }
        }
      }
    }
    // output layer, loss and output delta: every CTA computes all of it from the same data
    const float* A7 = s.A[kHid & 1];
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int r = warp; r < n; r += kCluThreads / 32) {
        float a = fmaf(A7[r * kAs + 32 + lane], s.w3[32 + lane], __fmul_rn(A7[r * kAs + lane], s.w3[lane]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) s.y[r] = a + s.w3[kH];
      }
    }
    __syncthreads();
    {
      float e = 0.f;
      if (tid < n) {
        const float d = s.y[tid] - s.t[tid];
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
    // stage i reads the delta at layer i's output from D[(i + 1) & 1] and the layer's input activations from A[i & 1]
    // (act[6] is still there from the forward pass; act[i - 1] is prefetched into the other buffer during stage i)
    {
      float* D7 = s.D[kHid & 1];
      for (int i = tid; i < n * kH; i += kCluThreads) {
        const int r = i >> 6, k = i & 63;
        D7[r * kDs + k] = A7[r * kAs + k] > 0.f ? s.y[r] * s.w3[k] : 0.f;
      }
      if (tid < kH) {
        float g = 0.f;
        for (int r = 0; r < n; ++r) g = fmaf(s.y[r], A7[r * kAs + tid], g);
        s.g3[tid] = g;
      } else if (tid == kH) {
        float g = 0.f;
        for (int r = 0; r < n; ++r) g += s.y[r];
        s.g3[kH] = g;
      }
    }
    __syncthreads();
    for (int i = kHid - 1; i >= 0; --i) {
      const float* Ain = s.A[i & 1];
      const float* Din = s.D[(i + 1) & 1];
      float* Dout = s.D[i & 1];
      if (i > 0) clu_prefetch(s.A[(i - 1) & 1], act + (size_t)(i - 1) * act_stride, n);   // next stage's layer inputs
      {   // gradients of the own rows: dW[c][k] = sum_r D[r][cs + c] A[r][k], thread = (k, c); the bias gradient
          // db[c] = sum_r D[r][c] rides along (rows in order, like the single-CTA kernel)
        const int k = tid & 63, c = tid >> 6;
        float gw = 0.f, gb = 0.f;
#pragma unroll 4
        for (int r = 0; r < n; ++r) {
          const float d = Din[r * kDs + cs + c];
          gw = fmaf(d, Ain[r * kAs + k], gw);
          gb += d;
        }
        s.dW[i][c * kH + k] = gw;
        if (k == 0) s.db[i][c] = gb;
      }
      {   // own columns of the input delta: dA[r][cs + kk] = sum_c D[r][c] W[c][cs + kk], masked by the input's ReLU
        const int kk = tid % kSl, rg = tid / kSl;
        const float* Wc = s.Wc[i];
        float dn[kRq];
#pragma unroll
        for (int q = 0; q < kRq; ++q) dn[q] = 0.f;
#pragma unroll 4
        for (int c = 0; c < kH; c += 4) {
          const float w0 = Wc[c * kSl + kk], w1 = Wc[(c + 1) * kSl + kk], w2 = Wc[(c + 2) * kSl + kk], w3 = Wc[(c + 3) * kSl + kk];
#pragma unroll
          for (int q = 0; q < kRq; ++q) {
            const float4 d = *reinterpret_cast<const float4*>(Din + (rg + kRg * q) * kDs + c);
            dn[q] = fmaf(d.x, w0, dn[q]);
            dn[q] = fmaf(d.y, w1, dn[q]);
            dn[q] = fmaf(d.z, w2, dn[q]);
            dn[q] = fmaf(d.w, w3, dn[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < kRq; ++q) {
          const int r = rg + kRg * q;
          if (r < n) {
            const float v = Ain[r * kAs + cs + kk] > 0.f ? dn[q] : 0.f;
            if (i > 0) clu_store_all<CLU>(Dout + r * kDs + cs + kk, v);
            else Dout[r * kDs + cs + kk] = v;                   // the input layer's gradient only needs the own columns
          }
        }
      }
      clu_prefetch_wait();
      clu_sync();        // deltas of all CTAs (and this CTA's prefetched tile) are in place; after stage 0: every CTA
                         // is done with its weight-column copies, which the Adam step below overwrites remotely
    }
    if (tid < kSl * 4) {   // first layer, own rows
      const int c = tid >> 2, d = tid & 3;
      const float* D0 = s.D[0];
      float g = 0.f;
      if (d < 3) { for (int r = 0; r < n; ++r) g = fmaf(D0[r * kDs + cs + c], s.x[r * 3 + d], g); }
      else { for (int r = 0; r < n; ++r) g += D0[r * kDs + cs + c]; }
      s.dW1[tid] = g;
    }
    __syncthreads();
    // ================= Adam, own parameters; new values go to every CTA that reads them =================
    for (int i = 0; i < kHid; ++i) {
      for (int q = tid; q < kSl * kH; q += kCluThreads) {
        const int c = q >> 6, k = q & 63;
        // the padded row image Wf is the parameter; Mo / Vo use the dense index
        float* w = s.Wf[i] + c * kWs + k;
        float mi = s.Mo[i][q], vi = s.Vo[i][q];
        const float pn = adam_step(*w, mi, vi, s.dW[i][q], as);
        s.Mo[i][q] = mi;
        s.Vo[i][q] = vi;
        *w = pn;
        // column k of row cs + c belongs to the backward product of CTA k / kSl
        clu_store(clu_map(&s.Wc[i][(cs + c) * kSl + (k % kSl)], (uint32_t)(k / kSl)), pn);
      }
      if (tid < kSl) adam_update(s.bown[i], s.mb[i], s.vb[i], tid, s.db[i][tid], as);
    }
    if (tid < kSl * 4) {
      adam_update(s.w1 + cs * 4, s.m1, s.v1, tid, s.dW1[tid], as);
      clu_store_all<CLU>(s.w1 + cs * 4 + tid, s.w1[cs * 4 + tid]);
    }
    if (tid <= kH) adam_update(s.w3, s.m3, s.v3, tid, s.g3[tid], as);   // output layer: identical private update in every CTA
    clu_sync();          // all remote parameter stores have landed before the next epoch reads them
  }
  // ---- write the fitted parameters back: own rows of the hidden layers and of the input layer; CTA 0 the output layer
  for (int i = 0; i < kHid; ++i) {
    const int j = used_layer(i);
    for (int q = tid; q < kSl * kH; q += kCluThreads) P[off_l2w(j) + (cs + (q >> 6)) * kH + (q & 63)] = s.Wf[i][(q >> 6) * kWs + (q & 63)];
    if (tid < kSl) P[off_l2b(j) + cs + tid] = s.bown[i][tid];
  }
  if (tid < kSl * 4) {
    const int c = cs + (tid >> 2), d = tid & 3;
    P[d < 3 ? off_l1w() + c * 3 + d : off_l1b() + c] = s.w1[cs * 4 + tid];
  }
  if (tid <= kH && cs == 0) P[off_l3w() + tid] = s.w3[tid];
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

// One cluster of CLU CTAs; false when the device cannot place it (the caller falls back to a smaller shape)
template <int CLU>
static bool launch_fit_cluster(float* params, const float* x, const float* target, int n, int epochs, float lr, float* act,
                               float* loss_out, cudaStream_t st) {
  auto kern = surrogate_fit_cluster_kernel<CLU>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CluSmem<CLU>)) != cudaSuccess ||
      (CLU > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
    cudaGetLastError();
    return false;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CLU, 1, 1);
  cfg.blockDim = dim3(clu_threads(CLU), 1, 1);
  cfg.dynamicSmemBytes = sizeof(CluSmem<CLU>);
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = CLU;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess || clusters < 1) {
    cudaGetLastError();
    return false;
  }
  if (cudaLaunchKernelEx(&cfg, kern, params, x, target, n, epochs, lr, act, loss_out) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}

int launch_surrogate_fit(float* params, float* m, float* v, const float* x, const float* target, int n, int epochs, float lr,
                         float* workspace, float* loss_out, cudaStream_t st) {
  if (n <= 0 || epochs <= 0) return 0;
  // kernel attributes are per device and one process may drive several: set them per launch
  if (cudaFuncSetAttribute(surrogate_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess) return -1;
  float* act = workspace;
  float* yws = workspace + (size_t)(kHid + 1) * n * kH;
  // EVPP_SG_CLUSTER: 0 = single CTA, 8 / 16 = that cluster size; default: 16 CTAs when the device can co-schedule such a
  // cluster (non-portable size, 177 KB of shared memory per CTA), else 8, else one CTA.  Process-wide preference; the
  // kernel attributes are set per launch.
  static int want = -1;
  if (want < 0) {
    const char* e = getenv("EVPP_SG_CLUSTER");
    want = e && *e ? atoi(e) : 16;
    if (want != 0 && want != 8 && want != 16) want = 16;
  }
  if (want == 16 && launch_fit_cluster<16>(params, x, target, n, epochs, lr, act, loss_out, st)) return 1;
  if (want >= 8 && launch_fit_cluster<8>(params, x, target, n, epochs, lr, act, loss_out, st)) return 1;
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
