// Ray generation, sampling, compositing and score reduction kernels.
// All are small next to the MLP (a ray costs 304 MFLOP of MLP and ~6 KB of traffic here); they are written for
// reference-identical arithmetic first (the serial cumprod / cumsum / searchsorted semantics of the reference's
// ATen CPU ops) and for coalesced HBM traffic second (one warp per ray, rows staged in shared memory).
#include "sampling.cuh"

namespace evpp {

namespace {

// get_rays (run_nerf_helpers.py:173-182) + viewdirs = d/||d|| (run_nerf.py:170).
// Arithmetic mirrors ATen-CPU fp32: (i-cx)/fx true division; the 3-term dot product is summed
// left to right with separately rounded products; the norm is sqrt(fma(z,z,fma(y,y,x*x))).
__global__ void get_rays_kernel(const float* __restrict__ c2w, int H, int W, float fx, float fy, float cx,
                                float cy, const int32_t* __restrict__ pix_idx, int R, long long ray_begin,
                                int n, float* __restrict__ rays_o, float* __restrict__ rays_d,
                                float* __restrict__ viewdirs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long g = ray_begin + t;
  const int v = (int)(g / R);
  const int r = (int)(g % R);
  const int p = pix_idx ? pix_idx[(long long)v * R + r] : r;
  const int j = p / W, i = p % W;
  const float* M = c2w + (long long)v * 12;
  const float d0 = __fdiv_rn(__fsub_rn((float)i, cx), fx);
  const float d1 = -__fdiv_rn(__fsub_rn((float)j, cy), fy);
  const float d2 = -1.f;
  float d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = __fmul_rn(d0, M[c * 4 + 0]);
    const float b = __fmul_rn(d1, M[c * 4 + 1]);
    const float e = __fmul_rn(d2, M[c * 4 + 2]);
    d[c] = __fadd_rn(__fadd_rn(a, b), e);
  }
  const float nrm = __fsqrt_rn(__fmaf_rn(d[2], d[2], __fmaf_rn(d[1], d[1], __fmul_rn(d[0], d[0]))));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (rays_o) rays_o[(long long)t * 3 + c] = M[c * 4 + 3];
    if (rays_d) rays_d[(long long)t * 3 + c] = d[c];
    if (viewdirs) viewdirs[(long long)t * 3 + c] = __fdiv_rn(d[c], nrm);
  }
}

// viewdirs = rays_d / ||rays_d|| for caller-provided rays (run_nerf.py:165-171)
__global__ void normalize_dirs_kernel(const float* __restrict__ rays_d, int n, float* __restrict__ viewdirs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float d0 = rays_d[(long long)t * 3 + 0], d1 = rays_d[(long long)t * 3 + 1], d2 = rays_d[(long long)t * 3 + 2];
  const float nrm = __fsqrt_rn(__fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0))));
  viewdirs[(long long)t * 3 + 0] = __fdiv_rn(d0, nrm);
  viewdirs[(long long)t * 3 + 1] = __fdiv_rn(d1, nrm);
  viewdirs[(long long)t * 3 + 2] = __fdiv_rn(d2, nrm);
}

// z_vals = near*(1-t) + far*t  (run_nerf.py:442-448)
__global__ void coarse_z_kernel(const float* __restrict__ t_vals, long long total, int S, float near_plane,
                                float far_plane, int lindisp, float* __restrict__ z) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const float t = t_vals[idx % S];
  const float omt = __fsub_rn(1.f, t);
  float v;
  if (!lindisp) {
    v = __fadd_rn(__fmul_rn(near_plane, omt), __fmul_rn(far_plane, t));
  } else {
    const float a = __fmul_rn(__fdiv_rn(1.f, near_plane), omt);
    const float b = __fmul_rn(__fdiv_rn(1.f, far_plane), t);
    v = __fdiv_rn(1.f, __fadd_rn(a, b));
  }
  z[idx] = v;
}

__device__ __forceinline__ float ray_norm(const float* d) {
  return __fsqrt_rn(__fmaf_rn(d[2], d[2], __fmaf_rn(d[1], d[1], __fmul_rn(d[0], d[0]))));
}

// ---------------------------------------------------------------------------------------------------------
// Per-ray kernels: ONE WARP PER RAY.  Global traffic is coalesced (a ray's raw / z rows are contiguous and are
// staged in the warp's shared-memory scratch); everything that is elementwise in the reference (alpha, sigmoid,
// pdf, inverse-cdf lookups, merge ranks) runs across the 32 lanes; everything the reference computes as a serial
// recurrence keeps its exact serial order so that results are bit-identical to a sequential walk:
//   - transmittance: torch.cumprod on CPU carries the product in double and rounds each output to fp32
//     (at::acc_type<float,false>) -> one lane runs the S-step double chain;
//   - the six per-ray sums (acc, depth, rgb, sum beta^2) are fp32 fused-multiply-add chains in sample order ->
//     six lanes run one chain each, in lockstep (the same FFMA instruction with per-lane operands).
constexpr int kRayWarps = 8;   // warps (= rays in flight) per CTA

// Per-warp scratch, planar so that the serial stages read four consecutive samples with one 16-byte load.
// Sp = S rounded up to a multiple of 4.
struct RayScratch {
  float* pl;    // [5][Sp]  planes 0..2 = sigmoid(rgb), 3 = alpha, then weight, 4 = beta
  float* sz;    // [Sp]     z
  double* dT;   // [Sp]     (double)((1 - alpha) + 1e-10), then the transmittance before each sample
  float* ones;  // [4]      1.0f
  int Sp;
};
__host__ __device__ constexpr int ray_sp(int S) { return (S + 3) & ~3; }
__host__ __device__ constexpr int ray_scratch_floats(int S) { return 8 * ray_sp(S) + 4; }
__host__ __device__ constexpr int resample_scratch_floats(int Sc, int Ni) {
  return ray_scratch_floats(Sc) + ((2 * Sc + 2 * Ni + 3) & ~3);   // + cdf[Sc], z_samples[Ni], merged[Sc+Ni]
}
__device__ __forceinline__ RayScratch ray_scratch(float* base, int S) {
  const int Sp = ray_sp(S);
  return RayScratch{base + 2 * Sp, base + 7 * Sp, reinterpret_cast<double*>(base), base + 8 * Sp, Sp};
}

// raw2outputs (run_nerf.py:344-388) for one ray by one warp.  On return plane 3 holds weights[i].
// SIGMA_ONLY: `raw` is the [S] row of raw sigmas a score-only coarse pass produces (mlp_tc.cu, HEADS = 1); only the
// weights are computed (all that sample_pdf needs) -- the alpha / transmittance arithmetic is the same instructions
// in the same order, so the weights are bit-identical to the full variant's.
template <bool SIGMA_ONLY>
__device__ __forceinline__ void composite_ray_warp(const float* __restrict__ raw, const float* __restrict__ z, float dnorm,
                                                   int S, int white_bkgd, long long t, const CompositeOut& co,
                                                   const RayScratch& sc, int lane) {
  const int Sp = sc.Sp;
  float* pl = sc.pl;
  float* sz = sc.sz;
  double* dT = sc.dT;
  for (int k = lane; k < S; k += 32) sz[k] = z[k];
  if (lane < 4) sc.ones[lane] = 1.f;
  __syncwarp();
  // stage 1, elementwise over the samples.  The fp32 -> fp64 conversion of the transmittance factor happens here,
  // in parallel: conversions run on the 16-lane XU pipe and cost a full warp slot each, so they must not sit in
  // the single-lane recurrence below.
#pragma unroll 2
  for (int i = lane; i < S; i += 32) {
    const float* rw = raw + i * 5;
    const float x3 = SIGMA_ONLY ? raw[i] : rw[3];
    const float zi = sz[i];
    float dist = (i + 1 < S) ? __fsub_rn(sz[i + 1], zi) : 1e10f;
    dist = __fmul_rn(dist, dnorm);
    const float sigma = fmaxf(x3, 0.f);
    const float alpha = __fsub_rn(1.f, expf(__fmul_rn(-sigma, dist)));
    dT[i] = (double)__fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
    pl[3 * Sp + i] = alpha;
    if (!SIGMA_ONLY) {
      const float x0 = rw[0], x1 = rw[1], x2 = rw[2], x4 = rw[4];
      pl[0 * Sp + i] = 1.f / (1.f + expf(-x0));
      pl[1 * Sp + i] = 1.f / (1.f + expf(-x1));
      pl[2 * Sp + i] = 1.f / (1.f + expf(-x2));
      pl[4 * Sp + i] = x4;
      if (co.unc) co.unc[t * S + i] = x4;
    }
  }
  __syncwarp();
  // stage 2, the transmittance recurrence in double (torch.cumprod's accumulation type on CPU), one lane:
  // dT[i] <- product of the factors before sample i
  if (lane == 0) {
    double T = 1.0;
    int i = 0;
    for (; i + 4 <= S; i += 4) {
      const double2 f01 = *reinterpret_cast<const double2*>(dT + i);
      const double2 f23 = *reinterpret_cast<const double2*>(dT + i + 2);
      double2 o01, o23;
      o01.x = T; T = T * f01.x;
      o01.y = T; T = T * f01.y;
      o23.x = T; T = T * f23.x;
      o23.y = T; T = T * f23.y;
      *reinterpret_cast<double2*>(dT + i) = o01;
      *reinterpret_cast<double2*>(dT + i + 2) = o23;
    }
    for (; i < S; ++i) { const double f = dT[i]; dT[i] = T; T = T * f; }
  }
  __syncwarp();
  // weights = alpha * (float)T, elementwise again (the fp64 -> fp32 rounding of every output, like cumprod's)
  for (int i = lane; i < S; i += 32) {
    const float w = __fmul_rn(pl[3 * Sp + i], (float)dT[i]);
    pl[3 * Sp + i] = w;
    if (co.weights) co.weights[t * S + i] = w;
  }
  __syncwarp();
  if (SIGMA_ONLY) return;
  // stage 3, six fp32 FMA chains in sample order: lane 0 acc (x = 1: fma(w, 1, acc) rounds exactly like acc + w),
  // 1 depth (x = z), 2..4 rgb (x = sigmoid), 5 sum of beta^2 (a = x = beta)
  float accum = 0.f;
  if (lane < 6) {
    const float* ap = pl + (lane == 5 ? 4 : 3) * Sp;
    const float* xp = lane == 0 ? sc.ones : lane == 1 ? sz : lane == 5 ? pl + 4 * Sp : pl + (lane - 2) * Sp;
    const int xs = lane == 0 ? 0 : 1;
    int i = 0;
    for (; i + 8 <= S; i += 8) {
      const float4 a0 = *reinterpret_cast<const float4*>(ap + i), a1 = *reinterpret_cast<const float4*>(ap + i + 4);
      const float4 x0 = *reinterpret_cast<const float4*>(xp + i * xs), x1 = *reinterpret_cast<const float4*>(xp + (i + 4) * xs);
      accum = __fmaf_rn(a0.x, x0.x, accum); accum = __fmaf_rn(a0.y, x0.y, accum);
      accum = __fmaf_rn(a0.z, x0.z, accum); accum = __fmaf_rn(a0.w, x0.w, accum);
      accum = __fmaf_rn(a1.x, x1.x, accum); accum = __fmaf_rn(a1.y, x1.y, accum);
      accum = __fmaf_rn(a1.z, x1.z, accum); accum = __fmaf_rn(a1.w, x1.w, accum);
    }
    for (; i < S; ++i) accum = __fmaf_rn(ap[i], xp[i * xs], accum);
  }
  const float acc = __shfl_sync(0xffffffffu, accum, 0);
  const float depth = __shfl_sync(0xffffffffu, accum, 1);
  if (lane >= 2 && lane <= 4 && co.rgb) {
    float c = accum;
    if (white_bkgd) c += 1.f - acc;
    co.rgb[t * 3 + (lane - 2)] = c;
  }
  if (lane == 0) {
    if (co.depth) co.depth[t] = depth;
    if (co.acc) co.acc[t] = acc;
    if (co.disp) {
      const float q = depth / acc;   // 0/0 -> NaN propagates like torch.max(1e-10, nan)
      co.disp[t] = 1.f / (isnan(q) ? q : fmaxf(1e-10f, q));
    }
  }
  if (lane == 5 && co.beta2) co.beta2[t] = accum;
}

__global__ void __launch_bounds__(32 * kRayWarps)
composite_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int n,
                 int S, int white_bkgd, CompositeOut co) {
  extern __shared__ __align__(16) float ray_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * kRayWarps + warp;
  if (t >= n) return;
  const RayScratch sc = ray_scratch(ray_smem + (size_t)warp * ray_scratch_floats(S), S);
  const float dn = ray_norm(rays_d + t * 3);
  composite_ray_warp<false>(raw + t * S * 5, z + t * S, dn, S, white_bkgd, t, co, sc, lane);
}

// Score-only final pass: sum over the samples of beta^2 per ray from the [n,S] beta rows of a beta-only fine pass
// (mlp_tc.cu, HEADS = 2).  One THREAD per ray: the sum is the same fp32 fused-multiply-add chain in sample order as
// lane 5 of composite_ray_warp runs (bit-identical), 32 independent chains per warp instead of one; a thread walks
// its row with 16-byte loads, every 128-byte line it touches is consumed from L1 within 8 steps.
__global__ void __launch_bounds__(128)
beta2_kernel(const float* __restrict__ beta, int n, int S, float* __restrict__ beta2) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float* b = beta + t * S;
  float accum = 0.f;
  int i = 0;
  if ((S & 3) == 0) {
    for (; i + 4 <= S; i += 4) {
      const float4 v = *reinterpret_cast<const float4*>(b + i);
      accum = __fmaf_rn(v.x, v.x, accum); accum = __fmaf_rn(v.y, v.y, accum);
      accum = __fmaf_rn(v.z, v.z, accum); accum = __fmaf_rn(v.w, v.w, accum);
    }
  }
  for (; i < S; ++i) accum = __fmaf_rn(b[i], b[i], accum);
  beta2[t] = accum;
}

// Coarse compositing + sample_pdf(det) (run_nerf_helpers.py:216-259) + sort(cat) (run_nerf.py:483), one warp per ray.
//
// sample_pdf's two double-precision accumulations (sum of the weights, cumsum of the pdf; torch accumulates fp32
// sums in double on CPU) are EXACT for these inputs, whatever the order: every addend is an fp32 number in
// [1e-5 / 2, 2) (the weights are >= 0 and get + 1e-5, the pdf entries are >= 1e-5 / (1 + 126e-5)), so all partial
// sums are multiples of 2^-41 below 2^7 -> at most 48 significant bits < 53.  A parallel reduction / scan in double
// therefore returns the very same values as the reference's serial loops.  The inverse-cdf lookup is
// searchsorted(right=True) = upper_bound on the non-decreasing cdf; the final sort(cat[...]) is the stable merge
// of two sorted runs (coarse entries first on ties), whose output positions are ranks found by binary search.
template <bool SIGMA_ONLY>
__global__ void __launch_bounds__(32 * kRayWarps)
coarse_resample_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                       const float* __restrict__ u_vals, int n, int Sc, int Ni, int white_bkgd, CompositeOut co,
                       float* __restrict__ z_fine, float* __restrict__ cdf_dump, int32_t* __restrict__ inds_dump,
                       float* __restrict__ zs_dump) {
  extern __shared__ __align__(16) float ray_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * kRayWarps + warp;
  if (t >= n) return;
  const int per_warp = resample_scratch_floats(Sc, Ni);
  float* base = ray_smem + (size_t)warp * per_warp;
  const RayScratch sc = ray_scratch(base, Sc);
  float* cdf = base + ray_scratch_floats(Sc);   // [Sc]
  float* zs = cdf + Sc;                          // [Ni]
  float* zf = zs + Ni;                           // [Sc + Ni]
  const float* wts = sc.pl + 3 * sc.Sp;          // coarse weights after composite_ray_warp
  const float* zc = sc.sz;
  const float dn = ray_norm(rays_d + t * 3);
  composite_ray_warp<SIGMA_ONLY>(raw + t * Sc * (SIGMA_ONLY ? 1 : 5), z + t * Sc, dn, Sc, white_bkgd, t, co, sc, lane);
  __syncwarp();

  // pdf over weights[1:-1] + 1e-5, cdf = [0, cumsum(pdf)]  -> nb = Sc-1 entries; lane L owns entries
  // k = 1 + L*B .. 1 + L*B + B-1 of the cumsum
  const int nb = Sc - 1;
  const int npdf = Sc - 2;
  const int B = (npdf + 31) / 32;      // <= 4 (Sc <= 128)
  float wk[4];
  double part = 0.0;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int k = 1 + lane * B + m;
    wk[m] = 0.f;
    if (m < B && k <= npdf) {
      wk[m] = __fadd_rn(wts[k], 1e-5f);
      part += (double)wk[m];
    }
  }
  double sumd = part;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sumd += __shfl_xor_sync(0xffffffffu, sumd, o);
  const float sum = (float)sumd;
  double loc[4];
  double run = 0.0;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int k = 1 + lane * B + m;
    if (m < B && k <= npdf) run += (double)__fdiv_rn(wk[m], sum);
    loc[m] = run;
  }
  double incl = run;   // inclusive scan of the per-lane totals
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const double excl = incl - run;   // exact: all of these are multiples of 2^-41 below 2^7
  if (lane == 0) cdf[0] = 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int k = 1 + lane * B + m;
    if (m < B && k <= npdf) cdf[k] = (float)(excl + loc[m]);
  }
  __syncwarp();
  if (cdf_dump) for (int k = lane; k < nb; k += 32) cdf_dump[t * nb + k] = cdf[k];

  // inverse-cdf lookups
  for (int j = lane; j < Ni; j += 32) {
    const float u = u_vals[j];
    int lo = 0, hi = nb;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int idx = lo;
    if (inds_dump) inds_dump[t * Ni + j] = idx;
    const int below = idx - 1 > 0 ? idx - 1 : 0;
    const int above = idx < nb - 1 ? idx : nb - 1;
    float denom = __fsub_rn(cdf[above], cdf[below]);
    if (denom < 1e-5f) denom = 1.f;
    const float tt = __fdiv_rn(__fsub_rn(u, cdf[below]), denom);
    const float bb = __fmul_rn(0.5f, __fadd_rn(zc[below + 1], zc[below]));   // z_vals_mid (run_nerf.py:479)
    const float ba = __fmul_rn(0.5f, __fadd_rn(zc[above + 1], zc[above]));
    const float s = __fadd_rn(bb, __fmul_rn(tt, __fsub_rn(ba, bb)));
    zs[j] = s;
    if (zs_dump) zs_dump[t * Ni + j] = s;
  }
  __syncwarp();
  if (co.z_std && lane == 0) {   // torch.std(z_samples, unbiased=False) (run_nerf.py:501)
    float mean = 0.f;
    for (int j = 0; j < Ni; ++j) mean += zs[j];
    mean /= (float)Ni;
    float var = 0.f;
    for (int j = 0; j < Ni; ++j) { const float dlt = zs[j] - mean; var = fmaf(dlt, dlt, var); }
    co.z_std[t] = sqrtf(var / (float)Ni);
  }
  // torch.sort(cat[z_vals, z_samples]): both runs are ascending up to rounding; make the sample run exactly
  // sorted (almost never needed), then merge by rank.
  bool unsorted = false;
  for (int j = lane + 1; j < Ni; j += 32) unsorted |= zs[j - 1] > zs[j];
  if (__any_sync(0xffffffffu, unsorted)) {
    if (lane == 0) {
      for (int j = 1; j < Ni; ++j) {
        const float v = zs[j];
        int k = j - 1;
        while (k >= 0 && zs[k] > v) { zs[k + 1] = zs[k]; --k; }
        zs[k + 1] = v;
      }
    }
    __syncwarp();
  }
  for (int a = lane; a < Sc; a += 32) {   // coarse entry a goes after the samples strictly below it
    const float v = zc[a];
    int lo = 0, hi = Ni;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (zs[mid] < v) lo = mid + 1; else hi = mid;
    }
    zf[a + lo] = v;
  }
  for (int b = lane; b < Ni; b += 32) {   // sample b goes after the coarse entries <= it
    const float v = zs[b];
    int lo = 0, hi = Sc;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (zc[mid] <= v) lo = mid + 1; else hi = mid;
    }
    zf[b + lo] = v;
  }
  __syncwarp();
  float* zg = z_fine + t * (Sc + Ni);
  for (int o = lane; o < Sc + Ni; o += 32) zg[o] = zf[o];
}

// Random pixel subsets drawn on the device: R DISTINCT pixels of each view's n-pixel image.  The reference draws
// them on the host with np.random.choice(n, R, replace=False) per pose (nerf.py:287, tsdf_gpu.py:497), which shuffles
// all n = 160 000 pixel ids for every pose (2 ms per view, seconds per planning step).  Here pixel k of view v is
// perm_v(k), k < R, where perm_v is a keyed pseudo-random permutation of [0, n): a 4-round Feistel network on the
// next even power of two with cycle walking back into the domain -- O(1) per pixel, distinct by construction.
// Not the reference's RNG stream (callers that need that pass pix_idx explicitly); same sampling model.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__global__ void pixel_subset_kernel(uint64_t seed, int V, int R, uint32_t n, int half_bits, int32_t* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)V * R) return;
  const uint32_t v = (uint32_t)(t / R), k = (uint32_t)(t % R);
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t key[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) key[r] = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9e3779b9u * (v * 4u + r + 1u)));
  uint32_t x = k;
  do {
    uint32_t l = x >> half_bits, rr = x & mask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t nl = rr;
      rr = l ^ (mix32(rr ^ key[r]) & mask);
      l = nl;
    }
    x = (l << half_bits) | rr;
  } while (x >= n);
  out[t] = (int32_t)x;
}

// score[v] = sum_rays beta2 / R   (nerf.py:307-309); fp64 tree, deterministic.
// A score that is not finite (16-bit activations overflowed somewhere upstream) raises *nonfinite.
__global__ void view_reduce_kernel(const float* __restrict__ beta2, int R, float* __restrict__ scores,
                                   int* __restrict__ nonfinite) {
  __shared__ double sh[256];
  const int v = blockIdx.x;
  double a = 0.0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) a += (double)beta2[(long long)v * R + r];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float sc = (float)(sh[0] / (double)R);
    scores[v] = sc;
    if (nonfinite && !isfinite(sc)) atomicAdd(nonfinite, 1);
  }
}

// plannerServer_Object/core/vpp.py:199-220 + interface.py:115-116, single CTA.
__global__ void select_nbv_kernel(const float* __restrict__ scores, const float* __restrict__ weight, int V,
                                  int dirs, float* __restrict__ norm_scores, int32_t* __restrict__ best_dir,
                                  int32_t* __restrict__ winner) {
  __shared__ double smax[256];
  __shared__ int sidx[256];
  const int L = V / dirs;
  double bestv = -1.0 / 0.0;
  int besti = -1;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    double m = 0.0;
    int mk = 0;
    for (int k = 0; k < dirs; ++k) {   // np.argmax: first maximum
      const int i = l * dirs + k;
      const double s = (double)scores[i] * (weight ? (double)weight[i] : 1.0);
      if (k == 0 || s > m) { m = s; mk = k; }
    }
    if (best_dir) best_dir[l] = mk;
    if (norm_scores) norm_scores[l] = (float)m;   // normalised below
    if (m > bestv || (m == bestv && l > besti)) { bestv = m; besti = l; }   // ties -> highest index
  }
  smax[threadIdx.x] = bestv;
  sidx[threadIdx.x] = besti;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double ov = smax[threadIdx.x + s];
      const int oi = sidx[threadIdx.x + s];
      if (ov > smax[threadIdx.x] || (ov == smax[threadIdx.x] && oi > sidx[threadIdx.x])) {
        smax[threadIdx.x] = ov;
        sidx[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  const double gmax = smax[0];
  const int gl = sidx[0];
  if (norm_scores)
    for (int l = threadIdx.x; l < L; l += blockDim.x) norm_scores[l] = (float)((double)norm_scores[l] / gmax);
  if (threadIdx.x == 0 && winner) {
    int mk = 0;
    double m = 0.0;
    for (int k = 0; k < dirs; ++k) {
      const int i = gl * dirs + k;
      const double s = (double)scores[i] * (weight ? (double)weight[i] : 1.0);
      if (k == 0 || s > m) { m = s; mk = k; }
    }
    winner[0] = gl * dirs + mk;
  }
}

}  // namespace

int launch_get_rays(const float* c2w, int V, int H, int W, float fx, float fy, float cx, float cy,
                    const int32_t* pix_idx, int R, long long ray_begin, int n, float* rays_o, float* rays_d,
                    float* viewdirs, cudaStream_t st) {
  (void)V;
  if (n <= 0) return 0;
  get_rays_kernel<<<(n + 255) / 256, 256, 0, st>>>(c2w, H, W, fx, fy, cx, cy, pix_idx, R, ray_begin, n, rays_o,
                                                    rays_d, viewdirs);
  return 1;
}

int launch_normalize_dirs(const float* rays_d, int n, float* viewdirs, cudaStream_t st) {
  if (n <= 0) return 0;
  normalize_dirs_kernel<<<(n + 255) / 256, 256, 0, st>>>(rays_d, n, viewdirs);
  return 1;
}

int launch_coarse_z(const float* t_vals, int n, int S, float near_plane, float far_plane, int lindisp, float* z,
                    cudaStream_t st) {
  const long long total = (long long)n * S;
  if (total <= 0) return 0;
  coarse_z_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(t_vals, total, S, near_plane, far_plane,
                                                                   lindisp, z);
  return 1;
}

template <typename K>
static bool ray_kernel_smem(K kern, size_t bytes) {
  if (bytes <= 48 * 1024) return true;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

int launch_coarse_resample(const float* raw, const float* z, const float* rays_d, const float* u_vals, int n,
                           int Sc, int Ni, int white_bkgd, CompositeOut co, float* z_fine, float* cdf_dump,
                           int32_t* inds_dump, float* zs_dump, cudaStream_t st, bool sigma_only) {
  if (n <= 0) return 0;
  const size_t smem = (size_t)kRayWarps * resample_scratch_floats(Sc, Ni) * sizeof(float);
  if (sigma_only) {
    if (!ray_kernel_smem(coarse_resample_kernel<true>, smem)) return -1;
    coarse_resample_kernel<true><<<(n + kRayWarps - 1) / kRayWarps, 32 * kRayWarps, smem, st>>>(
        raw, z, rays_d, u_vals, n, Sc, Ni, white_bkgd, co, z_fine, cdf_dump, inds_dump, zs_dump);
  } else {
    if (!ray_kernel_smem(coarse_resample_kernel<false>, smem)) return -1;
    coarse_resample_kernel<false><<<(n + kRayWarps - 1) / kRayWarps, 32 * kRayWarps, smem, st>>>(
        raw, z, rays_d, u_vals, n, Sc, Ni, white_bkgd, co, z_fine, cdf_dump, inds_dump, zs_dump);
  }
  return 1;
}

int launch_beta2(const float* beta, int n, int S, float* beta2, cudaStream_t st) {
  if (n <= 0) return 0;
  beta2_kernel<<<(n + 127) / 128, 128, 0, st>>>(beta, n, S, beta2);
  return 1;
}

int launch_composite(const float* raw, const float* z, const float* rays_d, int n, int S, int white_bkgd,
                     CompositeOut co, cudaStream_t st) {
  if (n <= 0) return 0;
  const size_t smem = (size_t)kRayWarps * ray_scratch_floats(S) * sizeof(float);
  if (!ray_kernel_smem(composite_kernel, smem)) return -1;
  composite_kernel<<<(n + kRayWarps - 1) / kRayWarps, 32 * kRayWarps, smem, st>>>(raw, z, rays_d, n, S, white_bkgd, co);
  return 1;
}

int launch_pixel_subset(uint64_t seed, int V, int R, int n_pixels, int32_t* out, cudaStream_t st) {
  const long long total = (long long)V * R;
  if (total <= 0) return 0;
  int half_bits = 1;
  while ((1ll << (2 * half_bits)) < n_pixels) ++half_bits;
  pixel_subset_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(seed, V, R, (uint32_t)n_pixels, half_bits, out);
  return 1;
}

int launch_view_reduce(const float* beta2, int V, int R, float* scores, int* nonfinite, cudaStream_t st) {
  if (V <= 0) return 0;
  view_reduce_kernel<<<V, 256, 0, st>>>(beta2, R, scores, nonfinite);
  return 1;
}

int launch_select_nbv(const float* scores, const float* weight, int V, int dirs, float* norm_scores,
                      int32_t* best_dir, int32_t* winner, cudaStream_t st) {
  select_nbv_kernel<<<1, 256, 0, st>>>(scores, weight, V, dirs, norm_scores, best_dir, winner);
  return 1;
}

}  // namespace evpp
