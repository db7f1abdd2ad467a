// Ray generation, sampling, compositing and score reduction kernels.
// All are tiny next to the MLP (a ray costs 304 MFLOP of MLP and ~6 KB of traffic here), so they
// are written for reference-identical arithmetic first: one thread walks one ray sequentially,
// exactly like the serial cumprod / cumsum / searchsorted of the reference's ATen CPU ops.
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

// raw2outputs (run_nerf.py:344-388) for one ray, sequentially.  w_out (local or global) receives
// the weights if non-null.
template <typename WOut>
__device__ __forceinline__ void composite_ray(const float* __restrict__ raw, const float* __restrict__ z,
                                              float dnorm, int S, int white_bkgd, int t,
                                              const CompositeOut& co, WOut w_store) {
  // torch.cumprod / cumsum on CPU accumulate fp32 inputs in double and round each output to
  // fp32 (at::acc_type<float,false>); T is carried the same way.
  double T = 1.0;
  float acc = 0.f, depth = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, b2 = 0.f;
  float zi = z[0];
  for (int i = 0; i < S; ++i) {
    const float* rw = raw + (long long)i * 5;
    const float zn = (i + 1 < S) ? z[i + 1] : 0.f;
    float dist = (i + 1 < S) ? __fsub_rn(zn, zi) : 1e10f;
    dist = __fmul_rn(dist, dnorm);
    const float sigma = fmaxf(rw[3], 0.f);
    const float alpha = __fsub_rn(1.f, expf(__fmul_rn(-sigma, dist)));
    const float w = __fmul_rn(alpha, (float)T);
    T = T * (double)__fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
    w_store(i, w);
    c0 = __fmaf_rn(w, 1.f / (1.f + expf(-rw[0])), c0);
    c1 = __fmaf_rn(w, 1.f / (1.f + expf(-rw[1])), c1);
    c2 = __fmaf_rn(w, 1.f / (1.f + expf(-rw[2])), c2);
    depth = __fmaf_rn(w, zi, depth);
    acc += w;
    const float beta = rw[4];
    b2 = __fmaf_rn(beta, beta, b2);
    if (co.unc) co.unc[(long long)t * S + i] = beta;
    zi = zn;
  }
  if (white_bkgd) { c0 += 1.f - acc; c1 += 1.f - acc; c2 += 1.f - acc; }
  if (co.rgb) { co.rgb[(long long)t * 3 + 0] = c0; co.rgb[(long long)t * 3 + 1] = c1; co.rgb[(long long)t * 3 + 2] = c2; }
  if (co.depth) co.depth[t] = depth;
  if (co.acc) co.acc[t] = acc;
  if (co.disp) {
    const float q = depth / acc;   // 0/0 -> NaN propagates like torch.max(1e-10, nan)
    co.disp[t] = 1.f / (isnan(q) ? q : fmaxf(1e-10f, q));
  }
  if (co.beta2) co.beta2[t] = b2;
}

__global__ void composite_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                 const float* __restrict__ rays_d, int n, int S, int white_bkgd,
                                 CompositeOut co) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float dn = ray_norm(rays_d + (long long)t * 3);
  float* wg = co.weights ? co.weights + (long long)t * S : nullptr;
  composite_ray(raw + (long long)t * S * 5, z + (long long)t * S, dn, S, white_bkgd, t, co,
                [wg](int i, float w) { if (wg) wg[i] = w; });
}

// Coarse compositing + sample_pdf(det) (run_nerf_helpers.py:216-259) + sort(cat) (run_nerf.py:483).
__global__ void coarse_resample_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                       const float* __restrict__ rays_d, const float* __restrict__ u_vals,
                                       int n, int Sc, int Ni, int white_bkgd, CompositeOut co,
                                       float* __restrict__ z_fine, float* __restrict__ cdf_dump,
                                       int32_t* __restrict__ inds_dump, float* __restrict__ zs_dump) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float w[kMaxSc];
  float cdf[kMaxSc];
  float zs[kMaxSf];
  const float* zc = z + (long long)t * Sc;
  const float dn = ray_norm(rays_d + (long long)t * 3);
  float* wg = co.weights ? co.weights + (long long)t * Sc : nullptr;
  composite_ray(raw + (long long)t * Sc * 5, zc, dn, Sc, white_bkgd, t, co,
                [&w, wg](int i, float v) { w[i] = v; if (wg) wg[i] = v; });

  // pdf over weights[1:-1] + 1e-5, cdf = [0, cumsum(pdf)]  -> nb = Sc-1 entries
  const int nb = Sc - 1;
  double sumd = 0.0;
  for (int k = 1; k < Sc - 1; ++k) { w[k] = __fadd_rn(w[k], 1e-5f); sumd += (double)w[k]; }
  const float sum = (float)sumd;
  cdf[0] = 0.f;
  double run = 0.0;
  for (int k = 1; k < nb; ++k) { run += (double)__fdiv_rn(w[k], sum); cdf[k] = (float)run; }
  if (cdf_dump) for (int k = 0; k < nb; ++k) cdf_dump[(long long)t * nb + k] = cdf[k];

  // inverse-cdf lookups; u ascending -> searchsorted(right=True) advances monotonically
  int idx = 0;
  float mean = 0.f;
  for (int j = 0; j < Ni; ++j) {
    const float u = u_vals[j];
    while (idx < nb && cdf[idx] <= u) ++idx;
    if (inds_dump) inds_dump[(long long)t * Ni + j] = idx;
    const int below = idx - 1 > 0 ? idx - 1 : 0;
    const int above = idx < nb - 1 ? idx : nb - 1;
    float denom = __fsub_rn(cdf[above], cdf[below]);
    if (denom < 1e-5f) denom = 1.f;
    const float tt = __fdiv_rn(__fsub_rn(u, cdf[below]), denom);
    const float bb = __fmul_rn(0.5f, __fadd_rn(zc[below + 1], zc[below]));   // z_vals_mid (run_nerf.py:479)
    const float ba = __fmul_rn(0.5f, __fadd_rn(zc[above + 1], zc[above]));
    const float s = __fadd_rn(bb, __fmul_rn(tt, __fsub_rn(ba, bb)));
    zs[j] = s;
    mean += s;
    if (zs_dump) zs_dump[(long long)t * Ni + j] = s;
  }
  if (co.z_std) {   // torch.std(z_samples, unbiased=False) (run_nerf.py:501)
    mean /= (float)Ni;
    float var = 0.f;
    for (int j = 0; j < Ni; ++j) { const float dlt = zs[j] - mean; var = fmaf(dlt, dlt, var); }
    co.z_std[t] = sqrtf(var / (float)Ni);
  }
  // torch.sort(cat[z_vals, z_samples]): both runs are ascending up to rounding; make the
  // sample run exactly sorted (no-op walk when it already is), then merge.
  for (int j = 1; j < Ni; ++j) {
    const float v = zs[j];
    int k = j - 1;
    while (k >= 0 && zs[k] > v) { zs[k + 1] = zs[k]; --k; }
    zs[k + 1] = v;
  }
  float* zf = z_fine + (long long)t * (Sc + Ni);
  int a = 0, b = 0;
  for (int o = 0; o < Sc + Ni; ++o) {
    const bool take_c = (b >= Ni) || (a < Sc && zc[a] <= zs[b]);
    zf[o] = take_c ? zc[a] : zs[b];
    if (take_c) ++a; else ++b;
  }
}

// score[v] = sum_rays beta2 / R   (nerf.py:307-309); fp64 tree, deterministic.
__global__ void view_reduce_kernel(const float* __restrict__ beta2, int R, float* __restrict__ scores) {
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
  if (threadIdx.x == 0) scores[v] = (float)(sh[0] / (double)R);
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

int launch_coarse_resample(const float* raw, const float* z, const float* rays_d, const float* u_vals, int n,
                           int Sc, int Ni, int white_bkgd, CompositeOut co, float* z_fine, float* cdf_dump,
                           int32_t* inds_dump, float* zs_dump, cudaStream_t st) {
  if (n <= 0) return 0;
  coarse_resample_kernel<<<(n + 127) / 128, 128, 0, st>>>(raw, z, rays_d, u_vals, n, Sc, Ni, white_bkgd, co,
                                                          z_fine, cdf_dump, inds_dump, zs_dump);
  return 1;
}

int launch_composite(const float* raw, const float* z, const float* rays_d, int n, int S, int white_bkgd,
                     CompositeOut co, cudaStream_t st) {
  if (n <= 0) return 0;
  composite_kernel<<<(n + 127) / 128, 128, 0, st>>>(raw, z, rays_d, n, S, white_bkgd, co);
  return 1;
}

int launch_view_reduce(const float* beta2, int V, int R, float* scores, cudaStream_t st) {
  if (V <= 0) return 0;
  view_reduce_kernel<<<V, 256, 0, st>>>(beta2, R, scores);
  return 1;
}

int launch_select_nbv(const float* scores, const float* weight, int V, int dirs, float* norm_scores,
                      int32_t* best_dir, int32_t* winner, cudaStream_t st) {
  select_nbv_kernel<<<1, 256, 0, st>>>(scores, weight, V, dirs, norm_scores, best_dir, winner);
  return 1;
}

}  // namespace evpp
