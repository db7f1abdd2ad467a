// Stage-1 candidate scorer of the planner: occupancy-state ray cast through the TSDF state volume.
// Replaces TSDF.render_rays + isectSegAABB_gpu + the per-view means of get_uncertainty_tsdf
// (plannerServer_Object/core/tsdf_gpu.py:198-354, :67-114, :473-530; Room copy: coefficients only).
// The reference runs ~60 ATen launches and 6 host round-trips per 32768-ray chunk; here one thread
// marches one ray through the 56 samples (the 0.5 MB uint8 state volume is L2/L1 resident) and one CTA
// reduces one view.  Integer outputs (voxel index, unknown count, first-hit sample) are bit-exact with the
// reference's fp32 arithmetic, quirks included (out-of-bounds cascade, first-hit rule, slab test).
#include "common.cuh"
#include "tsdf.cuh"

namespace evpp {

namespace {

__global__ void tsdf_pack_state_kernel(const float* __restrict__ state, long long n, uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint8_t)(int)state[i];
}

__global__ void tsdf_march_kernel(TsdfVolume vol, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                  int n, float* __restrict__ uncer, float* __restrict__ depth,
                                  int32_t* __restrict__ n_unknown, int32_t* __restrict__ first_hit,
                                  int32_t* __restrict__ voxel_idx) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float o[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { o[c] = rays_o[(long long)t * 3 + c]; d[c] = rays_d[(long long)t * 3 + c]; }
  const int S = vol.n_samples;
  int n0 = 0, k = 0;
  bool found = false;
  float nray_hit = 0.f;
  for (int s = 0; s < S; ++s) {
    const float tv = vol.t_vals[s];
    // z = near*(1-t) + far*t (:206), pts = o + d*z (:207), each op rounded like ATen
    const float z = __fadd_rn(__fmul_rn(vol.near_plane, __fsub_rn(1.f, tv)), __fmul_rn(vol.far_plane, tv));
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      v[c] = floorf(__fdiv_rn(__fsub_rn(__fadd_rn(o[c], __fmul_rn(d[c], z)), vol.origin[c]), vol.voxel));   // :211
    // sequential out-of-bounds rewrite (:233-247): lower cascade, then upper cascade to ZERO
    if (v[2] < 0.f) { v[0] = 0.f; v[1] = 0.f; v[2] = 0.f; }
    else if (v[1] < 0.f) { v[0] = 0.f; v[1] = 0.f; }
    else if (v[0] < 0.f) { v[0] = 0.f; }
    if (v[2] >= (float)(vol.dim[2] - 1)) { v[0] = 0.f; v[1] = 0.f; v[2] = 0.f; }
    else if (v[1] >= (float)(vol.dim[1] - 1)) { v[0] = 0.f; v[1] = 0.f; }
    else if (v[0] >= (float)(vol.dim[0] - 1)) { v[0] = 0.f; }
    const int ix = (int)v[0], iy = (int)v[1], iz = (int)v[2];
    if (voxel_idx) {
      int32_t* vi = voxel_idx + ((long long)t * S + s) * 3;
      vi[0] = ix; vi[1] = iy; vi[2] = iz;
    }
    const long long lin = ((long long)ix * vol.dim[1] + iy) * vol.dim[2] + iz;
    const uint8_t st = vol.state[lin];
    n0 += st == 0;                                    // unknown voxels along the ray (:276-279)
    if (st == 2 && !found) {                          // first occupied sample = argmin(3|10) (:285-287)
      found = true;
      k = s;
      nray_hit = vol.nray[lin];
    }
  }
  // unknown-space term (:280-282)
  float u = 0.f;
  if (n0 > 0) u = __fmul_rn(__fdiv_rn((float)n0, (float)S), vol.c_unknown);
  if (u > 1.f) u = 1.f;
  float dep = vol.far_plane;
  if (k != 0) {                                       // k == 0 counts as "no hit" (:288)
    dep = (float)((double)vol.near_plane + (double)k / (double)S * ((double)vol.far_plane - (double)vol.near_plane));   // :293
    u = __fdiv_rn(1.f, __fadd_rn(1.f, __fmul_rn(__fmul_rn(vol.c_surface, dep), nray_hit)));                              // :298
  }
  // slab test against the scoring box (:67-114); axes with |d| < 1e-6 are skipped entirely
  float lo = 0.f, hi = 10000.f;
  bool hit = true;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (fabsf(d[c]) >= 1e-6f) {
      const float ood = __fdiv_rn(1.f, d[c]);
      float t1 = __fmul_rn(__fsub_rn(vol.amin[c], o[c]), ood);
      float t2 = __fmul_rn(__fsub_rn(vol.amax[c], o[c]), ood);
      if (t1 > t2) { const float tmp = t1; t1 = t2; t2 = tmp; }
      if (t1 > lo) lo = t1;
      if (t2 < hi) hi = t2;
      if (lo > hi) hit = false;
    }
  }
  if (!hit) { u = 1e-10f; dep = vol.far_plane; }      // :299-300
  if (uncer) uncer[t] = u;
  if (depth) depth[t] = dep;
  if (n_unknown) n_unknown[t] = n0;
  if (first_hit) first_hit[t] = k;
}

// per view: mean of u over its R rays; mean of depths strictly inside (near, far), else far (:513-529)
__global__ void tsdf_view_reduce_kernel(const float* __restrict__ uncer, const float* __restrict__ depth, int R,
                                        float near_plane, float far_plane, float* __restrict__ view_uncer,
                                        float* __restrict__ view_depth) {
  __shared__ double su[256], sd[256];
  __shared__ int sc[256];
  const int v = blockIdx.x;
  double au = 0.0, ad = 0.0;
  int cnt = 0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    au += (double)uncer[(long long)v * R + r];
    const float dd = depth[(long long)v * R + r];
    if (dd < far_plane && dd > near_plane) { ad += (double)dd; ++cnt; }
  }
  su[threadIdx.x] = au; sd[threadIdx.x] = ad; sc[threadIdx.x] = cnt;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      su[threadIdx.x] += su[threadIdx.x + s];
      sd[threadIdx.x] += sd[threadIdx.x + s];
      sc[threadIdx.x] += sc[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    view_uncer[v] = (float)(su[0] / (double)R);
    view_depth[v] = sc[0] == 0 ? far_plane : (float)(sd[0] / (double)sc[0]);
  }
}

}  // namespace

int launch_tsdf_pack_state(const float* state, long long n, uint8_t* out, cudaStream_t st) {
  if (n <= 0) return 0;
  tsdf_pack_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(state, n, out);
  return 1;
}

int launch_tsdf_march(const TsdfVolume& vol, const float* rays_o, const float* rays_d, int n, float* uncer, float* depth,
                      int32_t* n_unknown, int32_t* first_hit, int32_t* voxel_idx, cudaStream_t st) {
  if (n <= 0) return 0;
  tsdf_march_kernel<<<(n + 127) / 128, 128, 0, st>>>(vol, rays_o, rays_d, n, uncer, depth, n_unknown, first_hit, voxel_idx);
  return 1;
}

int launch_tsdf_view_reduce(const float* uncer, const float* depth, int V, int R, float near_plane, float far_plane,
                            float* view_uncer, float* view_depth, cudaStream_t st) {
  if (V <= 0) return 0;
  tsdf_view_reduce_kernel<<<V, 256, 0, st>>>(uncer, depth, R, near_plane, far_plane, view_uncer, view_depth);
  return 1;
}

}  // namespace evpp
