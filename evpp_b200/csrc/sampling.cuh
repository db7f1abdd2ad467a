// Ray generation, stratified / hierarchical sampling, alpha compositing and score reduction.
#pragma once
#include "common.cuh"

namespace evpp {

constexpr int kMaxSc = 128;   // max coarse samples handled by the per-ray kernels
constexpr int kMaxSf = 384;   // max fine samples

struct CompositeOut {          // all optional (nullptr = skip)
  float* rgb; float* disp; float* acc; float* depth; float* unc; float* z_std;
  float* weights;              // [n,S]
  float* beta2;                // [n] sum over samples of raw[...,4]^2 (score numerator)
};

int launch_get_rays(const float* c2w, int V, int H, int W, float fx, float fy, float cx, float cy,
                    const int32_t* pix_idx, int R, long long ray_begin, int n, float* rays_o, float* rays_d,
                    float* viewdirs, cudaStream_t st);
int launch_normalize_dirs(const float* rays_d, int n, float* viewdirs, cudaStream_t st);
int launch_coarse_z(const float* t_vals, int n, int S, float near_plane, float far_plane, int lindisp, float* z,
                    cudaStream_t st);
// raw2outputs on the coarse pass + sample_pdf + sort (run_nerf.py:473-484)
int launch_coarse_resample(const float* raw, const float* z, const float* rays_d, const float* u_vals, int n,
                           int Sc, int Ni, int white_bkgd, CompositeOut co, float* z_fine, float* cdf_dump,
                           int32_t* inds_dump, float* zs_dump, cudaStream_t st, bool sigma_only = false);
// score-only final pass: per-ray sum of beta^2 from beta rows [n,S] (bit-identical to launch_composite's beta2)
int launch_beta2(const float* beta, int n, int S, float* beta2, cudaStream_t st);
// raw2outputs on the final pass, fused with the per-ray sum of beta^2
int launch_composite(const float* raw, const float* z, const float* rays_d, int n, int S, int white_bkgd,
                     CompositeOut co, cudaStream_t st);
// R distinct pseudo-random pixel ids in [0, n_pixels) per view (keyed Feistel permutation), out [V,R]
int launch_pixel_subset(uint64_t seed, int V, int R, int n_pixels, int32_t* out, cudaStream_t st);
// nonfinite (optional, device int): incremented for every view whose score is inf / NaN
int launch_view_reduce(const float* beta2, int V, int R, float* scores, int* nonfinite, cudaStream_t st);
int launch_select_nbv(const float* scores, const float* weight, int V, int dirs, float* norm_scores,
                      int32_t* best_dir, int32_t* winner, cudaStream_t st);

}  // namespace evpp
