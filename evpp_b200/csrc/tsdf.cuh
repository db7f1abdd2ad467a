// TSDF occupancy ray-cast scorer (stage 1 of the planner), see tsdf.cu.
#pragma once
#include "common.cuh"

namespace evpp {

struct TsdfVolume {          // device pointers + geometry of one state map
  const uint8_t* state;      // [X,Y,Z] 0 unknown, 1 empty, 2 occupied (State_vol_origin, tsdf_gpu.py:778)
  const float* nray;         // [X,Y,Z] observation count of surface voxels (Nray_vol)
  const float* t_vals;       // [n_samples] torch.linspace(0,1,n_samples)
  int dim[3];
  float origin[3];           // vol_origin (fusion.py:39)
  float voxel;               // voxel_size
  float amin[3], amax[3];    // aabb_bnds (tsdf_gpu.py:31)
  float near_plane, far_plane;
  int n_samples;             // int((far-near)/voxel)+1 = 56
  float c_unknown, c_surface;   // 0.2 / 0.1 (Object), unknown_cof / surface_cof (Room)
};

int launch_tsdf_pack_state(const float* state, long long n, uint8_t* out, cudaStream_t st);
int launch_tsdf_march(const TsdfVolume& vol, const float* rays_o, const float* rays_d, int n, float* uncer, float* depth,
                      int32_t* n_unknown, int32_t* first_hit, int32_t* voxel_idx, cudaStream_t st);
int launch_tsdf_view_reduce(const float* uncer, const float* depth, int V, int R, float near_plane, float far_plane,
                            float* view_uncer, float* view_depth, cudaStream_t st);

}  // namespace evpp
