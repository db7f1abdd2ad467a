// NGP-mode scorer: occupancy-bitfield marching, hash-grid encoding, small MLP, compositing (see ngp.cu).
#pragma once
#include "common.cuh"

namespace evpp {

constexpr int kNgpMaxLevels = 16;
// packed fp32 MLP: sigma0 [64][32], sigma1 [16][64], color0 [64][32] (31 inputs + zero column),
// color1 [64][64], heads [4][64] (r, g, b, uncertainty), uncertainty bias [1]
constexpr int kNgpMlpFloats = 64 * 32 + 16 * 64 + 64 * 32 + 64 * 64 + 4 * 64 + 1;

struct NgpLevel {
  float scale;          // 2^(l*log2(b)) * base - 1
  uint32_t res;         // ceil(scale) + 1
  uint32_t offset;      // first entry of the level in the table
  uint32_t size;        // entries of the level (multiple of 8, <= 2^log2_hashmap)
  uint32_t hashed;      // 1: (res+1)^3 exceeds the table -> spatial hash, size is a power of two; 0: dense, index < size
  uint32_t stride_y, stride_z;   // dense strides (res+1), (res+1)^2
};

// TensoRF VM backbone: packed fp32 colour network: basis [27][144], color0 [64][43] (27 basis + 16 SH inputs),
// color1 [64][64], heads [4][64] (r, g, b, uncertainty), uncertainty bias [1]
constexpr int kTfMlpFloats = 27 * 144 + 64 * 43 + 64 * 64 + 4 * 64 + 1;

struct NgpParams {      // device pointers + configuration, passed to kernels by value
  const uint16_t* tf_planes;    // TensoRF: fp16 [3][R][R][64] indexed [i][y][x][c] (16 density + 48 appearance channels)
  const uint16_t* tf_lines;     // TensoRF: fp16 [3][R][64]
  const float* tf_mlp;          // kTfMlpFloats
  int tf_res;                   // R
  const uint16_t* table;        // fp16 [entries][2]
  const float* mlp;             // kNgpMlpFloats
  const uint8_t* bitfield;      // grid_res^3 / 8 bytes, cell (x*G+y)*G+z -> bit (cell & 7) of byte cell >> 3
  float occ_lo[3], occ_hi[3];   // march_mode 0: world-space box of the occupied cells grown by one cell (evpp_ngp_set_occupancy);
                                // steps outside it cannot be kept, so the marcher does not test them.  occ_lo > occ_hi: nothing occupied
  NgpLevel levels[kNgpMaxLevels];
  int n_levels;
  int grid_res;
  float amin[3], amax[3], inv_size[3];
  float near_plane, far_plane, dt;
  int max_steps, max_samples;
  int white_bkgd;
  // march_mode 1 (torch-ngp raymarching semantics): cubic scene box mapped to [-bound, bound]^3, `cascades` nested
  // Morton-ordered H^3 bitfields, dt = clamp(t * dt_gamma, dt_min, dt_max), empty cells skipped to their far face
  int march_mode;
  int cascades;
  float bound, scale, center[3];   // x_n = (x_world - center) * scale
  float dt_gamma, dt_min, dt_max, min_near;
};

int launch_ngp_encode(const NgpParams& P, const float* u, long long M, float* enc, int32_t* idx, cudaStream_t st);
int ngp_mask_words(const NgpParams& P);   // ballot words per ray kept between the count and the fill pass
int ngp_scan_tiles(int n);                // the scan needs a workspace of 2 * ngp_scan_tiles(n) ints
int launch_ngp_march_count(const NgpParams& P, const float* ro, const float* rd, int n, int32_t* counts, uint32_t* masks,
                           float* t0, cudaStream_t st);
// total: device int (samples of this chunk); sample_counter: optional device int64 running sum over chunks
int launch_ngp_scan(const int32_t* counts, int n, int32_t* offsets, int32_t* total, int32_t* tile_ws,
                    unsigned long long* sample_counter, cudaStream_t st);
// torch-ngp marcher (march_mode 1): thread per ray, count pass then (after the scan) fill pass; delta [M] = step
// length of every kept sample in normalised units
int launch_tngp_march_count(const NgpParams& P, const float* ro, const float* rd, int n, int32_t* counts, cudaStream_t st);
int launch_tngp_march_fill(const NgpParams& P, const float* ro, const float* rd, int n, const int32_t* counts,
                           const int32_t* offsets, float* t, int32_t* ray_of_sample, float* delta, int32_t* cell_index,
                           cudaStream_t st);
int launch_ngp_march_fill(const NgpParams& P, int n, const int32_t* counts, const int32_t* offsets, const uint32_t* masks,
                          const float* t0, float* t, int32_t* ray_of_sample, cudaStream_t st);
// M = upper bound of the sample count (the buffers' capacity); total_dev (optional, device) = the actual count, read
// by the kernel itself, so that the host never has to wait for the marcher's total
int launch_ngp_field(const NgpParams& P, const float* ro, const float* rd, const float* t, const int32_t* ray_of_sample,
                     long long M, const int32_t* total_dev, float* raw, float* enc_dump, int num_sms, bool tensor_cores,
                     cudaStream_t st);
int launch_tensorf_field(const NgpParams& P, const float* ro, const float* rd, const float* t, const int32_t* ray_of_sample,
                         long long M, const int32_t* total_dev, float* raw, int num_sms, cudaStream_t st);
// delta: optional per-sample step lengths (march_mode 1); null = the constant P.dt
int launch_ngp_composite(const NgpParams& P, const float* raw, const float* t, const float* delta, const int32_t* counts,
                         const int32_t* offsets, int n, float* rgb, float* depth, float* acc, float* beta2, cudaStream_t st);

}  // namespace evpp
