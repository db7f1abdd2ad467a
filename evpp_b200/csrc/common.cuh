// Shared declarations of the evpp_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/evpp_b200.h"

namespace evpp {

// ---- fixed architecture of the reference NeRF (run_nerf_helpers.py:78-134, nerf_configs.py) ----
constexpr int kW = 256;          // netwidth
constexpr int kD = 8;            // netdepth
constexpr int kSkip = 4;         // skips=[4]
constexpr int kLpos = 10;        // multires
constexpr int kLdir = 4;         // multires_views
constexpr int kEmbPos = 63;      // 3 + 3*2*10
constexpr int kEmbDir = 27;      // 3 + 3*2*4
constexpr int kWv = 128;         // W/2 view branch
constexpr int kOut = 5;          // r,g,b,sigma,beta
constexpr double kFlopPerEval = 1187072.0;

// ---- fp32 (exact) MLP weight stream: 8 KB chunks of [kc][N] fp32, see mlp_simt.cu -------------
constexpr int kSimtChunkFloats = 2048;
// chunk counts per layer: L0 64/8=8, L1-4 32 each, L5 (64+256)/8=40, L6-7 32, feature 32,
// views (256+32)/16=18
constexpr int kSimtChunksTotal = 8 + 4 * 32 + 40 + 2 * 32 + 32 + 18;

struct SimtWeights {          // device pointers
  const float* stream;        // [kSimtChunksTotal][2048]
  const float* bias;          // [8*256 + 256(feature) + 128(views)]
  const float* head_w;        // alpha [256], rgb [3][128], unc [128]  -> 256 + 384 + 128
  const float* head_b;        // alpha, r, g, b, unc -> 5
};

// ---- tcgen05 MLP (fp16 or bf16 operands), see mlp_tc.cu ------------------------------------------
struct TcConsts {             // fp32 epilogue constants, passed to the kernel by value (constant bank)
  float bias[10][256];        // pts_linears 0..7, feature_linear (8), views_linears.0 (9, first 128)
  float w_alpha[256];
  float w_rgb[3][128];
  float w_unc[128];
  float head_b[8];            // alpha, r, g, b, uncertainty
  float inv_wscale[12];       // split-precision passes: 1 / (power-of-two scale of layer l's weight image)
};
struct TcWeights {
  const void* image;          // device: pre-tiled, pre-swizzled B-operand K-blocks, streamed with cp.async.bulk
  const TcConsts* consts;     // host
  int dtype;                  // EVPP_PREC_FP16 / EVPP_PREC_BF16 / EVPP_PREC_FP16X2 / EVPP_PREC_FP16X3
  float* dbg;                 // optional device [M,256] dump of one layer's activations (parity tests)
  int dbg_layer;
  long long* trace;           // optional device [128] clock64 stamps of one tile's pipeline (diagnostics)
  int folded;                 // image / consts hold feature_linear folded into views_linears.0 (score-only fine pass, heads = 3)
};

struct RayBatch {             // device pointers of one chunk
  const float* rays_o;        // [n,3]
  const float* rays_d;        // [n,3]
  const float* viewdirs;      // [n,3]
  const float* z;             // [n,S]
  int n;
  int S;
};

// launchers (each returns the number of kernels launched)
int launch_mlp_simt(const SimtWeights& w, const RayBatch& rb, float* raw, cudaStream_t st);
int launch_mlp_simt_points(const SimtWeights& w, const float* pts, const float* dirs, int64_t M, float* out,
                           cudaStream_t st);
// heads: 0 = raw [n,S,5]; 1 = sigma only -> raw [n,S] (pass stops after layer 7); 2 = beta only -> raw [n,S];
// 3 = beta only with feature_linear folded into the view layer (needs weights packed with fold = true);
// 4 = raw [n,S,5] with the folded view layer
int launch_mlp_tc(const TcWeights& w, const RayBatch& rb, float* raw, int num_sms, cudaStream_t st, int heads = 0);
int launch_mlp_tc_points(const TcWeights& w, const float* pts, const float* dirs, int64_t M, float* out,
                         int num_sms, cudaStream_t st);
size_t tc_weight_image_bytes();
// host-side repack of one network: fills `image` (tc_weight_image_bytes()) from fp32 state_dict tensors
// fold: multiply feature_linear into views_linears.0 (nine GEMM layers; the image of the score-only fine pass)
void tc_pack_weights(const std::vector<std::vector<float>>& tensors, int dtype, bool fold, std::vector<uint8_t>& image,
                     TcConsts& consts);

}  // namespace evpp
