/*
 * evpp_b200 -- C ABI of the B200-native candidate-view scoring engine.
 *
 * The reference (small-zeng/EVPP) has no FFI layer; its seam for this path is two Python
 * call signatures:
 *   render(H, W, K, chunk, rays, c2w, ndc, near, far, use_viewdirs, c2w_staticcam, dev, **kwargs)
 *        nerfServer/core/run_nerf.py:131-196
 *   Controller.get_uncertainty(self, poses) -> Tensor[V]
 *        nerfServer/core/nerf.py:266-323
 * Each entry point below names the reference code it replaces.  All functions return 0 on
 * success and a negative evpp_status on failure; evpp_last_error() returns a thread-local,
 * NUL-terminated description of the last failure on the calling thread.
 *
 * Pointers documented "device" are CUDA device pointers owned by the caller (e.g. PyTorch
 * tensors); "host" pointers are ordinary (ideally pinned) host memory.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  The engine keeps its
 * workspace per handle: calls on ONE handle must be serialised by the caller; different
 * handles are independent (the Django server of the reference is thread-per-request,
 * nerfServer/core/views.py:79 -- use one handle per request thread or a lock).
 */
#ifndef EVPP_B200_H_
#define EVPP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVPP_ABI_VERSION 2

typedef enum evpp_status {
  EVPP_OK = 0,
  EVPP_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  EVPP_ERR_CUDA = -2,      /* a CUDA runtime call or kernel failed */
  EVPP_ERR_STATE = -3,     /* e.g. weights not loaded */
  EVPP_ERR_NO_DEVICE = -4, /* no usable sm_100 device: there is NO CPU fallback */
  EVPP_ERR_NONFINITE = -5  /* a score came out inf / NaN (16-bit activations overflowed): retry in bf16 */
} evpp_status;

typedef enum evpp_precision {
  EVPP_PREC_FP32 = 0,  /* fp32 FFMA MLP (exact mode; used for the top-K re-score)            */
  EVPP_PREC_FP16 = 1,  /* tcgen05.mma kind::f16, fp16 operands, fp32 accumulation in TMEM   */
  EVPP_PREC_BF16 = 2,  /* tcgen05.mma kind::f16, bf16 operands, fp32 accumulation in TMEM   */
  EVPP_PREC_FP16X2 = 3, /* tcgen05, weights as an fp16 (hi, lo) pair = 22 bits, activations fp16: 2 MMAs per product */
  EVPP_PREC_FP16X3 = 4  /* tcgen05, weights AND activations as (hi, lo) pairs, 3 MMAs per product, fp32 accumulate:
                           fp32-class results on the tensor cores (the exact path of the NBV re-score)            */
} evpp_precision;

typedef enum evpp_net { EVPP_NET_COARSE = 0, EVPP_NET_FINE = 1 } evpp_net;

/* Mirrors the reference's flags: nerfServer/config.txt:6-13, core/nerf_configs.py:17-58,
 * Controller.near/far nerfServer/core/nerf.py:100-101. */
typedef struct evpp_config {
  int32_t n_samples;       /* N_samples (64)                                   */
  int32_t n_importance;    /* N_importance (128; 0 = coarse only)              */
  float   near_plane;      /* Controller.near (0.5)                            */
  float   far_plane;       /* Controller.far (80 live, 6 offline)              */
  int32_t white_bkgd;      /* white_bkgd (1)                                   */
  int32_t lindisp;         /* lindisp (0)                                      */
  int32_t precision;       /* evpp_precision of the bulk pass                  */
  int32_t max_chunk_rays;  /* rays per internal chunk (0 = default 131072)     */
} evpp_config;

typedef struct evpp_handle evpp_handle;

const char* evpp_last_error(void);
int evpp_abi_version(void);

/* Number of CUDA devices usable by the engine (compute capability 10.x). */
int evpp_device_count(void);

/* Replaces create_nerf's model construction (run_nerf.py:244-341) for inference. */
int evpp_create(int device_ordinal, const evpp_config* cfg, evpp_handle** out);
int evpp_destroy(evpp_handle* h);

/* Replaces Controller.copy_model (nerf.py:718-757): snapshot one network's weights.
 * `packed` (host or device memory) holds the fp32 tensors of NeRF.state_dict() in its
 * own order (run_nerf_helpers.py:90-104): pts_linears.{0..7}.{weight,bias},
 * views_linears.0, feature_linear, alpha_linear, rgb_linear, uncertainty_linear;
 * weight = [out,in] row-major.  offsets[i] = element offset of tensor i in `packed`,
 * n_tensors must be 26.  The engine keeps its own repacked copies. */
int evpp_load_nerf_weights(evpp_handle* h, int which_net, const float* packed,
                           const int64_t* offsets, int n_tensors);

/* Replaces Controller.get_uncertainty (nerf.py:266-309) = get_rays (run_nerf_helpers.py:
 * 173-182) + render (run_nerf.py:131-196) + sum(beta^2)/R.
 *   c2w      device [V,3,4] fp32 camera-to-world (views.get_pose, views.py:141-150)
 *   pix_idx  device [V,R] int32 flat pixel indices j*W+i, or NULL = full H*W grid (R ignored)
 *   scores   device [V] fp32 out */
int evpp_score_views(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy,
                     float cx, float cy, const int32_t* pix_idx, int R, float* scores,
                     void* stream);

/* Same at an explicit precision (< 0 = cfg.precision) -- e.g. EVPP_PREC_FP16X3 to re-score the contenders of a sweep
 * whose poses are already on the device. */
int evpp_score_views_ex(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy,
                        float cx, float cy, const int32_t* pix_idx, int R, int precision,
                        float* scores, void* stream);

/* How many scores of the LAST device-side scoring call on this handle were not finite (inf / NaN: 16-bit
 * activations overflowed).  Synchronises `stream`.  The host entry points check this themselves and return
 * EVPP_ERR_NONFINITE. */
int evpp_get_nonfinite(evpp_handle* h, int32_t* count, void* stream);

/* Same with HOST buffers (what the HTTP view nerfServer/core/views.py:94-113 holds):
 * host->device copy of poses / pixel indices and device->host copy of the scores included. */
int evpp_score_views_host(evpp_handle* h, const float* c2w_host, int V, int H, int W, float fx,
                          float fy, float cx, float cy, const int32_t* pix_idx_host, int R,
                          float* scores_host, void* stream);

/* Re-score a subset of views with the fp32 (exact) MLP regardless of cfg.precision; used to
 * make the arg-max of plannerServer_{Object,Room}/core/interface.py:115-116 identical to the fp32
 * reference.  view_ids: host [K] indices into c2w; scores_host: host [K] out. */
int evpp_rescore_views_fp32(evpp_handle* h, const float* c2w_host, const int32_t* view_ids, int K,
                            int H, int W, float fx, float fy, float cx, float cy,
                            const int32_t* pix_idx_host, int R, float* scores_host, void* stream);

/* The same re-score at any precision.  EVPP_PREC_FP16X3 is the tensor-core exact path: fp16 (hi, lo) split of
 * weights and activations, three tcgen05 MMAs per product, fp32 accumulation -- fp32-class scores (<= 2e-5
 * relative to the reference) at about a third of the bulk pass's rate instead of the FFMA kernel's 1/26. */
int evpp_rescore_views(evpp_handle* h, const float* c2w_host, const int32_t* view_ids, int K, int H, int W,
                       float fx, float fy, float cx, float cy, const int32_t* pix_idx_host, int R,
                       int precision, float* scores_host, void* stream);

/* How much of the network scoring calls (evpp_score_views*, evpp_rescore_views*) evaluate.  Controller.get_uncertainty's
 * result depends on raw[..., 4] of the final pass only (nerf.py:307-309):
 *   mode 0 (default)  the coarse network stops at its sigma, the fine network skips its alpha / rgb heads, and
 *                     feature_linear is multiplied into views_linears.0 on the host -- the reference applies no
 *                     activation between them (run_nerf_helpers.py:119-122), so the two are ONE linear map; the pass is
 *                     the same function of its inputs, equal to the reference's layer order at rounding level;
 *   mode 2            the same dead-output elimination in the reference's layer order: scores bit-identical to mode 1;
 *   mode 1            the full network like render() (parity tests, like-for-like timing). */
int evpp_set_score_mode(evpp_handle* h, int mode);

/* evpp_render_rays(_ex) on the tensor-core precisions evaluates feature_linear multiplied into views_linears.0 as well
 * (default 1: the same function with nine GEMM layers, outputs equal to the reference's layer order at the operand
 * type's rounding level); 0 = the reference's ten layers.  Stage capture and the fp32 FFMA kernel always use ten. */
int evpp_set_render_mode(evpp_handle* h, int fold_feature_layer);

/* Replaces render(rays=...) / batchify_rays / render_rays (run_nerf.py:116-196, 391-509) with
 * use_viewdirs=True, ndc=False, perturb=0, raw_noise_std=0.  All device pointers; any output
 * may be NULL.  S_f = n_samples + n_importance (or n_samples when n_importance == 0).
 *   rgb [N,3] disp [N] acc [N] depth [N] unc [N,S_f] (= raw[...,4], 'uncertainty_map')
 *   raw [N,S_f,5]; coarse-pass extras rgb0 [N,3] disp0 acc0 depth0 [N] unc0 [N,n_samples];
 *   z_std [N] */
typedef struct evpp_render_out {
  float* rgb; float* disp; float* acc; float* depth; float* unc; float* raw;
  float* rgb0; float* disp0; float* acc0; float* depth0; float* unc0; float* z_std;
} evpp_render_out;

int evpp_render_rays(evpp_handle* h, const float* rays_o, const float* rays_d, int64_t N,
                     float near_plane, float far_plane, int precision,
                     const evpp_render_out* out, void* stream);

/* Same with explicit view directions (device [N,3], used as given; NULL = rays_d / ||rays_d|| like render()).
 * Replaces batchify_rays on a caller-built [N,11] ray batch: Controller.get_rays_depth (nerf.py:353-392) appends
 * the raw, un-normalised directions as viewdirs. */
int evpp_render_rays_ex(evpp_handle* h, const float* rays_o, const float* rays_d, const float* viewdirs,
                        int64_t N, float near_plane, float far_plane, int precision,
                        const evpp_render_out* out, void* stream);

/* get_rays (run_nerf_helpers.py:173-182) + viewdir normalisation (run_nerf.py:170).
 * device out: rays_o/rays_d/viewdirs [V*R,3] (any may be NULL). */
int evpp_get_rays(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy,
                  float cx, float cy, const int32_t* pix_idx, int R, float* rays_o, float* rays_d,
                  float* viewdirs, void* stream);

/* The MLP alone: embed (run_nerf_helpers.py:22-55) + NeRF.forward (:108-134) on explicit
 * points.  pts/dirs device [M,3] (dirs already normalised), out device [M,5] =
 * [r,g,b,sigma_raw,beta]. M is padded internally. */
int evpp_mlp_forward(evpp_handle* h, int which_net, int precision, const float* pts,
                     const float* dirs, int64_t M, float* out, void* stream);

/* Same, and additionally dumps the fp32 activations one layer of the tensor-core path hands to the
 * next (after bias / ReLU, before the 16-bit conversion): layer 0..7 = pts_linears (run_nerf_helpers.py:
 * 111-116), 8 = feature_linear (:119), 9 = views_linears.0 (:122-124, 128 wide).  acts: device
 * [M,256] (layer 9 fills the first 128 columns of each row).  Parity-test hook. */
int evpp_mlp_layer_dump(evpp_handle* h, int which_net, int precision, const float* pts, const float* dirs,
                        int64_t M, int layer, float* out, float* acts, void* stream);

/* Diagnostics: runs the tensor-core MLP like evpp_mlp_forward and records SM-clock stamps of the third
 * tile processed by CTA 0 into trace (device int64[128]): [4l+0..3] = MMA issuer at layer l (layer start,
 * first K-block issued, last K-block issued, accumulator committed); [40+3l+0..2] = epilogue warp at layer
 * l (accumulator ready, all chunks stored, staging window done).  Needs M >= 3 * 128 * #SMs rows. */
int evpp_mlp_trace(evpp_handle* h, int which_net, int precision, const float* pts, const float* dirs,
                   int64_t M, float* out, int64_t* trace, void* stream);

/* Per-stage dumps of the LAST evpp_render_rays call's final chunk, for parity tests
 * (SURVEY.md appendix B).  `out` is a device buffer of at least *count elements
 * (call with out == NULL to query count).  Stages: */
typedef enum evpp_stage {
  EVPP_STAGE_Z_COARSE = 0,   /* float [n,n_samples]                 */
  EVPP_STAGE_RAW_COARSE = 1, /* float [n,n_samples,5]               */
  EVPP_STAGE_W_COARSE = 2,   /* float [n,n_samples] weights         */
  EVPP_STAGE_CDF = 3,        /* float [n,n_samples-1]               */
  EVPP_STAGE_INDS = 4,       /* int32 [n,n_importance] searchsorted */
  EVPP_STAGE_Z_SAMPLES = 5,  /* float [n,n_importance]              */
  EVPP_STAGE_Z_FINE = 6      /* float [n,S_f] sorted                */
} evpp_stage;
/* Enable (1) / disable (0) allocation and filling of the capture-only stages (W_COARSE, CDF,
 * INDS, Z_SAMPLES). Off by default. */
int evpp_set_stage_capture(evpp_handle* h, int enabled);
int evpp_stage_dump(evpp_handle* h, int stage, void* out, int64_t* count, void* stream);

/* Stage-level entry points on caller-provided tensors (device pointers), for parity tests that
 * feed the reference's own intermediate tensors to one stage at a time:
 * evpp_composite = raw2outputs (run_nerf.py:344-388) on raw [n,S,5], z [n,S], rays_d [n,3];
 *   outputs (any may be NULL): rgb [n,3], disp/acc/depth [n], weights [n,S], beta2 [n] (sum of
 *   raw[...,4]^2 per ray, the numerator of nerf.py:307-309).
 * evpp_resample = raw2outputs on the coarse pass + sample_pdf(det=True)
 *   (run_nerf_helpers.py:216-259) + sort(cat) (run_nerf.py:479-483) -> z_fine [n,S+n_importance];
 *   optional dumps cdf [n,S-1], inds int32 [n,n_importance], z_samples [n,n_importance]. */
int evpp_composite(evpp_handle* h, const float* raw, const float* z, const float* rays_d, int n, int S,
                   float* rgb, float* disp, float* acc, float* depth, float* weights, float* beta2,
                   void* stream);
int evpp_resample(evpp_handle* h, const float* raw_coarse, const float* z_coarse, const float* rays_d,
                  int n, float* z_fine, float* cdf, int32_t* inds, float* z_samples, void* stream);

/* Next-best-view selection tail (plannerServer_Object/core/vpp.py:199-220 +
 * interface.py:115-116): score*=weight (NULL = 1), per-location max over consecutive groups
 * of dirs_per_location views, normalise by the global max, arg-max (highest index on exact
 * ties).  Device in/out; norm_scores [V/dirs_per_location] and best_dir [same, int32] may be
 * NULL; winner_view (device int32[1]) receives the index into the V input views. */
int evpp_select_nbv(evpp_handle* h, const float* scores, const float* weight, int V,
                    int dirs_per_location, float* norm_scores, int32_t* best_dir,
                    int32_t* winner_view, void* stream);

/* Random pixel subsets on the device: pix_idx (device int32 [V,R]) receives R DISTINCT pixel ids in [0, H*W) per
 * view, the input evpp_score_views / evpp_tsdf_score_views take.  Replaces the per-pose
 * np.random.choice(H*W, R, replace=False) of Controller.get_uncertainty (nerf.py:284-290) and
 * get_uncertainty_tsdf (tsdf_gpu.py:497), which shuffles all H*W ids on the host for every pose.  Same sampling
 * model, a different random stream (keyed by `seed`); callers that need the reference's NumPy stream pass their own
 * pix_idx instead. */
int evpp_sample_pixels(evpp_handle* h, uint64_t seed, int V, int R, int n_pixels, int32_t* pix_idx, void* stream);

/* ---- Planner surrogate g_phi ----
 * The small MLP the planners fit to the (location -> normalised gain) rows of one planning step and then query
 * inside weighted A* (class MLP, plannerServer_Object/core/vpp.py:26-67; identical in plannerServer_Room).
 * params: device float[evpp_surrogate_param_count()] = the module's state_dict flattened in its own order
 * (layer1.weight [64,3], layer1.bias, layer2_1 .. layer2_8 (weight [64,64], bias), layer3.weight [1,64],
 * layer3.bias); layer2_2 is declared by the reference but never used by forward(): it is carried through untouched.
 *
 * evpp_surrogate_fit: Viewpath_planner.train_only (vpp.py:619-654): `epochs` full-batch steps of torch.optim.Adam
 * (lr, betas (0.9, 0.999), eps 1e-8, fresh moments) on F.mse_loss over N <= 128 points; xyz device [N,3], gain device
 * [N]; params is updated in place; loss_curve (device float[epochs], may be NULL) receives the loss BEFORE each
 * step, like the reference's mlp_loss list.  One kernel launch.
 * evpp_surrogate_query: mlp(x) for M points (vpp.py:233,242,487; localPlanner.py:126 via env.get_gain). */
int evpp_surrogate_param_count(void);
int evpp_surrogate_fit(evpp_handle* h, float* params, const float* xyz, const float* gain, int N, int epochs, float lr,
                       float* loss_curve, void* stream);
int evpp_surrogate_query(evpp_handle* h, const float* params, const float* xyz, int64_t M, float* out, void* stream);

/* ---- Stage-1 candidate scorer: occupancy-state ray cast through the planner's TSDF state volume ----
 * Replaces TSDF.render_rays + isectSegAABB_gpu + get_uncertainty_tsdf
 * (plannerServer_Object/core/tsdf_gpu.py:198-354, :67-114, :473-530; the Room copy differs only in the
 * two coefficients, plannerServer_Room/core/tsdf_gpu.py:281,298).
 *
 * evpp_tsdf_set_volume: snapshot of State_vol_origin (float [X,Y,Z]: 0 unknown, 1 empty, 2 occupied) and
 *   Nray_vol (float [X,Y,Z]), host or device memory; dims = tsdf_vol._vol_dim, origin = _vol_origin,
 *   voxel_size = _voxel_size (fusion.py:32-39); aabb_min/max = aabb_bnds columns (tsdf_gpu.py:31);
 *   near/far (tsdf_gpu.py:39); n_samples = int((far-near)/voxel_res)+1 evaluated in double by the caller
 *   (tsdf_gpu.py:41: 56, whereas the fp32 voxel size would give 55); c_unknown / c_surface = 0.2 / 0.1 (Object) or unknown_cof / surface_cof.
 * evpp_tsdf_render_rays: TSDF.render on device rays; any output may be NULL.  uncer/depth [N] float,
 *   n_unknown/first_hit [N] int32 (count of state==0 samples, index of the first state==2 sample, 0 = none),
 *   voxel_idx [N,n_samples,3] int32 (after the reference's out-of-bounds rewrite).
 * evpp_tsdf_score_views(_host): get_uncertainty_tsdf -- per view mean uncertainty and mean valid depth;
 *   pix_idx [V,R] int32 flat pixel indices (row*W + col) or NULL = all H*W pixels. */
int evpp_tsdf_set_volume(evpp_handle* h, const float* state, const float* nray, const int32_t* dims,
                         const float* origin, float voxel_size, const float* aabb_min, const float* aabb_max,
                         float near_plane, float far_plane, int32_t n_samples, float c_unknown, float c_surface);
int evpp_tsdf_render_rays(evpp_handle* h, const float* rays_o, const float* rays_d, int64_t N, float* uncer,
                          float* depth, int32_t* n_unknown, int32_t* first_hit, int32_t* voxel_idx, void* stream);
int evpp_tsdf_score_views(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy, float cx,
                          float cy, const int32_t* pix_idx, int R, float* view_uncer, float* view_depth,
                          void* stream);
int evpp_tsdf_score_views_host(evpp_handle* h, const float* c2w_host, int V, int H, int W, float fx, float fy,
                               float cx, float cy, const int32_t* pix_idx_host, int R, float* view_uncer_host,
                               float* view_depth_host, void* stream);

/* ---- NGP-mode scorer: occupancy-bitfield marching + multiresolution hash grid + small MLP ----------
 * BASELINE.json's north_star mandates these stages; the reference has NO counterpart (its README links
 * ashawkey/torch-ngp and environment.yml:271 pins tinycudann==1.7, but nothing imports them), so there is
 * no reference interface to cite: the entry points mirror the reference-parity ones above
 * (evpp_score_views <-> Controller.get_uncertainty nerf.py:266-309, evpp_render_rays <-> render run_nerf.py:131)
 * and the NeurAR parts keep the reference's definitions (beta = exp(linear(h)) run_nerf_helpers.py:128-129,
 * score = sum(beta^2)/rays nerf.py:307-309).  Arithmetic: torch-ngp gridencoder / instant-ngp conventions,
 * restated in oracle/ngp_oracle.py ("parity unpinned"). */
typedef struct evpp_ngp_config {
  float aabb_min[3], aabb_max[3];   /* scene box the occupancy grid and the hash grid span               */
  int32_t grid_res;                 /* occupancy resolution G (bitfield of G^3 bits, x-major cells)      */
  float near_plane, far_plane;      /* ray segment clip, like Controller.near / far                      */
  float dt;                         /* uniform step: t_i = t0 + (i + 0.5) * dt                           */
  int32_t max_steps, max_samples;   /* steps walked / samples kept per ray at most                       */
  int32_t n_levels;                 /* hash-grid levels (16; 2 features each)                            */
  int32_t white_bkgd;
  /* march_mode 0: uniform steps over one x-major G^3 bitfield, warp-per-ray ballot / popc compaction (above).
   * march_mode 1: the published torch-ngp `raymarching` rule: the (cubic) scene box is [-bound, bound]^3 with
   *   bound = 2^(cascades-1), `cascades` nested Morton-ordered grid_res^3 bitfields (cascades*G^3 bits),
   *   dt = clamp(t*dt_gamma, 2 sqrt3 / max_steps, 2 sqrt3 * 2^(cascades-1) / G) in normalised units, occupied cell ->
   *   sample + step dt, empty cell -> skip to its far face; rays start at max(box entry, min_near).  near_plane,
   *   far_plane and dt are not used in this mode. */
  int32_t march_mode;
  int32_t cascades;
  float dt_gamma;
  float min_near;
} evpp_ngp_config;

typedef struct evpp_ngp_out {       /* device pointers, any may be NULL                                  */
  float* rgb; float* depth; float* acc; float* beta2;   /* per ray [N,3] / [N]                           */
  int32_t* counts;                  /* kept samples per ray [N] (integer, bit-exact vs the restatement)  */
  float* t; float* raw; float* enc; /* per kept sample, packed ray after ray: [cap], [cap,5], [cap,32]   */
  int64_t sample_capacity;          /* capacity of t / raw / enc / delta / cell in samples               */
  int64_t total_so_far;             /* out: kept samples of the call                                     */
  float* delta;                     /* march_mode 1: step length of every kept sample [cap] (normalised) */
  int32_t* cell;                    /* march_mode 1: bitfield index (cascade*G^3 + Morton) of the sample's cell [cap] */
} evpp_ngp_out;

/* level tables computed by the caller (torch-ngp gridencoder): scale_l = 2^(l*log2 b)*base - 1,
 * res_l = ceil(scale_l)+1, size_l = min(2^log2_hashmap, (res_l+1)^3) rounded up to 8, offsets cumulative */
int evpp_ngp_configure(evpp_handle* h, const evpp_ngp_config* cfg, const float* level_scale, const int32_t* level_res,
                       const int64_t* level_offset, const int64_t* level_size);
int evpp_ngp_set_occupancy(evpp_handle* h, const uint8_t* bitfield, int64_t nbytes);
/* table: fp16 [entries][2]; mlp_packed: fp32 sigma0 [64][32], sigma1 [16][64], color0 [64][32] (31 inputs
 * + zero column), color1 [64][64], heads [4][64] (r,g,b,uncertainty), uncertainty bias [1] */
int evpp_ngp_load_params(evpp_handle* h, const uint16_t* table_fp16, int64_t n_entries, const float* mlp_packed,
                         int64_t n_mlp_floats);
/* TensoRF VM backbone instead of the hash grid (BASELINE configs[3]; same marching / compositing / scoring calls):
 * planes fp16 [3][R][R][64] indexed [i][y][x][c], lines fp16 [3][R][64] (channels 0..15 density, 16..63 appearance;
 * plane i spans axes (0,1) (0,2) (1,2), line i runs along axis 2, 1, 0), mlp_packed fp32: basis [27][144],
 * color0 [64][43] (27 basis + 16 SH inputs), color1 [64][64], heads [4][64], uncertainty bias [1] */
int evpp_tensorf_load_params(evpp_handle* h, const uint16_t* planes_fp16, const uint16_t* lines_fp16, int32_t res,
                             const float* mlp_packed, int64_t n_mlp_floats);
/* hash-grid encoding alone: u device [M,3] in [0,1] -> enc [M,32] fp32, indices [M,16,8] int32 (optional) */
int evpp_ngp_encode(evpp_handle* h, const float* u, int64_t M, float* enc, int32_t* indices, void* stream);
int evpp_ngp_render_rays(evpp_handle* h, const float* rays_o, const float* rays_d, int64_t N, evpp_ngp_out* out,
                         void* stream);
int evpp_ngp_score_views(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy, float cx, float cy,
                         const int32_t* pix_idx, int R, float* scores, void* stream);
int evpp_ngp_score_views_host(evpp_handle* h, const float* c2w_host, int V, int H, int W, float fx, float fy, float cx,
                              float cy, const int32_t* pix_idx_host, int R, float* scores_host, void* stream);
int evpp_ngp_get_sample_count(evpp_handle* h, int64_t* samples);

/* Counters since handle creation: kernels launched by this library and algorithmic MLP
 * evaluations issued (for bench.py's gpu_launches and roofline accounting). */
int evpp_get_counters(evpp_handle* h, int64_t* kernel_launches, int64_t* mlp_evals);

/* Self-check (compute-sanitizer is not available on the development pool): every device buffer the engine owns is
 * allocated with 256-byte guard bands of a known pattern on both sides; this synchronises the device and counts the
 * guard bytes any kernel has overwritten since the buffers were allocated (0 = no out-of-bounds write). */
int evpp_check_guards(evpp_handle* h, int32_t* corrupted_bytes, int32_t* buffers_checked);

/* Self-check hook: fills every per-call workspace buffer with 0xFF bytes (NaN as fp32), so that a kernel reading an
 * element that was not written earlier in the same call produces a visibly wrong result. */
int evpp_poison_workspace(evpp_handle* h);

/* Average duration in ms of the MLP kernel launches recorded (CUDA events on the launching
 * stream) since the last reset; enable with evpp_set_profiling(h, 1). Synchronises. */
int evpp_set_profiling(evpp_handle* h, int enabled);
int evpp_get_mlp_time_ms(evpp_handle* h, double* total_ms, int64_t* launches, int64_t* evals);
/* Same, plus the FLOPs those launches really executed (a score-only coarse evaluation is 982,528 FLOP, a
 * score-only fine evaluation 1,054,720 with the folded feature layer (1,185,792 without), a full evaluation
 * 1,187,072: SURVEY.md section 8). */
int evpp_get_mlp_profile(evpp_handle* h, double* total_ms, int64_t* launches, int64_t* evals, double* flops);

#ifdef __cplusplus
}
#endif
#endif /* EVPP_B200_H_ */
