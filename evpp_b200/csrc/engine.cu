// C ABI of the evpp_b200 engine (see include/evpp_b200.h): handle, weight snapshot, chunked
// orchestration of the per-ray pipeline
//   get_rays -> coarse z -> MLP(coarse) -> composite+resample -> MLP(fine) -> composite+beta^2
//   -> per-view reduction
// on the caller's stream.  There is deliberately no CPU path: without an sm_100 device every entry
// point fails with EVPP_ERR_NO_DEVICE.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in nsys / ncu --nvtx, cost nothing without a tool attached

#include "common.cuh"
#include "sampling.cuh"
#include "tsdf.cuh"
#include "ngp.cuh"
#include "surrogate.cuh"

using namespace evpp;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(EVPP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                  __LINE__);                                                                        \
  } while (0)

// Device buffer with guard bands.  compute-sanitizer is closed on the GPU pool this engine is developed on, so every
// buffer the engine owns carries kGuard bytes of a known pattern before and after the payload; evpp_check_guards()
// verifies them (a write past either end of any workspace, by any kernel, shows up there) -- see tests and
// tools/selfcheck.py.
constexpr size_t kGuard = 256;
constexpr int kGuardByte = 0xA5;
struct DevBuf {
  void* p = nullptr;       // payload (256-byte aligned: cudaMalloc alignment + kGuard)
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    release();
    void* base = nullptr;
    cudaError_t e = cudaMalloc(&base, need + 2 * kGuard);
    if (e != cudaSuccess) return (int)e;
    cudaMemset(base, kGuardByte, kGuard);
    cudaMemset(static_cast<char*>(base) + kGuard + need, kGuardByte, kGuard);
    p = static_cast<char*>(base) + kGuard;
    bytes = need;
    return 0;
  }
  void release() {
    if (p) cudaFree(static_cast<char*>(p) - kGuard);
    p = nullptr;
    bytes = 0;
  }
  // number of corrupted guard bytes (device must be idle)
  int check() const {
    if (!p) return 0;
    unsigned char host[2 * kGuard];
    if (cudaMemcpy(host, static_cast<char*>(p) - kGuard, kGuard, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (cudaMemcpy(host + kGuard, static_cast<char*>(p) + bytes, kGuard, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    int bad = 0;
    for (size_t i = 0; i < 2 * kGuard; ++i) bad += host[i] != kGuardByte;
    return bad;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int kNumPrec = 5;   // evpp_precision values

// FLOPs the MLP kernel really executes per evaluation: the full network, the sigma-only
// coarse pass (layers 0..7 + alpha head) and the beta-only fine pass (no alpha / rgb heads)
constexpr double kFlopFull = kFlopPerEval;
constexpr double kFlopSigmaOnly = 2.0 * (63 * 256 + 4 * 256 * 256 + 319 * 256 + 2 * 256 * 256 + 256);
constexpr double kFlopBetaOnly = kFlopPerEval - 2.0 * (256 + 3 * 128);
constexpr double kFlopBetaOnlyFolded = kFlopBetaOnly - 2.0 * 256 * 256;      // feature_linear multiplied into the view layer
constexpr double kFlopFullFolded = kFlopFull - 2.0 * 256 * 256;

// Every entry point runs on the handle's device and leaves the caller's current device as it found it
// (one process may drive several GPUs, nerf.py:269-273).
// NVTX range over a scope (tracing).  Names: evpp:score_views, evpp:chunk, evpp:mlp_coarse, ...
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};
#define EVPP_ON_DEVICE(dev)      \
  DeviceGuard _guard(dev);       \
  CUDA_TRY(_guard.err)

struct NetWeights {
  bool loaded = false;
  DevBuf simt_stream, bias, head_w, head_b, tc_image[kNumPrec][2];   // tc_image[precision 1..4][folded feature layer]
  bool tc_ready[kNumPrec][2] = {};
  TcConsts tc_consts[2] = {};
  std::vector<std::vector<float>> host;                   // fp32 state_dict tensors (for lazy tc repack)
};

}  // namespace

struct evpp_handle {
  int device = 0;
  int num_sms = 148;
  evpp_config cfg{};
  int Sc = 64, Ni = 128, Sf = 192;
  NetWeights net[2];
  DevBuf t_vals, u_vals;
  // chunk workspace
  int chunk_cap = 0;
  DevBuf rays_o, rays_d, viewdirs, z_c, raw_c, z_f, raw_f;
  DevBuf beta2, st_c2w, st_pix, st_scores, nonfinite;
  bool render_fold = true;   // evpp_set_render_mode: render() evaluates feature_linear folded into the view layer
  int score_mode = 0;   // evpp_set_score_mode: 0 score-only passes + folded feature layer, 1 full network like render(),
                        // 2 score-only passes in the reference's layer order (bit-identical to 1)
  // stage capture
  bool capture = false;
  DevBuf w_c, cdf, inds, zs;
  int last_n = 0;
  // stage-1 TSDF scorer
  bool tsdf_ready = false;
  TsdfVolume tsdf{};
  DevBuf tsdf_state, tsdf_nray, tsdf_t, tsdf_ro, tsdf_rd, tsdf_u, tsdf_d, tsdf_vu, tsdf_vd;
  // NGP-mode scorer
  bool ngp_configured = false, ngp_has_occ = false, ngp_has_params = false, ngp_tensorf = false;
  DevBuf tf_planes, tf_lines, tf_mlp;
  NgpParams ngp{};
  DevBuf ngp_table, ngp_mlp, ngp_bits, ngp_ro, ngp_rd, ngp_counts, ngp_offsets, ngp_total, sg_ws, ngp_masks, ngp_tiles, ngp_t0, ngp_t, ngp_ray, ngp_raw,
      ngp_beta2, ngp_scores, ngp_delta, ngp_cell, ngp_counter;
  int64_t ngp_samples = 0;       // kept samples since handle creation (host-visible part; the device counter holds the rest)
  int ngp_chunk_rays = 1 << 20;  // rays per marching chunk (one count -> scan -> fill round trip each)
  // counters / profiling
  int64_t launches = 0, evals = 0;
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  std::vector<int64_t> prof_evals;
  std::vector<double> prof_flops;
  double prof_ms = 0.0, prof_total_flops = 0.0;
  int64_t prof_launches = 0, prof_total_evals = 0;
};

namespace {

// torch.linspace on CPU (fp32): step = (end-start)/(steps-1); first half start + step*i, second
// half end - step*(steps-1-i), each with a fused multiply-add (verified against torch 2.11).
std::vector<float> torch_linspace(float start, float end, int steps) {
  std::vector<float> v(steps);
  if (steps == 1) { v[0] = start; return v; }
  const float step = (end - start) / (float)(steps - 1);
  const int half = steps / 2;
  for (int i = 0; i < steps; ++i)
    v[i] = i < half ? std::fmaf(step, (float)i, start) : std::fmaf(-step, (float)(steps - i - 1), end);
  return v;
}

const int kTensorSizes[26] = {
    256 * 63, 256, 256 * 256, 256, 256 * 256, 256, 256 * 256, 256, 256 * 256, 256, 256 * 319, 256,
    256 * 256, 256, 256 * 256, 256, 128 * 283, 128, 256 * 256, 256, 1 * 256, 1, 3 * 128, 3, 1 * 128, 1};

int ensure_workspace(evpp_handle* h, int cap) {
  if (cap <= h->chunk_cap) return 0;
  const size_t c = (size_t)cap;
  int e = 0;
  e |= h->rays_o.ensure(c * 3 * 4);
  e |= h->rays_d.ensure(c * 3 * 4);
  e |= h->viewdirs.ensure(c * 3 * 4);
  e |= h->z_c.ensure(c * h->Sc * 4);
  e |= h->raw_c.ensure(c * h->Sc * 5 * 4);
  if (h->Ni > 0) {
    e |= h->z_f.ensure(c * h->Sf * 4);
    e |= h->raw_f.ensure(c * h->Sf * 5 * 4);
  }
  if (h->capture) {
    e |= h->w_c.ensure(c * h->Sc * 4);
    e |= h->cdf.ensure(c * (h->Sc - 1) * 4);
    e |= h->inds.ensure(c * (h->Ni > 0 ? h->Ni : 1) * 4);
    e |= h->zs.ensure(c * (h->Ni > 0 ? h->Ni : 1) * 4);
  }
  if (e) return fail(EVPP_ERR_CUDA, "workspace allocation for %d rays failed: %s", cap,
                     cudaGetErrorString(cudaGetLastError()));
  h->chunk_cap = cap;
  return 0;
}

int ensure_tc(evpp_handle* h, int which, int prec, bool fold = false) {
  NetWeights& nw = h->net[which];
  if (nw.tc_ready[prec][fold]) return 0;
  std::vector<uint8_t> image;
  tc_pack_weights(nw.host, prec, fold, image, nw.tc_consts[fold]);
  if (nw.tc_image[prec][fold].ensure(image.size())) return fail(EVPP_ERR_CUDA, "tc weight image alloc failed");
  CUDA_TRY(cudaMemcpy(nw.tc_image[prec][fold].p, image.data(), image.size(), cudaMemcpyHostToDevice));
  nw.tc_ready[prec][fold] = true;
  return 0;
}

SimtWeights simt_of(const NetWeights& nw) {
  return SimtWeights{nw.simt_stream.as<float>(), nw.bias.as<float>(), nw.head_w.as<float>(), nw.head_b.as<float>()};
}
TcWeights tc_of(const NetWeights& nw, int prec, float* dbg = nullptr, int dbg_layer = -1, long long* trace = nullptr,
                bool fold = false) {
  return TcWeights{nw.tc_image[prec][fold].p, &nw.tc_consts[fold], prec, dbg, dbg_layer, trace, fold ? 1 : 0};
}

// heads: 0 = raw [n,S,5]; 1 = sigma only, 2 = beta only, 3 = beta only with the folded feature layer -> raw [n,S]
// (tensor-core precisions only)
int run_mlp(evpp_handle* h, int which, int prec, const RayBatch& rb, float* raw, cudaStream_t st, int heads = 0) {
  NetWeights& nw = h->net[which];
  if (!nw.loaded) return fail(EVPP_ERR_STATE, "weights of net %d not loaded (evpp_load_nerf_weights)", which);
  NvtxRange nvtx_mlp(which == EVPP_NET_COARSE ? "evpp:mlp_coarse" : "evpp:mlp_fine");
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->profiling) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
  }
  int nl = 0;
  if (prec == EVPP_PREC_FP32) {
    if (heads != 0) return fail(EVPP_ERR_INVALID, "internal: the fp32 FFMA kernel has no score-only variant");
    nl = launch_mlp_simt(simt_of(nw), rb, raw, st);
  } else {
    const bool fold = heads == 3 || heads == 4;
    int rc = ensure_tc(h, which, prec, fold);
    if (rc) return rc;
    nl = launch_mlp_tc(tc_of(nw, prec, nullptr, -1, nullptr, fold), rb, raw, h->num_sms, st, heads);
    if (nl < 0) return fail(EVPP_ERR_CUDA, "tcgen05 MLP launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (h->profiling) {
    cudaEventRecord(e1, st);
    h->prof_events.emplace_back(e0, e1);
    h->prof_evals.push_back((int64_t)rb.n * rb.S);
    h->prof_flops.push_back((double)rb.n * rb.S * (heads == 1 ? kFlopSigmaOnly : heads == 2 ? kFlopBetaOnly : heads == 3 ? kFlopBetaOnlyFolded : heads == 4 ? kFlopFullFolded : kFlopFull));
  }
  h->launches += nl;
  h->evals += (int64_t)rb.n * rb.S;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// One chunk of rays already in h->rays_*: coarse pass, resample, fine pass, composite.
// Score-only calls (out == nullptr: only the per-ray sum of beta^2 leaves the chunk) skip what the score does not
// depend on -- Controller.get_uncertainty reads raw[..., 4] of the final pass and nothing else (nerf.py:307-309):
// the coarse pass stops at its sigma (all sample_pdf needs), the final pass skips the alpha / rgb heads, and the
// two compositing kernels shrink to the weights and to the beta^2 chain.  Every value that reaches the score is
// computed by the same instructions in the same order as in the full path: the scores are bit-identical.
int render_chunk(evpp_handle* h, int n, float near_plane, float far_plane, int prec, const evpp_render_out* out,
                 long long off, float* beta2, cudaStream_t st) {
  const int Sc = h->Sc, Ni = h->Ni, Sf = h->Sf;
  NvtxRange nvtx_chunk("evpp:chunk");
  const bool score_only = !out && beta2 && !h->capture && h->score_mode != 1 && prec != EVPP_PREC_FP32;
  const int beta_heads = h->score_mode == 0 ? 3 : 2;      // the beta-only pass: with / without the folded feature layer
  // all five outputs (render(), or a scoring call in score mode 1 = like-for-like with the reference's layer order)
  const int full_heads = (out && h->render_fold && !h->capture && prec != EVPP_PREC_FP32) ? 4 : 0;
  h->launches += launch_coarse_z(h->t_vals.as<float>(), n, Sc, near_plane, far_plane, h->cfg.lindisp,
                                 h->z_c.as<float>(), st);
  RayBatch rb{h->rays_o.as<float>(), h->rays_d.as<float>(), h->viewdirs.as<float>(), h->z_c.as<float>(), n, Sc};
  int rc = run_mlp(h, EVPP_NET_COARSE, prec, rb, h->raw_c.as<float>(), st, score_only ? (Ni > 0 ? 1 : beta_heads) : full_heads);
  if (rc) return rc;
  int nl;
  CompositeOut co{};
  if (Ni > 0) {
    if (out) {
      co.rgb = out->rgb0 ? out->rgb0 + off * 3 : nullptr;
      co.disp = out->disp0 ? out->disp0 + off : nullptr;
      co.acc = out->acc0 ? out->acc0 + off : nullptr;
      co.depth = out->depth0 ? out->depth0 + off : nullptr;
      co.unc = out->unc0 ? out->unc0 + off * Sc : nullptr;
      co.z_std = out->z_std ? out->z_std + off : nullptr;
    }
    co.weights = h->capture ? h->w_c.as<float>() : nullptr;
    nl = launch_coarse_resample(
        h->raw_c.as<float>(), h->z_c.as<float>(), h->rays_d.as<float>(), h->u_vals.as<float>(), n, Sc, Ni,
        h->cfg.white_bkgd, co, h->z_f.as<float>(), h->capture ? h->cdf.as<float>() : nullptr,
        h->capture ? h->inds.as<int32_t>() : nullptr, h->capture ? h->zs.as<float>() : nullptr, st, score_only);
    if (nl < 0) return fail(EVPP_ERR_CUDA, "resampling kernel configuration failed (shared memory for %d + %d samples)", Sc, Ni);
    h->launches += nl;
    RayBatch rf{h->rays_o.as<float>(), h->rays_d.as<float>(), h->viewdirs.as<float>(), h->z_f.as<float>(), n, Sf};
    rc = run_mlp(h, EVPP_NET_FINE, prec, rf, h->raw_f.as<float>(), st, score_only ? beta_heads : full_heads);
    if (rc) return rc;
  }
  const float* raw = Ni > 0 ? h->raw_f.as<float>() : h->raw_c.as<float>();
  const float* z = Ni > 0 ? h->z_f.as<float>() : h->z_c.as<float>();
  const int S = Ni > 0 ? Sf : Sc;
  if (score_only) {
    h->launches += launch_beta2(raw, n, S, beta2, st);
    h->last_n = n;
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  CompositeOut cf{};
  if (out) {
    cf.rgb = out->rgb ? out->rgb + off * 3 : nullptr;
    cf.disp = out->disp ? out->disp + off : nullptr;
    cf.acc = out->acc ? out->acc + off : nullptr;
    cf.depth = out->depth ? out->depth + off : nullptr;
    cf.unc = out->unc ? out->unc + off * S : nullptr;
    if (out->raw)
      CUDA_TRY(cudaMemcpyAsync(out->raw + off * S * 5, raw, (size_t)n * S * 5 * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (Ni == 0 && h->capture) cf.weights = h->w_c.as<float>();
  cf.beta2 = beta2;
  nl = launch_composite(raw, z, h->rays_d.as<float>(), n, S, h->cfg.white_bkgd, cf, st);
  if (nl < 0) return fail(EVPP_ERR_CUDA, "compositing kernel configuration failed (shared memory for %d samples)", S);
  h->launches += nl;
  h->last_n = n;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

std::vector<DevBuf*> all_buffers(evpp_handle* h) {
  std::vector<DevBuf*> v = {&h->t_vals, &h->u_vals, &h->rays_o, &h->rays_d, &h->viewdirs, &h->z_c, &h->raw_c, &h->z_f,
                            &h->raw_f, &h->beta2, &h->nonfinite, &h->st_c2w, &h->st_pix, &h->st_scores, &h->w_c, &h->cdf, &h->inds, &h->zs,
                            &h->tsdf_state, &h->tsdf_nray, &h->tsdf_t, &h->tsdf_ro, &h->tsdf_rd, &h->tsdf_u, &h->tsdf_d,
                            &h->tsdf_vu, &h->tsdf_vd, &h->ngp_table, &h->ngp_mlp, &h->ngp_bits, &h->ngp_ro, &h->ngp_rd,
                            &h->ngp_counts, &h->ngp_offsets, &h->ngp_total, &h->ngp_masks, &h->ngp_tiles, &h->ngp_t0, &h->ngp_t, &h->ngp_ray, &h->ngp_raw, &h->ngp_beta2,
                            &h->ngp_scores, &h->ngp_delta, &h->ngp_cell, &h->ngp_counter, &h->tf_planes, &h->tf_lines, &h->tf_mlp, &h->sg_ws};
  for (int i = 0; i < 2; ++i) {
    NetWeights& nw = h->net[i];
    for (DevBuf* b : {&nw.simt_stream, &nw.bias, &nw.head_w, &nw.head_b}) v.push_back(b);
    for (auto& pr : nw.tc_image) { v.push_back(&pr[0]); v.push_back(&pr[1]); }
  }
  return v;
}

int check_device_ok(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(EVPP_ERR_NO_DEVICE, "no CUDA device visible; evpp_b200 has no CPU fallback");
  }
  if (device < 0 || device >= count) return fail(EVPP_ERR_INVALID, "device ordinal %d out of range (%d devices)", device, count);
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10)
    return fail(EVPP_ERR_NO_DEVICE, "device %d has compute capability %d.x; this build is sm_100a only", device, major);
  return 0;
}

int score_views_device(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy, float cx,
                       float cy, const int32_t* pix_idx, int R, float* scores, int prec, cudaStream_t st) {
  if (V < 0 || H <= 0 || W <= 0) return fail(EVPP_ERR_INVALID, "bad V/H/W");
  if (V == 0) return 0;
  if (!pix_idx) R = H * W;
  if (R <= 0) return fail(EVPP_ERR_INVALID, "R must be positive when pix_idx is given");
  NvtxRange nvtx_score("evpp:score_views");
  const long long N = (long long)V * R;
  const int cap = (int)std::min<long long>(N, h->cfg.max_chunk_rays);
  int rc = ensure_workspace(h, cap);
  if (rc) return rc;
  if (h->beta2.ensure((size_t)N * 4) || h->nonfinite.ensure(4)) return fail(EVPP_ERR_CUDA, "beta2 buffer alloc failed");
  CUDA_TRY(cudaMemsetAsync(h->nonfinite.p, 0, 4, st));
  for (long long g0 = 0; g0 < N; g0 += cap) {
    const int n = (int)std::min<long long>(cap, N - g0);
    h->launches += launch_get_rays(c2w, V, H, W, fx, fy, cx, cy, pix_idx, R, g0, n, h->rays_o.as<float>(),
                                   h->rays_d.as<float>(), h->viewdirs.as<float>(), st);
    rc = render_chunk(h, n, h->cfg.near_plane, h->cfg.far_plane, prec, nullptr, 0, h->beta2.as<float>() + g0, st);
    if (rc) return rc;
  }
  h->launches += launch_view_reduce(h->beta2.as<float>(), V, R, scores, h->nonfinite.as<int>(), st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

const char* evpp_last_error(void) { return g_err.c_str(); }
int evpp_abi_version(void) { return EVPP_ABI_VERSION; }

int evpp_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < count; ++d) {
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d);
    if (major == 10) ++ok;
  }
  return ok;
}

int evpp_create(int device_ordinal, const evpp_config* cfg, evpp_handle** out) {
  if (!cfg || !out) return fail(EVPP_ERR_INVALID, "null cfg/out");
  *out = nullptr;
  int rc = check_device_ok(device_ordinal);
  if (rc) return rc;
  if (cfg->n_samples < 4 || cfg->n_samples > kMaxSc)
    return fail(EVPP_ERR_INVALID, "n_samples %d out of range [4,%d]", cfg->n_samples, kMaxSc);
  if (cfg->n_importance < 0 || cfg->n_importance + cfg->n_samples > kMaxSf)
    return fail(EVPP_ERR_INVALID, "n_samples+n_importance %d exceeds %d", cfg->n_importance + cfg->n_samples, kMaxSf);
  if (cfg->precision < 0 || cfg->precision >= kNumPrec) return fail(EVPP_ERR_INVALID, "bad precision %d", cfg->precision);
  EVPP_ON_DEVICE(device_ordinal);
  evpp_handle* h = new evpp_handle();
  h->device = device_ordinal;
  h->cfg = *cfg;
  h->ngp_chunk_rays = h->cfg.max_chunk_rays > 0 ? h->cfg.max_chunk_rays : (1 << 20);   // few samples per ray: larger chunks
  if (h->cfg.max_chunk_rays <= 0) h->cfg.max_chunk_rays = 131072;
  h->Sc = cfg->n_samples;
  h->Ni = cfg->n_importance;
  h->Sf = h->Sc + h->Ni;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device_ordinal);
  auto t = torch_linspace(0.f, 1.f, h->Sc);
  auto u = torch_linspace(0.f, 1.f, h->Ni > 0 ? h->Ni : 1);
  if (h->t_vals.ensure(t.size() * 4) || h->u_vals.ensure(u.size() * 4)) {
    delete h;
    return fail(EVPP_ERR_CUDA, "table alloc failed");
  }
  cudaMemcpy(h->t_vals.p, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(h->u_vals.p, u.data(), u.size() * 4, cudaMemcpyHostToDevice);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    delete h;
    return fail(EVPP_ERR_CUDA, "create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return 0;
}

int evpp_destroy(evpp_handle* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  for (auto& pr : h->prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (DevBuf* b : all_buffers(h)) b->release();
  delete h;
  return 0;
}

int evpp_load_nerf_weights(evpp_handle* h, int which, const float* packed, const int64_t* offsets, int n) {
  if (!h || !packed || !offsets) return fail(EVPP_ERR_INVALID, "null argument");
  if (which != 0 && which != 1) return fail(EVPP_ERR_INVALID, "which_net must be 0 (coarse) or 1 (fine)");
  if (n != 26) return fail(EVPP_ERR_INVALID, "expected 26 state_dict tensors (NeRF D=8,W=256,use_viewdirs), got %d", n);
  EVPP_ON_DEVICE(h->device);
  NetWeights& nw = h->net[which];
  nw.host.assign(26, {});
  for (int i = 0; i < 26; ++i) {
    if (offsets[i] < 0) return fail(EVPP_ERR_INVALID, "negative offset for tensor %d", i);
    nw.host[i].resize(kTensorSizes[i]);
    CUDA_TRY(cudaMemcpy(nw.host[i].data(), packed + offsets[i], (size_t)kTensorSizes[i] * 4, cudaMemcpyDefault));
    for (float v : nw.host[i])
      if (!std::isfinite(v)) return fail(EVPP_ERR_INVALID, "non-finite weight in tensor %d", i);
  }
  // ---- fp32 stream: per layer [K_padded][N] (input-major), chunked in 2048 floats -------------
  std::vector<float> stream((size_t)kSimtChunksTotal * kSimtChunkFloats, 0.f);
  size_t pos = 0;
  auto put = [&](const std::vector<float>& Wt, int N, int in_dim, int in_begin, int in_count, int k_pad) {
    // appends rows k=0..k_pad-1: row k = W[:, in_begin+k] (zero if k >= in_count)
    for (int k = 0; k < k_pad; ++k) {
      if (k < in_count)
        for (int o = 0; o < N; ++o) stream[pos + (size_t)k * N + o] = Wt[(size_t)o * in_dim + in_begin + k];
    }
    pos += (size_t)k_pad * N;
  };
  put(nw.host[0], 256, 63, 0, 63, 64);
  for (int l = 1; l <= 4; ++l) put(nw.host[2 * l], 256, 256, 0, 256, 256);
  put(nw.host[10], 256, 319, 0, 63, 64);      // skip layer: cat[input_pts(63), h(256)]
  put(nw.host[10], 256, 319, 63, 256, 256);
  put(nw.host[12], 256, 256, 0, 256, 256);
  put(nw.host[14], 256, 256, 0, 256, 256);
  put(nw.host[18], 256, 256, 0, 256, 256);    // feature_linear
  put(nw.host[16], 128, 283, 0, 256, 256);    // views_linears.0: cat[feature(256), dirs(27)]
  put(nw.host[16], 128, 283, 256, 27, 32);
  if (pos != stream.size()) return fail(EVPP_ERR_INVALID, "internal: weight stream size mismatch");
  std::vector<float> bias(10 * 256, 0.f);
  for (int l = 0; l < 8; ++l) memcpy(&bias[l * 256], nw.host[2 * l + 1].data(), 256 * 4);
  memcpy(&bias[8 * 256], nw.host[19].data(), 256 * 4);
  memcpy(&bias[9 * 256], nw.host[17].data(), 128 * 4);
  std::vector<float> head_w(256 + 384 + 128), head_b(8, 0.f);
  memcpy(&head_w[0], nw.host[20].data(), 256 * 4);
  memcpy(&head_w[256], nw.host[22].data(), 384 * 4);
  memcpy(&head_w[640], nw.host[24].data(), 128 * 4);
  head_b[0] = nw.host[21][0];
  head_b[1] = nw.host[23][0]; head_b[2] = nw.host[23][1]; head_b[3] = nw.host[23][2];
  head_b[4] = nw.host[25][0];
  if (nw.simt_stream.ensure(stream.size() * 4) || nw.bias.ensure(bias.size() * 4) ||
      nw.head_w.ensure(head_w.size() * 4) || nw.head_b.ensure(head_b.size() * 4))
    return fail(EVPP_ERR_CUDA, "weight buffer alloc failed");
  CUDA_TRY(cudaMemcpy(nw.simt_stream.p, stream.data(), stream.size() * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(nw.bias.p, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(nw.head_w.p, head_w.data(), head_w.size() * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(nw.head_b.p, head_b.data(), head_b.size() * 4, cudaMemcpyHostToDevice));
  for (auto& pr : nw.tc_ready) pr[0] = pr[1] = false;
  nw.loaded = true;
  if (h->cfg.precision != EVPP_PREC_FP32) {
    int rc = ensure_tc(h, which, h->cfg.precision);
    if (rc) return rc;
  }
  return 0;
}

int evpp_score_views_ex(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy, float cx,
                        float cy, const int32_t* pix_idx, int R, int precision, float* scores, void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (V == 0) return 0;   // empty candidate list: nothing to score (empty tensors carry null pointers)
  if (!c2w || !scores) return fail(EVPP_ERR_INVALID, "null argument");
  if (precision < 0) precision = h->cfg.precision;
  if (precision >= kNumPrec) return fail(EVPP_ERR_INVALID, "bad precision %d", precision);
  EVPP_ON_DEVICE(h->device);
  return score_views_device(h, c2w, V, H, W, fx, fy, cx, cy, pix_idx, R, scores, precision, (cudaStream_t)stream);
}

int evpp_score_views(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy, float cx,
                     float cy, const int32_t* pix_idx, int R, float* scores, void* stream) {
  return evpp_score_views_ex(h, c2w, V, H, W, fx, fy, cx, cy, pix_idx, R, -1, scores, stream);
}

int evpp_get_nonfinite(evpp_handle* h, int32_t* count, void* stream) {
  if (!h || !count) return fail(EVPP_ERR_INVALID, "null argument");
  *count = 0;
  if (!h->nonfinite.p) return 0;
  EVPP_ON_DEVICE(h->device);
  CUDA_TRY(cudaMemcpyAsync(count, h->nonfinite.p, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

static int score_host_impl(evpp_handle* h, const float* c2w_host, int V, int H, int W, float fx, float fy,
                           float cx, float cy, const int32_t* pix_host, int R, float* scores_host, int prec,
                           cudaStream_t st) {
  if (V == 0) return 0;
  if (h->st_c2w.ensure((size_t)V * 12 * 4) || h->st_scores.ensure((size_t)V * 4))
    return fail(EVPP_ERR_CUDA, "staging alloc failed");
  CUDA_TRY(cudaMemcpyAsync(h->st_c2w.p, c2w_host, (size_t)V * 12 * 4, cudaMemcpyHostToDevice, st));
  const int32_t* pix_dev = nullptr;
  if (pix_host) {
    if (h->st_pix.ensure((size_t)V * R * 4)) return fail(EVPP_ERR_CUDA, "staging alloc failed");
    CUDA_TRY(cudaMemcpyAsync(h->st_pix.p, pix_host, (size_t)V * R * 4, cudaMemcpyHostToDevice, st));
    pix_dev = h->st_pix.as<int32_t>();
  }
  int rc = score_views_device(h, h->st_c2w.as<float>(), V, H, W, fx, fy, cx, cy, pix_dev, R,
                              h->st_scores.as<float>(), prec, st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(scores_host, h->st_scores.p, (size_t)V * 4, cudaMemcpyDeviceToHost, st));
  int32_t bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&bad, h->nonfinite.p, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (bad)
    return fail(EVPP_ERR_NONFINITE, "%d of %d view scores are not finite (precision %d: 16-bit activations overflowed; retry with "
                "EVPP_PREC_BF16 or EVPP_PREC_FP32)", (int)bad, V, prec);
  return 0;
}

int evpp_score_views_host(evpp_handle* h, const float* c2w_host, int V, int H, int W, float fx, float fy,
                          float cx, float cy, const int32_t* pix_idx_host, int R, float* scores_host,
                          void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (V == 0) return 0;
  if (!c2w_host || !scores_host) return fail(EVPP_ERR_INVALID, "null argument");
  EVPP_ON_DEVICE(h->device);
  return score_host_impl(h, c2w_host, V, H, W, fx, fy, cx, cy, pix_idx_host, R, scores_host, h->cfg.precision,
                         (cudaStream_t)stream);
}

int evpp_rescore_views(evpp_handle* h, const float* c2w_host, const int32_t* view_ids, int K, int H, int W,
                       float fx, float fy, float cx, float cy, const int32_t* pix_idx_host, int R, int precision,
                       float* scores_host, void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (precision < 0 || precision >= kNumPrec) return fail(EVPP_ERR_INVALID, "bad precision %d", precision);
  if (K <= 0) return 0;      // an empty re-score list (empty tensors carry null pointers)
  if (!c2w_host || !view_ids || !scores_host) return fail(EVPP_ERR_INVALID, "null argument");
  EVPP_ON_DEVICE(h->device);
  std::vector<float> poses((size_t)K * 12);
  std::vector<int32_t> pix;
  if (pix_idx_host) pix.resize((size_t)K * R);
  for (int k = 0; k < K; ++k) {
    if (view_ids[k] < 0) return fail(EVPP_ERR_INVALID, "negative view id");
    memcpy(&poses[(size_t)k * 12], c2w_host + (size_t)view_ids[k] * 12, 48);
    if (pix_idx_host) memcpy(&pix[(size_t)k * R], pix_idx_host + (size_t)view_ids[k] * R, (size_t)R * 4);
  }
  return score_host_impl(h, poses.data(), K, H, W, fx, fy, cx, cy, pix_idx_host ? pix.data() : nullptr, R,
                         scores_host, precision, (cudaStream_t)stream);
}

int evpp_rescore_views_fp32(evpp_handle* h, const float* c2w_host, const int32_t* view_ids, int K, int H, int W,
                            float fx, float fy, float cx, float cy, const int32_t* pix_idx_host, int R,
                            float* scores_host, void* stream) {
  return evpp_rescore_views(h, c2w_host, view_ids, K, H, W, fx, fy, cx, cy, pix_idx_host, R, EVPP_PREC_FP32,
                            scores_host, stream);
}

int evpp_set_render_mode(evpp_handle* h, int fold_feature_layer) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  h->render_fold = fold_feature_layer != 0;
  return 0;
}

int evpp_set_score_mode(evpp_handle* h, int mode) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (mode < 0 || mode > 2) return fail(EVPP_ERR_INVALID, "score mode must be 0 (lean, folded), 1 (full outputs) or 2 (lean, reference layer order)");
  h->score_mode = mode;
  return 0;
}

int evpp_render_rays_ex(evpp_handle* h, const float* rays_o, const float* rays_d, const float* viewdirs, int64_t N,
                        float near_plane, float far_plane, int precision, const evpp_render_out* out, void* stream) {
  if (!h || !out) return fail(EVPP_ERR_INVALID, "null argument");
  if (N < 0) return fail(EVPP_ERR_INVALID, "negative N");
  if (N == 0) return 0;
  if (!rays_o || !rays_d) return fail(EVPP_ERR_INVALID, "null rays");
  if (precision < 0) precision = h->cfg.precision;
  if (precision >= kNumPrec) return fail(EVPP_ERR_INVALID, "bad precision");
  EVPP_ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int cap = (int)std::min<int64_t>(N, h->cfg.max_chunk_rays);
  int rc = ensure_workspace(h, cap);
  if (rc) return rc;
  for (int64_t g0 = 0; g0 < N; g0 += cap) {
    const int n = (int)std::min<int64_t>(cap, N - g0);
    CUDA_TRY(cudaMemcpyAsync(h->rays_o.p, rays_o + g0 * 3, (size_t)n * 12, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->rays_d.p, rays_d + g0 * 3, (size_t)n * 12, cudaMemcpyDeviceToDevice, st));
    if (viewdirs) {
      // caller-provided view directions, used as they are (Controller.get_rays_depth hands batchify_rays the raw,
      // un-normalised directions: nerf.py:364-366)
      CUDA_TRY(cudaMemcpyAsync(h->viewdirs.p, viewdirs + g0 * 3, (size_t)n * 12, cudaMemcpyDeviceToDevice, st));
    } else {
      // viewdirs = rays_d / ||rays_d|| (run_nerf.py:165-171): reuse the ray kernel's normalisation
      h->launches += launch_normalize_dirs(h->rays_d.as<float>(), n, h->viewdirs.as<float>(), st);
    }
    rc = render_chunk(h, n, near_plane, far_plane, precision, out, g0, nullptr, st);
    if (rc) return rc;
  }
  return 0;
}

int evpp_render_rays(evpp_handle* h, const float* rays_o, const float* rays_d, int64_t N, float near_plane,
                     float far_plane, int precision, const evpp_render_out* out, void* stream) {
  return evpp_render_rays_ex(h, rays_o, rays_d, nullptr, N, near_plane, far_plane, precision, out, stream);
}

int evpp_get_rays(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy, float cx, float cy,
                  const int32_t* pix_idx, int R, float* rays_o, float* rays_d, float* viewdirs, void* stream) {
  if (!h || !c2w) return fail(EVPP_ERR_INVALID, "null argument");
  if (!pix_idx) R = H * W;
  const long long N = (long long)V * R;
  if (N <= 0) return 0;
  if (N > 0x7fffffffLL) return fail(EVPP_ERR_INVALID, "too many rays for one call");
  EVPP_ON_DEVICE(h->device);
  h->launches += launch_get_rays(c2w, V, H, W, fx, fy, cx, cy, pix_idx, R, 0, (int)N, rays_o, rays_d, viewdirs,
                                 (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int mlp_forward_impl(evpp_handle* h, int which, int precision, const float* pts, const float* dirs,
                            int64_t M, float* out, int dbg_layer, float* dbg, void* stream,
                            long long* trace = nullptr) {
  if (!h || !pts || !dirs || !out) return fail(EVPP_ERR_INVALID, "null argument");
  if (which != 0 && which != 1) return fail(EVPP_ERR_INVALID, "bad net");
  if (M <= 0) return 0;
  EVPP_ON_DEVICE(h->device);
  NetWeights& nw = h->net[which];
  if (!nw.loaded) return fail(EVPP_ERR_STATE, "weights of net %d not loaded", which);
  if (precision < 0) precision = h->cfg.precision;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == EVPP_PREC_FP32) {
    if (dbg) return fail(EVPP_ERR_INVALID, "layer dumps exist for the tensor-core path only");
    h->launches += launch_mlp_simt_points(simt_of(nw), pts, dirs, M, out, st);
  } else if (precision < kNumPrec) {
    int rc = ensure_tc(h, which, precision);
    if (rc) return rc;
    int nl = launch_mlp_tc_points(tc_of(nw, precision, dbg, dbg_layer, trace), pts, dirs, M, out, h->num_sms, st);
    if (nl < 0) return fail(EVPP_ERR_CUDA, "tcgen05 MLP launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->launches += nl;
  } else {
    return fail(EVPP_ERR_INVALID, "bad precision");
  }
  h->evals += M;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int evpp_mlp_forward(evpp_handle* h, int which, int precision, const float* pts, const float* dirs, int64_t M,
                     float* out, void* stream) {
  return mlp_forward_impl(h, which, precision, pts, dirs, M, out, -1, nullptr, stream);
}

int evpp_mlp_layer_dump(evpp_handle* h, int which, int precision, const float* pts, const float* dirs, int64_t M,
                        int layer, float* out, float* acts, void* stream) {
  if (layer < 0 || layer > 9 || !acts) return fail(EVPP_ERR_INVALID, "layer must be 0..9 and acts non-null");
  return mlp_forward_impl(h, which, precision, pts, dirs, M, out, layer, acts, stream);
}

int evpp_mlp_trace(evpp_handle* h, int which, int precision, const float* pts, const float* dirs, int64_t M,
                   float* out, int64_t* trace, void* stream) {
  if (!trace) return fail(EVPP_ERR_INVALID, "null trace buffer");
  if (precision == EVPP_PREC_FP32) return fail(EVPP_ERR_INVALID, "the pipeline trace exists for the tensor-core path only");
  return mlp_forward_impl(h, which, precision, pts, dirs, M, out, -1, nullptr, stream,
                          reinterpret_cast<long long*>(trace));
}

int evpp_set_stage_capture(evpp_handle* h, int enabled) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  h->capture = enabled != 0;
  h->chunk_cap = 0;   // force (re)allocation of the capture buffers
  return 0;
}

int evpp_stage_dump(evpp_handle* h, int stage, void* out, int64_t* count, void* stream) {
  if (!h || !count) return fail(EVPP_ERR_INVALID, "null argument");
  const int64_t n = h->last_n;
  const void* src = nullptr;
  int64_t cnt = 0;
  switch (stage) {
    case EVPP_STAGE_Z_COARSE: src = h->z_c.p; cnt = n * h->Sc; break;
    case EVPP_STAGE_RAW_COARSE: src = h->raw_c.p; cnt = n * h->Sc * 5; break;
    case EVPP_STAGE_W_COARSE: src = h->w_c.p; cnt = n * h->Sc; break;
    case EVPP_STAGE_CDF: src = h->cdf.p; cnt = n * (h->Sc - 1); break;
    case EVPP_STAGE_INDS: src = h->inds.p; cnt = n * h->Ni; break;
    case EVPP_STAGE_Z_SAMPLES: src = h->zs.p; cnt = n * h->Ni; break;
    case EVPP_STAGE_Z_FINE: src = h->z_f.p; cnt = n * h->Sf; break;
    default: return fail(EVPP_ERR_INVALID, "unknown stage %d", stage);
  }
  *count = cnt;
  if (!out) return 0;
  if (!src) return fail(EVPP_ERR_STATE, "stage %d not captured (evpp_set_stage_capture before rendering)", stage);
  EVPP_ON_DEVICE(h->device);
  CUDA_TRY(cudaMemcpyAsync(out, src, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int evpp_composite(evpp_handle* h, const float* raw, const float* z, const float* rays_d, int n, int S,
                   float* rgb, float* disp, float* acc, float* depth, float* weights, float* beta2, void* stream) {
  if (!h || !raw || !z || !rays_d) return fail(EVPP_ERR_INVALID, "null argument");
  if (n <= 0) return 0;
  if (S < 1 || S > kMaxSf) return fail(EVPP_ERR_INVALID, "S (%d) must be in [1,%d]", S, kMaxSf);
  EVPP_ON_DEVICE(h->device);
  CompositeOut co{};
  co.rgb = rgb; co.disp = disp; co.acc = acc; co.depth = depth; co.weights = weights; co.beta2 = beta2;
  const int nl = launch_composite(raw, z, rays_d, n, S, h->cfg.white_bkgd, co, (cudaStream_t)stream);
  if (nl < 0) return fail(EVPP_ERR_CUDA, "compositing kernel configuration failed (shared memory for %d samples)", S);
  h->launches += nl;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int evpp_resample(evpp_handle* h, const float* raw_coarse, const float* z_coarse, const float* rays_d, int n,
                  float* z_fine, float* cdf, int32_t* inds, float* z_samples, void* stream) {
  if (!h || !raw_coarse || !z_coarse || !rays_d || !z_fine) return fail(EVPP_ERR_INVALID, "null argument");
  if (h->Ni <= 0) return fail(EVPP_ERR_STATE, "engine configured with n_importance == 0");
  if (n <= 0) return 0;
  EVPP_ON_DEVICE(h->device);
  CompositeOut co{};
  const int nl = launch_coarse_resample(raw_coarse, z_coarse, rays_d, h->u_vals.as<float>(), n, h->Sc, h->Ni,
                                        h->cfg.white_bkgd, co, z_fine, cdf, inds, z_samples, (cudaStream_t)stream);
  if (nl < 0) return fail(EVPP_ERR_CUDA, "resampling kernel configuration failed");
  h->launches += nl;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int evpp_select_nbv(evpp_handle* h, const float* scores, const float* weight, int V, int dirs_per_location,
                    float* norm_scores, int32_t* best_dir, int32_t* winner_view, void* stream) {
  if (!h || !scores) return fail(EVPP_ERR_INVALID, "null argument");
  if (dirs_per_location <= 0 || V <= 0 || V % dirs_per_location)
    return fail(EVPP_ERR_INVALID, "V (%d) must be a positive multiple of dirs_per_location (%d)", V, dirs_per_location);
  EVPP_ON_DEVICE(h->device);
  h->launches += launch_select_nbv(scores, weight, V, dirs_per_location, norm_scores, best_dir, winner_view,
                                   (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// planner surrogate g_phi (plannerServer_*/core/vpp.py:26-67, 619-654)
// ---------------------------------------------------------------------------------------------------
int evpp_sample_pixels(evpp_handle* h, uint64_t seed, int V, int R, int n_pixels, int32_t* pix_idx, void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (V < 0 || R < 0 || n_pixels <= 0 || R > n_pixels) return fail(EVPP_ERR_INVALID, "need 0 <= R <= n_pixels");
  if (V == 0 || R == 0) return 0;
  if (!pix_idx) return fail(EVPP_ERR_INVALID, "null pix_idx");
  EVPP_ON_DEVICE(h->device);
  h->launches += launch_pixel_subset(seed, V, R, n_pixels, pix_idx, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int evpp_surrogate_param_count(void) { return kSurrogateParams; }

int evpp_surrogate_fit(evpp_handle* h, float* params, const float* xyz, const float* gain, int N, int epochs, float lr,
                       float* loss_curve, void* stream) {
  if (!h || !params || !xyz || !gain) return fail(EVPP_ERR_INVALID, "null argument");
  if (N < 1 || N > kSurrogateMaxN) return fail(EVPP_ERR_INVALID, "N (%d) must be in [1,%d]", N, kSurrogateMaxN);
  if (epochs < 0 || !(lr > 0.f)) return fail(EVPP_ERR_INVALID, "bad epochs / lr");
  if (epochs == 0) return 0;
  EVPP_ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t moments = (2 * (size_t)kSurrogateParams + 3) & ~(size_t)3;   // keeps the activation workspace 16-byte aligned
  if (h->sg_ws.ensure((moments + surrogate_workspace_floats(N)) * 4)) return fail(EVPP_ERR_CUDA, "surrogate workspace alloc failed");
  float* m = h->sg_ws.as<float>();
  CUDA_TRY(cudaMemsetAsync(m, 0, moments * 4, st));     // a fresh optimizer per fit, like train_only
  const int nl = launch_surrogate_fit(params, m, m + kSurrogateParams, xyz, gain, N, epochs, lr, m + moments, loss_curve, st);
  if (nl < 0) return fail(EVPP_ERR_CUDA, "surrogate kernel configuration failed");
  h->launches += nl;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int evpp_surrogate_query(evpp_handle* h, const float* params, const float* xyz, int64_t M, float* out, void* stream) {
  if (!h || !params) return fail(EVPP_ERR_INVALID, "null argument");
  if (M < 0 || M > (1ll << 30)) return fail(EVPP_ERR_INVALID, "bad M");
  if (M == 0) return 0;
  if (!xyz || !out) return fail(EVPP_ERR_INVALID, "null argument");
  EVPP_ON_DEVICE(h->device);
  const int nl = launch_surrogate_query(params, xyz, (int)M, out, (cudaStream_t)stream);
  if (nl < 0) return fail(EVPP_ERR_CUDA, "surrogate kernel configuration failed");
  h->launches += nl;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// stage-1 TSDF occupancy ray-cast scorer (plannerServer_*/core/tsdf_gpu.py)
// ---------------------------------------------------------------------------------------------------
int evpp_tsdf_set_volume(evpp_handle* h, const float* state, const float* nray, const int32_t* dims,
                         const float* origin, float voxel_size, const float* aabb_min, const float* aabb_max,
                         float near_plane, float far_plane, int32_t n_samples, float c_unknown, float c_surface) {
  if (!h || !state || !nray || !dims || !origin || !aabb_min || !aabb_max) return fail(EVPP_ERR_INVALID, "null argument");
  if (dims[0] < 2 || dims[1] < 2 || dims[2] < 2) return fail(EVPP_ERR_INVALID, "volume dims must be >= 2");
  if (!(voxel_size > 0.f) || !(far_plane > near_plane)) return fail(EVPP_ERR_INVALID, "bad voxel size / bounds");
  EVPP_ON_DEVICE(h->device);
  const long long n = (long long)dims[0] * dims[1] * dims[2];
  const int S = n_samples;   // the caller evaluates int((far-near)/voxel_res)+1 in double like tsdf_gpu.py:41
  if (S < 2 || S > 4096) return fail(EVPP_ERR_INVALID, "sample count %d out of range", S);
  DevBuf tmp;
  if (tmp.ensure((size_t)n * 4) || h->tsdf_state.ensure((size_t)n) || h->tsdf_nray.ensure((size_t)n * 4) ||
      h->tsdf_t.ensure((size_t)S * 4))
    return fail(EVPP_ERR_CUDA, "TSDF volume allocation failed");
  CUDA_TRY(cudaMemcpy(tmp.p, state, (size_t)n * 4, cudaMemcpyDefault));
  CUDA_TRY(cudaMemcpy(h->tsdf_nray.p, nray, (size_t)n * 4, cudaMemcpyDefault));
  h->launches += launch_tsdf_pack_state(tmp.as<float>(), n, h->tsdf_state.as<uint8_t>(), nullptr);
  CUDA_TRY(cudaDeviceSynchronize());
  tmp.release();
  auto t = torch_linspace(0.f, 1.f, S);
  CUDA_TRY(cudaMemcpy(h->tsdf_t.p, t.data(), (size_t)S * 4, cudaMemcpyHostToDevice));
  TsdfVolume& v = h->tsdf;
  v.state = h->tsdf_state.as<uint8_t>();
  v.nray = h->tsdf_nray.as<float>();
  v.t_vals = h->tsdf_t.as<float>();
  for (int c = 0; c < 3; ++c) { v.dim[c] = dims[c]; v.origin[c] = origin[c]; v.amin[c] = aabb_min[c]; v.amax[c] = aabb_max[c]; }
  v.voxel = voxel_size;
  v.near_plane = near_plane;
  v.far_plane = far_plane;
  v.n_samples = S;
  v.c_unknown = c_unknown;
  v.c_surface = c_surface;
  h->tsdf_ready = true;
  return 0;
}

int evpp_tsdf_render_rays(evpp_handle* h, const float* rays_o, const float* rays_d, int64_t N, float* uncer,
                          float* depth, int32_t* n_unknown, int32_t* first_hit, int32_t* voxel_idx, void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (N == 0) return 0;
  if (N < 0 || N > 0x7fffffffLL) return fail(EVPP_ERR_INVALID, "bad ray count");
  if (!rays_o || !rays_d) return fail(EVPP_ERR_INVALID, "null rays");
  if (!h->tsdf_ready) return fail(EVPP_ERR_STATE, "no TSDF state volume (evpp_tsdf_set_volume)");
  EVPP_ON_DEVICE(h->device);
  h->launches += launch_tsdf_march(h->tsdf, rays_o, rays_d, (int)N, uncer, depth, n_unknown, first_hit, voxel_idx,
                                   (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int evpp_tsdf_score_views(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy, float cx,
                          float cy, const int32_t* pix_idx, int R, float* view_uncer, float* view_depth,
                          void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (V == 0) return 0;
  if (!c2w || !view_uncer || !view_depth) return fail(EVPP_ERR_INVALID, "null argument");
  if (V < 0 || H <= 0 || W <= 0) return fail(EVPP_ERR_INVALID, "bad V/H/W");
  if (!h->tsdf_ready) return fail(EVPP_ERR_STATE, "no TSDF state volume (evpp_tsdf_set_volume)");
  if (!pix_idx) R = H * W;
  if (R <= 0) return fail(EVPP_ERR_INVALID, "R must be positive when pix_idx is given");
  const long long N = (long long)V * R;
  if (N > 0x7fffffffLL) return fail(EVPP_ERR_INVALID, "too many rays for one call");
  EVPP_ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->tsdf_ro.ensure((size_t)N * 12) || h->tsdf_rd.ensure((size_t)N * 12) || h->tsdf_u.ensure((size_t)N * 4) ||
      h->tsdf_d.ensure((size_t)N * 4))
    return fail(EVPP_ERR_CUDA, "TSDF ray workspace allocation failed");
  h->launches += launch_get_rays(c2w, V, H, W, fx, fy, cx, cy, pix_idx, R, 0, (int)N, h->tsdf_ro.as<float>(),
                                 h->tsdf_rd.as<float>(), nullptr, st);
  h->launches += launch_tsdf_march(h->tsdf, h->tsdf_ro.as<float>(), h->tsdf_rd.as<float>(), (int)N, h->tsdf_u.as<float>(),
                                   h->tsdf_d.as<float>(), nullptr, nullptr, nullptr, st);
  h->launches += launch_tsdf_view_reduce(h->tsdf_u.as<float>(), h->tsdf_d.as<float>(), V, R, h->tsdf.near_plane,
                                         h->tsdf.far_plane, view_uncer, view_depth, st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int evpp_tsdf_score_views_host(evpp_handle* h, const float* c2w_host, int V, int H, int W, float fx, float fy,
                               float cx, float cy, const int32_t* pix_idx_host, int R, float* view_uncer_host,
                               float* view_depth_host, void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (V == 0) return 0;
  if (!c2w_host || !view_uncer_host || !view_depth_host) return fail(EVPP_ERR_INVALID, "null argument");
  EVPP_ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->st_c2w.ensure((size_t)V * 48) || h->tsdf_vu.ensure((size_t)V * 4) || h->tsdf_vd.ensure((size_t)V * 4))
    return fail(EVPP_ERR_CUDA, "staging alloc failed");
  CUDA_TRY(cudaMemcpyAsync(h->st_c2w.p, c2w_host, (size_t)V * 48, cudaMemcpyHostToDevice, st));
  const int32_t* pix_dev = nullptr;
  if (pix_idx_host) {
    if (h->st_pix.ensure((size_t)V * R * 4)) return fail(EVPP_ERR_CUDA, "staging alloc failed");
    CUDA_TRY(cudaMemcpyAsync(h->st_pix.p, pix_idx_host, (size_t)V * R * 4, cudaMemcpyHostToDevice, st));
    pix_dev = h->st_pix.as<int32_t>();
  }
  int rc = evpp_tsdf_score_views(h, h->st_c2w.as<float>(), V, H, W, fx, fy, cx, cy, pix_dev, R, h->tsdf_vu.as<float>(),
                                 h->tsdf_vd.as<float>(), stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(view_uncer_host, h->tsdf_vu.p, (size_t)V * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(view_depth_host, h->tsdf_vd.p, (size_t)V * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// NGP-mode scorer (north_star subsystems 1, 2, 4; no counterpart in the reference)
// ---------------------------------------------------------------------------------------------------
int evpp_ngp_configure(evpp_handle* h, const evpp_ngp_config* cfg, const float* level_scale, const int32_t* level_res,
                       const int64_t* level_offset, const int64_t* level_size) {
  if (!h || !cfg || !level_scale || !level_res || !level_offset || !level_size) return fail(EVPP_ERR_INVALID, "null argument");
  if (cfg->n_levels < 1 || cfg->n_levels > kNgpMaxLevels) return fail(EVPP_ERR_INVALID, "n_levels must be 1..%d", kNgpMaxLevels);
  if (cfg->n_levels != 16) return fail(EVPP_ERR_INVALID, "the fused field kernel is built for 16 levels x 2 features (32 inputs)");
  if (cfg->grid_res < 8 || cfg->grid_res > 1024 || (cfg->grid_res & 7)) return fail(EVPP_ERR_INVALID, "grid_res must be a multiple of 8 in [8,1024]");
  if (cfg->march_mode != 0 && cfg->march_mode != 1) return fail(EVPP_ERR_INVALID, "march_mode must be 0 (uniform) or 1 (torch-ngp)");
  if (cfg->max_steps < 1 || cfg->max_samples < 1) return fail(EVPP_ERR_INVALID, "bad marching parameters");
  if (cfg->march_mode == 0 && (!(cfg->dt > 0.f) || !(cfg->far_plane > cfg->near_plane)))
    return fail(EVPP_ERR_INVALID, "bad marching parameters");
  if (cfg->march_mode == 1) {
    if (cfg->grid_res & (cfg->grid_res - 1)) return fail(EVPP_ERR_INVALID, "torch-ngp marching needs a power-of-two grid_res (Morton order)");
    if (cfg->cascades < 1 || cfg->cascades > 8) return fail(EVPP_ERR_INVALID, "cascades must be in [1,8]");
    if (!(cfg->dt_gamma >= 0.f) || !(cfg->min_near >= 0.f)) return fail(EVPP_ERR_INVALID, "dt_gamma and min_near must be >= 0");
    const float sx = cfg->aabb_max[0] - cfg->aabb_min[0];
    for (int c = 1; c < 3; ++c)
      if (std::fabs((cfg->aabb_max[c] - cfg->aabb_min[c]) - sx) > 1e-5f * sx)
        return fail(EVPP_ERR_INVALID, "torch-ngp marching needs a cubic scene box");
  }
  NgpParams& P = h->ngp;
  P.n_levels = cfg->n_levels;
  for (int l = 0; l < cfg->n_levels; ++l) {
    if (level_size[l] <= 0 || level_size[l] > 0x7fffffffLL || level_offset[l] < 0 || level_offset[l] + level_size[l] > 0xffffffffLL)
      return fail(EVPP_ERR_INVALID, "bad level table entry %d", l);
    // torch-ngp's index rule: add dimensions while the running stride fits the table, hash once it does not
    const uint64_t r1 = (uint64_t)level_res[l] + 1, size = (uint64_t)level_size[l];
    uint64_t stride = 1;
    int d = 0;
    for (; d < 3 && stride <= size; ++d) stride *= r1;
    const bool hashed = stride > size;
    if (hashed && (size & (size - 1))) return fail(EVPP_ERR_INVALID, "hashed level %d needs a power-of-two table size", l);
    if (!hashed && d != 3) return fail(EVPP_ERR_INVALID, "internal: dense level %d", l);
    P.levels[l] = NgpLevel{level_scale[l], (uint32_t)level_res[l], (uint32_t)level_offset[l], (uint32_t)level_size[l],
                           hashed ? 1u : 0u, (uint32_t)r1, (uint32_t)(r1 * r1)};
  }
  P.grid_res = cfg->grid_res;
  for (int c = 0; c < 3; ++c) {
    if (!(cfg->aabb_max[c] > cfg->aabb_min[c])) return fail(EVPP_ERR_INVALID, "empty scene box");
    P.amin[c] = cfg->aabb_min[c];
    P.amax[c] = cfg->aabb_max[c];
    P.inv_size[c] = 1.f / (cfg->aabb_max[c] - cfg->aabb_min[c]);
  }
  P.near_plane = cfg->near_plane; P.far_plane = cfg->far_plane; P.dt = cfg->dt;
  P.max_steps = cfg->max_steps; P.max_samples = cfg->max_samples; P.white_bkgd = cfg->white_bkgd;
  P.march_mode = cfg->march_mode;
  P.cascades = cfg->march_mode == 1 ? cfg->cascades : 1;
  if (cfg->march_mode == 1) {
    // fp32 constants exactly as the restatement computes them (the test suite holds the same arithmetic in numpy)
    P.bound = (float)(1 << (cfg->cascades - 1));
    const float half = 0.5f * (cfg->aabb_max[0] - cfg->aabb_min[0]);
    P.scale = P.bound / half;
    for (int c = 0; c < 3; ++c) P.center[c] = 0.5f * (cfg->aabb_min[c] + cfg->aabb_max[c]);
    const float two_sqrt3 = 2.f * 1.7320508075688772f;
    P.dt_gamma = cfg->dt_gamma;
    P.dt_min = two_sqrt3 / (float)cfg->max_steps;
    P.dt_max = two_sqrt3 * (float)(1 << (cfg->cascades - 1)) / (float)cfg->grid_res;
    P.min_near = cfg->min_near;
  }
  {
    DeviceGuard guard(h->device);      // (allocations must land on the handle's device, whatever the caller's current one is)
    if (guard.err != cudaSuccess || h->ngp_counter.ensure(8)) return fail(EVPP_ERR_CUDA, "counter alloc failed");
    cudaMemset(h->ngp_counter.p, 0, 8);
  }
  h->ngp_samples = 0;
  h->ngp_configured = true;
  h->ngp_has_occ = h->ngp_has_params = false;
  return 0;
}

int evpp_ngp_set_occupancy(evpp_handle* h, const uint8_t* bitfield, int64_t nbytes) {
  if (!h || !bitfield) return fail(EVPP_ERR_INVALID, "null argument");
  if (!h->ngp_configured) return fail(EVPP_ERR_STATE, "evpp_ngp_configure first");
  const int64_t G = h->ngp.grid_res;
  const int64_t want = G * G * G / 8 * (h->ngp.march_mode == 1 ? h->ngp.cascades : 1);
  if (nbytes != want) return fail(EVPP_ERR_INVALID, "bitfield must hold %sgrid_res^3 bits (%lld bytes)",
                                  h->ngp.march_mode == 1 ? "cascades x " : "", (long long)want);
  EVPP_ON_DEVICE(h->device);
  if (h->ngp_bits.ensure((size_t)nbytes)) return fail(EVPP_ERR_CUDA, "bitfield alloc failed");
  CUDA_TRY(cudaMemcpy(h->ngp_bits.p, bitfield, (size_t)nbytes, cudaMemcpyDefault));
  h->ngp.bitfield = h->ngp_bits.as<uint8_t>();
  for (int c = 0; c < 3; ++c) { h->ngp.occ_lo[c] = 1.f; h->ngp.occ_hi[c] = 0.f; }       // "nothing occupied"
  if (h->ngp.march_mode == 0) {
    // Box of the occupied cells (cell (x*G+y)*G+z <-> bit (cell & 7) of byte cell >> 3), grown by one cell: the marcher
    // only tests the steps of a ray that fall inside it (ngp.cu:ray_span).  One pass over the bitfield on the host.
    std::vector<uint8_t> host((size_t)nbytes);
    CUDA_TRY(cudaMemcpy(host.data(), h->ngp_bits.p, (size_t)nbytes, cudaMemcpyDeviceToHost));
    int64_t lo[3] = {G, G, G}, hi[3] = {-1, -1, -1};
    for (int64_t x = 0; x < G; ++x)
      for (int64_t y = 0; y < G; ++y) {
        const int64_t row = (x * G + y) * G;            // G is a multiple of 8 (checked by evpp_ngp_configure): whole bytes
        int64_t z0 = -1, z1 = -1;
        for (int64_t b = 0; b < G / 8; ++b) {
          const uint8_t v = host[(size_t)(row / 8 + b)];
          if (!v) continue;
          for (int k = 0; k < 8; ++k)
            if (v >> k & 1) { if (z0 < 0) z0 = b * 8 + k; z1 = b * 8 + k; }
        }
        if (z0 < 0) continue;
        lo[0] = std::min(lo[0], x); hi[0] = std::max(hi[0], x);
        lo[1] = std::min(lo[1], y); hi[1] = std::max(hi[1], y);
        lo[2] = std::min(lo[2], z0); hi[2] = std::max(hi[2], z1);
      }
    if (hi[0] >= 0)
      for (int c = 0; c < 3; ++c) {
        const double cell = ((double)h->ngp.amax[c] - (double)h->ngp.amin[c]) / (double)G;
        h->ngp.occ_lo[c] = (float)((double)h->ngp.amin[c] + (double)(lo[c] - 1) * cell);
        h->ngp.occ_hi[c] = (float)((double)h->ngp.amin[c] + (double)(hi[c] + 2) * cell);
      }
  }
  h->ngp_has_occ = true;
  return 0;
}

int evpp_ngp_load_params(evpp_handle* h, const uint16_t* table_fp16, int64_t n_entries, const float* mlp_packed,
                         int64_t n_mlp_floats) {
  if (!h || !table_fp16 || !mlp_packed) return fail(EVPP_ERR_INVALID, "null argument");
  if (!h->ngp_configured) return fail(EVPP_ERR_STATE, "evpp_ngp_configure first");
  const NgpLevel& last = h->ngp.levels[h->ngp.n_levels - 1];
  if (n_entries != (int64_t)last.offset + last.size) return fail(EVPP_ERR_INVALID, "table has %lld entries, level table needs %lld", (long long)n_entries, (long long)last.offset + last.size);
  if (n_mlp_floats != kNgpMlpFloats) return fail(EVPP_ERR_INVALID, "packed MLP must hold %d floats", kNgpMlpFloats);
  EVPP_ON_DEVICE(h->device);
  if (h->ngp_table.ensure((size_t)n_entries * 4) || h->ngp_mlp.ensure((size_t)kNgpMlpFloats * 4))
    return fail(EVPP_ERR_CUDA, "NGP parameter alloc failed");
  CUDA_TRY(cudaMemcpy(h->ngp_table.p, table_fp16, (size_t)n_entries * 4, cudaMemcpyDefault));
  CUDA_TRY(cudaMemcpy(h->ngp_mlp.p, mlp_packed, (size_t)kNgpMlpFloats * 4, cudaMemcpyDefault));
  h->ngp.table = h->ngp_table.as<uint16_t>();
  h->ngp.mlp = h->ngp_mlp.as<float>();
  h->ngp_has_params = true;
  h->ngp_tensorf = false;
  return 0;
}

int evpp_tensorf_load_params(evpp_handle* h, const uint16_t* planes_fp16, const uint16_t* lines_fp16, int32_t res,
                             const float* mlp_packed, int64_t n_mlp_floats) {
  if (!h || !planes_fp16 || !lines_fp16 || !mlp_packed) return fail(EVPP_ERR_INVALID, "null argument");
  if (!h->ngp_configured) return fail(EVPP_ERR_STATE, "evpp_ngp_configure first");
  if (res < 2 || res > 2048) return fail(EVPP_ERR_INVALID, "grid resolution out of range");
  if (n_mlp_floats != kTfMlpFloats) return fail(EVPP_ERR_INVALID, "packed TensoRF colour network must hold %d floats", kTfMlpFloats);
  EVPP_ON_DEVICE(h->device);
  const size_t pb = (size_t)3 * res * res * 64 * 2, lb = (size_t)3 * res * 64 * 2;
  if (h->tf_planes.ensure(pb) || h->tf_lines.ensure(lb) || h->tf_mlp.ensure((size_t)kTfMlpFloats * 4))
    return fail(EVPP_ERR_CUDA, "TensoRF parameter alloc failed");
  CUDA_TRY(cudaMemcpy(h->tf_planes.p, planes_fp16, pb, cudaMemcpyDefault));
  CUDA_TRY(cudaMemcpy(h->tf_lines.p, lines_fp16, lb, cudaMemcpyDefault));
  CUDA_TRY(cudaMemcpy(h->tf_mlp.p, mlp_packed, (size_t)kTfMlpFloats * 4, cudaMemcpyDefault));
  h->ngp.tf_planes = h->tf_planes.as<uint16_t>();
  h->ngp.tf_lines = h->tf_lines.as<uint16_t>();
  h->ngp.tf_mlp = h->tf_mlp.as<float>();
  h->ngp.tf_res = res;
  h->ngp_has_params = true;
  h->ngp_tensorf = true;
  return 0;
}

int evpp_ngp_encode(evpp_handle* h, const float* u, int64_t M, float* enc, int32_t* indices, void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (M == 0) return 0;
  if (!u || M < 0) return fail(EVPP_ERR_INVALID, "bad argument");
  if (!h->ngp_configured || !h->ngp_has_params || h->ngp_tensorf)
    return fail(EVPP_ERR_STATE, "hash-grid encoder not configured / no hash-grid parameters loaded");
  EVPP_ON_DEVICE(h->device);
  h->launches += launch_ngp_encode(h->ngp, u, M, enc, indices, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// march -> scan -> fill -> field -> composite for rays already on the device; leaves per-ray beta2 in `beta2`.
// No host round trip: the sample buffers are sized for the worst case (rays x max_samples) once, and the kernels
// after the scan read the chunk's sample total from device memory.  Only a caller that asks for per-sample dumps
// (evpp_ngp_render_rays with t / raw / enc / delta / cell) makes the host wait for the total.
static int ngp_render_device(evpp_handle* h, const float* ro, const float* rd, int n, evpp_ngp_out* out, long long off,
                             float* beta2, cudaStream_t st) {
  NvtxRange nvtx_chunk("evpp:ngp_chunk");
  const NgpParams& P = h->ngp;
  const bool tngp = P.march_mode == 1;
  const long long cap = (long long)n * P.max_samples + 1;
  if (h->ngp_counts.ensure((size_t)n * 4) || h->ngp_offsets.ensure((size_t)n * 4) || h->ngp_total.ensure(4) ||
      h->ngp_tiles.ensure((size_t)2 * ngp_scan_tiles(n) * 4) ||
      (!tngp && (h->ngp_masks.ensure((size_t)n * ngp_mask_words(P) * 4) || h->ngp_t0.ensure((size_t)n * 4))) ||
      (tngp && h->ngp_delta.ensure((size_t)cap * 4)) ||
      h->ngp_t.ensure((size_t)cap * 4) || h->ngp_ray.ensure((size_t)cap * 4) || h->ngp_raw.ensure((size_t)cap * 20))
    return fail(EVPP_ERR_CUDA, "NGP workspace alloc failed (%d rays x %d samples)", n, P.max_samples);
  const bool want_cells = tngp && out && out->cell;
  if (want_cells && h->ngp_cell.ensure((size_t)cap * 4)) return fail(EVPP_ERR_CUDA, "NGP workspace alloc failed");
  int32_t* counts = out && out->counts ? out->counts + off : h->ngp_counts.as<int32_t>();
  int32_t* offsets = h->ngp_offsets.as<int32_t>();
  const int32_t* total_dev = h->ngp_total.as<int32_t>();
  if (tngp)
    h->launches += launch_tngp_march_count(P, ro, rd, n, counts, st);
  else
    h->launches += launch_ngp_march_count(P, ro, rd, n, counts, h->ngp_masks.as<uint32_t>(), h->ngp_t0.as<float>(), st);
  h->launches += launch_ngp_scan(counts, n, offsets, h->ngp_total.as<int32_t>(), h->ngp_tiles.as<int32_t>(),
                                 h->ngp_counter.as<unsigned long long>(), st);
  if (tngp)
    h->launches += launch_tngp_march_fill(P, ro, rd, n, counts, offsets, h->ngp_t.as<float>(), h->ngp_ray.as<int32_t>(),
                                          h->ngp_delta.as<float>(), want_cells ? h->ngp_cell.as<int32_t>() : nullptr, st);
  else
    h->launches += launch_ngp_march_fill(P, n, counts, offsets, h->ngp_masks.as<uint32_t>(), h->ngp_t0.as<float>(),
                                          h->ngp_t.as<float>(), h->ngp_ray.as<int32_t>(), st);
  const bool dumps = out && (out->t || out->raw || out->enc || out->delta || out->cell);
  long long M = -1;
  if (dumps) {      // diagnostics path: the host needs the total to place the per-sample dumps
    int32_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, h->ngp_total.p, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    M = total;
  }
  const bool room = dumps && out->sample_capacity >= out->total_so_far + M;
  float* enc_dump = room && out->enc ? out->enc + out->total_so_far * 32 : nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->profiling) {       // the field kernel (encode + MLP) is this mode's dominant kernel: timed like the NeRF MLP
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
  }
  if (h->ngp_tensorf)
    h->launches += launch_tensorf_field(P, ro, rd, h->ngp_t.as<float>(), h->ngp_ray.as<int32_t>(), cap - 1, total_dev,
                                        h->ngp_raw.as<float>(), h->num_sms, st);
  else
    h->launches += launch_ngp_field(P, ro, rd, h->ngp_t.as<float>(), h->ngp_ray.as<int32_t>(), cap - 1, total_dev,
                                    h->ngp_raw.as<float>(), enc_dump, h->num_sms, h->cfg.precision != EVPP_PREC_FP32, st);
  if (h->profiling) {
    cudaEventRecord(e1, st);
    h->prof_events.emplace_back(e0, e1);
    h->prof_evals.push_back(0);      // the sample count lives on the device (evpp_ngp_get_sample_count)
    h->prof_flops.push_back(0.0);
  }
  h->launches += launch_ngp_composite(P, h->ngp_raw.as<float>(), h->ngp_t.as<float>(), tngp ? h->ngp_delta.as<float>() : nullptr,
                                      counts, offsets, n,
                                      out && out->rgb ? out->rgb + off * 3 : nullptr, out && out->depth ? out->depth + off : nullptr,
                                      out && out->acc ? out->acc + off : nullptr, beta2, st);
  if (room) {
    if (out->t) CUDA_TRY(cudaMemcpyAsync(out->t + out->total_so_far, h->ngp_t.p, (size_t)M * 4, cudaMemcpyDeviceToDevice, st));
    if (out->raw) CUDA_TRY(cudaMemcpyAsync(out->raw + out->total_so_far * 5, h->ngp_raw.p, (size_t)M * 20, cudaMemcpyDeviceToDevice, st));
    if (out->delta && tngp) CUDA_TRY(cudaMemcpyAsync(out->delta + out->total_so_far, h->ngp_delta.p, (size_t)M * 4, cudaMemcpyDeviceToDevice, st));
    if (want_cells) CUDA_TRY(cudaMemcpyAsync(out->cell + out->total_so_far, h->ngp_cell.p, (size_t)M * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (dumps) out->total_so_far += M;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// rays per marching chunk: the worst-case sample buffers (rays x max_samples x 32 B) stay below ~8.6 GB
static int ngp_chunk_cap(evpp_handle* h) {
  const long long by_samples = (1ll << 28) / std::max(1, h->ngp.max_samples);
  return (int)std::max<long long>(1, std::min<long long>(h->ngp_chunk_rays, by_samples));
}

static int ngp_ready(evpp_handle* h) {
  if (!h->ngp_configured || !h->ngp_has_occ || !h->ngp_has_params)
    return fail(EVPP_ERR_STATE, "NGP scorer needs evpp_ngp_configure, evpp_ngp_set_occupancy and evpp_ngp_load_params");
  return 0;
}

int evpp_ngp_render_rays(evpp_handle* h, const float* rays_o, const float* rays_d, int64_t N, evpp_ngp_out* out, void* stream) {
  if (!h || !out) return fail(EVPP_ERR_INVALID, "null argument");
  out->total_so_far = 0;
  if (N == 0) return 0;
  if (!rays_o || !rays_d || N < 0) return fail(EVPP_ERR_INVALID, "bad rays");
  int rc = ngp_ready(h);
  if (rc) return rc;
  EVPP_ON_DEVICE(h->device);
  const int cap = ngp_chunk_cap(h);
  for (int64_t g0 = 0; g0 < N; g0 += cap) {
    const int n = (int)std::min<int64_t>(cap, N - g0);
    rc = ngp_render_device(h, rays_o + g0 * 3, rays_d + g0 * 3, n, out, g0, out->beta2 ? out->beta2 + g0 : nullptr, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

int evpp_ngp_score_views(evpp_handle* h, const float* c2w, int V, int H, int W, float fx, float fy, float cx, float cy,
                         const int32_t* pix_idx, int R, float* scores, void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (V == 0) return 0;
  if (!c2w || !scores || V < 0 || H <= 0 || W <= 0) return fail(EVPP_ERR_INVALID, "bad argument");
  int rc = ngp_ready(h);
  if (rc) return rc;
  if (!pix_idx) R = H * W;
  if (R <= 0) return fail(EVPP_ERR_INVALID, "R must be positive when pix_idx is given");
  EVPP_ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = (long long)V * R;
  const int cap = (int)std::min<long long>(N, ngp_chunk_cap(h));
  if (h->ngp_ro.ensure((size_t)cap * 12) || h->ngp_rd.ensure((size_t)cap * 12) || h->ngp_beta2.ensure((size_t)N * 4))
    return fail(EVPP_ERR_CUDA, "NGP ray workspace alloc failed");
  for (long long g0 = 0; g0 < N; g0 += cap) {
    const int n = (int)std::min<long long>(cap, N - g0);
    h->launches += launch_get_rays(c2w, V, H, W, fx, fy, cx, cy, pix_idx, R, g0, n, h->ngp_ro.as<float>(), h->ngp_rd.as<float>(), nullptr, st);
    rc = ngp_render_device(h, h->ngp_ro.as<float>(), h->ngp_rd.as<float>(), n, nullptr, 0, h->ngp_beta2.as<float>() + g0, st);
    if (rc) return rc;
  }
  h->launches += launch_view_reduce(h->ngp_beta2.as<float>(), V, R, scores, nullptr, st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int evpp_ngp_score_views_host(evpp_handle* h, const float* c2w_host, int V, int H, int W, float fx, float fy, float cx,
                              float cy, const int32_t* pix_idx_host, int R, float* scores_host, void* stream) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (V == 0) return 0;
  if (!c2w_host || !scores_host) return fail(EVPP_ERR_INVALID, "null argument");
  EVPP_ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->st_c2w.ensure((size_t)V * 48) || h->ngp_scores.ensure((size_t)V * 4)) return fail(EVPP_ERR_CUDA, "staging alloc failed");
  CUDA_TRY(cudaMemcpyAsync(h->st_c2w.p, c2w_host, (size_t)V * 48, cudaMemcpyHostToDevice, st));
  const int32_t* pix_dev = nullptr;
  if (pix_idx_host) {
    if (h->st_pix.ensure((size_t)V * R * 4)) return fail(EVPP_ERR_CUDA, "staging alloc failed");
    CUDA_TRY(cudaMemcpyAsync(h->st_pix.p, pix_idx_host, (size_t)V * R * 4, cudaMemcpyHostToDevice, st));
    pix_dev = h->st_pix.as<int32_t>();
  }
  int rc = evpp_ngp_score_views(h, h->st_c2w.as<float>(), V, H, W, fx, fy, cx, cy, pix_dev, R, h->ngp_scores.as<float>(), stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(scores_host, h->ngp_scores.p, (size_t)V * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int evpp_ngp_get_sample_count(evpp_handle* h, int64_t* samples) {
  if (!h || !samples) return fail(EVPP_ERR_INVALID, "null argument");
  *samples = 0;
  if (!h->ngp_counter.p) return 0;
  EVPP_ON_DEVICE(h->device);
  unsigned long long v = 0;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(&v, h->ngp_counter.p, 8, cudaMemcpyDeviceToHost));
  *samples = (int64_t)v;
  return 0;
}

int evpp_check_guards(evpp_handle* h, int32_t* corrupted_bytes, int32_t* buffers_checked) {
  if (!h || !corrupted_bytes) return fail(EVPP_ERR_INVALID, "null argument");
  EVPP_ON_DEVICE(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  int bad = 0, n = 0;
  for (DevBuf* b : all_buffers(h)) {
    if (!b->p) continue;
    const int c = b->check();
    if (c < 0) return fail(EVPP_ERR_CUDA, "guard read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
    bad += c;
    ++n;
  }
  *corrupted_bytes = bad;
  if (buffers_checked) *buffers_checked = n;
  return 0;
}

int evpp_poison_workspace(evpp_handle* h) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  EVPP_ON_DEVICE(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  // 0xFF bytes = NaN as fp32, -1 as int32: a kernel that reads a workspace element nobody wrote in THIS call shows it
  for (DevBuf* b : {&h->rays_o, &h->rays_d, &h->viewdirs, &h->z_c, &h->raw_c, &h->z_f, &h->raw_f, &h->beta2, &h->st_scores,
                    &h->w_c, &h->cdf, &h->inds, &h->zs, &h->tsdf_ro, &h->tsdf_rd, &h->tsdf_u, &h->tsdf_d, &h->tsdf_vu, &h->tsdf_vd,
                    &h->ngp_ro, &h->ngp_rd, &h->ngp_counts, &h->ngp_offsets, &h->ngp_masks, &h->ngp_tiles, &h->ngp_t0, &h->ngp_t,
                    &h->ngp_ray, &h->ngp_raw, &h->ngp_beta2, &h->ngp_scores, &h->ngp_delta, &h->ngp_cell})
    if (b->p) CUDA_TRY(cudaMemset(b->p, 0xFF, b->bytes));
  return 0;
}

int evpp_get_counters(evpp_handle* h, int64_t* kernel_launches, int64_t* mlp_evals) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  if (kernel_launches) *kernel_launches = h->launches;
  if (mlp_evals) *mlp_evals = h->evals;
  return 0;
}

int evpp_set_profiling(evpp_handle* h, int enabled) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  h->profiling = enabled != 0;
  for (auto& pr : h->prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  h->prof_events.clear();
  h->prof_evals.clear();
  h->prof_flops.clear();
  h->prof_ms = 0.0;
  h->prof_total_flops = 0.0;
  h->prof_launches = 0;
  h->prof_total_evals = 0;
  return 0;
}

static int collect_profile(evpp_handle* h) {
  EVPP_ON_DEVICE(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  for (size_t i = 0; i < h->prof_events.size(); ++i) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_events[i].first, h->prof_events[i].second));
    h->prof_ms += ms;
    h->prof_launches += 1;
    h->prof_total_evals += h->prof_evals[i];
    h->prof_total_flops += h->prof_flops[i];
    cudaEventDestroy(h->prof_events[i].first);
    cudaEventDestroy(h->prof_events[i].second);
  }
  h->prof_events.clear();
  h->prof_evals.clear();
  h->prof_flops.clear();
  return 0;
}

int evpp_get_mlp_time_ms(evpp_handle* h, double* total_ms, int64_t* launches, int64_t* evals) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  int rc = collect_profile(h);
  if (rc) return rc;
  if (total_ms) *total_ms = h->prof_ms;
  if (launches) *launches = h->prof_launches;
  if (evals) *evals = h->prof_total_evals;
  return 0;
}

int evpp_get_mlp_profile(evpp_handle* h, double* total_ms, int64_t* launches, int64_t* evals, double* flops) {
  if (!h) return fail(EVPP_ERR_INVALID, "null handle");
  int rc = collect_profile(h);
  if (rc) return rc;
  if (total_ms) *total_ms = h->prof_ms;
  if (launches) *launches = h->prof_launches;
  if (evals) *evals = h->prof_total_evals;
  if (flops) *flops = h->prof_total_flops;
  return 0;
}

}  // extern "C"
