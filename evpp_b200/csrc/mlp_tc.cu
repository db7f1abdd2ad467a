// Fused positional encoding + NeRF/NeurAR MLP on the 5th-generation tensor cores (sm_100a).
// Replaces Embedder.embed (run_nerf_helpers.py:22-55), run_network (run_nerf.py:84-113) and
// NeRF.forward (run_nerf_helpers.py:108-134) of the reference for the bulk (fp16 / bf16 operand,
// fp32 accumulate) pass.  The fp32 twin is mlp_simt.cu.
//
// One persistent CTA per SM, in clusters of 2 CTAs sharing the weight stream by bulk-copy multicast.  A CTA owns one
// 128-row tile at a time and runs all ten GEMM layers on it without the activations ever leaving tensor memory:
//
//   shared memory   X      16 KB       the 63-ch position embedding (layers 0 and 5), K-major, 128-byte swizzle
//                   XD      8 KB       the 27-ch direction embedding (view layer), 64-byte swizzle
//                   ring   6 x 32 KB   weight K-blocks W[:, 64k..64k+63] ([N rows x 128 B], pre-swizzled on the host)
//                                      streamed from L2 with cp.async.bulk(.multicast) + mbarrier transaction counts
//                   bias   10 KB       fp32 biases of the ten layers
//   tensor memory   2 x 256 columns    layer l accumulates (fp32) into region l & 1; the epilogue packs the 16-bit
//                                      activations IN PLACE (16 fp32 columns -> 8 packed columns at the same offset),
//                                      which is the A-operand layout of layer l+1's TS-mode MMAs (D = other region)
//
//   warps 0-15        encode the tile, then per layer drain the accumulator 64 columns at a time (tcgen05.ld, bias,
//                     ReLU + pack in one F2FP.RELU, tcgen05.st); layer l+1's K-block k is issued as soon as chunk k
//                     is stored.  alpha / rgb / uncertainty heads are fp32 FFMA on the un-rounded activations.
//   warp 16 (1 lane)  weight producer: 39 bulk copies per tile through the 6-stage ring
//   warp 17 (1 lane)  tcgen05.mma issuer (elect.sync, so each MMA is a single uniform-datapath instruction):
//                     D[128 x N] += A[128 x 16] * W[N x 16]^T, 4 per K-block.  The last K-blocks of a 256-wide layer
//                     run as two N = 128 halves (columns 0..127 first), so the drain of the first half overlaps
//                     the second half's MMAs (SPLIT template parameter).
//
// Per row the kernel reads 4 B (z) and writes 20 B (render()) or 4 B (scoring calls).
//
// Template axes added in round 2 (documented at mlp_tc_kernel):
//   NP     1 = plain 16-bit operands (bulk pass); 2 / 3 = split-fp16 passes (W = W_hi + W_lo, A = A_hi + A_lo, two / three
//          MMAs per product, fp32 accumulate): fp32-class results on the tensor cores -- the exact path that re-scores the
//          near-tied candidate views so that the planner's arg-max is the fp32 reference's.  A_lo lives in the 8 tensor-
//          memory columns the in-place packing leaves free, W_lo is a second weight image in the same ring (5 stages).
//   HEADS  score-only variants for Controller.get_uncertainty (nerf.py:307-309 reads raw[..., 4] of the fine pass only):
//          1 = sigma only (coarse pass, stops after layer 7), 2 = beta only, 3 = beta only with feature_linear multiplied
//          into views_linears.0 on the host (no activation between them: one linear map, nine GEMM layers).
// Every weight image is scaled per layer by a power of two (undone in the epilogue's bias FMA) and the epilogue's ReLU
// keeps NaN, so neither an underflow of tiny trained weights nor an overflow of fp16 activations can pass silently.
#include <cmath>

#include "common.cuh"

namespace evpp {

namespace {

constexpr int kTileM = 128;
constexpr int kEpiWarps = 16;              // 4 per TMEM lane quarter
constexpr int kColSplit = kEpiWarps / 4;   // warps sharing one 64-column chunk of a row quarter
constexpr int kColsPerWarp = 64 / kColSplit;
constexpr int kThreads = 32 * (kEpiWarps + 2);
// Warp roles: 0..kEpiWarps-1 epilogue (TMEM lane quarter = warp & 3); the two control warps get the HIGHEST
// warp ids because the SM's issue arbiter favours high warp ids: the MMA issuer must never wait for an
// issue slot behind busy epilogue warps.
constexpr int kProducerWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kMaxStages = 6;
constexpr int kStageBytes = 32768;
constexpr int kBlkBytes = 16384;          // 128 rows x 128 B
constexpr int kLayers = 10;               // pts_linears 0..7, feature_linear (8), views_linears.0 (9)

struct Misc {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[4];   // [accumulator region][column half]
  uint64_t a_ready[4];
  uint64_t x_ready;
  uint64_t xd_ready;
  uint32_t tmem_base;
  uint32_t pad_;
};

// Shared-memory layout.  NP = number of tensor-core passes per product (see mlp_tc_kernel): with NP > 1 every
// operand exists as a (hi, lo) pair of fp16 images, so the embeddings take twice the room and the weight ring
// gives up one stage for them.
template <int NP>
struct Lay {
  static constexpr int kStages = NP > 1 ? 5 : 6;
  static constexpr int kOffX = 0;                                     // position embedding (hi): 128 rows x 128 B, 128-byte swizzle
  static constexpr int kOffXlo = kOffX + kBlkBytes;                   // its lo half (NP > 1 only)
  static constexpr int kOffXD = NP > 1 ? 2 * kBlkBytes : kBlkBytes;   // direction embedding (hi): 128 rows x 64 B, 64-byte swizzle
  static constexpr int kOffXDlo = kOffXD + kBlkBytes / 2;             // its lo half (NP > 1 only)
  static constexpr int kOffRing = kOffXD + (NP > 1 ? kBlkBytes : kBlkBytes / 2);
  static constexpr int kOffBias = kOffRing + kStages * kStageBytes;   // fp32 bias[10][256]
  static constexpr int kOffMisc = kOffBias + kLayers * 256 * 4;
  static constexpr int kSmemBytes = kOffMisc + (int)sizeof(Misc);
  static_assert(kOffRing % 1024 == 0, "weight K-blocks need 1024-byte alignment (128-byte swizzle atoms)");
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
};

// per-tile weight stream: number of 64-wide K-blocks per layer and their size
// `view` = index of the view layer: 9 in the reference's layer order, 8 when feature_linear is folded into it
__host__ __device__ constexpr int layer_has_x(int l, int view = 9) { return l == 0 || l == 5 || l == view; }
__host__ __device__ constexpr int layer_h_blocks(int l) { return l == 0 ? 0 : 4; }
__host__ __device__ constexpr int layer_n(int l, int view = 9) { return l == view ? 128 : 256; }
constexpr size_t kImageBytes = (size_t)34 * 32768 + (size_t)5 * 16384;

// ---------------------------------------------------------------- PTX helpers --------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
          "r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                   "r"(bar), "h"(mask)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M=128, K=16, kind::f16 (fp16 or bf16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory (lane = row, 8 columns = 16 packed 16-bit channels)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout_type SWIZZLE_128B=2 [61,64))
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// same for 64-byte rows (32 channels): 64-byte swizzle (layout_type 4), 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format f32 (1) [4,6), a/b format [7,10)/[10,13)
// (0 = f16, 1 = bf16), K-major A and B, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, int n) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// One lane of a converged warp.  Unlike `lane == 0`, elect.sync lets the compiler prove that a single thread runs
// the region, so each tcgen05.mma / tcgen05.commit becomes ONE uniform-datapath instruction instead of an
// ELECT / branch loop over the "possibly divergent" active lanes (5 instructions and a branch per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

// pack2 of max(x, 0): one F2FP.RELU instead of a conversion and a packed max
template <bool BF16>
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {
  uint32_t r;
  if (BF16) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// byte offset of 16-byte chunk `ch` (8 channels) of row `row` inside a swizzled 16 KB K-block
__device__ __forceinline__ uint32_t sw128(int row, int ch) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4));
}

// fp32 -> (hi, lo) fp16 pair with hi + lo == x to ~22 significant bits (lo is subnormal below |x| ~ 0.25, where its
// absolute resolution 2^-24 is still at the fp32 rounding level of O(1) activations)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Frequency encoding of one coordinate triple.
// EXACT = false: angle doubling from an accurate sincosf at the base frequency: the 2^k-th octave carries <= 2^k
// ulp-level phase error (<= 6e-5 at k = 9), below the 2.4e-4 rounding step of the 16-bit operand it is converted to.
// EXACT = true (the split-fp16 passes): one accurate sincosf per octave on the exactly scaled argument, like the
// reference's torch.sin(x * 2^k) (run_nerf_helpers.py:46-49), stored as a (hi, lo) pair of blocks lo_delta bytes apart.
template <bool BF16, int L, bool SW64, bool EXACT>
__device__ __forceinline__ void encode_store(const float (&p)[3], uint32_t row_base, uint32_t lo_delta, int row,
                                             int nchunks) {
  float e[64];
  e[0] = p[0]; e[1] = p[1]; e[2] = p[2];
  if (EXACT) {
#pragma unroll
    for (int l = 0; l < L; ++l) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float sn, cs;      // (not &e[..]: taking the array's address would pin all of e[] in local memory)
        sincosf(p[i] * (float)(1 << l), &sn, &cs);
        e[3 + l * 6 + i] = sn;
        e[3 + l * 6 + 3 + i] = cs;
      }
    }
  } else {
    float s[3], c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) sincosf(p[i], &s[i], &c[i]);
#pragma unroll
    for (int l = 0; l < L; ++l) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        e[3 + l * 6 + i] = s[i];
        e[3 + l * 6 + 3 + i] = c[i];
        const float s2 = 2.f * s[i] * c[i];
        const float c2 = fmaf(-2.f * s[i], s[i], 1.f);
        s[i] = s2; c[i] = c2;
      }
    }
  }
#pragma unroll
  for (int k = 3 + 6 * L; k < 64; ++k) e[k] = 0.f;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    if (ch < nchunks) {
      const int sw = SW64 ? (ch ^ ((row >> 1) & 3)) : (ch ^ (row & 7));
      if (EXACT) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(e[ch * 8 + 2 * j], e[ch * 8 + 2 * j + 1], hi[j], lo[j]);
        st_shared_v4(row_base + (sw << 4), hi[0], hi[1], hi[2], hi[3]);
        st_shared_v4(row_base + lo_delta + (sw << 4), lo[0], lo[1], lo[2], lo[3]);
      } else {
        st_shared_v4(row_base + (sw << 4), pack2<BF16>(e[ch * 8 + 0], e[ch * 8 + 1]),
                     pack2<BF16>(e[ch * 8 + 2], e[ch * 8 + 3]), pack2<BF16>(e[ch * 8 + 4], e[ch * 8 + 5]),
                     pack2<BF16>(e[ch * 8 + 6], e[ch * 8 + 7]));
      }
    }
  }
}

struct TcArgs {
  const uint8_t* image;
  RayBatch rb;
  const float* pts;
  const float* dirs;
  long long M;
  float* out;
  float* dbg;        // optional [M,256] dump of the activations written by layer dbg_layer
  int dbg_layer;
  int tiles_per_cta;
  long long tile0;     // this launch owns tiles [tile0, tile_end): CTA b takes tile0 + b + it * gridDim.x
  long long tile_end;
  int pdl;             // 1: main grid of a hybrid launch (releases its dependent), 2: the dependent filler grid, 0: neither
  long long* trace;  // optional [128] clock64 stamps of CTA 0's third tile (kernels built with DM != 0 only)
};

__device__ __forceinline__ long long clk() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}
#define EVPP_TRACE(idx)                                                              \
  do {                                                                               \
    if (TR && args.trace && blockIdx.x == 0 && it == 2) args.trace[(idx)] = clk();   \
  } while (0)

#ifdef EVPP_TC_DIAG
__device__ unsigned long long g_diag[2][160][3];   // [4-cluster | 2-cluster launch][CTA][start ns, end ns, smid]
#endif

struct RowData {     // what one epilogue thread knows about its row of the current tile
  float p[3];
  float dv[3];
  long long m;
  bool valid;
};

template <bool RAYS>
__device__ __forceinline__ void load_row(const TcArgs& args, long long tile, long long ntiles, int row, RowData& rd) {
  rd.m = tile * kTileM + row;
  rd.valid = tile < ntiles && rd.m < args.M;
  rd.p[0] = rd.p[1] = rd.p[2] = 0.f;
  rd.dv[0] = rd.dv[1] = 0.f; rd.dv[2] = 1.f;
  if (rd.valid) {
    if (RAYS) {
      const long long ray = rd.m / args.rb.S;
      const float z = args.rb.z[rd.m];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // pts = rays_o + rays_d * z (run_nerf.py:466), product and sum rounded separately like ATen
        rd.p[c] = __fadd_rn(args.rb.rays_o[ray * 3 + c], __fmul_rn(args.rb.rays_d[ray * 3 + c], z));
        rd.dv[c] = args.rb.viewdirs[ray * 3 + c];
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) { rd.p[c] = args.pts[rd.m * 3 + c]; rd.dv[c] = args.dirs[rd.m * 3 + c]; }
    }
  }
}

// ReLU that keeps NaN (fmaxf(NaN, 0) would return 0 and hide an overflow of the 16-bit activations upstream: the
// non-finite score check of the engine relies on NaN / inf reaching the outputs)
__device__ __forceinline__ float relu_nan(float x) {
  float r;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x));
  return r;
}
// two fp32 fused multiply-adds in one instruction: a * s + b
__device__ __forceinline__ float2 fma2(float2 a, float s, float2 b) {
  float2 c;
  asm("{\n\t.reg .b64 ra, rs, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rs, {%4, %4};\n\tmov.b64 rb, {%5, %6};\n\t"
      "fma.rn.f32x2 rc, ra, rs, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(c.x), "=f"(c.y)
      : "f"(a.x), "f"(a.y), "f"(s), "f"(b.x), "f"(b.y));
  return c;
}
// Drains one layer's accumulator.  KIND 0: ReLU layer -> next A operand; 1: layer 7 (ReLU + alpha head);
// 2: feature_linear (no ReLU); 3: view layer (ReLU + rgb / uncertainty heads, nothing stored);
// 4: layer 7 of a sigma-only pass (ReLU + alpha head, nothing stored: it is the last layer);
// 5: view layer of a beta-only pass (ReLU + uncertainty head only).
// A warp owns kColsPerWarp columns of each 64-column chunk of its 32 rows; the TMEM load of chunk k+1 is in
// flight while chunk k is converted and stored.
// HFC >= 0 makes the warp's column slice a compile-time constant, so that the head weights of layers 7 and 9 become
// immediate constant-bank operands of the FFMAs instead of indexed constant loads.
// The accumulator carries the layer's power-of-two weight scale (inv_ws undoes it, exactly, in the same FMA that adds
// the bias).  NP == 3 stores the activations as a (hi, lo) fp16 pair: hi in the 8 packed columns at c0, lo in the
// 8 columns behind them.
template <bool BF16, int KIND, bool DBG, bool TR, int NP, int HFC = -1>
__device__ __forceinline__ void drain_layer(const TcConsts& cst, const TcArgs& args, const float* bias_l, int l,
                                            uint32_t t_acc, int hf, int lane, uint32_t bar_a, uint32_t bar_acc_l,
                                            uint32_t acc_parity, const RowData& rd, float& alpha, float (&hd)[4],
                                            long long* tr = nullptr, long long* tr0 = nullptr) {
  if (HFC >= 0) hf = HFC;
  constexpr bool kStore = KIND <= 2;
  constexpr bool kRelu = KIND != 2;
  constexpr int NCH = (KIND == 3 || KIND == 5) ? 2 : 4;
  constexpr int NC = kColsPerWarp;   // 16
  // The tensor-memory load of chunk k+1 is in flight while chunk k is converted (two register buffers).  Measured for
  // the fp16x3 pass too: a single buffer saves its ~80 bytes of spills and costs 9 % (115 vs 126 views/s).
  constexpr bool kPrefetch = true;
  uint32_t r[kPrefetch ? 2 : 1][NC];
  const float inv_ws = cst.inv_wscale[l];
  // the first chunk is on the critical path (the next layer's MMAs wait for it): fetch its biases before the
  // accumulator is ready
  float4 bpre[NC / 4];
#pragma unroll
  for (int j = 0; j < NC / 4; ++j) bpre[j] = reinterpret_cast<const float4*>(bias_l + hf * NC)[j];
  // columns 0..127 and 128..255 of the accumulator become final separately (bar_acc_l, bar_acc_l + 8): the MMA
  // issuer runs the last K-blocks of a layer as two N = 128 halves so that this drain starts while the second
  // half is still on the tensor pipe
  mbar_wait(bar_acc_l, acc_parity);
  tc_fence_after();
  if (TR && tr && lane == 0) tr[0] = clk();
  tmem_ld16(t_acc + (uint32_t)(hf * NC), r[0]);
  bool half1_early = false;
#pragma unroll
  for (int kb = 0; kb < NCH; ++kb) {
    const int rb = kPrefetch ? (kb & 1) : 0;
    tmem_ld_wait();
    if (kPrefetch && kb + 1 < NCH) {
      if (kb + 1 == 2) {
        half1_early = __all_sync(0xffffffffu, mbar_test(bar_acc_l + 8, acc_parity));
        if (half1_early) {
          tc_fence_after();
          tmem_ld16(t_acc + (uint32_t)((kb + 1) * 64 + hf * NC), r[kPrefetch ? ((kb + 1) & 1) : 0]);
        }
      } else {
        tmem_ld16(t_acc + (uint32_t)((kb + 1) * 64 + hf * NC), r[kPrefetch ? ((kb + 1) & 1) : 0]);
      }
    }
    const int c0 = kb * 64 + hf * NC;
    float2 v[NC / 2];
    const float4* b4 = reinterpret_cast<const float4*>(bias_l + c0);   // shared memory, warp-uniform address
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      const float4 b = kb == 0 ? bpre[j] : b4[j];
      const float2 a0 = make_float2(__uint_as_float(r[rb][4 * j]), __uint_as_float(r[rb][4 * j + 1]));
      const float2 a1 = make_float2(__uint_as_float(r[rb][4 * j + 2]), __uint_as_float(r[rb][4 * j + 3]));
      v[2 * j] = fma2(a0, inv_ws, make_float2(b.x, b.y));
      v[2 * j + 1] = fma2(a1, inv_ws, make_float2(b.z, b.w));
    }
    if (kRelu && (KIND != 0 || DBG || NP == 3)) {
#pragma unroll
      for (int j = 0; j < NC / 2; ++j) { v[j].x = relu_nan(v[j].x); v[j].y = relu_nan(v[j].y); }
    }
    if (DBG) {
      if (l == args.dbg_layer && rd.valid) {
#pragma unroll
        for (int j = 0; j < NC / 2; ++j) {
          args.dbg[rd.m * 256 + c0 + 2 * j] = v[j].x;
          args.dbg[rd.m * 256 + c0 + 2 * j + 1] = v[j].y;
        }
      }
    }
    if (KIND == 1 || KIND == 4) {
#pragma unroll
      for (int j = 0; j < NC / 2; ++j) {
        alpha = fmaf(v[j].x, cst.w_alpha[c0 + 2 * j], alpha);
        alpha = fmaf(v[j].y, cst.w_alpha[c0 + 2 * j + 1], alpha);
      }
    }
    if (KIND == 3 || KIND == 5) {
#pragma unroll
      for (int j = 0; j < NC / 2; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float x = e ? v[j].y : v[j].x;
          const int c = c0 + 2 * j + e;
          if (KIND == 3) {
            hd[0] = fmaf(x, cst.w_rgb[0][c], hd[0]);
            hd[1] = fmaf(x, cst.w_rgb[1][c], hd[1]);
            hd[2] = fmaf(x, cst.w_rgb[2][c], hd[2]);
          }
          hd[3] = fmaf(x, cst.w_unc[c], hd[3]);
        }
      }
    }
    if (kStore) {
      // next layer's A operand, written IN PLACE over the accumulator columns this thread just read:
      // channels c0..c0+15 of this row -> 8 packed columns at c0 (k-step c0/16 of the next layer's TS MMAs)
      uint32_t pk[NC / 2];
      if (NP == 3) {
        uint32_t pl[NC / 2];
#pragma unroll
        for (int j = 0; j < NC / 2; ++j) split2(v[j].x, v[j].y, pk[j], pl[j]);
        tmem_st8(t_acc + (uint32_t)c0, pk);
        tmem_st8(t_acc + (uint32_t)(c0 + NC / 2), pl);
      } else {
#pragma unroll
        for (int j = 0; j < NC / 2; ++j) {
          pk[j] = (KIND == 0 && !DBG) ? pack2_relu<BF16>(v[j].x, v[j].y) : pack2<BF16>(v[j].x, v[j].y);
        }
        tmem_st8(t_acc + (uint32_t)c0, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a + 8 * kb);
      if (TR && tr && lane == 0) tr[1 + kb] = clk();
      if (TR && tr0 && lane == 0 && kb == 0) tr0[0] = clk();
    }
    if (kPrefetch) {
      if (kb == 1 && NCH > 2 && !half1_early) {
        mbar_wait(bar_acc_l + 8, acc_parity);
        tc_fence_after();
        tmem_ld16(t_acc + (uint32_t)(2 * 64 + hf * NC), r[0]);
      }
    } else if (kb + 1 < NCH) {
      if (kb + 1 == 2) {
        mbar_wait(bar_acc_l + 8, acc_parity);
        tc_fence_after();
      }
      tmem_ld16(t_acc + (uint32_t)((kb + 1) * 64 + hf * NC), r[0]);
    }
  }
  if (NCH == 2) mbar_wait(bar_acc_l + 8, acc_parity);   // keeps the phase of the (unused) second-half barrier in step
}

// SPLIT: hidden K-blocks kb >= SPLIT of a 256-wide layer are issued as two N = 128 halves (4 = none)
// DM: 0 = production, 1 = per-layer activation dumps + pipeline trace (fp32 ReLU path), 2 = pipeline trace only
// (production arithmetic and timing)
// NP: tensor-core passes per product.  1 = plain 16-bit operands (the bulk pass).  With fp16 operands split as
// x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 22 significant bits), 2 = A_hi * (W_hi + W_lo): weights at
// fp32-class precision, activations rounded to fp16; 3 = A_hi * W_hi + A_hi * W_lo + A_lo * W_hi: the whole
// product at fp32-class precision (the dropped lo * lo term is 2^-22 relative), accumulated in fp32 in tensor
// memory -- the "exact" path the planner uses to re-score near-tied candidates (plannerServer_*/core/
// interface.py:115-116 must pick the view the fp32 reference picks).  W_lo is a second weight image streamed
// through the same ring; A_lo lives in the 8 tensor-memory columns that the in-place packing leaves free.
// HEADS: dead-output elimination for score-only calls.  0 = all five outputs -> out[m * 5 + ..] (render());
// 1 = sigma only -> out[m] (the coarse pass of Controller.get_uncertainty feeds nothing but sample_pdf's
// weights: the pass stops after layer 7 + alpha head); 2 = beta only -> out[m] (nerf.py:307-309 reads
// raw[..., 4] of the fine pass and nothing else: the alpha and rgb heads are skipped); 3 = like 2 with feature_linear
// FOLDED into views_linears.0: the reference applies no activation between them (run_nerf_helpers.py:119-122:
// feature = feature_linear(h); h = relu(views_linears[0](cat[feature, dirs]))), so
// W_v[:, :256] (W_f h + b_f) + W_v[:, 256:] d + b_v = (W_v[:, :256] W_f) h + W_v[:, 256:] d + (W_v[:, :256] b_f + b_v):
// one 283 -> 128 layer whose weights are multiplied out on the host in fp64 (tc_pack_weights, fold = true).  The
// pass runs 9 GEMM layers instead of 10 (11 % fewer tensor cycles per fine evaluation) and skips one rounding of
// the activations to 16 bits; it is the same function of its inputs, equal to the unfolded pass at rounding level.
// 4 = all five outputs with the same folded view layer (render() when evpp_set_render_mode allows it).
template <bool BF16, bool RAYS, int C, int DM, int SPLIT, int NP, int HEADS>
__global__ void __launch_bounds__(kThreads, 1)
mlp_tc_kernel(const __grid_constant__ TcConsts cst, const __grid_constant__ TcArgs args) {
  static_assert(NP == 1 || (!BF16 && SPLIT == 4 && DM == 0), "split passes: fp16 operands, no half-split K-blocks");
  static_assert(HEADS == 0 || (RAYS && DM == 0), "score-only variants exist for ray batches only");
  static_assert(HEADS >= 0 && HEADS <= 4, "HEADS: 0 full, 1 sigma only, 2 beta only, 3 beta only + folded feature layer, 4 full + folded");
  using L_ = Lay<NP>;
  constexpr int kStages = L_::kStages;
  constexpr int NB = NP > 1 ? 2 : 1;                  // ring stages per K-block: W_hi (, W_lo)
  constexpr bool kFold = HEADS == 3 || HEADS == 4;     // feature_linear multiplied into the view layer on the host
  constexpr bool kAll = HEADS == 0 || HEADS == 4;      // all five outputs (render())
  constexpr bool kBeta = HEADS == 2 || HEADS == 3;     // beta only
  constexpr int kView = kFold ? 8 : 9;                 // index of the view layer
  constexpr int kLayersRun = HEADS == 1 ? 8 : kView + 1;
  constexpr bool DBG = DM == 1;
  constexpr bool TR = DM != 0;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t smem_base = smem_u32(smem_dyn);
  if (smem_base & 1023u) __trap();   // the 128-byte swizzle atoms need 1024-byte aligned blocks
  uint8_t* smem_gen = smem_dyn;
  Misc& misc = *reinterpret_cast<Misc*>(smem_gen + L_::kOffMisc);
  float* partial = reinterpret_cast<float*>(smem_gen + L_::kOffXD);   // [128][kColSplit-1][5] head partial sums; aliases XD,
                                                                       // which is idle between the view layer and the next tile
  float* bias_s = reinterpret_cast<float*>(smem_gen + L_::kOffBias);
  for (int i = threadIdx.x; i < kLayers * 256; i += kThreads) bias_s[i] = cst.bias[i >> 8][i & 255];
  const uint32_t sX = smem_base + L_::kOffX, sXD = smem_base + L_::kOffXD, sRing = smem_base + L_::kOffRing;
  constexpr uint32_t kXloDelta = L_::kOffXlo - L_::kOffX, kXDloDelta = L_::kOffXDlo - L_::kOffXD;
  const uint32_t bar_full = smem_u32(&misc.full[0]), bar_empty = smem_u32(&misc.empty[0]);
  const uint32_t bar_acc = smem_u32(&misc.acc_full[0]), bar_a = smem_u32(&misc.a_ready[0]);
  const uint32_t bar_x = smem_u32(&misc.x_ready), bar_xd = smem_u32(&misc.xd_ready);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = C > 1 ? cluster_ctarank() : 0u;
  const uint16_t mc_mask = (uint16_t)((1u << C) - 1u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, C);
    }
    for (int k = 0; k < 4; ++k) mbar_init(bar_acc + 8 * k, 1);
    for (int k = 0; k < 4; ++k) mbar_init(bar_a + 8 * k, kEpiWarps);
    mbar_init(bar_x, kEpiWarps);
    mbar_init(bar_xd, kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {   // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc.tmem_base)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = misc.tmem_base;

  const int ntile_iters = args.tiles_per_cta;
  const long long ntiles = args.tile_end;
  // Hybrid launch (launch_any): a dependent launch that fills the SMs this grid's clusters cannot tile may start as
  // soon as every CTA of this grid is resident.
  if (args.pdl == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef EVPP_TC_DIAG
  if (threadIdx.x == 0) {
    unsigned long long t; unsigned sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    g_diag[C == 4 ? 0 : 1][blockIdx.x][0] = t;
    g_diag[C == 4 ? 0 : 1][blockIdx.x][2] = sm;
  }
#endif

  if (warp == kProducerWarp) {
    // ================= weight producer =================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < ntile_iters; ++it) {
        const uint8_t* src = args.image;
        for (int l = 0; l < kLayersRun; ++l) {
          const int nblk = (layer_has_x(l, kView) + layer_h_blocks(l)) * NB;   // NP > 1: every K-block is a (W_hi, W_lo) pair
          const uint32_t bytes = (uint32_t)layer_n(l, kView) * 128u;
          for (int b = 0; b < nblk; ++b) {
            if (l == 4 && b == 0) EVPP_TRACE(117);
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            if (l == 4 && b == 0) EVPP_TRACE(118);
            mbar_expect_tx(bar_full + 8 * stage, bytes);
            if (C == 1) {
              bulk_g2s(sRing + stage * kStageBytes, src, bytes, bar_full + 8 * stage);
            } else {
              const uint32_t part = bytes / C;
              bulk_g2s_mcast(sRing + stage * kStageBytes + cta_rank * part, src + cta_rank * part, part,
                             bar_full + 8 * stage, mc_mask);
            }
            src += bytes;
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    if (TR && lane == 31 && args.trace && blockIdx.x == 0) {
      // diagnostics: when each accumulator half of each layer really becomes final (tile 2 of CTA 0)
      uint32_t ph[2] = {0u, 0u};
      for (int it = 0; it < ntile_iters; ++it)
        for (int l = 0; l < kLayers; ++l) {
          mbar_wait(bar_acc + 16 * (l & 1), ph[l & 1]);
          EVPP_TRACE(70 + 2 * l);
          mbar_wait(bar_acc + 16 * (l & 1) + 8, ph[l & 1]);
          EVPP_TRACE(71 + 2 * l);
          ph[l & 1] ^= 1u;
        }
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, a_phase = 0, x_phase = 0, xd_phase = 0;
      bool pre_full = false;   // the next weight stage's arrival was already awaited (static after unrolling)
      for (int it = 0; it < ntile_iters; ++it) {
        for (int l = 0; l < kLayersRun; ++l) {
          // Accumulator region of layer l.  With an ODD number of layers per tile (the folded pass runs nine) the first
          // and the last layer of a tile would share a region, and the next tile's layer 0 -- which depends on nothing
          // but its staged embedding -- could overwrite the accumulator the epilogue is still draining: the parity
          // alternates from tile to tile there, so that a tile's first layer always lands in the other region.
          const uint32_t reg = (uint32_t)((l + ((kLayersRun & 1) ? it : 0)) & 1);
          const uint32_t d_tmem = tmem + reg * 256u;
          const uint32_t a_tmem = tmem + (reg ^ 1u) * 256u;
          const uint32_t idesc = make_idesc(BF16, layer_n(l, kView));
          uint32_t accumulate = 0;
          EVPP_TRACE(l * 4 + 0);
          if constexpr (NP > 1) {
            // ---- split passes: per K step A_hi * W_hi (+ A_hi * W_lo (+ A_lo * W_hi)); K-block = 2 ring stages ----
            const int nblk = layer_has_x(l, kView) + layer_h_blocks(l);
            for (int b = 0; b < nblk; ++b) {
              const bool xblk = layer_has_x(l, kView) && b == 0;
              const int kb = b - layer_has_x(l, kView);
              uint32_t s_hi = stage, p_hi = phase;
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
              uint32_t s_lo = stage, p_lo = phase;
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
              mbar_wait(bar_full + 8 * s_hi, p_hi);   // weights first: prefetched long ago
              mbar_wait(bar_full + 8 * s_lo, p_lo);
              const uint64_t b_hi = make_desc(sRing + s_hi * kStageBytes), b_lo = make_desc(sRing + s_lo * kStageBytes);
              if (xblk) {
                if (l == 0) { mbar_wait(bar_x, x_phase); x_phase ^= 1u; }
                if (l == kView) { mbar_wait(bar_xd, xd_phase); xd_phase ^= 1u; }
                tc_fence_after();
                const int ksteps = l == kView ? 2 : 4;
                for (int k = 0; k < ksteps; ++k) {
                  const uint64_t a_hi = l == kView ? make_desc_sw64(sXD + k * 32) : make_desc(sX + k * 32);
                  const uint64_t a_lo = l == kView ? make_desc_sw64(sXD + kXDloDelta + k * 32) : make_desc(sX + kXloDelta + k * 32);
                  umma_f16(d_tmem, a_hi, b_hi + (uint64_t)(k * 2), idesc, accumulate);
                  umma_f16(d_tmem, a_hi, b_lo + (uint64_t)(k * 2), idesc, 1u);
                  if (NP == 3) umma_f16(d_tmem, a_lo, b_hi + (uint64_t)(k * 2), idesc, 1u);
                  accumulate = 1;
                }
              } else {
                mbar_wait(bar_a + 8 * kb, a_phase);     // then the activations of this K-block
                tc_fence_after();
                const uint32_t a_addr0 = a_tmem + (uint32_t)(kb * 64);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t a_hi = a_addr0 + (uint32_t)(k * 16);
                  umma_f16_ts(d_tmem, a_hi, b_hi + (uint64_t)(k * 2), idesc, accumulate);
                  umma_f16_ts(d_tmem, a_hi, b_lo + (uint64_t)(k * 2), idesc, 1u);
                  if (NP == 3) umma_f16_ts(d_tmem, a_hi + 8u, b_hi + (uint64_t)(k * 2), idesc, 1u);
                  accumulate = 1;
                }
              }
              if (C == 1) { umma_commit(bar_empty + 8 * s_hi); umma_commit(bar_empty + 8 * s_lo); }
              else { umma_commit_mcast(bar_empty + 8 * s_hi, mc_mask); umma_commit_mcast(bar_empty + 8 * s_lo, mc_mask); }
            }
            umma_commit(bar_acc + 16 * reg);
            if (l >= 1) a_phase ^= 1u;
            umma_commit(bar_acc + 16 * reg + 8);
            continue;
          }
          if (layer_has_x(l, kView)) {
            // layers 0 and 5 read the position embedding (X), the view layer the direction embedding (XD)
            if (l == 0) { mbar_wait(bar_x, x_phase); x_phase ^= 1u; }
            if (l == kView) { mbar_wait(bar_xd, xd_phase); xd_phase ^= 1u; }
            if (!pre_full) mbar_wait(bar_full + 8 * stage, phase);
            pre_full = false;
            tc_fence_after();
            const int ksteps = l == kView ? 2 : 4;
            const uint32_t xa = l == kView ? sXD : sX;
            for (int k = 0; k < ksteps; ++k) {
              umma_f16(d_tmem, l == kView ? make_desc_sw64(xa + k * 32) : make_desc(xa + k * 32),
                       make_desc(sRing + stage * kStageBytes + k * 32), idesc, accumulate);
              accumulate = 1;
            }
            if (C == 1) umma_commit(bar_empty + 8 * stage); else umma_commit_mcast(bar_empty + 8 * stage, mc_mask);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          const int nfull = (layer_n(l, kView) == 256 && layer_h_blocks(l) == 4) ? SPLIT : layer_h_blocks(l);
          for (int kb = 0; kb < nfull; ++kb) {
            if (!pre_full) mbar_wait(bar_full + 8 * stage, phase);   // weights first: prefetched long ago, off the critical path
            pre_full = false;
            if (kb == 0 && l == 4) EVPP_TRACE(116);
            const uint64_t b_desc0 = make_desc(sRing + stage * kStageBytes);
            const uint32_t a_addr0 = a_tmem + (uint32_t)(kb * 64);
            mbar_wait(bar_a + 8 * kb, a_phase);         // then the activations of this K-block
            tc_fence_after();
            if (kb == 0) EVPP_TRACE(l * 4 + 1);
            if (kb == 3) EVPP_TRACE(l * 4 + 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // A = the previous layer's activations, packed in place in the other accumulator region
              umma_f16_ts(d_tmem, a_addr0 + (uint32_t)(k * 16), b_desc0 + (uint64_t)(k * 2), idesc, accumulate);   // +32 B per K step
              accumulate = 1;
            }
            if (C == 1) umma_commit(bar_empty + 8 * stage); else umma_commit_mcast(bar_empty + 8 * stage, mc_mask);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          if (nfull < layer_h_blocks(l)) {
            // The remaining K-blocks as two N = 128 halves (weight rows 0..127 -> accumulator columns 0..127 first):
            // the first half's columns are final, and its drain starts, while the second half is still computing.
            const uint32_t idesc_h = make_idesc(BF16, 128);
            uint32_t s2 = stage, p2 = phase;
            for (int kb = nfull; kb < 4; ++kb) {
              mbar_wait(bar_full + 8 * s2, p2);
              const uint64_t b_desc0 = make_desc(sRing + s2 * kStageBytes);
              const uint32_t a_addr0 = a_tmem + (uint32_t)(kb * 64);
              mbar_wait(bar_a + 8 * kb, a_phase);
              tc_fence_after();
              if (kb == 3) EVPP_TRACE(l * 4 + 2);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(d_tmem, a_addr0 + (uint32_t)(k * 16), b_desc0 + (uint64_t)(k * 2), idesc_h, 1u);
              if (++s2 == kStages) { s2 = 0; p2 ^= 1u; }
            }
            umma_commit(bar_acc + 16 * reg);
            // The issuer crawls while the epilogue warps are busy (it loses most issue slots to them), and after
            // the second half it is on the critical path to the next layer: look at the next layer's first
            // weight stage now, while the tensor pipe still has queued work.
            if (l + 1 < kLayersRun || it + 1 < ntile_iters) {
              mbar_wait(bar_full + 8 * s2, p2);
              pre_full = true;
            }
            for (int kb = nfull; kb < 4; ++kb) {
              const uint64_t b_desc0 = make_desc(sRing + stage * kStageBytes + 16384);
              const uint32_t a_addr0 = a_tmem + (uint32_t)(kb * 64);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(d_tmem + 128u, a_addr0 + (uint32_t)(k * 16), b_desc0 + (uint64_t)(k * 2), idesc_h, 1u);
              if (C == 1) umma_commit(bar_empty + 8 * stage); else umma_commit_mcast(bar_empty + 8 * stage, mc_mask);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
          } else {
            umma_commit(bar_acc + 16 * reg);
          }
          if (l >= 1) a_phase ^= 1u;
          umma_commit(bar_acc + 16 * reg + 8);
          EVPP_TRACE(l * 4 + 3);
        }
      }
    }
  } else {
    // ================= encode + epilogue warps =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int hf = warp >> 2;         // which kColsPerWarp-column slice of every 64-column chunk
    const int row = q * 32 + lane;
    const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t acc_phase = 0u;   // bit r = phase of accumulator region r (NOT an array: l & 1 is a runtime index, and a
                               // dynamically indexed array would live in local memory -- 320 LDL / STL per tile)
    RowData cur, nxt;
    load_row<RAYS>(args, args.tile0 + (long long)blockIdx.x, ntiles, row, cur);
    if (hf == 0) {
      encode_store<BF16, kLpos, false, NP == 3>(cur.p, sX + row_off, kXloDelta, row, 8);
      fence_proxy_async();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_x);
    for (int it = 0; it < ntile_iters; ++it) {
      float alpha = 0.f;
      float hd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int l = 0; l < kLayersRun; ++l) {
        const int ab = (l + ((kLayersRun & 1) ? it : 0)) & 1;      // accumulator region, see the MMA issuer
        const uint32_t bar_l = bar_acc + 16 * ab, par_l = (acc_phase >> ab) & 1u;
        acc_phase ^= 1u << ab;
        if (warp == 0 && lane == 0) EVPP_TRACE(40 + l * 3 + 0);   // (before the accumulator wait)
        const uint32_t t_acc = t_lane + (uint32_t)(ab * 256);
        long long* tr = nullptr;
        if (TR && args.trace && blockIdx.x == 0 && it == 2 && l == 3 && (warp == 0 || warp == kEpiWarps - 1))
          tr = args.trace + (warp == 0 ? 90 : 95);
        long long* tr0 = (TR && args.trace && blockIdx.x == 0 && it == 2 && l == 3) ? args.trace + 100 + warp : nullptr;
#define EVPP_DRAIN(KIND_, HFC_) \
  drain_layer<BF16, KIND_, DBG, TR, NP, HFC_>(cst, args, bias_s + l * 256, l, t_acc, hf, lane, bar_a, bar_l, par_l, cur, alpha, hd, tr, tr0)
#define EVPP_DRAIN_HF(KIND_) \
  do { if (hf == 0) EVPP_DRAIN(KIND_, 0); else if (hf == 1) EVPP_DRAIN(KIND_, 1); else if (hf == 2) EVPP_DRAIN(KIND_, 2); else EVPP_DRAIN(KIND_, 3); } while (0)
        if (HEADS != 1 && l == kView) {
          if constexpr (kBeta) EVPP_DRAIN_HF(5); else if constexpr (kAll) EVPP_DRAIN_HF(3);
        } else if (HEADS != 1 && !kFold && l == 8) {
          if constexpr (HEADS != 1 && !kFold) EVPP_DRAIN(2, -1);
        } else if (l == 7) {
          if constexpr (HEADS == 1) EVPP_DRAIN_HF(4); else if constexpr (kBeta) EVPP_DRAIN(0, -1); else EVPP_DRAIN_HF(1);
        } else {
          EVPP_DRAIN(0, -1);
        }
#undef EVPP_DRAIN_HF
#undef EVPP_DRAIN
        if (warp == 0 && lane == 0) EVPP_TRACE(40 + l * 3 + 1);
        // The epilogue warps now idle until the next layer's MMAs finish: use the window to stage operands.
        if (l == 0 && HEADS != 1) {
          // XD <- direction embedding of this tile (the previous tile's view-layer MMAs, its last readers,
          // completed before that tile's final accumulator wait)
          if (hf == 1) {
            encode_store<BF16, kLdir, true, NP == 3>(cur.dv, sXD + (uint32_t)(row * 64), kXDloDelta, row, 4);
            fence_proxy_async();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_xd);
        } else if (l == 5) {
          // X <- position embedding of the NEXT tile (layer 5's MMAs were the last readers of this tile's)
          load_row<RAYS>(args, args.tile0 + (long long)blockIdx.x + (long long)(it + 1) * gridDim.x,
                         it + 1 < ntile_iters ? ntiles : 0, row, nxt);
          if (hf == 0) {
            encode_store<BF16, kLpos, false, NP == 3>(nxt.p, sX + row_off, kXloDelta, row, 8);
            fence_proxy_async();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_x);
        }
        if (warp == 0 && lane == 0) EVPP_TRACE(40 + l * 3 + 2);
      }
      // ---- heads: combine the column slices, add biases, exp on the uncertainty ----
      if (hf != 0) {
        float* pr = partial + (row * (kColSplit - 1) + hf - 1) * 5;
        if (kAll) { pr[0] = hd[0]; pr[1] = hd[1]; pr[2] = hd[2]; }
        if (kAll || HEADS == 1) pr[3] = alpha;
        if (HEADS != 1) pr[4] = hd[3];
      }
      epi_bar_sync();
      if (hf == 0 && cur.valid) {
        float t[5] = {hd[0], hd[1], hd[2], alpha, hd[3]};
#pragma unroll
        for (int k = 0; k < kColSplit - 1; ++k) {
          const float* pr = partial + (row * (kColSplit - 1) + k) * 5;
#pragma unroll
          for (int e = 0; e < 5; ++e)
            if (kAll || (HEADS == 1 && e == 3) || (kBeta && e == 4)) t[e] += pr[e];
        }
        if (kAll) {
          float* o = args.out + cur.m * 5;
          o[0] = t[0] + cst.head_b[1];
          o[1] = t[1] + cst.head_b[2];
          o[2] = t[2] + cst.head_b[3];
          o[3] = t[3] + cst.head_b[0];
          o[4] = expf(t[4] + cst.head_b[4]);   // run_nerf_helpers.py:128-129
        } else if (HEADS == 1) {
          args.out[cur.m] = t[3] + cst.head_b[0];
        } else {
          args.out[cur.m] = expf(t[4] + cst.head_b[4]);
        }
      }
      epi_bar_sync();   // partial[] (XD) is rewritten in the next tile's layer-0 window
      cur = nxt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();   // peers may still multicast into / commit to this CTA until here
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
#ifdef EVPP_TC_DIAG
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_diag[C == 4 ? 0 : 1][blockIdx.x][1] = t;
  }
#endif
  // A dependent (filler) launch must not complete before the grid it was launched behind: the next kernel in the
  // stream reads the rows of both.
  if (args.pdl == 2) asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <bool BF16, bool RAYS, int C, int DM, int SPLIT, int NP, int HEADS>
cudaError_t launch_one(const TcConsts& cst, const TcArgs& a, int grid, cudaStream_t st, bool dependent = false) {
  auto kern = mlp_tc_kernel<BF16, RAYS, C, DM, SPLIT, NP, HEADS>;
  constexpr int smem = Lay<NP>::kSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = dependent ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, cst, a);
}

// Launch options, read from the environment once (thread-safe: initialised as a function-local static).
struct TcEnv {
  int cluster = 2;             // CTAs sharing one weight stream by bulk-copy multicast (1, 2, 4, or 42 = hybrid 4 + 2); EVPP_TC_CLUSTER
  int split = 2;               // hidden K-blocks >= split run as two N = 128 halves (2, or 4 = off); EVPP_TC_SPLIT
  float hybrid_ratio = 1.066f; // per-SM speed of a 4-CTA cluster relative to a 2-CTA cluster; EVPP_TC_HYBRID_R
  TcEnv() {
    if (const char* e = getenv("EVPP_TC_CLUSTER")) {
      const int c = atoi(e);
      if (c == 1 || c == 2 || c == 4 || c == 42) cluster = c;
    }
    if (const char* e = getenv("EVPP_TC_SPLIT")) {
      const int c = atoi(e);
      if (c == 2 || c == 4) split = c;
    }
    if (const char* e = getenv("EVPP_TC_HYBRID_R")) {
      const float r = (float)atof(e);
      if (r > 0.5f && r < 2.f) hybrid_ratio = r;
    }
  }
};
const TcEnv& tc_env() {
  static const TcEnv env;
  return env;
}

// Split-precision passes (fp16x2 / fp16x3) and the score-only variants are built for clusters of 1 and 2 CTAs and
// one K-block schedule each (the experiment knobs EVPP_TC_CLUSTER=4/42 and EVPP_TC_SPLIT apply to the plain pass).
template <bool RAYS, int NP, int HEADS>
cudaError_t launch_np(const TcWeights& w, const TcArgs& a, int C, long long grid, cudaStream_t st) {
  if (C >= 2) return launch_one<false, RAYS, 2, 0, 4, NP, HEADS>(*w.consts, a, (int)grid, st, false);
  return launch_one<false, RAYS, 1, 0, 4, NP, HEADS>(*w.consts, a, (int)grid, st, false);
}

template <bool RAYS, int SPLIT>
cudaError_t launch_split(const TcWeights& w, const TcArgs& a, int C, long long grid, cudaStream_t st,
                         bool dependent = false) {
  const bool bf = w.dtype == EVPP_PREC_BF16;
  cudaError_t e;
#define EVPP_TC_DISPATCH(CC, DD)                                                                    \
  e = bf ? launch_one<true, RAYS, CC, DD, SPLIT, 1, 0>(*w.consts, a, (int)grid, st, dependent)      \
         : launch_one<false, RAYS, CC, DD, SPLIT, 1, 0>(*w.consts, a, (int)grid, st, dependent)
  if (a.dbg) {
    EVPP_TC_DISPATCH(1, RAYS ? 0 : 1);
  } else if (a.trace) {          // production kernel (cluster, packed epilogue) with the clock stamps compiled in
    if (C == 2) { EVPP_TC_DISPATCH(2, RAYS ? 0 : 2); } else { EVPP_TC_DISPATCH(1, RAYS ? 0 : 2); }
  } else if (C == 4) { EVPP_TC_DISPATCH(4, 0); }
  else if (C == 2) { EVPP_TC_DISPATCH(2, 0); }
  else { EVPP_TC_DISPATCH(1, 0); }
#undef EVPP_TC_DISPATCH
  return e;
}

// score-only variants of the plain pass (HEADS 1 / 2), ray batches only
template <int HEADS>
cudaError_t launch_heads(const TcWeights& w, const TcArgs& a, int C, long long grid, cudaStream_t st) {
  const bool bf = w.dtype == EVPP_PREC_BF16;
  if (C >= 2)
    return bf ? launch_one<true, true, 2, 0, 2, 1, HEADS>(*w.consts, a, (int)grid, st, false)
              : launch_one<false, true, 2, 0, 2, 1, HEADS>(*w.consts, a, (int)grid, st, false);
  return bf ? launch_one<true, true, 1, 0, 2, 1, HEADS>(*w.consts, a, (int)grid, st, false)
            : launch_one<false, true, 1, 0, 2, 1, HEADS>(*w.consts, a, (int)grid, st, false);
}

// How many 4-CTA clusters of this kernel are co-resident (they do not tile every GPC: 33 on a 148-SM B200).
int max_clusters_of_4(int num_sms) {
  static const int max_clusters = [num_sms] {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Lay<1>::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    auto kern = mlp_tc_kernel<false, true, 4, 0, 2, 1, 0>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<1>::kSmemBytes);
    int mc = 0;
    if (cudaOccupancyMaxActiveClusters(&mc, kern, &cfg) != cudaSuccess || mc <= 0) mc = num_sms / 4;
    if (getenv("EVPP_TC_VERBOSE")) fprintf(stderr, "evpp: %d co-resident 4-CTA clusters\n", mc);
    return mc;
  }();
  return max_clusters;
}

template <bool RAYS>
int launch_any(const TcWeights& w, TcArgs a, int num_sms, cudaStream_t st, int heads = 0) {
  if (a.M <= 0) return 0;
  const TcEnv& env = tc_env();
  const int g_cluster = env.cluster, g_split = env.split;
  const float g_hybrid_ratio = env.hybrid_ratio;
  const long long ntiles = (a.M + kTileM - 1) / kTileM;
  a.image = reinterpret_cast<const uint8_t*>(w.image);
  a.dbg = w.dbg;
  a.dbg_layer = w.dbg_layer;
  a.trace = w.trace;
  a.tile0 = 0;
  a.tile_end = ntiles;
  a.pdl = 0;
  if ((a.dbg || a.trace) && RAYS) return -1;
  const int np = w.dtype == EVPP_PREC_FP16X3 ? 3 : w.dtype == EVPP_PREC_FP16X2 ? 2 : 1;
  if ((np > 1 || heads != 0) && (a.dbg || a.trace)) return -1;
  if (heads != 0 && !RAYS) return -1;
  if ((heads == 3 || heads == 4) != (w.folded != 0)) return -1;      // the folded passes need the folded weight image, and only they may use it

  // Hybrid 4 + 2: 4-CTA clusters share one weight stream four ways (L2 reads quartered) but leave 16 of the 148 SMs
  // idle; a second launch of 2-CTA clusters, released by griddepcontrol.launch_dependents as soon as every 4-CTA
  // cluster is resident, takes those SMs and a proportional slice of the tiles.
  if (g_cluster == 42 && np == 1 && heads == 0 && !a.dbg && !a.trace && ntiles >= 32LL * num_sms) {
    const long long grid_a = (long long)max_clusters_of_4(num_sms) * 4;
    const long long grid_b = (num_sms - grid_a) / 2 * 2;
    if (grid_a > 0 && grid_b > 0) {
      const double share_a = grid_a * (double)g_hybrid_ratio / (grid_a * (double)g_hybrid_ratio + grid_b);
      long long tiles_a = (long long)(ntiles * share_a) / grid_a * grid_a;
      TcArgs aa = a, ab = a;
      aa.tile_end = tiles_a;
      aa.tiles_per_cta = (int)(tiles_a / grid_a);
      ab.tile0 = tiles_a;
      aa.pdl = 1;
      ab.pdl = 2;
      ab.tiles_per_cta = (int)((ntiles - tiles_a + grid_b - 1) / grid_b);
      cudaError_t e = g_split == 2 ? launch_split<RAYS, 2>(w, aa, 4, grid_a, st) : launch_split<RAYS, 4>(w, aa, 4, grid_a, st);
      if (e != cudaSuccess) return -1;
      e = g_split == 2 ? launch_split<RAYS, 2>(w, ab, 2, grid_b, st, true) : launch_split<RAYS, 4>(w, ab, 2, grid_b, st, true);
#ifdef EVPP_TC_DIAG
      {
        static int shown = 0;
        if (shown < 4 && ntiles > 100000) {
          ++shown;
          cudaStreamSynchronize(st);
          static unsigned long long hd[2][160][3];
          cudaMemcpyFromSymbol(hd, g_diag, sizeof(hd));
          unsigned long long t0 = ~0ull;
          for (int k = 0; k < 2; ++k)
            for (int b = 0; b < (k ? grid_b : grid_a); ++b) t0 = hd[k][b][0] < t0 ? hd[k][b][0] : t0;
          for (int k = 0; k < 2; ++k) {
            unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0;
            for (int b = 0; b < (k ? grid_b : grid_a); ++b) {
              s0 = hd[k][b][0] < s0 ? hd[k][b][0] : s0; s1 = hd[k][b][0] > s1 ? hd[k][b][0] : s1;
              e0 = hd[k][b][1] < e0 ? hd[k][b][1] : e0; e1 = hd[k][b][1] > e1 ? hd[k][b][1] : e1;
            }
            fprintf(stderr, "diag %s: start %.3f..%.3f ms, end %.3f..%.3f ms, tiles/cta %d; smids:", k ? "B(2)" : "A(4)",
                    (s0 - t0) * 1e-6, (s1 - t0) * 1e-6, (e0 - t0) * 1e-6, (e1 - t0) * 1e-6, k ? ab.tiles_per_cta : aa.tiles_per_cta);
            if (k) for (int b = 0; b < grid_b; ++b) fprintf(stderr, " %llu", hd[k][b][2]);
            fprintf(stderr, "\n");
          }
        }
      }
#endif
      return e == cudaSuccess ? 2 : -1;
    }
  }

  int C = g_cluster == 42 ? 2 : g_cluster;
  if ((np > 1 || heads != 0) && C > 2) C = 2;
  long long grid = ntiles < num_sms ? ntiles : num_sms;
  grid = (grid + C - 1) / C * C;
  if (grid > num_sms) grid = num_sms / C * C;
  if (C == 4) {
    // 4-CTA clusters do not tile every GPC: keep the persistent grid to what is co-resident
    const long long cap = (long long)max_clusters_of_4(num_sms) * 4;
    if (grid > cap) grid = cap;
  }
  a.tiles_per_cta = (int)((ntiles + grid - 1) / grid);
  cudaError_t e;
  if (np > 1) {
    if constexpr (RAYS) {
      if (np == 3) e = heads == 1 ? launch_np<true, 3, 1>(w, a, C, grid, st) : heads == 2 ? launch_np<true, 3, 2>(w, a, C, grid, st)
                     : heads == 3 ? launch_np<true, 3, 3>(w, a, C, grid, st) : heads == 4 ? launch_np<true, 3, 4>(w, a, C, grid, st)
                     : launch_np<true, 3, 0>(w, a, C, grid, st);
      else e = heads == 1 ? launch_np<true, 2, 1>(w, a, C, grid, st) : heads == 2 ? launch_np<true, 2, 2>(w, a, C, grid, st)
               : heads == 3 ? launch_np<true, 2, 3>(w, a, C, grid, st) : heads == 4 ? launch_np<true, 2, 4>(w, a, C, grid, st)
               : launch_np<true, 2, 0>(w, a, C, grid, st);
    } else {
      e = np == 3 ? launch_np<false, 3, 0>(w, a, C, grid, st) : launch_np<false, 2, 0>(w, a, C, grid, st);
    }
  } else if (heads != 0) {
    if constexpr (RAYS) e = heads == 1 ? launch_heads<1>(w, a, C, grid, st) : heads == 2 ? launch_heads<2>(w, a, C, grid, st)
                          : heads == 3 ? launch_heads<3>(w, a, C, grid, st) : launch_heads<4>(w, a, C, grid, st);
    else e = cudaErrorInvalidValue;
  } else {
    switch (g_split) {
      case 2: e = launch_split<RAYS, 2>(w, a, C, grid, st); break;
      default: e = launch_split<RAYS, 4>(w, a, C, grid, st); break;
    }
  }
#ifdef EVPP_TC_DIAG
  {
    static int shown = 0;
    if (shown < 6 && ntiles > 50000 && e == cudaSuccess) {
      ++shown;
      cudaStreamSynchronize(st);
      static unsigned long long hd[2][160][3];
      cudaMemcpyFromSymbol(hd, g_diag, sizeof(hd));
      const int k = C == 4 ? 0 : 1;
      unsigned long long s0 = ~0ull, e0 = ~0ull, e1 = 0;
      for (int b = 0; b < grid; ++b) {
        s0 = hd[k][b][0] < s0 ? hd[k][b][0] : s0;
        e0 = hd[k][b][1] < e0 ? hd[k][b][1] : e0; e1 = hd[k][b][1] > e1 ? hd[k][b][1] : e1;
      }
      fprintf(stderr, "diag C=%d: end %.3f..%.3f ms, tiles/cta %d; (smid:end)", C, (e0 - s0) * 1e-6, (e1 - s0) * 1e-6, a.tiles_per_cta);
      for (int b = 0; b < grid; b += C) fprintf(stderr, " %llu:%.2f", hd[k][b][2], (hd[k][b][1] - s0) * 1e-6);
      fprintf(stderr, "\n");
    }
  }
#endif
  return e == cudaSuccess ? 1 : -1;
}

template <typename T>
T to_op(float v);
template <>
__half to_op<__half>(float v) { return __float2half_rn(v); }
template <>
__nv_bfloat16 to_op<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Appends one pre-swizzled K-block: n_rows x 64 channels; channel k of row n = W[n*in_dim + col0 + k]
// for k < count, else 0.  Layout = what make_desc()/SWIZZLE_128B expects (see sw128()).
template <typename T>
void put_block(std::vector<uint8_t>& img, const std::vector<float>& W, int n_rows, int in_dim, int col0, int count,
               float scale = 1.f) {
  const size_t base = img.size();
  img.resize(base + (size_t)n_rows * 128, 0);
  for (int n = 0; n < n_rows; ++n)
    for (int k = 0; k < 64; ++k) {
      const float v = k < count ? W[(size_t)n * in_dim + col0 + k] * scale : 0.f;
      const T hv = to_op<T>(v);
      const size_t off = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
      memcpy(&img[base + off], &hv, 2);
    }
}

// One K-block of a weight image: rows [0, n_rows) of state_dict tensor `tensor` ([n_rows][in_dim]), input columns
// [col0, col0 + count), belonging to GEMM layer `layer` (its power-of-two scale applies).
struct Blk { int tensor, n_rows, in_dim, col0, count, layer; };

// Consumption order of the kernel.  fold = false: the reference's ten GEMM layers (pts_linears 0..7, feature_linear,
// views_linears.0).  fold = true: nine layers, feature_linear multiplied into views_linears.0 (tensor 16 then holds
// [W_v[:, :256] W_f | W_v[:, 256:]], see fold_feature_layer) as layer 8.
std::vector<Blk> image_blocks(bool fold) {
  std::vector<Blk> v;
  v.push_back({0, 256, 63, 0, 63, 0});                                          // layer 0 : X
  for (int l = 1; l <= 4; ++l)
    for (int kb = 0; kb < 4; ++kb) v.push_back({2 * l, 256, 256, kb * 64, 64, l});
  v.push_back({10, 256, 319, 0, 63, 5});                                        // layer 5 : X part of cat[pts, h]
  for (int kb = 0; kb < 4; ++kb) v.push_back({10, 256, 319, 63 + kb * 64, 64, 5});
  for (int l = 6; l <= 7; ++l)
    for (int kb = 0; kb < 4; ++kb) v.push_back({2 * l, 256, 256, kb * 64, 64, l});
  const int lv = fold ? 8 : 9;
  if (!fold)
    for (int kb = 0; kb < 4; ++kb) v.push_back({18, 256, 256, kb * 64, 64, 8});   // feature_linear
  v.push_back({16, 128, 283, 256, 27, lv});                                     // views : direction part
  for (int kb = 0; kb < 4; ++kb) v.push_back({16, 128, 283, kb * 64, 64, lv});  // views : feature part
  return v;
}

// state_dict tensor of GEMM layer l (0..7 pts_linears, then feature_linear (8) and views_linears.0 (9), or the folded
// view layer as 8)
int layer_tensor(int l, bool fold) { return l < 8 ? 2 * l : (l == 8 && !fold) ? 18 : 16; }
int n_layers(bool fold) { return fold ? 9 : 10; }

// Power-of-two scale that brings the largest |w| of a layer into [2^13, 2^14): a trained layer with tiny weights
// does not flush to zero in fp16 (subnormals start at 6e-5), a layer with weights beyond 65504 does not overflow,
// the lo halves of the split weights sit in fp16's normal range, and the epilogue undoes the scale exactly
// (TcConsts::inv_wscale, folded into the bias FMA).
float layer_wscale(const std::vector<float>& W) {
  float m = 0.f;
  for (float v : W) m = std::fmax(m, std::fabs(v));
  if (!(m > 0.f)) return 1.f;
  int e = 0;
  std::frexp(m, &e);                       // m = f * 2^e, f in [0.5, 1)
  return std::ldexp(1.f, 14 - e);          // m * scale in [2^13, 2^14)
}

// feature_linear folded into views_linears.0 (exact algebra, products accumulated in double):
//   W'[o][k] = sum_j W_v[o][j] W_f[j][k]   (k < 256),   W'[o][256 + k] = W_v[o][256 + k],   b'[o] = sum_j W_v[o][j] b_f[j] + b_v[o]
void fold_feature_layer(std::vector<std::vector<float>>& t) {
  const std::vector<float>&Wv = t[16], &bv = t[17], &Wf = t[18], &bf = t[19];
  std::vector<float> W2(Wv.size()), b2(bv.size());
  for (int o = 0; o < 128; ++o) {
    std::vector<double> acc(256, 0.0);
    double bacc = (double)bv[o];
    for (int j = 0; j < 256; ++j) {
      const double w = (double)Wv[(size_t)o * 283 + j];
      bacc += w * (double)bf[j];
      const float* row = &Wf[(size_t)j * 256];
      for (int k = 0; k < 256; ++k) acc[k] += w * (double)row[k];
    }
    for (int k = 0; k < 256; ++k) W2[(size_t)o * 283 + k] = (float)acc[k];
    for (int k = 256; k < 283; ++k) W2[(size_t)o * 283 + k] = Wv[(size_t)o * 283 + k];
    b2[o] = (float)bacc;
  }
  t[16] = W2;
  t[17] = b2;
}

template <typename T>
void pack_image(const std::vector<std::vector<float>>& t, bool fold, std::vector<uint8_t>& img) {
  img.clear();
  float sc[kLayers];
  for (int l = 0; l < n_layers(fold); ++l) sc[l] = layer_wscale(t[layer_tensor(l, fold)]);
  for (const Blk& b : image_blocks(fold)) put_block<T>(img, t[b.tensor], b.n_rows, b.in_dim, b.col0, b.count, sc[b.layer]);
}

// Split image: every K-block as a (hi, lo) pair, hi = fp16(w * s), lo = fp16(w * s - hi).
void pack_image_split(const std::vector<std::vector<float>>& t, bool fold, std::vector<uint8_t>& img) {
  std::vector<std::vector<float>> hi(t.size()), lo(t.size());
  for (int l = 0; l < n_layers(fold); ++l) {
    const int ti = layer_tensor(l, fold);
    const float s = layer_wscale(t[ti]);
    hi[ti].resize(t[ti].size());
    lo[ti].resize(t[ti].size());
    for (size_t i = 0; i < t[ti].size(); ++i) {
      const float w = t[ti][i] * s;
      const float h = __half2float(__float2half_rn(w));
      hi[ti][i] = h;
      lo[ti][i] = w - h;
    }
  }
  img.clear();
  for (const Blk& b : image_blocks(fold)) {
    put_block<__half>(img, hi[b.tensor], b.n_rows, b.in_dim, b.col0, b.count);
    put_block<__half>(img, lo[b.tensor], b.n_rows, b.in_dim, b.col0, b.count);
  }
}

}  // namespace

size_t tc_weight_image_bytes() { return kImageBytes; }

void tc_pack_weights(const std::vector<std::vector<float>>& t_in, int dtype, bool fold, std::vector<uint8_t>& image, TcConsts& c) {
  std::vector<std::vector<float>> t = t_in;
  if (fold) fold_feature_layer(t);
  if (dtype == EVPP_PREC_BF16) pack_image<__nv_bfloat16>(t, fold, image);
  else if (dtype == EVPP_PREC_FP16X2 || dtype == EVPP_PREC_FP16X3) pack_image_split(t, fold, image);
  else pack_image<__half>(t, fold, image);
  memset(&c, 0, sizeof(c));
  for (int l = 0; l < 8; ++l) memcpy(c.bias[l], t[2 * l + 1].data(), 256 * 4);
  if (fold) {
    memcpy(c.bias[8], t[17].data(), 128 * 4);     // layer 8 = the folded view layer
  } else {
    memcpy(c.bias[8], t[19].data(), 256 * 4);
    memcpy(c.bias[9], t[17].data(), 128 * 4);
  }
  memcpy(c.w_alpha, t[20].data(), 256 * 4);
  memcpy(c.w_rgb, t[22].data(), 384 * 4);
  memcpy(c.w_unc, t[24].data(), 128 * 4);
  c.head_b[0] = t[21][0];
  c.head_b[1] = t[23][0]; c.head_b[2] = t[23][1]; c.head_b[3] = t[23][2];
  c.head_b[4] = t[25][0];
  for (int l = 0; l < n_layers(fold); ++l) c.inv_wscale[l] = 1.f / layer_wscale(t[layer_tensor(l, fold)]);
}

int launch_mlp_tc(const TcWeights& w, const RayBatch& rb, float* raw, int num_sms, cudaStream_t st, int heads) {
  TcArgs a{};
  a.rb = rb;
  a.M = (long long)rb.n * rb.S;
  a.out = raw;
  return launch_any<true>(w, a, num_sms, st, heads);
}

int launch_mlp_tc_points(const TcWeights& w, const float* pts, const float* dirs, int64_t M, float* out,
                         int num_sms, cudaStream_t st) {
  TcArgs a{};
  a.pts = pts;
  a.dirs = dirs;
  a.M = (long long)M;
  a.out = out;
  return launch_any<false>(w, a, num_sms, st);
}

}  // namespace evpp
