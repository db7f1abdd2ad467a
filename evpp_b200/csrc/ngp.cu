// NGP-mode scorer (BASELINE.json north_star subsystems 1, 2 and 4): occupancy-bitfield ray marching with
// warp-level sample compaction, multiresolution hash-grid encoding, small density / colour / uncertainty
// MLP, compositing fused with the per-ray sum of beta^2.
//
// The reference contains none of this (parity unpinned): arithmetic follows the
// published torch-ngp / instant-ngp conventions and is mirrored by the CPU restatement used in the tests;
// the NeurAR parts keep the reference's definitions: beta = exp(linear(h)) (run_nerf_helpers.py:128-129),
// w = alpha * prod(1 - alpha + 1e-10) (run_nerf.py:376-378), score = sum(beta^2) / rays (nerf.py:307-309).
//
// Kernels
//   ngp_march_kernel         count pass, a warp per 32 rays: span per ray on its own lane, then the warp walks each ray
//                            32 steps per batch inside the box of the occupied cells: occupancy bit test, ballot +
//                            popc -> per-ray sample count (integer, bit-exact with the restatement), ballots stored
//   ngp_tile_sum/scan        exclusive scan of the counts -> sample offsets
//   ngp_fill_kernel          replays the stored ballots, writes the kept t values at offset + rank (compaction)
//   tngp_march_kernel        the torch-ngp marching rule (cascades, Morton bitfield, dt_gamma), thread per ray
//   ngp_encode_kernel        hash-grid encoding alone (parity + gather-bandwidth measurement)
//   ngp_field_kernel         per sample: encode (16 levels x 8 corners, half2 gathers from the L2-resident
//                            table) + density net + colour net + uncertainty head, fp32 FFMA (exact mode)
//   ngp_field_tc_kernel      same on the tensor cores (fp16 operands, fp32 accumulate): one warp per 32 samples,
//                            the five layers chained through mma.m16n8k16 register fragments
//   ngp_composite_kernel     one warp per ray: transmittance by warp scan, rgb / depth / acc and sum(beta^2)
#include "ngp.cuh"

namespace evpp {

namespace {

__device__ __forceinline__ uint32_t grid_index(uint32_t x, uint32_t y, uint32_t z, const NgpLevel& lv) {
  // torch-ngp gridencoder rule: dense index x + y*(res+1) + z*(res+1)^2 while the table holds (res+1)^3 entries,
  // else the instant-ngp spatial hash modulo the (power-of-two) table size.  Dense indices are < size by
  // construction, so the reference formula's "% size" is the mask below / a no-op.
  if (lv.hashed) return ((x * 1u) ^ (y * 2654435761u) ^ (z * 805459861u)) & (lv.size - 1u);
  return x + y * lv.stride_y + z * lv.stride_z;
}

// Encodes one normalised position; enc[2*l], enc[2*l+1] = trilinear blend of the level's 8 corners.
template <bool DUMP>
__device__ __forceinline__ void encode_point(const NgpParams& P, float ux, float uy, float uz, float* enc,
                                             int32_t* idx_out) {
  const __half2* __restrict__ table = reinterpret_cast<const __half2*>(P.table);
#pragma unroll
  for (int l = 0; l < kNgpMaxLevels; ++l) {      // n_levels == 16 is enforced at configuration time
    const NgpLevel lv = P.levels[l];
    const float px = __fadd_rn(__fmul_rn(ux, lv.scale), 0.5f);
    const float py = __fadd_rn(__fmul_rn(uy, lv.scale), 0.5f);
    const float pz = __fadd_rn(__fmul_rn(uz, lv.scale), 0.5f);
    const float gx = floorf(px), gy = floorf(py), gz = floorf(pz);
    const float wx = __fsub_rn(px, gx), wy = __fsub_rn(py, gy), wz = __fsub_rn(pz, gz);
    const uint32_t ix = (uint32_t)gx, iy = (uint32_t)gy, iz = (uint32_t)gz;
    uint32_t idx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      idx[c] = grid_index(ix + (c & 1), iy + ((c >> 1) & 1), iz + ((c >> 2) & 1), lv) + lv.offset;
    __half2 f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = __ldg(table + idx[c]);    // 8 independent 4-byte gathers in flight
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = __fmul_rn(__fmul_rn((c & 1) ? wx : __fsub_rn(1.f, wx), ((c >> 1) & 1) ? wy : __fsub_rn(1.f, wy)),
                                ((c >> 2) & 1) ? wz : __fsub_rn(1.f, wz));
      const float2 v = __half22float2(f[c]);
      a0 = fmaf(w, v.x, a0);
      a1 = fmaf(w, v.y, a1);
      if (DUMP) idx_out[l * 8 + c] = (int32_t)idx[c];
    }
    enc[2 * l] = a0;
    enc[2 * l + 1] = a1;
  }
}

__global__ void ngp_encode_kernel(NgpParams P, const float* __restrict__ u, long long M, float* __restrict__ enc_out,
                                  int32_t* __restrict__ idx_out) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float enc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) enc[k] = 0.f;
  if (idx_out) encode_point<true>(P, u[m * 3], u[m * 3 + 1], u[m * 3 + 2], enc, idx_out + m * (long long)P.n_levels * 8);
  else encode_point<false>(P, u[m * 3], u[m * 3 + 1], u[m * 3 + 2], enc, nullptr);
  if (enc_out) {
    float4* o = reinterpret_cast<float4*>(enc_out + m * 32);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = make_float4(enc[4 * k], enc[4 * k + 1], enc[4 * k + 2], enc[4 * k + 3]);
  }
}

// ---------------------------------------------------------------- marching ------------------------
struct RaySpan { float t0; int n_steps; int i_begin, i_end; };

__device__ __forceinline__ RaySpan ray_span(const NgpParams& P, const float* o, const float* d) {
  float lo = P.near_plane, hi = P.far_plane;
  float blo = -3.0e38f, bhi = 3.0e38f;      // the ray inside the (grown) box of the occupied cells
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float inv = __fdiv_rn(1.f, d[c]);
    const float ta = __fmul_rn(__fsub_rn(P.amin[c], o[c]), inv);
    const float tb = __fmul_rn(__fsub_rn(P.amax[c], o[c]), inv);
    lo = fmaxf(lo, fminf(ta, tb));      // fminf / fmaxf drop NaNs (0 * inf on axis-parallel rays)
    hi = fminf(hi, fmaxf(ta, tb));
    const float oa = (P.occ_lo[c] - o[c]) * inv, ob = (P.occ_hi[c] - o[c]) * inv;
    blo = fmaxf(blo, fminf(oa, ob));
    bhi = fminf(bhi, fmaxf(oa, ob));
    if (d[c] == 0.f && (o[c] < P.occ_lo[c] || o[c] > P.occ_hi[c])) bhi = -3.0e38f;   // parallel to the slab and outside it
  }
  RaySpan s;
  s.t0 = lo;
  s.n_steps = 0;
  s.i_begin = s.i_end = 0;
  if (hi > lo) {
    const int n = (int)ceilf(__fdiv_rn(__fsub_rn(hi, lo), P.dt));
    s.n_steps = n < P.max_steps ? n : P.max_steps;
    // Steps that can possibly be kept: those whose sample lies inside the grown occupied box, with two steps of slack
    // for the rounding of this estimate (the box itself is one whole cell larger than the occupied cells on every
    // side).  Everything outside has an empty cell by construction, so skipping it changes no count and no sample.
    if (bhi >= blo && P.occ_lo[0] <= P.occ_hi[0]) {
      const float inv_dt = __frcp_rn(P.dt);
      const float fb = floorf((blo - lo) * inv_dt) - 2.f, fe = ceilf((bhi - lo) * inv_dt) + 2.f;
      s.i_begin = fb > 0.f ? (fb < (float)s.n_steps ? (int)fb : s.n_steps) : 0;
      s.i_end = fe < (float)s.n_steps ? (fe > 0.f ? (int)fe : 0) : s.n_steps;
      if (s.i_end < s.i_begin) s.i_end = s.i_begin;
    }
  }
  return s;
}

// Count pass.  A warp owns 32 consecutive rays: lane L loads ray L and computes its span (six divisions and two slab
// tests per ray -- done once per ray instead of 32 times, with coalesced loads), then the whole warp walks the rays one
// after the other, 32 steps per lane batch: it tests the occupancy bit of each step's cell and ballots the kept lanes.
// The ballots are stored (masks[ray][word]) so that the fill pass replays them instead of recomputing cells and
// bitfield lookups.
__global__ void ngp_march_kernel(NgpParams P, const float* __restrict__ rays_o, const float* __restrict__ rays_d, int n,
                                 int32_t* __restrict__ counts, uint32_t* __restrict__ masks, int words_per_ray,
                                 float* __restrict__ t0_out) {
  const long long ray0 = ((((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5)) << 5;
  const int lane = threadIdx.x & 31;
  if (ray0 >= n) return;
  const long long mine = ray0 + lane;
  float mo[3] = {0.f, 0.f, 0.f}, md[3] = {1.f, 1.f, 1.f};
  RaySpan msp;
  msp.t0 = 0.f;
  msp.n_steps = msp.i_begin = msp.i_end = 0;
  if (mine < n) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { mo[c] = rays_o[mine * 3 + c]; md[c] = rays_d[mine * 3 + c]; }
    msp = ray_span(P, mo, md);
  }
  int my_count = 0;
  const int nr = (int)(n - ray0 < 32 ? n - ray0 : 32);
  for (int r = 0; r < nr; ++r) {
    float o[3], d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { o[c] = __shfl_sync(0xffffffffu, mo[c], r); d[c] = __shfl_sync(0xffffffffu, md[c], r); }
    const float t0 = __shfl_sync(0xffffffffu, msp.t0, r);
    const int i_begin = __shfl_sync(0xffffffffu, msp.i_begin, r), i_end = __shfl_sync(0xffffffffu, msp.i_end, r);
    uint32_t* mrow = masks + (ray0 + r) * words_per_ray;
    int count = 0;
    const int first = (i_begin >> 5) << 5;       // the walk starts at the ballot word that holds the first candidate step
    for (int w = lane; w < (first >> 5) && w < words_per_ray; w += 32) mrow[w] = 0u;   // skipped leading words: nothing kept
    // Four 32-step batches per iteration: their cell computations and bitfield loads are independent, so they are
    // issued together; the sequential rule "stop once the sample budget is reached" is applied afterwards, batch by
    // batch, exactly as a one-batch-at-a-time walk would.
    for (int base0 = first; base0 < i_end && count < P.max_samples; base0 += 128) {
      unsigned bal[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = base0 + j * 32 + lane;
        bool keep = false;
        if (i < i_end) {
          const float t = __fadd_rn(t0, __fmul_rn(__fadd_rn((float)i, 0.5f), P.dt));
          // cell = floor(u * G) per axis, inside the grid iff 0 <= cell < G: one round-down conversion and one
          // unsigned compare per axis (grid_res <= 1024, so the linear index fits 32 bits)
          uint32_t cell[3];
          bool ok = true;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float p = __fadd_rn(o[c], __fmul_rn(d[c], t));
            const float u = __fmul_rn(__fsub_rn(p, P.amin[c]), P.inv_size[c]);
            cell[c] = (uint32_t)__float2int_rd(__fmul_rn(u, (float)P.grid_res));
            ok = ok && cell[c] < (uint32_t)P.grid_res;
          }
          if (ok) {
            const uint32_t lin = (cell[0] * (uint32_t)P.grid_res + cell[1]) * (uint32_t)P.grid_res + cell[2];
            keep = (P.bitfield[lin >> 3] >> (lin & 7u)) & 1;
          }
        }
        bal[j] = __ballot_sync(0xffffffffu, keep);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int base = base0 + j * 32;
        if (base < i_end && count < P.max_samples) {
          if (lane == 0 && (base >> 5) < words_per_ray) mrow[base >> 5] = bal[j];
          count += __popc(bal[j]);
        }
      }
    }
    if (lane == r) my_count = count;
  }
  if (mine < n) {
    counts[mine] = my_count < P.max_samples ? my_count : P.max_samples;
    t0_out[mine] = msp.t0;
  }
}

// Fill pass: kept step i of a ray (bit i & 31 of ballot word i >> 5) gets slot offsets[ray] + (number of kept steps
// before it) -- the warp-level compaction of the count pass, replayed across the words.  Eight lanes per ray, lane
// w owns word w of a group of eight words, a shuffle scan of the popcounts gives its first slot.  Words the count
// pass never wrote (after the ray's last step, or after the sample budget was reached) only produce ranks
// >= counts[ray], which are dropped, so neither the step count nor the span has to be recomputed: the ray's first
// sample position t0 is all that is carried over.
__global__ void ngp_fill_kernel(NgpParams P, int n, const int32_t* __restrict__ counts, const int32_t* __restrict__ offsets,
                                const uint32_t* __restrict__ masks, int words_per_ray, const float* __restrict__ t0_in,
                                float* __restrict__ t_out, int32_t* __restrict__ ray_of_sample) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const bool live = gid < n;
  const int ray = live ? (int)gid : 0;
  const uint32_t* mrow = masks + (long long)ray * words_per_ray;
  const int cnt = live ? counts[ray] : 0;
  const long long base_out = offsets[ray];
  const float t0 = t0_in[ray];
  uint32_t m_next = (live && sub < words_per_ray) ? mrow[sub] : 0u;
  // groups of a warp run the same number of rounds (the shuffles need all 32 lanes)
  int rounds = cnt > 0 ? (words_per_ray + 7) >> 3 : 0;
#pragma unroll
  for (int k = 16; k >= 8; k >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, k));
  int carry = 0;
  for (int r = 0; r < rounds; ++r) {
    const int w = r * 8 + sub;
    uint32_t m = carry < cnt ? m_next : 0u;
    if (r + 1 < rounds) m_next = (live && w + 8 < words_per_ray) ? mrow[w + 8] : 0u;
    const int pc = __popc(m);
    int incl = pc;
#pragma unroll
    for (int k = 1; k < 8; k <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, k, 8);
      if (sub >= k) incl += v;
    }
    int rank = carry + incl - pc;
    while (m && rank < cnt) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const int i = w * 32 + b;
      t_out[base_out + rank] = __fadd_rn(t0, __fmul_rn(__fadd_rn((float)i, 0.5f), P.dt));
      ray_of_sample[base_out + rank] = ray;
      ++rank;
    }
    carry += __shfl_sync(0xffffffffu, incl, 7, 8);
  }
}

// ---------------------------------------------------------------- torch-ngp marcher (march_mode 1) ----------
// Restatement of the PUBLISHED ray-marching rule of ashawkey/torch-ngp (`raymarching`: near_far_from_aabb +
// march_rays): the scene cube is [-bound, bound]^3 in normalised coordinates, rays have unit directions and t is a
// normalised distance; `cascades` nested occupancy grids of H^3 cells (cascade c covers [-2^c, 2^c]^3, capped at the
// bound), each stored as a Morton-ordered bitfield; the step is dt = clamp(t * dt_gamma, dt_min, dt_max) with
// dt_min = 2 sqrt(3) / max_steps, dt_max = 2 sqrt(3) 2^(C-1) / H; the cascade of a sample is the larger of the one
// its position falls in and the one whose cells are at least dt wide; an occupied cell yields a sample and advances
// by dt, an empty cell is skipped to its far face in dt-sized steps.  The dependency is not in the reference tree
// (README.md:25 only links it): parity is UNPINNED; the test suite restates the same arithmetic in numpy,
// operation by operation (every product and sum below is a separately rounded fp32 operation for that reason) and
// the tests demand bit-equal counts, cell indices and sample positions.
// One THREAD per ray: the step recurrence is serial (dt depends on t); the count pass and the fill pass (after the
// exclusive scan of the counts -- a deterministic layout instead of torch-ngp's atomicAdd order) walk it twice.
__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
__device__ __forceinline__ uint32_t morton3d(uint32_t x, uint32_t y, uint32_t z) {
  return expand_bits10(x) | (expand_bits10(y) << 1) | (expand_bits10(z) << 2);
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ int mip_of(float mx, int C) {      // frexp exponent clamped to [0, C-1]
  int e;
  frexpf(mx, &e);
  return min(C - 1, max(0, e));
}

template <bool FILL>
__global__ void __launch_bounds__(128)
tngp_march_kernel(NgpParams P, const float* __restrict__ rays_o, const float* __restrict__ rays_d, int n,
                  int32_t* __restrict__ counts, const int32_t* __restrict__ offsets, float* __restrict__ t_out,
                  int32_t* __restrict__ ray_of_sample, float* __restrict__ delta_out, int32_t* __restrict__ cell_out) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= n) return;
  const int limit = FILL ? counts[ray] : min(P.max_steps, P.max_samples);
  if (FILL && limit == 0) return;
  float o[3], d[3], rd[3];
  const float wx = rays_d[(long long)ray * 3], wy = rays_d[(long long)ray * 3 + 1], wz = rays_d[(long long)ray * 3 + 2];
  const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)), __fmul_rn(wz, wz)));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = __fmul_rn(__fsub_rn(rays_o[(long long)ray * 3 + c], P.center[c]), P.scale);
    d[c] = __fdiv_rn(rays_d[(long long)ray * 3 + c], len);
    rd[c] = __fdiv_rn(1.f, d[c]);
  }
  // near_far_from_aabb on the cube [-bound, bound]^3 (slab test axis by axis, a miss leaves no samples)
  float near = 0.f, far = 0.f;
  bool hit = true;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float a = __fmul_rn(__fsub_rn(-P.bound, o[c]), rd[c]);
    float b = __fmul_rn(__fsub_rn(P.bound, o[c]), rd[c]);
    if (a > b) { const float tmp = a; a = b; b = tmp; }
    if (c == 0) {
      near = a; far = b;
    } else {
      if (near > b || a > far) hit = false;
      if (a > near) near = a;
      if (b < far) far = b;
    }
  }
  if (near < P.min_near) near = P.min_near;
  int steps = 0;
  if (hit) {
    const int H = P.grid_res, C = P.cascades;
    const float fH = (float)H, rH = __fdiv_rn(1.f, fH);
    const uint32_t H3 = (uint32_t)H * H * H;
    const float t_to_world = __fdiv_rn(1.f, __fmul_rn(P.scale, len));    // world parameter along the un-normalised ray
    const long long base_out = FILL ? offsets[ray] : 0;
    float t = near;
    while (t < far && steps < limit) {
      float x[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) x[c] = clampf(__fadd_rn(o[c], __fmul_rn(t, d[c])), -P.bound, P.bound);
      const float dt = clampf(__fmul_rn(t, P.dt_gamma), P.dt_min, P.dt_max);
      const int level = max(mip_of(fmaxf(fabsf(x[0]), fmaxf(fabsf(x[1]), fabsf(x[2]))), C),
                            mip_of(__fmul_rn(__fmul_rn(dt, fH), 0.5f), C));
      const float mip_bound = fminf(scalbnf(1.f, level), P.bound);
      const float mip_rbound = __fdiv_rn(1.f, mip_bound);
      int nc[3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        nc[c] = (int)clampf(__fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(x[c], mip_rbound), 1.f), 0.5f), fH), 0.f, (float)(H - 1));
      const uint32_t index = (uint32_t)level * H3 + morton3d((uint32_t)nc[0], (uint32_t)nc[1], (uint32_t)nc[2]);
      const bool occ = (P.bitfield[index >> 3] >> (index & 7u)) & 1;
      if (occ) {
        if (FILL) {
          t_out[base_out + steps] = __fmul_rn(t, t_to_world);
          ray_of_sample[base_out + steps] = ray;
          delta_out[base_out + steps] = dt;
          if (cell_out) cell_out[base_out + steps] = (int32_t)index;
        }
        ++steps;
        t = __fadd_rn(t, dt);
      } else {
        float tt = 3.4e38f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float face = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nc[c], 0.5f), copysignf(0.5f, d[c])), rH), 2.f), 1.f), mip_bound);
          tt = fminf(tt, __fmul_rn(__fsub_rn(face, x[c]), rd[c]));       // fminf drops the NaN of an axis-parallel ray
        }
        tt = __fadd_rn(t, fmaxf(0.f, tt));
        do {
          t = __fadd_rn(t, clampf(__fmul_rn(t, P.dt_gamma), P.dt_min, P.dt_max));
        } while (t < tt && t < far);      // (t < far: same outputs -- nothing is emitted past far -- and no endless walk when tt = inf)
      }
    }
  }
  if (!FILL) counts[ray] = steps;
}

// Exclusive scan of the per-ray counts in three small launches: per-tile sums (4096 counts per block), a single
// block scanning the (<= a few hundred) tile sums with the kernel below, and per-tile local scans.
constexpr int kScanTile = 4096;   // 256 threads x 16 counts

__global__ void __launch_bounds__(256) ngp_tile_sum_kernel(const int32_t* __restrict__ counts, int n, int32_t* __restrict__ tile_sums) {
  __shared__ int warp_sum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * kScanTile + threadIdx.x * 16;
  int s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += i0 + j < n ? counts[i0 + j] : 0;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if (lane == 0) warp_sum[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += warp_sum[k];
    tile_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) ngp_tile_scan_kernel(const int32_t* __restrict__ counts, int n,
                                                            const int32_t* __restrict__ tile_offsets,
                                                            int32_t* __restrict__ offsets) {
  __shared__ int warp_sum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * kScanTile + threadIdx.x * 16;
  int v[16];
  int mine = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) { v[j] = i0 + j < n ? counts[i0 + j] : 0; mine += v[j]; }
  int incl = mine;
#pragma unroll
  for (int k = 1; k < 32; k <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, k);
    if (lane >= k) incl += t;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  int run = tile_offsets[blockIdx.x] + incl - mine;
  for (int k = 0; k < warp; ++k) run += warp_sum[k];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (i0 + j < n) offsets[i0 + j] = run;
    run += v[j];
  }
}

// exclusive scan by ONE block: tiles of 4096 entries, warp-shuffle scans, running carry (used on the tile sums)
__global__ void ngp_scan_kernel(const int32_t* __restrict__ counts, int n, int32_t* __restrict__ offsets,
                                int32_t* __restrict__ total, unsigned long long* __restrict__ sample_counter) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 4096) {
    const int i0 = base + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = i0 + j < n ? counts[i0 + j] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int incl = mine;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, k);
      if (lane >= k) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int k = 1; k < 32; k <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, k);
        if (lane >= k) w += t;
      }
      warp_sum[lane] = w;     // inclusive over warps
    }
    __syncthreads();
    int run = carry_s + (warp ? warp_sum[warp - 1] : 0) + incl - mine;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j < n) offsets[i0 + j] = run;
      run += v[j];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *total = carry_s;
    if (sample_counter) *sample_counter += (unsigned long long)carry_s;   // single block, stream-ordered: no atomics needed
  }
}

// ---------------------------------------------------------------- field (encode + MLP) -------------
__device__ __forceinline__ void sh16(float x, float y, float z, float* o) {
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y; o[2] = 0.48860251190291987f * z; o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy; o[5] = -1.0925484305920792f * yz; o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz; o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2); o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2); o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2); o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// One dense layer for TWO samples of a thread: y[o] = sum_k WT[k][o] * x[k].  WT ([K][OUT], shared memory) is
// read as broadcast float4s, the inputs come from the thread's own shared-memory column (stride kFieldThreads),
// the 2 x OUT accumulators stay in registers: 2 * OUT FFMA per (OUT/4 + 2) shared loads.
constexpr int kFieldThreads = 128;
template <int K, int OUT>
__device__ __forceinline__ void dense2(const float* __restrict__ WT, const float* x0, const float* x1, float (&y0)[OUT],
                                       float (&y1)[OUT]) {
#pragma unroll
  for (int o = 0; o < OUT; ++o) { y0[o] = 0.f; y1[o] = 0.f; }
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const float a0 = x0[k * kFieldThreads], a1 = x1[k * kFieldThreads];
    const float4* w4 = reinterpret_cast<const float4*>(WT + k * OUT);
#pragma unroll
    for (int o = 0; o < OUT / 4; ++o) {
      const float4 w = w4[o];
      y0[4 * o] = fmaf(a0, w.x, y0[4 * o]); y0[4 * o + 1] = fmaf(a0, w.y, y0[4 * o + 1]);
      y0[4 * o + 2] = fmaf(a0, w.z, y0[4 * o + 2]); y0[4 * o + 3] = fmaf(a0, w.w, y0[4 * o + 3]);
      y1[4 * o] = fmaf(a1, w.x, y1[4 * o]); y1[4 * o + 1] = fmaf(a1, w.y, y1[4 * o + 1]);
      y1[4 * o + 2] = fmaf(a1, w.z, y1[4 * o + 2]); y1[4 * o + 3] = fmaf(a1, w.w, y1[4 * o + 3]);
    }
  }
}

// shared memory: transposed weights (kNgpMlpFloats) + per-thread activation columns x[2][64][kFieldThreads]
constexpr int kFieldSmemBytes = (kNgpMlpFloats + 3) / 4 * 16 + 2 * 64 * kFieldThreads * 4;

__global__ void __launch_bounds__(kFieldThreads, 2)
ngp_field_kernel(NgpParams P, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                 const float* __restrict__ t_s, const int32_t* __restrict__ ray_of_sample, long long M_cap,
                 const int32_t* __restrict__ total_dev, float* __restrict__ raw, float* __restrict__ enc_dump) {
  extern __shared__ __align__(16) float smem_w[];
  const long long M = total_dev ? (((long long)*total_dev) < M_cap ? (long long)*total_dev : M_cap) : M_cap;   // the marcher's total, read on the device
  float* WT_s0 = smem_w;                     // [32][64]
  float* WT_s1 = WT_s0 + 32 * 64;            // [64][16]
  float* WT_c0 = WT_s1 + 64 * 16;            // [32][64] (input 31 is the zero pad)
  float* WT_c1 = WT_c0 + 32 * 64;            // [64][64]
  float* WT_hd = WT_c1 + 64 * 64;            // [64][4]: r, g, b, uncertainty
  float* xs = smem_w + (kNgpMlpFloats + 3) / 4 * 4;
  {   // transpose the packed [OUT][K] matrices into [K][OUT]
    const float* g = P.mlp;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) WT_s0[(i % 32) * 64 + i / 32] = g[i];
    g += 64 * 32;
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) WT_s1[(i % 64) * 16 + i / 64] = g[i];
    g += 16 * 64;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) WT_c0[(i % 32) * 64 + i / 32] = g[i];
    g += 64 * 32;
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) WT_c1[(i % 64) * 64 + i / 64] = g[i];
    g += 64 * 64;
    for (int i = threadIdx.x; i < 4 * 64; i += blockDim.x) WT_hd[(i % 64) * 4 + i / 64] = g[i];
  }
  const float unc_b = P.mlp[kNgpMlpFloats - 1];
  __syncthreads();
  float* x0 = xs + threadIdx.x;                               // sample slot 0: x0[k * kFieldThreads]
  float* x1 = xs + 64 * kFieldThreads + threadIdx.x;          // sample slot 1
  for (long long base = (long long)blockIdx.x * (2 * kFieldThreads); base < M; base += (long long)gridDim.x * (2 * kFieldThreads)) {
    float sh[2][16];
    long long ms[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long m = base + j * kFieldThreads + threadIdx.x;
      ms[j] = m;
      float* x = j ? x1 : x0;
      float enc[32];
      if (m < M) {
        const int ray = ray_of_sample[m];
        const float t = t_s[m];
        float d[3], u[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          d[c] = rays_d[(long long)ray * 3 + c];
          u[c] = __fmul_rn(__fsub_rn(__fadd_rn(rays_o[(long long)ray * 3 + c], __fmul_rn(d[c], t)), P.amin[c]), P.inv_size[c]);
        }
        encode_point<false>(P, u[0], u[1], u[2], enc, nullptr);
        const float inv_n = rsqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        sh16(d[0] * inv_n, d[1] * inv_n, d[2] * inv_n, sh[j]);
        if (enc_dump) {
#pragma unroll
          for (int k = 0; k < 32; ++k) enc_dump[m * 32 + k] = enc[k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) enc[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) sh[j][k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k * kFieldThreads] = enc[k];
    }
    float sigma[2];
    {
      float h0[64], h1[64];
      dense2<32, 64>(WT_s0, x0, x1, h0, h1);                  // density net, layer 0 (+ReLU)
#pragma unroll
      for (int k = 0; k < 64; ++k) { x0[k * kFieldThreads] = fmaxf(h0[k], 0.f); x1[k * kFieldThreads] = fmaxf(h1[k], 0.f); }
    }
    {
      float g0[16], g1[16];
      dense2<64, 16>(WT_s1, x0, x1, g0, g1);                  // density net, layer 1: sigma + 15 geometric features
      sigma[0] = expf(fminf(fmaxf(g0[0], -15.f), 15.f));      // trunc_exp
      sigma[1] = expf(fminf(fmaxf(g1[0], -15.f), 15.f));
#pragma unroll
      for (int k = 0; k < 16; ++k) { x0[k * kFieldThreads] = sh[0][k]; x1[k * kFieldThreads] = sh[1][k]; }
#pragma unroll
      for (int k = 0; k < 15; ++k) { x0[(16 + k) * kFieldThreads] = g0[1 + k]; x1[(16 + k) * kFieldThreads] = g1[1 + k]; }
      x0[31 * kFieldThreads] = 0.f; x1[31 * kFieldThreads] = 0.f;
    }
    {
      float h0[64], h1[64];
      dense2<32, 64>(WT_c0, x0, x1, h0, h1);                  // colour net, layer 0 (+ReLU)
#pragma unroll
      for (int k = 0; k < 64; ++k) { x0[k * kFieldThreads] = fmaxf(h0[k], 0.f); x1[k * kFieldThreads] = fmaxf(h1[k], 0.f); }
    }
    {
      float h0[64], h1[64];
      dense2<64, 64>(WT_c1, x0, x1, h0, h1);                  // colour net, layer 1 (+ReLU)
#pragma unroll
      for (int k = 0; k < 64; ++k) { x0[k * kFieldThreads] = fmaxf(h0[k], 0.f); x1[k * kFieldThreads] = fmaxf(h1[k], 0.f); }
    }
    float hd0[4], hd1[4];
    dense2<64, 4>(WT_hd, x0, x1, hd0, hd1);                   // rgb + uncertainty heads
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (ms[j] < M) {
        const float* hd = j ? hd1 : hd0;
        float* r = raw + ms[j] * 5;
        r[0] = 1.f / (1.f + expf(-hd[0]));
        r[1] = 1.f / (1.f + expf(-hd[1]));
        r[2] = 1.f / (1.f + expf(-hd[2]));
        r[3] = sigma[j];
        r[4] = expf(hd[3] + unc_b);                            // NeurAR head, run_nerf_helpers.py:128-129
      }
    }
  }
}

// ---------------------------------------------------------------- field on the tensor cores ---------
// Same network, fp16 operands / fp32 accumulation, one warp per 32 samples.  The five layers are chained in
// REGISTERS: the fp32 accumulator fragments of mma.m16n8k16 (rows g / g+8, columns 2t, 2t+1 of an 8-wide tile)
// of two neighbouring output tiles are exactly the A-operand fragment of the next layer's 16-wide K step, so
// after ReLU + pack the activations never leave the register file.  Only the two network inputs (the 32 hash-grid
// features and the 16 SH coefficients, which each lane computes for its own sample) are transposed into fragment
// layout through a per-warp shared-memory staging tile.  Weights live in shared memory as fp16 [out][in] rows
// padded by 8 halves (conflict-free 32-bit fragment loads).
// (These small GEMMs -- N <= 64, K <= 64 -- cannot fill a 128-row tcgen05 tile without a second staging pass
// through shared or tensor memory per layer; the warp-level MMA keeps the chain in registers instead.)
#ifndef EVPP_NGP_TC_CTAS
#define EVPP_NGP_TC_CTAS 2
#endif
constexpr int kTcWarps = 8;
constexpr int kLd32 = 40, kLd64 = 72;            // padded row strides (halves) for K = 32 / K = 64 matrices
constexpr int kTcW_s0 = 0;                       // [64][kLd32]
constexpr int kTcW_s1 = kTcW_s0 + 64 * kLd32;    // [16][kLd64]
constexpr int kTcW_c0 = kTcW_s1 + 16 * kLd64;    // [64][kLd32]: k 0..15 = SH, 16 = sigma column (zero), 17..31 = geo
constexpr int kTcW_c1 = kTcW_c0 + 64 * kLd32;    // [64][kLd64]
constexpr int kTcW_hd = kTcW_c1 + 64 * kLd64;    // [8][kLd64]: r, g, b, uncertainty, 4 zero rows
constexpr int kTcW_end = kTcW_hd + 8 * kLd64;
constexpr int kTcStageHalves = 32 * kLd32 + 32 * 24;                    // per warp: features [32][kLd32] + SH [32][24]
constexpr int kFieldTcSmemBytes = (kTcW_end + kTcWarps * kTcStageHalves) * 2;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// acc[mt][nt] (+)= A[mt][ks] * W^T for one layer; W: shared fp16 [N][LDW] (row = output unit)
template <int KS, int NT, int LDW>
__device__ __forceinline__ void layer_mma(const __half* __restrict__ W, const uint32_t (&a)[2][KS][4], float (&acc)[2][NT][4],
                                          int g, int t, int k_off = 0) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const __half* w = W + (nt * 8 + g) * LDW + k_off + ks * 16 + 2 * t;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(w), b1 = *reinterpret_cast<const uint32_t*>(w + 8);
      mma16816(acc[0][nt], a[0][ks], b0, b1);
      mma16816(acc[1][nt], a[1][ks], b0, b1);
    }
  }
}
// accumulator tiles (2j, 2j+1) -> A fragment of K step j of the next layer (optionally through ReLU)
template <int NT, bool RELU>
__device__ __forceinline__ void to_operand(const float (&acc)[2][NT][4], uint32_t (&a)[2][NT / 2][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < NT / 2; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float* c = acc[mt][2 * j + (q >> 1)];
        float x = c[(q & 1) * 2], y = c[(q & 1) * 2 + 1];
        if (RELU) { x = fmaxf(x, 0.f); y = fmaxf(y, 0.f); }
        a[mt][j][q] = pack_h2(x, y);
      }
}
template <int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[2][NT][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;
}
// fragment loads of a staged [32][LD] fp16 tile: rows = samples of the warp, K step ks
template <int LD>
__device__ __forceinline__ void load_a(const __half* st, int ks, int g, int t, uint32_t (&a0)[4], uint32_t (&a1)[4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    uint32_t* a = mt ? a1 : a0;
    const __half* r0 = st + (mt * 16 + g) * LD + ks * 16 + 2 * t;
    const __half* r1 = r0 + 8 * LD;
    a[0] = *reinterpret_cast<const uint32_t*>(r0);
    a[1] = *reinterpret_cast<const uint32_t*>(r1);
    a[2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
    a[3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
  }
}

__global__ void __launch_bounds__(32 * kTcWarps, EVPP_NGP_TC_CTAS)
ngp_field_tc_kernel(NgpParams P, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                    const float* __restrict__ t_s, const int32_t* __restrict__ ray_of_sample, long long M_cap,
                    const int32_t* __restrict__ total_dev, float* __restrict__ raw, float* __restrict__ enc_dump) {
  extern __shared__ __align__(16) __half smem_h[];
  const long long M = total_dev ? (((long long)*total_dev) < M_cap ? (long long)*total_dev : M_cap) : M_cap;   // the marcher's total, read on the device
  {   // fp32 packed weights -> padded fp16 [out][in]
    const float* g = P.mlp;
    for (int i = threadIdx.x; i < kTcW_end; i += blockDim.x) smem_h[i] = __float2half_rn(0.f);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) smem_h[kTcW_s0 + (i / 32) * kLd32 + i % 32] = __float2half_rn(g[i]);
    g += 64 * 32;
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) smem_h[kTcW_s1 + (i / 64) * kLd64 + i % 64] = __float2half_rn(g[i]);
    g += 16 * 64;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
      // packed colour layer 0 columns: 0..15 SH, 16..30 geo, 31 zero  ->  k 0..15 SH, 16 zero (sigma), 17..31 geo
      const int o = i / 32, k = i % 32;
      if (k < 16) smem_h[kTcW_c0 + o * kLd32 + k] = __float2half_rn(g[i]);
      else if (k < 31) smem_h[kTcW_c0 + o * kLd32 + k + 1] = __float2half_rn(g[i]);
    }
    g += 64 * 32;
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) smem_h[kTcW_c1 + (i / 64) * kLd64 + i % 64] = __float2half_rn(g[i]);
    g += 64 * 64;
    for (int i = threadIdx.x; i < 4 * 64; i += blockDim.x) smem_h[kTcW_hd + (i / 64) * kLd64 + i % 64] = __float2half_rn(g[i]);
  }
  const float unc_b = P.mlp[kNgpMlpFloats - 1];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  __half* st_enc = smem_h + kTcW_end + warp * kTcStageHalves;
  __half* st_sh = st_enc + 32 * kLd32;
  const long long n_groups = (M + 31) / 32;
  for (long long grp = (long long)blockIdx.x * kTcWarps + warp; grp < n_groups; grp += (long long)gridDim.x * kTcWarps) {
    const long long m = grp * 32 + lane;
    {   // each lane: encoding + SH of its own sample -> staging rows
      float enc[32], sh[16];
#pragma unroll
      for (int k = 0; k < 32; ++k) enc[k] = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) sh[k] = 0.f;
      if (m < M) {
        const int ray = ray_of_sample[m];
        const float tt = t_s[m];
        float d[3], u[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          d[c] = rays_d[(long long)ray * 3 + c];
          u[c] = __fmul_rn(__fsub_rn(__fadd_rn(rays_o[(long long)ray * 3 + c], __fmul_rn(d[c], tt)), P.amin[c]), P.inv_size[c]);
        }
        encode_point<false>(P, u[0], u[1], u[2], enc, nullptr);
        const float inv_n = rsqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        sh16(d[0] * inv_n, d[1] * inv_n, d[2] * inv_n, sh);
        if (enc_dump) {
#pragma unroll
          for (int k = 0; k < 32; ++k) enc_dump[m * 32 + k] = enc[k];
        }
      }
      __syncwarp();   // previous group's fragment loads are done
      uint32_t* e32 = reinterpret_cast<uint32_t*>(st_enc + lane * kLd32);
#pragma unroll
      for (int k = 0; k < 16; ++k) e32[k] = pack_h2(enc[2 * k], enc[2 * k + 1]);
      uint32_t* s32 = reinterpret_cast<uint32_t*>(st_sh + lane * 24);
#pragma unroll
      for (int k = 0; k < 8; ++k) s32[k] = pack_h2(sh[2 * k], sh[2 * k + 1]);
      __syncwarp();
    }
    // ---- density net ----
    uint32_t aE[2][2][4];
    load_a<kLd32>(st_enc, 0, g, t, aE[0][0], aE[1][0]);
    load_a<kLd32>(st_enc, 1, g, t, aE[0][1], aE[1][1]);
    uint32_t aH[2][4][4];
    {
      float acc[2][8][4];
      zero_acc<8>(acc);
      layer_mma<2, 8, kLd32>(smem_h + kTcW_s0, aE, acc, g, t);
      to_operand<8, true>(acc, aH);
    }
    uint32_t aX[2][2][4];       // colour-net input: K step 0 = SH, K step 1 = [sigma_raw, geo15]
    float sig[2][2];
    {
      float acc[2][2][4];
      zero_acc<2>(acc);
      layer_mma<4, 2, kLd64>(smem_h + kTcW_s1, aH, acc, g, t);
      uint32_t aG[2][1][4];
      to_operand<2, false>(acc, aG);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        sig[mt][0] = acc[mt][0][0];      // column 0 of rows g / g+8 lives on the lanes with t == 0
        sig[mt][1] = acc[mt][0][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) aX[mt][1][q] = aG[mt][0][q];
      }
    }
    load_a<24>(st_sh, 0, g, t, aX[0][0], aX[1][0]);
    // ---- colour net ----
    uint32_t aC[2][4][4];
    {
      float acc[2][8][4];
      zero_acc<8>(acc);
      layer_mma<2, 8, kLd32>(smem_h + kTcW_c0, aX, acc, g, t);
      to_operand<8, true>(acc, aC);
    }
    uint32_t aD[2][4][4];
    {
      float acc[2][8][4];
      zero_acc<8>(acc);
      layer_mma<4, 8, kLd64>(smem_h + kTcW_c1, aC, acc, g, t);
      to_operand<8, true>(acc, aD);
    }
    float hd[2][1][4];
    zero_acc<1>(hd);
    layer_mma<4, 1, kLd64>(smem_h + kTcW_hd, aD, hd, g, t);
    // ---- outputs: lanes t == 0 hold (r, g) and sigma, lanes t == 1 hold (b, uncertainty) of rows g and g + 8
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const long long row = grp * 32 + mt * 16 + g + hrow * 8;
        if (row < M) {
          const float c0 = hd[mt][0][hrow * 2], c1 = hd[mt][0][hrow * 2 + 1];
          float* r = raw + row * 5;
          if (t == 0) {
            r[0] = 1.f / (1.f + expf(-c0));
            r[1] = 1.f / (1.f + expf(-c1));
            r[3] = expf(fminf(fmaxf(sig[mt][hrow], -15.f), 15.f));      // trunc_exp
          } else if (t == 1) {
            r[2] = 1.f / (1.f + expf(-c0));
            r[4] = expf(c1 + unc_b);                                     // NeurAR head, run_nerf_helpers.py:128-129
          }
        }
      }
  }
}

// ---------------------------------------------------------------- TensoRF VM field ------------------
// BASELINE configs[3]: the hash grid replaced by vector-matrix decomposed grids (three R x R planes and three
// R-long lines, 16 density + 48 appearance components fused into 64-channel fp16 texels = one 128-byte line per
// bilinear corner).  Eight lanes share a sample and read one 16-byte slice of each texel, so every corner is a
// single coalesced 128-byte access; density = sum of plane x line products, the 144 appearance products go
// through a per-warp staging tile into the same register-chained warp-MMA colour network as the hash-grid
// backbone (basis 144 -> 27, [basis, SH16] -> 64 -> 64 -> rgb + NeurAR head).
constexpr int kTfLdApp = 152;                                    // 144 appearance features + 8 pad (halves)
constexpr int kTfLdC0 = 56;                                      // 48 inputs + 8 pad
constexpr int kTfW_bs = 0;                                       // [32][kTfLdApp]  (27 rows used)
constexpr int kTfW_c0 = kTfW_bs + 32 * kTfLdApp;                 // [64][kTfLdC0]: k 0..26 basis, 27..31 zero, 32..47 SH
constexpr int kTfW_c1 = kTfW_c0 + 64 * kTfLdC0;                  // [64][kLd64]
constexpr int kTfW_hd = kTfW_c1 + 64 * kLd64;                    // [8][kLd64]
constexpr int kTfW_end = kTfW_hd + 8 * kLd64;
constexpr int kTfStageHalves = 32 * kTfLdApp + 32 * 24 + 64;     // app [32][152] + SH [32][24] + sigma_feat [32] fp32
// warps per CTA (one CTA per SM): 16 = 204.6 KB of shared memory, 124 registers, no spills.  The kernel is latency-bound
// (ncu: L1 hit rate 87 %, L2 4 %, issue-active 48 % at 8 warps), so residency pays: 8 / 12 / 16 warps = 25.8 / 31.0 /
// 32.9 k views/s on the Room sweep shape.  EVPP_TF_WARPS = 8 | 12 | 16 selects the shape at run time.
constexpr int tensorf_smem_bytes(int warps) { return (kTfW_end + warps * kTfStageHalves) * 2; }

__device__ __forceinline__ void tf_split(float u, int R, int& i0, float& f) {
  const float pix = __fmul_rn(u, (float)(R - 1));              // align_corners = True
  i0 = min(max((int)floorf(pix), 0), R - 2);
  f = __fsub_rn(pix, (float)i0);
}
__device__ __forceinline__ void h8_to_f(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 x = __half22float2(h[k]); f[2 * k] = x.x; f[2 * k + 1] = x.y; }
}

template <int kTfWarps>
__global__ void __launch_bounds__(32 * kTfWarps, 1)
tensorf_field_kernel(NgpParams P, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                     const float* __restrict__ t_s, const int32_t* __restrict__ ray_of_sample, long long M_cap,
                     const int32_t* __restrict__ total_dev, float* __restrict__ raw) {
  extern __shared__ __align__(16) __half smem_h[];
  const long long M = total_dev ? (((long long)*total_dev) < M_cap ? (long long)*total_dev : M_cap) : M_cap;   // the marcher's total, read on the device
  {
    const float* g = P.tf_mlp;
    for (int i = threadIdx.x; i < kTfW_end; i += blockDim.x) smem_h[i] = __float2half_rn(0.f);
    __syncthreads();
    for (int i = threadIdx.x; i < 27 * 144; i += blockDim.x) smem_h[kTfW_bs + (i / 144) * kTfLdApp + i % 144] = __float2half_rn(g[i]);
    g += 27 * 144;
    for (int i = threadIdx.x; i < 64 * 43; i += blockDim.x) {
      const int o = i / 43, k = i % 43;          // packed columns: 0..26 basis, 27..42 SH -> k 0..26, 32..47
      smem_h[kTfW_c0 + o * kTfLdC0 + (k < 27 ? k : k + 5)] = __float2half_rn(g[i]);
    }
    g += 64 * 43;
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) smem_h[kTfW_c1 + (i / 64) * kLd64 + i % 64] = __float2half_rn(g[i]);
    g += 64 * 64;
    for (int i = threadIdx.x; i < 4 * 64; i += blockDim.x) smem_h[kTfW_hd + (i / 64) * kLd64 + i % 64] = __float2half_rn(g[i]);
  }
  const float unc_b = P.tf_mlp[kTfMlpFloats - 1];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  __half* st_app = smem_h + kTfW_end + warp * kTfStageHalves;
  __half* st_sh = st_app + 32 * kTfLdApp;
  float* st_sig = reinterpret_cast<float*>(st_sh + 32 * 24);
  const int R = P.tf_res;
  const uint4* planes = reinterpret_cast<const uint4*>(P.tf_planes);   // 8 uint4 per texel
  const uint4* lines = reinterpret_cast<const uint4*>(P.tf_lines);
  const long long n_groups = (M + 31) / 32;
  for (long long grp = (long long)blockIdx.x * kTfWarps + warp; grp < n_groups; grp += (long long)gridDim.x * kTfWarps) {
    __syncwarp();
    // ---- SH of the lane's own sample ----
    {
      const long long m = grp * 32 + lane;
      float sh[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) sh[k] = 0.f;
      if (m < M) {
        const int ray = ray_of_sample[m];
        const float dx = rays_d[(long long)ray * 3], dy = rays_d[(long long)ray * 3 + 1], dz = rays_d[(long long)ray * 3 + 2];
        const float inv_n = rsqrtf(dx * dx + dy * dy + dz * dz);
        sh16(dx * inv_n, dy * inv_n, dz * inv_n, sh);
      }
      uint32_t* s32 = reinterpret_cast<uint32_t*>(st_sh + lane * 24);
#pragma unroll
      for (int k = 0; k < 8; ++k) s32[k] = pack_h2(sh[2 * k], sh[2 * k + 1]);
    }
    // ---- VM features: 8 lanes per sample (16-byte channel slices), 4 samples per pass ----
    const int sub = lane & 7;
#pragma unroll 1
    for (int pass = 0; pass < 8; ++pass) {
      const int srow = pass * 4 + (lane >> 3);
      const long long m = grp * 32 + srow;
      float u[3] = {0.f, 0.f, 0.f};
      if (m < M) {
        const int ray = ray_of_sample[m];
        const float tt = t_s[m];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          u[c] = __fmul_rn(__fsub_rn(__fadd_rn(rays_o[(long long)ray * 3 + c], __fmul_rn(rays_d[(long long)ray * 3 + c], tt)), P.amin[c]), P.inv_size[c]);
      }
      float dsum = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int m0 = i == 2 ? 1 : 0, m1 = i == 0 ? 1 : 2, mv = 2 - i;     // matMode (0,1) (0,2) (1,2); vecMode 2,1,0
        int x0, y0, z0;
        float fx, fy, fz;
        tf_split(u[m0], R, x0, fx);
        tf_split(u[m1], R, y0, fy);
        tf_split(u[mv], R, z0, fz);
        const uint4* pl = planes + (((long long)i * R + y0) * R + x0) * 8 + sub;
        const uint4 v00 = __ldg(pl), v01 = __ldg(pl + 8), v10 = __ldg(pl + (long long)R * 8), v11 = __ldg(pl + (long long)R * 8 + 8);
        const uint4* ln = lines + ((long long)i * R + z0) * 8 + sub;
        const uint4 l0 = __ldg(ln), l1 = __ldg(ln + 8);
        float a[8], b[8], c[8], d[8], p0[8], p1[8];
        h8_to_f(v00, a); h8_to_f(v01, b); h8_to_f(v10, c); h8_to_f(v11, d); h8_to_f(l0, p0); h8_to_f(l1, p1);
        float prod[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float top = fmaf(fx, b[k] - a[k], a[k]), bot = fmaf(fx, d[k] - c[k], c[k]);
          const float pv = fmaf(fy, bot - top, top);
          const float lv = fmaf(fz, p1[k] - p0[k], p0[k]);
          prod[k] = pv * lv;
        }
        if (sub < 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) dsum += prod[k];
        } else {
          uint4 o;
          o.x = pack_h2(prod[0], prod[1]); o.y = pack_h2(prod[2], prod[3]);
          o.z = pack_h2(prod[4], prod[5]); o.w = pack_h2(prod[6], prod[7]);
          *reinterpret_cast<uint4*>(st_app + srow * kTfLdApp + i * 48 + (sub - 2) * 8) = o;
        }
      }
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);      // lanes sub 0 and 1 hold the 16 density channels
      if (sub == 0) st_sig[srow] = dsum;
    }
    __syncwarp();
    // ---- colour network on the tensor cores ----
    uint32_t aX[2][3][4];       // [basis32 (2 K steps) | SH16]
    {
      uint32_t aA[2][9][4];
#pragma unroll
      for (int ks = 0; ks < 9; ++ks) load_a<kTfLdApp>(st_app, ks, g, t, aA[0][ks], aA[1][ks]);
      float acc[2][4][4];
      zero_acc<4>(acc);
      layer_mma<9, 4, kTfLdApp>(smem_h + kTfW_bs, aA, acc, g, t);
      uint32_t aB[2][2][4];
      to_operand<4, false>(acc, aB);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int q = 0; q < 4; ++q) { aX[mt][0][q] = aB[mt][0][q]; aX[mt][1][q] = aB[mt][1][q]; }
    }
    load_a<24>(st_sh, 0, g, t, aX[0][2], aX[1][2]);
    uint32_t aC[2][4][4];
    {
      float acc[2][8][4];
      zero_acc<8>(acc);
      layer_mma<3, 8, kTfLdC0>(smem_h + kTfW_c0, aX, acc, g, t);
      to_operand<8, true>(acc, aC);
    }
    uint32_t aD[2][4][4];
    {
      float acc[2][8][4];
      zero_acc<8>(acc);
      layer_mma<4, 8, kLd64>(smem_h + kTfW_c1, aC, acc, g, t);
      to_operand<8, true>(acc, aD);
    }
    float hd[2][1][4];
    zero_acc<1>(hd);
    layer_mma<4, 1, kLd64>(smem_h + kTfW_hd, aD, hd, g, t);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int srow = mt * 16 + g + hrow * 8;
        const long long row = grp * 32 + srow;
        if (row < M) {
          const float c0 = hd[mt][0][hrow * 2], c1 = hd[mt][0][hrow * 2 + 1];
          float* r = raw + row * 5;
          if (t == 0) {
            r[0] = 1.f / (1.f + expf(-c0));
            r[1] = 1.f / (1.f + expf(-c1));
            r[3] = expf(fminf(fmaxf(st_sig[srow], -15.f), 15.f));
          } else if (t == 1) {
            r[2] = 1.f / (1.f + expf(-c0));
            r[4] = expf(c1 + unc_b);                                     // NeurAR head, run_nerf_helpers.py:128-129
          }
        }
      }
  }
}

// ---------------------------------------------------------------- compositing ----------------------
// Eight lanes per ray (four rays per warp): kept samples per ray average below ten, so a full warp per ray would
// idle most lanes.  Transmittance = exclusive product of (1 - alpha + 1e-10) by an 8-wide shuffle scan, carried
// across groups of eight samples in T.
__global__ void ngp_composite_kernel(NgpParams P, const float* __restrict__ raw, const float* __restrict__ t_s,
                                     const float* __restrict__ delta, const int32_t* __restrict__ counts, const int32_t* __restrict__ offsets, int n,
                                     float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ acc,
                                     float* __restrict__ beta2) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const bool live = gid < n;
  const int ray = live ? (int)gid : 0;
  const int cnt = live ? counts[ray] : 0;
  const long long off = live ? offsets[ray] : 0;
  int maxcnt = cnt;
#pragma unroll
  for (int k = 16; k >= 8; k >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, k));
  float T = 1.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, dep = 0.f, a = 0.f, b2 = 0.f;
  for (int base = 0; base < maxcnt; base += 8) {
    const int s = base + sub;
    float alpha = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, tt = 0.f, beta = 0.f;
    if (s < cnt) {
      const float* r = raw + (off + s) * 5;
      r0 = r[0]; r1 = r[1]; r2 = r[2];
      alpha = 1.f - expf(-r[3] * (delta ? delta[off + s] : P.dt));
      beta = r[4];
      tt = t_s[off + s];
    }
    const float f = s < cnt ? (1.f - alpha + 1e-10f) : 1.f;   // run_nerf.py:378
    float incl = f;
#pragma unroll
    for (int k = 1; k < 8; k <<= 1) {
      const float v = __shfl_up_sync(0xffffffffu, incl, k, 8);
      if (sub >= k) incl *= v;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1, 8);
    if (sub == 0) excl = 1.f;
    const float w = alpha * T * excl;
    c0 = fmaf(w, r0, c0); c1 = fmaf(w, r1, c1); c2 = fmaf(w, r2, c2);
    dep = fmaf(w, tt, dep);
    a += w;
    b2 = fmaf(beta, beta, b2);
    T *= __shfl_sync(0xffffffffu, incl, 7, 8);
  }
#pragma unroll
  for (int k = 4; k > 0; k >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, k); c1 += __shfl_xor_sync(0xffffffffu, c1, k);
    c2 += __shfl_xor_sync(0xffffffffu, c2, k); dep += __shfl_xor_sync(0xffffffffu, dep, k);
    a += __shfl_xor_sync(0xffffffffu, a, k); b2 += __shfl_xor_sync(0xffffffffu, b2, k);
  }
  if (live && sub == 0) {
    if (P.white_bkgd) { c0 += 1.f - a; c1 += 1.f - a; c2 += 1.f - a; }
    if (rgb) { rgb[(long long)ray * 3] = c0; rgb[(long long)ray * 3 + 1] = c1; rgb[(long long)ray * 3 + 2] = c2; }
    if (depth) depth[ray] = dep;
    if (acc) acc[ray] = a;
    if (beta2) beta2[ray] = b2;
  }
}

}  // namespace

int launch_ngp_encode(const NgpParams& P, const float* u, long long M, float* enc, int32_t* idx, cudaStream_t st) {
  if (M <= 0) return 0;
  ngp_encode_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(P, u, M, enc, idx);
  return 1;
}

int ngp_mask_words(const NgpParams& P) {
  // n_steps <= min(max_steps, ceil((far - near) / dt)) up to rounding of the fp32 span: + 2 steps of margin
  const double span = ((double)P.far_plane - (double)P.near_plane) / (double)P.dt;
  long long bound = (long long)std::ceil(span) + 2;
  if (bound > P.max_steps) bound = P.max_steps;
  if (bound < 1) bound = 1;
  return (int)((bound + 31) / 32);
}

int ngp_scan_tiles(int n) { return (n + kScanTile - 1) / kScanTile; }

int launch_ngp_march_count(const NgpParams& P, const float* ro, const float* rd, int n, int32_t* counts, uint32_t* masks,
                           float* t0, cudaStream_t st) {
  if (n <= 0) return 0;
  const long long warps = ((long long)n + 31) / 32;       // a warp per 32 rays
  ngp_march_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(P, ro, rd, n, counts, masks, ngp_mask_words(P), t0);
  return 1;
}

// tile_ws: 2 * ngp_scan_tiles(n) ints
int launch_ngp_scan(const int32_t* counts, int n, int32_t* offsets, int32_t* total, int32_t* tile_ws,
                    unsigned long long* sample_counter, cudaStream_t st) {
  if (n <= 0) return 0;
  const int tiles = ngp_scan_tiles(n);
  ngp_tile_sum_kernel<<<tiles, 256, 0, st>>>(counts, n, tile_ws);
  ngp_scan_kernel<<<1, 1024, 0, st>>>(tile_ws, tiles, tile_ws + tiles, total, sample_counter);
  ngp_tile_scan_kernel<<<tiles, 256, 0, st>>>(counts, n, tile_ws + tiles, offsets);
  return 3;
}

int launch_ngp_march_fill(const NgpParams& P, int n, const int32_t* counts, const int32_t* offsets, const uint32_t* masks,
                          const float* t0, float* t, int32_t* ray_of_sample, cudaStream_t st) {
  if (n <= 0) return 0;
  ngp_fill_kernel<<<(unsigned)(((long long)n * 8 + 255) / 256), 256, 0, st>>>(P, n, counts, offsets, masks, ngp_mask_words(P), t0,
                                                                              t, ray_of_sample);
  return 1;
}

int launch_ngp_field(const NgpParams& P, const float* ro, const float* rd, const float* t, const int32_t* ray_of_sample,
                     long long M, const int32_t* total_dev, float* raw, float* enc_dump, int num_sms, bool tensor_cores,
                     cudaStream_t st) {
  if (M <= 0) return 0;
  if (tensor_cores) {
    cudaFuncSetAttribute(ngp_field_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFieldTcSmemBytes);
    long long blocks = ((M + 31) / 32 + kTcWarps - 1) / kTcWarps;
    const long long cap = (long long)num_sms * EVPP_NGP_TC_CTAS;
    if (blocks > cap) blocks = cap;
    ngp_field_tc_kernel<<<(unsigned)blocks, 32 * kTcWarps, kFieldTcSmemBytes, st>>>(P, ro, rd, t, ray_of_sample, M, total_dev, raw, enc_dump);
    return 1;
  }
  cudaFuncSetAttribute(ngp_field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFieldSmemBytes);
  long long blocks = (M + 2 * kFieldThreads - 1) / (2 * kFieldThreads);
  const long long cap = (long long)num_sms * 2;      // 2 resident CTAs per SM, grid-stride over the rest
  if (blocks > cap) blocks = cap;
  ngp_field_kernel<<<(unsigned)blocks, kFieldThreads, kFieldSmemBytes, st>>>(P, ro, rd, t, ray_of_sample, M, total_dev, raw, enc_dump);
  return 1;
}

template <int WARPS>
static void launch_tensorf(const NgpParams& P, const float* ro, const float* rd, const float* t, const int32_t* ray_of_sample,
                           long long M, const int32_t* total_dev, float* raw, int num_sms, cudaStream_t st) {
  constexpr int smem = tensorf_smem_bytes(WARPS);
  cudaFuncSetAttribute(tensorf_field_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long blocks = ((M + 31) / 32 + WARPS - 1) / WARPS;
  if (blocks > num_sms) blocks = num_sms;
  tensorf_field_kernel<WARPS><<<(unsigned)blocks, 32 * WARPS, smem, st>>>(P, ro, rd, t, ray_of_sample, M, total_dev, raw);
}

int launch_tensorf_field(const NgpParams& P, const float* ro, const float* rd, const float* t, const int32_t* ray_of_sample,
                         long long M, const int32_t* total_dev, float* raw, int num_sms, cudaStream_t st) {
  if (M <= 0) return 0;
  const char* e = getenv("EVPP_TF_WARPS");
  const int w = e && *e ? atoi(e) : 16;
  if (w == 8) launch_tensorf<8>(P, ro, rd, t, ray_of_sample, M, total_dev, raw, num_sms, st);
  else if (w == 12) launch_tensorf<12>(P, ro, rd, t, ray_of_sample, M, total_dev, raw, num_sms, st);
  else launch_tensorf<16>(P, ro, rd, t, ray_of_sample, M, total_dev, raw, num_sms, st);
  return 1;
}

int launch_ngp_composite(const NgpParams& P, const float* raw, const float* t, const float* delta, const int32_t* counts,
                         const int32_t* offsets, int n, float* rgb, float* depth, float* acc, float* beta2, cudaStream_t st) {
  if (n <= 0) return 0;
  ngp_composite_kernel<<<(unsigned)(((long long)n * 8 + 255) / 256), 256, 0, st>>>(P, raw, t, delta, counts, offsets, n, rgb, depth, acc, beta2);
  return 1;
}

int launch_tngp_march_count(const NgpParams& P, const float* ro, const float* rd, int n, int32_t* counts, cudaStream_t st) {
  if (n <= 0) return 0;
  tngp_march_kernel<false><<<(n + 127) / 128, 128, 0, st>>>(P, ro, rd, n, counts, nullptr, nullptr, nullptr, nullptr, nullptr);
  return 1;
}

int launch_tngp_march_fill(const NgpParams& P, const float* ro, const float* rd, int n, const int32_t* counts,
                           const int32_t* offsets, float* t, int32_t* ray_of_sample, float* delta, int32_t* cell_index,
                           cudaStream_t st) {
  if (n <= 0) return 0;
  tngp_march_kernel<true><<<(n + 127) / 128, 128, 0, st>>>(P, ro, rd, n, const_cast<int32_t*>(counts), offsets, t, ray_of_sample, delta,
                                                           cell_index);
  return 1;
}

}  // namespace evpp
