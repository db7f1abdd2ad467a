"""CPU oracle for the NGP-mode scorer (occupancy-bitfield marching, multiresolution hash-grid
encoding, small density / colour / uncertainty MLP, compositing, per-view score).
TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference (small-zeng/EVPP) contains no hash grid, occupancy grid or
``cuda_ray`` path at all: its README links ashawkey/torch-ngp and environment.yml:271 pins
tinycudann==1.7, but no source file imports either.
BASELINE.json's north_star nevertheless mandates these stages, so this file *defines* their
arithmetic, following the published torch-ngp / instant-ngp conventions:

* grid encoder (torch-ngp ``gridencoder``): level scale ``2^(l*log2(b)) * base - 1``, resolution
  ``ceil(scale) + 1``, ``pos = x*scale + 0.5``, dense index while the running stride fits the table,
  otherwise the instant-ngp spatial hash (primes 1, 2654435761, 805459861), trilinear blend of
  2 features per level, 16 levels, table size 2^19;
* occupancy bitfield over the scene box, uniform steps ``t_i = t0 + (i+0.5)*dt`` between the box
  entry/exit clipped to [near, far], a sample is kept where its cell's bit is set;
* density net 32 -> 64 -> 16 (sigma = exp(h[0]), 15 geometric features), colour net
  [SH16(dir), geo] 31 -> 64 -> 64 -> 3 (sigmoid);
and, pinned by the reference, the NeurAR parts: uncertainty head ``beta = exp(linear(h))``
(nerfServer/core/run_nerf_helpers.py:128-129), alpha compositing ``w = alpha * prod(1 - alpha + 1e-10)``
(run_nerf.py:376-378) and the per-view score ``sum(beta^2) / rays`` over all evaluated samples
(nerfServer/core/nerf.py:307-309).

Integer quantities (sample counts, offsets, cell and table indices) are computed with exactly the fp32
operation order of the CUDA kernels (numpy float32 arithmetic rounds every operation), so the GPU tests
demand bit-exact equality for them; features and network outputs are compared with tolerances.
"""
from __future__ import annotations

import numpy as np

F = np.float32
PRIMES = (np.uint32(1), np.uint32(2654435761), np.uint32(805459861))


# ---------------------------------------------------------------------------------------------------
# configuration shared with evpp_b200/ngp.py (which holds its own copy of the level table builder)
# ---------------------------------------------------------------------------------------------------
def level_table(n_levels=16, base_res=16, log2_hashmap=19, desired_res=2048):
    """Per level: scale (fp32), resolution, table offset and size (entries).  torch-ngp gridencoder."""
    per_level_scale = np.exp2(np.log2(desired_res / base_res) / (n_levels - 1))
    S = np.log2(per_level_scale)
    max_params = 2 ** log2_hashmap
    out, offset = [], 0
    for l in range(n_levels):
        scale = F(np.exp2(l * S) * base_res - 1.0)
        res = int(np.ceil(scale)) + 1
        n = min(max_params, (res + 1) ** 3)
        n = int(np.ceil(n / 8) * 8)
        out.append(dict(scale=scale, res=res, offset=offset, size=n))
        offset += n
    return out, offset


def hash_index(pg, res, size):
    """pg: uint32 [..., 3] corner coordinates.  Dense while the stride fits, else spatial hash."""
    pg = pg.astype(np.uint32)
    stride = np.uint64(1)
    index = np.zeros(pg.shape[:-1], dtype=np.uint32)
    d = 0
    while d < 3 and stride <= np.uint64(size):
        index = (index + pg[..., d] * np.uint32(stride)).astype(np.uint32)
        stride = stride * np.uint64(res + 1)
        d += 1
    if stride > np.uint64(size):
        index = (pg[..., 0] * PRIMES[0]) ^ (pg[..., 1] * PRIMES[1]) ^ (pg[..., 2] * PRIMES[2])
    return (index % np.uint32(size)).astype(np.uint32)


def grid_encode(u, table, levels, return_indices=False):
    """u: fp32 [M,3] in [0,1]; table: fp16 [entries,2].  Returns fp32 [M, 2*L] (and int64 [M,L,8] indices)."""
    M = u.shape[0]
    out = np.zeros((M, 2 * len(levels)), dtype=F)
    all_idx = np.zeros((M, len(levels), 8), dtype=np.int64)
    with np.errstate(over="ignore"):
        for l, lv in enumerate(levels):
            pos = (u * lv["scale"]).astype(F) + F(0.5)
            pg = np.floor(pos).astype(F)
            w = (pos - pg).astype(F)
            pgi = pg.astype(np.uint32)
            acc = np.zeros((M, 2), dtype=F)
            for c in range(8):
                off = np.array([(c >> 0) & 1, (c >> 1) & 1, (c >> 2) & 1], dtype=np.uint32)
                wx = np.where(off[0], w[:, 0], F(1) - w[:, 0]).astype(F)
                wy = np.where(off[1], w[:, 1], F(1) - w[:, 1]).astype(F)
                wz = np.where(off[2], w[:, 2], F(1) - w[:, 2]).astype(F)
                wgt = ((wx * wy).astype(F) * wz).astype(F)
                idx = hash_index(pgi + off, lv["res"], lv["size"]).astype(np.int64) + lv["offset"]
                all_idx[:, l, c] = idx
                f = table[idx].astype(F)
                acc = (acc + wgt[:, None] * f).astype(F)          # fma on the GPU: compare with tolerance
            out[:, 2 * l:2 * l + 2] = acc
    return (out, all_idx) if return_indices else out


# ---------------------------------------------------------------------------------------------------
# occupancy-bitfield marching
# ---------------------------------------------------------------------------------------------------
def synthetic_occupancy(G=128, aabb_min=(-5.0, 0.01, -5.0), aabb_max=(5.0, 5.01, 6.0), seed=0):
    """Room shell (1-voxel walls of a 5 x 2.5 x 5 m box around the centre) + a 1 m cube + sparse specks,
    rasterised at G^3 over the scene box.  Returns uint8 bitfield [G^3/8]
    (bit i of byte j = cell 8j+i, cells in x-major (x*G+y)*G+z order)."""
    rng = np.random.RandomState(seed)
    amin, amax = np.asarray(aabb_min, dtype=np.float64), np.asarray(aabb_max, dtype=np.float64)
    c = (np.arange(G) + 0.5) / G
    X, Y, Z = np.meshgrid(amin[0] + c * (amax[0] - amin[0]), amin[1] + c * (amax[1] - amin[1]),
                          amin[2] + c * (amax[2] - amin[2]), indexing="ij")
    vox = (amax - amin) / G
    room_lo, room_hi = np.array([-2.5, 0.01, -2.0]), np.array([2.5, 2.51, 3.0])
    inside = (X > room_lo[0]) & (X < room_hi[0]) & (Y > room_lo[1]) & (Y < room_hi[1]) & (Z > room_lo[2]) & (Z < room_hi[2])
    core = (X > room_lo[0] + vox[0]) & (X < room_hi[0] - vox[0]) & (Y > room_lo[1] + vox[1]) & (Y < room_hi[1] - vox[1]) & \
           (Z > room_lo[2] + vox[2]) & (Z < room_hi[2] - vox[2])
    occ = inside & ~core
    occ |= (np.abs(X) < 0.5) & (Y < 1.0) & (np.abs(Z - 0.5) < 0.5)
    occ |= rng.rand(G, G, G) < 0.0005
    return np.packbits(occ.reshape(-1), bitorder="little"), occ


def march(rays_o, rays_d, bitfield, G, aabb_min, aabb_max, near, far, dt, max_steps, max_samples):
    """Returns (counts int32 [N], t list per ray).  fp32 op order == ngp.cu::march_ray."""
    o, d = rays_o.astype(F), rays_d.astype(F)
    amin, amax = np.asarray(aabb_min, dtype=F), np.asarray(aabb_max, dtype=F)
    inv_size = (F(1) / (amax - amin)).astype(F)
    dt, near, far = F(dt), F(near), F(far)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv_d = (F(1) / d).astype(F)
        ta = ((amin - o).astype(F) * inv_d).astype(F)
        tb = ((amax - o).astype(F) * inv_d).astype(F)
        tlo = np.fmin(ta, tb)
        thi = np.fmax(ta, tb)
        t0 = np.fmax(near, np.fmax(np.fmax(tlo[:, 0], tlo[:, 1]), tlo[:, 2])).astype(F)
        t1 = np.fmin(far, np.fmin(np.fmin(thi[:, 0], thi[:, 1]), thi[:, 2])).astype(F)
    N = o.shape[0]
    counts = np.zeros(N, dtype=np.int32)
    ts = []
    for r in range(N):
        kept = []
        if t1[r] > t0[r]:
            n_steps = int(np.ceil(((t1[r] - t0[r]).astype(F) / dt).astype(F)))
            n_steps = min(n_steps, max_steps)
            i = np.arange(n_steps, dtype=F)
            t = (t0[r] + ((i + F(0.5)).astype(F) * dt).astype(F)).astype(F)
            p = (o[r][None, :] + (d[r][None, :] * t[:, None]).astype(F)).astype(F)
            u = ((p - amin).astype(F) * inv_size).astype(F)
            cell = np.floor((u * F(G)).astype(F)).astype(np.int64)
            ok = ((cell >= 0) & (cell < G)).all(1)
            lin = (cell[:, 0] * G + cell[:, 1]) * G + cell[:, 2]
            lin = np.where(ok, lin, 0)
            bit = (bitfield[lin >> 3] >> (lin & 7)) & 1
            keep = ok & (bit == 1)
            kept = t[keep][:max_samples]
        counts[r] = len(kept)
        ts.append(np.asarray(kept, dtype=F))
    return counts, ts


# ---------------------------------------------------------------------------------------------------
# torch-ngp marching (march_mode 1)
# ---------------------------------------------------------------------------------------------------
# Restatement of the PUBLISHED algorithm of ashawkey/torch-ngp's `raymarching` extension (near_far_from_aabb +
# march_rays; README.md:25 of the reference links the project, nothing of it is in the reference tree and it is not
# installed here: PARITY UNPINNED -- there is no copy to execute or diff against).  What is restated: the scene cube
# [-bound, bound]^3 with bound = 2^(C-1); C nested occupancy grids of H^3 cells, Morton-ordered bitfield with
# index = cascade * H^3 + morton3D(x, y, z); the step dt = clamp(t * dt_gamma, dt_min, dt_max),
# dt_min = 2 sqrt3 / max_steps, dt_max = 2 sqrt3 * 2^(C-1) / H; the cascade of a sample =
# max(frexp-exponent of max|x|, frexp-exponent of dt * H / 2) clamped to [0, C-1]; an occupied cell gives a sample and
# advances by dt; an empty cell is left through its far face in dt-sized steps.  Every fp32 operation is rounded
# separately, in the order of evpp_b200/csrc/ngp.cu::tngp_march_kernel, so counts, cell indices, sample positions and
# step lengths can be compared bit for bit.
def morton3d(x, y, z):
    def expand(v):
        v = np.uint32(v)
        v = np.uint32((int(v) * 0x00010001) & 0xFF0000FF)
        v = np.uint32((int(v) * 0x00000101) & 0x0F00F00F)
        v = np.uint32((int(v) * 0x00000011) & 0xC30C30C3)
        v = np.uint32((int(v) * 0x00000005) & 0x49249249)
        return int(v)
    return expand(x) | (expand(y) << 1) | (expand(z) << 2)


def tngp_constants(aabb_min, aabb_max, H, cascades, max_steps, dt_gamma, min_near):
    amin, amax = np.asarray(aabb_min, dtype=F), np.asarray(aabb_max, dtype=F)
    bound = F(1 << (cascades - 1))
    half = F(0.5) * (amax[0] - amin[0])
    two_sqrt3 = F(2.0) * F(1.7320508075688772)
    return dict(bound=bound, scale=F(bound / half), center=(F(0.5) * (amin + amax)).astype(F), H=int(H), C=int(cascades),
                dt_gamma=F(dt_gamma), dt_min=F(two_sqrt3 / F(max_steps)), dt_max=F(F(two_sqrt3 * F(1 << (cascades - 1))) / F(H)),
                min_near=F(min_near), max_steps=int(max_steps))


def cascade_occupancy(H=32, cascades=2, seed=0, fill=0.03):
    """Synthetic nested occupancy: a solid blob shell + specks, bool [C, H, H, H] in (x, y, z) order per cascade;
    cascade c covers [-2^c, 2^c]^3 (capped at the bound)."""
    rng = np.random.RandomState(seed)
    occ = np.zeros((cascades, H, H, H), dtype=bool)
    g = (np.arange(H) + 0.5) / H * 2 - 1
    for c in range(cascades):
        b = min(2.0 ** c, 2.0 ** (cascades - 1))
        X, Y, Z = np.meshgrid(g * b, g * b, g * b, indexing="ij")
        r = np.sqrt(X ** 2 + Y ** 2 + Z ** 2)
        occ[c] = (np.abs(r - 0.6) < 0.08) | ((np.abs(X - 0.3) < 0.15) & (np.abs(Y) < 0.4) & (np.abs(Z + 0.2) < 0.1))
        occ[c] |= rng.rand(H, H, H) < fill
    return occ


def _mip(mx, C):
    e = int(np.frexp(F(mx))[1])
    return min(C - 1, max(0, e))


def march_tngp(rays_o, rays_d, bitfield, K, max_samples):
    """rays in WORLD coordinates (un-normalised directions).  Returns (counts int32 [N], and per ray the lists
    t_world fp32, delta fp32 (normalised step), cell int (bitfield index))."""
    o_w, d_w = rays_o.astype(F), rays_d.astype(F)
    N = o_w.shape[0]
    H, C = K["H"], K["C"]
    fH, rH = F(H), F(F(1) / F(H))
    H3 = H * H * H
    bound = K["bound"]
    counts = np.zeros(N, dtype=np.int32)
    ts, deltas, cells = [], [], []
    limit = min(K["max_steps"], max_samples)

    def clampf(v, lo, hi):
        return F(min(max(F(v), F(lo)), F(hi)))

    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for r in range(N):
            wx, wy, wz = d_w[r]
            ln = F(np.sqrt(F(F(F(wx * wx) + F(wy * wy)) + F(wz * wz))))
            o = [F(F(o_w[r][c] - K["center"][c]) * K["scale"]) for c in range(3)]
            d = [F(d_w[r][c] / ln) for c in range(3)]
            rd = [F(F(1) / d[c]) for c in range(3)]
            near = far = F(0)
            hit = True
            for c in range(3):
                a = F(F(-bound - o[c]) * rd[c])
                b = F(F(bound - o[c]) * rd[c])
                if a > b:
                    a, b = b, a
                if c == 0:
                    near, far = a, b
                else:
                    if near > b or a > far:
                        hit = False
                    if a > near:
                        near = a
                    if b < far:
                        far = b
            if near < K["min_near"]:
                near = K["min_near"]
            t_list, dl_list, ce_list = [], [], []
            if hit:
                t_to_world = F(F(1) / F(K["scale"] * ln))
                t = F(near)
                while t < far and len(t_list) < limit:
                    x = [clampf(F(o[c] + F(t * d[c])), -bound, bound) for c in range(3)]
                    dt = clampf(F(t * K["dt_gamma"]), K["dt_min"], K["dt_max"])
                    level = max(_mip(max(abs(x[0]), abs(x[1]), abs(x[2])), C), _mip(F(F(dt * fH) * F(0.5)), C))
                    mip_bound = F(min(F(2.0 ** level), bound))
                    mip_rbound = F(F(1) / mip_bound)
                    nc = [int(clampf(F(F(F(F(x[c] * mip_rbound) + F(1)) * F(0.5)) * fH), 0.0, F(H - 1))) for c in range(3)]
                    index = level * H3 + morton3d(nc[0], nc[1], nc[2])
                    occ = (int(bitfield[index >> 3]) >> (index & 7)) & 1
                    if occ:
                        t_list.append(F(t * t_to_world))
                        dl_list.append(dt)
                        ce_list.append(index)
                        t = F(t + dt)
                    else:
                        tt = F(3.4e38)
                        for c in range(3):
                            sg = F(np.copysign(F(0.5), d[c]))
                            face = F(F(F(F(F(F(nc[c]) + F(0.5)) + sg) * rH) * F(2)) - F(1))
                            face = F(face * mip_bound)
                            cand = F(F(face - x[c]) * rd[c])
                            if not np.isnan(cand):
                                tt = F(min(tt, cand))
                        tt = F(t + F(max(F(0), tt)))
                        while True:
                            t = F(t + clampf(F(t * K["dt_gamma"]), K["dt_min"], K["dt_max"]))
                            if not (t < tt and t < far):
                                break
            counts[r] = len(t_list)
            ts.append(np.asarray(t_list, dtype=F))
            deltas.append(np.asarray(dl_list, dtype=F))
            cells.append(np.asarray(ce_list, dtype=np.int64))
    return counts, ts, deltas, cells


# ---------------------------------------------------------------------------------------------------
# networks
# ---------------------------------------------------------------------------------------------------
def sh16(dirs):
    """Real spherical harmonics up to degree 3 (16 coefficients), instant-ngp / torch-ngp ordering."""
    x, y, z = dirs[:, 0].astype(np.float64), dirs[:, 1].astype(np.float64), dirs[:, 2].astype(np.float64)
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    out = np.stack([
        0.28209479177387814 * np.ones_like(x),
        -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2)], -1)
    return out


def init_ngp_params(seed=0, n_levels=16, base_res=16, log2_hashmap=19, desired_res=2048):
    """Random-init NGP scorer: tables U(-1e-4, 1e-4) (fp16), nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) weights."""
    rng = np.random.RandomState(seed)
    levels, total = level_table(n_levels, base_res, log2_hashmap, desired_res)
    table = rng.uniform(-1e-4, 1e-4, size=(total, 2)).astype(np.float16)

    def lin(o, i, bias=False):
        k = 1.0 / np.sqrt(i)
        return rng.uniform(-k, k, size=(o, i)).astype(F), (rng.uniform(-k, k, size=(o,)).astype(F) if bias else np.zeros(o, F))

    p = {"table": table}
    p["sigma0.w"], _ = lin(64, 32)
    p["sigma1.w"], _ = lin(16, 64)
    p["color0.w"], _ = lin(64, 31)
    p["color1.w"], _ = lin(64, 64)
    p["rgb.w"], _ = lin(3, 64)
    p["unc.w"], p["unc.b"] = lin(1, 64, bias=True)
    return p, levels


def spread_params(p):
    """Make scores of different views separate (like the reference-parity 'spread' variant): larger table
    amplitudes so the encoding carries signal, a positive density offset through the first feature."""
    q = {k: v.copy() for k, v in p.items()}
    q["table"] = (q["table"].astype(F) * F(2000.0)).astype(np.float16)
    q["unc.w"] = q["unc.w"] * F(4.0)
    return q


def mlp_forward(enc, dirs, p):
    """enc fp32 [M,32], dirs fp32 [M,3] (normalised).  Returns raw [M,5] = (r,g,b sigmoid, sigma, beta), float64 math."""
    e = enc.astype(np.float64)
    h = np.maximum(e @ p["sigma0.w"].astype(np.float64).T, 0.0)
    h = h @ p["sigma1.w"].astype(np.float64).T
    sigma = np.exp(np.clip(h[:, 0], -15.0, 15.0))                       # trunc_exp
    geo = h[:, 1:]
    x = np.concatenate([sh16(dirs), geo], -1)
    c = np.maximum(x @ p["color0.w"].astype(np.float64).T, 0.0)
    c = np.maximum(c @ p["color1.w"].astype(np.float64).T, 0.0)
    rgb = 1.0 / (1.0 + np.exp(-(c @ p["rgb.w"].astype(np.float64).T)))
    beta = np.exp(c @ p["unc.w"].astype(np.float64).T + p["unc.b"].astype(np.float64))   # run_nerf_helpers.py:128-129
    return np.concatenate([rgb, sigma[:, None], beta], -1)


def composite(raw, t, dt, white_bkgd=True):
    """One ray.  raw [S,5], t [S].  w = alpha * prod(1-alpha+1e-10) (run_nerf.py:376-378), constant step dt."""
    if len(t) == 0:
        bg = 1.0 if white_bkgd else 0.0
        return np.array([bg, bg, bg]), 0.0, 0.0, 0.0
    alpha = 1.0 - np.exp(-raw[:, 3] * float(dt))
    T = np.cumprod(np.concatenate([[1.0], 1.0 - alpha + 1e-10]))[:-1]
    w = alpha * T
    rgb = (w[:, None] * raw[:, :3]).sum(0)
    acc = w.sum()
    if white_bkgd:
        rgb = rgb + (1.0 - acc)
    return rgb, float((w * t).sum()), float(acc), float((raw[:, 4] ** 2).sum())


# ---------------------------------------------------------------------------------------------------
# TensoRF VM backbone (BASELINE configs[3]; parity unpinned like the rest of this file)
# ---------------------------------------------------------------------------------------------------
TF_MAT = ((0, 1), (0, 2), (1, 2))     # plane i spans axes (x = u[m0], y = u[m1])   (torch-ngp tensoRF: matMode)
TF_VEC = (2, 1, 0)                    # line i runs along this axis                  (vecMode)


def init_tensorf_params(seed=0, R=300):
    """VM-decomposed grids, 16 density + 48 appearance components fused into 64-channel texels (fp16):
    planes [3][R][R][64] indexed [i][y][x][c], lines [3][R][64]; basis 144 -> 27 (no bias), colour net
    [basis27, SH16] 43 -> 64 -> 64 -> 3, NeurAR head on the last hidden layer."""
    rng = np.random.RandomState(seed)
    p = {"planes": (rng.uniform(-1, 1, size=(3, R, R, 64)) * 0.35).astype(np.float16),
         "lines": (rng.uniform(-1, 1, size=(3, R, 64)) * 0.35).astype(np.float16), "R": R}

    def lin(o, i):
        k = 1.0 / np.sqrt(i)
        return rng.uniform(-k, k, size=(o, i)).astype(F)

    p["basis.w"] = lin(27, 144) * F(4.0)
    p["color0.w"] = lin(64, 43)
    p["color1.w"] = lin(64, 64)
    p["rgb.w"] = lin(3, 64)
    p["unc.w"] = lin(1, 64) * F(4.0)
    p["unc.b"] = rng.uniform(-0.125, 0.125, size=(1,)).astype(F)
    return p


def tensorf_features(u, p):
    """u fp32 [M,3] in [0,1].  Returns (sigma_feat [M], app [M,144]); bilinear plane x linear line, align_corners=True:
    pix = u*(R-1), i0 = clamp(floor(pix), 0, R-2), f = pix - i0 (fp32 like the kernel; blends in float64)."""
    R = p["R"]
    planes, lines = p["planes"].astype(np.float64), p["lines"].astype(np.float64)

    def split(x):
        pix = (x.astype(F) * F(R - 1)).astype(F)
        i0 = np.clip(np.floor(pix).astype(np.int64), 0, R - 2)
        return i0, (pix - i0.astype(F)).astype(F).astype(np.float64)

    M = u.shape[0]
    sig = np.zeros(M)
    app = np.zeros((M, 144))
    for i in range(3):
        x0, fx = split(u[:, TF_MAT[i][0]])
        y0, fy = split(u[:, TF_MAT[i][1]])
        z0, fz = split(u[:, TF_VEC[i]])
        P = planes[i]
        top = (1 - fx)[:, None] * P[y0, x0] + fx[:, None] * P[y0, x0 + 1]
        bot = (1 - fx)[:, None] * P[y0 + 1, x0] + fx[:, None] * P[y0 + 1, x0 + 1]
        pv = (1 - fy)[:, None] * top + fy[:, None] * bot
        lv = (1 - fz)[:, None] * lines[i][z0] + fz[:, None] * lines[i][z0 + 1]
        prod = pv * lv
        sig += prod[:, :16].sum(-1)
        app[:, i * 48:(i + 1) * 48] = prod[:, 16:]
    return sig, app


def tensorf_forward(u, dirs, p):
    """raw [M,5] = (r,g,b sigmoid, sigma = exp(clip(sigma_feat)), beta = exp(unc))."""
    sig, app = tensorf_features(u, p)
    b = app @ p["basis.w"].astype(np.float64).T
    x = np.concatenate([b, sh16(dirs)], -1)
    c = np.maximum(x @ p["color0.w"].astype(np.float64).T, 0.0)
    c = np.maximum(c @ p["color1.w"].astype(np.float64).T, 0.0)
    rgb = 1.0 / (1.0 + np.exp(-(c @ p["rgb.w"].astype(np.float64).T)))
    beta = np.exp(c @ p["unc.w"].astype(np.float64).T + p["unc.b"].astype(np.float64))
    return np.concatenate([rgb, np.exp(np.clip(sig, -15.0, 15.0))[:, None], beta], -1)


def render_rays(rays_o, rays_d, p, levels, bitfield, cfg):
    """Full NGP-mode pipeline for a ray list.  Returns dict(counts, offsets, t, rgb, depth, acc, beta2)."""
    amin, amax = np.asarray(cfg["aabb_min"], dtype=F), np.asarray(cfg["aabb_max"], dtype=F)
    counts, ts = march(rays_o, rays_d, bitfield, cfg["G"], amin, amax, cfg["near"], cfg["far"], cfg["dt"],
                       cfg["max_steps"], cfg["max_samples"])
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    N = rays_o.shape[0]
    rgb, depth, acc, beta2 = np.zeros((N, 3)), np.zeros(N), np.zeros(N), np.zeros(N)
    inv_size = (F(1) / (amax - amin)).astype(F)
    nrm = np.sqrt((rays_d.astype(np.float64) ** 2).sum(-1, keepdims=True))
    vd = (rays_d / nrm).astype(F)
    raws = []
    for r in range(N):
        t = ts[r]
        if len(t):
            pts = (rays_o[r].astype(F)[None, :] + (rays_d[r].astype(F)[None, :] * t[:, None]).astype(F)).astype(F)
            u = ((pts - amin).astype(F) * inv_size).astype(F)
            if "planes" in p:
                raw = tensorf_forward(u, np.repeat(vd[r][None, :], len(t), 0), p)
            else:
                enc = grid_encode(u, p["table"], levels)
                raw = mlp_forward(enc, np.repeat(vd[r][None, :], len(t), 0), p)
        else:
            raw = np.zeros((0, 5))
        raws.append(raw)
        rgb[r], depth[r], acc[r], beta2[r] = composite(raw, t, cfg["dt"], cfg.get("white_bkgd", True))
    return dict(counts=counts, offsets=offsets, t=np.concatenate(ts) if N else np.zeros(0, F), rgb=rgb, depth=depth,
                acc=acc, beta2=beta2, raw=np.concatenate(raws) if N else np.zeros((0, 5)))


def score_views(rays_o, rays_d, V, p, levels, bitfield, cfg):
    """Per-view score = sum over the view's rays and kept samples of beta^2 / rays per view (nerf.py:307-309)."""
    out = render_rays(rays_o, rays_d, p, levels, bitfield, cfg)
    R = rays_o.shape[0] // V
    return out["beta2"].reshape(V, R).sum(-1) / R, out
