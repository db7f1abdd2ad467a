"""NGP-mode scorer: occupancy-bitfield ray marching + multiresolution hash-grid encoding + small density /
colour / NeurAR-uncertainty MLP, behind the same scoring signature as the reference-parity path
(``get_uncertainty(poses) -> Tensor[V]``, nerfServer/core/nerf.py:266).

The reference contains no such scorer; BASELINE.json's north_star mandates it.
Arithmetic follows the torch-ngp gridencoder / instant-ngp conventions; all of it runs in the CUDA library."""
import ctypes as C

import numpy as np
import torch

from ._lib import EvppNgpConfig, EvppNgpOut, check
from .engine import Engine, _ptr, _stream_ptr
from .run_nerf import _intrinsics

MLP_SHAPES = [("sigma0.w", (64, 32)), ("sigma1.w", (16, 64)), ("color0.w", (64, 31)), ("color1.w", (64, 64)),
              ("rgb.w", (3, 64)), ("unc.w", (1, 64)), ("unc.b", (1,))]


def level_table(n_levels=16, base_res=16, log2_hashmap=19, desired_res=2048):
    """Per-level (scale, resolution, offset, size) of the hash grid (torch-ngp gridencoder):
    scale = 2^(l*log2 b)*base - 1 with b = (desired/base)^(1/(L-1)), res = ceil(scale)+1,
    size = min(2^log2_hashmap, (res+1)^3) rounded up to a multiple of 8."""
    S = np.log2(desired_res / base_res) / (n_levels - 1)
    scale, res, offset, size, off = [], [], [], [], 0
    for l in range(n_levels):
        sc = np.float32(np.exp2(l * S) * base_res - 1.0)
        r = int(np.ceil(sc)) + 1
        n = int(np.ceil(min(2 ** log2_hashmap, (r + 1) ** 3) / 8) * 8)
        scale.append(sc); res.append(r); offset.append(off); size.append(n)
        off += n
    return (np.asarray(scale, np.float32), np.asarray(res, np.int32), np.asarray(offset, np.int64),
            np.asarray(size, np.int64), off)


def init_params(seed=0, table_entries=None):
    """Random-init NGP scorer parameters: fp16 table U(-1e-4, 1e-4), nn.Linear-style weights."""
    rng = np.random.RandomState(seed)
    if table_entries is None:
        table_entries = level_table()[4]
    p = {"table": rng.uniform(-1e-4, 1e-4, size=(table_entries, 2)).astype(np.float16)}
    for name, shape in MLP_SHAPES:
        if name == "unc.b":
            continue
        k = 1.0 / np.sqrt(shape[1])
        p[name] = rng.uniform(-k, k, size=shape).astype(np.float32)
    p["unc.b"] = rng.uniform(-0.125, 0.125, size=(1,)).astype(np.float32)
    return p


def pack_mlp(p):
    """fp32 [sigma0 64x32 | sigma1 16x64 | color0 64x32 (zero column 31) | color1 64x64 | heads 4x64 | unc bias]."""
    c0 = np.zeros((64, 32), np.float32)
    c0[:, :31] = p["color0.w"]
    heads = np.concatenate([p["rgb.w"], p["unc.w"]], 0).astype(np.float32)
    return np.concatenate([p["sigma0.w"].reshape(-1), p["sigma1.w"].reshape(-1), c0.reshape(-1), p["color1.w"].reshape(-1),
                           heads.reshape(-1), np.asarray(p["unc.b"], np.float32).reshape(-1)]).astype(np.float32)


def morton3d(x, y, z):
    """Interleave the low 10 bits of x, y, z (x in bit 0): torch-ngp's cell order inside one cascade."""
    def expand(v):
        v = np.asarray(v, dtype=np.uint32)
        v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
        v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
        v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
        v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
        return v
    with np.errstate(over="ignore"):
        return expand(x) | (expand(y) << np.uint32(1)) | (expand(z) << np.uint32(2))


def pack_cascade_bitfield(occ):
    """bool [C, H, H, H] (cascade, x, y, z) -> uint8 [C*H^3/8] bitfield in torch-ngp's layout: bit (index & 7) of byte
    index >> 3 with index = cascade * H^3 + morton3d(x, y, z)."""
    occ = np.asarray(occ, dtype=bool)
    Cc, H = occ.shape[0], occ.shape[1]
    g = np.arange(H, dtype=np.uint32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    m = morton3d(X, Y, Z).reshape(-1).astype(np.int64)
    flat = np.zeros((Cc, H ** 3), dtype=bool)
    for c in range(Cc):
        flat[c, m] = occ[c].reshape(-1)
    return np.packbits(flat.reshape(-1), bitorder="little")


class NGPScorer:
    def __init__(self, engine=None, device=0, H=400, W=400, focal=300.0, aabb_min=(-5.0, 0.01, -5.0),
                 aabb_max=(5.0, 5.01, 6.0), grid_res=128, near=0.5, far=6.0, dt=0.02, max_steps=1024, max_samples=256,
                 n_levels=16, base_res=16, log2_hashmap=19, desired_res=2048, white_bkgd=True, max_chunk_rays=0,
                 precision="fp16", march="uniform", cascades=1, dt_gamma=0.0, min_near=0.2):
        """precision: "fp16" = small MLP on the tensor cores (fp16 operands, fp32 accumulate); "fp32" = exact FFMA.
        march: "uniform" = constant step dt over one x-major grid_res^3 bitfield, one warp per ray with ballot / popc
        compaction; "torch_ngp" = the published torch-ngp marching rule (cubic scene box = [-2^(cascades-1), ..]^3,
        ``cascades`` Morton-ordered bitfields -- see pack_cascade_bitfield --, dt = clamp(t * dt_gamma, dt_min, dt_max),
        empty cells skipped to their far face, rays start at max(box entry, min_near) in normalised units)."""
        if march not in ("uniform", "torch_ngp"):
            raise ValueError("march must be 'uniform' or 'torch_ngp'")
        self.march, self.cascades = march, cascades
        self.engine = engine if engine is not None else Engine(device=device, precision=precision, near=near, far=far,
                                                               max_chunk_rays=max_chunk_rays)
        self.H, self.W, self.focal = H, W, focal
        self.K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
        self.grid_res, self.dt, self.max_samples = grid_res, dt, max_samples
        self.levels = level_table(n_levels, base_res, log2_hashmap, desired_res)
        self.cfg = EvppNgpConfig((C.c_float * 3)(*aabb_min), (C.c_float * 3)(*aabb_max), grid_res, near, far, dt,
                                 max_steps, max_samples, n_levels, int(bool(white_bkgd)),
                                 1 if march == "torch_ngp" else 0, int(cascades), float(dt_gamma), float(min_near))
        e = self.engine
        sc, rs, of, sz, _ = self.levels
        check(e.lib.evpp_ngp_configure(e._h, C.byref(self.cfg), sc.ctypes.data_as(C.c_void_p), rs.ctypes.data_as(C.c_void_p),
                                       of.ctypes.data_as(C.c_void_p), sz.ctypes.data_as(C.c_void_p)))

    @property
    def table_entries(self):
        return self.levels[4]

    def set_occupancy(self, bitfield):
        b = np.ascontiguousarray(bitfield, dtype=np.uint8)
        check(self.engine.lib.evpp_ngp_set_occupancy(self.engine._h, b.ctypes.data_as(C.c_void_p), b.size))

    def load_params(self, p):
        table = np.ascontiguousarray(p["table"], dtype=np.float16)
        mlp = pack_mlp(p)
        check(self.engine.lib.evpp_ngp_load_params(self.engine._h, table.ctypes.data_as(C.c_void_p), table.shape[0],
                                                   mlp.ctypes.data_as(C.c_void_p), mlp.size))

    def encode(self, u, want_indices=False):
        e = self.engine
        u = torch.as_tensor(u).to(e.device, torch.float32).reshape(-1, 3).contiguous()
        M = u.shape[0]
        enc = torch.empty(M, 32, device=e.device)
        idx = torch.empty(M, 16, 8, device=e.device, dtype=torch.int32) if want_indices else None
        with torch.cuda.device(e.device):
            check(e.lib.evpp_ngp_encode(e._h, _ptr(u), M, _ptr(enc), _ptr(idx), _stream_ptr(e.device)))
        return (enc, idx) if want_indices else enc

    def render_rays(self, rays_o, rays_d, want_samples=False):
        """Returns dict(rgb, depth, acc, beta2, counts[, t, raw, enc, total; torch_ngp march: delta, cell])."""
        e = self.engine
        ro = torch.as_tensor(rays_o).to(e.device, torch.float32).reshape(-1, 3).contiguous()
        rd = torch.as_tensor(rays_d).to(e.device, torch.float32).reshape(-1, 3).contiguous()
        n = ro.shape[0]
        o = {"rgb": torch.empty(n, 3, device=e.device), "depth": torch.empty(n, device=e.device),
             "acc": torch.empty(n, device=e.device), "beta2": torch.empty(n, device=e.device),
             "counts": torch.empty(n, device=e.device, dtype=torch.int32)}
        cap = n * self.max_samples if want_samples else 0
        if want_samples:
            o.update(t=torch.empty(cap, device=e.device), raw=torch.empty(cap, 5, device=e.device),
                     enc=torch.empty(cap, 32, device=e.device))
            if self.march == "torch_ngp":
                o.update(delta=torch.empty(cap, device=e.device), cell=torch.empty(cap, device=e.device, dtype=torch.int32))
        out = EvppNgpOut(*[(o[k].data_ptr() if k in o else None) for k in ("rgb", "depth", "acc", "beta2", "counts", "t", "raw", "enc")],
                         cap, 0, o["delta"].data_ptr() if "delta" in o else None, o["cell"].data_ptr() if "cell" in o else None)
        with torch.cuda.device(e.device):
            check(e.lib.evpp_ngp_render_rays(e._h, _ptr(ro), _ptr(rd), n, C.byref(out), _stream_ptr(e.device)))
        o["total"] = int(out.total_so_far)
        if want_samples:
            for k in ("t", "raw", "enc", "delta", "cell"):
                if k in o:
                    o[k] = o[k][:o["total"]]
        return o

    def get_uncertainty(self, poses, pix_idx=None):
        """Same contract as Controller.get_uncertainty (nerf.py:266-309): poses [V,4,4] -> cuda float32 [V]."""
        poses = torch.Tensor(np.asarray(poses, dtype=np.float64))[:, :3, :4].contiguous()
        V = poses.shape[0]
        fx, fy, cx, cy = _intrinsics(self.K)
        pix_t = None if pix_idx is None else torch.as_tensor(np.ascontiguousarray(pix_idx, dtype=np.int32))
        R = 0 if pix_t is None else pix_t.shape[1]
        scores = torch.empty(V, dtype=torch.float32)
        e = self.engine
        with torch.cuda.device(e.device):
            check(e.lib.evpp_ngp_score_views_host(e._h, _ptr(poses), V, self.H, self.W, fx, fy, cx, cy, _ptr(pix_t), R,
                                                  _ptr(scores), _stream_ptr(e.device)))
        return scores.to(e.device)

    def score_views(self, c2w, pix_idx=None):
        """Device poses [V,3,4] -> device scores [V] (no host copies)."""
        e = self.engine
        c2w = c2w.to(e.device, torch.float32).contiguous()
        V = c2w.shape[0]
        fx, fy, cx, cy = _intrinsics(self.K)
        R = 0
        if pix_idx is not None:
            pix_idx = pix_idx.to(e.device, torch.int32).contiguous()
            R = pix_idx.shape[1]
        scores = torch.empty(V, device=e.device)
        with torch.cuda.device(e.device):
            check(e.lib.evpp_ngp_score_views(e._h, _ptr(c2w), V, self.H, self.W, fx, fy, cx, cy, _ptr(pix_idx), R, _ptr(scores),
                                             _stream_ptr(e.device)))
        return scores

    def sample_count(self):
        n = C.c_int64(0)
        check(self.engine.lib.evpp_ngp_get_sample_count(self.engine._h, C.byref(n)))
        return n.value


# ---------------------------------------------------------------------------------------------------------
# TensoRF VM backbone (BASELINE configs[3]): same marching / compositing / scoring, other field
# ---------------------------------------------------------------------------------------------------------
TF_MLP_SHAPES = [("basis.w", (27, 144)), ("color0.w", (64, 43)), ("color1.w", (64, 64)), ("rgb.w", (3, 64)),
                 ("unc.w", (1, 64)), ("unc.b", (1,))]


def init_tensorf_params(seed=0, R=300):
    """Random-init VM grids (fp16 [3][R][R][64] planes indexed [i][y][x][c], [3][R][64] lines; 16 density + 48
    appearance channels per texel) and colour network (basis 144 -> 27, [basis, SH16] -> 64 -> 64 -> rgb + uncertainty)."""
    rng = np.random.RandomState(seed)
    p = {"planes": (rng.uniform(-1, 1, size=(3, R, R, 64)) * 0.35).astype(np.float16),
         "lines": (rng.uniform(-1, 1, size=(3, R, 64)) * 0.35).astype(np.float16), "R": R}
    for name, shape in TF_MLP_SHAPES:
        if name == "unc.b":
            continue
        k = 1.0 / np.sqrt(shape[1])
        p[name] = rng.uniform(-k, k, size=shape).astype(np.float32)
    p["unc.b"] = rng.uniform(-0.125, 0.125, size=(1,)).astype(np.float32)
    return p


class TensoRFScorer(NGPScorer):
    """NGPScorer with the hash grid replaced by vector-matrix decomposed grids."""

    def load_params(self, p):
        planes = np.ascontiguousarray(p["planes"], dtype=np.float16)
        lines = np.ascontiguousarray(p["lines"], dtype=np.float16)
        R = int(p["R"])
        assert planes.shape == (3, R, R, 64) and lines.shape == (3, R, 64)
        heads = np.concatenate([p["rgb.w"], p["unc.w"]], 0).astype(np.float32)
        mlp = np.concatenate([p["basis.w"].reshape(-1), p["color0.w"].reshape(-1), p["color1.w"].reshape(-1),
                              heads.reshape(-1), np.asarray(p["unc.b"], np.float32).reshape(-1)]).astype(np.float32)
        check(self.engine.lib.evpp_tensorf_load_params(self.engine._h, planes.ctypes.data_as(C.c_void_p),
                                                       lines.ctypes.data_as(C.c_void_p), R,
                                                       mlp.ctypes.data_as(C.c_void_p), mlp.size))
