"""CPU oracle for the planner's stage-1 TSDF occupancy ray-cast scorer.  TEST INFRASTRUCTURE ONLY.

Restates ``TSDF.render_rays`` / ``isectSegAABB_gpu`` / ``get_uncertainty_tsdf``
(plannerServer_Object/core/tsdf_gpu.py:198-354, :67-114, :473-530; the Room copy differs only in the
two coefficients, plannerServer_Room/core/tsdf_gpu.py:281,298) in torch-CPU fp32 ops, quirks included:

* 56 uniform samples in [near, far] (:41, :205-207), voxel = floor((p - origin) / voxel_size) (:211);
* out-of-bounds rewrite through SEQUENTIAL ``where``s that re-read already rewritten coordinates
  (:233-247): z < 0 zeroes x, y, z; else y < 0 zeroes x, y; else x < 0 zeroes x -- and the upper bound
  (>= dim-1) rewrites to **0**, not dim-1, with the same cascade;
* unknown count n0 = #(state == 0), u = c_unk * n0 / 56 clipped to 1 (:276-282);
* first hit k = argmin(where(state == 2, 3, 10)) (first minimum), k == 0 counts as "no hit" (:285-290);
* hit rays: depth = near + k/56 * (far - near) evaluated in float64 then cast (:293),
  u = 1 / (1 + c_s * depth * Nray[hit]) (:298);
* rays missing the box ``aabb_bnds`` (slab test that skips axes with |d| < 1e-6 entirely, :67-114):
  u = 1e-10, depth = far (:295-300);
* per view: mean of u over the rays, mean of the depths strictly inside (near, far) else far (:513-529).

Pin: the reference module cannot be imported here (open3d / skimage / imageio missing) and its
``render_rays`` relies on torch-1.x "sequence of index arrays" indexing that torch 2.x dropped
(SURVEY.md section 8a row a13).  ``run_reference_render_rays`` below therefore executes the UNMODIFIED
source text of ``isectSegAABB_gpu`` and ``TSDF.render_rays``, cut out of the reference file with ``ast``,
under a tensor subclass that restores the legacy indexing rule -- nothing else is changed -- and
``oracle/make_golden_tsdf.py`` asserts this restatement equals it bit for bit before writing
``tests/golden/tsdf_*.npz``.  Parity status: pinned by outputs of the reference's own function bodies.
"""
from __future__ import annotations

import ast
import math
import os
import time
import types

import numpy as np
import torch

REF_FILE = "/root/reference/plannerServer_Object/core/tsdf_gpu.py"

NEAR, FAR, VOXEL = 0.5, 6.0, 0.1                     # tsdf_gpu.py:39-40
N_SAMPLES = int((FAR - NEAR) / VOXEL) + 1            # :41 -> 56
VOL_BNDS_CABIN = np.array([[-5, 5], [0.01, 5.0], [-5, 6]], dtype=np.float64)      # :30
AABB_CABIN = np.array([[-1.0, 1.0], [0.02, 2.5], [-1.5, 2.0]], dtype=np.float64)  # :31


def volume_geometry(vol_bnds=VOL_BNDS_CABIN, voxel=VOXEL):
    """fusion.py:37-39: dims = ceil(extent / voxel), origin = lower bounds as float32."""
    dims = np.ceil((vol_bnds[:, 1] - vol_bnds[:, 0]) / voxel).astype(int)
    origin = vol_bnds[:, 0].astype(np.float32)
    return dims, origin


def isect_aabb(rays_o, rays_d, amin, amax, tmin=0.0, tmax=10000.0):
    """tsdf_gpu.py:67-114 (fp32; the float64 box corners enter as 0-dim tensors and do not promote)."""
    n = rays_o.shape[0]
    lo = torch.full((n,), tmin, dtype=torch.float32)
    hi = torch.full((n,), tmax, dtype=torch.float32)
    hit = torch.ones(n, dtype=torch.float32)
    for i in range(3):
        valid = torch.abs(rays_d[:, i]) >= 1e-6
        ood = 1.0 / rays_d[:, i]
        a = torch.tensor(float(amin[i]), dtype=torch.float64)
        b = torch.tensor(float(amax[i]), dtype=torch.float64)
        t1 = (a - rays_o[:, i]) * ood
        t2 = (b - rays_o[:, i]) * ood
        t1, t2 = torch.minimum(t1, t2), torch.maximum(t1, t2)          # swap where t1 > t2 (:89-92)
        lo = torch.where(valid & (t1 > lo), t1, lo)
        hi = torch.where(valid & (t2 < hi), t2, hi)
        hit = torch.where(valid & (lo > hi), torch.zeros_like(hit), hit)
    return hit, torch.stack([lo, hi], -1)


def voxel_indices(pts, origin, voxel, dims):
    """tsdf_gpu.py:210-249: floor + the cascading out-of-bounds rewrite.  pts [..., 3] fp32."""
    idx = torch.floor((pts - torch.from_numpy(np.asarray(origin, dtype=np.float32))) / torch.tensor(voxel, dtype=torch.float32))
    x, y, z = idx[..., 0], idx[..., 1], idx[..., 2]
    zero = torch.zeros_like(x)
    x = torch.where((x < 0) | (y < 0) | (z < 0), zero, x)
    y = torch.where((x < 0) | (y < 0) | (z < 0), zero, y)
    z = torch.where((x < 0) | (y < 0) | (z < 0), zero, z)
    x = torch.where((x >= dims[0] - 1) | (y >= dims[1] - 1) | (z >= dims[2] - 1), zero, x)
    y = torch.where((x >= dims[0] - 1) | (y >= dims[1] - 1) | (z >= dims[2] - 1), zero, y)
    z = torch.where((x >= dims[0] - 1) | (y >= dims[1] - 1) | (z >= dims[2] - 1), zero, z)
    return torch.stack([x, y, z], -1).long()


def render_rays(rays_o, rays_d, state, nray, origin, voxel=VOXEL, near=NEAR, far=FAR, aabb=AABB_CABIN,
                c_unknown=0.2, c_surface=0.1, stages=None):
    """tsdf_gpu.py:198-354.  state / nray: float32 [X,Y,Z] volumes.  Returns (uncer [N], depth [N])."""
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    n_samples = int((far - near) / voxel) + 1
    t_vals = torch.linspace(0., 1., steps=n_samples)
    z_vals = torch.full((rays_o.shape[0], 1), near) * (1. - t_vals) + torch.full((rays_o.shape[0], 1), far) * t_vals
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    vi = voxel_indices(pts, origin, voxel, state.shape)
    st = state[vi[..., 0], vi[..., 1], vi[..., 2]]
    nr = nray[vi[..., 0], vi[..., 1], vi[..., 2]]
    N = st.shape[0]
    depth = torch.ones(N) * far
    uncer = torch.zeros(N)
    n0 = torch.sum((st == 0).float(), 1)
    has = n0 != 0
    uncer[has] = c_unknown * (n0[has] / n_samples)
    uncer[uncer > 1] = 1
    buff = torch.where(st == 2, torch.full_like(st, 3.), torch.full_like(st, 10.))
    k = torch.min(buff, 1)[1].numpy().astype(np.int32)
    occ_ray = np.nonzero(k)[0]
    kk = k[occ_ray]
    depth[occ_ray] = torch.tensor(near + kk / n_samples * (far - near)).float()
    hit, _ = isect_aabb(rays_o, rays_d, aabb[:, 0], aabb[:, 1])
    uncer[occ_ray] = torch.div(1.0, 1.0 + c_surface * depth[occ_ray] * nr[occ_ray, kk])
    miss = hit == 0
    uncer[miss] = 1e-10
    depth[miss] = far
    if stages is not None:
        stages.update(voxel_idx=vi, n_unknown=n0.int(), first_hit=torch.from_numpy(k), aabb_hit=hit.int())
    return uncer, depth


def view_means(uncer, depth, V, R, near=NEAR, far=FAR):
    """tsdf_gpu.py:513-529."""
    u = uncer.numpy().reshape(V, R)
    d = depth.numpy().reshape(V, R)
    uncers = np.mean(u, axis=1).tolist()
    depths = []
    for i in range(V):
        sel = d[i][(d[i] < far) & (d[i] > near)]
        depths.append(float(far) if sel.shape[0] == 0 else float(np.mean(sel, axis=0)))
    return uncers, depths


def synthetic_state_volume(dims, seed=0):
    """Synthetic TSDF state map for tests and benchmarks: 0 = unknown, 1 = empty, 2 = occupied.
    An observed (empty) ball around the scene centre, an occupied box shell and floor inside it, noise
    specks; voxel [0,0,0] forced empty as the reference does (tsdf_gpu.py:779).  Nray = hits per surface voxel."""
    rng = np.random.RandomState(seed)
    X, Y, Z = [int(v) for v in dims]
    gx, gy, gz = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    state = np.zeros((X, Y, Z), dtype=np.float32)
    c = np.array([X / 2, Y / 3, Z / 2])
    r = np.sqrt((gx - c[0]) ** 2 + (gy - c[1]) ** 2 + (gz - c[2]) ** 2)
    state[r < 0.38 * min(X, Z)] = 1
    box = (np.abs(gx - c[0]) < 10) & (gy < 25) & (np.abs(gz - c[2]) < 15)
    inner = (np.abs(gx - c[0]) < 9) & (gy < 24) & (np.abs(gz - c[2]) < 14)
    state[box & ~inner] = 2
    state[(gy <= 1) & (state == 1)] = 2
    speck = rng.rand(X, Y, Z) < 0.002
    state[speck] = 2
    state[0, 0, 0] = 1
    nray = np.zeros_like(state)
    nray[state == 2] = rng.randint(0, 40, size=int((state == 2).sum())).astype(np.float32)
    return state, nray


# ---------------------------------------------------------------------------------------------------
# Execution of the reference's own function bodies (build container only)
# ---------------------------------------------------------------------------------------------------
class _Legacy(torch.Tensor):
    """torch-1.x indexing rule the reference relies on: a 2-D numpy array or a list of lists used as an
    index means "one index array per dimension" (tsdf_gpu.py:256-259, :273, :284, :291, :298-300)."""

    @staticmethod
    def _fix(idx):
        if isinstance(idx, np.ndarray) and idx.ndim == 2:
            return tuple(torch.from_numpy(np.ascontiguousarray(r)).long() for r in idx)
        if isinstance(idx, list) and len(idx) > 0 and isinstance(idx[0], (list, np.ndarray)):
            return tuple(torch.as_tensor(np.asarray(r)).long() for r in idx)
        return idx

    def __getitem__(self, idx):
        return super().__getitem__(_Legacy._fix(idx))

    def __setitem__(self, idx, val):
        return super().__setitem__(_Legacy._fix(idx), val)


def _legacy(t):
    return t.as_subclass(_Legacy)


def reference_available():
    return os.path.exists(REF_FILE)


def run_reference_render_rays(rays_o, rays_d, state, nray, origin, voxel=VOXEL, near=NEAR, far=FAR, aabb=AABB_CABIN):
    """Runs the unmodified source of ``isectSegAABB_gpu`` and ``TSDF.render_rays`` from the reference."""
    src = open(REF_FILE).read()
    tree = ast.parse(src)
    fn_src = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "isectSegAABB_gpu":
            fn_src[node.name] = ast.get_source_segment(src, node)
        if isinstance(node, ast.ClassDef) and node.name == "TSDF":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == "render_rays":
                    fn_src[sub.name] = ast.get_source_segment(src, sub)
    import textwrap

    class _TorchShim:
        """torch, with factory functions returning legacy-indexing tensors."""
        def __getattr__(self, name):
            f = getattr(torch, name)
            if name in ("zeros", "ones", "tensor", "cat", "clone", "stack"):
                return lambda *a, **k: _legacy(f(*a, **k))
            return f

    ns = {"torch": _TorchShim(), "np": np, "time": time, "math": math, "device": torch.device("cpu"),
          "N_samples": int((far - near) / voxel) + 1, "near": near, "far": far, "aabb_bnds": np.asarray(aabb)}
    exec(fn_src["isectSegAABB_gpu"], ns)
    exec(textwrap.dedent(fn_src["render_rays"]), ns)
    self = types.SimpleNamespace(
        vol_origin=_legacy(torch.from_numpy(np.asarray(origin, dtype=np.float32))),
        voxel_size=_legacy(torch.tensor(float(voxel))),
        index_bnd=np.asarray(state.shape),
        State_vol_origin=_legacy(state.clone()), Nray_vol=_legacy(nray.clone()),
        Color_vol=_legacy(torch.ones(*state.shape, 3) * 255))
    n = rays_o.reshape(-1, 3).shape[0]
    batch = _legacy(torch.cat([rays_o.reshape(-1, 3).float(), rays_d.reshape(-1, 3).float(),
                               torch.full((n, 1), near), torch.full((n, 1), far)], -1))
    ret = ns["render_rays"](self, batch)
    return torch.Tensor(ret["uncer_map"]).clone(), torch.Tensor(ret["depth_map"]).clone()
