"""CPU oracle for EVPP's candidate-view scoring path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain PyTorch-CPU fp32 ops, of the reference's
algorithm for the hot path (pose -> rays -> stratified + hierarchical sampling
-> NeRF+uncertainty MLP -> alpha compositing -> per-view score -> next-best
view).  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product (``evpp_b200``) never does, and fails loudly without its CUDA library.

Every function cites the reference file:line it follows (paths are relative
to the reference root, e.g. ``nerfServer/core/run_nerf.py``).  It deliberately
uses the *same ATen ops in the same order* as the reference so that on one
machine its outputs are bit-identical to the reference's; that equality is
pinned by ``oracle/make_golden.py`` (run where ``/root/reference`` exists) and
by the committed fixtures under ``tests/golden/``.

Parity status: the reference has no tests or golden vectors for this path
(SURVEY.md section 4), so the pin is "outputs of the reference itself run
here": tests/golden/*.npz were produced by importing the unmodified reference
modules (oracle/ref_import.py) and this oracle is checked against them.

The NGP-mode stages (occupancy bitfield marching, hash-grid encoding, small
MLP) at the bottom of this file have NO counterpart in the reference
("parity unpinned"): they restate the public
torch-ngp semantics and only the NeurAR head / score / compositing parts are
pinned by the reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# live constants of the reference
# ----------------------------------------------------------------------------
N_SAMPLES = 64        # nerfServer/config.txt:10
N_IMPORTANCE = 128    # nerfServer/config.txt:11
MULTIRES = 10         # nerfServer/core/nerf_configs.py (multires)
MULTIRES_VIEWS = 4    # nerfServer/core/nerf_configs.py (multires_views)
NETDEPTH = 8
NETWIDTH = 256
SKIPS = (4,)
CHUNK = 1024 * 32     # nerf_configs.py chunk
NETCHUNK = 1024 * 64  # nerf_configs.py netchunk


# ----------------------------------------------------------------------------
# a1: pose
# ----------------------------------------------------------------------------
def get_pose(location, u, v):
    """nerfServer/core/views.py:141-150 (twin: plannerServer_*/core/tsdf_gpu.py:149-157)."""
    sx = np.sin(u)
    cx = np.cos(u)
    sy = np.sin(v)
    cy = np.cos(v)
    return [[cy, sy * sx, -sy * cx, location[0]],
            [0, cx, sx, location[1]],
            [-sy, cy * sx, -cy * cx, location[2]],
            [0, 0, 0, 1]]


def look_at_uv(position, target):
    """plannerServer_Object/core/interface.py:71-73 / vpp.py:150-153."""
    dx, dy, dz = target[0] - position[0], target[1] - position[1], target[2] - position[2]
    u = np.arctan2(-dy, np.sqrt(dx ** 2 + dz ** 2))
    v = np.arctan2(dx, dz)
    return u, v


def make_K(H, W, focal):
    """nerfServer/core/nerf.py:112-116: K = [[f,0,W/2],[0,f,H/2],[0,0,1]] (numpy fp64)."""
    return np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])


# ----------------------------------------------------------------------------
# a2: rays
# ----------------------------------------------------------------------------
def get_rays(H, W, K, c2w):
    """nerfServer/core/run_nerf_helpers.py:173-182 (c2w: fp32 tensor [3,4])."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = i.t()
    j = j.t()
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


# ----------------------------------------------------------------------------
# a6: positional encoding
# ----------------------------------------------------------------------------
def embed(x, multires):
    """nerfServer/core/run_nerf_helpers.py:22-55: cat[x, sin(x f0), cos(x f0), ...], f=2**linspace(0,L-1,L)."""
    freq_bands = 2. ** torch.linspace(0., multires - 1, steps=multires)
    out = [x]
    for freq in freq_bands:
        for p_fn in (torch.sin, torch.cos):
            out.append(p_fn(x * freq))
    return torch.cat(out, -1)


# ----------------------------------------------------------------------------
# a7: the NeRF + NeurAR-uncertainty MLP, on a plain state_dict
# ----------------------------------------------------------------------------
def nerf_state_dict_keys(D=NETDEPTH):
    keys = []
    for i in range(D):
        keys += [f"pts_linears.{i}.weight", f"pts_linears.{i}.bias"]
    for n in ("views_linears.0", "feature_linear", "alpha_linear", "rgb_linear", "uncertainty_linear"):
        keys += [n + ".weight", n + ".bias"]
    return keys


def init_nerf_state_dict(seed=None, D=NETDEPTH, W=NETWIDTH, input_ch=63, input_ch_views=27, skips=SKIPS):
    """Random-init weights exactly as the reference's ``NeRF.__init__`` would draw them
    (nerfServer/core/run_nerf_helpers.py:79-106): nn.Linear default init, layers
    constructed in the same order so one ``torch.manual_seed`` reproduces them."""
    if seed is not None:
        torch.manual_seed(seed)
    lin = torch.nn.Linear
    pts = [lin(input_ch, W)] + [lin(W, W) if i not in skips else lin(W + input_ch, W) for i in range(D - 1)]
    views = lin(input_ch_views + W, W // 2)
    feature = lin(W, W)
    alpha = lin(W, 1)
    rgb = lin(W // 2, 3)
    unc = lin(W // 2, 1)
    sd = OrderedDict()
    for i, l in enumerate(pts):
        sd[f"pts_linears.{i}.weight"] = l.weight.detach().clone()
        sd[f"pts_linears.{i}.bias"] = l.bias.detach().clone()
    for name, l in (("views_linears.0", views), ("feature_linear", feature), ("alpha_linear", alpha),
                    ("rgb_linear", rgb), ("uncertainty_linear", unc)):
        sd[name + ".weight"] = l.weight.detach().clone()
        sd[name + ".bias"] = l.bias.detach().clone()
    return sd


def spread_variant(sd):
    """SURVEY.md section 8(d) config-1 'spread' variant: scores of different views separate."""
    sd = OrderedDict((k, v.clone()) for k, v in sd.items())
    sd["uncertainty_linear.weight"] *= 8
    sd["alpha_linear.bias"] += 0.5
    return sd


def sparse_variant(sd, k=2000.0, q=0.03):
    """Random-init weights reshaped into a scene with empty space and solid blobs: sigma_raw' = k * (sigma_raw - q)
    (random-init sigma_raw spans about +-0.04, so ~2 % of space is dense) and rgb logits x 30 (saturated colours).
    Rays then miss everything (white background), stop early or stop late: all three masks of the depth /
    surface queries (nerf.py:344-350, 386-391) get exercised."""
    sd = {k_: v.clone() for k_, v in sd.items()}
    sd["alpha_linear.weight"] = sd["alpha_linear.weight"] * k
    sd["alpha_linear.bias"] = sd["alpha_linear.bias"] * k - k * q
    sd["rgb_linear.weight"] = sd["rgb_linear.weight"] * 30
    return sd


def nerf_forward(sd, x, input_ch=63, input_ch_views=27, skips=SKIPS, D=NETDEPTH):
    """nerfServer/core/run_nerf_helpers.py:108-134 (use_viewdirs=True branch)."""
    input_pts, input_views = torch.split(x, [input_ch, input_ch_views], dim=-1)
    h = input_pts
    for i in range(D):
        h = F.linear(h, sd[f"pts_linears.{i}.weight"], sd[f"pts_linears.{i}.bias"])
        h = F.relu(h)
        if i in skips:
            h = torch.cat([input_pts, h], -1)
    alpha = F.linear(h, sd["alpha_linear.weight"], sd["alpha_linear.bias"])
    feature = F.linear(h, sd["feature_linear.weight"], sd["feature_linear.bias"])
    h = torch.cat([feature, input_views], -1)
    h = F.relu(F.linear(h, sd["views_linears.0.weight"], sd["views_linears.0.bias"]))
    rgb = F.linear(h, sd["rgb_linear.weight"], sd["rgb_linear.bias"])
    uncertainty = torch.exp(F.linear(h, sd["uncertainty_linear.weight"], sd["uncertainty_linear.bias"]))
    return torch.cat([rgb, alpha, uncertainty], -1)


def run_network(sd, inputs, viewdirs, netchunk=NETCHUNK):
    """nerfServer/core/run_nerf.py:84-113 + batchify :73-81."""
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed(inputs_flat, MULTIRES)
    input_dirs = viewdirs[:, None].expand(inputs.shape)
    input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
    embedded = torch.cat([embedded, embed(input_dirs_flat, MULTIRES_VIEWS)], -1)
    outs = [nerf_forward(sd, embedded[i:i + netchunk]) for i in range(0, embedded.shape[0], netchunk)]
    outputs_flat = torch.cat(outs, 0)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


# ----------------------------------------------------------------------------
# a8: compositing
# ----------------------------------------------------------------------------
def raw2outputs(raw, z_vals, rays_d, white_bkgd=False):
    """nerfServer/core/run_nerf.py:344-388 with raw_noise_std=0."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.Tensor([1e10]).expand(dists[..., :1].shape)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = 1. - torch.exp(-F.relu(raw[..., 3]) * dists)
    weights = alpha * torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / torch.sum(weights, -1))
    acc_map = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map


# ----------------------------------------------------------------------------
# a9: hierarchical resampling (deterministic branch)
# ----------------------------------------------------------------------------
def sample_pdf(bins, weights, N_samples, return_stages=False):
    """nerfServer/core/run_nerf_helpers.py:216-259 with det=True, pytest=False."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0., 1., steps=N_samples)
    u = u.expand(list(cdf.shape[:-1]) + [N_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], -1)
    matched_shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(matched_shape), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(matched_shape), 2, inds_g)
    denom = (cdf_g[..., 1] - cdf_g[..., 0])
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    samples = bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])
    if return_stages:
        return samples, {"cdf": cdf, "inds": inds, "below": below, "above": above}
    return samples


# ----------------------------------------------------------------------------
# a5 + a10: render_rays, a3/a4: render
# ----------------------------------------------------------------------------
def render_rays(ray_batch, sd_coarse, sd_fine, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE,
                white_bkgd=True, lindisp=False, retraw=True, stages=None):
    """nerfServer/core/run_nerf.py:391-509 with perturb=0, raw_noise_std=0 (inference, nerf.py:750-751)."""
    N_rays = ray_batch.shape[0]
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:]
    bounds = torch.reshape(ray_batch[..., 6:8], [-1, 1, 2])
    near, far = bounds[..., 0], bounds[..., 1]
    t_vals = torch.linspace(0., 1., steps=N_samples)
    if not lindisp:
        z_vals = near * (1. - t_vals) + far * (t_vals)
    else:
        z_vals = 1. / (1. / near * (1. - t_vals) + 1. / far * (t_vals))
    z_vals = z_vals.expand([N_rays, N_samples])
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    raw = run_network(sd_coarse, pts, viewdirs)
    uncertainty_map = raw[..., 4]
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, white_bkgd)
    if stages is not None:
        stages.setdefault("z_coarse", []).append(z_vals.clone())
        stages.setdefault("raw_coarse", []).append(raw.clone())
        stages.setdefault("weights_coarse", []).append(weights.clone())
    ret0 = None
    if N_importance > 0:
        ret0 = (rgb_map, disp_map, acc_map, uncertainty_map, depth_map)
        z_vals_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        z_samples, st = sample_pdf(z_vals_mid, weights[..., 1:-1], N_importance, return_stages=True)
        z_samples = z_samples.detach()
        z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        raw = run_network(sd_fine if sd_fine is not None else sd_coarse, pts, viewdirs)
        uncertainty_map = raw[..., 4]
        rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, white_bkgd)
        if stages is not None:
            for k, v in st.items():
                stages.setdefault(k, []).append(v.clone())
            stages.setdefault("z_samples", []).append(z_samples.clone())
            stages.setdefault("z_fine", []).append(z_vals.clone())
            stages.setdefault("weights_fine", []).append(weights.clone())
    ret = {"rgb_map": rgb_map, "disp_map": disp_map, "acc_map": acc_map,
           "uncertainty_map": uncertainty_map, "depth_map": depth_map}
    if retraw:
        ret["raw"] = raw
    if N_importance > 0:
        ret["rgb0"], ret["disp0"], ret["acc0"], ret["uncertainty_map0"], ret["depth_map0"] = ret0
        ret["z_std"] = torch.std(z_samples, dim=-1, unbiased=False)
    return ret


def render(H, W, K, chunk=CHUNK, rays=None, c2w=None, near=0., far=1., sd_coarse=None, sd_fine=None,
           N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, white_bkgd=True, lindisp=False, retraw=True,
           stages=None):
    """nerfServer/core/run_nerf.py:131-196 with ndc=False, use_viewdirs=True, c2w_staticcam=None;
    batchify_rays :116-128."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    viewdirs = rays_d
    viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
    viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    sh = rays_d.shape
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near_t, far_t = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    rays_cat = torch.cat([rays_o, rays_d, near_t, far_t, viewdirs], -1)
    all_ret = {}
    for i in range(0, rays_cat.shape[0], chunk):
        ret = render_rays(rays_cat[i:i + chunk], sd_coarse, sd_fine, N_samples=N_samples,
                          N_importance=N_importance, white_bkgd=white_bkgd, lindisp=lindisp,
                          retraw=retraw, stages=stages)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    all_ret = {k: torch.cat(all_ret[k], 0) for k in all_ret}
    for k in all_ret:
        k_sh = list(sh[:-1]) + list(all_ret[k].shape[1:])
        all_ret[k] = torch.reshape(all_ret[k], k_sh)
    k_extract = ["rgb_map", "disp_map", "acc_map", "uncertainty_map", "depth_map"]
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]


# ----------------------------------------------------------------------------
# a11: per-view score
# ----------------------------------------------------------------------------
def build_view_rays(poses, H, W, K, pix_idx=None):
    """nerfServer/core/nerf.py:278-296: per pose full-grid get_rays then a row subset.
    The reference draws the subset with an unseeded np.random.choice (nerf.py:287);
    the harness passes explicit flat pixel indices ``pix_idx [V,R]`` (row-major j*W+i,
    i.e. ``coords`` order of nerf.py:141-142) or None for the full grid."""
    poses = torch.Tensor(np.asarray(poses, dtype=np.float64)) if not torch.is_tensor(poses) else poses
    poses = poses[:, :3, :4]
    all_o, all_d = [], []
    for v, pose in enumerate(poses):
        rays_o, rays_d = get_rays(H, W, K, pose)
        rays_o = rays_o.reshape(-1, 3)
        rays_d = rays_d.reshape(-1, 3)
        if pix_idx is not None:
            sel = torch.as_tensor(np.asarray(pix_idx[v]), dtype=torch.long)
            rays_o, rays_d = rays_o[sel], rays_d[sel]
        all_o.append(rays_o)
        all_d.append(rays_d)
    return torch.stack([torch.cat(all_o, 0), torch.cat(all_d, 0)], 0)


def score_views(poses, H, W, K, near, far, sd_coarse, sd_fine, pix_idx=None, chunk=CHUNK,
                N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, white_bkgd=True, stages=None):
    """nerfServer/core/nerf.py:266-309: uncers[v] = sum_{rays,samples} beta^2 / rays_per_view,
    beta = fine-pass raw[...,4] (un-composited)."""
    V = len(poses)
    batch_rays = build_view_rays(poses, H, W, K, pix_idx)
    R = batch_rays.shape[1] // V
    S_f = N_samples + N_importance
    with torch.no_grad():
        rgb, disp, acc, uncertainty, depth, extras = render(
            H, W, K, chunk=chunk, rays=batch_rays, near=near, far=far, sd_coarse=sd_coarse, sd_fine=sd_fine,
            N_samples=N_samples, N_importance=N_importance, white_bkgd=white_bkgd, retraw=True, stages=stages)
        uncertainty = uncertainty ** 2
        uncers = uncertainty.reshape((V, R * S_f))
        uncers = torch.sum(uncers, -1) / R
    return uncers


# ----------------------------------------------------------------------------
# f2: depth / surface queries through the same renderer
# ----------------------------------------------------------------------------
def sphere_dirs(step):
    """Ray fan of Controller.get_surface_points (nerfServer/core/nerf.py:332-336)."""
    rays_d = []
    for u in np.arange(0, 2 * np.pi, step):
        for v in np.arange(0, np.pi, step):
            rays_d.append([np.cos(u), np.sin(u) * np.cos(v), np.sin(u) * np.sin(v)])
    return torch.Tensor(rays_d)


def get_surface_points(location, radius, step, H, W, K, near, far, sd_coarse, sd_fine, chunk=CHUNK,
                       return_all=False):
    """nerfServer/core/nerf.py:325-351: sphere of rays from ``location``; white pixels (mean(rgb)*255 > 250) get
    depth 0; keep the points with near < depth < radius."""
    rays_d = sphere_dirs(step)
    location = torch.Tensor(location)
    rays_o = location.expand([rays_d.shape[0], rays_d.shape[1]])
    batch_rays = torch.stack([rays_o, rays_d], 0)
    with torch.no_grad():
        rgb, disp, acc, uncertainty, depth, extras = render(H, W, K, chunk=chunk, rays=batch_rays, near=near, far=far,
                                                            sd_coarse=sd_coarse, sd_fine=sd_fine, retraw=True)
        depth_raw = depth.clone()
        rgbs = torch.mean(rgb, -1) * 255
        depth[rgbs > 250] = 0
        white = rgbs > 250
        points = rays_o + rays_d * depth[..., None]
        keep = (depth < radius) & (depth > near)
        points = points[depth < radius]
        depth = depth[depth < radius]
        points = points[depth > near]
    if return_all:
        return points, {"depth_raw": depth_raw, "white": white, "keep": keep, "rgb": rgb}
    return points


def get_rays_depth(locations, dirs, near, far, sd_coarse, sd_fine, chunk=CHUNK, return_all=False):
    """nerfServer/core/nerf.py:353-392: depth along caller rays through batchify_rays on a hand-built [N,11] batch
    whose view directions are the RAW ``dirs`` (not normalised: ``rays = torch.cat([rays, dirs], -1)``, :365);
    white pixels and depths below near are zeroed."""
    inputs = torch.Tensor(locations)
    dirs = torch.Tensor(dirs)
    with torch.no_grad():
        sh = dirs.shape
        rays_o = torch.reshape(inputs, [-1, 3]).float()
        rays_d = torch.reshape(dirs, [-1, 3]).float()
        near_t, far_t = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
        rays = torch.cat([rays_o, rays_d, near_t, far_t], -1)
        rays = torch.cat([rays, dirs.reshape(-1, 3)], -1)
        all_ret = {}
        for i in range(0, rays.shape[0], chunk):
            ret = render_rays(rays[i:i + chunk], sd_coarse, sd_fine, retraw=True)
            for k in ret:
                all_ret.setdefault(k, []).append(ret[k])
        all_ret = {k: torch.cat(all_ret[k], 0) for k in all_ret}
        for k in all_ret:
            all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
        rgb_map, depth_map = all_ret["rgb_map"], all_ret["depth_map"]
        depth_raw = depth_map.clone()
        rgbs = torch.mean(rgb_map, -1) * 255
        depth_map[rgbs > 250] = 0
    depth_map[depth_map < near] = 0
    if return_all:
        return depth_map, {"depth_raw": depth_raw, "white": rgbs > 250, "rgb": rgb_map}
    return depth_map


# ----------------------------------------------------------------------------
# a12: next-best-view selection
# ----------------------------------------------------------------------------
def depth_clip_object(select_depth):
    """plannerServer_Object/core/vpp.py:193-196."""
    select_depth = np.asarray(select_depth, dtype=np.float64)
    depth_clip = np.ones(select_depth.shape)
    depth_clip[select_depth < 1.5] = np.exp(-2 * np.abs(select_depth[select_depth < 1.5] - 3.0))
    depth_clip[select_depth > 4.5] = np.exp(-2 * np.abs(select_depth[select_depth > 4.5] - 3.0))
    return depth_clip


def select_nbv(views, uncers, dirs_per_location=3, depth_clip=None):
    """plannerServer_Object/core/vpp.py:199-210 + get_maxdirdata :214-220 + interface.py:115-116.
    views [V,5] = (x,y,z,u,v), consecutive groups of ``dirs_per_location`` share a location.
    Returns (data [L,6] normalised, nbv_row, index of the winning view in ``views``).
    Exact ties: the reference sorts with numpy's default (unstable) argsort and takes the
    last row, so its winner among exact ties is unspecified; here: highest index."""
    views = np.asarray(views, dtype=np.float64)
    uncer = np.asarray(uncers, dtype=np.float64)
    if depth_clip is not None:
        uncer = uncer * depth_clip
    data_buff = np.concatenate((views[:, 0:5], uncer.reshape(-1, 1)), axis=1)
    L = data_buff.shape[0] // dirs_per_location
    maxdir, maxidx = [], []
    for i in range(L):
        dir_buff = data_buff[i * dirs_per_location:(i + 1) * dirs_per_location]
        k = int(np.argmax(dir_buff[:, 5]))
        maxdir.append(list(dir_buff[k]))
        maxidx.append(i * dirs_per_location + k)
    data = np.array(maxdir)
    data[:, 5] = data[:, 5] / np.max(data[:, 5])
    order = np.argsort(data[:, 5], kind="stable")
    win = int(order[-1])
    return data, data[win], maxidx[win]


# ----------------------------------------------------------------------------
# synthetic workloads
# ----------------------------------------------------------------------------
def circle_views(V=32, r=3.0, y=1.0, center=(0.0, 1.0, 0.0)):
    """Config 1: V poses on a circle radius r at height y, looking at ``center``."""
    views = []
    for k in range(V):
        a = 2 * np.pi * k / V
        pos = (center[0] + r * np.cos(a), y, center[2] + r * np.sin(a))
        u, v = look_at_uv(pos, center)
        views.append([pos[0], pos[1], pos[2], u, v])
    return np.array(views)


def hemisphere_views(V=512, center=(0.0, 1.0, 0.0), rmin=3.0, rmax=4.5, seed=0):
    """Config 2: Fibonacci upper-hemisphere positions, radius in [rmin,rmax], look-at + the
    planner's +-5 deg / +-10 deg jitters (plannerServer_Object/core/vpp.py:154-155)."""
    rng = np.random.RandomState(seed)
    views = []
    ga = np.pi * (3.0 - np.sqrt(5.0))
    us = np.linspace(-5 / 180 * np.pi, 5 / 180 * np.pi, 3)
    vs = np.linspace(-10 / 180 * np.pi, 10 / 180 * np.pi, 5)
    for k in range(V):
        yy = (k + 0.5) / V
        rad = np.sqrt(max(0.0, 1 - yy * yy))
        th = ga * k
        r = rmin + (rmax - rmin) * rng.rand()
        pos = (center[0] + r * rad * np.cos(th), center[1] + r * yy, center[2] + r * rad * np.sin(th))
        u, v = look_at_uv(pos, center)
        u = float(np.clip(u + us[rng.randint(3)], -np.pi / 2, np.pi / 2))
        v = v + vs[rng.randint(5)]
        views.append([pos[0], pos[1], pos[2], u, v])
    return np.array(views)


def room_views(V=4096, seed=0, aoi=((-1.5, 1.0), (0.7, 1.9), (0.0, 2.5)), dirs_per_loc=4):
    """Config 3: positions uniform in aoi_bnds (plannerServer_Room/core/tsdf_gpu.py:25) x
    ``dirs_per_loc`` directions drawn from the Room fan (plannerServer_Room/core/vpp.py:151-154)."""
    rng = np.random.RandomState(seed)
    L = V // dirs_per_loc
    views = []
    for _ in range(L):
        x = rng.uniform(*aoi[0]); y = rng.uniform(*aoi[1]); z = rng.uniform(*aoi[2])
        dy = 1.0 - y
        u_c = np.arctan2(-dy, 0.6)
        ufan = np.linspace(u_c - 30 / 180 * np.pi, u_c + 30 / 180 * np.pi, 5)
        vfan = np.linspace(0, 2 * np.pi, 12)
        for _d in range(dirs_per_loc):
            u = float(np.clip(ufan[rng.randint(5)], -np.pi / 2, np.pi / 2))
            v = float(vfan[rng.randint(12)])
            views.append([x, y, z, u, v])
    return np.array(views)


def views_to_poses(views):
    return [get_pose(list(v[0:3]), v[3], v[4]) for v in views]
