"""Drop-in for the renderer seam of the reference: same names, argument meaning and return
structure as nerfServer/core/run_nerf.py (render, :131-196; render_path, :199-241) and run_nerf_helpers.py
(get_rays, :173-182), backed by the CUDA engine instead of PyTorch ops.

Supported configuration = what the server runs at inference (nerf.py:727-752): ndc=False,
use_viewdirs=True, perturb=0, raw_noise_std=0.  Anything else raises (no silent fallback).
"""
import hashlib
import os

import numpy as np
import torch

from .engine import Engine

_ENGINE_CACHE = {}             # insertion-ordered: least recently used first
_FINGERPRINT_MEMO = {}
_ENGINE_CACHE_MAX = 2          # e.g. one snapshot per scoring device of the reference's uncertainty_kargs list


def get_pose(location, u, v):
    """c2w from position + pitch u + yaw v; same matrix as nerfServer/core/views.py:141-150."""
    su, cu, sv, cv = np.sin(u), np.cos(u), np.sin(v), np.cos(v)
    return [[cv, sv * su, -sv * cu, location[0]],
            [0, cu, su, location[1]],
            [-sv, cv * su, -cv * cu, location[2]],
            [0, 0, 0, 1]]


def get_poses(locations, us, vs):
    """get_pose for many views at once: float64 [V,4,4], element for element what get_pose returns (the same NumPy
    sin / cos loops run on arrays instead of scalars; the CPU test suite checks the equality), without the
    per-view Python overhead (16 us per pose, 25 ms for the 1500 stage-1 views of a planning step)."""
    loc = np.asarray(locations, dtype=np.float64).reshape(-1, 3)
    u, v = np.asarray(us, dtype=np.float64).reshape(-1), np.asarray(vs, dtype=np.float64).reshape(-1)
    su, cu, sv, cv = np.sin(u), np.cos(u), np.sin(v), np.cos(v)
    P = np.zeros((len(u), 4, 4))
    P[:, 0, 0], P[:, 0, 1], P[:, 0, 2], P[:, 0, 3] = cv, sv * su, -sv * cu, loc[:, 0]
    P[:, 1, 1], P[:, 1, 2], P[:, 1, 3] = cu, su, loc[:, 1]
    P[:, 2, 0], P[:, 2, 1], P[:, 2, 2], P[:, 2, 3] = -sv, cv * su, -cv * cu, loc[:, 2]
    P[:, 3, 3] = 1
    return P


def _intrinsics(K):
    return float(np.float32(K[0][0])), float(np.float32(K[1][1])), float(np.float32(K[0][2])), float(np.float32(K[1][2]))


def make_render_kwargs(network_fn, network_fine, N_samples=64, N_importance=128, white_bkgd=True, lindisp=False,
                       device=0, precision="fp16", near=0.5, far=80.0, max_chunk_rays=0):
    """Counterpart of create_nerf's render_kwargs_test (run_nerf.py:316-333) / copy_model's kwargs
    (nerf.py:727-752): snapshots both networks into an Engine and returns the kwargs dict that
    ``render(**kwargs)`` consumes.  network_fn / network_fine: reference NeRF modules or state_dicts."""
    eng = Engine(device=device, n_samples=N_samples, n_importance=N_importance, near=near, far=far,
                 white_bkgd=white_bkgd, lindisp=lindisp, precision=precision, max_chunk_rays=max_chunk_rays)
    eng.load_weights(0, network_fn)
    eng.load_weights(1, network_fine if network_fine is not None else network_fn)
    return {"network_query_fn": None, "perturb": False, "N_importance": N_importance, "network_fine": network_fine,
            "N_samples": N_samples, "network_fn": network_fn, "use_viewdirs": True, "white_bkgd": white_bkgd,
            "raw_noise_std": 0., "device": eng.device, "ndc": False, "lindisp": lindisp, "engine": eng}


def _weights_fingerprint(net):
    """Content hash of a network's parameters (module or state_dict): the cache must notice new weights behind a
    recycled id() and must not pin one engine per deep copy of the same weights (copy_model, nerf.py:718-757,
    deep-copies the training networks at every hand-off)."""
    if net is None:
        return None
    sd = net.state_dict() if hasattr(net, "state_dict") else net
    # unchanged tensors (same storage, same in-place version counters) -> the fingerprint computed last time
    quick = tuple((k, v.data_ptr(), v._version) for k, v in sd.items())
    memo = _FINGERPRINT_MEMO.get(id(net))
    if memo is not None and memo[0] == quick:
        return memo[1]
    fp = _hash_state_dict(sd)
    if len(_FINGERPRINT_MEMO) > 16:
        _FINGERPRINT_MEMO.clear()
    _FINGERPRINT_MEMO[id(net)] = (quick, fp)
    return fp


def _hash_state_dict(sd):
    h = hashlib.blake2b(digest_size=16)
    for k in sorted(sd.keys()):
        t = sd[k].detach().to("cpu", torch.float32).contiguous()
        h.update(k.encode())
        h.update(t.numpy().tobytes())
    return h.hexdigest()


def _engine_from_kwargs(kwargs, near, far):
    eng = kwargs.get("engine")
    if eng is not None:
        return eng
    net, fine = kwargs.get("network_fn"), kwargs.get("network_fine")
    if net is None:
        raise ValueError("render(): kwargs need 'engine' (make_render_kwargs) or 'network_fn'")
    dev = kwargs.get("device", 0)
    dev = dev.index if isinstance(dev, torch.device) else dev
    key = (_weights_fingerprint(net), _weights_fingerprint(fine), kwargs.get("N_samples", 64), kwargs.get("N_importance", 128),
           bool(kwargs.get("white_bkgd", False)), bool(kwargs.get("lindisp", False)), dev)
    eng = _ENGINE_CACHE.get(key)
    if eng is None:
        # new weights replace the engine that held this configuration's previous snapshot: at most
        # _ENGINE_CACHE_MAX engines (0.7 GB of workspace each) stay alive, least recently used first out
        while len(_ENGINE_CACHE) >= _ENGINE_CACHE_MAX:
            _ENGINE_CACHE.pop(next(iter(_ENGINE_CACHE))).close()
        kw = make_render_kwargs(net, fine, N_samples=key[2], N_importance=key[3], white_bkgd=key[4], lindisp=key[5],
                                device=dev or 0, near=near, far=far)
        eng = kw["engine"]
    else:
        del _ENGINE_CACHE[key]          # re-insert: most recently used last
    _ENGINE_CACHE[key] = eng
    return eng


def get_rays(H, W, K, c2w, dev=None, engine=None):
    """Same contract as run_nerf_helpers.get_rays: returns rays_o, rays_d of shape [H,W,3]."""
    if engine is None:
        raise ValueError("get_rays needs engine= (the CUDA engine that generates the rays)")
    c2w = torch.as_tensor(np.asarray(c2w.cpu() if torch.is_tensor(c2w) else c2w), dtype=torch.float32)[:3, :4]
    fx, fy, cx, cy = _intrinsics(K)
    o, d, _ = engine.get_rays(c2w[None].contiguous(), H, W, fx, fy, cx, cy)
    return o.reshape(H, W, 3), d.reshape(H, W, 3)


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, dev=None, **kwargs):
    """Render rays.  Signature and return value of run_nerf.render (run_nerf.py:131-196):
    returns [rgb_map, disp_map, acc_map, uncertainty_map, depth_map, extras_dict] with extras
    holding raw (if retraw), rgb0, disp0, acc0, uncertainty_map0, depth_map0, z_std.
    ``chunk`` is accepted for compatibility; the engine chunks internally."""
    if ndc:
        raise NotImplementedError("ndc=True (LLFF forward-facing) is not on the scoring path (nerf.py:745 sets ndc=False)")
    if not use_viewdirs:
        raise NotImplementedError("use_viewdirs=False: the reference's networks are built with use_viewdirs=True (config.txt:9)")
    if kwargs.get("perturb", 0.) not in (0., False) or kwargs.get("raw_noise_std", 0.) not in (0., False):
        raise NotImplementedError("perturb / raw_noise_std are training-time options (nerf.py:750-751 zeroes them)")
    if c2w_staticcam is not None:
        raise NotImplementedError("c2w_staticcam is a visualisation-only option")
    eng = _engine_from_kwargs(kwargs, near, far)
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w, engine=eng)
    else:
        rays_o, rays_d = rays
    sh = tuple(rays_d.shape)
    want = ["rgb", "disp", "acc", "depth", "unc"]
    if eng.n_importance > 0:
        want += ["rgb0", "disp0", "acc0", "depth0", "unc0", "z_std"]
    if kwargs.get("retraw", False):
        want.append("raw")
    o = eng.render_rays(rays_o, rays_d, near, far, precision=kwargs.get("precision"), want=tuple(want))

    def rs(t):
        return t.reshape(sh[:-1] + tuple(t.shape[1:]))

    extras = {}
    if "raw" in o:
        extras["raw"] = rs(o["raw"])
    if eng.n_importance > 0:
        extras.update({"rgb0": rs(o["rgb0"]), "disp0": rs(o["disp0"]), "acc0": rs(o["acc0"]),
                       "uncertainty_map0": rs(o["unc0"]), "depth_map0": rs(o["depth0"]), "z_std": rs(o["z_std"])})
    return [rs(o["rgb"]), rs(o["disp"]), rs(o["acc"]), rs(o["unc"]), rs(o["depth"]), extras]


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0, dev=None):
    """Full-image renders of a list of poses: signature and return value of run_nerf.render_path (run_nerf.py:199-241,
    the caller of BASELINE configs[4]): returns (rgbs, disps, uncers, depths), four lists of per-pose numpy arrays
    ([H,W,3], [H,W], [H,W,S_f], [H,W]).  ``render_factor`` downsamples H, W and the focal length like the reference
    (K is used as given, as there); with ``savedir`` every rgb map is also written as ``%03d.png`` (8-bit, the
    reference's to8b).  ``gt_imgs`` is unused there too (its PSNR print is commented out)."""
    H, W, focal = hwf
    if render_factor != 0:
        H = H // render_factor
        W = W // render_factor
        focal = focal / render_factor
    rgbs, disps, uncers, depths = [], [], [], []
    for i, c2w in enumerate(render_poses):
        c2w = torch.as_tensor(np.asarray(c2w.cpu() if torch.is_tensor(c2w) else c2w), dtype=torch.float32)
        rgb, disp, acc, uncer, depth, _ = render(H, W, K, chunk=chunk, c2w=c2w[:3, :4], dev=dev, **render_kwargs)
        rgbs.append(rgb.cpu().numpy())
        disps.append(disp.cpu().numpy())
        uncers.append(uncer.cpu().numpy())
        depths.append(depth.cpu().numpy())
        if savedir is not None:
            import cv2
            rgb8 = (255 * np.clip(rgbs[-1], 0, 1)).astype(np.uint8)
            cv2.imwrite(os.path.join(savedir, "{:03d}.png".format(i)), rgb8[..., ::-1])
    return rgbs, disps, uncers, depths
