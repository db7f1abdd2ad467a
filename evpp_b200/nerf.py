"""Drop-in for the scoring seam of the reference's ``Controller`` (nerfServer/core/nerf.py:73):
``get_uncertainty(poses)`` (:266-309), ``get_surface_points`` (:325-351), ``get_rays_depth``
(:353-392) and the weight hand-off ``copy_model`` (:718-757).  The training loop, image store
and HTTP glue of the reference Controller are out of scope."""
import logging

import numpy as np
import torch

from ._lib import EvppError, EVPP_ERR_NONFINITE, PRECISIONS
from .engine import Engine
from .planner import contenders
from .run_nerf import _intrinsics


log = logging.getLogger("evpp_b200")


class Controller:
    def __init__(self, H=400, W=400, focal=300.0, near=0.5, far=80.0, N_samples=64, N_importance=128,
                 white_bkgd=True, device=0, precision="fp16", sample_num=1000, max_chunk_rays=0,
                 pixel_sampler="numpy", seed=0, rescore="auto", rescore_margin=2e-3, rescore_precision="fp16x3"):
        """Defaults are the live constants: views.py:21-23 (H, W, focal), nerf.py:100-101 (near, far),
        config.txt:10-13, nerf.py:276 (sample_num).

        rescore: how ``get_uncertainty`` protects the planner's arg-max (interface.py:115-116) from the bulk pass's
        16-bit rounding.  "auto" (default): every view whose fast score is within ``rescore_margin`` (relative) of
        the best one is scored again on the tensor-core exact path (``rescore_precision``, fp16x3: fp32-class) and
        that score is returned for it; a view outside the band carries <= 1e-3 relative error (the bulk pass's
        tolerance) and cannot overtake the leader, so the arg-max of the returned vector is the fp32 reference's.
        "all": every view on the exact path (needed when the caller weights the scores before its arg-max by factors
        this call cannot see, e.g. the Object planner's depth clip, vpp.py:193-200 -- or use
        ``planner.score_and_select``, which knows the weights).  None / "off": the bulk pass only."""
        self.H, self.W, self.focal = H, W, focal
        self.K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])   # nerf.py:112-116
        self.near, self.far = near, far
        self.sample_num = sample_num
        # "numpy": the reference's protocol, one np.random.choice(H*W, sample_num, replace=False) per pose from NumPy's
        # global RNG (nerf.py:287) -- same pixels as the reference for the same RNG state, 2 ms of host time per pose.
        # "device": distinct pseudo-random pixels drawn by a kernel (evpp_sample_pixels), no host work per pose.
        self.pixel_sampler, self._seed = pixel_sampler, int(seed)
        self.rescore = None if rescore in (None, "off", False) else rescore
        if self.rescore not in (None, "auto", "all"):
            raise ValueError("rescore must be 'auto', 'all' or None")
        self.rescore_margin, self.rescore_precision = float(rescore_margin), rescore_precision
        self.last_rescored = 0            # how many views the last get_uncertainty call re-scored
        self.engine = Engine(device=device, n_samples=N_samples, n_importance=N_importance, near=near, far=far,
                             white_bkgd=white_bkgd, precision=precision, max_chunk_rays=max_chunk_rays)
        self.uncertainty_kargs = None

    def copy_model(self, network_fn, network_fine=None):
        """Snapshot the (training) networks into the scorer; never aliases their parameters.  Accepts reference
        ``NeRF`` modules, state_dicts, or -- as ``network_fn`` alone -- the path of / the dict inside a checkpoint
        written by ``Controller.save_model`` (nerf.py:612-621)."""
        if isinstance(network_fn, (str, bytes)) or (isinstance(network_fn, dict) and "network_fn_state_dict" in network_fn):
            self.engine.load_checkpoint(network_fn)
        else:
            self.engine.load_weights(0, network_fn)
            self.engine.load_weights(1, network_fine if network_fine is not None else network_fn)
        self.uncertainty_kargs = [{"engine": self.engine, "device": self.engine.device}]

    def _pixel_subset(self, V):
        if self.sample_num is None:
            return None
        n = self.H * self.W
        # nerf.py:287 -- one np.random.choice per pose from numpy's global RNG
        return np.stack([np.random.choice(n, size=[self.sample_num], replace=False) for _ in range(V)]).astype(np.int32)

    def _score_host(self, poses, pix_t, precision=None):
        """Bulk pass with HOST poses / pixel ids -> host scores; one retry in bf16 when the fp16 activations of a
        trained checkpoint overflow (EVPP_ERR_NONFINITE), logged."""
        fx, fy, cx, cy = _intrinsics(self.K)
        try:
            return self.engine.score_views_host(poses, self.H, self.W, fx, fy, cx, cy, pix_t, precision=precision)
        except EvppError as e:
            if e.code != EVPP_ERR_NONFINITE or (precision or self.engine.precision) == PRECISIONS["bf16"]:
                raise
            log.warning("evpp_b200: %s -- retrying this call with bf16 operands", e)
            return self.engine.score_views_host(poses, self.H, self.W, fx, fy, cx, cy, pix_t, precision="bf16")

    def get_uncertainty(self, poses, pix_idx=None):
        """poses: list/array [V,4,4] (views.get_pose).  Returns a cuda float32 tensor [V] with
        sum over rays and the 192 fine samples of beta^2, divided by rays per view (nerf.py:307-309)."""
        if self.uncertainty_kargs is None:
            raise TypeError("'NoneType' object is not subscriptable (weights not handed off: call copy_model)")
        poses = torch.Tensor(np.asarray(poses, dtype=np.float64))[:, :3, :4].contiguous()   # fp64 -> fp32, nerf.py:278
        V = poses.shape[0]
        self.last_rescored = 0
        if V == 0:
            return torch.empty(0, device=self.engine.device)
        if pix_idx is None and self.sample_num is not None:
            if self.pixel_sampler == "device":
                self._seed += 1
                pix_idx = self.engine.sample_pixels(V, self.sample_num, self.H * self.W, self._seed).cpu().numpy()
            else:
                pix_idx = self._pixel_subset(V)
        pix_t = None if pix_idx is None else torch.as_tensor(np.ascontiguousarray(pix_idx, dtype=np.int32))
        exact = PRECISIONS[self.rescore_precision]
        if self.rescore == "all" and self.engine.precision != PRECISIONS["fp32"]:
            self.last_rescored = V
            return self._score_host(poses, pix_t, precision=exact).to(self.engine.device)
        scores = self._score_host(poses, pix_t).clone()
        if self.rescore == "auto" and self.engine.precision not in (PRECISIONS["fp32"], exact) and V > 1:
            band = contenders(scores.numpy().astype(np.float64), None, self.rescore_margin)
            fx, fy, cx, cy = _intrinsics(self.K)
            scores[torch.from_numpy(band)] = self.engine.rescore_views(poses, band, self.H, self.W, fx, fy, cx, cy, pix_t,
                                                                       precision=exact)
            self.last_rescored = len(band)
        return scores.to(self.engine.device)

    def _render(self, rays_o, rays_d, viewdirs=None):
        # the depth / surface queries threshold their outputs (white-pixel mask, radius): they run on the exact path
        prec = self.rescore_precision if self.engine.precision != PRECISIONS["fp32"] else "fp32"
        return self.engine.render_rays(rays_o, rays_d, self.near, self.far, precision=prec, want=("rgb", "depth"),
                                       viewdirs=viewdirs)

    def get_surface_points(self, location, radius, step):
        """nerf.py:325-351: sphere of rays from ``location``; keep hits with near < depth < radius."""
        rays_d = []
        for u in np.arange(0, 2 * np.pi, step):
            for v in np.arange(0, np.pi, step):
                rays_d.append([np.cos(u), np.sin(u) * np.cos(v), np.sin(u) * np.sin(v)])
        rays_d = torch.Tensor(rays_d).to(self.engine.device)
        rays_o = torch.Tensor(location).to(self.engine.device).expand(rays_d.shape).contiguous()
        o = self._render(rays_o, rays_d)
        depth = o["depth"].clone()
        depth[torch.mean(o["rgb"], -1) * 255 > 250] = 0
        points = rays_o + rays_d * depth[..., None]
        points = points[depth < radius]
        depth = depth[depth < radius]
        return points[depth > self.near]

    def get_rays_depth(self, locations, dirs):
        """nerf.py:353-392: depth along arbitrary rays; white pixels and depth<near -> 0.  Like the reference, the
        view directions handed to the network are the RAW ``dirs`` (``rays = torch.cat([rays, dirs], -1)``, :365),
        not their normalisation."""
        rays_o = torch.Tensor(locations).reshape(-1, 3)
        rays_d = torch.Tensor(dirs).reshape(-1, 3)
        sh = torch.Tensor(dirs).shape
        o = self._render(rays_o, rays_d, viewdirs=rays_d)
        depth = o["depth"].clone()
        depth[torch.mean(o["rgb"], -1) * 255 > 250] = 0
        depth[depth < self.near] = 0
        return depth.reshape(sh[:-1])
