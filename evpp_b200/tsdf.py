"""Drop-in for the scoring half of the planner's ``TSDF`` class (plannerServer_*/core/tsdf_gpu.py):
``render`` (:356-382) and ``get_uncertainty_tsdf`` (:473-530) on the CUDA engine.  Map building
(``tsdf_test``, fusion, dilation, marching cubes) stays with the planner and hands its state volume over
through ``set_state`` -- the same tensors the reference keeps in ``State_vol_origin`` / ``Nray_vol``."""
import ctypes as C

import numpy as np
import torch

from ._lib import check
from .engine import Engine, _ptr, _stream_ptr
from .run_nerf import _intrinsics, get_pose, get_poses

# scene constants of the Object (cabin) planner, tsdf_gpu.py:30-41
VOL_BNDS = np.array([[-5, 5], [0.01, 5.0], [-5, 6]], dtype=np.float64)
AABB_BNDS = np.array([[-1.0, 1.0], [0.02, 2.5], [-1.5, 2.0]], dtype=np.float64)
NEAR, FAR, VOXEL_RES = 0.5, 6.0, 0.1


class TSDF:
    def __init__(self, engine=None, device=0, H=400, W=400, K=None, vol_bnds=VOL_BNDS, aabb_bnds=AABB_BNDS,
                 voxel_res=VOXEL_RES, near=NEAR, far=FAR, unknown_cof=0.2, surface_cof=0.1, sample_num=500,
                 pixel_sampler="numpy", seed=0):
        self.engine = engine if engine is not None else Engine(device=device, precision="fp32", near=near, far=far)
        self.H, self.W = H, W
        self.K = np.array([[300, 0, 200], [0, 300, 200], [0, 0, 1]], dtype=np.float64) if K is None else np.asarray(K)
        self.vol_bnds = np.asarray(vol_bnds, dtype=np.float64)
        self.aabb_bnds = np.asarray(aabb_bnds, dtype=np.float64)
        self.voxel_res, self.near, self.far = float(voxel_res), float(near), float(far)
        self.unknown_cof, self.surface_cof = float(unknown_cof), float(surface_cof)
        self.sample_num = sample_num
        self.pixel_sampler, self._seed = pixel_sampler, int(seed)   # "numpy": tsdf_gpu.py:497's protocol; "device": evpp_sample_pixels
        # fusion.py:37-39
        self.index_bnd = np.ceil((self.vol_bnds[:, 1] - self.vol_bnds[:, 0]) / self.voxel_res).astype(int)
        self.vol_origin = self.vol_bnds[:, 0].astype(np.float32)
        self.N_samples = int((self.far - self.near) / self.voxel_res) + 1

    def get_pose(self, location, u, v):
        return get_pose(location, u, v)

    def set_state(self, State_vol_origin, Nray_vol):
        """Snapshot the planner's state map (float volumes of shape index_bnd, host or cuda tensors)."""
        st = torch.as_tensor(State_vol_origin, dtype=torch.float32).contiguous()
        nr = torch.as_tensor(Nray_vol, dtype=torch.float32).contiguous()
        if tuple(st.shape) != tuple(self.index_bnd) or tuple(nr.shape) != tuple(self.index_bnd):
            raise ValueError(f"state volumes must have shape {tuple(self.index_bnd)}")
        dims = np.asarray(self.index_bnd, dtype=np.int32)
        amin = self.aabb_bnds[:, 0].astype(np.float32)
        amax = self.aabb_bnds[:, 1].astype(np.float32)
        e = self.engine
        with torch.cuda.device(e.device):
            check(e.lib.evpp_tsdf_set_volume(e._h, _ptr(st), _ptr(nr), dims.ctypes.data_as(C.c_void_p),
                                             self.vol_origin.ctypes.data_as(C.c_void_p), self.voxel_res,
                                             amin.ctypes.data_as(C.c_void_p), amax.ctypes.data_as(C.c_void_p),
                                             self.near, self.far, self.N_samples, self.unknown_cof, self.surface_cof))

    def render(self, H, W, K, chunk=1024 * 32, rays=None, stages=False):
        """tsdf_gpu.py:356-382.  Returns [rgb_map, depth_map, uncer_map, extras]; rgb_map (the debug colour of
        the first hit) is not produced by the scorer and is None.  ``stages=True`` adds the integer per-ray
        quantities (n_unknown, first_hit, voxel_idx) to extras for parity tests."""
        e = self.engine
        rays_o, rays_d = rays
        sh = tuple(rays_d.shape)
        ro = torch.as_tensor(rays_o).to(e.device, torch.float32).reshape(-1, 3).contiguous()
        rd = torch.as_tensor(rays_d).to(e.device, torch.float32).reshape(-1, 3).contiguous()
        n = ro.shape[0]
        uncer = torch.empty(n, device=e.device)
        depth = torch.empty(n, device=e.device)
        extras = {}
        if stages:
            extras = {"n_unknown": torch.empty(n, device=e.device, dtype=torch.int32),
                      "first_hit": torch.empty(n, device=e.device, dtype=torch.int32),
                      "voxel_idx": torch.empty(n, self.N_samples, 3, device=e.device, dtype=torch.int32)}
        with torch.cuda.device(e.device):
            check(e.lib.evpp_tsdf_render_rays(e._h, _ptr(ro), _ptr(rd), n, _ptr(uncer), _ptr(depth),
                                              _ptr(extras.get("n_unknown")), _ptr(extras.get("first_hit")),
                                              _ptr(extras.get("voxel_idx")), _stream_ptr(e.device)))
        return [None, depth.reshape(sh[:-1]), uncer.reshape(sh[:-1]), extras]

    def get_uncertainty_tsdf(self, locations, us, vs, pix_idx=None):
        """tsdf_gpu.py:473-530: (uncer_list, depth_list) for the candidate views; ``sample_num`` random pixels
        per view drawn from NumPy's global RNG like the reference (:497), or explicit ``pix_idx`` [V,R]."""
        V = len(locations)
        if V == 0:
            return [], []
        poses = torch.Tensor(get_poses(locations, us, vs))[:, :3, :4].contiguous()
        if pix_idx is None and self.sample_num is not None and self.pixel_sampler == "device":
            e = self.engine
            self._seed += 1
            pix_dev = e.sample_pixels(V, self.sample_num, self.H * self.W, self._seed)
            fx, fy, cx, cy = _intrinsics(self.K)
            c2w = poses.to(e.device)
            u = torch.empty(V, device=e.device, dtype=torch.float32)
            d = torch.empty(V, device=e.device, dtype=torch.float32)
            with torch.cuda.device(e.device):
                check(e.lib.evpp_tsdf_score_views(e._h, _ptr(c2w), V, self.H, self.W, fx, fy, cx, cy, _ptr(pix_dev),
                                                  self.sample_num, _ptr(u), _ptr(d), _stream_ptr(e.device)))
            return u.tolist(), d.tolist()
        if pix_idx is None and self.sample_num is not None:
            pix_idx = np.stack([np.random.choice(self.H * self.W, size=[self.sample_num], replace=False) for _ in range(V)])
        pix_t = None if pix_idx is None else torch.as_tensor(np.ascontiguousarray(pix_idx, dtype=np.int32))
        R = 0 if pix_t is None else pix_t.shape[1]
        fx, fy, cx, cy = _intrinsics(self.K)
        u = torch.empty(V, dtype=torch.float32)
        d = torch.empty(V, dtype=torch.float32)
        e = self.engine
        with torch.cuda.device(e.device):
            check(e.lib.evpp_tsdf_score_views_host(e._h, _ptr(poses), V, self.H, self.W, fx, fy, cx, cy, _ptr(pix_t), R,
                                                   _ptr(u), _ptr(d), _stream_ptr(e.device)))
        return u.tolist(), d.tolist()
