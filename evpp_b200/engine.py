"""Python face of the C-ABI engine: owns an evpp_handle, moves torch tensors across the boundary.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in the CUDA library.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import EvppConfig, EvppRenderOut, check, PRECISIONS

STATE_DICT_KEYS = [f"pts_linears.{i}.{p}" for i in range(8) for p in ("weight", "bias")] + \
    [f"{n}.{p}" for n in ("views_linears.0", "feature_linear", "alpha_linear", "rgb_linear", "uncertainty_linear")
     for p in ("weight", "bias")]
STATE_DICT_SHAPES = [(256, 63), (256,)] + [(256, 256), (256,)] * 4 + [(256, 319), (256,)] + [(256, 256), (256,)] * 2 + \
    [(128, 283), (128,), (256, 256), (256,), (1, 256), (1,), (3, 128), (3,), (1, 128), (1,)]

STAGES = {"z_coarse": (0, torch.float32), "raw_coarse": (1, torch.float32), "weights_coarse": (2, torch.float32),
          "cdf": (3, torch.float32), "inds": (4, torch.int32), "z_samples": (5, torch.float32),
          "z_fine": (6, torch.float32)}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """One evpp_handle on one GPU.  Calls on one Engine must be serialised by the caller."""

    def __init__(self, device=0, n_samples=64, n_importance=128, near=0.5, far=80.0, white_bkgd=True,
                 lindisp=False, precision="fp16", max_chunk_rays=0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.EvppError(-4, "no CUDA device visible; evpp_b200 has no CPU fallback")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.precision = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        self.cfg = EvppConfig(n_samples, n_importance, near, far, int(bool(white_bkgd)), int(bool(lindisp)),
                              self.precision, max_chunk_rays)
        self.n_samples, self.n_importance = n_samples, n_importance
        self.S_f = n_samples + n_importance
        self._h = C.c_void_p()
        # (no torch.cuda.device here: the C side validates the ordinal and restores the caller's current device)
        check(self.lib.evpp_create(self.device.index, C.byref(self.cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            with torch.cuda.device(self.device):
                self.lib.evpp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights (Controller.copy_model, nerf.py:718-757) -------------------------------------------
    def load_weights(self, which, state_dict):
        """which: 0/'coarse' (network_fn) or 1/'fine' (network_fine); state_dict: reference NeRF keys."""
        which = {"coarse": 0, "fine": 1}.get(which, which)
        if hasattr(state_dict, "state_dict"):
            state_dict = state_dict.state_dict()
        flat, offsets, off = [], [], 0
        for k, shape in zip(STATE_DICT_KEYS, STATE_DICT_SHAPES):
            if k not in state_dict:
                raise KeyError(f"state_dict lacks {k} (expected reference NeRF(D=8,W=256,use_viewdirs=True))")
            t = state_dict[k].detach().to("cpu", torch.float32).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"{k}: shape {tuple(t.shape)} != {shape}")
            flat.append(t.reshape(-1))
            offsets.append(off)
            off += t.numel()
        packed = torch.cat(flat).contiguous()
        offs = np.asarray(offsets, dtype=np.int64)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_load_nerf_weights(self._h, which, C.c_void_p(packed.data_ptr()),
                                                  offs.ctypes.data_as(C.c_void_p), len(offsets)))

    def set_render_mode(self, fold_feature_layer):
        """True (default): render_rays on the tensor-core precisions evaluates feature_linear folded into the view
        layer (same function, nine GEMM layers); False: the reference's ten layers (evpp_set_render_mode)."""
        check(self.lib.evpp_set_render_mode(self._h, int(bool(fold_feature_layer))))

    def load_checkpoint(self, ckpt):
        """Both networks from a checkpoint of the reference: the ``.tar`` written by ``Controller.save_model``
        (nerfServer/core/nerf.py:612-621; reloaded by create_nerf, run_nerf.py:292-312) or the dict inside it, keys
        ``network_fn_state_dict`` / ``network_fine_state_dict`` (``global_step`` and ``optimizer_state_dict`` are
        ignored).  Returns ``global_step``."""
        if isinstance(ckpt, (str, bytes)):
            ckpt = torch.load(ckpt, map_location="cpu", weights_only=True)
        for key in ("network_fn_state_dict", "network_fine_state_dict"):
            if key not in ckpt or ckpt[key] is None:
                raise KeyError(f"checkpoint lacks {key} (expected the dict of Controller.save_model, nerf.py:612-621)")
        self.load_weights(0, ckpt["network_fn_state_dict"])
        self.load_weights(1, ckpt["network_fine_state_dict"])
        return ckpt.get("global_step")

    # ---- scoring (Controller.get_uncertainty, nerf.py:266-309) -------------------------------------
    def score_views(self, c2w, H, W, fx, fy, cx, cy, pix_idx=None, precision=None):
        """c2w: cuda float32 [V,3,4]; pix_idx: cuda int32 [V,R] or None.  Returns cuda float32 [V] (asynchronous:
        a non-finite score is not reported here, see nonfinite_count()).  precision: None = the bulk precision."""
        c2w = c2w.to(self.device, torch.float32).contiguous()
        V = c2w.shape[0]
        R = 0
        if pix_idx is not None:
            pix_idx = pix_idx.to(self.device, torch.int32).contiguous()
            R = pix_idx.shape[1]
        scores = torch.empty(V, device=self.device, dtype=torch.float32)
        prec = -1 if precision is None else (PRECISIONS[precision] if isinstance(precision, str) else int(precision))
        with torch.cuda.device(self.device):
            check(self.lib.evpp_score_views_ex(self._h, _ptr(c2w), V, H, W, fx, fy, cx, cy, _ptr(pix_idx), R, prec,
                                               _ptr(scores), _stream_ptr(self.device)))
        return scores

    def nonfinite_count(self):
        """Number of inf / NaN scores of the last device-side scoring call (synchronises)."""
        n = C.c_int32(0)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_get_nonfinite(self._h, C.byref(n), _stream_ptr(self.device)))
        return n.value

    def score_views_host(self, c2w_host, H, W, fx, fy, cx, cy, pix_idx_host=None, out=None, precision=None):
        """HOST in / HOST out (pinned torch CPU tensors recommended).  Synchronises.  Raises EvppError(-5,
        EVPP_ERR_NONFINITE) when a score is inf / NaN (16-bit activations overflowed); ``out`` is filled anyway.
        precision: None = the engine's bulk precision, else any precision name / code for this call."""
        assert c2w_host.device.type == "cpu" and c2w_host.dtype == torch.float32 and c2w_host.is_contiguous()
        V = c2w_host.shape[0]
        R = 0
        if pix_idx_host is not None:
            assert pix_idx_host.dtype == torch.int32 and pix_idx_host.is_contiguous()
            R = pix_idx_host.shape[1]
        if out is None:
            out = torch.empty(V, dtype=torch.float32).pin_memory()
        with torch.cuda.device(self.device):
            if precision is None:
                check(self.lib.evpp_score_views_host(self._h, _ptr(c2w_host), V, H, W, fx, fy, cx, cy,
                                                     _ptr(pix_idx_host), R, _ptr(out), _stream_ptr(self.device)))
            else:
                prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
                ids = torch.arange(V, dtype=torch.int32)
                check(self.lib.evpp_rescore_views(self._h, _ptr(c2w_host), _ptr(ids), V, H, W, fx, fy, cx, cy,
                                                  _ptr(pix_idx_host), R, prec, _ptr(out), _stream_ptr(self.device)))
        return out

    def rescore_views(self, c2w_host, view_ids, H, W, fx, fy, cx, cy, pix_idx_host=None, precision="fp16x3"):
        """Scores of the views ``view_ids`` (indices into the HOST pose array) at ``precision`` regardless of the
        engine's bulk precision.  "fp16x3" = the tensor-core exact path (split-fp16 operands, three MMAs per
        product, fp32-class results), "fp32" = the FFMA kernel (numerical anchor).  Returns a host tensor [K]."""
        ids = torch.as_tensor(np.asarray(view_ids), dtype=torch.int32).contiguous()
        R = 0 if pix_idx_host is None else pix_idx_host.shape[1]
        out = torch.empty(len(ids), dtype=torch.float32)
        prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_rescore_views(self._h, _ptr(c2w_host), _ptr(ids), len(ids), H, W, fx, fy, cx, cy,
                                              _ptr(pix_idx_host), R, prec, _ptr(out), _stream_ptr(self.device)))
        return out

    def rescore_views_fp32(self, c2w_host, view_ids, H, W, fx, fy, cx, cy, pix_idx_host=None):
        return self.rescore_views(c2w_host, view_ids, H, W, fx, fy, cx, cy, pix_idx_host, precision="fp32")

    def set_score_mode(self, mode):
        """What scoring calls evaluate (evpp_set_score_mode).  False / 0 / "lean" (default): only what the score
        depends on, with feature_linear folded into the view layer (same function, one GEMM layer less);
        2 / "lean_unfolded": the same elimination in the reference's layer order (bit-identical to the full
        evaluation); True / 1 / "full": the full network like render()."""
        code = {"lean": 0, "full": 1, "lean_unfolded": 2}.get(mode, mode) if isinstance(mode, str) else int(mode)
        check(self.lib.evpp_set_score_mode(self._h, int(code)))

    def render_rays(self, rays_o, rays_d, near, far, precision=None, want=("rgb", "disp", "acc", "depth", "unc"),
                    capture_stages=False, viewdirs=None):
        """viewdirs: optional [N,3] view directions used as given (None = rays_d / ||rays_d||, run_nerf.py:170)."""
        rays_o = rays_o.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        rays_d = rays_d.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        if viewdirs is not None:
            viewdirs = viewdirs.to(self.device, torch.float32).reshape(-1, 3).contiguous()
            assert viewdirs.shape == rays_d.shape
        N = rays_o.shape[0]
        S = self.S_f if self.n_importance > 0 else self.n_samples
        shapes = {"rgb": (N, 3), "disp": (N,), "acc": (N,), "depth": (N,), "unc": (N, S), "raw": (N, S, 5),
                  "rgb0": (N, 3), "disp0": (N,), "acc0": (N,), "depth0": (N,), "unc0": (N, self.n_samples),
                  "z_std": (N,)}
        outs = {k: torch.empty(shapes[k], device=self.device, dtype=torch.float32) for k in want}
        ro = EvppRenderOut(**{k: (outs[k].data_ptr() if k in outs else None) for k in shapes})
        prec = self.precision if precision is None else (PRECISIONS[precision] if isinstance(precision, str) else precision)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_set_stage_capture(self._h, int(capture_stages)))
            check(self.lib.evpp_render_rays_ex(self._h, _ptr(rays_o), _ptr(rays_d), _ptr(viewdirs), N, float(near),
                                               float(far), prec, C.byref(ro), _stream_ptr(self.device)))
        return outs

    def stage(self, name):
        sid, dt = STAGES[name]
        cnt = C.c_int64(0)
        check(self.lib.evpp_stage_dump(self._h, sid, C.c_void_p(0), C.byref(cnt), C.c_void_p(0)))
        out = torch.empty(cnt.value, device=self.device, dtype=dt)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_stage_dump(self._h, sid, _ptr(out), C.byref(cnt), _stream_ptr(self.device)))
        return out

    def get_rays(self, c2w, H, W, fx, fy, cx, cy, pix_idx=None):
        c2w = c2w.to(self.device, torch.float32).contiguous()
        V = c2w.shape[0]
        R = H * W
        if pix_idx is not None:
            pix_idx = pix_idx.to(self.device, torch.int32).contiguous()
            R = pix_idx.shape[1]
        o = torch.empty(V * R, 3, device=self.device)
        d = torch.empty_like(o)
        vd = torch.empty_like(o)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_get_rays(self._h, _ptr(c2w), V, H, W, fx, fy, cx, cy, _ptr(pix_idx), R, _ptr(o),
                                         _ptr(d), _ptr(vd), _stream_ptr(self.device)))
        return o, d, vd

    def mlp_forward(self, which, pts, dirs, precision=None):
        which = {"coarse": 0, "fine": 1}.get(which, which)
        pts = pts.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        dirs = dirs.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        out = torch.empty(pts.shape[0], 5, device=self.device)
        prec = self.precision if precision is None else (PRECISIONS[precision] if isinstance(precision, str) else precision)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_mlp_forward(self._h, which, prec, _ptr(pts), _ptr(dirs), pts.shape[0], _ptr(out),
                                            _stream_ptr(self.device)))
        return out

    def mlp_layer_dump(self, which, pts, dirs, layer, precision="fp16"):
        """Tensor-core MLP + fp32 dump of the activations layer `layer` produces; returns (out [M,5], acts [M,256])."""
        which = {"coarse": 0, "fine": 1}.get(which, which)
        pts = pts.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        dirs = dirs.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        out = torch.empty(pts.shape[0], 5, device=self.device)
        acts = torch.zeros(pts.shape[0], 256, device=self.device)
        prec = PRECISIONS[precision] if isinstance(precision, str) else precision
        with torch.cuda.device(self.device):
            check(self.lib.evpp_mlp_layer_dump(self._h, which, prec, _ptr(pts), _ptr(dirs), pts.shape[0], int(layer),
                                               _ptr(out), _ptr(acts), _stream_ptr(self.device)))
        return out, acts

    def mlp_trace(self, which, pts, dirs, precision="fp16"):
        """Tensor-core MLP + SM-clock stamps of one tile's pipeline (see evpp_mlp_trace); returns int64 [128]."""
        which = {"coarse": 0, "fine": 1}.get(which, which)
        pts = pts.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        dirs = dirs.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        out = torch.empty(pts.shape[0], 5, device=self.device)
        trace = torch.zeros(128, device=self.device, dtype=torch.int64)
        prec = PRECISIONS[precision] if isinstance(precision, str) else precision
        with torch.cuda.device(self.device):
            check(self.lib.evpp_mlp_trace(self._h, which, prec, _ptr(pts), _ptr(dirs), pts.shape[0], _ptr(out),
                                          _ptr(trace), _stream_ptr(self.device)))
        return trace

    def composite(self, raw, z, rays_d):
        """raw2outputs on caller tensors; returns dict(rgb, disp, acc, depth, weights, beta2)."""
        raw = raw.to(self.device, torch.float32).contiguous()
        z = z.to(self.device, torch.float32).contiguous()
        rays_d = rays_d.to(self.device, torch.float32).contiguous()
        n, S = z.shape
        o = {"rgb": torch.empty(n, 3, device=self.device), "disp": torch.empty(n, device=self.device),
             "acc": torch.empty(n, device=self.device), "depth": torch.empty(n, device=self.device),
             "weights": torch.empty(n, S, device=self.device), "beta2": torch.empty(n, device=self.device)}
        with torch.cuda.device(self.device):
            check(self.lib.evpp_composite(self._h, _ptr(raw), _ptr(z), _ptr(rays_d), n, S, _ptr(o["rgb"]),
                                          _ptr(o["disp"]), _ptr(o["acc"]), _ptr(o["depth"]), _ptr(o["weights"]),
                                          _ptr(o["beta2"]), _stream_ptr(self.device)))
        return o

    def resample(self, raw_coarse, z_coarse, rays_d):
        """coarse compositing + sample_pdf + sort on caller tensors; returns dict(z_fine, cdf, inds, z_samples)."""
        raw = raw_coarse.to(self.device, torch.float32).contiguous()
        z = z_coarse.to(self.device, torch.float32).contiguous()
        rays_d = rays_d.to(self.device, torch.float32).contiguous()
        n = z.shape[0]
        o = {"z_fine": torch.empty(n, self.S_f, device=self.device),
             "cdf": torch.empty(n, self.n_samples - 1, device=self.device),
             "inds": torch.empty(n, self.n_importance, device=self.device, dtype=torch.int32),
             "z_samples": torch.empty(n, self.n_importance, device=self.device)}
        with torch.cuda.device(self.device):
            check(self.lib.evpp_resample(self._h, _ptr(raw), _ptr(z), _ptr(rays_d), n, _ptr(o["z_fine"]),
                                         _ptr(o["cdf"]), _ptr(o["inds"]), _ptr(o["z_samples"]),
                                         _stream_ptr(self.device)))
        return o

    def select_nbv(self, scores, weight=None, dirs_per_location=1):
        scores = scores.to(self.device, torch.float32).contiguous()
        V = scores.shape[0]
        L = V // dirs_per_location
        if weight is not None:
            weight = weight.to(self.device, torch.float32).contiguous()
        norm = torch.empty(L, device=self.device)
        best = torch.empty(L, device=self.device, dtype=torch.int32)
        win = torch.empty(1, device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_select_nbv(self._h, _ptr(scores), _ptr(weight), V, dirs_per_location, _ptr(norm),
                                           _ptr(best), _ptr(win), _stream_ptr(self.device)))
        return norm, best, win

    def sample_pixels(self, V, R, n_pixels, seed):
        """R distinct pseudo-random pixel ids per view on the device (evpp_sample_pixels), int32 [V, R]."""
        out = torch.empty(V, R, device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_sample_pixels(self._h, C.c_uint64(int(seed) & (2 ** 64 - 1)), V, R, n_pixels, _ptr(out),
                                              _stream_ptr(self.device)))
        return out

    def check_guards(self):
        """(corrupted guard bytes, buffers checked): every engine-owned device buffer has guard bands on both sides;
        a non-zero count means some kernel wrote out of bounds (evpp_check_guards).  Synchronises."""
        bad, n = C.c_int32(0), C.c_int32(0)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_check_guards(self._h, C.byref(bad), C.byref(n)))
        return bad.value, n.value

    def poison_workspace(self):
        """Fill the per-call workspace with NaN bit patterns (self-check hook, evpp_poison_workspace)."""
        with torch.cuda.device(self.device):
            check(self.lib.evpp_poison_workspace(self._h))

    # ---- counters -----------------------------------------------------------------------------------
    def counters(self):
        a, b = C.c_int64(0), C.c_int64(0)
        check(self.lib.evpp_get_counters(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_profiling(self, on):
        check(self.lib.evpp_set_profiling(self._h, int(on)))

    def mlp_time(self):
        ms, n, ev = C.c_double(0), C.c_int64(0), C.c_int64(0)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_get_mlp_time_ms(self._h, C.byref(ms), C.byref(n), C.byref(ev)))
        return ms.value, n.value, ev.value

    def mlp_profile(self):
        """(total ms, launches, evaluations, executed FLOPs) of the MLP launches recorded since set_profiling(True)."""
        ms, n, ev, fl = C.c_double(0), C.c_int64(0), C.c_int64(0), C.c_double(0)
        with torch.cuda.device(self.device):
            check(self.lib.evpp_get_mlp_profile(self._h, C.byref(ms), C.byref(n), C.byref(ev), C.byref(fl)))
        return ms.value, n.value, ev.value, fl.value
