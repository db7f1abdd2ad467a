"""Golden fixture for the Controller seam (rows a1, a11, f2): runs the UNMODIFIED source text of the reference's
``Controller.get_uncertainty`` / ``get_surface_points`` / ``get_rays_depth`` (nerfServer/core/nerf.py:266-392) and
``views.get_pose`` (nerfServer/core/views.py:141-150), cut out with ``ast`` (their modules import Django /
configargparse and cannot be imported), over the unmodified ``render`` / ``batchify_rays`` / ``get_rays`` of the
reference, asserts that the oracle restatements reproduce every output bit for bit, and writes
tests/golden/controller_queries.npz.  Build container only (needs /root/reference).

TEST INFRASTRUCTURE -- never imported by the product.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import evpp_oracle as O, ref_import  # noqa: E402

REF_NERF = "/root/reference/nerfServer/core/nerf.py"
REF_VIEWS = "/root/reference/nerfServer/core/views.py"


def reference_controller_methods():
    """{name: function} of the three Controller methods and get_pose, compiled from the reference's own text."""
    H, R = ref_import.load()
    ns = {"torch": torch, "np": np, "render": R.render, "batchify_rays": R.batchify_rays, "get_rays": H.get_rays,
          "device": torch.device("cpu"), "print": lambda *a, **k: None}
    tree = ast.parse(open(REF_NERF, encoding="utf-8").read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Controller"][0]
    fns = [f for f in cls.body if isinstance(f, ast.FunctionDef) and
           f.name in ("get_uncertainty", "get_surface_points", "get_rays_depth")]
    assert len(fns) == 3
    exec(compile(ast.Module(fns, []), REF_NERF, "exec"), ns)
    tree = ast.parse(open(REF_VIEWS, encoding="utf-8").read())
    gp = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_pose"]
    exec(compile(ast.Module(gp, []), REF_VIEWS, "exec"), ns)
    return ns, H, R


def stand_in_controller(Himg, Wimg, focal, near, far, kwargs):
    """What the three methods read of ``self`` (nerf.py:86-127)."""
    K = np.array([[focal, 0, 0.5 * Wimg], [0, focal, 0.5 * Himg], [0, 0, 1]])
    coords = torch.stack(torch.meshgrid(torch.linspace(0, Himg - 1, Himg), torch.linspace(0, Wimg - 1, Wimg), indexing="ij"), -1)
    return types.SimpleNamespace(H=Himg, W=Wimg, K=K, near=near, far=far, args=types.SimpleNamespace(chunk=1024 * 32), i=100,
                                 uncertainty_kargs=[kwargs], sapmle_model_index=0, devices=[torch.device("cpu")],
                                 coords=torch.reshape(coords, [-1, 2]))


def main():
    ns, H, R = reference_controller_methods()
    torch.manual_seed(0)
    c, f = O.init_nerf_state_dict(), O.init_nerf_state_dict()
    c, f = O.spread_variant(c), O.spread_variant(f)
    Himg = Wimg = 40                      # >= 1000 pixels: get_uncertainty draws 1000 per pose (nerf.py:276)
    focal, near, far = 30.0, 0.5, 6.0
    kwargs = ref_import.make_reference_kwargs(H, R, c, f)
    kwargs.update({"near": near, "far": far})     # bds_dict, nerf.py:168-174
    ctl = stand_in_controller(Himg, Wimg, focal, near, far, kwargs)
    out = {"H": Himg, "W": Wimg, "focal": focal, "near": near, "far": far}

    # ---- a1 + a11: get_pose and get_uncertainty with the reference's own RNG protocol ----
    views = O.hemisphere_views(3, seed=5)
    poses = [ns["get_pose"](list(v[0:3]), v[3], v[4]) for v in views]
    assert np.array_equal(np.asarray(poses), np.asarray(O.views_to_poses(views))), "oracle get_pose != reference"
    np.random.seed(0)
    uncers = ns["get_uncertainty"](ctl, poses)
    np.random.seed(0)
    pix = np.stack([np.random.choice(Himg * Wimg, size=[1000], replace=False) for _ in range(len(views))]).astype(np.int32)
    mine = O.score_views(poses, Himg, Wimg, ctl.K, near, far, c, f, pix_idx=pix)
    assert torch.equal(uncers, mine), "oracle score_views != reference Controller.get_uncertainty"
    out.update({"views": np.asarray(views), "poses": np.asarray(poses, dtype=np.float64), "pix": pix, "uncers": uncers.numpy()})

    # ---- f2: depth / surface queries on the "sparse" weights (empty space + solid blobs: every mask is exercised) ----
    torch.manual_seed(0)
    c, f = O.init_nerf_state_dict(), O.init_nerf_state_dict()
    c, f = O.sparse_variant(c), O.sparse_variant(f)
    kwargs = ref_import.make_reference_kwargs(H, R, c, f)
    kwargs.update({"near": near, "far": far})
    ctl = stand_in_controller(Himg, Wimg, focal, near, far, kwargs)
    location, radius, step = [0.3, 1.0, -0.2], 3.0, 0.35
    pts_ref = ns["get_surface_points"](ctl, location, radius, step)
    pts, extra = O.get_surface_points(location, radius, step, Himg, Wimg, ctl.K, near, far, c, f, return_all=True)
    assert torch.equal(pts_ref, pts), "oracle get_surface_points != reference"
    out.update({"sp_location": np.asarray(location), "sp_radius": radius, "sp_step": step, "sp_points": pts_ref.numpy(),
                "sp_depth_raw": extra["depth_raw"].numpy(), "sp_white": extra["white"].numpy(), "sp_keep": extra["keep"].numpy(),
                "sp_rgb": extra["rgb"].numpy()})

    # ---- f2: get_rays_depth (non-unit directions on purpose: they ARE the view directions, nerf.py:365) ----
    rng = np.random.RandomState(3)
    # [N,3] inputs: the reference concatenates the [N,8] batch with ``dirs`` as given, so it only accepts 2-D lists
    locs = rng.uniform(-0.5, 0.5, size=(63, 3)) + np.array([0.0, 1.0, 0.0])
    dirs = rng.normal(size=(63, 3)) * rng.uniform(0.5, 2.0, size=(63, 1))
    d_ref = ns["get_rays_depth"](ctl, locs.tolist(), dirs.tolist())
    d, extra = O.get_rays_depth(locs.tolist(), dirs.tolist(), near, far, c, f, return_all=True)
    assert torch.equal(d_ref, d), "oracle get_rays_depth != reference"
    out.update({"rd_locations": locs, "rd_dirs": dirs, "rd_depths": d_ref.numpy(), "rd_depth_raw": extra["depth_raw"].numpy(),
                "rd_white": extra["white"].numpy(), "rd_rgb": extra["rgb"].numpy()})
    print("get_rays_depth: white", int(extra["white"].sum()), "of", extra["white"].numel(), "| zero depths", int((d_ref == 0).sum()),
          "| get_surface_points: kept", len(pts_ref), "of", int(out["sp_keep"].size), "white", int(out["sp_white"].sum()))
    path = os.path.join(ROOT, "tests", "golden", "controller_queries.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
