"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

Usage:  python oracle/make_golden.py
Writes small fixtures (inputs + reference outputs + per-stage tensors) that pin both the
oracle restatement (``-m "not gpu"`` tests) and the CUDA path (``-m gpu`` tests).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import evpp_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sd_to_np(sd, prefix):
    return {prefix + k: v.numpy() for k, v in sd.items()}


def main():
    torch.set_num_threads(os.cpu_count())
    H, R = ref_import.load()
    os.makedirs(OUT, exist_ok=True)

    # ---- weights: the reference's own NeRF constructor under manual_seed(0) --------------------
    torch.manual_seed(0)
    ref_c = H.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
    ref_f = H.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
    torch.manual_seed(0)
    sd_c = O.init_nerf_state_dict()
    sd_f = O.init_nerf_state_dict()
    for k, v in ref_c.state_dict().items():
        assert torch.equal(v, sd_c[k]), k
    for k, v in ref_f.state_dict().items():
        assert torch.equal(v, sd_f[k]), k
    assert list(ref_c.state_dict().keys()) == O.nerf_state_dict_keys()
    print("weight init identical to the reference constructor")

    for variant in ("plain", "spread"):
        c, f = (sd_c, sd_f) if variant == "plain" else (O.spread_variant(sd_c), O.spread_variant(sd_f))
        kwargs = ref_import.make_reference_kwargs(H, R, c, f)

        # ---- golden A: 6 views x (8x6) rays, near .5 far 6, full per-stage dump ------------------
        Himg, Wimg, focal = 6, 8, 6.0
        K = O.make_K(Himg, Wimg, focal)
        views = O.circle_views(6)
        poses = O.views_to_poses(views)
        batch_rays = O.build_view_rays(poses, Himg, Wimg, K)
        # the reference's own ray generator must agree with the restated one
        p32 = torch.Tensor(np.asarray(poses))[:, :3, :4]
        ro, rd = H.get_rays(Himg, Wimg, K, p32[2], torch.device("cpu"))
        assert torch.equal(rd.reshape(-1, 3), batch_rays[1][2 * 48:3 * 48])
        with torch.no_grad():
            rgb, disp, acc, unc, depth, extras = R.render(
                Himg, Wimg, K, chunk=1024 * 32, rays=batch_rays, near=0.5, far=6.0,
                retraw=True, dev=torch.device("cpu"), **kwargs)
            V, Rv = len(poses), Himg * Wimg
            uncers = torch.sum((unc ** 2).reshape(V, Rv * 192), -1) / Rv     # nerf.py:307-309
        stages = {}
        with torch.no_grad():
            o_uncers = O.score_views(poses, Himg, Wimg, K, 0.5, 6.0, c, f, stages=stages)
            o_out = O.render(Himg, Wimg, K, rays=batch_rays, near=0.5, far=6.0, sd_coarse=c, sd_fine=f)
        assert torch.equal(o_uncers, uncers), (o_uncers, uncers)
        for a, b in zip(o_out[:5], (rgb, disp, acc, unc, depth)):
            assert torch.equal(a, b)
        for k in extras:
            assert torch.equal(o_out[5][k], extras[k]), k
        print(variant, "oracle == reference (bit-exact) on golden A; scores", uncers.numpy())
        fx = dict(views=views, poses=np.asarray(poses), H=Himg, W=Wimg, focal=focal, near=0.5, far=6.0,
                  rays_o=batch_rays[0].numpy(), rays_d=batch_rays[1].numpy(),
                  rgb=rgb.numpy(), disp=disp.numpy(), acc=acc.numpy(), unc=unc.numpy(), depth=depth.numpy(),
                  raw=extras["raw"].numpy(), rgb0=extras["rgb0"].numpy(), disp0=extras["disp0"].numpy(),
                  acc0=extras["acc0"].numpy(), unc0=extras["uncertainty_map0"].numpy(),
                  depth0=extras["depth_map0"].numpy(), z_std=extras["z_std"].numpy(), uncers=uncers.numpy())
        for k, v in stages.items():
            fx["st_" + k] = torch.cat(v, 0).numpy()
        np.savez_compressed(os.path.join(OUT, f"render_small_{variant}.npz"), **fx)

        # ---- golden B: far=80 (live bounds, nerf.py:100-101), pixel subset, 4 views ---------------
        Himg, Wimg, focal = 40, 40, 30.0
        K = O.make_K(Himg, Wimg, focal)
        views = O.hemisphere_views(4, seed=1)
        poses = O.views_to_poses(views)
        rng = np.random.RandomState(3)
        pix = np.stack([rng.choice(Himg * Wimg, size=20, replace=False) for _ in range(4)]).astype(np.int32)
        batch_rays = O.build_view_rays(poses, Himg, Wimg, K, pix)
        with torch.no_grad():
            rgb, disp, acc, unc, depth, extras = R.render(
                Himg, Wimg, K, chunk=1024 * 32, rays=batch_rays, near=0.5, far=80.0,
                retraw=True, dev=torch.device("cpu"), **kwargs)
            uncers = torch.sum((unc ** 2).reshape(4, 20 * 192), -1) / 20
            o_uncers = O.score_views(poses, Himg, Wimg, K, 0.5, 80.0, c, f, pix_idx=pix)
        assert torch.equal(o_uncers, uncers)
        np.savez_compressed(os.path.join(OUT, f"score_far80_{variant}.npz"), views=views,
                            poses=np.asarray(poses), H=Himg, W=Wimg, focal=focal, near=0.5, far=80.0,
                            pix=pix, uncers=uncers.numpy(), depth=depth.numpy(), rgb=rgb.numpy(),
                            unc=unc.numpy())
        print(variant, "golden B scores", uncers.numpy())

    # ---- golden C: config-1 shape (32 views x 40x30), scores + NBV only, spread weights ------------
    c, f = O.spread_variant(sd_c), O.spread_variant(sd_f)
    kwargs = ref_import.make_reference_kwargs(H, R, c, f)
    Himg, Wimg, focal = 30, 40, 30.0
    K = O.make_K(Himg, Wimg, focal)
    views = O.circle_views(32)
    poses = O.views_to_poses(views)
    batch_rays = O.build_view_rays(poses, Himg, Wimg, K)
    with torch.no_grad():
        rgb, disp, acc, unc, depth, extras = R.render(
            Himg, Wimg, K, chunk=1024 * 32, rays=batch_rays, near=0.5, far=6.0,
            retraw=True, dev=torch.device("cpu"), **kwargs)
        uncers = torch.sum((unc ** 2).reshape(32, 1200 * 192), -1) / 1200
    data, nbv, win = O.select_nbv(views, uncers.numpy(), dirs_per_location=1)
    np.savez_compressed(os.path.join(OUT, "config1_spread.npz"), views=views, poses=np.asarray(poses),
                        H=Himg, W=Wimg, focal=focal, near=0.5, far=6.0, uncers=uncers.numpy(),
                        nbv=nbv, win=win, depth_mean=depth.reshape(32, -1).mean(-1).numpy())
    print("config1 spread scores", uncers.numpy(), "winner", win)

    # ---- golden D: MLP-only known answers (reference NeRF.forward on fixed embedded inputs) ---------
    torch.manual_seed(5)
    pts = (torch.rand(256, 3) - 0.5) * 8
    dirs = torch.nn.functional.normalize(torch.randn(256, 3), dim=-1)
    emb_fn, _, _ = H.get_embedder(10, 0, torch.device("cpu"))
    embd_fn, _, _ = H.get_embedder(4, 0, torch.device("cpu"))
    x = torch.cat([emb_fn(pts, torch.device("cpu")), embd_fn(dirs, torch.device("cpu"))], -1)
    with torch.no_grad():
        y = ref_c(x)
        y_o = O.nerf_forward(sd_c, torch.cat([O.embed(pts, 10), O.embed(dirs, 4)], -1))
    assert torch.equal(y, y_o)
    np.savez_compressed(os.path.join(OUT, "mlp_kat.npz"), pts=pts.numpy(), dirs=dirs.numpy(),
                        emb=x.numpy(), out=y.numpy())
    # weights are regenerated from the seed on the GPU box; store a checksum to pin them
    chk = {k: float(v.double().sum()) for k, v in sd_c.items()}
    np.savez_compressed(os.path.join(OUT, "weights_seed0_checksum.npz"),
                        keys=np.array(list(chk.keys())), sums=np.array(list(chk.values())),
                        first=sd_c["pts_linears.0.weight"][:4, :8].numpy(),
                        fine_first=sd_f["pts_linears.0.weight"][:4, :8].numpy())
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
