"""Generate tests/golden/tsdf_small.npz by running the reference's own ``render_rays`` /
``isectSegAABB_gpu`` function bodies (build container only; see oracle/tsdf_oracle.py).

Usage:  python oracle/make_golden_tsdf.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import evpp_oracle as O  # noqa: E402
from oracle import tsdf_oracle as T  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    assert T.reference_available(), "needs /root/reference"
    dims, origin = T.volume_geometry()
    state, nray = T.synthetic_state_volume(dims, seed=0)
    st, nr = torch.from_numpy(state), torch.from_numpy(nray)
    H = W = 400
    K = np.array([[300, 0, 200], [0, 300, 200], [0, 0, 1.]])
    # 16 candidate views: 12 around the object + 4 from corners of the sampling box (rays leaving the map,
    # negative voxel indices, rays missing the scoring box), 250 random pixels each
    views = np.concatenate([O.hemisphere_views(12, seed=3, center=(0., 1., 0.5), rmin=2.5, rmax=5.5),
                            np.array([[-4.8, 0.3, -4.8, 0.2, 0.7], [4.7, 4.6, 5.7, -0.4, 3.5],
                                      [0.0, 0.05, 0.0, 1.2, 0.0], [4.9, 2.0, 0.0, 0.0, 1.5708]])])
    poses = O.views_to_poses(views)
    rng = np.random.RandomState(1)
    pix = np.stack([rng.choice(H * W, size=250, replace=False) for _ in range(len(poses))]).astype(np.int32)
    ro, rd = O.build_view_rays(poses, H, W, K, pix)
    stages = {}
    u, d = T.render_rays(ro, rd, st, nr, origin, stages=stages)
    ur, dr = T.run_reference_render_rays(ro, rd, st, nr, origin)
    assert torch.equal(u, ur) and torch.equal(d, dr), "restatement differs from the reference's function bodies"
    # axis-parallel rays (slab test skips the axis) and a ray starting inside the box
    ro2 = torch.tensor([[0., 1., -4.], [0., 1., 0.], [3., 1., 0.], [0., 4.9, 0.5]])
    rd2 = torch.tensor([[0., 0., 1.], [1., 0., 0.], [0., 0., 1.], [0., -1., 0.]])
    u2, d2 = T.render_rays(ro2, rd2, st, nr, origin)
    ur2, dr2 = T.run_reference_render_rays(ro2, rd2, st, nr, origin)
    assert torch.equal(u2, ur2) and torch.equal(d2, dr2)
    vu, vd = T.view_means(u, d, len(poses), pix.shape[1])
    np.savez_compressed(os.path.join(OUT, "tsdf_small.npz"), views=views, pix=pix, H=H, W=W, K=K,
                        rays_o=ro.numpy(), rays_d=rd.numpy(), uncer=u.numpy(), depth=d.numpy(),
                        n_unknown=stages["n_unknown"].numpy(), first_hit=stages["first_hit"].numpy(),
                        aabb_hit=stages["aabb_hit"].numpy(), voxel_idx=stages["voxel_idx"].numpy().astype(np.int16),
                        view_uncer=np.asarray(vu), view_depth=np.asarray(vd),
                        rays_o2=ro2.numpy(), rays_d2=rd2.numpy(), uncer2=u2.numpy(), depth2=d2.numpy(),
                        state_seed=0, dims=dims, origin=origin)
    print("tsdf_small.npz written:", len(u), "rays,", int((stages['first_hit'] != 0).sum()), "hits,",
          int((stages['aabb_hit'] == 0).sum()), "box misses; reference function bodies == restatement")


if __name__ == "__main__":
    main()
