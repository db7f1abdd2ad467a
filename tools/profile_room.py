"""Short single-GPU run of the scoring hot path at the BENCHMARKED launch size, for ncu captures: 32 views of the
Room sweep shape = 393 216 rays = exactly three 131072-ray chunks, i.e. per scoring call three sigma-only coarse
launches (8.39 M evaluations) and three beta-only fine launches (25.17 M evaluations) of mlp_tc_kernel.

    ncu --set full --clock-control none --import-source on -k regex:mlp_tc -c 2 -o rep python tools/profile_room.py fp16
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import evpp_b200                                  # noqa: E402
from evpp_b200 import synthetic as S              # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
V, H, W, focal = 32, 96, 128, 96.0
torch.manual_seed(0)
c, f = S.init_nerf_state_dict(), S.init_nerf_state_dict()
eng = evpp_b200.Engine(device=0, precision=prec, near=0.5, far=6.0)
eng.load_weights(0, c)
eng.load_weights(1, f)
poses = torch.Tensor(np.asarray(S.views_to_poses(S.room_views(V, seed=0))))[:, :3, :4].contiguous().cuda()
for i in range(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = eng.score_views(poses, H, W, focal, focal, 0.5 * W, 0.5 * H)
    torch.cuda.synchronize()
    print(f"iter {i}: {1e3 * (time.perf_counter() - t0):.2f} ms for {V} views -> {V / (time.perf_counter() - t0):.1f} views/s")
print("scores", s[:4].cpu().numpy(), "launches", eng.counters())
