"""GPU-box measurement: views/s of the Object workload (512 views x 80x60 rays) as a function of max_chunk_rays."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evpp_b200 import Engine, synthetic as S      # noqa: E402

torch.manual_seed(0)
c, f = S.init_nerf_state_dict(), S.init_nerf_state_dict()
poses = torch.Tensor(np.asarray(S.views_to_poses(S.hemisphere_views(512, seed=0))))[:, :3, :4].contiguous().cuda()
for chunk in (65536, 131072, 262144, 524288, 131072):
    e = Engine(device=0, near=0.5, far=6.0, precision="fp16", max_chunk_rays=chunk)
    e.load_weights(0, c)
    e.load_weights(1, f)
    for _ in range(2):
        e.score_views(poses, 60, 80, 60.0, 60.0, 40.0, 30.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        e.score_views(poses, 60, 80, 60.0, 60.0, 40.0, 30.0)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"max_chunk_rays {chunk:7d}: {512 / dt:7.1f} views/s")
    e.close()
