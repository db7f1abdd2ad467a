"""Short single-GPU run of the scoring hot path for ncu captures (see profiles/README.md)."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import evpp_b200                                  # noqa: E402
from evpp_b200 import synthetic as S              # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H, W, focal = 60, 80, 60.0
torch.manual_seed(0)
c, f = S.init_nerf_state_dict(), S.init_nerf_state_dict()
eng = evpp_b200.Engine(device=0, precision=prec, near=0.5, far=6.0)
eng.load_weights(0, c)
eng.load_weights(1, f)
poses = torch.Tensor(np.asarray(S.views_to_poses(S.hemisphere_views(V, seed=0))))[:, :3, :4].contiguous().cuda()
import time
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = eng.score_views(poses, H, W, focal, focal, 0.5 * W, 0.5 * H)
    torch.cuda.synchronize()
    print(f"iter {i}: {1e3 * (time.perf_counter() - t0):.2f} ms for {V} views -> {V / (time.perf_counter() - t0):.1f} views/s")
print("scores", s[:4].cpu().numpy(), "launches", eng.counters())
