"""GPU-box measurement: MLP kernel rate per precision / score-only variant on one 131072-ray chunk shape
(evaluations per second and effective TFLOP/s at the full network's 1,187,072 FLOP per evaluation)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evpp_b200 import Engine                       # noqa: E402
from evpp_b200 import synthetic as S               # noqa: E402

torch.manual_seed(0)
c, f = S.init_nerf_state_dict(), S.init_nerf_state_dict()
H, W, focal, V = 96, 128, 96.0, int(os.environ.get("V", 64))
views = S.room_views(V, seed=0)
poses = torch.Tensor(np.asarray(S.views_to_poses(views)))[:, :3, :4].contiguous().cuda()
for prec in sys.argv[1:] or ("fp16", "bf16", "fp16x2", "fp16x3"):
    for full in (1, 2, 0):
        eng = Engine(device=0, near=0.5, far=6.0, precision=prec)
        eng.load_weights(0, c)
        eng.load_weights(1, f)
        eng.set_score_mode(full)
        eng.score_views(poses, H, W, focal, focal, 0.5 * W, 0.5 * H)
        torch.cuda.synchronize()
        eng.set_profiling(True)
        t0 = time.perf_counter()
        for _ in range(3):
            s = eng.score_views(poses, H, W, focal, focal, 0.5 * W, 0.5 * H)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        ms, n, ev, fl = eng.mlp_profile()
        print(f"{prec:7s} { {1: 'full outputs ', 2: 'score-only   ', 0: 'score-only+fold'}[full] }: {V / dt:8.1f} views/s  step {1e3 * dt:8.1f} ms  MLP {ms / 3:8.1f} ms "
              f"({fl / ms / 1e9:7.1f} TFLOP/s executed, {ev * 1187072.0 / ms / 1e9:7.1f} at 1,187,072 FLOP/eval)  score[0] {float(s[0]):.6f}", flush=True)
        eng.close()
