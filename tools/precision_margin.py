"""GPU-box measurement: max relative error of the per-view scores against the reference's golden outputs for the
16-bit operand types of the tensor-core path (tolerance in north_star: 1e-3)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evpp_b200 import Engine                       # noqa: E402
from oracle import evpp_oracle as O                # noqa: E402

torch.manual_seed(0)
c0, f0 = O.init_nerf_state_dict(), O.init_nerf_state_dict()
W8 = {"plain": (c0, f0), "spread": (O.spread_variant(c0), O.spread_variant(f0))}
G = os.path.join(ROOT, "tests", "golden")
for prec in ("fp32", "fp16x3", "fp16x2", "fp16", "bf16"):
    for name, variant, far in (("render_small_plain", "plain", 6.0), ("render_small_spread", "spread", 6.0),
                               ("score_far80_plain", "plain", 80.0), ("score_far80_spread", "spread", 80.0),
                               ("config1_spread", "spread", 6.0)):
        g = np.load(os.path.join(G, name + ".npz"))
        eng = Engine(device=0, near=0.5, far=far, precision=prec)
        eng.load_weights(0, W8[variant][0])
        eng.load_weights(1, W8[variant][1])
        H, W = int(g["H"]), int(g["W"])
        K = O.make_K(H, W, float(g["focal"]))
        poses = g["poses"] if "poses" in g.files else np.asarray(O.views_to_poses(g["views"]))
        poses = torch.Tensor(poses)[:, :3, :4].contiguous()
        pix = torch.from_numpy(g["pix"]).contiguous() if "pix" in g.files else None
        s = eng.score_views_host(poses, H, W, K[0][0], K[1][1], K[0][2], K[1][2], pix_idx_host=pix).numpy()
        rel = np.abs(s - g["uncers"]) / np.abs(g["uncers"])
        print(f"{prec} {name:22s} V={len(s):3d} max rel err {rel.max():.2e}  mean {rel.mean():.2e}  argmax same: {int(np.argmax(s)) == int(np.argmax(g['uncers']))}")
        eng.close()
