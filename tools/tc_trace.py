"""GPU-box diagnostic: timeline (SM clocks) of one tile through the tcgen05 MLP pipeline."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import evpp_b200                      # noqa: E402
from evpp_b200 import synthetic as S  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
torch.manual_seed(0)
c = S.init_nerf_state_dict()
eng = evpp_b200.Engine(device=0, precision=prec, near=0.5, far=6.0)
eng.load_weights(0, c)
eng.load_weights(1, c)
M = 148 * 128 * 6
pts = (torch.rand(M, 3) - 0.5) * 10
dirs = torch.nn.functional.normalize(torch.randn(M, 3), dim=-1)
for _ in range(2):
    t = eng.mlp_trace(0, pts, dirs, precision=prec).cpu().numpy()
t0 = t[0]
print("layer | MMA: start  first-issue  last-issue  commit | EPI: acc-ready  stored  window-done   (cycles from the tile's layer-0 start)")
for l in range(10):
    m = t[4 * l:4 * l + 4] - t0
    e = t[40 + 3 * l:40 + 3 * l + 3] - t0
    print(f"{l:5d} | {m[0]:8d} {m[1]:8d} {m[2]:8d} {m[3]:8d} | {e[0]:8d} {e[1]:8d} {e[2]:8d}")
print("tile total (layer-0 start -> layer-9 stored):", t[40 + 28] - t0)
print("accumulator halves final (monitor thread):")
for l in range(10):
    print(f"{l:5d} | half0 {t[70 + 2 * l] - t0:8d}  half1 {t[71 + 2 * l] - t0:8d}")
for name, b in (("warp 0", 90), ("warp 15", 95)):
    print(f"layer-3 drain, {name}: acc wait done {t[b] - t0}, chunk arrivals {[int(x - t0) for x in t[b + 1:b + 5]]}")
print("layer-3 chunk-0 arrival per epilogue warp:", [int(x - t0) for x in t[100:116]])
print("layer-4 issuer: weights of K-block 0 present at", int(t[116] - t0))
print("layer-4 K-block 0 weights: producer reached the stage at", int(t[117] - t0), "stage free at", int(t[118] - t0))
