"""GPU-box diagnostic: per-layer activations of the tcgen05 MLP against a torch fp32 restatement
(computed on the GPU with the same weights).  Usage: python tests/diag_tc_layers.py [fp16|bf16]"""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import evpp_oracle as O   # noqa: E402  (diagnostic tool, not the product path)
import evpp_b200                      # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
torch.manual_seed(0)
c = O.init_nerf_state_dict()
eng = evpp_b200.Engine(device=0, precision=prec, near=0.5, far=6.0)
eng.load_weights(0, c)
eng.load_weights(1, c)
torch.manual_seed(4)
M = 128 * 3 + 5
pts = (torch.rand(M, 3) - 0.5) * 10
dirs = torch.nn.functional.normalize(torch.randn(M, 3), dim=-1)
x = torch.cat([O.embed(pts, 10), O.embed(dirs, 4)], -1)
W = {k: v.double() for k, v in c.items()}
xp, xd = x[:, :63].double(), x[:, 63:].double()
acts = []
h = xp
for i in range(8):
    h = torch.relu(h @ W[f"pts_linears.{i}.weight"].T + W[f"pts_linears.{i}.bias"])
    acts.append(h)
    if i == 4:
        h = torch.cat([xp, h], -1)
feat = h @ W["feature_linear.weight"].T + W["feature_linear.bias"]
acts.append(feat)
hv = torch.relu(torch.cat([feat, xd], -1) @ W["views_linears.0.weight"].T + W["views_linears.0.bias"])
acts.append(hv)
with torch.no_grad():
    ref = O.nerf_forward(c, x).numpy()
for layer in range(10):
    out, a = eng.mlp_layer_dump(0, pts, dirs, layer, precision=prec)
    torch.cuda.synchronize()
    a = a.cpu().double()[:, :acts[layer].shape[1]]
    err = (a - acts[layer]).abs()
    scale = acts[layer].abs().max().item()
    bad_rows = (err.max(1).values > 0.02 * scale).sum().item()
    bad_cols = (err.max(0).values > 0.02 * scale).sum().item()
    print(f"layer {layer}: max|err| {err.max().item():.3e} (scale {scale:.3e}) rows>2% {bad_rows}/{M} cols>2% {bad_cols}/{a.shape[1]}"
          f"  finite={bool(torch.isfinite(a).all())}", flush=True)
y = out.cpu().numpy()
err = np.abs(y - ref).max(0) / (np.abs(ref).max(0) + 1e-6)
print("final out rel err per channel", err, "finite", np.isfinite(y).all())
print("sample out", y[:2], "ref", ref[:2])
