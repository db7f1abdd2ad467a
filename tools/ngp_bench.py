"""GPU-box measurement of the NGP-mode scorer: views/s on the Room sweep shape (BASELINE configs[2]:
128x96 rays per view, occupancy skipping) and the hash-grid gather bandwidth (512 B per sample)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import evpp_b200                                  # noqa: E402
from evpp_b200 import ngp, synthetic as S         # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H, W, focal = 96, 128, 96.0
sc = evpp_b200.NGPScorer(device=0, H=H, W=W, focal=focal)
G = 128
# synthetic room shell + cube rasterised on the device side of the test helper (same rule as the test oracle)
amin, amax = np.array([-5.0, 0.01, -5.0]), np.array([5.0, 5.01, 6.0])
c = (np.arange(G) + 0.5) / G
X, Y, Z = np.meshgrid(amin[0] + c * 10, amin[1] + c * 5, amin[2] + c * 11, indexing="ij")
vox = (amax - amin) / G
lo, hi = np.array([-2.5, 0.01, -2.0]), np.array([2.5, 2.51, 3.0])
inside = (X > lo[0]) & (X < hi[0]) & (Y > lo[1]) & (Y < hi[1]) & (Z > lo[2]) & (Z < hi[2])
core = (X > lo[0] + vox[0]) & (X < hi[0] - vox[0]) & (Y > lo[1] + vox[1]) & (Y < hi[1] - vox[1]) & (Z > lo[2] + vox[2]) & (Z < hi[2] - vox[2])
occ = (inside & ~core) | ((np.abs(X) < 0.5) & (Y < 1.0) & (np.abs(Z - 0.5) < 0.5))
sc.set_occupancy(np.packbits(occ.reshape(-1), bitorder="little"))
p = ngp.init_params(0)
p["table"] = (p["table"].astype(np.float32) * 2000).astype(np.float16)
sc.load_params(p)
views = S.room_views(V, seed=0)
poses = torch.Tensor(np.asarray(S.views_to_poses(views)))[:, :3, :4].contiguous().cuda()
for i in range(3):
    torch.cuda.synchronize()
    n0 = sc.sample_count()
    t0 = time.perf_counter()
    s = sc.score_views(poses)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ns = sc.sample_count() - n0
    print(f"iter {i}: {V} views x {H * W} rays in {dt * 1e3:.1f} ms -> {V / dt:.0f} views/s, {V * H * W / dt / 1e6:.1f} Mrays/s, "
          f"{ns / dt / 1e6:.1f} Msamples/s ({ns / (V * H * W):.1f} samples/ray), gather {ns * 512 / dt / 1e9:.0f} GB/s algorithmic")
print("scores", s[:4].cpu().numpy())
# hash-grid encode alone on spatially coherent points (consecutive samples of rays) and on random points
M = 1 << 23
for name, u in (("random points", torch.rand(M, 3, device="cuda")),
                ("ray-coherent points", (torch.rand(M // 64, 1, 3, device="cuda") + torch.linspace(0, 0.1, 64, device="cuda")[None, :, None] *
                                         torch.nn.functional.normalize(torch.randn(M // 64, 1, 3, device="cuda"), dim=-1)).clamp(0, 1).reshape(-1, 3))):
    u = u.contiguous()
    for _ in range(2):
        sc.encode(u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        sc.encode(u)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"encode only, {name}: {M / ms / 1e3:.0f} Msamples/s, gather {M * 512 / ms / 1e6:.0f} GB/s algorithmic "
          f"(+{M * (12 + 128) / ms / 1e6:.0f} GB/s streamed in/out)")
