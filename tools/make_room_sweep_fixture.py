"""GPU-box tool: builds tests/golden/room_sweep_nbv.json, the stored next-best view of bench.py's default workload
(the 4096-view Room sweep, 128x96 rays, random-init seed 0).

  1. all 4096 views on the tensor-core exact path (fp16x3) and on the bulk path (fp16);
  2. the 32 best of (1) again with the fp32 FFMA kernel (the numerical anchor, 2e-5 of the reference);
  3. the N_ORACLE best on the HOST with the unmodified reference render() (oracle/_ref; else the oracle port):
     the fp32 oracle's own scores of the leaders.  A view outside the N_ORACLE best cannot be the oracle's arg-max
     when its exact-path score trails the leader by more than the exact path's error bound (2e-5) -- checked here.

Writes gpurun_out/room_sweep_nbv.json (copy it to tests/golden/).  TEST INFRASTRUCTURE (imports oracle/)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                   # noqa: E402
from evpp_b200 import Engine                                   # noqa: E402

N_ORACLE = int(os.environ.get("N_ORACLE", 8))
c, f, views, poses = bench.make_workload(0, 1)
V, H, W, focal = bench.V_TOTAL, bench.IMG_H, bench.IMG_W, bench.FOCAL
intr = (focal, focal, 0.5 * W, 0.5 * H)
eng = Engine(device=0, near=bench.NEAR, far=bench.FAR, precision="fp16")
eng.load_weights(0, c)
eng.load_weights(1, f)
pd = poses.cuda()
t0 = time.time()
fast = eng.score_views(pd, H, W, *intr).cpu().numpy().astype(np.float64)
t1 = time.time()
x3 = eng.score_views(pd, H, W, *intr, precision="fp16x3").cpu().numpy().astype(np.float64)
t2 = time.time()
order = np.argsort(x3, kind="stable")[::-1]
top32 = np.sort(order[:32]).astype(np.int32)
f32 = eng.rescore_views(poses, top32, H, W, *intr, precision="fp32").numpy().astype(np.float64)
t3 = time.time()
print(f"fp16 {t1 - t0:.1f} s, fp16x3 {t2 - t1:.1f} s, fp32 x32 {t3 - t2:.1f} s")
rel = np.abs(fast - x3) / x3
print(f"fp16 vs fp16x3 over {V} views: max rel {rel.max():.2e}, mean {rel.mean():.2e}")
print(f"fp32 vs fp16x3 over the 32 leaders: max rel {np.abs(f32 - x3[top32]).max() / x3.max():.2e}")
lead = x3[order[0]]
for m in (2e-3, 1e-3, 5e-4, 2e-4):
    print(f"views within {m:g} of the leader: fp16x3 {int((x3 >= lead * (1 - m)).sum())}, fp16 {int((fast >= fast.max() * (1 - m)).sum())}")
print("score spread: min %.4f median %.4f max %.4f" % (x3.min(), np.median(x3), x3.max()))
print("leaders (id, fp16x3, gap to #1):", [(int(i), round(float(x3[i]), 5), f"{(lead - x3[i]) / lead:.2e}") for i in order[:10]])

ref = bench.HostReference(c, f, views)
ids = [int(i) for i in order[:N_ORACLE]]
t0 = time.time()
orc = ref.score(ids, None).numpy().astype(np.float64)
print(f"host oracle ({ref.kind}) on {N_ORACLE} leaders: {time.time() - t0:.1f} s")
win = ids[int(np.argsort(orc, kind='stable')[-1])]
err = np.abs(orc - x3[ids]) / orc
print("oracle vs fp16x3 on the leaders: max rel", f"{err.max():.2e}", "| oracle winner", win, "| fp16x3 winner", int(order[0]))
bound = 2e-5
outside_gap = (lead - x3[order[N_ORACLE]]) / lead
out = {"source": f"host {ref.kind} (unmodified reference render() when 'reference') on the {N_ORACLE} leaders of the fp16x3 sweep; "
                 f"view #{N_ORACLE + 1} trails the leader by {outside_gap:.2e} relative on the exact path (error bound {bound:g})",
       "views": V, "H": H, "W": W, "focal": focal, "near": bench.NEAR, "far": bench.FAR, "seed": 0,
       "winner": win, "oracle_ids": ids, "oracle_scores": [float(x) for x in orc],
       "fp16x3_scores_of_oracle_ids": [float(x3[i]) for i in ids],
       "fp32_ffma_ids": [int(i) for i in top32], "fp32_ffma_scores": [float(x) for x in f32],
       "fp16x3_top64_ids": [int(i) for i in order[:64]], "fp16x3_top64_scores": [float(x3[i]) for i in order[:64]],
       "leaders_outside_oracle_set_are_out_of_reach": bool(outside_gap > 2 * bound),
       "max_rel_fp16_vs_fp16x3": float(rel.max())}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "room_sweep_nbv.json"), "w"), indent=1)
np.save(os.path.join(ROOT, "gpurun_out", "room_sweep_scores.npy"), np.stack([fast, x3]))
print("wrote gpurun_out/room_sweep_nbv.json")
