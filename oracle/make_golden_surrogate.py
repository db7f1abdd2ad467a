"""Writes tests/golden/surrogate.npz from the UNMODIFIED reference code (class MLP and Viewpath_planner.train_only cut out
of /root/reference/plannerServer_Object/core/vpp.py), after asserting that oracle/surrogate_oracle.py reproduces it bit
for bit.  Run in the build container: python oracle/make_golden_surrogate.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import surrogate_oracle as SO   # noqa: E402

RefMLP, ref_train_only = SO.reference_symbols()
data = SO.synthetic_gain_samples(0)
torch.manual_seed(0)
ref = RefMLP()
init = SO.pack_state_dict(ref.state_dict())
torch.manual_seed(0)
mine = SO.MLP()
assert np.array_equal(SO.pack_state_dict(mine.state_dict()), init)
torch.set_num_threads(1)
x = torch.tensor(data[:, 0:3]).to(torch.float32)
pred0 = ref(x).detach().numpy().reshape(-1)
# the reference keeps its loss list local; the restatement returns it -- identical trajectories are checked on the weights
ref = ref_train_only(None, ref, data, N=100)
mine, losses = SO.train_only(mine, data, N=100)
assert np.array_equal(SO.pack_state_dict(ref.state_dict()), SO.pack_state_dict(mine.state_dict())), "restatement != reference"
pred = ref(x).detach().numpy().reshape(-1)
# one- and ten-step checkpoints (well conditioned: used for the tight GPU comparisons)
ck = {}
for steps in (1, 10):
    torch.manual_seed(0)
    m = RefMLP()
    m2, l2 = SO.train_only(m, data, N=100, epochs=steps)
    ck[f"params_after_{steps}"] = SO.pack_state_dict(m2.state_dict())
out = os.path.join(ROOT, "tests", "golden", "surrogate.npz")
np.savez_compressed(out, data=data, init_params=init, pred_init=pred0, loss_curve=np.asarray(losses, dtype=np.float64),
                    params_trained=SO.pack_state_dict(ref.state_dict()), pred_trained=pred, **ck)
print("wrote", out, "loss", losses[0], "->", losses[-1])
