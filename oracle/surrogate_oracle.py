"""CPU restatement of the planners' surrogate g_phi.  TEST INFRASTRUCTURE ONLY: imported by
tests/ (and by oracle/make_golden_surrogate.py); the product path (evpp_b200/csrc/surrogate.cu) never touches it.

Follows, op for op:
  class MLP                      plannerServer_Object/core/vpp.py:26-67  (plannerServer_Room/core/vpp.py holds the same class)
  Viewpath_planner.train_only    plannerServer_Object/core/vpp.py:619-654

Parity status: PINNED by outputs of the reference itself -- `reference_symbols()` cuts the two definitions out of the
unmodified reference file with `ast` and executes them here; tests/test_oracle_cpu.py demands bit-equality between this
restatement and that code (same torch build, same seed), and tests/golden/surrogate.npz holds inputs, initial
parameters, the 300-entry loss list and the trained predictions produced by the reference code.

What "parity" can mean for the CUDA fit: a 300-step Adam trajectory is chaotic in the last bits -- the reference's own
fp32 run and the same code in fp64 end 0.06 apart in predictions (measured) -- so the GPU test pins the
algorithm where it is well conditioned (forward values, the first optimizer steps, the early loss curve) and bounds the
end state by that reference-intrinsic spread."""
import ast
import os
from collections import OrderedDict

import numpy as np
import torch

REF_VPP = "/root/reference/plannerServer_Object/core/vpp.py"
LAYERS = ["layer1"] + [f"layer2_{i}" for i in range(1, 9)] + ["layer3"]
USED_HIDDEN = (1, 3, 4, 5, 6, 7, 8)          # forward() skips layer2_2 (vpp.py:43-62)


class MLP(torch.nn.Module):
    """vpp.py:26-67"""

    def __init__(self):
        super().__init__()
        self.layer1 = torch.nn.Linear(3, 64)
        self.layer2_1 = torch.nn.Linear(64, 64)
        self.layer2_2 = torch.nn.Linear(64, 64)
        self.layer2_3 = torch.nn.Linear(64, 64)
        self.layer2_4 = torch.nn.Linear(64, 64)
        self.layer2_5 = torch.nn.Linear(64, 64)
        self.layer2_6 = torch.nn.Linear(64, 64)
        self.layer2_7 = torch.nn.Linear(64, 64)
        self.layer2_8 = torch.nn.Linear(64, 64)
        self.layer3 = torch.nn.Linear(64, 1)

    def forward(self, x):
        x = torch.nn.functional.relu(self.layer1(x))
        for i in USED_HIDDEN:
            x = torch.nn.functional.relu(getattr(self, f"layer2_{i}")(x))
        return self.layer3(x)


def train_only(mlp, data, N=100, epochs=300, lr=0.001):
    """vpp.py:619-654; returns (mlp, loss list).  data[:, 0:3] = location, data[:, 5] = normalised gain."""
    data = np.asarray(data)
    xyz = np.concatenate((data[0:N, 0].reshape(N, 1), data[0:N, 1].reshape(N, 1), data[0:N, 2].reshape(N, 1)), axis=1)
    input_x = torch.tensor(xyz).to(torch.float32)
    labels = torch.tensor(data[0:N, 5].reshape(N, 1)).to(torch.float32)
    opt = torch.optim.Adam(mlp.parameters(), lr=lr)
    losses = []
    for _ in range(epochs):
        opt.zero_grad()
        preds = mlp(input_x)
        loss = torch.nn.functional.mse_loss(preds, labels)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return mlp, losses


def pack_state_dict(sd):
    """state_dict -> flat float32 vector in the module's own order (the C ABI's `params`)."""
    return np.concatenate([np.asarray(sd[f"{l}.{p}"].detach().cpu().numpy(), dtype=np.float32).reshape(-1)
                           for l in LAYERS for p in ("weight", "bias")])


def unpack_state_dict(flat):
    flat = np.asarray(flat, dtype=np.float32)
    sd, o = OrderedDict(), 0
    for l in LAYERS:
        fin, fout = (3, 64) if l == "layer1" else ((64, 1) if l == "layer3" else (64, 64))
        sd[f"{l}.weight"] = torch.from_numpy(flat[o:o + fin * fout].reshape(fout, fin).copy()); o += fin * fout
        sd[f"{l}.bias"] = torch.from_numpy(flat[o:o + fout].copy()); o += fout
    assert o == flat.size
    return sd


def synthetic_gain_samples(seed=0, N=100):
    """One planning step's training rows in the layout the planners build (vpp.py:205-210): x, y, z, u, v, gain with
    the gain normalised by its maximum; locations on the Object planner's hemisphere shell."""
    rng = np.random.RandomState(seed)
    d = rng.normal(size=(N, 3))
    d[:, 1] = np.abs(d[:, 1])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    loc = np.array([0.0, 1.0, 0.0]) + d * rng.uniform(3.0, 4.5, size=(N, 1))
    g = 0.6 + 0.4 * np.sin(1.3 * loc[:, 0]) * np.cos(0.9 * loc[:, 2]) + 0.1 * rng.normal(size=N)
    g = np.clip(g, 0.02, None)
    data = np.zeros((N, 6))
    data[:, 0:3] = loc
    data[:, 3:5] = rng.uniform(-0.5, 0.5, size=(N, 2))
    data[:, 5] = g / g.max()
    return data


def reference_symbols():
    """(MLP class, train_only function) cut out of the unmodified reference source; None when /root/reference is absent."""
    if not os.path.exists(REF_VPP):
        return None
    tree = ast.parse(open(REF_VPP, encoding="utf-8").read())
    ns = {"torch": torch, "np": np, "time": __import__("time"), "device": torch.device("cpu"), "print": lambda *a, **k: None}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "MLP":
            exec(compile(ast.Module([node], []), REF_VPP, "exec"), ns)
        if isinstance(node, ast.ClassDef) and node.name == "Viewpath_planner":
            fn = [f for f in node.body if isinstance(f, ast.FunctionDef) and f.name == "train_only"][0]
            exec(compile(ast.Module([fn], []), REF_VPP, "exec"), ns)
    return ns["MLP"], ns["train_only"]
