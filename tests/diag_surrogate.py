"""GPU-box diagnostic: deviations of the CUDA surrogate fit from the reference run in tests/golden/surrogate.npz."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import evpp_b200
from oracle import surrogate_oracle as SO
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "surrogate.npz")))
eng = evpp_b200.Engine(device=0, precision="fp32")
data = g["data"]
mlp = evpp_b200.MLP(eng, SO.unpack_state_dict(g["init_params"]))
p0 = mlp(data[:, 0:3]).cpu().numpy().reshape(-1)
print("forward max abs diff", np.abs(p0 - g["pred_init"]).max(), "single", mlp(data[0, 0:3]).cpu().numpy(), p0[0])
for steps in (1, 10):
    m = evpp_b200.MLP(eng, SO.unpack_state_dict(g["init_params"]))
    evpp_b200.train_only(m, data, N=100, epochs=steps)
    d = np.abs(m.params.cpu().numpy() - g[f"params_after_{steps}"])
    lc = m.loss_curve.cpu().numpy()
    print(f"{steps} steps: params max abs diff {d.max():.3e}, mean {d.mean():.3e}; loss rel diff max", np.abs(lc / g["loss_curve"][:steps] - 1).max())
m = evpp_b200.MLP(eng, SO.unpack_state_dict(g["init_params"]))
for i in range(3):
    m.load_state_dict(SO.unpack_state_dict(g["init_params"]))
    torch.cuda.synchronize(); t0 = time.time()
    evpp_b200.train_only(m, data, N=100)
    torch.cuda.synchronize(); print("fit 300 epochs: %.2f ms" % (1e3 * (time.time() - t0)))
lc = m.loss_curve.cpu().numpy()
pred = m(data[:, 0:3]).cpu().numpy().reshape(-1)
print("loss first/last", lc[0], lc[-1], "reference", g["loss_curve"][0], g["loss_curve"][-1])
print("loss curve rel diff at epochs 10,30,100,299:", [float(abs(lc[i] / g["loss_curve"][i] - 1)) for i in (10, 30, 100, 299)])
print("trained predictions max abs diff vs reference run", np.abs(pred - g["pred_trained"]).max(), "rms", np.sqrt(np.mean((pred - g["pred_trained"]) ** 2)))
xs = torch.rand(5000, 3) * 4 - 2
torch.cuda.synchronize(); t0 = time.time(); q = m(xs); torch.cuda.synchronize(); print("5000 queries: %.3f ms" % (1e3 * (time.time() - t0)))
ref = SO.MLP(); ref.load_state_dict(m.state_dict())
print("query vs torch forward of the same weights", (q.cpu() - ref(xs).detach()).abs().max().item())
