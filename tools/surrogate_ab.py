"""GPU-box check of the surrogate fit: the cluster kernels (16 CTAs = the default where the device can place such a
cluster, 8 CTAs; slices exchanged through distributed shared memory) against the single-CTA kernel (EVPP_SG_CLUSTER=0):
parameters and loss curves must be bit-identical; prints the fit times.  The switch is read once per process, so
every variant runs in its own process."""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(tag):
    import evpp_b200
    from oracle import surrogate_oracle as SO
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "surrogate.npz")))
    eng = evpp_b200.Engine(device=0, precision="fp32")
    out = {}
    for n, epochs in ((100, 300), (128, 7), (37, 12), (1, 3)):
        m = evpp_b200.MLP(eng, SO.unpack_state_dict(g["init_params"]))
        data = np.tile(g["data"], (2, 1))[:max(n, 1)]
        ms = []
        for _ in range(3):
            m.load_state_dict(SO.unpack_state_dict(g["init_params"]))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            evpp_b200.train_only(m, data, N=n, epochs=epochs)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        out[f"params_{n}_{epochs}"] = m.params.cpu().numpy()
        out[f"loss_{n}_{epochs}"] = m.loss_curve.cpu().numpy()
        print(f"[{tag}] n={n} epochs={epochs}: {min(ms):.3f} ms (best of 3), final loss {out[f'loss_{n}_{epochs}'][-1]:.6g}")
    np.savez(os.path.join(ROOT, "gpurun_out", f"surrogate_{tag}.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for tag, env in (("default", {}), ("cluster16", {"EVPP_SG_CLUSTER": "16"}), ("cluster8", {"EVPP_SG_CLUSTER": "8"}),
                     ("single", {"EVPP_SG_CLUSTER": "0"})):
        env_ = {k: v for k, v in os.environ.items() if k != "EVPP_SG_CLUSTER"}
        subprocess.run([sys.executable, os.path.abspath(__file__), tag], env=dict(env_, **env), check=True)
    b = np.load(os.path.join(ROOT, "gpurun_out", "surrogate_single.npz"))
    ok = True
    for tag in ("default", "cluster16", "cluster8"):
        a = np.load(os.path.join(ROOT, "gpurun_out", f"surrogate_{tag}.npz"))
        for k in a.files:
            same = np.array_equal(a[k], b[k])
            ok &= same
            print(tag, k, "bit-identical" if same else f"DIFFERENT: max abs {np.abs(a[k] - b[k]).max():.3e}")
    print("SURROGATE A/B", "OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
