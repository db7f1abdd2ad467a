"""Import the UNMODIFIED reference renderer: from /root/reference in the build container, from the byte-identical
vendored copy ``oracle/_ref/`` (written by ``oracle/build_ref.py``, git-ignored, shipped like the ``.so``) elsewhere.

TEST / BENCH INFRASTRUCTURE.  Used by ``oracle/make_golden*.py`` and the ``not gpu`` tests that pin
``oracle/evpp_oracle.py`` against the real code, and by ``bench.py --impl reference`` / its ``cpu_baseline`` leg to
time the reference itself on the host cores.  Never imported by the product, ``smoke()`` or the ``-m gpu`` tests.

Recipe: ``run_nerf.py`` imports imageio / matplotlib.pyplot /
colorlog and opens ``./logs/test.log`` at import time, so three empty stub modules are
placed in ``sys.modules`` and the import happens from a scratch cwd that has ``logs/``.
"""
import logging
import os
import sys
import tempfile
import types

REF_ROOTS = ["/root/reference/nerfServer",
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "nerfServer")]   # vendored copy (build_ref.py)
REF_ROOT = next((r for r in REF_ROOTS if os.path.isfile(os.path.join(r, "core", "run_nerf.py"))), REF_ROOTS[0])


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "core", "run_nerf.py"))


def source():
    """'tree' = /root/reference itself, 'vendored' = the byte-identical copy under oracle/_ref (GPU box), None."""
    if not available():
        return None
    return "tree" if REF_ROOT == REF_ROOTS[0] else "vendored"


def load():
    """Returns (run_nerf_helpers, run_nerf) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    for name in ["imageio", "matplotlib", "matplotlib.pyplot", "colorlog"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    class _Fmt(logging.Formatter):
        def __init__(self, fmt=None, *a, **k):
            super().__init__((fmt or "").replace("%(log_color)s", ""))

    sys.modules["colorlog"].ColoredFormatter = _Fmt
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="evpp_ref_")
    os.makedirs(os.path.join(scratch, "logs"), exist_ok=True)
    os.chdir(scratch)
    try:
        from core import run_nerf_helpers as H  # noqa
        from core import run_nerf as R  # noqa
    finally:
        os.chdir(cwd)
    import torch
    torch.autograd.set_detect_anomaly(False)  # switched on by the helpers at import (run_nerf_helpers.py:3)
    logging.getLogger().setLevel(logging.WARNING)
    return H, R


def make_reference_kwargs(H, R, sd_coarse, sd_fine, N_samples=64, N_importance=128, white_bkgd=True,
                          netchunk=1024 * 64):
    """Build the kwargs dict of create_nerf (run_nerf.py:316-333) around reference NeRF modules
    holding the given weights."""
    import torch
    embed_fn, input_ch, _ = H.get_embedder(10, 0, torch.device("cpu"))
    embeddirs_fn, input_ch_views, _ = H.get_embedder(4, 0, torch.device("cpu"))
    nets = []
    for sd in (sd_coarse, sd_fine):
        m = H.NeRF(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4],
                   input_ch_views=input_ch_views, use_viewdirs=True)
        m.load_state_dict(sd)
        nets.append(m)
    query = lambda inputs, viewdirs, network_fn: R.run_network(  # noqa: E731
        inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=netchunk)
    return {
        "network_query_fn": query, "perturb": False, "N_importance": N_importance,
        "network_fine": nets[1], "N_samples": N_samples, "network_fn": nets[0],
        "use_viewdirs": True, "white_bkgd": white_bkgd, "raw_noise_std": 0.,
        "device": torch.device("cpu"), "ndc": False, "lindisp": False,
    }
