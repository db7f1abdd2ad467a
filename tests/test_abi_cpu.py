"""CPU tests: the C-ABI library builds, loads and exports every symbol the header declares; it
refuses to work without a GPU (no CPU fallback); host-side logic of the drop-in layer."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import evpp_oracle as O


def test_library_exports_every_declared_symbol(build_lib):
    from evpp_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "evpp_b200.h")).read()
    declared = set(re.findall(r"\b(evpp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.evpp_abi_version() == 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(build_lib):
    import evpp_b200
    with pytest.raises(evpp_b200.EvppError):
        evpp_b200.Engine()
    from evpp_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    cfg = _lib.EvppConfig(64, 128, 0.5, 6.0, 1, 0, 0, 0)
    h = C.c_void_p()
    rc = lib.evpp_create(0, C.byref(cfg), C.byref(h))
    assert rc == -4 and b"no CPU fallback" in lib.evpp_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "evpp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_planner_tail_matches_oracle():
    from evpp_b200 import planner, run_nerf
    rng = np.random.RandomState(0)
    for dirs in (1, 3, 4):
        views = rng.randn(12 * dirs, 5)
        s = rng.rand(12 * dirs) * 100
        s[5] = s[7] = s.max() + 1     # exact tie
        clip = O.depth_clip_object(rng.rand(12 * dirs) * 6)
        for dc in (None, clip):
            d0, n0, w0 = O.select_nbv(views, s, dirs, dc)
            d1, n1, w1 = planner.select_nbv(views, s, dirs, dc)
            assert w0 == w1
            np.testing.assert_array_equal(d0, d1)
            np.testing.assert_array_equal(n0, n1)
    np.testing.assert_array_equal(planner.depth_clip_object([1.0, 2.0, 5.0]), O.depth_clip_object([1.0, 2.0, 5.0]))
    for (u, v) in [(0.3, -0.7), (1.2, 2.9)]:
        np.testing.assert_array_equal(np.array(run_nerf.get_pose([1, 2, 3], u, v)), np.array(O.get_pose([1, 2, 3], u, v)))


def test_render_rejects_unsupported_modes():
    from evpp_b200 import render
    K = O.make_K(4, 4, 3.0)
    with pytest.raises(NotImplementedError):
        render(4, 4, K, rays=(torch.zeros(1, 3), torch.ones(1, 3)), ndc=True, use_viewdirs=True)
    with pytest.raises(NotImplementedError):
        render(4, 4, K, rays=(torch.zeros(1, 3), torch.ones(1, 3)), ndc=False, use_viewdirs=False)
    with pytest.raises(NotImplementedError):
        render(4, 4, K, rays=(torch.zeros(1, 3), torch.ones(1, 3)), ndc=False, use_viewdirs=True, perturb=1.0)


def test_get_mlp_maxdata_host_logic_matches_reference_lines():
    """planner.get_mlp_maxdata (vpp.py:223-244) over a stand-in module on the CPU: the same lines as the reference's
    method (sort by predicted gain, mean of the N_MAX best locations, gain of that location)."""
    from evpp_b200 import planner
    from oracle import surrogate_oracle as SO
    torch.manual_seed(3)
    mlp = SO.MLP()
    cand = np.random.RandomState(1).uniform(-4, 4, size=(200, 6))
    data_all, init_node, init_gain = planner.get_mlp_maxdata(cand, mlp)
    # the reference's lines
    N_test = cand.shape[0]
    xyz = np.concatenate((cand[:, 0].reshape(N_test, 1), cand[:, 1].reshape(N_test, 1), cand[:, 2].reshape(N_test, 1)), axis=1)
    test_input = torch.tensor(xyz).to(torch.float32)
    g = mlp(test_input).detach()
    ref_all = torch.cat((test_input, g), axis=1).cpu().numpy()
    ref_sort = ref_all[ref_all[:, 3].argsort()]
    ref_node = np.array([float(np.mean(ref_sort[N_test - 1:, k])) for k in range(3)])
    ref_gain = mlp(torch.tensor(ref_node).to(torch.float32)).detach().cpu().numpy()[0]
    assert np.array_equal(data_all, ref_all) and np.array_equal(init_node, ref_node) and init_gain == ref_gain
    _, node3, _ = planner.get_mlp_maxdata(cand, mlp, N_MAX=3)
    np.testing.assert_allclose(node3, ref_sort[N_test - 3:, 0:3].mean(0), rtol=1e-6)


def test_render_path_passes_the_reference_arguments_through():
    """render_path (run_nerf.py:199-241) forwards to render(): unsupported modes are refused the same way, before any
    device work."""
    from evpp_b200 import render_path
    K = O.make_K(4, 4, 3.0)
    with pytest.raises(NotImplementedError):
        render_path([np.eye(4)], (4, 4, 3.0), K, 1024, {"ndc": True, "use_viewdirs": True})
    assert render_path([], (4, 4, 3.0), K, 1024, {"ndc": False, "use_viewdirs": True}) == ([], [], [], [])
