import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def weights():
    """Seed-0 random-init weights of the reference NeRF (coarse, fine), plain and spread variants."""
    from oracle import evpp_oracle as O
    torch.manual_seed(0)
    c = O.init_nerf_state_dict()
    f = O.init_nerf_state_dict()
    return {"plain": (c, f), "spread": (O.spread_variant(c), O.spread_variant(f))}


@pytest.fixture(scope="session")
def build_lib():
    from evpp_b200 import build
    return build.build()


_ENGINES = {}


@pytest.fixture(scope="session")
def engine_factory(build_lib, weights):
    """Engines cached per (variant, precision, far, N_importance) -- GPU tests only."""
    from evpp_b200 import Engine

    def make(variant="plain", precision="fp32", near=0.5, far=6.0, n_importance=128, max_chunk_rays=0):
        key = (variant, precision, near, far, n_importance, max_chunk_rays)
        if key not in _ENGINES:
            e = Engine(device=0, n_samples=64, n_importance=n_importance, near=near, far=far, white_bkgd=True,
                       precision=precision, max_chunk_rays=max_chunk_rays)
            c, f = weights[variant]
            e.load_weights(0, c)
            e.load_weights(1, f)
            _ENGINES[key] = e
        return _ENGINES[key]

    return make
