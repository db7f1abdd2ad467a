"""evpp_b200 -- B200-native (sm_100a) candidate-view scoring engine for EVPP's planner hot path.

Drop-in for the reference's ``render(...)`` (nerfServer/core/run_nerf.py:131) and
``Controller.get_uncertainty(poses)`` (nerfServer/core/nerf.py:266); everything numerical runs in
the hand-written CUDA library behind the C ABI of include/evpp_b200.h.  No CPU fallback.
"""
from ._lib import EvppError, LIB_PATH  # noqa: F401
from .engine import Engine  # noqa: F401
from .run_nerf import render, render_path, get_rays, get_pose, get_poses, make_render_kwargs  # noqa: F401
from .nerf import Controller  # noqa: F401
from .planner import select_nbv, score_and_select, contenders, MLP, train_only, get_mlp_maxdata, sample_direction, direction_fan  # noqa: F401
from .tsdf import TSDF  # noqa: F401
from .ngp import NGPScorer, TensoRFScorer  # noqa: F401

__all__ = ["Engine", "EvppError", "render", "render_path", "get_rays", "get_pose", "get_poses", "make_render_kwargs", "Controller",
           "select_nbv", "score_and_select", "contenders", "MLP", "train_only", "get_mlp_maxdata", "sample_direction", "direction_fan", "TSDF", "NGPScorer", "TensoRFScorer"]
