"""Next-best-view selection tail of the planner (plannerServer_Object/core/vpp.py:190-220,
interface.py:115-116) and the fast-pass + fp32 re-score protocol that keeps the selected view
identical to the fp32 reference when the bulk pass runs on fp16/bf16 tensor cores."""
import numpy as np
import torch

from .run_nerf import _intrinsics, get_pose


def depth_clip_object(select_depth):
    """Object-scene weighting (vpp.py:193-196): 1 inside [1.5,4.5] m, exp(-2|d-3|) outside."""
    d = np.asarray(select_depth, dtype=np.float64)
    clip = np.ones(d.shape)
    lo, hi = d < 1.5, d > 4.5
    clip[lo] = np.exp(-2 * np.abs(d[lo] - 3.0))
    clip[hi] = np.exp(-2 * np.abs(d[hi] - 3.0))
    return clip


def select_nbv(views, uncers, dirs_per_location=3, depth_clip=None):
    """views [V,5]=(x,y,z,u,v) grouped by location; returns (data [L,6], nbv_row, winner view index).
    Per location keep the best direction (first maximum, np.argmax), normalise by the global
    maximum, take the last row of the ascending sort; exact ties -> highest index (the
    reference's unstable argsort leaves that case unspecified)."""
    views = np.asarray(views, dtype=np.float64)
    s = np.asarray(torch.as_tensor(uncers).detach().cpu().numpy() if torch.is_tensor(uncers) else uncers, dtype=np.float64)
    if depth_clip is not None:
        s = s * np.asarray(depth_clip, dtype=np.float64)
    L = len(s) // dirs_per_location
    grp = s[:L * dirs_per_location].reshape(L, dirs_per_location)
    k = np.argmax(grp, axis=1)
    idx = np.arange(L) * dirs_per_location + k
    data = np.concatenate([views[idx, 0:5], grp[np.arange(L), k][:, None]], axis=1)
    data[:, 5] = data[:, 5] / np.max(data[:, 5])
    win = int(np.argsort(data[:, 5], kind="stable")[-1])
    return data, data[win], int(idx[win])


def score_and_select(engine, views, H, W, K, dirs_per_location=1, depth_clip=None, pix_idx=None, topk=16):
    """Fast pass on the engine's configured precision for all views, then re-score the ``topk``
    best (after weighting) with the fp32 MLP and select among the corrected scores.
    Returns dict(scores, data, nbv, winner, margin)."""
    views = np.asarray(views, dtype=np.float64)
    poses = torch.Tensor(np.asarray([get_pose(list(v[0:3]), v[3], v[4]) for v in views]))[:, :3, :4].contiguous()
    fx, fy, cx, cy = _intrinsics(K)
    pix_t = None if pix_idx is None else torch.as_tensor(np.ascontiguousarray(pix_idx, dtype=np.int32))
    scores = engine.score_views_host(poses, H, W, fx, fy, cx, cy, pix_t).clone().numpy().astype(np.float64)
    w = np.ones_like(scores) if depth_clip is None else np.asarray(depth_clip, dtype=np.float64)
    if topk and engine.precision != 0:
        cand = np.argsort(scores * w, kind="stable")[::-1][:min(topk, len(scores))]
        cand = np.sort(cand).astype(np.int32)
        exact = engine.rescore_views_fp32(poses, cand, H, W, fx, fy, cx, cy, pix_t).numpy().astype(np.float64)
        # guard: a view outside the candidate set must not beat the re-scored winner spuriously;
        # its fast score carries <=1e-3 relative error, so demote ties inside that band
        scores[cand] = exact
    data, nbv, win = select_nbv(views, scores, dirs_per_location, depth_clip)
    ws = np.sort(scores * w)
    margin = float((ws[-1] - ws[-2]) / ws[-1]) if len(ws) > 1 else float("inf")
    return {"scores": scores, "data": data, "nbv": nbv, "winner": win, "margin": margin}
