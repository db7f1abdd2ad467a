"""Next-best-view selection tail of the planner (plannerServer_Object/core/vpp.py:190-220,
interface.py:115-116) and the fast-pass + fp32 re-score protocol that keeps the selected view
identical to the fp32 reference when the bulk pass runs on fp16/bf16 tensor cores."""
import numpy as np
import torch

from ._lib import PRECISIONS
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


def contenders(scores, weight=None, margin=2e-3):
    """Sorted indices (int32) of the views whose weighted fast score is within ``margin`` (relative) of the best.
    With a bulk pass accurate to tol = margin / 2 per view (north_star: 1e-3), no view outside this band can be the
    true arg-max: true_v <= fast_v (1 + tol) < best_fast (1 - 2 tol)(1 + tol) <= true_leader (1 + tol)^2 (1 - 2 tol)
    < true_leader."""
    s = np.asarray(scores, dtype=np.float64)
    if weight is not None:
        s = s * np.asarray(weight, dtype=np.float64)
    best = np.max(s)
    return np.flatnonzero(s >= best - margin * abs(best)).astype(np.int32)


def score_and_select(engine, views, H, W, K, dirs_per_location=1, depth_clip=None, pix_idx=None, topk=0, margin=2e-3,
                     exact="fp16x3"):
    """Fast pass at the engine's bulk precision for all views, then the views that could still be the arg-max are
    scored again on the exact path (``exact``: "fp16x3" = split-fp16 tensor-core passes, fp32-class; "fp32" = the
    FFMA kernel) and the selection runs on the corrected scores, so that the next-best view is the fp32
    reference's (plannerServer_*/core/interface.py:115-116).

    Candidates = every view whose weighted fast score is within ``margin`` of the best (``contenders``) plus the
    ``topk`` best.  After the re-score the guard loop adds any view left outside whose weighted fast score, inflated
    by the bulk pass's tolerance margin / 2, still reaches the best exact score, and repeats until there is none:
    a non-candidate can therefore never win on an inflated 16-bit score.
    Returns dict(scores, data, nbv, winner, margin, rescored)."""
    views = np.asarray(views, dtype=np.float64)
    poses = torch.Tensor(np.asarray([get_pose(list(v[0:3]), v[3], v[4]) for v in views]))[:, :3, :4].contiguous()
    fx, fy, cx, cy = _intrinsics(K)
    pix_t = None if pix_idx is None else torch.as_tensor(np.ascontiguousarray(pix_idx, dtype=np.int32))
    scores = engine.score_views_host(poses, H, W, fx, fy, cx, cy, pix_t).clone().numpy().astype(np.float64)
    w = np.ones_like(scores) if depth_clip is None else np.asarray(depth_clip, dtype=np.float64)
    done = np.zeros(len(scores), dtype=bool)
    exact_code = PRECISIONS[exact] if isinstance(exact, str) else int(exact)
    if engine.precision not in (PRECISIONS["fp32"], exact_code) and len(scores) > 1:
        fast = scores.copy()
        cand = set(contenders(fast, w, margin).tolist())
        if topk:
            cand |= set(np.argsort(fast * w, kind="stable")[::-1][:min(topk, len(fast))].tolist())
        tol = 0.5 * margin
        while cand:
            ids = np.sort(np.fromiter(cand, dtype=np.int32))
            scores[ids] = engine.rescore_views(poses, ids, H, W, fx, fy, cx, cy, pix_t, precision=exact).numpy().astype(np.float64)
            done[ids] = True
            best = np.max((scores * w)[done])
            cand = set(np.flatnonzero(~done & (fast * w * (1.0 + tol) >= best)).tolist())
    data, nbv, win = select_nbv(views, scores, dirs_per_location, depth_clip)
    ws = np.sort((scores * w)[done]) if done.sum() > 1 else np.sort(scores * w)
    margin_out = float((ws[-1] - ws[-2]) / ws[-1]) if len(ws) > 1 else float("inf")
    return {"scores": scores, "data": data, "nbv": nbv, "winner": win, "margin": margin_out, "rescored": int(done.sum())}


# ---------------------------------------------------------------------------------------------------------------------
# surrogate g_phi: class MLP (plannerServer_*/core/vpp.py:26-67) and
# Viewpath_planner.train_only (vpp.py:619-654), fitted and queried on the GPU (evpp_b200/csrc/surrogate.cu)
# ---------------------------------------------------------------------------------------------------------------------
_SG_LAYERS = ["layer1"] + [f"layer2_{i}" for i in range(1, 9)] + ["layer3"]


class MLP:
    """Drop-in for the planners' ``MLP`` module: same parameter names (``state_dict()`` / ``load_state_dict()``),
    same default initialisation (torch.nn.Linear's, drawn in the reference's construction order, so an identical
    ``torch.manual_seed`` gives identical initial weights), ``mlp(x)`` for ``x`` of shape [3] or [M, 3].
    The parameters live on the engine's device as one packed vector; queries and the fit are CUDA kernels."""

    def __init__(self, engine, state_dict=None):
        self.engine = engine
        if state_dict is None:
            mods = [torch.nn.Linear(3, 64)] + [torch.nn.Linear(64, 64) for _ in range(8)] + [torch.nn.Linear(64, 1)]
            state_dict = {f"{n}.{p}": getattr(m, p).detach() for n, m in zip(_SG_LAYERS, mods) for p in ("weight", "bias")}
        self.params = torch.empty(engine.lib.evpp_surrogate_param_count(), device=engine.device, dtype=torch.float32)
        self.loss_curve = None
        self.load_state_dict(state_dict)

    def load_state_dict(self, sd):
        flat = torch.cat([torch.as_tensor(sd[f"{n}.{p}"]).detach().to(torch.float32).reshape(-1).cpu()
                          for n in _SG_LAYERS for p in ("weight", "bias")])
        if flat.numel() != self.params.numel():
            raise ValueError(f"state_dict holds {flat.numel()} values, the surrogate has {self.params.numel()}")
        self.params.copy_(flat)

    def state_dict(self):
        from collections import OrderedDict
        flat, sd, o = self.params.detach().cpu(), OrderedDict(), 0
        for n in _SG_LAYERS:
            fin, fout = (3, 64) if n == "layer1" else ((64, 1) if n == "layer3" else (64, 64))
            sd[f"{n}.weight"] = flat[o:o + fin * fout].reshape(fout, fin).clone(); o += fin * fout
            sd[f"{n}.bias"] = flat[o:o + fout].clone(); o += fout
        return sd

    def __call__(self, x):
        from .engine import _ptr, _stream_ptr
        x = torch.as_tensor(x, dtype=torch.float32)
        single = x.dim() == 1
        pts = x.reshape(-1, 3).to(self.engine.device).contiguous()
        out = torch.empty(pts.shape[0], device=self.engine.device, dtype=torch.float32)
        with torch.cuda.device(self.engine.device):
            _lib_check(self.engine.lib.evpp_surrogate_query(self.engine._h, _ptr(self.params), _ptr(pts), pts.shape[0],
                                                            _ptr(out), _stream_ptr(self.engine.device)))
        return out.reshape(1) if single else out.reshape(-1, 1)

    # the reference moves the module around with .to(device) / .cpu(); placement is the engine's here
    def to(self, *_a, **_k):
        return self

    def cpu(self):
        return self


def _lib_check(rc):
    from ._lib import check
    check(rc)


def train_only(mlp, data, N=100, epochs=300, lr=0.001):
    """``Viewpath_planner.train_only(mlp, data, N)`` (vpp.py:619-654): 300 full-batch Adam steps (lr 1e-3) on the MSE
    between ``mlp(data[:N, 0:3])`` and ``data[:N, 5]``; one kernel launch.  Returns ``mlp`` (trained in place);
    ``mlp.loss_curve`` holds the loss before each step (the reference's ``mlp_loss`` list)."""
    from .engine import _ptr, _stream_ptr
    data = np.asarray(data, dtype=np.float64)
    e = mlp.engine
    xyz = torch.tensor(np.ascontiguousarray(data[0:N, 0:3])).to(torch.float32).to(e.device).contiguous()
    gain = torch.tensor(np.ascontiguousarray(data[0:N, 5])).to(torch.float32).to(e.device).contiguous()
    loss = torch.empty(int(epochs), device=e.device, dtype=torch.float32)
    with torch.cuda.device(e.device):
        _lib_check(e.lib.evpp_surrogate_fit(e._h, _ptr(mlp.params), _ptr(xyz), _ptr(gain), int(N), int(epochs), float(lr),
                                            _ptr(loss), _stream_ptr(e.device)))
    mlp.loss_curve = loss
    return mlp


def get_mlp_maxdata(init_test_data, mlp, N_MAX=1):
    """The query half of ``Viewpath_planner.get_mlp_maxdata`` (vpp.py:223-244): evaluate the fitted surrogate on the
    sampled candidate locations ``init_test_data[:, 0:3]`` in one batched kernel launch, sort by predicted gain and
    return ``(data_all [N, 4] = (x, y, z, g), init_node [3] = mean of the N_MAX best locations, init_node_gain)``.
    Drawing the locations (``sample_location``: the planner's collision state) stays on the caller's side."""
    xyz = np.ascontiguousarray(np.asarray(init_test_data, dtype=np.float64)[:, 0:3])
    test_input = torch.tensor(xyz).to(torch.float32)
    g = mlp(test_input).detach().cpu()
    data_all = torch.cat((test_input, g), axis=1).numpy()
    data_sort = data_all[data_all[:, 3].argsort()]
    n = data_all.shape[0]
    init_node = np.array([float(np.mean(data_sort[n - N_MAX:, k])) for k in range(3)])
    init_node_gain = mlp(torch.tensor(init_node).to(torch.float32)).detach().cpu().numpy()[0]
    return data_all, init_node, init_node_gain


# ---------------------------------------------------------------------------------------------------------------------
# Viewpath_planner.sample_direction (plannerServer_Object/core/vpp.py:144-210, plannerServer_Room/core/vpp.py:141-200):
# the two-stage candidate scoring of one planning step, with both stages on the GPU scorers
# ---------------------------------------------------------------------------------------------------------------------
def direction_fan(locations, object_center, scene="object"):
    """The candidate directions of every location: Object 3 pitches x 5 yaws around the look-at direction
    (vpp.py:150-161), Room 5 pitches (around atan2(-dy, 0.6)) x 12 yaws over the full circle (Room vpp.py:147-159).
    Returns views [L * fan, 5] = (x, y, z, u, v) and the fan size."""
    locations = np.asarray(locations, dtype=np.float64)
    views = []
    for i in range(locations.shape[0]):
        x, y, z = locations[i, 0], locations[i, 1], locations[i, 2]
        dx, dy, dz = object_center[0] - x, object_center[1] - y, object_center[2] - z
        if scene == "object":
            u_center = np.arctan2(-dy, np.sqrt(dx ** 2 + dz ** 2))
            v_center = np.arctan2(dx, dz)
            us = np.linspace(u_center - 5 / 180 * np.pi, u_center + 5 / 180 * np.pi, 3)
            vs = np.linspace(v_center - 10 / 180 * np.pi, v_center + 10 / 180 * np.pi, 5)
        else:
            u_center = np.arctan2(-dy, 0.6)
            us = np.linspace(u_center - 30 / 180 * np.pi, u_center + 30 / 180 * np.pi, 5)
            vs = np.linspace((0) / 180 * np.pi, (360) / 180 * np.pi, 12)
        for u in us:
            for v in vs:
                if u < -np.pi / 2.0:
                    u = -np.pi / 2.0
                if u > np.pi / 2.0:
                    u = np.pi / 2.0
                views.append([x, y, z, u, v])
    fan = 15 if scene == "object" else 60
    return np.array(views), fan


def sample_direction(locations, tsdf, get_uncertainty, object_center, scene="object"):
    """Drop-in for ``Viewpath_planner.sample_direction(locations)``: stage 1 scores the whole direction fan with the
    TSDF occupancy ray cast (``tsdf.get_uncertainty_tsdf``, evpp_b200.TSDF), keeps the 3 best directions per location,
    stage 2 scores those with the NeRF uncertainty (``get_uncertainty(locations, us, vs) -> (uncers, _, _)``, the
    reference's interface2 signature: ``evpp_b200.views.planner_client(controller)`` builds it from a Controller), the
    Object scene weights by the depth clip, the best direction per location is kept and the gains are normalised by
    their maximum.  Returns data [L, 6] = (x, y, z, u, v, gain) like the reference."""
    locations = np.asarray(locations, dtype=np.float64)
    sample_num = locations.shape[0]
    views, fan = direction_fan(locations, object_center, scene)
    uncer, depth = tsdf.get_uncertainty_tsdf(views[:, 0:3].tolist(), views[:, 3].tolist(), views[:, 4].tolist())
    depth = np.array(depth)
    data_buff = np.concatenate((views, np.array(uncer).reshape(sample_num * fan, 1)), axis=1)
    select_views, select_depth = [], []
    for i in range(sample_num):
        dir_buff = data_buff[i * fan:(i + 1) * fan]
        depth_buff = depth[i * fan:(i + 1) * fan]
        index = (np.argsort(dir_buff[:, 5]))[::-1]
        for j in range(3):
            select_views.append(dir_buff[index[j]].tolist())
            select_depth.append(depth_buff[index[j]])
    select_views = np.array(select_views)
    select_depth = np.array(select_depth)
    uncer, _, _ = get_uncertainty(select_views[:, 0:3].tolist(), select_views[:, 3].tolist(), select_views[:, 4].tolist())
    uncer = np.array(uncer)
    if scene == "object":
        uncer = uncer * depth_clip_object(select_depth)      # the Room planner computes a clip and does not apply it
    uncer = np.array(uncer).reshape((sample_num * 3, 1))
    data_buff = np.concatenate((select_views[:, 0:5], uncer), axis=1)
    data = np.array([list(data_buff[i * 3:(i + 1) * 3][np.argmax(data_buff[i * 3:(i + 1) * 3][:, 5])])
                     for i in range(int(data_buff.shape[0] / 3))])           # get_maxdirdata, vpp.py:214-220
    data[:, 5] = data[:, 5] / np.max(data[:, 5])
    return data
