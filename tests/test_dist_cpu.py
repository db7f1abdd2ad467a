"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sharding + score gather + replicated
arg-max give the single-process answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from evpp_b200 import dist as edist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, V, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    poses = torch.arange(V, dtype=torch.float32)

    def score_fn(block):                       # stand-in scorer: deterministic function of the pose
        return torch.sin(block * 0.37) * 10 + block * 0.01

    full = edist.score_views_sharded(score_fn, poses, rank, world)
    q.put((rank, full.numpy(), edist.replicated_argmax(full)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly():
    for V in (0, 1, 7, 512, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [edist.shard_range(V, r, world) for r in range(world)]
            covered = [i for b, e in spans for i in range(b, e)]
            assert covered == list(range(V))


def test_two_rank_gather_matches_single_process():
    V = 37                                     # ragged: 19 + 18
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, V, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    poses = torch.arange(V, dtype=torch.float32)
    ref = (torch.sin(poses * 0.37) * 10 + poses * 0.01).numpy()
    for rank, full, win in res:
        np.testing.assert_array_equal(full, ref)
        assert win == int(np.argmax(ref))


def _render_worker(rank, world, port, N, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ro = torch.arange(N * 3, dtype=torch.float32).reshape(N, 3)
    rd = torch.flip(ro, [0])

    def render_fn(o, d):                      # stand-in renderer: per-ray maps as functions of the ray
        return {"depth": o[:, 0] * 2 + d[:, 1], "rgb": torch.stack([o[:, 0], d[:, 0], o[:, 2] - d[:, 2]], -1),
                "beta2": (o[:, 1] * 0.5) ** 2}

    out = edist.render_sharded(render_fn, ro, rd, rank, world)
    q.put((rank, {k: v.numpy() for k, v in out.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ray_sharded_render_matches_single_process():
    """BASELINE configs[4] caller shape (full-resolution render, ray-sharded): shards reassemble to the full maps."""
    N = 53
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_render_worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ro = torch.arange(N * 3, dtype=torch.float32).reshape(N, 3)
    rd = torch.flip(ro, [0])
    ref = {"depth": (ro[:, 0] * 2 + rd[:, 1]).numpy(), "rgb": torch.stack([ro[:, 0], rd[:, 0], ro[:, 2] - rd[:, 2]], -1).numpy(),
           "beta2": ((ro[:, 1] * 0.5) ** 2).numpy()}
    for rank, out in res:
        for k in ref:
            np.testing.assert_array_equal(out[k], ref[k])


def test_wire_shim_round_trip():
    """views.get_uncertainty wire format (nerfServer/core/views.py:94-113) against a stand-in controller."""
    import json
    from evpp_b200 import views

    class FakeController:
        def get_uncertainty(self, poses):
            assert np.asarray(poses).shape == (2, 4, 4)
            return torch.tensor([float(np.asarray(p)[0][3]) for p in poses])

        def get_rays_depth(self, locations, directions):
            return torch.tensor([1.5] * len(locations))

    body = json.dumps({"locations": [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], "us": [0.1, 0.2], "vs": [0.3, 0.4]})
    assert views.get_uncertainty(FakeController(), body) == {"uncers": [1.0, 4.0]}
    assert views.get_ray_depth(FakeController(), json.dumps({"locations": [[0, 0, 0]], "directions": [[0, 0, 1]]})) == {"depths": [1.5]}
    import pytest
    with pytest.raises(ValueError):
        views.get_uncertainty(FakeController(), body, method="GET")


def _sweep_block(V, b, e):
    """Stand-in per-ray maps of pixels [b, e) of V views: [V, e-b, 3] = (depth-like, acc-like, beta^2 sum)."""
    pix = torch.arange(b, e, dtype=torch.float32)[None, :]
    v = torch.arange(V, dtype=torch.float32)[:, None]
    return torch.stack([pix * 0.5 + v, (pix + 1) / (v + 2), (0.01 * pix + 0.1 * v) ** 2], -1)


def _sweep_worker(rank, world, port, V, n_pixels, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    maps, scores = edist.render_views_ray_sharded(lambda b, e: _sweep_block(V, b, e), V, n_pixels, rank, world)
    q.put((rank, maps.numpy(), scores.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pixels", [37, 2])
def test_two_rank_ray_sharded_view_sweep(n_pixels):
    """BASELINE configs[4]: every rank renders a pixel block of every view; maps are all-gathered, the per-view
    score is the all-reduced sum of beta^2 / rays (odd pixel count: the last block is short)."""
    V = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, V, n_pixels, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = _sweep_block(V, 0, n_pixels)
    ref_scores = (full[..., -1].sum(1, dtype=torch.float64) / n_pixels).to(torch.float32).numpy()
    one_maps, one_scores = edist.render_views_ray_sharded(lambda b, e: _sweep_block(V, b, e), V, n_pixels, 0, 1)
    np.testing.assert_array_equal(one_maps.numpy(), full.numpy())
    for rank, maps, scores in res:
        np.testing.assert_array_equal(maps, full.numpy())
        np.testing.assert_allclose(scores, ref_scores, rtol=1e-6)
        np.testing.assert_allclose(scores, one_scores.numpy(), rtol=1e-6)


def test_planner_client_is_the_callable_sample_direction_expects():
    """views.planner_client: interface2.get_uncertainty's call shape (locations, us, vs) -> (uncers, [], []) over the
    JSON bodies of the two services (plannerServer_*/core/interface2.py:46-60, nerfServer/core/views.py:94-113)."""
    import evpp_b200
    from evpp_b200 import views

    class FakeController:
        def get_uncertainty(self, poses):
            p = np.asarray(poses)
            return torch.tensor(1.0 + p[:, 0, 3] ** 2 + 0.1 * p[:, 1, 1])

    client = views.planner_client(FakeController())
    u, a, b = client([[1.0, 2.0, 3.0], [0.5, 0.0, 1.0]], [0.1, 0.2], [0.3, 0.4])
    assert a == [] and b == [] and len(u) == 2 and abs(u[0] - (2.0 + 0.1 * np.cos(0.1))) < 1e-6

    class Tsdf:
        @staticmethod
        def get_uncertainty_tsdf(locs, us, vs):
            a = np.asarray(locs)
            return list(0.5 + 0.5 * np.sin(a[:, 0] + np.asarray(us))), list(2.0 + np.cos(np.asarray(vs)))

    locs = np.random.RandomState(1).uniform(-1, 1, (5, 3))
    data = evpp_b200.sample_direction(locs, Tsdf, client, [0.0, 1.0, 0.0], scene="room")
    assert data.shape == (5, 6) and data[:, 5].max() == 1.0


class _SweepEngine:
    """Host stand-in for the CUDA engine in sweep_select: a bulk pass with a bounded relative error, an exact pass,
    and the selection tail (first maximum per location, ties of the global maximum -> highest index)."""
    precision = 1          # fp16
    device = torch.device("cpu")

    def __init__(self):
        self.exact_calls = []

    @staticmethod
    def _truth(p):
        x = p[:, 0, 3].double()
        return 100.0 * (1.0 + 3e-4 * torch.sin(x * 12.9898)) * torch.where(x % 5 == 0, 1.0, 0.9)

    def score_views(self, poses, H, W, fx, fy, cx, cy, pix=None, precision=None):
        t = self._truth(poses)
        if precision is None:
            t = t * (1.0 + 4e-4 * torch.cos(poses[:, 0, 3].double() * 78.233))      # 16-bit pass: |error| <= 4e-4
        else:
            self.exact_calls.append(poses[:, 0, 3].long().tolist())
        return t.float()

    def select_nbv(self, scores, weight, dirs):
        s = scores.double() * (1.0 if weight is None else weight.double())
        best = np.flatnonzero(s.numpy() == s.numpy().max())[-1]
        return None, None, torch.tensor([int(best)], dtype=torch.int32)


def _sweep_select_worker(rank, world, port, V, q, speeds=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    poses = torch.zeros(V, 3, 4)
    poses[:, 0, 3] = torch.arange(V, dtype=torch.float32)
    eng = _SweepEngine()
    ranges = None
    if speeds is not None:        # unequal, speed-proportional blocks; the speeds travel through the collective
        sp = edist.exchange_speeds(speeds[rank], 1.0, rank, world, "cpu")
        ranges = edist.balanced_ranges(V, sp)
        assert ranges[0][0] == 0 and ranges[-1][1] == V and ranges[0][1] == ranges[1][0]
    scores, win, band = edist.sweep_select(eng, poses, 6, 8, 6.0, 6.0, 4.0, 3.0, rank, world, margin=2e-3, ranges=ranges)
    q.put((rank, scores.numpy(), win, band, eng.exact_calls))
    dist.barrier()
    dist.destroy_process_group()


def test_balanced_ranges_cover_exactly_and_follow_the_speeds():
    for V in (1, 7, 4096):
        for sp in ([1.0], [1.0, 1.0], [1.0, 1.03, 0.98, 1.0], [3.0, 1.0], [1.0] * 8):
            r = edist.balanced_ranges(V, sp)
            assert r[0][0] == 0 and r[-1][1] == V and all(r[i][1] == r[i + 1][0] for i in range(len(r) - 1))
            assert all(e >= b for b, e in r)
    r = edist.balanced_ranges(4096, [1.0, 1.02])
    assert r[1][1] - r[1][0] > r[0][1] - r[0][0] and abs((r[1][1] - r[1][0]) / (r[0][1] - r[0][0]) - 1.02) < 2e-3
    assert edist.balanced_ranges(10, [0.0, 1.0]) == [edist.shard_range(10, 0, 2), edist.shard_range(10, 1, 2)]   # bad timing: equal blocks


@pytest.mark.parametrize("V,speeds", [(41, None), (200, None), (200, (1.0, 1.7))])
def test_two_rank_sweep_select_matches_single_process_and_exact_argmax(V, speeds):
    """dist.sweep_select at world size 2 (gloo): bulk pass sharded in blocks, contenders dealt out round robin for
    the exact pass, both gathers, replicated selection -- every rank ends with the single-process scores and with
    the arg-max of the EXACT scores although the leaders are closer than the bulk pass's error."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_select_worker, args=(r, 2, port, V, q, speeds)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    poses = torch.zeros(V, 3, 4)
    poses[:, 0, 3] = torch.arange(V, dtype=torch.float32)
    one = _SweepEngine()
    s1, w1, b1 = edist.sweep_select(one, poses, 6, 8, 6.0, 6.0, 4.0, 3.0, 0, 1, margin=2e-3)
    truth = _SweepEngine._truth(poses).float().numpy()
    assert w1 == int(np.flatnonzero(truth == truth.max())[-1])
    fast = one.score_views(poses, 6, 8, 6.0, 6.0, 4.0, 3.0).numpy()
    assert 1 < b1 < V and int(np.argmax(fast)) != w1 or True          # (the bulk arg-max may or may not differ)
    seen = []
    for rank, scores, win, band, calls in res:
        np.testing.assert_array_equal(scores, s1.numpy())
        assert win == w1 and band == b1
        seen += [i for c in calls for i in c]
    assert sorted(seen) == sorted(i for c in one.exact_calls for i in c)      # every contender re-scored exactly once overall
    assert abs(len(res[0][4][0]) - len(res[1][4][0])) <= 1                    # ... and evenly over the ranks
