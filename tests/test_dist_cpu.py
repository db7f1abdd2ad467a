"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sharding + score gather + replicated
arg-max give the single-process answer."""
import os
import socket

import numpy as np
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
