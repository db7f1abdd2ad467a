"""Multi-GPU layer: one process per GPU, candidate views sharded in contiguous blocks, one small
all_gather of per-view scores (NCCL over NVLink on the GPU box, gloo in CPU tests), arg-max
replicated on every rank.  Views are independent, so there is no data-path collective besides
the score gather."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(V, rank, world):
    """Contiguous block [begin,end) of ceil(V/world) views for ``rank`` (last blocks may be short/empty)."""
    per = (V + world - 1) // world
    b = min(V, rank * per)
    return b, min(V, b + per)


def gather_scores(local_scores, V, rank=None, world=None, group=None):
    """all_gather of equally padded per-rank score blocks -> full [V] tensor on every rank."""
    world = dist.get_world_size(group) if world is None else world
    per = (V + world - 1) // world
    pad = torch.full((per,), float("-inf"), dtype=torch.float32, device=local_scores.device)
    pad[:local_scores.numel()] = local_scores.to(torch.float32)
    out = torch.empty(world * per, dtype=torch.float32, device=local_scores.device)
    dist.all_gather_into_tensor(out, pad, group=group) if hasattr(dist, "all_gather_into_tensor") and \
        local_scores.device.type == "cuda" else _all_gather_list(out, pad, world, group)
    return out[:V]


def _all_gather_list(out, pad, world, group):
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out.copy_(torch.cat(parts))


def score_views_sharded(score_fn, poses, rank, world, device="cpu", group=None):
    """score_fn(poses_block) -> scores tensor for that block.  Returns the full [V] scores (every rank)."""
    V = len(poses)
    b, e = shard_range(V, rank, world)
    local = score_fn(poses[b:e]) if e > b else torch.empty(0, dtype=torch.float32, device=device)
    return gather_scores(local.to(device), V, rank, world, group)


def replicated_argmax(scores):
    """Highest index among exact ties, like the host tail (planner.select_nbv with 1 dir/location)."""
    s = scores.detach().cpu().numpy().astype(np.float64)
    return int(np.argsort(s, kind="stable")[-1])
