"""Multi-GPU layer: one process per GPU, candidate views sharded in contiguous blocks, one small
all_gather of per-view scores (NCCL over NVLink on the GPU box, gloo in CPU tests), arg-max
replicated on every rank.  Views are independent, so there is no data-path collective besides
the score gather.  The full-resolution render sweep (BASELINE configs[4]) shards the RAYS of
every image instead: one all_gather of the per-ray maps and one all_reduce of the per-view partial sums of beta^2."""
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


def render_sharded(render_fn, rays_o, rays_d, rank, world, group=None):
    """Ray-sharded full-image render (render_path's caller shape, run_nerf.py:199-241): rank r renders the
    contiguous block of rays ``shard_range(N, r, world)``; the per-ray maps are all-gathered so that every rank
    holds the full image, and the per-image sum of beta^2 is the sum of the shards' partial sums.
    render_fn(rays_o_block, rays_d_block) -> dict of per-ray tensors ([n] or [n,k]); returns the same dict for all
    N rays."""
    N = rays_o.shape[0]
    b, e = shard_range(N, rank, world)
    per = (N + world - 1) // world
    local = render_fn(rays_o[b:e], rays_d[b:e]) if e > b else None
    keys = sorted(local.keys()) if local is not None else None
    if world > 1:
        holder = [keys]
        src = 0
        dist.broadcast_object_list(holder, src=src, group=group)     # rank 0 always owns a non-empty block
        keys = holder[0]
    out = {}
    for k in keys:
        t = local[k] if local is not None else None
        if world == 1:
            out[k] = t
            continue
        shape_tail = list(t.shape[1:]) if t is not None else None
        holder = [shape_tail, str(t.dtype) if t is not None else None]
        dist.broadcast_object_list(holder, src=0, group=group)
        shape_tail, dtype = holder[0], getattr(torch, holder[1].split(".")[-1])
        dev = t.device if t is not None else rays_o.device
        pad = torch.zeros([per] + shape_tail, dtype=dtype, device=dev)
        if t is not None:
            pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[k] = torch.cat(parts)[:N]
    return out


def render_views_ray_sharded(render_block, V, n_pixels, rank, world, device="cpu", group=None):
    """Full-resolution sweep, ray-sharded: rank r renders pixels
    ``shard_range(n_pixels, r, world)`` of EVERY one of the V views.

    render_block(b, e) -> float32 [V, e-b, C] per-ray maps of pixel block [b, e); the LAST channel is the ray's
    sum over samples of beta^2.  Returns (maps [V, n_pixels, C] on every rank, scores [V] float32) with
    scores = sum over all rays and samples of beta^2 / n_pixels (Controller.get_uncertainty, nerf.py:307-309):
    one all_gather of the padded blocks and one all_reduce(sum) of the fp64 per-view partial sums."""
    b, e = shard_range(n_pixels, rank, world)
    per = (n_pixels + world - 1) // world
    local = render_block(b, e) if e > b else None
    if world == 1:
        return local, (local[..., -1].sum(1, dtype=torch.float64) / n_pixels).to(torch.float32)
    C = torch.tensor([local.shape[-1] if local is not None else 0], device=device)
    dist.all_reduce(C, op=dist.ReduceOp.MAX, group=group)
    C = int(C.item())
    pad = torch.zeros(V, per, C, dtype=torch.float32, device=device)
    if local is not None:
        pad[:, :e - b] = local
    if pad.device.type == "cuda":
        out = torch.empty(world, V, per, C, dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.stack(parts)
    maps = out.permute(1, 0, 2, 3).reshape(V, world * per, C)[:, :n_pixels]
    partial = pad[..., -1].sum(1, dtype=torch.float64)        # the padding rows are zero
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return maps, (partial / n_pixels).to(torch.float32)
