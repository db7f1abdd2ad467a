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


def balanced_ranges(V, speeds):
    """Contiguous view blocks proportional to per-rank speeds (views / ms measured on each rank's own previous pass):
    the GPUs of one box settle at different clocks under the power cap (1-3 % apart), and a strong-scaled sweep ends
    when the slowest rank does.  Returns [(begin, end)] per rank covering [0, V) exactly."""
    sp = np.asarray(speeds, dtype=np.float64)
    if len(sp) == 0 or not np.all(np.isfinite(sp)) or np.any(sp <= 0):
        return [shard_range(V, r, len(sp)) for r in range(len(sp))]
    edges = np.rint(np.concatenate([[0.0], np.cumsum(sp / sp.sum())]) * V).astype(np.int64)
    edges[0], edges[-1] = 0, V
    edges = np.maximum.accumulate(edges)
    return [(int(edges[r]), int(edges[r + 1])) for r in range(len(sp))]


def gather_blocks(local, ranges, V, rank, world, group=None):
    """all_gather of per-rank score blocks of UNEQUAL sizes (ranges[r] = (begin, end) of rank r) -> full [V] tensor."""
    per = max(e - b for b, e in ranges)
    pad = torch.full((per,), float("-inf"), dtype=torch.float32, device=local.device)
    pad[:local.numel()] = local.to(torch.float32)
    out = torch.empty(world * per, dtype=torch.float32, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group) if hasattr(dist, "all_gather_into_tensor") and \
        local.device.type == "cuda" else _all_gather_list(out, pad, world, group)
    full = torch.empty(V, dtype=torch.float32, device=local.device)
    for r, (b, e) in enumerate(ranges):
        full[b:e] = out[r * per:r * per + (e - b)]
    return full


def exchange_speeds(local_views, local_ms, rank, world, device, group=None):
    """Every rank's bulk-pass speed (views / ms) on every rank: the input of balanced_ranges."""
    t = torch.zeros(world, dtype=torch.float64, device=device)
    t[rank] = (local_views / local_ms) if local_ms > 0 and local_views > 0 else 0.0
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def scatter_round_robin(ids, rank, world):
    """The share of ``ids`` (sorted global view ids of the contenders) that ``rank`` re-scores: every world-th one.
    Contenders cluster wherever the scene is uncertain, i.e. inside one rank's contiguous block of the sweep;
    dealing them out keeps the exact pass balanced (strong scaling) whatever their positions."""
    return ids[rank::world]


def sweep_select(engine, poses, H, W, fx, fy, cx, cy, rank=0, world=1, margin=2e-3, exact="fp16x3", weight=None,
                 dirs_per_location=1, group=None, host_out=None, ranges=None, timing=None):
    """One candidate sweep with identical-NBV guarantee, sharded over ``world`` ranks:

      1. bulk pass: rank r scores its contiguous block of the V poses at the engine's 16-bit precision;
      2. all_gather of the V scores (NCCL on the GPU box; 16 KB at V = 4096);
      3. contenders = views whose (weighted) score is within ``margin`` of the best (planner.contenders) -- the same
         set on every rank; they are dealt out round robin and re-scored on the tensor-core exact path (``exact``);
      4. all_gather of the exact scores, written over the fast ones;
      5. the planner's selection tail (vpp.py:199-220 + interface.py:115-116) on the device (evpp_select_nbv).

    poses: float32 [V,3,4] on the engine's device (already resident) or on the host (pinned: copied every call).
    ranges: optional [(begin, end)] per rank (balanced_ranges) instead of equal blocks; timing: optional dict that
    receives 'bulk_ms' (CUDA-event time of this rank's bulk pass) and 'bulk_views' for the next call's balancing.
    Returns (scores [V] on the device, winner view id, number of re-scored views).  With ``host_out`` (a pinned
    float32 [V + 1] tensor) the scores and the winner id are also landed in host memory before returning."""
    from .planner import contenders
    dev = engine.device
    V = poses.shape[0]
    b, e = ranges[rank] if ranges is not None else shard_range(V, rank, world)
    on_host = poses.device.type == "cpu"
    poses_dev = poses.to(dev, non_blocking=True) if on_host else poses
    ev = None
    if timing is not None and dev.type == "cuda":
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    local = engine.score_views(poses_dev[b:e], H, W, fx, fy, cx, cy) if e > b else torch.empty(0, device=dev)
    if ev is not None:
        ev[1].record()
    if world == 1:
        scores = local
    elif ranges is not None:
        scores = gather_blocks(local, ranges, V, rank, world, group)
    else:
        scores = gather_scores(local, V, rank, world, group)
    fast_host = scores.cpu().numpy().astype(np.float64)
    if ev is not None:
        timing["bulk_ms"], timing["bulk_views"] = ev[0].elapsed_time(ev[1]), e - b
    if not np.isfinite(fast_host).all():
        raise FloatingPointError("non-finite view scores in the bulk pass (16-bit activations overflowed): use precision='bf16'")
    band = contenders(fast_host, None if weight is None else weight.cpu().numpy(), margin) if margin > 0 else np.zeros(0, np.int32)
    n_band = len(band)
    if margin > 0 and engine.precision not in (0, 4) and V > 1:
        mine = scatter_round_robin(band, rank, world)
        per = (n_band + world - 1) // world
        ex = torch.full((per,), float("-inf"), device=dev)
        if len(mine):
            ids = torch.from_numpy(mine.astype(np.int64)).to(dev)
            ex[:len(mine)] = engine.score_views(poses_dev[ids], H, W, fx, fy, cx, cy, precision=exact)
        if world > 1:
            allx = torch.empty(world * per, device=dev)
            dist.all_gather_into_tensor(allx, ex, group=group) if dev.type == "cuda" else _all_gather_list(allx, ex, world, group)
            # allx[r * per + k] is the exact score of band[r + k * world]
            k = torch.arange(per, device=dev)[None, :].expand(world, per)
            r = torch.arange(world, device=dev)[:, None].expand(world, per)
            slot = (r + k * world).reshape(-1)
            ok = slot < n_band
            band_t = torch.from_numpy(band.astype(np.int64)).to(dev)
            scores = scores.clone()
            scores[band_t[slot[ok]]] = allx[ok]
        else:
            scores = scores.clone()
            scores[torch.from_numpy(band.astype(np.int64)).to(dev)] = ex[:n_band]
    _, _, win = engine.select_nbv(scores, weight, dirs_per_location)
    if host_out is not None:
        host_out[:V].copy_(scores, non_blocking=True)
        host_out[V:V + 1].copy_(win.to(torch.float32), non_blocking=True)
        torch.cuda.synchronize(dev)
        return scores, int(host_out[V].item()), n_band
    return scores, int(win.item()), n_band
