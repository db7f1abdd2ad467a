#!/usr/bin/env python
"""Benchmark of the candidate-view scoring hot path (BASELINE.json metric: candidate views scored/sec).

  python bench.py --gpus N --steps K --warmup W            # the CUDA engine (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own renderer on the host CPU cores

A step = one pass of the hot path over the north-star workload, configs[2] of BASELINE.json: the 4096-view Room
sweep (1024 positions x 4 directions of the Room fan, 128x96 rays per view, 64+128 samples per ray, random-init
reference NeRF x2) in reference-parity mode (frequency encoding + 8x256 NeRF + NeurAR head -- the only scorer the
reference implements).  The 4096 views are split over the N ranks (STRONG scaling).  A step
is everything the planner needs from the sweep, i.e. an NBV identical to the fp32 reference's:

  bulk pass (fp16 tensor cores, score-only network outputs) -> all_gather of the 4096 scores -> every view within
  2e-3 of the leader re-scored on the tensor-core exact path (fp16x3), dealt out over the ranks -> all_gather ->
  the planner's selection tail on the device -> winner,

and the winner is checked against tests/golden/room_sweep_nbv.json (the fp32 oracle's arg-max over the leaders).
`--workload object` keeps round 1's weak-scaled configs[1] (512 views x 80x60 rays per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_EVAL = 1187072.0
V_TOTAL, IMG_H, IMG_W, FOCAL = 4096, 96, 128, 96.0       # room: views of the whole sweep; object: views PER GPU
WORKLOAD = "room"
NEAR, FAR = 0.5, 6.0
S_C, S_I = 64, 128
EVALS_PER_RAY = S_C + (S_C + S_I)
MARGIN = 2e-3
NBV_FIXTURE = os.path.join(ROOT, "tests", "golden", "room_sweep_nbv.json")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 1400.0, 1590.0, "fallback"


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_workload(rank, world=1):
    """Weights (seed 0) and the candidate views.  room: the SAME 4096 views on every rank (seed 0; each rank scores a
    block of them); object: 512 views per rank, seeded by the rank (weak scaling)."""
    from evpp_b200 import synthetic as S
    torch.manual_seed(0)
    c = S.init_nerf_state_dict()
    f = S.init_nerf_state_dict()
    views = S.room_views(V_TOTAL, seed=0) if WORKLOAD == "room" else S.hemisphere_views(V_TOTAL, seed=rank)
    poses = torch.Tensor(np.asarray(S.views_to_poses(views)))[:, :3, :4].contiguous()
    return c, f, views, poses


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


class HostReference:
    """The reference path on the host cores: the UNMODIFIED reference render() (oracle/_ref or /root/reference,
    kind 'reference') when it can be imported, else the oracle port (bit-identical to it on CPU, kind 'port').
    score(view_ids, n_pix) = Controller.get_uncertainty's arithmetic (nerf.py:278-309) on a pixel subset."""

    def __init__(self, c, f, views):
        from oracle import evpp_oracle as O, ref_import
        self.O, self.views, self.c, self.f = O, views, c, f
        self.K = O.make_K(IMG_H, IMG_W, FOCAL)
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.kind, self.kwargs, self.R = "port", None, None
        if ref_import.available():
            try:
                H, R = ref_import.load()
                self.kwargs = ref_import.make_reference_kwargs(H, R, c, f, N_samples=S_C, N_importance=S_I, white_bkgd=True)
                self.R, self.Hm = R, H
                self.kind = "reference"
            except Exception as e:      # e.g. cv2 missing on the box: fall back to the port, say so
                print(f"bench: reference import failed ({e}); timing the oracle port", file=sys.stderr)
        self.desc = ("unmodified reference render() (nerfServer/core/run_nerf.py:131, vendored by oracle/build_ref.py) + sum(beta^2)/R (nerf.py:307-309)"
                     if self.kind == "reference" else "oracle port, bit-identical to the reference's render() on CPU (oracle/make_golden.py)")

    def score(self, ids, pix):
        """ids: view indices; pix: int [len(ids), R] flat pixel ids (or None = full grid).  Returns scores [len(ids)]."""
        O = self.O
        poses = O.views_to_poses([self.views[i] for i in ids])
        with torch.no_grad():
            if self.kind == "port":
                return O.score_views(poses, IMG_H, IMG_W, self.K, NEAR, FAR, self.c, self.f, pix_idx=pix)
            rays = O.build_view_rays(poses, IMG_H, IMG_W, self.K, pix)           # get_rays + row subset, nerf.py:282-296
            R = rays.shape[1] // len(ids)
            out = self.R.render(IMG_H, IMG_W, self.K, chunk=1024 * 32, rays=rays, retraw=True, dev=torch.device("cpu"),
                                near=NEAR, far=FAR, **{k: v for k, v in self.kwargs.items() if k not in ("near", "far")})
            unc = out[3] ** 2
            return torch.sum(unc.reshape(len(ids), R * (S_C + S_I)), -1) / R


def reference_sample(step):
    """Bounded sample of the sweep for one host step: 32 views of the workload x 1/32 of each view's pixels = one
    view-equivalent of rays (work is exactly linear in rays: views and rays share no state)."""
    n_pix = IMG_H * IMG_W
    rng = np.random.RandomState(1000 + step)
    ids = rng.choice(V_TOTAL, size=32, replace=False)
    per = n_pix // 32
    pix = np.stack([rng.choice(n_pix, size=per, replace=False) for _ in ids]).astype(np.int32)
    return ids, pix, 32 * per / n_pix


def cpu_reference_rate(c, f, views, seconds_budget=20.0):
    """cpu_baseline leg of the default line: the reference on a bounded sample (about 20 s), all host threads."""
    ref = HostReference(c, f, views)
    ids, pix, veq = reference_sample(0)
    t0 = time.time()
    ref.score(ids, pix)
    t1 = time.time() - t0
    n = int(max(1, min(8, seconds_budget // max(t1, 1e-3))))
    best, tot = 1e30, 0.0
    for k in range(n):
        ids, pix, veq = reference_sample(1 + k)
        t0 = time.time()
        ref.score(ids, pix)
        dt = time.time() - t0
        best, tot = min(best, dt), tot + dt
    return veq * n / tot, ref.threads, ref.kind, (
        f"{n} samples of {int(veq * IMG_H * IMG_W)} rays (32 views of the {V_TOTAL}-view sweep x {IMG_H * IMG_W // 32} pixels each = {veq:.2f} "
        f"view-equivalents), 64+128 samples/ray, {tot:.1f} s wall (warm-up {t1:.1f} s, best sample {best:.1f} s); {ref.desc}; "
        f"{ref.threads} threads of {cpu_model()}; rate extrapolated linearly in rays")


def run_reference(args, rank, world):
    if rank != 0:
        return
    c, f, views, _ = make_workload(0, world)
    ref = HostReference(c, f, views)
    on_gpu = args.ref_device == "cuda"
    if on_gpu:
        # informational: the reference's torch path (same ops, same order) with its tensors on cuda:0
        import warnings
        from oracle import evpp_oracle as O
        warnings.simplefilter("ignore")
        torch.set_default_device("cuda")
        torch.set_default_tensor_type(torch.cuda.FloatTensor)      # the legacy torch.Tensor(...) constructors too
        c = {k: v.cuda() for k, v in c.items()}
        f = {k: v.cuda() for k, v in f.items()}
        ref.kind, ref.c, ref.f = "port", c, f
        ref.desc = "oracle port (the reference's torch ops in the reference's order) with every tensor on cuda:0"
    times, veq = [], 1.0
    for s in range(args.warmup + args.steps):
        ids, pix, veq = reference_sample(s)
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.time()
        sc = ref.score(ids, pix)
        if on_gpu:
            sc = sc.cpu()
        if s >= args.warmup:
            times.append(time.time() - t0)
    ms = 1e3 * float(np.mean(times))
    val = veq / (ms / 1e3)
    best3 = veq / min(times[:3]) if times else None
    sample = (f"per step 32 views of the {V_TOTAL}-view workload x {IMG_H * IMG_W // 32} pixels each = {veq:.2f} view-equivalents of rays "
              f"({int(veq * IMG_H * IMG_W)} rays x 256 MLP evaluations), rate extrapolated linearly in rays (BASELINE.md section 3); {ref.desc}; "
              f"{ref.threads} threads of {cpu_model()}; best of the first 3 timed steps: {best3:.4f} views/s"
              + ("; THIS RUN: the same torch ops on cuda:0, not the host" if on_gpu else ""))
    line = {"impl": "reference", "metric": "candidate views scored/sec", "value": val, "unit": "views/s",
            "rays_per_s": val * IMG_H * IMG_W, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if WORKLOAD == "room" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, "reference torch path on cuda:0 (informational)" if on_gpu else "host CPU, torch fp32"),
            "cpu_baseline": {"value": val, "unit": "views/s", "cores": ref.threads, "kind": ref.kind, "sample": sample,
                             "cpu": cpu_model(), "best_of_3": best3},
            "e2e": {"value": val, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ngp(args, rank, local_rank, world):
    """NGP-mode scorer on the Room sweep shape: 512 views x 128x96 rays per GPU (4096 views at 8 GPUs), synthetic
    room-shell occupancy grid, 16-level 2^19 hash grid, random-init parameters.  Roofline = hash-grid gather."""
    import torch.distributed as dist
    import evpp_b200
    from evpp_b200 import dist as edist, ngp, synthetic as S
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    dev = torch.device("cuda", local_rank)
    H, W, focal, V = 96, 128, 96.0, 512
    tensorf = args.mode == "tensorf"
    if tensorf:      # BASELINE configs[3]: VM-decomposed grids (res 300, 16 + 48 components) instead of the hash grid
        sc = evpp_b200.TensoRFScorer(device=local_rank, H=H, W=W, focal=focal)
        sc.set_occupancy(S.room_occupancy_bitfield(128))
        sc.load_params(ngp.init_tensorf_params(0, R=300))
    else:
        sc = evpp_b200.NGPScorer(device=local_rank, H=H, W=W, focal=focal)
        sc.set_occupancy(S.room_occupancy_bitfield(128))
        p = ngp.init_params(0)
        p["table"] = (p["table"].astype(np.float32) * 2000).astype(np.float16)
        sc.load_params(p)
    bytes_per_sample = 2304 if tensorf else 512      # 3 x (4 plane + 2 line corners) x 128 B  |  16 levels x 8 corners x 4 B
    views = S.room_views(V, seed=rank)
    poses_host = torch.Tensor(np.asarray(S.views_to_poses(views)))[:, :3, :4].contiguous()
    poses_dev, poses_pinned = poses_host.to(dev), poses_host.pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    V_total = V * world

    def run(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms, wall, n0, l0 = 0.0, 0.0, sc.sample_count(), sc.engine.counters()[0]
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            wall += time.perf_counter() - t0
            ms += e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item() / steps, t[1].item() / steps, (sc.sample_count() - n0) / steps, (sc.engine.counters()[0] - l0)

    def step_dev():
        s = sc.score_views(poses_dev)
        return edist.gather_scores(s, V_total, rank, world) if world > 1 else s

    poses44 = np.concatenate([poses_host.numpy(), np.tile([[[0, 0, 0, 1.0]]], (V, 1, 1))], 1)

    def step_e2e():
        s = sc.get_uncertainty(poses44)
        return edist.gather_scores(s, V_total, rank, world).cpu() if world > 1 else s.cpu()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sc.engine.set_profiling(True)
    ms_dev, _, samples, launches = run(step_dev, args.steps, args.warmup)
    field_ms, field_launches, _, _ = sc.engine.mlp_profile()          # CUDA events around the field kernel launches
    sc.engine.set_profiling(False)
    field_ms_per_step = field_ms / max(1, args.steps + args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    _, ms_e2e, _, _ = run(step_e2e, max(2, min(args.steps, 3)), 1)
    if rank == 0:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        gather_gbs = samples * bytes_per_sample / (ms_dev / 1e3) / 1e9
        field_gbs = samples * bytes_per_sample / (field_ms_per_step / 1e3) / 1e9 if field_ms_per_step > 0 else None
        # ceilings measured on this chip for 4-byte gathers from an L2-resident table (tools/gather_bw.cu + the stand-alone
        # encode kernel, profiles/r1_microbench.txt): the hash-grid table (24.5 MB) and the VM grids (34.6 MB) live in the
        # 126 MB L2, so HBM bandwidth is not the bound of this mode
        peaks_g = {"random_pairs_GBs": 1763.0, "ray_coherent_GBs": 3411.0, "source": "profiles/r1_microbench.txt (440.8 G random-pair "
                   "gathers/s x 4 B; stand-alone encode kernel on ray-coherent points 3411 GB/s)"}
        # TensoRF reads whole 128-byte lines (one per bilinear corner) that neighbouring samples share through L1: no
        # gather-rate microbenchmark applies; its line is set against the measured HBM copy bandwidth, the only measured
        # memory peak a coalesced-line stream can be compared with (the 34.6 MB of grids are L2-resident, so the
        # algorithmic rate may approach or exceed it: the kernel is bound by L1/L2 and instruction issue, not by HBM)
        peak_gather = d.get("hbm_gbs", 6650.0) if tensorf else peaks_g["ray_coherent_GBs"]
        line = {"metric": "candidate views scored/sec", "value": V_total / (ms_dev / 1e3), "unit": "views/s",
                "rays_per_s": V_total * H * W / (ms_dev / 1e3), "samples_per_s": samples * world / (ms_dev / 1e3),
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16 tables and MLP operands, f32 accumulate", "data": "synthetic",
                "config": {"workload": (f"BASELINE configs[3] TensoRF variant of the Room sweep (no reference counterpart): {V} views x {W}x{H} rays "
                                        "per GPU, 128^3 occupancy bitfield, VM grids res 300 with 16+48 components, dt 0.02 m") if tensorf else
                                       (f"BASELINE configs[2] Room scene shape, NGP mode (no reference counterpart): {V} views x {W}x{H} rays per "
                                        "GPU, 128^3 occupancy bitfield (room shell + cube), 16-level 2^19 hash grid, dt 0.02 m"),
                           "views_per_gpu": V, "rays_per_view": H * W, "samples_per_ray": samples / (V * H * W),
                           "l2": "256 MB L2 flush between timed steps"},
                "e2e": {"value": V_total / (ms_e2e / 1e3), "unit": "views/s", "h2d_bytes_per_step": V * 48, "d2h_bytes_per_step": V * 4,
                        "api": "NGPScorer.get_uncertainty == evpp_ngp_score_views_host"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": field_gbs if field_gbs else gather_gbs, "peak": peak_gather, "unit": "GB/s",
                             "frac": (field_gbs if field_gbs else gather_gbs) / peak_gather, "traffic": None,
                             "kernel": "tensorf_field_kernel" if tensorf else "ngp_field_tc_kernel (hash-grid gathers + small MLP on warp-level tensor-core MMA)",
                             "achieved_whole_step": gather_gbs, "achieved_field_kernel_only": field_gbs,
                             "field_kernel_ms_per_step": field_ms_per_step, "field_share_of_step": field_ms_per_step / ms_dev if ms_dev > 0 else None,
                             "peak_definition": ("measured HBM copy bandwidth (MEASURED_PEAKS.json), for scale only: the 34.6 MB of grids are L2-resident and neighbouring samples share their 128-byte lines through L1 (ncu, profiles/r2_tensorf_ncu_metrics.txt: L1 hit rate 87 %, L2 throughput 4 %, DRAM < 1 %), so the algorithmic rate can exceed it; the kernel is bound by instruction latency at 16 warps per SM (issue-active 48 %), not by memory" if tensorf else
                                                 "memory-system bound, but by the L2 gather rate, not HBM: measured rate of the stand-alone encode kernel on "
                                                 "ray-coherent points (the access pattern of this step)"), "peaks_measured": peaks_g,
                             "frac_of_random_pair_rate": (field_gbs if field_gbs else gather_gbs) / peaks_g["random_pairs_GBs"],
                             "hbm_peak_GBs": d.get("hbm_gbs", 6650.0),
                             "note": "algorithmic feature-gather bytes (hash grid: 16 levels x 8 corners x 4 B = 512 B, VM grids: 2304 B per kept sample) / "
                                     "CUDA-event time of the field kernel; the kernel also evaluates the MLP and writes 20 B per sample"},
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_fullres(args, rank, local_rank, world):
    """BASELINE configs[4]: full-resolution uncertainty render sweep (640x480 rays per view, 64+128 samples), the rays
    of every image sharded over the ranks (strong scaling: a step renders the same 8 views of the 256-view sweep at
    any N).  Per-ray rgb / depth / acc / sum(beta^2) maps are all-gathered, per-view scores all-reduced."""
    import torch.distributed as dist
    import evpp_b200
    from evpp_b200 import dist as edist, synthetic as S
    from __graft_entry__ import build
    if local_rank == 0:
        build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    dev = torch.device("cuda", local_rank)
    H, W, focal, V = 480, 640, 480.0, 8
    n_pix = H * W
    torch.manual_seed(0)
    c, f = S.init_nerf_state_dict(), S.init_nerf_state_dict()
    views = S.hemisphere_views(256, seed=0)[:V]
    poses_host = torch.Tensor(np.asarray(S.views_to_poses(views)))[:, :3, :4].contiguous()
    eng = evpp_b200.Engine(device=local_rank, n_samples=S_C, n_importance=S_I, near=NEAR, far=FAR, white_bkgd=True,
                           precision=args.precision)
    eng.load_weights(0, c)
    eng.load_weights(1, f)
    poses_dev, poses_pinned = poses_host.to(dev), poses_host.pin_memory()
    maps_pinned = torch.empty(V, n_pix, 6, dtype=torch.float32).pin_memory()
    scores_pinned = torch.empty(V, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sweep(poses):
        def block(b, e):
            pix = torch.arange(b, e, dtype=torch.int32, device=dev)[None, :].expand(V, e - b).contiguous()
            ro, rd, _ = eng.get_rays(poses, H, W, focal, focal, 0.5 * W, 0.5 * H, pix)
            o = eng.render_rays(ro, rd, NEAR, FAR, want=("rgb", "depth", "acc", "unc"))
            beta2 = o["unc"].square().sum(-1, keepdim=True)          # per-ray sum over the 192 samples
            return torch.cat([o["rgb"], o["depth"][:, None], o["acc"][:, None], beta2], -1).view(V, e - b, 6)
        return edist.render_views_ray_sharded(block, V, n_pix, rank, world, device=dev)

    def step_device():
        return sweep(poses_dev)

    def step_e2e():
        maps, scores = sweep(poses_pinned.to(dev, non_blocking=True))
        maps_pinned.copy_(maps, non_blocking=True)
        scores_pinned.copy_(scores, non_blocking=True)
        torch.cuda.synchronize()
        return scores_pinned

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if profile:
            eng.set_profiling(True)
        l0, _ = eng.counters()
        total_ms, wall = 0.0, 0.0
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            wall += time.perf_counter() - t0
            total_ms += ev0.elapsed_time(ev1)
        l1, _ = eng.counters()
        t = torch.tensor([total_ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item() / steps, t[1].item() / steps, l1 - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, _, launches = timed(step_device, args.steps, args.warmup, profile=True)
    mlp_ms, mlp_launches, mlp_evals, mlp_flops = eng.mlp_profile()
    eng.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    _, ms_e2e, _ = timed(step_e2e, max(2, min(args.steps, 3)), 1)
    if rank == 0:
        sustained, _, src = peaks()
        # the same numbers must come out at every N (rays are independent): fp64 sums of the gathered maps and scores
        check = {"scores": [float(x) for x in scores_pinned[:4]], "maps_sum": float(maps_pinned.sum(dtype=torch.float64)),
                 "depth_sum": float(maps_pinned[..., 3].sum(dtype=torch.float64))}
        flops_per_launch = mlp_flops / max(1, mlp_launches)        # executed: render() folds feature_linear into the view layer
        avg_launch_ms = mlp_ms / max(1, mlp_launches)
        achieved = flops_per_launch / (avg_launch_ms / 1e3) / 1e12 if avg_launch_ms > 0 else 0.0
        bpe, _ = measured_traffic_per_eval()
        line = {"metric": "candidate views scored/sec", "value": V / (ms_dev / 1e3), "unit": "views/s",
                "rays_per_s": V * n_pix / (ms_dev / 1e3), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16 operands, f32 accumulate (tcgen05)" if args.precision == "fp16" else args.precision, "data": "synthetic",
                "config": {"workload": f"BASELINE configs[4] full-resolution uncertainty render sweep: {V} of the 256 views per step x {W}x{H} rays, "
                                       f"{S_C}+{S_I} samples/ray, reference-parity renderer; per-ray rgb/depth/acc/sum(beta^2) maps + per-view scores",
                           "views_per_step": V, "rays_per_view": n_pix,
                           "parallelism": f"rays of every image sharded over {world} GPU(s): all_gather of the maps "
                                          f"({V * n_pix * 24 / 1e6:.0f} MB), all_reduce of the per-view partial sums",
                           "l2": "per-step working set exceeds L2; 256 MB L2 flush between timed steps"},
                "e2e": {"value": V / (ms_e2e / 1e3), "unit": "views/s", "h2d_bytes_per_step": V * 48,
                        "d2h_bytes_per_step": V * n_pix * 24 + V * 4,
                        "api": "Engine.get_rays + Engine.render_rays (== render()) + dist.render_views_ray_sharded; maps and scores to pinned host memory"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                             "traffic": None,      # (the committed DRAM capture is of the score-only launch: 4 + 4 B per evaluation; render() writes 20 B)
                             "peak_source": f"{src} bf16_tflops_sustained", "kernel": "mlp_tc_kernel",
                             "flop_per_launch": flops_per_launch, "avg_launch_ms": avg_launch_ms,
                             "mlp_share_of_step": mlp_ms / (ms_dev * args.steps) if ms_dev > 0 else None},
                "check": check, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_planner(args, rank, local_rank, world):
    """One planning step of the Object planner in the reference's live protocol (SURVEY.md section 8, rows a11-a13 + f3):
    100 locations x 15 directions = 1500 stage-1 views x 500 random pixels through the TSDF state volume, the 3 best
    directions per location = 300 stage-2 views x 1000 random pixels of the 400x400 image through the NeRF
    (Controller.get_uncertainty), depth clip / per-location maximum / normalisation, the surrogate fit (300 Adam epochs
    on the 100 gains) and 1000 surrogate queries.  A step is host-driven by nature (the top-3 selection sits between the
    stages), so the timed region is wall clock around the whole step with a device synchronise on both sides: `value`
    and `e2e` are the same measurement.  N > 1 runs independent replicas (a planning step does not shard)."""
    import evpp_b200
    from evpp_b200 import synthetic as S
    torch.cuda.set_device(local_rank)
    H = W = 400
    focal = 300.0
    c, f = S.init_nerf_state_dict(0), S.init_nerf_state_dict(1)
    ctl = evpp_b200.Controller(H=H, W=W, focal=focal, near=0.5, far=6.0, precision=args.precision, sample_num=1000, device=local_rank,
                               pixel_sampler=args.pixel_sampler)
    ctl.copy_model(c, f)
    ts = evpp_b200.TSDF(engine=ctl.engine, pixel_sampler=args.pixel_sampler)
    state, nray = S.state_volume(ts.index_bnd, seed=0)
    ts.set_state(state, nray)
    locations = np.asarray(S.hemisphere_views(100, seed=rank))[:, 0:3]
    center = [0.0, 1.0, 0.0]
    mlp = evpp_b200.MLP(ctl.engine)
    t_stage = {"stage1_tsdf": 0.0, "stage2_nerf": 0.0, "fit": 0.0, "queries": 0.0}

    def timed(key, fn):
        def wrapper(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize()
            t_stage[key] += time.perf_counter() - t0
            return r
        return wrapper

    class TsdfTimed:
        get_uncertainty_tsdf = staticmethod(timed("stage1_tsdf", ts.get_uncertainty_tsdf))

    def get_uncertainty(locs, us, vs):
        return ctl.get_uncertainty(evpp_b200.get_poses(locs, us, vs)).cpu().numpy().tolist(), [], []

    get_uncertainty = timed("stage2_nerf", get_uncertainty)
    test_pts = torch.tensor(np.random.RandomState(1).uniform(-3, 3, size=(1000, 3))).to(torch.float32)

    def step():
        data = evpp_b200.sample_direction(locations, TsdfTimed, get_uncertainty, center, scene="object")
        timed("fit", evpp_b200.train_only)(mlp, data[0:100], N=100)
        g = timed("queries", lambda: mlp(test_pts).detach().cpu())()
        return data, g

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", local_rank))
    for _ in range(args.warmup):
        step()
    for k in t_stage:
        t_stage[k] = 0.0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctl.engine.counters()[0]
    ctl.engine.set_profiling(True)
    wall = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        wall += time.perf_counter() - t0
    mlp_ms, mlp_launches, mlp_evals, mlp_flops = ctl.engine.mlp_profile()
    ctl.engine.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctl.engine.counters()[0] - l0
    if rank != 0:
        return
    sec = wall / args.steps
    views = 1500 + 300
    tensor = args.precision != "fp32"
    d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = d.get("bf16_tflops_sustained", 1399.2)
    achieved = mlp_flops / (mlp_ms / 1e3) / 1e12 if mlp_ms > 0 else 0.0        # FLOPs the launches executed (score-only passes)
    cpu = None if args.no_cpu_baseline else planner_cpu_baseline(c, f, locations, center, state, nray, ts)
    line = {"metric": "candidate views scored/sec", "value": views * world / sec, "unit": "views/s",
            "planning_steps_per_s": world / sec, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
            "higher_is_better": True, "scaling": "weak (independent replicas: a planning step does not shard)", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate (stage 2); u8 / f32 (stage 1); f32 (surrogate)" if tensor else "f32", "data": "synthetic",
            "config": {"workload": "one planning step of the Object planner, live protocol: 100 locations x 15 directions = 1500 stage-1 "
                                   "views x 500 px (TSDF state volume 100x50x110), top 3 per location = 300 stage-2 views x 1000 px of "
                                   "400x400 (64+128 samples/ray), selection, surrogate fit (300 Adam epochs, 100 points), 1000 queries",
                       "stage_ms": {k: 1e3 * v / args.steps for k, v in t_stage.items()},
                       "host_logic_ms": 1e3 * (sec - sum(t_stage.values()) / args.steps),
                       "pixel_sampler": args.pixel_sampler + (" (pixel subsets drawn on the device, evpp_sample_pixels)" if args.pixel_sampler == "device" else
                                                             " (the reference's np.random.choice(160000, R, replace=False) per pose on the host: ~2 ms per view)"),
                       "l2": "256 MB L2 flush between timed steps"},
            "e2e": {"value": views * world / sec, "unit": "views/s", "h2d_bytes_per_step": 1500 * (48 + 2000) + 300 * (48 + 4000) + 100 * 16,
                    "d2h_bytes_per_step": 1500 * 8 + 300 * 4 + 1000 * 4,
                    "api": "evpp_b200.sample_direction / train_only / MLP.__call__ over evpp_tsdf_score_views_host, evpp_score_views, evpp_surrogate_fit, evpp_surrogate_query"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "mlp_tc_kernel inside stage 2 (300 views x 1000 rays = 2 bulk launches of 19.2 M and 57.6 M evaluations + the exact "
                                   "re-score of the views within 2e-3 of the leader); short launches run near the maximum clock, hence close to the sustained peak",
                         "mlp_share_of_step": mlp_ms / 1e3 / (sec * args.steps) if sec > 0 else None},
            "clocks": clocks}
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


def planner_cpu_baseline(c, f, locations, center, state, nray, ts):
    """The same planning step with the oracle ports on the host (all threads), stages timed on bounded samples and scaled:
    stage 1 on 30 of the 1500 views, stage 2 on 2 of the 300 views, the surrogate fit and the queries in full."""
    import evpp_b200
    from oracle import evpp_oracle as O, tsdf_oracle as T, surrogate_oracle as SO
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    K = O.make_K(400, 400, 300.0)
    views, _ = evpp_b200.direction_fan(locations, center, "object")
    rs = np.random.RandomState(0)
    with torch.no_grad():
        v1 = views[:30]
        pix = np.stack([rs.choice(400 * 400, size=[500], replace=False) for _ in range(len(v1))])
        t0 = time.time()
        ro, rd = O.build_view_rays(O.views_to_poses(v1), 400, 400, K, pix)
        T.render_rays(ro, rd, torch.from_numpy(state), torch.from_numpy(nray), ts.vol_origin)
        t1 = (time.time() - t0) / len(v1)
        v2 = views[:2]
        pix = np.stack([rs.choice(400 * 400, size=[1000], replace=False) for _ in range(len(v2))])
        O.score_views(O.views_to_poses(v2[:1]), 400, 400, K, 0.5, 6.0, c, f, pix_idx=pix[:1])
        t0 = time.time()
        O.score_views(O.views_to_poses(v2), 400, 400, K, 0.5, 6.0, c, f, pix_idx=pix)
        t2 = (time.time() - t0) / len(v2)
    data = SO.synthetic_gain_samples(0)
    t0 = time.time()
    m, _ = SO.train_only(SO.MLP(), data, N=100)
    t3 = time.time() - t0
    t0 = time.time()
    with torch.no_grad():
        m(torch.rand(1000, 3))
    t4 = time.time() - t0
    sec = 1500 * t1 + 300 * t2 + t3 + t4
    return {"value": 1800 / sec, "unit": "views/s", "cores": threads, "kind": "port",
            "sample": f"stage 1: 30 of 1500 views ({1e3 * t1:.1f} ms/view), stage 2: 2 of 300 views ({1e3 * t2:.0f} ms/view), surrogate fit in full "
                      f"({t3:.2f} s), 1000 queries ({1e3 * t4:.1f} ms); scaled step {sec:.1f} s"}


def workload_config(world, engine_desc):
    if WORKLOAD == "room":
        return {"workload": f"BASELINE configs[2] Room scene sweep (north star): {V_TOTAL} candidate poses x {IMG_W}x{IMG_H} rays, split over "
                            f"{world} GPU(s), {S_C}+{S_I} samples/ray ({EVALS_PER_RAY} MLP evals/ray), reference-parity scorer "
                            "(freq. encoding + 8x256 NeRF + NeurAR uncertainty head, random-init seed 0); a step = bulk pass + exact "
                            "re-score of the near-tied leaders + NBV selection",
                "views": V_TOTAL, "views_per_gpu": (V_TOTAL + world - 1) // world, "rays_per_view": IMG_H * IMG_W, "near": NEAR, "far": FAR,
                "parallelism": f"{V_TOTAL} views sharded over {world} GPU(s) in contiguous blocks (sized in proportion to each rank's measured "
                               "speed after two calibration sweeps), all_gather of the scores, contenders "
                               "dealt out round robin for the exact pass, second all_gather, replicated selection",
                "engine": engine_desc,
                "l2": "per-step working set (~0.7 GB of streamed ray/sample intermediates per 131072-ray chunk, >300 chunks) exceeds L2; "
                      "256 MB L2 flush between timed steps"}
    return {"workload": f"BASELINE configs[1] Object scene: {V_TOTAL} candidate views x {IMG_W}x{IMG_H} rays per GPU, "
                        f"{S_C}+{S_I} samples/ray ({EVALS_PER_RAY} MLP evals/ray), reference-parity scorer "
                        "(freq. encoding + 8x256 NeRF + NeurAR uncertainty head, random-init seed 0)",
            "views_per_gpu": V_TOTAL, "rays_per_view": IMG_H * IMG_W, "near": NEAR, "far": FAR,
            "parallelism": f"views sharded over {world} GPU(s), all_gather of scores", "engine": engine_desc,
            "l2": "per-step working set (~14 GB of streamed ray/sample intermediates) exceeds L2; 256 MB L2 flush between timed steps"}


def measured_traffic_per_eval():
    """DRAM bytes per MLP evaluation from the committed ncu --set full capture of one benchmark-sized launch
    (profiles/r2_mlp_tc_dram.json, written by tools/ncu_metrics.py from the .ncu-rep); None when absent."""
    path = os.path.join(ROOT, "profiles", "r2_mlp_tc_dram.json")
    try:
        d = json.load(open(path))
        return float(d["dram_bytes"]) / float(d["evals"]), d
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="evpp_b200")
    ap.add_argument("--precision", default="fp16", choices=["fp32", "fp16", "bf16", "fp16x2", "fp16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-balance", action="store_true", help="equal view blocks per rank (no speed-proportional split)")
    ap.add_argument("--no-rescore", action="store_true", help="bulk pass only (no exact re-score of the leaders, no NBV identity)")
    ap.add_argument("--no-fold", action="store_true", help="score-only passes in the reference's layer order (feature_linear not folded into the view layer)")
    ap.add_argument("--full-outputs", action="store_true", help="scoring calls evaluate the full network like render() (round-1 behaviour)")
    ap.add_argument("--pixel-sampler", default="device", choices=["device", "numpy"],
                    help="--mode planner only: where the random pixel subsets of the live protocol are drawn")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' runs the same reference torch path on cuda:0 (informational row; "
                         "the reference arm the driver compares against is the CPU one)")
    ap.add_argument("--workload", default="room", choices=["object", "room"],
                    help="room (default): the 4096-view Room sweep of BASELINE configs[2], strong-scaled over the ranks, with NBV "
                         "selection inside the step; object: BASELINE configs[1], 512 hemisphere views x 80x60 rays per GPU (weak)")
    ap.add_argument("--views", type=int, default=0, help="override the number of views (room: whole sweep; object: per GPU)")
    ap.add_argument("--mode", default="parity", choices=["parity", "ngp", "tensorf", "planner", "fullres"],
                    help="parity: the reference's scorer (headline); ngp: occupancy grid + hash grid "
                         "scorer on the Room sweep shape (BASELINE configs[2]; no reference counterpart); fullres: BASELINE configs[4], "
                         "640x480 renders with the rays of every image sharded over the ranks (strong scaling)")
    args = ap.parse_args()
    global IMG_H, IMG_W, FOCAL, WORKLOAD, V_TOTAL
    if args.workload == "object":
        IMG_H, IMG_W, FOCAL, WORKLOAD, V_TOTAL = 60, 80, 60.0, "object", 512
    if args.views > 0:
        V_TOTAL = args.views
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device visible -- the evpp_b200 engine has no CPU fallback "
                 "(--impl reference times the reference algorithm on the host cores)")
    if args.mode in ("ngp", "tensorf"):
        run_ngp(args, rank, local_rank, world)
        return
    if args.mode == "planner":
        run_planner(args, rank, local_rank, world)
        return
    if args.mode == "fullres":
        run_fullres(args, rank, local_rank, world)
        return

    import torch.distributed as dist
    import evpp_b200
    from evpp_b200 import dist as edist
    from evpp_b200.run_nerf import _intrinsics
    from __graft_entry__ import build
    if local_rank == 0:
        build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    dev = torch.device("cuda", local_rank)
    c, f, views, poses_host = make_workload(rank, world)
    K = np.array([[FOCAL, 0, 0.5 * IMG_W], [0, FOCAL, 0.5 * IMG_H], [0, 0, 1]])
    fx, fy, cx, cy = _intrinsics(K)
    eng = evpp_b200.Engine(device=local_rank, n_samples=S_C, n_importance=S_I, near=NEAR, far=FAR,
                           white_bkgd=True, precision=args.precision)
    eng.load_weights(0, c)
    eng.load_weights(1, f)
    eng.set_score_mode(1 if args.full_outputs else 2 if args.no_fold else 0)
    room = WORKLOAD == "room"
    V = V_TOTAL                                   # views this rank holds poses for (room: the whole sweep)
    V_job = V_TOTAL if room else V_TOTAL * world  # views scored per step by the whole job
    poses_dev = poses_host.to(dev)
    poses_pinned = poses_host.pin_memory()
    out_pinned = torch.empty(V_job + 1, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    margin = 0.0 if args.no_rescore else MARGIN
    state = {"winner": None, "band": 0, "scores": None, "ranges": None, "calibrating": world > 1 and not args.no_balance}

    def sweep(poses, host_out):
        if room:
            timing = {} if state["calibrating"] else None
            s, win, band = edist.sweep_select(eng, poses, IMG_H, IMG_W, fx, fy, cx, cy, rank, world, margin=margin, host_out=host_out,
                                              ranges=state["ranges"], timing=timing)
            if timing:
                # load balance: the next sweep's blocks are proportional to the speeds the ranks
                # just measured on their own bulk pass (the GPUs of a box run 1-3 % apart under the power cap)
                sp = edist.exchange_speeds(timing["bulk_views"], timing["bulk_ms"], rank, world, dev)
                state["ranges"] = edist.balanced_ranges(V_TOTAL, sp)
        else:       # weak scaling: every rank scores its own 512 views, then the same gather / selection tail
            pd = poses.to(dev, non_blocking=True) if poses.device.type == "cpu" else poses
            s = eng.score_views(pd, IMG_H, IMG_W, fx, fy, cx, cy)
            if world > 1:
                s = edist.gather_scores(s, V_job, rank, world)
            _, _, w = eng.select_nbv(s, None, 1)
            if host_out is not None:
                host_out[:V_job].copy_(s, non_blocking=True)
                host_out[V_job:].copy_(w.to(torch.float32), non_blocking=True)
                torch.cuda.synchronize()
            win, band = int(w.item()), 0
        state.update(winner=win, band=band, scores=s)
        return s

    def step_device():
        return sweep(poses_dev, None)

    def step_e2e():
        return sweep(poses_pinned, out_pinned)

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if profile:
            eng.set_profiling(True)
        l0, e0 = eng.counters()
        total_ms, wall = 0.0, 0.0
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            wall += time.perf_counter() - t0
            total_ms += ev0.elapsed_time(ev1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l1, e1 = eng.counters()
        t = torch.tensor([total_ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item() / steps, t[1].item() / steps, (l1 - l0), (e1 - e0)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if state["calibrating"]:
        for _ in range(2):          # two untimed calibration sweeps fix the block sizes; the timed steps do not re-balance
            step_device()
        state["calibrating"] = False
    ms_dev, _, launches, evals = timed(step_device, args.steps, args.warmup, profile=True)
    mlp_ms, mlp_launches, mlp_evals, mlp_flops = eng.mlp_profile()
    eng.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    _, ms_e2e_wall, _, _ = timed(step_e2e, max(2, min(args.steps, 3)), min(args.warmup, 3))
    # transparency: the same step with the score-only passes in the reference's ten-layer order (feature_linear NOT folded
    # into the view layer), one untimed + one timed sweep
    ms_unfolded = None
    if not args.full_outputs and not args.no_fold and args.precision != "fp32":
        eng.set_score_mode(2)
        ms_unfolded, _, _, _ = timed(step_device, 1, 1)
        eng.set_score_mode(0)

    if rank == 0:
        sustained, burst, src = peaks()
        value = V_job / (ms_dev / 1e3)
        e2e_val = V_job / (ms_e2e_wall / 1e3)
        avg_launch_ms = mlp_ms / max(1, mlp_launches)
        achieved = mlp_flops / (mlp_ms / 1e3) / 1e12 if mlp_ms > 0 else 0.0            # FLOPs the launches really executed
        nominal = FLOP_PER_EVAL * mlp_evals / (mlp_ms / 1e3) / 1e12 if mlp_ms > 0 else 0.0
        tensor = args.precision != "fp32"
        bpe, cap = measured_traffic_per_eval()
        # NBV check: the winner of the timed sweep against the stored fp32-oracle winner of this exact workload
        check = {"winner": state["winner"], "rescored_views": state["band"], "expected": None, "ok": None}
        if room and os.path.exists(NBV_FIXTURE):
            fx_ = json.load(open(NBV_FIXTURE))
            if fx_.get("views") == V_TOTAL and fx_.get("H") == IMG_H and fx_.get("W") == IMG_W:
                check["expected"] = fx_["winner"]
                check["ok"] = bool(state["winner"] == fx_["winner"]) if not args.no_rescore else None
                check["source"] = fx_.get("source")
        if check["ok"] is False:
            sys.exit(f"bench.py: NBV mismatch: the sweep selected view {state['winner']}, the fp32 oracle selects {check['expected']}")
        line = {"metric": "candidate views scored/sec", "value": value, "unit": "views/s",
                "rays_per_s": value * IMG_H * IMG_W, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong" if room else "weak", "vs_baseline": None,
                "dtype": {"fp16": "f16 operands, f32 accumulate (tcgen05); leaders re-scored with split-f16 operands (3 MMAs per product, f32-class)",
                          "bf16": "bf16 operands, f32 accumulate (tcgen05); leaders re-scored with split-f16 operands",
                          "fp16x2": "split-f16 weights, f16 activations, f32 accumulate (tcgen05)", "fp16x3": "split-f16 operands, f32 accumulate (tcgen05)",
                          "fp32": "f32"}[args.precision],
                "data": "synthetic", "config": workload_config(world, f"evpp_b200 CUDA engine, precision={args.precision}, "
                                                                      f"{'full' if args.full_outputs else 'score-only'} network outputs{'' if args.full_outputs or args.no_fold else ', feature_linear folded into the view layer'}, rescore margin {margin:g}"),
                "e2e": {"value": e2e_val, "unit": "views/s", "h2d_bytes_per_step": int(V * 12 * 4),
                        "d2h_bytes_per_step": int(V_job * 4 + 4 + (V_job * 4 if room else 0)), "ms_per_step": ms_e2e_wall,
                        "api": "evpp_b200.dist.sweep_select (pinned host poses in every step; scores + winner id land in pinned host memory; the "
                               "contender logic reads the gathered scores on the host)" if room else
                               "Engine.score_views + select_nbv (pinned host poses in, host scores + winner out)"},
                "gpu_launches": int(launches),
                "nbv_check": check,
                "variants": {"score_only_reference_layer_order": {"value": V_job / (ms_unfolded / 1e3), "unit": "views/s", "ms_per_step": ms_unfolded,
                                                                  "note": "1 warm-up + 1 timed sweep with --no-fold: ten GEMM layers per fine evaluation as in the reference; "
                                                                          "the headline folds feature_linear into views_linears.0 (no activation between them: one linear map)"}}
                if ms_unfolded else None,
                "view_blocks": [e_ - b_ for b_, e_ in state["ranges"]] if state["ranges"] else None,
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                             "frac": achieved / sustained,
                             "traffic": (bpe * mlp_evals / max(1, mlp_launches)) if (tensor and bpe) else None,
                             "traffic_unit": "bytes/launch: DRAM bytes per evaluation of the committed ncu --set full capture of one 131072-ray fine "
                                             "launch (profiles/r2_mlp_tc_dram.json) x the evaluations of an average launch of this run" if bpe else None,
                             "peak_source": f"{src} bf16_tflops_sustained (cuBLAS bf16 under the power cap; the kernel is timed inside a long step)",
                             "kernel": "mlp_tc_kernel (fused encode + NeRF/NeurAR MLP, tcgen05 TS-mode, activations in tensor memory; all launches of the "
                                       "step: bulk fp16 and exact fp16x3, score-only variants)" if tensor else "mlp_simt_kernel (fp32 FFMA; tensor peak shown for scale only)",
                             "achieved_definition": "FLOPs the launches executed (score-only coarse evaluation 982,528, fine 1,054,720 with the folded feature layer; an fp16x3 launch "
                                                    "counts its evaluations once, not its three MMA passes) / CUDA-event time of the launches",
                             "achieved_at_1187072_flop_per_eval": nominal,
                             "flop_per_launch": mlp_flops / max(1, mlp_launches), "avg_launch_ms": avg_launch_ms,
                             "mlp_share_of_step": mlp_ms / (ms_dev * args.steps) if ms_dev > 0 else None},
                "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            v, cores, kind, sample = cpu_reference_rate(c, f, views)
            line["cpu_baseline"] = {"value": v, "unit": "views/s", "cores": cores, "kind": kind, "sample": sample}
            # informational: the reference's own torch path with its tensors on this GPU -- what the
            # reference delivers in deployment -- so that the speed-up decomposes into "GPU vs CPU" and "fused sm_100a
            # vs PyTorch eager" (BASELINE.md section 3).  Bounded sample, separate process.
            try:
                res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--ref-device", "cuda", "--steps", "3",
                                      "--warmup", "1", "--workload", WORKLOAD], capture_output=True, text=True, timeout=300)
                rl = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
                line["reference_torch_on_this_gpu"] = {"value": rl["value"], "unit": "views/s", "sample": rl["cpu_baseline"]["sample"],
                                                       "note": "informational; the reference arm the driver compares against is the host-CPU one"}
            except Exception as e:
                line["reference_torch_on_this_gpu"] = {"value": None, "note": f"not measured: {e}"}
            # informational: the north_star's NGP-mode scorer (occupancy bitfield + hash grid; no reference counterpart, parity
            # unpinned) and its TensoRF variant on the same sweep shape, so that the driver-run line carries them too
            line["other_modes"] = {}
            for mode in ("ngp", "tensorf"):
                try:
                    res = subprocess.run([sys.executable, os.path.abspath(__file__), "--mode", mode, "--steps", "3", "--warmup", "3"],
                                         capture_output=True, text=True, timeout=300)
                    ml = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
                    line["other_modes"][mode] = {"value": ml["value"], "unit": "views/s", "samples_per_s": ml.get("samples_per_s"),
                                                 "e2e": ml["e2e"]["value"], "workload": ml["config"]["workload"],
                                                 "field_kernel_gather_GBs": ml["roofline"].get("achieved_field_kernel_only"),
                                                 "roofline_frac": ml["roofline"]["frac"], "roofline_peak_GBs": ml["roofline"]["peak"],
                                                 "cmd": f"python bench.py --mode {mode} --steps 3 --warmup 3"}
                except Exception as e:
                    line["other_modes"][mode] = {"value": None, "note": f"not measured: {e}"}
            try:    # one planning step of the Object planner (rows f1 + a11 + a12 + f3 behind the reference's call shapes)
                res = subprocess.run([sys.executable, os.path.abspath(__file__), "--mode", "planner", "--steps", "3", "--warmup", "3"],
                                     capture_output=True, text=True, timeout=300)
                ml = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
                line["other_modes"]["planner"] = {"ms_per_planning_step": ml["ms_per_step"], "value": ml["value"], "unit": "views/s",
                                                  "stage_ms": ml["config"].get("stage_ms"), "workload": ml["config"]["workload"],
                                                  "cmd": "python bench.py --mode planner --steps 3 --warmup 3"}
            except Exception as e:
                line["other_modes"]["planner"] = {"value": None, "note": f"not measured: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
