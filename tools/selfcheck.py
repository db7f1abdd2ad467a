"""GPU-box self-check that stands in for compute-sanitizer (closed on this pool: "runs under it have left GPUs needing
a reset"): memcheck-like and racecheck-like evidence from the engine's own instrumentation.

  * out-of-bounds writes: every engine-owned device buffer has 256-byte guard bands (evpp_check_guards); caller-owned
    outputs are over-allocated here with a sentinel tail that must stay intact;
  * races / missing synchronisation: every kernel path is run REPEATS times on the same inputs and must return the
    same bits every time (a race between the producer / MMA / epilogue roles, a stale weight stage or an early
    tensor-memory overwrite changes some output of some run), for every MLP variant (precision x score-only x points /
    rays), ragged sizes, and every cluster launch mode (EVPP_TC_CLUSTER = 1, 2, 4, 42; one subprocess each);
  * uninitialised reads: the chunk workspace is poisoned with NaN bit patterns between runs (evpp buffers are reused
    chunk after chunk), so a read of a stale or never-written element turns up as a NaN / changed score.

    python tools/selfcheck.py            # all cluster modes, prints a summary, exit code 1 on any finding
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REPEATS = int(os.environ.get("REPEATS", 12))


def one_mode():
    import numpy as np
    import torch
    from evpp_b200 import Engine, NGPScorer, TSDF, MLP, train_only, ngp, synthetic as S   # noqa: F401
    from evpp_b200.engine import _ptr, _stream_ptr
    torch.manual_seed(0)
    c, f = S.init_nerf_state_dict(), S.init_nerf_state_dict()
    findings, runs = [], 0
    H, W, focal = 24, 32, 24.0
    views = S.room_views(12, seed=3)
    poses = torch.Tensor(np.asarray(S.views_to_poses(views)))[:, :3, :4].contiguous().cuda()
    M = 128 * 301 + 77
    pts = ((torch.rand(M, 3) - 0.5) * 8).cuda()
    dirs = torch.nn.functional.normalize(torch.randn(M, 3), dim=-1).cuda()
    SENT = 12345.0
    for prec in ("fp16", "bf16", "fp16x2", "fp16x3", "fp32"):
        eng = Engine(device=0, near=0.5, far=6.0, precision=prec, max_chunk_rays=4000)      # several ragged chunks
        eng.load_weights(0, c)
        eng.load_weights(1, f)
        lib = eng.lib
        for full in (False, True):
            eng.set_score_mode(full)
            ref = None
            for r in range(REPEATS):
                eng.poison_workspace()
                out = torch.full((poses.shape[0] + 64,), SENT, device="cuda")
                rc = lib.evpp_score_views_ex(eng._h, _ptr(poses), poses.shape[0], H, W, focal, focal, 0.5 * W, 0.5 * H, C.c_void_p(0), 0, -1,
                                             _ptr(out), _stream_ptr(eng.device))
                assert rc == 0, lib.evpp_last_error()
                s = out.cpu()
                runs += 1
                if not torch.isfinite(s[:poses.shape[0]]).all():
                    findings.append(f"{prec} full={full}: non-finite score (stale workspace read?)")
                if not (s[poses.shape[0]:] == SENT).all():
                    findings.append(f"{prec} full={full}: caller's score buffer overrun")
                if ref is None:
                    ref = s
                elif not torch.equal(ref, s):
                    findings.append(f"{prec} full={full}: run {r} differs from run 0 (max {float((ref - s).abs().max()):.3e})")
        ref = None
        for r in range(REPEATS):
            out = torch.full((M + 64, 5), SENT, device="cuda")
            rc = lib.evpp_mlp_forward(eng._h, 1, -1, _ptr(pts), _ptr(dirs), M, _ptr(out), _stream_ptr(eng.device))
            assert rc == 0, lib.evpp_last_error()
            y = out.cpu()
            runs += 1
            if not (y[M:] == SENT).all():
                findings.append(f"{prec} mlp_forward: caller's output overrun")
            if ref is None:
                ref = y
            elif not torch.equal(ref, y):
                findings.append(f"{prec} mlp_forward: run {r} differs from run 0")
        o = eng.render_rays(*eng.get_rays(poses[:2], H, W, focal, focal, 0.5 * W, 0.5 * H)[:2], 0.5, 6.0,
                            want=("rgb", "disp", "acc", "depth", "unc", "raw", "rgb0", "disp0", "acc0", "depth0", "unc0", "z_std"))
        if not all(torch.isfinite(v).all() for k, v in o.items() if k not in ("disp", "disp0")):
            findings.append(f"{prec} render_rays: non-finite output")
        bad, n = eng.check_guards()
        if bad:
            findings.append(f"{prec}: {bad} guard bytes overwritten around the engine's {n} buffers")
        print(f"  {prec:7s}: guards of {n} buffers intact = {bad == 0}", flush=True)
        eng.close()
    # the other kernel families on one handle: TSDF, NGP march / field / composite, surrogate fit
    eng = Engine(device=0, near=0.5, far=6.0, precision="fp16")
    eng.load_weights(0, c)
    eng.load_weights(1, f)
    sc = NGPScorer(engine=eng, H=H, W=W, focal=focal)
    sc.set_occupancy(S.room_occupancy_bitfield(128))
    sc.load_params(ngp.init_params(0))
    ref = None
    for r in range(REPEATS):
        eng.poison_workspace()
        s = sc.score_views(poses).cpu()
        runs += 1
        if ref is None:
            ref = s
        elif not torch.equal(ref, s):
            findings.append(f"ngp score_views: run {r} differs from run 0")
    ts = TSDF(engine=eng)
    state, nray = S.state_volume(ts.index_bnd, seed=0)
    ts.set_state(state, nray)
    u0 = None
    for r in range(REPEATS):
        eng.poison_workspace()
        np.random.seed(1)
        u, d = ts.get_uncertainty_tsdf([v[0:3] for v in views], [v[3] for v in views], [v[4] for v in views])
        runs += 1
        if u0 is None:
            u0 = (u, d)
        elif u0 != (u, d):
            findings.append(f"tsdf: run {r} differs from run 0")
    from oracle import surrogate_oracle as SO
    data = SO.synthetic_gain_samples(0)
    p0 = None
    for r in range(3):
        torch.manual_seed(5)
        m = train_only(MLP(eng), data, N=100)
        p = m.params.cpu()
        runs += 1
        if p0 is None:
            p0 = p
        elif not torch.equal(p0, p):
            findings.append(f"surrogate fit: run {r} differs from run 0")
    bad, n = eng.check_guards()
    if bad:
        findings.append(f"ngp/tsdf/surrogate: {bad} guard bytes overwritten ({n} buffers)")
    print(f"  other kernels: guards of {n} buffers intact = {bad == 0}", flush=True)
    eng.close()
    print(f"MODE EVPP_TC_CLUSTER={os.environ.get('EVPP_TC_CLUSTER', '2 (default)')}: {runs} runs, {len(findings)} findings")
    for x in findings:
        print("  FINDING:", x)
    return len(findings)


if __name__ == "__main__":
    if "--one" in sys.argv:
        sys.exit(1 if one_mode() else 0)
    total = 0
    for mode in ("1", "2", "4", "42"):
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=dict(os.environ, EVPP_TC_CLUSTER=mode),
                             capture_output=True, text=True, timeout=1500)
        print(res.stdout.rstrip())
        if res.returncode != 0:
            total += 1
            print(res.stderr[-3000:])
    print("SELF-CHECK", "CLEAN" if total == 0 else f"FAILED in {total} mode(s)")
    sys.exit(1 if total else 0)
