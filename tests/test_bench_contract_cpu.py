"""The reference arm of bench.py (CPU only, the one leg that may execute oracle/) prints one JSON line with the
contract's keys; the product arm must refuse to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_reference_arm_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "views/s" and line["value"] > 0
    from oracle import ref_import
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_import.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["scaling"] == "strong" and "4096" in line["config"]["workload"]      # the north-star Room sweep is the default
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode != 0                      # fails loudly: no CUDA device, no CPU fallback
    assert "no CPU fallback" in (res.stderr + res.stdout)


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ (measured on the B200 boxes) carry every key of the contract, a roofline
    whose fraction is achieved / peak, an e2e block with the host<->device bytes and clocks without thermal or
    hardware slowdown."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*.jsonl")))
    assert files
    n = 0
    for f in files:
        for raw in open(f):
            line = json.loads(raw)
            n += 1
            for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                      "vs_baseline", "dtype", "data", "config", "e2e"):
                assert k in line, (f, k)
            assert line["value"] > 0 and "workload" in line["config"] and line["vs_baseline"] is None
            if line.get("impl") == "reference":
                assert line["gpu_launches"] == 0 and line["cpu_baseline"]["kind"] in ("port", "reference")
                continue
            assert line["gpu_launches"] > 0
            assert line["steps"] >= 1 and line["warmup"] >= 3
            r = line["roofline"]
            assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
            e = line["e2e"]
            assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
            bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
            assert not bad, (f, bad)
    assert n >= 8


def test_reference_arm_times_the_reference_itself_and_agrees_with_the_port():
    """bench.HostReference: the vendored / in-tree UNMODIFIED reference render() and the oracle port give the same bits
    on a bench sample (so the arm's kind is a label of provenance, not a different computation)."""
    import numpy as np
    import pytest
    import torch
    import bench
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference not importable here")
    old = (bench.IMG_H, bench.IMG_W, bench.FOCAL, bench.V_TOTAL)
    try:
        bench.IMG_H, bench.IMG_W, bench.FOCAL, bench.V_TOTAL = 12, 16, 12.0, 64
        c, f, views, _ = bench.make_workload(0, 1)
        ref = bench.HostReference(c, f, views)
        assert ref.kind == "reference"
        ids, pix, veq = bench.reference_sample(0)
        assert len(ids) == 32 and pix.shape == (32, 12 * 16 // 32) and abs(veq - 1.0) < 1e-9
        a = ref.score(ids, pix)
        ref.kind = "port"
        b = ref.score(ids, pix)
        assert torch.equal(a, b)
    finally:
        bench.IMG_H, bench.IMG_W, bench.FOCAL, bench.V_TOTAL = old


def test_room_sweep_nbv_fixture_is_consistent():
    """tests/golden/room_sweep_nbv.json (tools/make_room_sweep_fixture.py): the stored winner is the arg-max of the
    host reference's scores of the leaders, the exact tensor-core path agrees with them to 2e-5, and no view outside
    the oracle set is within reach of the leader."""
    import numpy as np
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "room_sweep_nbv.json")))
    orc, x3 = np.array(g["oracle_scores"]), np.array(g["fp16x3_scores_of_oracle_ids"])
    assert g["winner"] == g["oracle_ids"][int(np.argmax(orc))]
    assert np.abs(orc - x3).max() / orc.max() < 2e-5
    assert g["leaders_outside_oracle_set_are_out_of_reach"] and "reference" in g["source"]
    assert g["views"] == 4096 and (g["H"], g["W"]) == (96, 128)
