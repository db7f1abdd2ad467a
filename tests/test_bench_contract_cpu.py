"""The reference arm of bench.py (CPU only, the one leg that may execute oracle/) prints one JSON line with the
contract's keys; the product arm must refuse to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "views/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
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
                assert line["gpu_launches"] == 0 and line["cpu_baseline"]["kind"] == "port"
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
