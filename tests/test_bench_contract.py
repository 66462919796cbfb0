"""bench.py contract checks that do not need a GPU: the reference arm (`--impl reference`) prints exactly one JSON line on
stdout with the keys the driver reads, and the product arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    # C2 keeps the CPU suite short; the driver runs the default workload (C4, all ~5M observations) on the GPU box's host
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "c2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gpba_observations_per_sec" and d["unit"] == "obs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"]
    # the reference arm runs the workload itself, not a sample: the config object is the one the device arm prints
    sys.path.insert(0, ROOT)
    import bench
    P = bench.load_problem("c2")
    assert d["config"] == bench.workload_config("c2", P) and d["config"]["n_obs"] == P.n_obs
    assert set(d["g2o_batch_stats_s_per_step"]) == {"timeResiduals", "timeQuadraticForm", "timeSchurComplement", "timeLinearSolver", "timeUpdate"}
    assert d["host"]["nproc"] == cb["cores"] and d["lm_iters_per_step"] == 2


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu", "--workload", "c2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "usage:" not in r.stderr          # a real refusal, not an argument error
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]    # no number without a device
