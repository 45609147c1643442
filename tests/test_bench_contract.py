"""CPU checks of bench.py's contract: the reference arm's JSON line (the one leg that needs no GPU), that both arms
describe the workload with the same `config` object, and that the product arm refuses to run without a CUDA device
instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline")


def run_bench(*argv, env=None, timeout=600):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *argv], capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


@pytest.fixture(scope="module")
def reference_line(oracle):
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "the arm prints exactly ONE JSON line"
    return json.loads(lines[0])


def test_reference_arm_line(reference_line):
    d = reference_line
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference"
    assert d["metric"] == "features_depth_associated_per_s" and d["unit"] == "features/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    # the CPU arm copies nothing and its e2e is the line's own value
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "keyframes" in cb["sample"]


def test_both_arms_name_the_same_workload(reference_line):
    """run_gpu and run_reference both build `config` with workload_config(batch, args.keyframes): same keys, and the values
    follow BASELINE.json configs[2] (the largest single-GPU configuration)."""
    sys.path.insert(0, ROOT)
    import bench

    src = open(BENCH).read()
    assert src.count('"config": workload_config(batch, ') == 2
    cfg = reference_line["config"]
    assert cfg["workload"] == bench.workload_name(bench.DEFAULT_KEYFRAMES)
    assert cfg["keyframes_per_gpu"] == bench.DEFAULT_KEYFRAMES
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "configs" in base and len(base["configs"]) >= 3
    assert cfg["scans_per_keyframe"] == 20 and cfg["points_per_scan"] == 131072 and cfg["resolution"] == [1440, 1080]
    assert cfg["features_per_keyframe"] == 200 and abs(cfg["td_b_l_s"] + 0.48) < 1e-12
    assert "larger than L2" in cfg["l2"]


def test_reference_arm_other_ranks_exit_quietly():
    p = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the product arm would run")
    p = run_bench("--steps", "1", "--warmup", "0", timeout=300)
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
    assert not any(ln.startswith("{") for ln in p.stdout.splitlines())


def test_gpus_flag_must_match_the_launcher():
    p = run_bench("--gpus", "2", "--steps", "1", "--warmup", "0", timeout=300)
    assert p.returncode != 0 and "WORLD_SIZE" in (p.stderr + p.stdout)
