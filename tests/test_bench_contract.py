"""bench.py's reference arm runs on the CPU, so its side of the JSON contract can be checked without a GPU; the GPU arm
must refuse to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run_bench("--impl", "reference", "--steps", "4", "--warmup", "1", "--batch", "64")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "blocks/s" and d["higher_is_better"] is True
    assert d["steps"] == 4 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == ("reference" if ol.ref_path() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"] and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["return_codes_or"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "2", "--warmup", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_both_arms_describe_the_same_workload():
    """The driver compares the two arms' `config` dicts (`same_config`): both come from one function of the command line.
    Default = BASELINE configs[3], 262144 instances strong-scaled over the GPUs of the run."""
    import argparse
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
        for n in (1, 2, 4, 8):
            sys.argv = ["bench.py", "--gpus", str(n), "--steps", "20", "--warmup", "5"]
            a = bench.parse_args()
            cfg = bench.config_dict(a, n)
            assert cfg["total_batch"] == 262144 and cfg["batch_per_gpu"] == 262144 // n and cfg["scaling"] == "strong"
            assert "configs[3]" in cfg["workload"] and "model" not in cfg
            assert bench.workload(a, n) == (262144 // n, "strong")
        sys.argv = ["bench.py", "--batch", "4096", "--gpus", "2"]
        assert bench.workload(bench.parse_args(), 2) == (4096, "weak")
    finally:
        sys.argv = argv
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--total-batch", "128")
    d = json.loads(r.stdout.strip().splitlines()[-1])
    sys.argv = ["bench.py", "--steps", "2", "--warmup", "1", "--total-batch", "128"]
    try:
        assert d["config"] == bench.config_dict(bench.parse_args(), 1)
    finally:
        sys.argv = argv
