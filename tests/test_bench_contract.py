"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the required keys;
the GPU arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from tests.parity_utils import ROOT


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_contract_keys():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--n-azimuth", "96", "--ref-sample", "4000"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 1e3 and abs(d["value"] - d["e2e"]["value"]) < 1e-6 * d["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_gpu_arm_refuses_to_run_without_cuda():
    r = _run(["--steps", "1", "--warmup", "1", "--n-azimuth", "64"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_both_arms_describe_the_same_config():
    """The reference arm steps the same workload with the same points per step as our arm: identical `config` objects
    (same_config must be true)."""
    import bench
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--n-azimuth", "64"])
    assert r.returncode == 0, r.stderr[-2000:]
    ref = json.loads(r.stdout.strip().splitlines()[-1])
    cfg, octree, decoder, pool = bench.build_workload("cpu", 0, 1, 64)
    ours = bench.shared_config(cfg, 64, len(pool), len(pool), 1, [int(p.shape[0]) for p in octree.hier_features])
    assert ref["config"] == ours
    assert ref["config"]["points_per_step_per_gpu"] == len(pool) == ref["config"]["global_points_per_step"]
    assert str(len(pool)) in ref["cpu_baseline"]["sample"]
