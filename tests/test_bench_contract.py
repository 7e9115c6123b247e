"""bench.py contract checks that need no GPU: the reference arm (the oracle port timed on the host cores) prints ONE
JSON line with the keys the driver reads, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          cwd=ROOT, env=e)


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "2", "--critic-iters", "1", "--vocab", "50")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "images/s"
    assert d["metric"].startswith("train images/sec")
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["vs_baseline"] is None                       # BASELINE.json publishes no number for this metric
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_non_zero_ranks_exit_quietly():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "2", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = _run("--steps", "1", "--warmup", "0")
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_workload_label_names_the_baseline_config_of_the_shape():
    """config.workload must name BASELINE.json's configs[] entry of the SHAPE that ran (round-1 verdict: every shape was
    labelled configs[1], and the reference arm claimed batch 256 while running 32)."""
    import argparse
    b = _bench_module()
    assert b.config_index(256, 3, 2000, 5) == 1 and b.config_index(32, 3, 2000, 5) == 0
    assert b.config_index(256, 30, 5000, 5) == 2 and b.config_index(64, 3, 2000, 5) is None
    a = argparse.Namespace(batch=256, timesteps=3, vocab=2000, critic_iters=5, cpu_batch=32)
    assert b.config_name(a).startswith("BASELINE configs[1]") and "batch 256/GPU" in b.config_name(a)
    assert b.config_name(a, 32).startswith("BASELINE configs[0]") and "batch 32/GPU" in b.config_name(a, 32)
    a3 = argparse.Namespace(batch=256, timesteps=30, vocab=5000, critic_iters=5, cpu_batch=32)
    assert b.config_name(a3).startswith("BASELINE configs[2]") and "10 triple" in b.config_name(a3)
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "2", "--critic-iters", "1", "--vocab", "50")
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert "batch 2/GPU" in d["config"]["workload"] and d["config"]["cpu_sample_batch"] == 2


def test_numa_binding_never_fails_the_run():
    """Host placement is best effort: without NVML / sysfs NUMA information it reports why and changes nothing."""
    b = _bench_module()
    before = os.sched_getaffinity(0)
    info = b.bind_to_gpu_numa_node(0)
    assert set(info) >= {"numa_node", "cpus"}
    if info["cpus"] is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
