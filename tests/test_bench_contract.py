"""The reference arm of bench.py (`--impl reference`): runs on host cores only, so its JSON contract is checked here on
the CPU.  The product arm needs a GPU; its line is checked by the gpu-marked test at the end of this file."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def _run(extra_env=None, *flags):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *flags],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["unit"] == "evals/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["metric"].startswith("pose-beam") and base["metric"].startswith("pose-beam")  # BASELINE.json's headline metric
    assert d["vs_baseline"] is None  # BASELINE.md publishes no number for this metric
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d.get("gpu_launches", 0) == 0


def test_reference_arm_only_rank_zero_works_under_torchrun():
    # ranks other than 0 exit 0 without printing; rank 0 prints the line for the N it was launched with
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2") == []
    lines = _run({"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2


import pytest  # noqa: E402


@pytest.mark.gpu
def test_product_arm_line_on_the_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-secondary"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert (BASE_KEYS - {"impl"}) | {"roofline", "clocks", "gpu_launches"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["dtype"] == "f64" and d["unit"] == "evals/s"
    assert d["gpu_launches"] >= 5 and d["value"] > 1e10
    e = d["e2e"]
    assert 0 < e["value"] <= d["value"] * 1.05 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 * max(1.0, r["frac"])
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
