"""The bench line's contract, checked on the CPU: the committed round-2 records carry every key the driver reads, their
derived figures follow from their own fields, and the reference arm prints the same metric / config."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.load(f)


@pytest.mark.parametrize("name,n", [("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4), ("r02_bench_n8.json", 8)])
def test_committed_records_follow_the_contract(name, n):
    d = _load(name)
    for k in KEYS:
        assert k in d, k
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    px = d["config"]["images_per_gpu"] * d["config"]["label_hw"][0] * d["config"]["label_hw"][1]
    assert abs(d["value"] - n * px / d["ms_per_step"] / 1e6) <= 1e-6 * d["value"]          # whole-job Gpixel/s from ms_per_step
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["frac"] < 1
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / r["launch_ms"] / 1e6) <= 1e-6 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] >= 0.95 * r["algorithmic_bytes_per_launch"]
    # the one-call step is two kernels (+ one flush kernel at the closing join when sharded)
    assert d["gpu_launches"] == 2 * d["steps"] + (1 if n > 1 else 0)
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"] and 0 < c["value"] < e["value"]
        p = d["cpu_baseline"]["confusion_hist_port"]
        assert p["matrix_bit_exact_vs_gpu"] is True and p["miou_bit_exact_vs_gpu"] is True
    else:
        assert d["stats_check"]["ok"] and d["cfg3_multi_level"]["check"]["ok"] and d["cfg5_crosscity"]["check"]["ok"]


def test_reference_arm_prints_the_same_metric_and_config():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    rec = _load("r02_bench_n1.json")
    assert line["impl"] == "reference" and line["metric"] == rec["metric"] and line["unit"] == rec["unit"]
    assert line["config"] == rec["config"] and line["higher_is_better"] is True
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
