"""bench.py's reference arm runs without a GPU: its JSON line must carry the contract's keys and the very metric / config the
GPU arm prints (the driver computes the ratio of the two lines itself)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-log2n", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config",
                "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["unit"] == "Gelem/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["config"]["workload"] == "form1_sum_f32_int64labels_N1073741824_G100000_uniform"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["vs_baseline"] is None


import pytest


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--log2n", "24", "--steps", "4", "--warmup", "3", "--no-sweep",
                          "--cpu-log2n", "18"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "verified"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 4 and line["unit"] == "Gelem/s" and line["value"] > 0
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["kernel_ms"] > 0 and rf["kernel_ms"] <= line["ms_per_step"] * 1.05
    assert line["gpu_launches"] > 0
    assert line["verified"]["ok"] is True
    assert line["e2e"]["h2d_bytes_per_step"] == (1 << 24) * 12 and line["e2e"]["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] > 0
