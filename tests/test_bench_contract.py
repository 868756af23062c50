"""CPU: the benchmark's JSON contract. (1) the committed line of the last GPU run (profiles/r01_bench_n1.json) carries
every key the driver reads, with consistent arithmetic; (2) `bench.py --impl reference` — the CPU arm — runs here and
prints a line of the same shape; (3) the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _check_common(line):
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["unit"] == "patches/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert "16k" in line["metric"] and "workload" in line["config"] and "model" not in line["config"]
    assert line["vs_baseline"] is None            # BASELINE.md publishes no number for this metric
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])


def test_committed_gpu_line_has_every_contract_key():
    with open(os.path.join(ROOT, "profiles", "r01_bench_n1.json")) as f:
        line = json.load(f)
    _check_common(line)
    assert line["dtype"] == "bf16" and line["data"] == "synthetic" and line["n_gpus"] == 1
    assert line["warmup"] >= 3 and line["gpu_launches"] >= 4 * line["steps"]
    bag = line["config"]["bag"]
    assert bag == [16384, 1024]
    # value is whole-job throughput: patches per step / time per step
    assert abs(line["value"] - bag[0] / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]
    # e2e moves the whole bf16 bag host -> device every step and is a distinct measurement
    assert line["e2e"]["h2d_bytes_per_step"] == bag[0] * bag[1] * 2 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["value"] < line["value"]
    roof = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof)
    assert roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s"
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")          # driver-written per pod, not tracked
    if os.path.exists(peaks):
        with open(peaks) as f:
            measured = json.load(f)
        if measured.get("gpu_name") == "NVIDIA B200" and "measured" in roof.get("peak_source", ""):
            assert abs(roof["peak"] - measured["bf16_tflops"]) <= 0.15 * measured["bf16_tflops"]   # same pool, other box
    # algorithmic FLOPs of one forward launch: 2 N (1024 L + 2 L D), big preset
    assert roof["flops_per_launch"] == 2 * 16384 * (1024 * 512 + 2 * 512 * 384)
    assert roof["traffic"] is None or roof["traffic"] >= 16384 * 2048      # at least the bag itself
    clocks = line["clocks"]
    assert not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1


def test_reference_arm_runs_on_the_host_cores():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    _check_common(line)
    assert line["impl"] == "reference" and line["dtype"] == "f32"
    assert line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_gpu_arm_fails_loudly_without_a_device():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
