"""bench.py prints ONE JSON line with the keys the driver and the judge read (task contract:
metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling /
vs_baseline / dtype / data / config.workload, roofline, cpu_baseline, e2e with byte counts,
gpu_launches, clocks) -- for both arms.  Small batch, few steps: seconds."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3", "--worlds", "4096",
                        "--e2e-steps", "3", *extra], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


@pytest.mark.gpu
def test_our_arm_prints_the_contract_line(kernel_variant):
    if kernel_variant != "auto":
        pytest.skip("one run is enough")
    d = _line("--no-extra")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["config"]["workload"] == "navigation_graph_n3_b65536" and d["steps"] == 4 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["gpu_launches"] == 4 and d["dtype"] == "f32" and d["vs_baseline"] is None
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "frac_survey_formula"):
        assert k in rf, k
    assert rf["bound"] == "hbm" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4096 * 3 * 5 * 8 and e["d2h_bytes_per_step"] > 4096 * 800
    assert 0 < e["pcie_ceiling"]["frac_of_it"] <= 1.05
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert len(d["per_rank_step_us"]) == 1


@pytest.mark.gpu
def test_reference_arm_prints_the_contract_line(kernel_variant):
    if kernel_variant != "auto":
        pytest.skip("one run is enough")
    d = _line("--impl", "reference")
    assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "navigation_graph_n3_b65536"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
