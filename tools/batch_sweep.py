#!/usr/bin/env python
"""Throughput of the step kernels against the batch size (SURVEY.md 8d asks for the sweep on
configs 4 and 5).  Run under gpurun: python tools/batch_sweep.py > gpurun_out/batch_sweep.txt"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWEEP = {
    "navigation_graph_n3_b65536": (16384, 65536, 262144, 1048576),
    "july_air_taxi_n3_b131072": (32768, 131072, 524288, 1048576),
    "three_phase_air_taxi_n3_b131072": (131072, 524288),
    "two_phase_air_taxi_n5": (65536, 262144),
    "three_phase_air_taxi_n10": (32768, 131072),
    "navigation_graph_n16_o8_w4": (4096, 16384, 65536, 262144),
    "july_air_taxi_n64": (1024, 4096, 16384, 32768),
}

print(f"{'workload':34s} {'worlds':>8s} {'us/step':>9s} {'agent-steps/s':>14s} {'GB/s':>7s} {'frac':>6s} {'hot us':>8s}", flush=True)
for wl, sizes in SWEEP.items():
    for b in sizes:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "30", "--warmup", "5", "--no-cpu", "--workload", wl,
                            "--worlds", str(b), "--e2e-steps", "1"], capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"{wl:34s} {b:8d} {d['ms_per_step'] * 1e3:9.1f} {d['value']:14.4g} {d['roofline']['achieved']:7.0f} {d['roofline']['frac']:6.3f}"
                  f" {d['hot_l2']['ms_per_step'] * 1e3:8.1f}", flush=True)
        except Exception:
            print(f"{wl} {b}: FAILED {r.stderr[-300:]}", flush=True)
