#!/usr/bin/env python
"""Tuning sweep (run under gpurun): build launch-bound variants of the step kernel and time
the default workload with each.  Usage: python tools/sweep.py [workload]"""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "contracts-marl-aam-corridors_b200/csrc/build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
VARIANTS = [(256, 1, 256), (256, 2, 256), (256, 3, 256), (256, 4, 256), (128, 4, 128), (128, 6, 128), (128, 8, 128),
            (64, 8, 64), (64, 12, 64)]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "navigation_graph_n3_b65536"
    build_only = "--build-only" in sys.argv
    os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
    for lbt, minb, thr in VARIANTS:
        so = os.path.join(ROOT, "build", "variants", f"lib_{lbt}_{minb}.so")
        if not os.path.exists(so):
            b.build(force=True, out=so, defines=(f"GMPE_LB_THREADS={lbt}", f"GMPE_LB_MINBLOCKS={minb}"))
        if build_only:
            continue
        env = dict(os.environ, GMPE_LIB=so, GMPE_THREADS=str(thr))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "40", "--warmup", "5", "--no-cpu",
                            "--workload", wl, "--e2e-steps", "3"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"lb=({lbt},{minb}) threads={thr}: {d['ms_per_step'] * 1e3:8.1f} us/step  hot {d['hot_l2']['ms_per_step'] * 1e3:8.1f} us"
                  f"  frac {d['roofline']['frac']:.3f}", flush=True)
        except Exception:
            print(f"lb=({lbt},{minb}) threads={thr}: FAILED {r.stderr[-300:]}", flush=True)


if __name__ == "__main__":
    main()
