#!/usr/bin/env python
"""A/B timing of build variants of the CUDA library (run under gpurun).
Usage: python tools/ab.py [--workload W] name1:DEF1=V,DEF2=V name2: ...
Each variant is built with its -D defines into build/variants/ and timed with bench.py."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "contracts-marl-aam-corridors_b200/csrc/build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)


def main():
    args = sys.argv[1:]
    wls = ["navigation_graph_n3_b65536"]
    build_only = False
    env_extra = {}
    variants = []
    while args:
        a = args.pop(0)
        if a == "--workload":
            wls = args.pop(0).split(",")
        elif a == "--build-only":
            build_only = True
        elif a == "--env":
            k, v = args.pop(0).split("=")
            env_extra[k] = v
        else:
            name, _, defs = a.partition(":")
            variants.append((name, tuple(d for d in defs.split(",") if d)))
    os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
    for name, defs in variants:
        so = os.path.join(ROOT, "build", "variants", f"lib_{name}.so")
        if not os.path.exists(so):
            b.build(force=True, out=so, defines=defs)
        if build_only:
            continue
        for wl in wls:
            env = dict(os.environ, GMPE_LIB=so, **env_extra)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "60", "--warmup", "5", "--no-cpu",
                                "--workload", wl, "--e2e-steps", "3"], env=env, capture_output=True, text=True)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                print(f"{name:24s} {wl:28s} {d['ms_per_step'] * 1e3:8.1f} us/step  hot {d['hot_l2']['ms_per_step'] * 1e3:8.1f} us"
                      f"  frac {d['roofline']['frac']:.3f}", flush=True)
            except Exception:
                print(f"{name}: FAILED {r.stderr[-400:]}", flush=True)


if __name__ == "__main__":
    main()
