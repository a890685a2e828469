#!/usr/bin/env python
"""Sweep of the host-facing step's slicing (gmpe_step_host) -- run under gpurun.

    python tools/e2e_sweep.py            # builds a GMPE_EXPERIMENTS variant, prints ms per venv.step

Times venv.step for every number of upload blocks (pipeline_chunks; 1 = the plain torch path) and
prints the PCIe ceiling of the box: the same byte counts moved by plain pinned copies, both
directions at once.  --trace prints the timeline of single steps from a GMPE_EXPERIMENTS build
(profiles/r2_e2e_timeline.txt holds the one that showed why copies in opposite directions must not
overlap on these boxes)."""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "contracts-marl-aam-corridors_b200/csrc/build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
if "--trace" in sys.argv:  # the timeline print-out lives in GMPE_EXPERIMENTS builds only
    so = os.path.join(ROOT, "build", "variants", "lib_experiments.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if "--no-build" not in sys.argv:
        b.build(force=True, out=so, defines=("GMPE_EXPERIMENTS",))
    os.environ["GMPE_LIB"] = so

import torch  # noqa: E402

import gmpe_b200  # noqa: E402,F401
from gmpe_b200 import default_args  # noqa: E402
from gmpe_b200.env_wrappers import GraphSubprocVecEnv  # noqa: E402

B, N, A = 65536, 3, 5
args = default_args(scenario_name="navigation_graph", num_agents=3, num_landmarks=3, episode_length=25, collaborative=True)
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
onehots = [np.eye(A)[rs.randint(A, size=(B, N))] for _ in range(4)]


def time_steps(venv, n=40):
    for t in range(5):
        venv.step(onehots[t % 4], 0)
    torch.cuda.synchronize()
    ts = []
    for t in range(n):
        t0 = time.perf_counter()
        o = venv.step(onehots[t % 4], 0)
        float(o[4].reshape(-1)[0])
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    return float(np.median(ts)), float(ts.mean()), float(ts.min())


def pcie_ceiling(d2h_bytes, h2d_bytes, n=30):
    pin_o = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    pin_i = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    d_o = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    d_i = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ts = []
    for t in range(n + 3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            pin_o.copy_(d_o, non_blocking=True)
        with torch.cuda.stream(s2):
            d_i.copy_(pin_i, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts[3:])) * 1e3


venv = GraphSubprocVecEnv(args=args, num_envs=B, device=dev, seed=1, info=True)
venv.reset()
venv.step(onehots[0], 0)
d2h, h2d = venv.d2h_bytes_last, venv.h2d_bytes_last
venv.close()
ms = pcie_ceiling(d2h, h2d)
print(f"pcie ceiling: {d2h} B down + {h2d} B up concurrently: {ms:.3f} ms -> {B * N / ms / 1e3:.1f} M agent-steps/s "
      f"({d2h / ms / 1e6:.1f} GB/s down)", flush=True)
if "--trace" in sys.argv:  # timeline of single steps (GMPE_EXPERIMENTS build prints it from gmpe_host_wait)
    for chunks in (2, 4, 8):
        venv = GraphSubprocVecEnv(args=args, num_envs=B, device=dev, seed=1, info=True, pipeline_chunks=chunks)
        venv.reset()
        med, mean, mn = time_steps(venv, 20)
        print(f"--- upload blocks {chunks}: median {med:.3f} ms", flush=True)
        os.environ["GMPE_TRACE"] = "1"
        for t in range(3):
            t0 = time.perf_counter()
            venv.step(onehots[t % 4], 0)
            print(f"    venv.step wall {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr, flush=True)
        os.environ["GMPE_TRACE"] = "0"
        venv.close()
    sys.exit(0)
for rep in range(2):
    for chunks in (1, 2, 3, 4, 6, 8, 16):
        venv = GraphSubprocVecEnv(args=args, num_envs=B, device=dev, seed=1, info=True, pipeline_chunks=chunks)
        venv.reset()
        med, mean, mn = time_steps(venv)
        venv.close()
        print(f"upload blocks {chunks:2d}: median {med:.3f} ms  mean {mean:.3f}  min {mn:.3f}  "
              f"-> {B * N / med / 1e3:.1f} M agent-steps/s", flush=True)
