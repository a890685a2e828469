#!/usr/bin/env python
"""Time the UNMODIFIED Python reference step (build container only: needs /root/reference).
The reference cannot travel to the GPU box, so bench.py's CPU arms time the C port of it;
this script records what the real thing does on the container's cores for context."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_harness import build_july_env, install_shims, july_args  # noqa: E402


def main():
    install_shims()
    for name, T in (("nav_metered_one_goal_graph_rotate_tube_july", 300), ("three_phase_graph", 300)):
        args = july_args(scenario_name=name, episode_length=10 ** 6)
        env, sc = build_july_env(args, seed=1)
        env.reset()
        rs = np.random.RandomState(0)
        acts = [[np.eye(25)[rs.randint(25)] for _ in range(3)] for _ in range(T)]
        for t in range(20):
            env.step(acts[t])
        t0 = time.perf_counter()
        for t in range(T):
            env.step(acts[t])
        dt = time.perf_counter() - t0
        print(f"{name}: MultiAgentGraphEnv.step, N=3, 1 process: {T / dt:8.1f} env-steps/s = {3 * T / dt:8.1f} agent-steps/s "
              f"({dt / T * 1e3:.2f} ms/step) on {os.cpu_count()} vCPU box, python {sys.version.split()[0]}", flush=True)


if __name__ == "__main__":
    main()
