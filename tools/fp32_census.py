#!/usr/bin/env python
"""Census of fp32-vs-fp64 threshold flips at BASELINE batch sizes.

    python tools/fp32_census.py [--steps 200] [--out profiles/r2_fp32_census.json]

Teacher-forced comparison (tests/fp32_census.py) of the production fp32 kernels with the fp64
oracle over the WHOLE batch -- 65,536 navigation_graph worlds and 131,072 july worlds, 200
steps each, auto-resets included.  Prints one JSON object per workload: world-steps compared,
world-steps whose discrete outputs (done / edge_index / num_edges / gnn_*) differ, the largest
margin to a threshold among those (each one is asserted to be within MARGIN), how many
world-steps sat inside MARGIN at all, the excused real-valued rows, and the worst relative
error of every real output everywhere else (asserted <= 1e-5).  Needs a GPU; uses oracle/
as the checker (test infrastructure, like tests/)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    import fp32_census
    from bench import WORKLOADS
    from gmpe_b200 import default_args, make_config

    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--workloads", default="navigation_graph_n3_b65536,july_air_taxi_n3_b131072")
    ap.add_argument("--worlds", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_fp32_census.json"))
    opts = ap.parse_args()
    res = {}
    for name in opts.workloads.split(","):
        wl = WORKLOADS[name]
        B = opts.worlds or wl["worlds_per_gpu"]
        args = default_args(**wl["args"])
        kw = dict(seed=1, auto_reset=True, env_edges=True, gnn_edges=True, info=True)
        cfg64 = make_config(args, B, precision="fp64", **kw)
        cfg32 = make_config(args, B, precision="fp32", **kw)
        rs = np.random.RandomState(17)
        acts = rs.randint(cfg64.total_actions, size=(opts.steps, B, cfg64.num_agents)).astype(np.int32)
        t0 = time.perf_counter()
        cs = fp32_census.run_teacher_forced(cfg64, cfg32, acts, tol=1e-5, nthreads=os.cpu_count() or 8, label=name)
        d = cs.as_dict()
        d.update(worlds=B, steps=opts.steps, seconds=round(time.perf_counter() - t0, 1))
        res[name] = d
        # north_star: "edge_index, adjacency and done flags must match bit-exactly" -- in fp32 that holds
        # except at a threshold; the exceptions stay below 5e-6 of the world-steps at E = 6, scaled by the
        # number of entity pairs (the bound of tests/test_gpu_batch_vs_oracle.py)
        E = cfg64.num_agents + cfg64.num_landmarks + cfg64.num_obstacles + cfg64.num_walls
        d["int_mismatch_bound"] = 5e-6 * (E / 6.0) ** 2
        assert d["int_mismatch_rate"] <= d["int_mismatch_bound"], d
        print(json.dumps({name: d}), flush=True)
    with open(opts.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
