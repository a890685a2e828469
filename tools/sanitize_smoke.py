#!/usr/bin/env python
"""Small end-to-end exercise of every kernel mapping for compute-sanitizer (run under gpurun:
`compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmpe_b200  # noqa: E402,F401
from gmpe_b200 import default_args, make_config  # noqa: E402
from gmpe_b200.engine import BatchedGraphMPE  # noqa: E402

CASES = [
    dict(scenario_name="navigation_graph", episode_length=4, collaborative=True),
    dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", world_size=3, episode_length=4,
         goal_rew=30, collaborative=False),
    dict(scenario_name="navigation_graph", num_agents=5, num_landmarks=6, num_obstacles=3, num_walls=4, world_size=2,
         episode_length=4),
    dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", num_agents=64,
         num_landmarks=64, world_size=12, episode_length=3, goal_rew=30, collaborative=False),
]
from gmpe_b200 import config as _config  # noqa: E402

CASES.append(dict(scenario_name="three_phase_graph", dynamics_type="air_taxi", world_size=3, num_agents=5, num_landmarks=5,
                  episode_length=4, goal_rew=20, collaborative=False))
for kern in ("auto", "tile"):
    _config.FORCE_TILE_DEFAULT = kern == "tile"  # the lane-per-agent kernel where gmpe_create picks it / the tile kernel forced
    for kw in CASES:
        for prec in ("fp32", "fp64"):
            B = 77 if kw.get("num_agents", 3) < 64 else 5
            cfg = make_config(default_args(**kw), B, precision=prec, seed=1, gnn_edges=True)
            eng = BatchedGraphMPE(cfg, 0)
            eng.reset()
            g = torch.Generator().manual_seed(0)
            for t in range(7):
                a = torch.randint(0, cfg.total_actions, (B, cfg.num_agents), generator=g, dtype=torch.int32).cuda()
                eng.step(a)
            m = torch.zeros(B, dtype=torch.uint8)
            m[::3] = 1
            eng.reset(env_mask=m.cuda())
            eng.step_range(a, B // 5, B // 2)
            ei, ea = eng.gnn_flat_edges()  # scan + gather kernels (gmpe_flat.cuh)
            torch.cuda.synchronize()
            print(kern, kw["scenario_name"][:14], kw.get("num_agents", 3), prec, eng.kernel_name, float(eng.out["reward"].sum()))
            eng.close()
print("sanitize smoke done")
