#!/usr/bin/env python
"""Back-to-back step time of one configuration (no L2 flush), e.g. to compare kernel mappings:
   python tools/quick_step_time.py nav_metered_one_goal_graph_rotate_tube_july 8 32768 [tile]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmpe_b200  # noqa: E402,F401
from gmpe_b200 import default_args, make_config  # noqa: E402
from gmpe_b200.engine import BatchedGraphMPE  # noqa: E402

scen, n, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4 and sys.argv[4] == "tile":
    from gmpe_b200 import config as _config
    _config.FORCE_TILE_DEFAULT = True
dyn = "double_integrator" if scen == "navigation_graph" else "air_taxi"
args = default_args(scenario_name=scen, dynamics_type=dyn, num_agents=n, num_landmarks=n, world_size=max(3, n // 2),
                    episode_length=50, collaborative=False)
cfg = make_config(args, B, precision="fp32", seed=1, env_edges=True, info=True)
eng = BatchedGraphMPE(cfg, 0)
eng.reset()
act = torch.randint(0, cfg.total_actions, (B, n), dtype=torch.int32, device="cuda")
for _ in range(5):
    eng.step(act)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 40
for _ in range(K):
    eng.step(act)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / K * 1e3
gb = eng.algorithmic_bytes_per_env_step() * B / 1e9
print(f"{eng.kernel_name:24s} {scen} N={n} B={B}: {us:8.1f} us/step  {B * n / us * 1e6:.3e} agent-steps/s  {gb / us * 1e6:7.0f} GB/s")
