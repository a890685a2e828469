#!/usr/bin/env python
"""A few steps of one bench workload, for ncu (run under gpurun):
   ncu --set full --clock-control none --import-source on -k regex:gmpe_step -s 5 -c 2 -o gpurun_out/prof \\
       python tools/prof_step.py navigation_graph_n3_b65536"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from gmpe_b200 import default_args, make_config  # noqa: E402
from gmpe_b200.engine import BatchedGraphMPE  # noqa: E402

name = sys.argv[1]
wl = WORKLOADS[name]
args = default_args(**wl["args"])
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl["worlds_per_gpu"]
cfg = make_config(args, B, precision="fp32", seed=1, auto_reset=True, env_edges=True, gnn_edges=False, info=True)
eng = BatchedGraphMPE(cfg, 0)
eng.reset()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(3)
for t in range(12):
    a = torch.randint(0, cfg.total_actions, (B, cfg.num_agents), generator=g, device="cuda", dtype=torch.int32)
    flush.zero_()
    eng.step(a)
torch.cuda.synchronize()
print(name, eng.kernel_name, "done")
