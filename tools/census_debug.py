#!/usr/bin/env python
"""Replay ONE world of a tools/fp32_census.py run up to step t on the oracle (worlds are independent and keyed by
their global id), then take that step on the GPU in fp32 and print where it differs most.
   python tools/census_debug.py two_phase_air_taxi_n5 55452 78 [65536]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import WORKLOADS  # noqa: E402
from gmpe_b200 import default_args, make_config  # noqa: E402
from gmpe_b200.engine import BatchedGraphMPE  # noqa: E402
from oracle.oracle import OracleEnv  # noqa: E402

name, world, t_fail = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
wl = WORKLOADS[name]
B = int(sys.argv[4]) if len(sys.argv) > 4 else wl["worlds_per_gpu"]
args = default_args(**wl["args"])
kw = dict(seed=1, auto_reset=True, env_edges=True, gnn_edges=True, info=True, world_offset=world)
c64, c32 = make_config(args, 1, precision="fp64", **kw), make_config(args, 1, precision="fp32", **kw)
rs = np.random.RandomState(17)
steps = t_fail + 1
acts = rs.randint(c64.total_actions, size=(max(steps, int(os.environ.get("CENSUS_STEPS", "100"))), B, c64.num_agents)).astype(np.int32)[:, world:world + 1]
orc = OracleEnv(c64, use_rk45=False)
orc.reset()
eng = BatchedGraphMPE(c32, 0)
for t in range(steps):
    if t == t_fail:
        eng.load_state_dict(orc.state)
        pre = {k: v.copy() for k, v in orc.state.items()}
    o = orc.step(acts[t])
g = eng.step(torch.as_tensor(acts[t_fail], device="cuda"), reset_state=orc.state)
np.set_printoptions(precision=9, suppress=False, linewidth=200)
print("pre agent_state\n", pre["agent_state"][0], "\nflags", pre["agent_flags"][0].tolist(), "world_i", pre["world_i"][0], "actions", acts[t_fail][0])
for k in ("obs", "node_obs", "adj", "reward", "info"):
    a, b = g[k].cpu().numpy().astype(np.float64)[0], o[k][0]
    err = np.abs(a - b) / np.maximum(1, np.abs(b))
    idx = np.unravel_index(np.nanargmax(err), err.shape)
    print(f"{k}: worst rel {np.nanmax(err):.3e} at {idx}: got {a[idx]!r} want {b[idx]!r}")
    if k == "node_obs":
        print(" got row ", a[idx[0], idx[1]], "\n want row", b[idx[0], idx[1]])
print("post agent_state (oracle)\n", orc.state["agent_state"][0], "\n(gpu)\n", eng.state["agent_state"][0].cpu().numpy())
