#!/usr/bin/env python
"""Reset parity fixture (VERDICT r1 next-round #2b): what the LIVE reference's
`MultiAgentGraphEnv.reset()` -> `Scenario.reset_world / random_scenario / setup_tube_params`
(july:339-613, three_phase_graph.py:416-587, utils.py:77-262) places, together with the uniform
[0, 1) numbers it drew from np.random to do so.

    python tests/golden/make_reset_golden.py        # build container only (needs /root/reference)

Every random call of the reset path is `np.random.uniform`, which consumes one double of the
global MT19937 stream per element (low + (high - low) * random_sample()): the draws of a reset
seeded with `s` are therefore `RandomState(s).random_sample(K)`.  The fixture stores, per case,
R resets for the exact comparison (seed of the draws in, placement out) and -- for two cases -- a few thousand
placements for the two-sample KS test of the Philox-driven reset against the reference's
distribution.  Written to tests/golden/reset_parity.npz."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import build_july_env, install_shims, july_args  # noqa: E402

K = 96  # draws kept per reset (N = 8 with rejections uses < 40)

CASES = {
    "july_point_n3": dict(),
    "july_line_n3": dict(formation_type="line"),
    "july_circle_n3": dict(formation_type="circle"),
    "july_point_n8": dict(num_agents=8, num_landmarks=8, world_size=5),
    "july_unicycle_n3": dict(dynamics_type="unicycle_vehicle"),
    "three_phase_n3": dict(scenario_name="three_phase_graph"),
    "two_phase_n5": dict(scenario_name="two_phase_graph", num_agents=5, num_landmarks=5),
    "corridor_rot_inv_n3": dict(scenario_name="nav_graph_metered_single_corridor_rot_inv"),
}
KS_CASES = {"july_point_n3": 2000, "three_phase_n3": 2000}


def tube_vector(w):
    tp = w.tube_params
    return np.array([tp["entrance"][0], tp["entrance"][1], tp["exit"][0], tp["exit"][1], tp["e"][0], tp["e"][1],
                     float(tp["n"][0]), float(tp["n"][1]), tp["L"], tp["half_width"], tp["width"], tp["angle"]], dtype=np.float64)


def one_reset(env, seed):
    np.random.seed(seed)
    env.reset()
    w = env.world
    return dict(agent_state=np.array([a.state.values.copy() for a in w.agents]),
                landmarks=np.array([l.state.p_pos.copy() for l in w.landmarks]), tube=tube_vector(w),
                goal_min_time=np.array([a.goal_min_time for a in w.agents], dtype=np.float64))


def main():
    install_shims()
    from multiagent.config import UnicycleVehicleConfig
    UnicycleVehicleConfig.COORDINATION_RANGE = UnicycleVehicleConfig.COMMUNICATION_RANGE  # P3
    out = {}
    cfgs = {}
    for name, over in CASES.items():
        args = july_args(episode_length=50, **over)
        env, sc = build_july_env(args, seed=1000)
        R = 64
        recs = [one_reset(env, 5000 + r) for r in range(R)]
        out[f"{name}/seeds"] = np.arange(5000, 5000 + R, dtype=np.int64)  # draws = RandomState(seed).random_sample(K)
        for k in recs[0]:
            out[f"{name}/{k}"] = np.stack([x[k] for x in recs])
        cfgs[name] = vars(args)
        # how many draws the reference really consumed: the global stream's position after the reset
        used = []
        for r in range(R):
            np.random.seed(5000 + r)
            env.reset()
            nxt = np.random.random_sample()
            stream = np.random.RandomState(5000 + r).random_sample(K + 1)
            used.append(int(np.where(stream == nxt)[0][0]))
        out[f"{name}/draws_used"] = np.array(used, dtype=np.int32)
        if name in KS_CASES:
            M = KS_CASES[name]
            recs = [one_reset(env, 100000 + r) for r in range(M)]
            out[f"{name}/ks_agent_state"] = np.stack([x["agent_state"][:, :3] for x in recs]).astype(np.float32)  # x, y, theta
            out[f"{name}/ks_tube"] = np.stack([x["tube"][[11, 8]] for x in recs]).astype(np.float32)            # angle, L
        print(name, "resets", R, "draws used", min(used), "..", max(used), flush=True)
    out["cfg_json"] = np.array(json.dumps(cfgs))
    path = os.path.join(HERE, "reset_parity.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
