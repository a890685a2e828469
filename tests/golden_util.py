"""Helpers shared by the golden-trace tests: load a trace, rebuild the config the
reference ran with, and expose teacher-forcing inputs in the C-ABI array shapes."""
from __future__ import annotations

import glob
import json
import os
from types import SimpleNamespace

import numpy as np

import gmpe_b200
from gmpe_b200 import abi, make_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def traces(pattern="july_*.npz"):
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, pattern)))


class Trace:
    def __init__(self, path):
        self.path = path
        self.name = os.path.basename(path)
        z = np.load(path, allow_pickle=False)
        self.z = {k: z[k] for k in z.files}
        self.args = SimpleNamespace(**json.loads(str(self.z["cfg_json"])))
        self.T = int(self.z["T"])
        self.episodes = int(self.z["episodes"])
        self.N = self.args.num_agents

    def cfg(self, precision="fp64", **kw):
        kw.setdefault("auto_reset", False)
        kw.setdefault("gnn_edges", True)
        return make_config(self.args, 1, precision=precision, **kw)

    def reset_state(self, ep, real=np.float64):
        """Initial condition of episode `ep` as C-ABI state arrays (B=1)."""
        z = self.z
        N = self.N
        aux = np.zeros((1, N, 2), dtype=real)
        aux[0, :, 1] = z["init_goal_min_time"][ep]
        wf = np.zeros((1, abi.WORLD_F), dtype=real)
        wf[0, :12] = z["init_tube"][ep]
        L = self.args.num_landmarks
        return {
            "agent_state": z["init_agent_state"][ep][None].astype(real),
            "agent_aux": aux,
            "world_f": wf,
            "landmarks": z["init_landmarks"][ep][None].astype(real),
            "obstacles": np.zeros((1, self.args.num_obstacles, 2), dtype=real),
            "walls": np.zeros((1, getattr(self.args, "num_walls", 0), abi.WALL_F), dtype=real),
        }

    def step_rows(self, ep):
        return range(ep * self.T, (ep + 1) * self.T)


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor): the 1e-5 relative fp32 tolerance of BASELINE.json's
    north_star, with an absolute floor for near-zero coordinates."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
