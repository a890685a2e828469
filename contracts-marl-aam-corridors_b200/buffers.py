"""Shapes of every state / output array of the C ABI and allocators for them.

One table, two backends: numpy (host; used by the CPU oracle wrapper in tests) and
torch (device; used by the engine).  Shapes follow ``include/gmpe.h``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi

STATE_FIELDS = ("agent_state", "agent_aux", "agent_flags", "agent_cnt", "world_f", "world_i",
                "landmarks", "obstacles", "walls", "agent_ext")
OUT_FIELDS = ("obs", "node_obs", "adj", "reward", "done", "info", "edge_index", "edge_attr",
              "num_edges", "gnn_edge_index", "gnn_edge_attr", "gnn_num_edges", "indiv_reward", "phase_reached")
#: optional float32 outputs (runner masks); only requested by DeviceGraphRollout
MASK_FIELDS = ("mask", "active_mask")


def shapes(cfg: abi.GmpeConfig) -> dict:
    """name -> (shape, kind) with kind in {'real', 'i32', 'u8'}."""
    lay = abi.layout_for(cfg)
    B, N, L, O, W = cfg.num_worlds, cfg.num_agents, cfg.num_landmarks, cfg.num_obstacles, cfg.num_walls
    E, cap = lay.num_entities, lay.edge_cap
    return {
        "agent_state": ((B, N, 4), "real"),
        "agent_aux": ((B, N, 2), "real"),
        "agent_flags": ((B, N, 4), "u8"),
        "agent_cnt": ((B, N, abi.AGENT_CNT), "i32"),
        "world_f": ((B, abi.WORLD_F), "real"),
        "world_i": ((B, abi.WORLD_I), "i32"),
        "landmarks": ((B, L, 2), "real"),
        "obstacles": ((B, O, 2), "real"),
        "walls": ((B, W, abi.WALL_F), "real"),
        "agent_ext": ((B, N, 2), "i32"),
        "obs": ((B, N, lay.obs_dim), "real"),
        "node_obs": ((B, N, E, lay.node_dim), "real"),
        "adj": ((B, E, E), "real"),
        "reward": ((B, N), "real"),
        "done": ((B, N), "u8"),
        "info": ((B, N, abi.INFO_K), "real"),
        "edge_index": ((B, 2, cap), "i32"),
        "edge_attr": ((B, cap), "real"),
        "num_edges": ((B,), "i32"),
        "gnn_edge_index": ((B, 2, cap), "i32"),
        "gnn_edge_attr": ((B, cap), "real"),
        "gnn_num_edges": ((B,), "i32"),
        "indiv_reward": ((B, N), "real"),
        "mask": ((B, N), "f32"),
        "active_mask": ((B, N), "f32"),
        "phase_reached": ((B, N), "u8"),
    }


def wanted_outputs(cfg: abi.GmpeConfig) -> tuple:
    names = ["obs", "node_obs", "adj", "reward", "done"]
    if cfg.flags & abi.FLAG_INFO:
        names += ["indiv_reward", "info"]
        if cfg.scenario in abi.ROT_INV_SCENARIOS:  # info['Phase_reached'] (three_phase_graph.py:806)
            names += ["phase_reached"]
    if cfg.flags & abi.FLAG_ENV_EDGES:
        names += ["edge_index", "edge_attr", "num_edges"]
    if cfg.flags & abi.FLAG_GNN_EDGES:
        names += ["gnn_edge_index", "gnn_edge_attr", "gnn_num_edges"]
    return tuple(names)


def alloc_numpy(cfg: abi.GmpeConfig, names, real=np.float64) -> dict:
    kinds = {"real": real, "i32": np.int32, "u8": np.uint8, "f32": np.float32}
    sh = shapes(cfg)
    return {n: np.zeros(sh[n][0], dtype=kinds[sh[n][1]]) for n in names}


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data if a.size else None
    # torch tensor
    assert a.is_contiguous()
    return a.data_ptr() if a.numel() else None


def make_struct(cls, arrays: dict):
    s = cls()
    for name, _ in cls._fields_:
        setattr(s, name, _ptr(arrays.get(name)))
    return s


def make_inject(theta=None, reset_state: dict | None = None, reset_draws=None):
    """Returns (GmpeStepInject, keepalive).  ``reset_draws``: float64 (B, K) uniform [0,1) numbers
    (numpy for the oracle, a device tensor for the engine) a reset consumes instead of Philox."""
    inj = abi.GmpeStepInject()
    keep = [theta, reset_draws]
    inj.theta_draw = _ptr(theta)
    if reset_draws is not None:
        assert reset_draws.ndim == 2
        inj.reset_draws = _ptr(reset_draws)
        inj.reset_draws_per_world = int(reset_draws.shape[1])
    if reset_state is not None:
        rs = make_struct(abi.GmpeState, reset_state)
        keep += [rs, reset_state]
        inj.reset_state = C.pointer(rs)
    return inj, keep
