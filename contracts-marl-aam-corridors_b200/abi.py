"""ctypes mirror of ``include/gmpe.h`` and loader of the CUDA library.

The product path has NO CPU fallback: :func:`load_library` raises if the in-tree
``libgmpe_b200.so`` (built by ``__graft_entry__.build()`` / ``csrc/build.py``) is
missing.  Struct layouts here must match the header byte for byte; the ``not gpu``
test ``tests/test_abi.py`` checks sizes and every exported symbol.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 13

# error codes
OK, E_INVALID, E_CUDA, E_NULL = 0, -1, -2, -3

SCEN_JULY_TUBE, SCEN_NAV_GRAPH, SCEN_THREE_PHASE, SCEN_TWO_PHASE, SCEN_CORRIDOR_ROT_INV = 0, 1, 2, 3, 4
ROT_INV_SCENARIOS = (SCEN_THREE_PHASE, SCEN_TWO_PHASE, SCEN_CORRIDOR_ROT_INV)
# the lane-per-agent kernel is instantiated for these world sizes (README.md:88 and
# eval_scripts/commands_test_run.txt:7,30 of the reference: 3, 5 and 10 agents)
ROT_INV_AGENT_COUNTS = tuple(range(2, 17))  # gmpe_coop_table.h: COOP_MAX_AGENTS_ROT_INV
DYN_DOUBLE_INTEGRATOR, DYN_UNICYCLE, DYN_AIR_TAXI = 0, 1, 2
PHYS_KINEMATIC, PHYS_FORCE = 0, 1
F32, F64 = 0, 1

FLAG_USE_DONES = 1 << 0
FLAG_COLLABORATIVE = 1 << 1
FLAG_AUTO_RESET = 1 << 2
FLAG_ENV_EDGES = 1 << 3
FLAG_GNN_EDGES = 1 << 4
FLAG_INFO = 1 << 5
FLAG_WALL_POINT_QUIRK = 1 << 6
FLAG_CLASSIC_CONTACT = 1 << 7
FLAG_FORCE_TILE = 1 << 8
FORMATION_POINT, FORMATION_LINE, FORMATION_CIRCLE = 0, 1, 2
FORMATIONS = {"point": FORMATION_POINT, "line": FORMATION_LINE, "circle": FORMATION_CIRCLE}

MAX_AGENTS = 64
MAX_ENTITIES = 160
WORLD_F, WORLD_I, AGENT_CNT, INFO_K, WALL_F = 16, 4, 8, 16, 4

# world_f slots
WF_ENT_X, WF_ENT_Y, WF_EXIT_X, WF_EXIT_Y, WF_E_X, WF_E_Y, WF_N_X, WF_N_Y = range(8)
WF_L, WF_HALF_W, WF_WIDTH, WF_ANGLE, WF_DELTA_SPACING = 8, 9, 10, 11, 12
WI_STEP, WI_EPISODE = 0, 1
AF_STATUS, AF_PREV_PHASE, AF_PHASE_REACHED, AF_GOAL_REACHED = 0, 1, 2, 3
(AC_SPACING_VIOL, AC_STEPS_IN_CORRIDOR, AC_CONFORMANCE, AC_AGENT_COLL, AC_OBST_COLL,
 AC_TIME_REQ_I, AC_DIST_TRAV_I, AC_DIST_LEFT_I) = range(8)

#: info columns in reference dict order (july:806-829)
INFO_KEYS = (
    "Dist_to_goal", "Time_req_to_goal", "Num_agent_collisions", "Num_obst_collisions",
    "Distance_mean", "Distance_variance", "Mean_by_variance", "Dists_traveled",
    "Time_taken", "Time_mean", "Time_stddev", "Time_mean_by_stddev", "Conformance",
    "Delta_spacing", "Spacing_violations", "Min_time_to_goal",
)
#: info dict of the rot-inv family (three_phase_graph.py:788-811, two_phase_graph.py:786-809,
#: nav_graph_metered_single_corridor_rot_inv.py:822-845): 'Phase_reached' sits before 'Min_time_to_goal';
#: its value is GmpeStepOut.phase_reached, not a column of the 16-wide info row
INFO_KEYS_ROT_INV = INFO_KEYS[:-1] + ("Phase_reached",) + INFO_KEYS[-1:]



class GmpeConfig(C.Structure):
    _fields_ = [
        ("scenario", C.c_int32), ("dynamics", C.c_int32), ("physics", C.c_int32), ("precision", C.c_int32),
        ("num_worlds", C.c_int32), ("num_agents", C.c_int32), ("num_landmarks", C.c_int32),
        ("num_obstacles", C.c_int32), ("num_walls", C.c_int32), ("episode_length", C.c_int32),
        ("total_actions", C.c_int32), ("flags", C.c_uint32),
        ("formation", C.c_int32), ("reserved0", C.c_int32),
        ("seed", C.c_uint64), ("world_offset", C.c_int64),
        ("world_size", C.c_double),
        ("goal_rew", C.c_double), ("collision_rew", C.c_double), ("formation_rew", C.c_double),
        ("max_speed", C.c_double), ("dt", C.c_double), ("v_min", C.c_double), ("v_max", C.c_double),
        ("accel_min", C.c_double), ("accel_max", C.c_double), ("ang_rate_max", C.c_double),
        ("goal_dist_thresh", C.c_double), ("separation_dist", C.c_double),
        ("coordination_range", C.c_double), ("gnn_max_edge_dist", C.c_double),
        ("entity_size", C.c_double), ("sensitivity", C.c_double),
        ("damping", C.c_double), ("contact_force", C.c_double), ("contact_margin", C.c_double),
        ("wall_contact_force", C.c_double), ("wall_contact_margin", C.c_double), ("wall_width", C.c_double),
        ("min_reward", C.c_double), ("max_reward", C.c_double),
    ]

    def copy(self) -> "GmpeConfig":
        out = GmpeConfig()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(GmpeConfig))
        return out

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class GmpeLayout(C.Structure):
    _fields_ = [
        ("real_bytes", C.c_int32), ("num_entities", C.c_int32), ("obs_dim", C.c_int32),
        ("node_dim", C.c_int32), ("edge_cap", C.c_int32), ("info_dim", C.c_int32),
    ] + [(n, C.c_int64) for n in (
        "agent_state", "agent_aux", "agent_flags", "agent_cnt", "world_f", "world_i",
        "landmarks", "obstacles", "walls", "agent_ext", "obs", "node_obs", "adj", "reward", "done", "info",
        "edge_index", "edge_attr", "num_edges")]


class GmpeState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "agent_state", "agent_aux", "agent_flags", "agent_cnt", "world_f", "world_i",
        "landmarks", "obstacles", "walls", "agent_ext")]


class GmpeStepOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "obs", "node_obs", "adj", "reward", "done", "info", "edge_index", "edge_attr", "num_edges",
        "gnn_edge_index", "gnn_edge_attr", "gnn_num_edges", "indiv_reward", "mask", "active_mask", "phase_reached")]


class GmpeStepInject(C.Structure):
    _fields_ = [("theta_draw", C.c_void_p), ("reset_state", C.POINTER(GmpeState)),
                ("reset_draws", C.c_void_p), ("reset_draws_per_world", C.c_int32), ("reserved0", C.c_int32)]


#: every symbol include/gmpe.h declares
EXPORTED_SYMBOLS = (
    "gmpe_abi_version", "gmpe_create", "gmpe_destroy", "gmpe_last_error", "gmpe_layout",
    "gmpe_reset", "gmpe_step", "gmpe_decode_onehot", "gmpe_launch_count", "gmpe_set_seed",
    "gmpe_kernel_name", "gmpe_step_range", "gmpe_decode_onehot_range", "gmpe_step_host", "gmpe_host_wait",
    "gmpe_placement_overflows", "gmpe_gnn_flatten", "gmpe_gnn_flat_scratch_elems",
)

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "csrc", "libgmpe_b200.so")

_lib = None


class GmpeError(RuntimeError):
    pass


def load_library(path: str | None = None) -> C.CDLL:
    """Load the in-tree CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("GMPE_LIB") or LIB_PATH  # GMPE_LIB: tuning variants (tools/sweep.py)
    if not os.path.exists(p):
        raise GmpeError(
            f"CUDA extension not built: {p} is missing. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the GraphMPE step path.")
    lib = C.CDLL(p)
    lib.gmpe_abi_version.restype = C.c_int
    lib.gmpe_create.argtypes = [C.POINTER(GmpeConfig), C.c_int, C.POINTER(C.c_void_p)]
    lib.gmpe_create.restype = C.c_int
    lib.gmpe_destroy.argtypes = [C.c_void_p]
    lib.gmpe_destroy.restype = C.c_int
    lib.gmpe_last_error.argtypes = [C.c_void_p]
    lib.gmpe_last_error.restype = C.c_char_p
    lib.gmpe_layout.argtypes = [C.POINTER(GmpeConfig), C.POINTER(GmpeLayout)]
    lib.gmpe_layout.restype = C.c_int
    lib.gmpe_reset.argtypes = [C.c_void_p, C.POINTER(GmpeState), C.c_void_p, C.POINTER(GmpeStepInject),
                               C.POINTER(GmpeStepOut), C.c_void_p]
    lib.gmpe_reset.restype = C.c_int
    lib.gmpe_step.argtypes = [C.c_void_p, C.POINTER(GmpeState), C.c_void_p, C.POINTER(GmpeStepInject),
                              C.POINTER(GmpeStepOut), C.c_void_p]
    lib.gmpe_step.restype = C.c_int
    lib.gmpe_decode_onehot.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.gmpe_decode_onehot.restype = C.c_int
    lib.gmpe_step_range.argtypes = [C.c_void_p, C.POINTER(GmpeState), C.c_void_p, C.POINTER(GmpeStepInject),
                                    C.POINTER(GmpeStepOut), C.c_int64, C.c_int64, C.c_void_p]
    lib.gmpe_step_range.restype = C.c_int
    lib.gmpe_decode_onehot_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.gmpe_decode_onehot_range.restype = C.c_int
    lib.gmpe_step_host.argtypes = [C.c_void_p, C.POINTER(GmpeState), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(GmpeStepOut), C.POINTER(GmpeStepOut), C.c_int, C.c_void_p]
    lib.gmpe_step_host.restype = C.c_int
    lib.gmpe_host_wait.argtypes = [C.c_void_p]
    lib.gmpe_host_wait.restype = C.c_int
    lib.gmpe_launch_count.argtypes = [C.c_void_p]
    lib.gmpe_launch_count.restype = C.c_int64
    lib.gmpe_kernel_name.argtypes = [C.c_void_p]
    lib.gmpe_kernel_name.restype = C.c_char_p
    lib.gmpe_placement_overflows.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    lib.gmpe_placement_overflows.restype = C.c_int
    lib.gmpe_gnn_flat_scratch_elems.argtypes = [C.c_void_p]
    lib.gmpe_gnn_flat_scratch_elems.restype = C.c_int64
    lib.gmpe_gnn_flatten.argtypes = [C.c_void_p, C.POINTER(GmpeStepOut), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    lib.gmpe_gnn_flatten.restype = C.c_int
    lib.gmpe_set_seed.argtypes = [C.c_void_p, C.c_uint64, C.c_int64]
    lib.gmpe_set_seed.restype = C.c_int
    v = lib.gmpe_abi_version()
    if v != ABI_VERSION:
        raise GmpeError(f"ABI mismatch: library reports {v}, python mirror expects {ABI_VERSION}; rebuild")
    if path is None:
        _lib = lib
    return lib


def layout_for(cfg: GmpeConfig) -> GmpeLayout:
    """Array sizes for ``cfg``.  Pure arithmetic mirrored from ``gmpe_layout`` so that
    hosts (and the CPU-only tests) can size buffers without the CUDA library."""
    lay = GmpeLayout()
    B, N, L, O, W = cfg.num_worlds, cfg.num_agents, cfg.num_landmarks, cfg.num_obstacles, cfg.num_walls
    E = N + L + O + W
    lay.real_bytes = 8 if cfg.precision == F64 else 4
    lay.num_entities = E
    lay.obs_dim = {SCEN_JULY_TUBE: 19, SCEN_NAV_GRAPH: 6, SCEN_THREE_PHASE: 15, SCEN_TWO_PHASE: 15, SCEN_CORRIDOR_ROT_INV: 13}[cfg.scenario]
    lay.node_dim = 7 if cfg.scenario in ROT_INV_SCENARIOS else 8
    lay.edge_cap = E * (E - 1)
    lay.info_dim = INFO_K
    lay.agent_state = B * N * 4
    lay.agent_aux = B * N * 2
    lay.agent_flags = B * N * 4
    lay.agent_cnt = B * N * AGENT_CNT
    lay.world_f = B * WORLD_F
    lay.world_i = B * WORLD_I
    lay.landmarks = B * L * 2
    lay.obstacles = B * O * 2
    lay.walls = B * W * WALL_F
    lay.agent_ext = B * N * 2
    lay.obs = B * N * lay.obs_dim
    lay.node_obs = B * N * E * lay.node_dim
    lay.adj = B * E * E
    lay.reward = B * N
    lay.done = B * N
    lay.info = B * N * INFO_K
    lay.edge_index = B * 2 * lay.edge_cap
    lay.edge_attr = B * lay.edge_cap
    lay.num_edges = B
    return lay
