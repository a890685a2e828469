"""Constants of the reference that are not CLI flags, and the Namespace -> POD mapping.

Mirrors ``multiagent/config.py`` (AirTaxiConfig:4-33, UnicycleVehicleConfig:36-53,
DoubleIntegratorConfig:94-116, RewardWeightConfig:132-143) and the world constants of
``multiagent/core.py:537-566``.  :func:`make_config` turns the argparse ``Namespace``
the reference hands to ``Scenario.make_world(args)`` into the frozen ``GmpeConfig``
that crosses the C ABI.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

from . import abi


class AirTaxiConfig:  # multiagent/config.py:4-33
    V_MIN = 60 * 0.514444 * 0.001
    V_MAX = 175 * 0.514444 * 0.001
    ACCEL_MIN = -0.001
    ACCEL_MAX = 0.002
    ANGULAR_RATE_MAX = 0.1
    MOTION_PRIM_ACCEL_OPTIONS = 5
    MOTION_PRIM_ANGRATE_OPTIONS = 5
    DT = 1.0
    DISTANCE_TO_GOAL_THRESHOLD = 0.35
    SEPARATION_DISTANCE = 1500 * 0.0003048
    COLLISION_DISTANCE = SEPARATION_DISTANCE
    COORDINATION_RANGE = 3 * 1.60934


class UnicycleVehicleConfig:  # multiagent/config.py:36-53
    V_MIN = 0.4
    V_MAX = 0.75
    ACCEL_MIN = -0.5
    ACCEL_MAX = 0.5
    ANGULAR_RATE_MAX = 0.5
    MOTION_PRIM_ACCEL_OPTIONS = 5
    MOTION_PRIM_ANGRATE_OPTIONS = 5
    DT = 0.1
    DISTANCE_TO_GOAL_THRESHOLD = 0.2
    COLLISION_DISTANCE = 0.4
    COMMUNICATION_RANGE = 5
    # the reference scenarios read COORDINATION_RANGE, which only AirTaxiConfig defines
    # (SURVEY G5); the harness patch P3 and this build use COMMUNICATION_RANGE for it.
    COORDINATION_RANGE = COMMUNICATION_RANGE


class DoubleIntegratorConfig:  # multiagent/config.py:94-116
    VX_MAX = 1.0
    V_MIN = 0.0  # DoubleIntegratorXYState.min_speed (core.py:164)
    DT = 0.1
    DISTANCE_TO_GOAL_THRESHOLD = 0.2
    COLLISION_DISTANCE = 0.5
    COMMUNICATION_RANGE = 5
    COORDINATION_RANGE = COMMUNICATION_RANGE


class RewardWeightConfig:  # multiagent/config.py:132-143
    MIN_REWARD = -40
    MAX_REWARD = 50


DYNAMICS = {
    "double_integrator": (abi.DYN_DOUBLE_INTEGRATOR, DoubleIntegratorConfig),
    "unicycle_vehicle": (abi.DYN_UNICYCLE, UnicycleVehicleConfig),
    "air_taxi": (abi.DYN_AIR_TAXI, AirTaxiConfig),
}

#: scenario file names (``args.scenario_name``) this build serves
JULY_SCENARIOS = ("nav_metered_one_goal_graph_rotate_tube_july",)
NAV_SCENARIOS = ("navigation_graph",)
THREE_PHASE_SCENARIOS = ("three_phase_graph",)
TWO_PHASE_SCENARIOS = ("two_phase_graph",)
CORRIDOR_ROT_INV_SCENARIOS = ("nav_graph_metered_single_corridor_rot_inv",)


def default_args(**over) -> SimpleNamespace:
    """Defaults of the reference CLI for the flags that parameterise the hot path
    (onpolicy/config.py:172-484, onpolicy/scripts/train_mpe.py:71-116)."""
    d = dict(
        scenario_name="navigation_graph", dynamics_type="double_integrator",
        num_agents=3, num_landmarks=3, num_obstacles=0, num_walls=0, num_scripted_agents=0,
        world_size=2, episode_length=25, collaborative=True, use_dones=False,
        max_speed=2, collision_rew=5, goal_rew=5, formation_rew=1, fair_wt=1, fair_rew=1,
        formation_type="point", total_actions=5, zeroshift=5, min_obs_dist=0.5,
        num_env_steps=5_000_000, n_rollout_threads=1, render_episodes=None,
        graph_feat_type="relative", max_edge_dist=1, discrete_action=True,
        use_safety_filter=False, seed=1,
    )
    d.update(over)
    return SimpleNamespace(**d)


def num_discrete_actions(args) -> int:
    """environment.py:60-78: 25 motion primitives for unicycle / air_taxi regardless of
    ``--total_actions``; ``total_actions`` sizes the DI space only."""
    if args.dynamics_type in ("unicycle_vehicle", "air_taxi"):
        return 25
    return int(args.total_actions)


#: testing / A-B timing: make_config adds GMPE_FLAG_FORCE_TILE when set (tests/conftest.py's
#: ``kernel_variant`` fixture flips it so that both kernel mappings stay pinned to the oracle)
FORCE_TILE_DEFAULT = False


def make_config(args, num_worlds: int, *, precision: str = "fp32", seed: int | None = None,
                world_offset: int = 0, auto_reset: bool = True, env_edges: bool = True,
                gnn_edges: bool = False, info: bool = True, physics: str | None = None,
                wall_point_quirk: bool = False, classic_contact: bool = False,
                force_tile: bool | None = None) -> abi.GmpeConfig:
    name = args.scenario_name
    if name in JULY_SCENARIOS:
        scen = abi.SCEN_JULY_TUBE
    elif name in NAV_SCENARIOS:
        scen = abi.SCEN_NAV_GRAPH
    elif name in THREE_PHASE_SCENARIOS:
        scen = abi.SCEN_THREE_PHASE
    elif name in TWO_PHASE_SCENARIOS:
        scen = abi.SCEN_TWO_PHASE
    elif name in CORRIDOR_ROT_INV_SCENARIOS:
        scen = abi.SCEN_CORRIDOR_ROT_INV
    else:
        raise ValueError(f"scenario {name!r} is not served by this build (have "
                         f"{JULY_SCENARIOS + NAV_SCENARIOS + THREE_PHASE_SCENARIOS + TWO_PHASE_SCENARIOS + CORRIDOR_ROT_INV_SCENARIOS})")
    if args.dynamics_type not in DYNAMICS:
        raise NotImplementedError(args.dynamics_type)  # environment.py:77-78
    dyn, K = DYNAMICS[args.dynamics_type]
    if getattr(args, "use_safety_filter", False):
        raise NotImplementedError("the HJ/CBF safety filter (multiagent/safety_filter.py) is out of scope")
    if getattr(args, "num_scripted_agents", 0):
        raise NotImplementedError("scripted agents are out of scope")
    feat = getattr(args, "graph_feat_type", "relative")
    if feat != "relative":
        # july:1612-1616 / _get_entity_feat_global (:1672): 7 global-frame features per node while the
        # declared node space stays (E, 8) (environment.py:1004); no shipped command or checkpoint uses it
        raise NotImplementedError(f"graph_feat_type={feat!r}: only 'relative' node features are served "
                                  "(the reference's 'global' branch, july:1612-1616, is not built)")
    formation = getattr(args, "formation_type", "point")
    if formation not in abi.FORMATIONS:
        # july:492-499: anything else leaves the landmarks where the last episode put them
        raise NotImplementedError(f"formation_type={formation!r}: the reference places landmarks for "
                                  f"{tuple(abi.FORMATIONS)} only (july:492-497)")
    if scen == abi.SCEN_NAV_GRAPH and formation != "point":
        raise NotImplementedError("navigation_graph places its goals at random (DESIGN.md D1); formation_type must stay 'point'")
    if physics is None:
        physics = "force" if scen == abi.SCEN_NAV_GRAPH else "kinematic"
    c = abi.GmpeConfig()
    c.scenario = scen
    c.dynamics = dyn
    c.physics = abi.PHYS_FORCE if physics == "force" else abi.PHYS_KINEMATIC
    if c.physics == abi.PHYS_FORCE and dyn != abi.DYN_DOUBLE_INTEGRATOR:
        raise ValueError("force physics integrates (x,y,vx,vy): needs dynamics_type='double_integrator'")
    c.precision = {"fp32": abi.F32, "fp64": abi.F64}[precision]
    c.num_worlds = int(num_worlds)
    c.num_agents = int(args.num_agents)
    c.num_landmarks = int(args.num_landmarks)
    c.num_obstacles = int(args.num_obstacles)
    c.num_walls = int(getattr(args, "num_walls", 0))
    c.episode_length = int(args.episode_length)
    c.total_actions = num_discrete_actions(args)
    flags = 0
    if getattr(args, "use_dones", False):
        flags |= abi.FLAG_USE_DONES
    if getattr(args, "collaborative", False):
        flags |= abi.FLAG_COLLABORATIVE
    if auto_reset:
        flags |= abi.FLAG_AUTO_RESET
    if env_edges:
        flags |= abi.FLAG_ENV_EDGES
    if gnn_edges:
        flags |= abi.FLAG_GNN_EDGES
    if info:
        flags |= abi.FLAG_INFO
    if wall_point_quirk:
        flags |= abi.FLAG_WALL_POINT_QUIRK
    if classic_contact:
        flags |= abi.FLAG_CLASSIC_CONTACT
    if FORCE_TILE_DEFAULT if force_tile is None else force_tile:
        flags |= abi.FLAG_FORCE_TILE
    c.flags = flags
    c.formation = abi.FORMATIONS[formation]
    c.seed = int(args.seed if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
    c.world_offset = int(world_offset)
    c.world_size = float(args.world_size)
    c.goal_rew = float(args.goal_rew)
    c.collision_rew = float(args.collision_rew)
    c.formation_rew = float(getattr(args, "formation_rew", 1))
    c.max_speed = float(args.max_speed) if args.max_speed is not None else 0.0
    c.dt = float(K.DT)
    if dyn == abi.DYN_DOUBLE_INTEGRATOR:
        c.v_min, c.v_max = 0.0, float(K.VX_MAX)
        c.accel_min = c.accel_max = c.ang_rate_max = 0.0
    else:
        c.v_min, c.v_max = float(K.V_MIN), float(K.V_MAX)
        c.accel_min, c.accel_max = float(K.ACCEL_MIN), float(K.ACCEL_MAX)
        c.ang_rate_max = float(K.ANGULAR_RATE_MAX)
    c.goal_dist_thresh = float(K.DISTANCE_TO_GOAL_THRESHOLD)
    c.separation_dist = float(K.COLLISION_DISTANCE)
    c.coordination_range = float(K.COORDINATION_RANGE)
    c.gnn_max_edge_dist = float(getattr(args, "max_edge_dist", 1))
    c.entity_size = 0.06      # Entity.size, core.py:385
    c.sensitivity = 5.0       # environment.py:460
    c.damping = 0.25          # core.py:542
    if classic_contact:       # onpolicy/envs/mpe/core.py:129-130
        c.contact_force, c.contact_margin = 1e2, 1e-3
    else:                     # core.py:544-548
        c.contact_force, c.contact_margin = 3e2, 2e-2
    c.wall_contact_force, c.wall_contact_margin = 2.2e2, 2.4e-2
    c.wall_width = 0.1        # Wall.width, core.py:355
    c.min_reward = float(RewardWeightConfig.MIN_REWARD)
    c.max_reward = float(RewardWeightConfig.MAX_REWARD)
    validate(c)
    return c


def validate(c: abi.GmpeConfig) -> None:
    """Host-side mirror of the checks ``gmpe_create`` performs."""
    N, L, O, W = c.num_agents, c.num_landmarks, c.num_obstacles, c.num_walls
    if not (1 <= N <= abi.MAX_AGENTS):
        raise ValueError(f"num_agents must be in [1, {abi.MAX_AGENTS}]")
    if N + L + O + W > abi.MAX_ENTITIES:
        raise ValueError(f"too many entities (max {abi.MAX_ENTITIES})")
    if L < N:
        raise ValueError("num_landmarks must be >= num_agents (landmark_poses[agent.id], july:1740)")
    if c.scenario in abi.ROT_INV_SCENARIOS and not (N in abi.ROT_INV_AGENT_COUNTS and L == N and O == 0 and W == 0):
        raise ValueError(f"the rot-inv scenarios are served for num_agents == num_landmarks in {abi.ROT_INV_AGENT_COUNTS} "
                         "without obstacles / walls")
    if c.scenario in abi.ROT_INV_SCENARIOS and c.physics != abi.PHYS_KINEMATIC:
        raise ValueError("the rot-inv scenarios use the kinematic step")
    if c.scenario == abi.SCEN_JULY_TUBE and (O or W):
        # the reference graph builder raises for obstacles / walls in this scenario
        raise ValueError("the july tube scenario supports num_obstacles == num_walls == 0 only")
    if c.num_worlds < 1 or c.episode_length < 1:
        raise ValueError("num_worlds and episode_length must be positive")
    if c.dynamics == abi.DYN_DOUBLE_INTEGRATOR and c.total_actions not in (5, 9):
        raise ValueError("double_integrator supports total_actions 5 or 9 (environment.py:404-409)")
    if not math.isfinite(c.world_size) or c.world_size <= 0:
        raise ValueError("world_size must be positive")
    # bounded rejection sampling on the device: refuse what cannot be packed (gmpe_api.cu validate())
    rot = c.scenario in abi.ROT_INV_SCENARIOS
    side = 0.9 * c.world_size if c.scenario == abi.SCEN_NAV_GRAPH else (0.6 if rot else 0.4) * c.world_size
    # tube scenarios stagger the spawn points by 1/5 (july:470) or 1/3 (tp:430) per agent along one axis
    length = side if c.scenario == abi.SCEN_NAV_GRAPH else side + (N - 1) / (3.0 if rot else 5.0)
    need = N * math.pi * 0.25 * c.separation_dist ** 2 + O * math.pi * 16.0 * c.entity_size ** 2
    if need > 0.5 * side * length:
        raise ValueError("agents cannot be packed: N * pi * (separation_dist / 2)^2 exceeds half of the spawn area "
                         "(raise world_size)")
