"""Scenario plugin surface backed by the batched CUDA engine.

Mirrors the reference's ``Scenario`` class API (``multiagent/scenario.py:4-12`` and
``multiagent/custom_scenarios/nav_metered_one_goal_graph_rotate_tube_july.py:104``):
``make_world / reset_world / reward / observation / graph_observation / update_graph /
get_id / info_callback / done``.  The reference evaluates these per agent in Python; here
one kernel launch computes all of them for every agent of every world
(``World.step`` below), and the per-agent callbacks are O(1) views into that step's
outputs, in the reference's dtypes' shapes.  ``World`` / ``Agent`` expose the attributes
the reference's env classes and runner read (``environment.py:35-54, 968-970``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import abi
from .config import make_config
from .engine import BatchedGraphMPE


class BaseScenario:  # multiagent/scenario.py:4-12
    def make_world(self, args):
        raise NotImplementedError()

    def reset_world(self, world, num_current_episode: int = 0):
        raise NotImplementedError()

    def info(self, agent, world):
        return {}


class _AgentState:
    def __init__(self, world, idx):
        self._w, self._i = world, idx

    @property
    def values(self):
        return self._w._host_state()[self._i]

    @property
    def p_pos(self):
        return self.values[:2]

    @property
    def p_vel(self):
        v = self.values
        if self._w.cfg.dynamics == abi.DYN_DOUBLE_INTEGRATOR:
            return v[2:4]
        return np.array([v[3] * np.cos(v[2]), v[3] * np.sin(v[2])])

    @property
    def theta(self):
        v = self.values
        return v[2] if self._w.cfg.dynamics != abi.DYN_DOUBLE_INTEGRATOR else np.arctan2(v[3], v[2])

    @property
    def speed(self):
        v = self.values
        return v[3] if self._w.cfg.dynamics != abi.DYN_DOUBLE_INTEGRATOR else float(np.hypot(v[2], v[3]))


class _Action:
    def __init__(self):
        self.u = None
        self.c = None


class Agent:
    """View of agent ``idx`` of a single-world engine (core.py:420-468)."""

    def __init__(self, world, idx):
        self._w = world
        self.id = idx
        self.global_id = idx
        self.name = f"agent {idx}"
        self.movable, self.collide, self.silent = True, True, True
        self.size = world.cfg.entity_size
        self.max_speed = world.cfg.max_speed
        self.state = _AgentState(world, idx)
        self.action = _Action()

    @property
    def status(self):
        return bool(self._w._host_flags()[self.id, abi.AF_STATUS])

    @property
    def previous_phase(self):
        return int(self._w._host_flags()[self.id, abi.AF_PREV_PHASE])


class _StaticEntity:
    def __init__(self, world, kind, idx, ent_index):
        self._w, self.id, self._e = world, idx, ent_index
        self.name = f"{kind} {idx}"
        self.movable = False
        self.collide = kind != "landmark"
        self.size = world.cfg.entity_size


class World:
    """Single world (B = 1 engine) with the attribute surface of ``multiagent.core.World``
    that ``MultiAgentBaseEnv.__init__`` and the runner read."""

    def __init__(self, args, cfg: abi.GmpeConfig, device=0):
        self.all_args = args
        self.cfg = cfg
        self.engine = BatchedGraphMPE(cfg, device)
        self.world_length = cfg.episode_length
        self.world_size = cfg.world_size
        self.world_aspect_ratio = 1.0
        self.with_background = False
        self.dim_p, self.dim_c = 2, 2
        self.total_actions = getattr(args, "total_actions", 5)
        self.dt = cfg.dt
        self.coordination_range = cfg.coordination_range
        self.collaborative = bool(cfg.flags & abi.FLAG_COLLABORATIVE)
        self.graph_mode = True
        self.graph_feat_type = getattr(args, "graph_feat_type", "relative")
        self.cache_dists = True
        self.current_time_step = 0
        self.agents = [Agent(self, i) for i in range(cfg.num_agents)]
        n = cfg.num_agents
        self.landmarks = [_StaticEntity(self, "landmark", i, n + i) for i in range(cfg.num_landmarks)]
        self.obstacles = [_StaticEntity(self, "obstacle", i, n + cfg.num_landmarks + i) for i in range(cfg.num_obstacles)]
        self.walls = [_StaticEntity(self, "wall", i, n + cfg.num_landmarks + cfg.num_obstacles + i) for i in range(cfg.num_walls)]
        self.scripted_agents = []
        self.edge_list = None
        self.edge_weight = None
        self._cache = {}
        self._host = None

    # -- reference attribute surface ------------------------------------------------
    @property
    def policy_agents(self):
        return self.agents

    @property
    def entities(self):
        return self.agents + self.landmarks + self.obstacles + self.walls

    @property
    def cached_dist_mag(self):
        return self._outputs()["adj"][0]

    @property
    def tube_params(self):
        wf = self.engine.state["world_f"][0].cpu().numpy().astype(np.float64)
        return {"entrance": wf[0:2], "exit": wf[2:4], "e": wf[4:6], "n": wf[6:8].astype(np.float32), "L": wf[8],
                "half_width": wf[9], "width": wf[10], "angle": wf[11]}

    # -- engine plumbing ---------------------------------------------------------------
    def _invalidate(self):
        self._cache = {}
        self._host = None

    def _outputs(self):
        if self._host is None:
            torch.cuda.synchronize(self.engine.device)
            self._host = {k: v.cpu().numpy() for k, v in self.engine.out.items()}
        return self._host

    def _host_state(self):
        if "as" not in self._cache:
            self._cache["as"] = self.engine.state["agent_state"][0].cpu().numpy().astype(np.float64)
        return self._cache["as"]

    def _host_flags(self):
        if "fl" not in self._cache:
            self._cache["fl"] = self.engine.state["agent_flags"][0].cpu().numpy()
        return self._cache["fl"]

    def step_actions(self, action_idx) -> None:
        """_set_action for every agent + World.step + every per-agent callback, one launch."""
        a = torch.as_tensor(np.asarray(action_idx, dtype=np.int32).reshape(1, -1), device=self.engine.device)
        self.engine.step(a)
        self.current_time_step += 1
        self._invalidate()

    def step(self):
        raise RuntimeError("World.step() is fused with the action decode and the scenario callbacks: "
                           "use MultiAgentGraphEnv.step(action_n) / World.step_actions(idx)")


class Scenario(BaseScenario):
    """Serves both scenario families of this build; the family is picked from
    ``args.scenario_name`` exactly like the reference picks a scenario file
    (multiagent/custom_scenarios/__init__.py:5-7)."""

    def __init__(self, device=0, precision="fp32"):
        self.device = device
        self.precision = precision

    def make_world(self, args) -> World:
        self.args = args
        self.num_agents = args.num_agents
        cfg = make_config(args, 1, precision=self.precision, auto_reset=False, gnn_edges=True)
        world = World(args, cfg, self.device)
        self.reset_world(world)
        return world

    def reset_world(self, world: World, num_current_episode: int = 0) -> None:
        world.engine.reset()
        world.current_time_step = 0
        world._invalidate()
        self.update_graph(world)

    # per-agent callbacks: O(1) views into the fused step's outputs -------------------
    def reward(self, agent: Agent, world: World) -> float:
        o = world._outputs()
        key = "indiv_reward" if "indiv_reward" in o else "reward"
        return float(o[key][0, agent.id])

    def observation(self, agent: Agent, world: World) -> np.ndarray:
        return world._outputs()["obs"][0, agent.id]

    def graph_observation(self, agent: Agent, world: World):
        o = world._outputs()
        return o["node_obs"][0, agent.id], o["adj"][0]  # one adjacency shared by all agents

    def update_graph(self, world: World) -> None:
        o = world._outputs()
        n = int(o["num_edges"][0])
        world.edge_list = o["edge_index"][0][:, :n].copy()
        world.edge_weight = o["edge_attr"][0][:n].copy()

    def get_id(self, agent: Agent) -> np.ndarray:
        return np.array([agent.global_id])

    def info_callback(self, agent: Agent, world: World) -> dict:
        o = world._outputs()
        row = o["info"][0, agent.id]
        d = {k: row[j] for j, k in enumerate(abi.INFO_KEYS)}
        if "phase_reached" in o:  # rot-inv family (three_phase_graph.py:806): same key order as the reference dict
            d = {k: (int(o["phase_reached"][0, agent.id]) if k == "Phase_reached" else d[k]) for k in abi.INFO_KEYS_ROT_INV}
        return d

    def done(self, agent: Agent, world: World) -> bool:
        return bool(world._outputs()["done"][0, agent.id])

    def get_agent_reached_goal(self, agent: Agent) -> int:
        return int(agent.status)


def load(name: str):
    """Plugin loader with the reference's calling convention:
    ``load(args.scenario_name + '.py').Scenario()`` (multiagent/MPE_env.py:63)."""
    from . import config

    base = name[:-3] if name.endswith(".py") else name
    if base not in config.JULY_SCENARIOS + config.NAV_SCENARIOS + config.THREE_PHASE_SCENARIOS + config.TWO_PHASE_SCENARIOS + config.CORRIDOR_ROT_INV_SCENARIOS:
        raise FileNotFoundError(f"no scenario {name!r} in this build")
    import sys
    return sys.modules[__name__]
