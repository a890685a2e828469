"""``MultiAgentGraphEnv`` with the reference's constructor and reset()/step() contract
(multiagent/environment.py:904-1091), driving a single-world (B = 1) engine.

Use this class where the reference is used one env at a time (the scenario ``__main__``
blocks, eval with ``n_rollout_threads == 1``).  Throughput runs go through
``env_wrappers.GraphSubprocVecEnv``, which keeps thousands of worlds in one batch.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np

from . import abi
from .config import num_discrete_actions
from .spaces import Box, Discrete


def build_spaces(args, cfg: abi.GmpeConfig):
    """Space declarations of MultiAgentBaseEnv.__init__ / set_graph_obs_space
    (environment.py:92-208, 986-1018)."""
    lay = abi.layout_for(cfg)
    N, E = cfg.num_agents, lay.num_entities
    A = num_discrete_actions(args)
    f32 = np.float32
    return dict(
        action_space=[Discrete(A) for _ in range(N)],
        observation_space=[Box(-np.inf, np.inf, (lay.obs_dim,), f32) for _ in range(N)],
        share_observation_space=[Box(-np.inf, np.inf, (lay.obs_dim * N,), f32) for _ in range(N)],
        node_observation_space=[Box(-np.inf, np.inf, (E, lay.node_dim), f32) for _ in range(N)],
        adj_observation_space=[Box(-np.inf, np.inf, (E, E), f32) for _ in range(N)],
        edge_observation_space=[Box(-np.inf, np.inf, (1,), f32) for _ in range(N)],
        agent_id_observation_space=[Box(-np.inf, np.inf, (1,), f32) for _ in range(N)],
        share_agent_id_observation_space=[Box(-np.inf, np.inf, (N,), f32) for _ in range(N)],
    )


def actions_to_index(action_n, num_actions: int, discrete_action_input: bool = False) -> np.ndarray:
    """Front half of _set_action (environment.py:336-475): a length-A one-hot per agent (or
    ``[int]`` when ``discrete_action_input``) -> action index.  Same error behaviour: the
    reference asserts that the whole action vector is consumed (environment.py:475)."""
    out = np.zeros(len(action_n), dtype=np.int32)
    for i, a in enumerate(action_n):
        a = np.asarray(a)
        if discrete_action_input:
            out[i] = int(a.reshape(-1)[0])
        else:
            assert a.ndim == 1 and a.shape[0] == num_actions, \
                f"action for agent {i} must be a one-hot of length {num_actions}, got shape {a.shape}"
            out[i] = int(np.argmax(a))
        if not (0 <= out[i] < num_actions):
            raise ValueError(f"action index {out[i]} out of range for Discrete({num_actions})")
    return out


class MultiAgentGraphEnv:
    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, world, reset_callback: Callable = None, reward_callback: Callable = None,
                 observation_callback: Callable = None, graph_observation_callback: Callable = None,
                 id_callback: Callable = None, info_callback: Callable = None, done_callback: Callable = None,
                 agent_reached_goal_callback: Callable = None, update_graph: Callable = None,
                 shared_viewer: bool = True, discrete_action: bool = True,
                 scenario_name: str = "navigation", dynamics_type: str = "airtaxi") -> None:
        self.world = world
        self.world_length = world.world_length
        self.world_aspect_ratio = world.world_aspect_ratio
        self.world_size = world.world_size
        self.with_background = world.with_background
        self.current_step = 0
        self.agents = world.policy_agents
        self.n = self.num_agents = len(world.policy_agents)
        self.reset_callback = reset_callback
        self.reward_callback = reward_callback
        self.observation_callback = observation_callback
        self.graph_observation_callback = graph_observation_callback
        self.id_callback = id_callback
        self.info_callback = info_callback
        self.done_callback = done_callback
        self.agent_reached_goal_callback = agent_reached_goal_callback
        self.update_graph = update_graph
        self.scenario_name = scenario_name
        self.discrete_action_space = discrete_action
        self.discrete_action_input = False
        self.shared_reward = world.collaborative
        self.dt = world.dt
        self.episode_length = world.world_length
        self.coordination_range = world.coordination_range
        self.num_discrete_action = num_discrete_actions(world.all_args)
        for k, v in build_spaces(world.all_args, world.cfg).items():
            setattr(self, k, v)
        self._seed = None

    def seed(self, seed=None):  # environment.py:219-223 (np.random there; Philox key here)
        self._seed = 1 if seed is None else int(seed)
        self.world.engine.set_seed(self._seed)
        return [self._seed]

    def reset(self, num_current_episode: Optional[int] = 0):  # environment.py:1066-1081
        self.current_step = 0
        self.reset_callback(self.world, num_current_episode)
        obs_n, node_obs_n, adj_n, agent_id_n = [], [], [], []
        self.agents = self.world.policy_agents
        for agent in self.agents:
            obs_n.append(self.observation_callback(agent, self.world))
            agent_id_n.append(self.id_callback(agent))
            node_obs, adj = self.graph_observation_callback(agent, self.world)
            node_obs_n.append(node_obs)
            adj_n.append(adj)
        return obs_n, agent_id_n, node_obs_n, adj_n

    def step(self, action_n: List):  # environment.py:1021-1063
        self.current_step += 1
        idx = actions_to_index(action_n, self.num_discrete_action, self.discrete_action_input)
        # environment.py:1022-1023: update_graph runs BEFORE _set_action / World.step, on the masked
        # distances of the previous step -- world.edge_list after step() describes the pre-step graph
        if self.update_graph is not None:
            self.update_graph(self.world)
        self.world.step_actions(idx)  # _set_action + World.step + all scenario callbacks, fused
        obs_n, reward_n, done_n, info_n = [], [], [], []
        node_obs_n, adj_n, agent_id_n = [], [], []
        for agent in self.agents:
            obs_n.append(self.observation_callback(agent, self.world))
            agent_id_n.append(self.id_callback(agent))
            reward = self.reward_callback(agent, self.world)
            reward_n.append(reward)
            node_obs, adj = self.graph_observation_callback(agent, self.world)
            node_obs_n.append(node_obs)
            adj_n.append(adj)
            done_n.append(bool(self.world._outputs()["done"][0, agent.id]))
            info = {"individual_reward": reward}
            if self.info_callback is not None:
                info.update(self.info_callback(agent, self.world))
            info_n.append(info)
        if self.shared_reward:
            reward = float(self.world._outputs()["reward"][0, 0])
            reward_n = [[reward]] * self.n
        return obs_n, agent_id_n, node_obs_n, adj_n, reward_n, done_n, info_n

    def render(self, mode: str = "human", close: bool = False):
        raise NotImplementedError("rendering (pyglet) is out of scope of this build")

    def close(self):
        self.world.engine.close()


def GraphMPEEnv(args, device=0, precision="fp32") -> MultiAgentGraphEnv:
    """Factory with the reference's signature (multiagent/MPE_env.py:56-84)."""
    from .scenario import Scenario

    assert "graph" in args.scenario_name, "Only use graph env for graph scenarios"
    scenario = Scenario(device=device, precision=precision)
    world = scenario.make_world(args=args)
    return MultiAgentGraphEnv(
        world=world, reset_callback=scenario.reset_world, reward_callback=scenario.reward,
        observation_callback=scenario.observation, graph_observation_callback=scenario.graph_observation,
        update_graph=scenario.update_graph, id_callback=scenario.get_id, info_callback=scenario.info_callback,
        done_callback=scenario.done, agent_reached_goal_callback=scenario.get_agent_reached_goal,
        scenario_name=args.scenario_name, discrete_action=getattr(args, "discrete_action", True),
        dynamics_type=args.dynamics_type)
