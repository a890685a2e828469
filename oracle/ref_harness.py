"""TEST INFRASTRUCTURE ONLY -- import shims for running the *unmodified* Python
reference (`/root/reference`) inside this container.

Nothing here is product code and nothing here travels to the GPU box:
`/root/reference` exists only in the build container.  The only consumers are
`tests/golden/make_golden.py` (which writes the committed golden traces) and the
`-m "not gpu"` tests that re-validate the C oracle against the live reference
when it happens to be present.

Shim list follows SURVEY.md section 8(c): every stub is import-time only and touches
no arithmetic.  Harness patches (each documented where applied):
  P1  build `MultiAgentGraphEnv` directly (the `GraphMPEEnv` factory passes a
      callback the july scenario does not define, `multiagent/MPE_env.py:79`).
  P2  seed `np.random` *before* `make_world` and retry when the first unseeded
      spawn lands inside the tube (`july:714` reads `agent.previous_phase`
      before it exists).
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GMPE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "multiagent"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Space:
    pass


class _Discrete(_Space):
    def __init__(self, n):
        self.n = int(n)

    def __repr__(self):
        return f"Discrete({self.n})"


class _Box(_Space):
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f"Box{self.shape}"


class _Tuple(_Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)


class _Env:
    def close(self):
        pass


class _EnvSpec:
    def __init__(self, *a, **k):
        pass


_installed = False


def install_shims() -> None:
    """Insert stub modules for everything the reference imports but this image
    lacks (gym, pyglet, jax, hj_reachability, cvxpy, casadi, imp)."""
    global _installed
    if _installed:
        return
    _installed = True

    spaces = _mod("gym.spaces", Discrete=_Discrete, Box=_Box, Tuple=_Tuple, Space=_Space)
    registration = _mod("gym.envs.registration", register=lambda *a, **k: None, EnvSpec=_EnvSpec)
    envs = _mod("gym.envs", registration=registration)
    gym = _mod("gym", Env=_Env, Space=_Space, spaces=spaces, envs=envs)
    gym.__path__ = []  # mark as package
    envs.__path__ = []

    pyglet = _mod("pyglet", image=types.SimpleNamespace(), window=types.SimpleNamespace())
    pyglet.__path__ = []
    _mod("pyglet.window")
    _mod("pyglet.image")

    jax = _mod("jax")
    jax.__path__ = []
    jax.numpy = _mod("jax.numpy")
    hj = _mod("hj_reachability")
    hj.__path__ = []
    hj.sets = types.SimpleNamespace(Box=lambda **k: None)
    hju = _mod("hj_reachability_utils")
    hju.__path__ = []
    _mod(
        "hj_reachability_utils.common",
        get_hj_grid_from_meta_data=lambda *a, **k: None,
        ControlAndDisturbanceAffineDynamics=object,
    )
    _mod("cvxpy")
    casadi = _mod("casadi")
    casadi.__all__ = []

    def load_source(name, path):
        loader = importlib.machinery.SourceFileLoader(name or "_gmpe_ref_scenario", path)
        spec = importlib.util.spec_from_loader(loader.name, loader)
        module = importlib.util.module_from_spec(spec)
        loader.exec_module(module)
        return module

    _mod("imp", load_source=load_source)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def july_args(**over):
    """The Namespace the reference's CLI would hand to `Scenario.make_world` for
    BASELINE config C3 (README.md:84-105 eval command + onpolicy/config.py defaults)."""
    import argparse

    d = dict(
        scenario_name="nav_metered_one_goal_graph_rotate_tube_july",
        dynamics_type="air_taxi",
        num_agents=3,
        num_landmarks=3,
        num_obstacles=0,
        num_walls=0,
        num_scripted_agents=0,
        world_size=3,
        episode_length=200,
        collaborative=False,
        use_dones=False,
        max_speed=2,
        collision_rew=5,
        goal_rew=30,
        formation_rew=1,
        fair_wt=1,
        fair_rew=1,
        formation_type="point",
        total_actions=5,
        zeroshift=5,
        min_obs_dist=0.5,
        num_env_steps=5_000_000,
        n_rollout_threads=1,
        render_episodes=None,
        graph_feat_type="relative",
        max_edge_dist=1,
        discrete_action=True,
        use_safety_filter=False,
        seed=1,
    )
    d.update(over)
    return argparse.Namespace(**d)


def build_july_env(args, seed: int):
    """P1 + P2: returns (env, scenario) for the july scenario, np.random seeded
    with `seed` before `make_world`."""
    install_shims()
    import numpy as np

    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        from multiagent.custom_scenarios import load
        from multiagent.environment import MultiAgentGraphEnv
    finally:
        os.chdir(cwd)

    attempt = 0
    while True:
        np.random.seed(seed + 7919 * attempt)
        scenario = load(args.scenario_name + ".py").Scenario()
        try:
            world = scenario.make_world(args)
            env = MultiAgentGraphEnv(
                world=world,
                reset_callback=scenario.reset_world,
                reward_callback=scenario.reward,
                observation_callback=scenario.observation,
                graph_observation_callback=scenario.graph_observation,
                update_graph=scenario.update_graph,
                id_callback=scenario.get_id,
                info_callback=scenario.info_callback,
                done_callback=scenario.done,
                scenario_name=args.scenario_name,
                discrete_action=args.discrete_action,
                dynamics_type=args.dynamics_type,
            )
            return env, scenario
        except AttributeError as e:  # G13: previous_phase read before creation
            if "previous_phase" not in str(e):
                raise
            attempt += 1
            if attempt > 50:
                raise
