#!/usr/bin/env python
"""Golden fixture for the rollout glue, made by running the
UNMODIFIED reference code in the build container:

  * ``GraphReplayBuffer`` (onpolicy/utils/graph_buffer.py:45-287) is imported as is;
  * ``GMPERunner.warmup`` / ``GMPERunner.insert`` (onpolicy/runner/shared/graph_mpe_runner.py:
    213-239, 384-427) are executed from their own source text against a stand-in ``self``
    (the real runner needs the whole trainer).

Inputs are a seeded synthetic env-output stream (shapes of the july scenario, N = 3); the
fixture stores the inputs and the env-produced buffer fields after T steps and
``after_update``.  Writes tests/golden/rollout_glue_n3.npz."""
import argparse
import ast
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_harness import REFERENCE_ROOT, install_shims  # noqa: E402


def runner_methods(*names):
    src = open(os.path.join(REFERENCE_ROOT, "onpolicy/runner/shared/graph_mpe_runner.py")).read()
    out = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []
            ns = {"np": np}
            exec(textwrap.dedent(ast.unparse(node)), ns)
            out[node.name] = ns[node.name]
    return out


def main():
    install_shims()
    from gym import spaces
    from onpolicy.utils.graph_buffer import GraphReplayBuffer

    B, N, E, F, A, T = 5, 3, 6, 19, 25, 7
    args = argparse.Namespace(episode_length=T, n_rollout_threads=B, hidden_size=8, recurrent_N=1, gamma=0.99,
                              gae_lambda=0.95, use_gae=True, use_popart=False, use_valuenorm=True,
                              use_proper_time_limits=False, use_centralized_V=True)
    # the reference dispatches on space.__class__.__name__ (onpolicy/utils/util.py:31-52); the
    # harness stubs of gym.spaces are private classes, so give them the real names
    Box = type("Box", (spaces.Box,), {})
    Discrete = type("Discrete", (spaces.Discrete,), {})
    box = lambda *s: Box(low=-np.inf, high=np.inf, shape=s, dtype=np.float32)
    buf = GraphReplayBuffer(args, N, box(F), box(N * F), box(E, 8), box(1), box(N), box(E, E), Discrete(A))
    m = runner_methods("warmup", "insert")
    rs = np.random.RandomState(3)
    stream = dict(obs=rs.randn(T + 1, B, N, F), node_obs=rs.randn(T + 1, B, N, E, 8), adj_world=np.abs(rs.randn(T + 1, B, E, E)),
                  rewards=rs.randn(T, B, N), dones=rs.rand(T, B, N) < 0.35)
    stream["dones"][3, 1] = True   # a world whose agents are all done (active_masks back to 1)
    stream["dones"][5, 4] = True
    ids = np.broadcast_to(np.arange(N).reshape(1, N, 1), (B, N, 1))

    class Envs:
        def reset(self_):
            return stream["obs"][0], ids, stream["node_obs"][0], np.broadcast_to(stream["adj_world"][0][:, None], (B, N, E, E))

    me = types.SimpleNamespace(envs=Envs(), buffer=buf, use_centralized_V=True, n_rollout_threads=B, num_agents=N,
                               recurrent_N=1, hidden_size=8)
    m["warmup"](me)
    for t in range(T):
        dones = stream["dones"][t]
        adj = np.broadcast_to(stream["adj_world"][t + 1][:, None], (B, N, E, E))
        data = (stream["obs"][t + 1], ids, stream["node_obs"][t + 1], adj, ids, stream["rewards"][t][:, :, None], dones,
                None, np.zeros((B, N, 1), np.float32), np.zeros((B, N, 1), np.float32), np.zeros((B, N, 1), np.float32),
                np.zeros((B, N, 1, 8), np.float32), np.zeros((B, N, 1, 8), np.float32), np.ones((B, N, A), np.float32))
        m["insert"](me, data)
    before = {k: getattr(buf, k).copy() for k in ("obs", "share_obs", "node_obs", "adj", "agent_id", "share_agent_id",
                                                   "rewards", "masks", "active_masks", "bad_masks", "available_actions")}
    buf.after_update()
    after = {"after_" + k: getattr(buf, k)[0].copy() for k in ("obs", "node_obs", "adj", "masks", "active_masks")}
    path = os.path.join(HERE, "rollout_glue_n3.npz")
    np.savez_compressed(path, **{"in_" + k: v for k, v in stream.items()}, **before, **after, T=T)
    print(path, os.path.getsize(path) // 1024, "KB", "masks zeros:", int((before["masks"] == 0).sum()),
          "active zeros:", int((before["active_masks"] == 0).sum()))


if __name__ == "__main__":
    main()


def make_infos_golden():
    """Runner.process_infos + log_env (onpolicy/runner/shared/base_runner.py:194-302) executed
    from their own source text on a seeded synthetic info stream -> tests/golden/process_infos_n3.npz."""
    import json

    src = open(os.path.join(REFERENCE_ROOT, "onpolicy/runner/shared/base_runner.py")).read()
    fn = None
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == "process_infos":
            ns = {"np": np}
            exec(textwrap.dedent(ast.unparse(node)), ns)
            fn = ns["process_infos"]
    from gmpe_b200 import abi

    B, N, T, dt = 11, 3, 200, 1.0
    rs = np.random.RandomState(9)
    info = rs.rand(B, N, len(abi.INFO_KEYS)) * 10
    j = abi.INFO_KEYS.index("Time_req_to_goal")
    info[..., j] = np.where(rs.rand(B, N) < 0.4, -1.0, np.floor(info[..., j]))
    rew = rs.randn(B, N)
    infos = [[dict({"individual_reward": rew[e, a]}, **{k: info[e, a, c] for c, k in enumerate(abi.INFO_KEYS)}) for a in range(N)]
             for e in range(B)]
    me = types.SimpleNamespace(num_agents=N, all_args=types.SimpleNamespace(episode_length=T), dt=dt)
    env_infos = fn(me, infos)
    logged = {k: float(np.mean(v)) for k, v in env_infos.items() if len(v) > 0}  # log_env, base_runner.py:298-302
    path = os.path.join(HERE, "process_infos_n3.npz")
    np.savez_compressed(path, info=info, indiv_reward=rew, episode_length=T, dt=dt, logged_json=json.dumps(logged))
    print(path, len(logged), "logged keys")


if __name__ == "__main__":
    make_infos_golden()
