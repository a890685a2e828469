"""The UNMODIFIED reference runner in the loop (SURVEY.md section 7 step 8 / 8b "drop-in test";
VERDICT r1 next-round #10): `GMPERunner.run()` (onpolicy/runner/shared/graph_mpe_runner.py:40-207) with its
own warmup / collect / collect_with_mask / get_finished / insert / compute, `Runner.train / save /
process_infos / log_*` (base_runner.py:150-302) and the unmodified `GraphReplayBuffer`
(onpolicy/utils/graph_buffer.py), stepping THIS repo's `GraphSubprocVecEnv` for a few episodes.

The reference exists only in the build container, which has no GPU; the GPU box has no reference.
So the façade runs here with its engine swapped for the oracle-backed stand-in
(tests/oracle_engine.py) and torch's pinned-memory / stream calls stubbed: what is proved is the
drop-in claim at the API level -- every shape, dtype, tuple arity, mask and `infos[e][a][key]`
access the runner makes -- not the kernels (tests/test_gpu_vec_env.py replays golden episodes
through the same façade on the GPU).  The policy is a stub with random logits (the real one needs
torch_geometric); the trainer only does what the rollout needs."""
import argparse
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle.ref_harness import REFERENCE_ROOT, install_shims, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not present (GPU box)")


class _Stream:
    def synchronize(self):
        pass

    cuda_stream = 0


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a, **k):
        pass

    def synchronize(self):
        pass


class StubPolicy:
    """R_MAPPOPolicy's call surface (graph_MAPPOPolicy.py): random logits, zero values."""

    def __init__(self, B, N, E, F, A, hidden, seed=0):
        self.B, self.N, self.E, self.F, self.A, self.hidden = B, N, E, F, A, hidden
        self.rs = np.random.RandomState(seed)
        self.actor, self.critic = torch.nn.Linear(1, 1), torch.nn.Linear(1, 1)
        self.calls = 0

    def get_actions(self, share_obs, obs, node_obs, adj, agent_id, share_agent_id, rnn_states, rnn_states_critic, masks,
                    available_actions=None, deterministic=False):
        n = self.B * self.N
        assert obs.shape == (n, self.F) and share_obs.shape == (n, self.N * self.F), (obs.shape, share_obs.shape)
        assert node_obs.shape == (n, self.E, 8) and adj.shape == (n, self.E, self.E), (node_obs.shape, adj.shape)
        assert agent_id.shape == (n, 1) and share_agent_id.shape == (n, self.N) and masks.shape == (n, 1)
        logits = self.rs.randn(n, self.A)
        if available_actions is not None:
            assert available_actions.shape == (n, self.A)
            logits = np.where(available_actions > 0, logits, -1e9)
        a = logits.argmax(-1)[:, None]
        self.calls += 1
        return (torch.zeros(n, 1), torch.as_tensor(a), torch.zeros(n, 1), torch.as_tensor(rnn_states), torch.as_tensor(rnn_states_critic))

    def get_values(self, share_obs, node_obs, adj, share_agent_id, rnn_states_critic, masks):
        return torch.zeros(self.B * self.N, 1)

    def lr_decay(self, *a):
        pass


class StubTrainer:
    value_normalizer = None

    def __init__(self, policy):
        self.policy = policy
        self.trained = 0

    def prep_rollout(self):
        pass

    def prep_training(self):
        pass

    def train(self, buffer, accumulation_steps=1, **kw):
        self.trained += 1
        self.last_rewards = buffer.rewards.copy()
        self.last_obs = buffer.obs.copy()
        self.last_actions = buffer.actions.copy()
        buffer.after_update()  # GR_MAPPO.train ends with it
        return {}


@pytest.mark.parametrize("scenario,dyn,A", [("nav_metered_one_goal_graph_rotate_tube_july", "air_taxi", 25),
                                            ("navigation_graph", "double_integrator", 5)])
def test_unmodified_gmpe_runner_runs_against_the_vec_env(monkeypatch, tmp_path, scenario, dyn, A):
    install_shims()
    for name, attrs in (("wandb", dict(log=lambda *a, **k: None)), ("imageio", {}),
                        ("tensorboardX", dict(SummaryWriter=type("SummaryWriter", (), {"__init__": lambda s, *a, **k: None})))):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, m)
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        from onpolicy.runner.shared.graph_mpe_runner import GMPERunner
        from onpolicy.utils.graph_buffer import GraphReplayBuffer
    finally:
        os.chdir(cwd)

    import oracle_engine
    from gmpe_b200 import default_args, env_wrappers

    monkeypatch.setattr(env_wrappers, "BatchedGraphMPE", oracle_engine.OracleEngine)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())

    B, T, episodes, hidden = 4, 6, 3, 8
    args = default_args(scenario_name=scenario, dynamics_type=dyn, world_size=3, episode_length=T, goal_rew=30,
                        collaborative=False, total_actions=5)
    envs = env_wrappers.GraphSubprocVecEnv(args=args, num_envs=B, device="cpu", seed=5)
    N, E, F = envs.num_agents, envs.engine.E, envs.engine.layout.obs_dim
    assert envs.action_space[0].n == A

    # the reference dispatches on space.__class__.__name__ (onpolicy/utils/util.py:31-52)
    for sp in ("observation_space", "share_observation_space", "node_observation_space", "adj_observation_space",
               "agent_id_observation_space", "share_agent_id_observation_space", "action_space"):
        assert getattr(envs, sp)[0].__class__.__name__ in ("Box", "Discrete"), sp

    all_args = argparse.Namespace(**vars(args))
    all_args.__dict__.update(episode_length=T, n_rollout_threads=B, hidden_size=hidden, recurrent_N=1, gamma=0.99, gae_lambda=0.95,
                             use_gae=True, use_popart=False, use_valuenorm=False, use_proper_time_limits=False,
                             use_centralized_V=True, increase_fairness=False, use_profiler=False, num_agents=N)
    runner = GMPERunner.__new__(GMPERunner)  # Runner.__init__ builds the GNN policy (torch_geometric): bypassed
    policy = StubPolicy(B, N, E, F, A, hidden)
    runner.__dict__.update(
        all_args=all_args, envs=envs, eval_envs=None, num_agents=N, n_rollout_threads=B, episode_length=T,
        num_env_steps=episodes * T * B, use_linear_lr_decay=False, use_centralized_V=True, recurrent_N=1, hidden_size=hidden,
        save_interval=1000, log_interval=1, eval_interval=1000, use_eval=False, use_wandb=True, save_dir=str(tmp_path),
        trainer=StubTrainer(policy), device=torch.device("cpu"),
        buffer=GraphReplayBuffer(all_args, N, envs.observation_space[0], envs.share_observation_space[0],
                                 envs.node_observation_space[0], envs.agent_id_observation_space[0],
                                 envs.share_agent_id_observation_space[0], envs.adj_observation_space[0], envs.action_space[0]))
    runner.run()

    assert runner.trainer.trained == episodes and policy.calls == episodes * T
    assert os.path.exists(os.path.join(str(tmp_path), "actor.pt"))
    # what the runner put into its buffer in the last episode is what the env produced for the
    # actions the runner chose: replay them through a fresh oracle batch
    from gmpe_b200 import make_config
    from oracle.oracle import OracleEnv

    eng = envs.engine
    assert len(eng.steps) == episodes * T
    orc = OracleEnv(make_config(args, B, precision="fp32", seed=5, auto_reset=True, info=True), use_rk45=False)
    orc.reset()
    for t, a in enumerate(eng.steps):
        o = orc.step(a)
        if t >= (episodes - 1) * T:
            k = t - (episodes - 1) * T
            assert np.array_equal(runner.trainer.last_actions[k][..., 0].astype(np.int32), a)
            assert np.allclose(runner.trainer.last_rewards[k][..., 0], o["reward"], rtol=0, atol=1e-5)
            assert np.allclose(runner.trainer.last_obs[k + 1], o["obs"], rtol=0, atol=1e-5)
