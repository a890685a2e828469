"""SURVEY.md section 8(f) rank 1 -- the rollout glue either side of the step path.
CPU: the numpy restatement (tests/rollout_ref.py) against the fixture produced by the
unmodified reference classes.  GPU: DeviceGraphRollout (step outputs written in place into
slot t+1 of device buffers) against the restatement fed with the same engine's outputs."""
import os

import numpy as np
import pytest

from rollout_ref import RolloutRef

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_glue_n3.npz")
FIELDS = ("obs", "share_obs", "node_obs", "adj", "agent_id", "share_agent_id", "rewards", "masks", "active_masks",
          "bad_masks", "available_actions")


def test_restatement_matches_the_reference_buffer_and_runner():
    z = np.load(GOLD)
    T = int(z["T"])
    _, B, N, F = z["in_obs"].shape
    E = z["in_adj_world"].shape[-1]
    A = z["available_actions"].shape[-1]
    ref = RolloutRef(T, B, N, E, F, A)
    ids = np.broadcast_to(np.arange(N).reshape(1, N, 1), (B, N, 1))
    adj = lambda t: np.broadcast_to(z["in_adj_world"][t][:, None], (B, N, E, E))
    ref.warmup(z["in_obs"][0], ids, z["in_node_obs"][0], adj(0))
    for t in range(T):
        ref.insert(z["in_obs"][t + 1], ids, z["in_node_obs"][t + 1], adj(t + 1), z["in_rewards"][t][:, :, None], z["in_dones"][t])
    for k in FIELDS:
        assert getattr(ref, k).dtype == z[k].dtype and np.array_equal(getattr(ref, k), z[k]), k
    ref.after_update()
    for k in ("obs", "node_obs", "adj", "masks", "active_masks"):
        assert np.array_equal(getattr(ref, k)[0], z["after_" + k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("scen", ["july", "nav_collab"])
def test_device_rollout_matches_the_restatement(scen):
    import torch

    from gmpe_b200 import default_args, make_config
    from gmpe_b200.engine import BatchedGraphMPE
    from gmpe_b200.rollout import DeviceGraphRollout

    if scen == "july":
        kw = dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", world_size=3,
                  episode_length=9, goal_rew=30, collaborative=False)
    else:
        kw = dict(scenario_name="navigation_graph", episode_length=6, collaborative=True, use_dones=True, total_actions=9)
    B, T = 777, 14  # rollout longer than an episode: auto-resets land inside the buffer
    cfg = make_config(default_args(**kw), B, precision="fp32", seed=3)
    eng, eng2 = BatchedGraphMPE(cfg, 0), BatchedGraphMPE(cfg, 0)
    N, E, F, A = eng.N, eng.E, eng.layout.obs_dim, cfg.total_actions
    roll = DeviceGraphRollout(eng, T)
    ref = RolloutRef(T, B, N, E, F, A)
    ids = np.broadcast_to(np.arange(N).reshape(1, N, 1), (B, N, 1))
    h = lambda o, k: o[k].cpu().numpy()
    wadj = lambda o: np.broadcast_to(h(o, "adj")[:, None], (B, N, E, E))
    for rnd in range(2):  # two rollouts: the second starts from after_update()
        if rnd == 0:
            roll.warmup()
            o = eng2.reset()
            ref.warmup(h(o, "obs"), ids, h(o, "node_obs"), wadj(o))
        g = torch.Generator().manual_seed(rnd)
        for t in range(T):
            a = torch.randint(0, A, (B, N), generator=g, dtype=torch.int32).cuda()
            roll.insert_env_step(a)
            o = eng2.step(a)
            rew = h(o, "reward")[:, :, None]
            ref.insert(h(o, "obs"), ids, h(o, "node_obs"), wadj(o), rew, h(o, "done").astype(bool))
        for k in FIELDS:
            got = getattr(roll, k).cpu().numpy()
            assert got.shape == getattr(ref, k).shape, (k, got.shape, getattr(ref, k).shape)
            assert np.array_equal(got, getattr(ref, k)), (rnd, k)
        assert (ref.masks == 0).any() and (ref.active_masks[1:] != ref.masks[1:]).any()  # both mask paths exercised
        roll.after_update()
        ref.after_update()
        for k in ("obs", "node_obs", "adj", "masks", "active_masks"):
            assert np.array_equal(getattr(roll, k)[0].cpu().numpy(), getattr(ref, k)[0]), k
    eng.close()
    eng2.close()
