"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the
oracle takes minutes at these sizes): graph outputs are mutually consistent, the batch is
invariant to how it is sharded, and a sampled subset of worlds matches the CPU oracle."""
import numpy as np
import pytest
import torch

from gmpe_b200 import abi, default_args, make_config
from oracle.oracle import OracleEnv

pytestmark = pytest.mark.gpu

NAV = dict(scenario_name="navigation_graph", num_agents=3, num_landmarks=3, episode_length=25, collaborative=True)
JULY = dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", num_agents=3,
            num_landmarks=3, world_size=3, episode_length=200, goal_rew=30, collaborative=False)


def _run(args, B, T, seed=5, world_offset=0, precision="fp32"):
    from gmpe_b200.engine import BatchedGraphMPE

    cfg = make_config(default_args(**args), B, precision=precision, seed=seed, world_offset=world_offset, gnn_edges=True)
    eng = BatchedGraphMPE(cfg, 0)
    eng.reset()
    return cfg, eng


def _actions(cfg, t, B0, B):
    # world b always receives the same action stream, whatever batch it is part of
    g = torch.Generator().manual_seed(1000 + t)
    a = torch.randint(0, cfg.total_actions, (1 << 17, cfg.num_agents), generator=g, dtype=torch.int32)
    return a[B0:B0 + B].cuda()


@pytest.mark.parametrize("name,args,B", [("config1_nav_65536", NAV, 65536), ("config2_july_131072", JULY, 131072)])
def test_graph_outputs_are_mutually_consistent_at_full_size(name, args, B):
    cfg, eng = _run(args, B, 0)
    N, E = cfg.num_agents, eng.E
    cap = E * (E - 1)
    for t in range(30):  # crosses an auto-reset of every navigation world (episode_length 25)
        out = eng.step(_actions(cfg, t, 0, B))
    adj, node, obs = out["adj"], out["node_obs"], out["obs"]
    st = eng.state["agent_state"]
    # adjacency: symmetric, zero diagonal, non-negative
    assert torch.equal(adj, adj.transpose(1, 2))
    assert (torch.diagonal(adj, dim1=1, dim2=2) == 0).all() and (adj >= 0).all()
    # node features: rel_pos is antisymmetric between agents and matches the state
    rel = node[:, :, :N, 2:4]
    assert torch.equal(rel, -rel.transpose(1, 2))
    pos = st[..., :2]
    assert torch.equal(rel, pos[:, None, :, :] - pos[:, :, None, :])
    # adjacency entries are the distances of those relative positions wherever unmasked
    d = torch.sqrt(rel[..., 0] ** 2 + rel[..., 1] ** 2)
    aa = adj[:, :N, :N]
    live = aa > 0
    assert torch.allclose(aa[live], d[live], rtol=2e-6, atol=0)
    # env-side edge list == row-major nonzero of (0 < adj <= coordination_range), -1 / 0 padded
    on = (adj > 0) & (adj <= float(np.float32(cfg.coordination_range)))
    ne = on.flatten(1).sum(1).to(torch.int32)
    assert torch.equal(ne, out["num_edges"])
    ei, ea = out["edge_index"], out["edge_attr"]
    slot = torch.arange(cap, device="cuda")[None, :]
    pad = slot >= ne[:, None]
    assert (ei[:, 0][pad] == -1).all() and (ei[:, 1][pad] == -1).all() and (ea[pad] == 0).all()
    r, c = ei[:, 0].clamp(min=0).long(), ei[:, 1].clamp(min=0).long()
    got = adj.flatten(1).gather(1, r * E + c)
    assert torch.equal(got[~pad], ea[~pad])
    key = (r * E + c).masked_fill(pad, E * E)
    assert (key[:, 1:] > key[:, :-1])[~pad[:, 1:]].all()  # strictly row-major
    # consumer-side list (process_adj): strict < max_edge_dist on the fp32 values
    gon = (adj > 0) & (adj < float(np.float32(cfg.gnn_max_edge_dist)))
    assert torch.equal(gon.flatten(1).sum(1).to(torch.int32), out["gnn_num_edges"])
    # done flags follow _get_done; observations carry the state
    step = eng.state["world_i"][:, abi.WI_STEP]
    status = eng.state["agent_flags"][..., abi.AF_STATUS]
    assert ((out["done"] != 0) | (step[:, None] < cfg.episode_length)).all()
    assert (status <= 1).all()
    if cfg.scenario == abi.SCEN_NAV_GRAPH:
        fresh = step == 0  # worlds that auto-reset this step show the new episode's observation
        assert torch.equal(obs[~fresh][..., 2:4], pos[~fresh])
    else:
        assert torch.equal(obs[..., 0:2], pos)
    eng.close()


@pytest.mark.parametrize("args", [NAV, JULY], ids=["nav", "july"])
def test_results_do_not_depend_on_the_sharding(args):
    """World b draws from Philox(seed, world_offset + b): a shard of the batch, run alone with
    its offset (what rank k of a multi-GPU job does), reproduces the same worlds bit for bit."""
    B, B0, Bs, T = 65536, 40960, 4096, 28
    cfg, big = _run(args, B, T)
    cfg_s, shard = _run(args, Bs, T, world_offset=B0)
    for t in range(T):
        ob = big.step(_actions(cfg, t, 0, B))
        osh = shard.step(_actions(cfg, t, B0, Bs))
    for k in ("obs", "node_obs", "adj", "reward", "done", "info", "edge_index", "edge_attr", "num_edges"):
        a, b = ob[k][B0:B0 + Bs], osh[k]
        assert torch.equal(a, b) or torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), k
    for k in ("agent_state", "agent_flags", "agent_cnt", "world_f", "landmarks"):
        assert torch.equal(big.state[k][B0:B0 + Bs], shard.state[k]), k
    big.close()
    shard.close()


WIDE = dict(scenario_name="nav_graph_metered_single_corridor_rot_inv", dynamics_type="air_taxi", num_agents=16, num_landmarks=16,
            world_size=6, episode_length=6, goal_rew=30, collaborative=False)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_wide_rot_inv_worlds_do_not_depend_on_the_sharding(precision):
    """The wide lane-per-agent instantiations (N = 16: two worlds per warp, node rows staged one world at a time):
    an odd batch (partial last group, transposed tail) and a shard of it agree bit for bit through auto-resets."""
    B, B0, Bs, T = 2051, 1025, 513, 14
    cfg, big = _run(WIDE, B, T, precision=precision)
    cfg_s, shard = _run(WIDE, Bs, T, world_offset=B0, precision=precision)
    for t in range(T):
        ob = big.step(_actions(cfg, t, 0, B))
        osh = shard.step(_actions(cfg, t, B0, Bs))
        for k in ("obs", "node_obs", "adj", "reward", "done", "edge_index", "edge_attr", "num_edges", "gnn_edge_index", "gnn_num_edges"):
            a, b = ob[k][B0:B0 + Bs], osh[k]
            assert torch.equal(a, b) or torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), (t, k)
    for k in ("agent_state", "agent_flags", "agent_cnt", "agent_ext", "world_f", "landmarks"):
        assert torch.equal(big.state[k][B0:B0 + Bs], shard.state[k]), k
    big.close()
    shard.close()


def test_sampled_worlds_of_the_full_batch_match_the_oracle():
    """fp64 instantiation at 65,536 worlds, free-running with auto-resets; 512 worlds spread
    over the batch are replayed by the CPU oracle with their global ids and compared exactly
    (discrete) / to 1e-9 (reals)."""
    B, T = 65536, 27
    cfg, eng = _run(NAV, B, T, precision="fp64")
    pick = np.arange(0, B, 128)
    outs = []
    for t in range(T):
        o = eng.step(_actions(cfg, t, 0, B))
        outs.append({k: o[k][pick].cpu().numpy() for k in ("obs", "node_obs", "adj", "reward", "done", "edge_index", "num_edges")})
    for j in range(0, len(pick), 64):  # the oracle replays each picked world as a 1-world batch with its id
        for b in pick[j:j + 2]:
            c1 = make_config(default_args(**NAV), 1, precision="fp64", seed=5, world_offset=int(b), gnn_edges=True)
            orc = OracleEnv(c1, use_rk45=False)
            orc.reset()
            i = int(np.where(pick == b)[0][0])
            for t in range(T):
                a = _actions(cfg, t, int(b), 1).cpu().numpy()
                oo = orc.step(a)
                for k in ("obs", "node_obs", "adj", "reward"):
                    assert np.abs(outs[t][k][i] - oo[k][0]).max() <= 1e-9, (b, t, k)
                for k in ("done", "edge_index", "num_edges"):
                    assert (outs[t][k][i] == oo[k][0]).all(), (b, t, k)
    eng.close()
