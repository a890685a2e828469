"""GPU tests of the reference-facing plugin surfaces: GraphSubprocVecEnv / GraphDummyVecEnv
(onpolicy/envs/env_wrappers.py:903-1037) and MultiAgentGraphEnv (multiagent/environment.py:904)."""
import numpy as np
import pytest
import torch

from golden_util import Trace, traces
from gmpe_b200 import abi, default_args, make_config
from oracle.oracle import OracleEnv

pytestmark = pytest.mark.gpu

JULY = dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", world_size=3,
            episode_length=6, goal_rew=30, collaborative=False, seed=3)


def test_subproc_vec_env_reference_calling_convention_and_shapes():
    from gmpe_b200.env_wrappers import GraphDummyVecEnv, GraphSubprocVecEnv
    from gmpe_b200.environment import GraphMPEEnv

    args = default_args(**JULY)

    def get_env_fn(rank):  # onpolicy/scripts/train_mpe.py:22-33
        def init_env():
            env = GraphMPEEnv(args)
            env.seed(args.seed + rank * 1000)
            return env
        return init_env

    B, N, E, A = 5, 3, 6, 25
    envs = GraphSubprocVecEnv([get_env_fn(i) for i in range(B)])
    assert envs.num_envs == B and envs.action_space[0].n == A
    assert envs.node_observation_space[0].shape == (E, 8) and envs.adj_observation_space[0].shape == (E, E)
    obs, ag, node, adj = envs.reset()
    assert obs.shape == (B, N, 19) and ag.shape == (B, N, 1) and node.shape == (B, N, E, 8) and adj.shape == (B, N, E, E)
    assert (ag[:, :, 0] == np.arange(N)).all()
    acts = np.eye(A)[np.random.RandomState(0).randint(A, size=(B, N))]  # graph_mpe_runner.py:375-377
    out = envs.step(acts, 0)
    assert len(out) == 7
    obs, ag, node, adj, rews, dones, infos = out
    assert rews.shape == (B, N) and dones.shape == (B, N) and dones.dtype == np.bool_
    assert len(infos) == B and len(infos[0]) == N and "Dist_to_goal" in infos[0][0].keys()
    assert infos[B - 1][N - 1]["individual_reward"] == rews[B - 1, N - 1]
    assert (adj[:, 0] == adj[:, 1]).all() and (adj[:, 0] == adj[:, 2]).all()  # one matrix per world (G12)
    with pytest.raises(ValueError):
        envs.step(np.zeros((B, N, A + 1)), 0)
    envs.close()

    dummy = GraphDummyVecEnv([get_env_fn(0)])
    dummy.reset()
    o8 = dummy.step(np.eye(A)[np.zeros((1, N), dtype=int)], 0)
    assert len(o8) == 8 and o8[7] in (0, 1)  # reset_count appended
    dummy.close()


@pytest.mark.parametrize("kw", [JULY, dict(scenario_name="navigation_graph", episode_length=5, num_obstacles=2, seed=9)])
def test_vec_env_matches_oracle_through_host_one_hot_actions_with_auto_reset(kw):
    from gmpe_b200.env_wrappers import GraphSubprocVecEnv

    args = default_args(**kw)
    B, T = 40, 14
    envs = GraphSubprocVecEnv(args=args, num_envs=B, precision="fp64", seed=args.seed, env_edges=True)
    cfg = envs.cfg
    orc = OracleEnv(cfg, use_rk45=False)
    N, A, E = cfg.num_agents, cfg.total_actions, envs.engine.E
    o = orc.reset()
    obs, ag, node, adj = envs.reset()
    np.testing.assert_allclose(obs, o["obs"], atol=1e-9)
    np.testing.assert_allclose(node, o["node_obs"], atol=1e-9)
    rs = np.random.RandomState(1)
    saw_reset = False
    for t in range(T):
        idx = rs.randint(A, size=(B, N))
        o = orc.step(idx.astype(np.int32))
        obs, ag, node, adj, rews, dones, infos = envs.step(np.eye(A)[idx], t)
        np.testing.assert_allclose(obs, o["obs"], atol=1e-9)
        np.testing.assert_allclose(node, o["node_obs"], atol=1e-9)
        np.testing.assert_allclose(adj, np.broadcast_to(o["adj"][:, None], (B, N, E, E)), atol=1e-9)
        want_r = o["reward"][..., None] if cfg.flags & abi.FLAG_COLLABORATIVE else o["reward"]
        np.testing.assert_allclose(rews, want_r, atol=1e-9)
        assert (dones == o["done"].astype(bool)).all()
        np.testing.assert_allclose(np.array([[infos[e][a]["Min_time_to_goal"] for a in range(N)] for e in range(3)]),
                                   o["info"][:3, :, abi.INFO_KEYS.index("Min_time_to_goal")], atol=1e-9)
        if dones.all(axis=1).any():
            saw_reset = True
            e = int(np.argmax(dones.all(axis=1)))
            # the observation of a finished env is the FIRST observation of its next episode
            # (env_wrappers.py:867-873) while reward/done stay terminal
            st = envs.engine.state["agent_state"][e].cpu().numpy()
            col = (0, 1) if cfg.scenario == abi.SCEN_JULY_TUBE else (2, 3)
            np.testing.assert_allclose(obs[e][:, col], st[:, :2], atol=1e-12)
            assert int(envs.engine.state["world_i"][e, abi.WI_STEP]) == 0
    assert saw_reset
    envs.close()


def test_vec_env_device_output_mode_is_zero_copy():
    from gmpe_b200.env_wrappers import GraphSubprocVecEnv

    args = default_args(**JULY)
    envs = GraphSubprocVecEnv(args=args, num_envs=16, output="torch")
    obs, ag, node, adj = envs.reset()
    assert obs.is_cuda and node.is_cuda and adj.shape == (16, 3, 6, 6) and adj.stride(1) == 0
    idx = torch.zeros((16, 3), dtype=torch.int32, device="cuda")
    obs, ag, node, adj, rews, dones, infos = envs.step(idx, 0)
    assert obs.data_ptr() == envs.engine.out["obs"].data_ptr() and dones.dtype == torch.bool
    assert envs.h2d_bytes_last == 0
    st = envs.episode_stats()
    assert st.shape == (5,) and float(st[4]) == 48.0
    envs.close()


@pytest.mark.parametrize("which", ["july_n3_pursuit_s1", "three_phase_n3_pursuit_s21", "two_phase_n3_pursuit_s31"])
def test_multi_agent_graph_env_replays_a_golden_episode_through_the_scenario_api(which):
    from gmpe_b200.environment import MultiAgentGraphEnv
    from gmpe_b200.scenario import Scenario

    tr = Trace([p for p in traces("*.npz") if which in p][0])
    F, ND, tol = (19, 8, 2e-7) if which.startswith("july") else (15, 7, 1e-6)
    scenario = Scenario(precision="fp64")
    world = scenario.make_world(tr.args)
    env = MultiAgentGraphEnv(world=world, reset_callback=scenario.reset_world, reward_callback=scenario.reward,
                             observation_callback=scenario.observation,
                             graph_observation_callback=scenario.graph_observation, update_graph=scenario.update_graph,
                             id_callback=scenario.get_id, info_callback=scenario.info_callback,
                             done_callback=scenario.done, scenario_name=tr.args.scenario_name,
                             dynamics_type=tr.args.dynamics_type)
    assert env.n == 3 and env.action_space[0].n == 25 and env.observation_space[0].shape == (F,)
    obs_n, id_n, node_n, adj_n = env.reset()
    assert len(obs_n) == 3 and node_n[0].shape == (6, ND) and adj_n[0].shape == (6, 6)
    z = tr.z
    world.engine.reset(reset_state=tr.reset_state(0))  # teacher-forced initial condition of the trace
    world.engine.state["agent_flags"][0, :, abi.AF_PREV_PHASE] = torch.as_tensor(z["init_prev_phase"][0], dtype=torch.uint8)
    world._invalidate()
    for t in range(60):
        if not np.isnan(z["theta_draw"][t]).all():
            break  # the B=1 env draws its own headings; stop before the first goal reach
        obs_n, id_n, node_n, adj_n, rew_n, done_n, info_n = env.step([np.eye(25)[a] for a in z["actions"][t]])
        np.testing.assert_allclose(np.array(obs_n), z["obs"][t], atol=tol)
        np.testing.assert_allclose(np.array(node_n), z["node_obs"][t], atol=tol)
        np.testing.assert_allclose(adj_n[0], z["adj"][t], atol=2e-7)
        np.testing.assert_allclose(np.array(rew_n), z["reward"][t], atol=tol)
        assert done_n == list(z["done"][t])
        assert info_n[1]["individual_reward"] == rew_n[1] and "Num_agent_collisions" in info_n[0]
        if which.startswith("july"):
            assert "Phase_reached" not in info_n[0]
        else:  # the rot-inv family's extra info entry, in the reference's key order (three_phase_graph.py:788-811)
            assert [d["Phase_reached"] for d in info_n] == [int(x) for x in z["phase_reached"][t]]
            assert list(info_n[0])[-2:] == ["Phase_reached", "Min_time_to_goal"]
        # update_graph runs at the top of step() on the previous step's masked distances
        # (environment.py:1022-1023); the trace stores the list of step t under index t
        if t > 0:
            assert (world.edge_list == z["edge_index"][t - 1][:, :z["num_edges"][t - 1]]).all()
        assert [int(i[0]) for i in id_n] == [0, 1, 2]
    assert t > 10
    with pytest.raises(AssertionError):
        env.step([np.eye(24)[0]] * 3)
    env.close()


@pytest.mark.parametrize("scen", ["nav", "july"])
def test_sliced_pipelined_step_equals_the_whole_batch_step(scen):
    """gmpe_step_range / gmpe_decode_onehot_range: stepping the batch in slices on several
    streams (what the host-facing vec env does to overlap PCIe copies) is bit-identical to one
    whole-batch launch, through auto-resets."""
    from gmpe_b200 import default_args
    from gmpe_b200.env_wrappers import GraphSubprocVecEnv

    if scen == "nav":
        args = default_args(scenario_name="navigation_graph", episode_length=5, collaborative=True)
    else:
        args = default_args(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi",
                            world_size=3, episode_length=6, goal_rew=30, collaborative=False)
    B = 9001  # not a multiple of the slice size
    a = GraphSubprocVecEnv(args=args, num_envs=B, device=0, seed=4, pipeline_chunks=1)
    b = GraphSubprocVecEnv(args=args, num_envs=B, device=0, seed=4, pipeline_chunks=5)
    assert b._chunks > 1
    ra, rb = a.reset(), b.reset()
    for x, y in zip(ra, rb):
        assert np.array_equal(x, y)
    A = a.cfg.total_actions
    rs = np.random.RandomState(1)
    for t in range(9):
        idx = rs.randint(A, size=(B, a.num_agents))
        act = np.eye(A)[idx] if t % 2 == 0 else idx.astype(np.int64)
        oa, ob = a.step(act, 0), b.step(act, 0)
        for k in range(6):
            assert np.array_equal(np.asarray(oa[k]), np.asarray(ob[k]), equal_nan=True), (t, k)
        assert np.array_equal(oa[6].info, ob[6].info, equal_nan=True) and np.array_equal(oa[6].rew, ob[6].rew)
    a.close()
    b.close()


def test_lazy_infos_are_fetched_on_access_and_guard_against_stale_reads():
    from gmpe_b200 import default_args
    from gmpe_b200.env_wrappers import GraphSubprocVecEnv

    args = default_args(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi",
                        world_size=3, episode_length=20, goal_rew=30, collaborative=False)
    lazy = GraphSubprocVecEnv(args=args, num_envs=300, device=0, seed=4)
    eager = GraphSubprocVecEnv(args=args, num_envs=300, device=0, seed=4, lazy_info=False)
    lazy.reset(), eager.reset()
    rs = np.random.RandomState(2)
    first = None
    for t in range(5):
        act = np.eye(25)[rs.randint(25, size=(300, 3))]
        ol, oe = lazy.step(act, 0), eager.step(act, 0)
        before = lazy.d2h_bytes_last
        assert before < eager.d2h_bytes_last
        assert ol[6][7][1]["Dist_to_goal"] == oe[6][7][1]["Dist_to_goal"]      # the runner's access pattern
        assert np.array_equal(ol[6].info, oe[6].info, equal_nan=True) and np.array_equal(ol[6].rew, oe[6].rew)
        assert lazy.d2h_bytes_last == eager.d2h_bytes_last > before           # the info rows came over on access
        if first is None:
            first = lazy.step(act, 0)[6]  # never read ...
            eager.step(act, 0)
    with pytest.raises(RuntimeError, match="earlier step"):
        first[0][0]["Dist_to_goal"]       # ... and stale by now
    lazy.close(), eager.close()


@pytest.mark.parametrize("output", ["numpy", "torch"])
def test_rot_inv_infos_carry_phase_reached(output):
    """info['Phase_reached'] (three_phase_graph.py:806) through the vec env: lazy and eager host paths and the
    device-resident path agree with each other and with the state flags of worlds that did not auto-reset."""
    from gmpe_b200 import default_args
    from gmpe_b200.env_wrappers import GraphSubprocVecEnv

    args = default_args(scenario_name="three_phase_graph", dynamics_type="air_taxi", world_size=3, episode_length=15,
                        goal_rew=30, collaborative=False)
    B = 257
    kw = dict(args=args, num_envs=B, device=0, seed=9, output=output)
    lazy, eager = GraphSubprocVecEnv(**kw), GraphSubprocVecEnv(lazy_info=False, **kw)
    lazy.reset(), eager.reset()
    rs = np.random.RandomState(3)
    seen = set()
    for t in range(32):
        idx = rs.randint(25, size=(B, 3)).astype(np.int32)
        act = torch.as_tensor(idx, device="cuda") if output == "torch" else np.eye(25)[idx]
        ol, oe = lazy.step(act, 0), eager.step(act, 0)
        if output == "torch":
            pl, pe = ol[6]["phase_reached"].cpu().numpy(), oe[6]["phase_reached"].cpu().numpy()
        else:
            pl, pe = ol[6].phase_reached, oe[6].phase_reached
            assert ol[6][B - 1][2]["Phase_reached"] == pe[B - 1, 2] and "Phase_reached" in oe[6][0][0].keys()
        assert pl.dtype == np.uint8 and pl.shape == (B, 3) and np.array_equal(pl, pe)
        fl = lazy.engine.state["agent_flags"].cpu().numpy()
        step = lazy.engine.state["world_i"][:, abi.WI_STEP].cpu().numpy()
        live = step > 0  # worlds that auto-reset this step show the terminal value in the info, 0 in the state
        assert np.array_equal(pl[live], fl[live][:, :, abi.AF_PHASE_REACHED])
        seen |= set(np.unique(pl).tolist())
    assert 0 in seen and max(seen) <= 2, seen
    lazy.close(), eager.close()


@pytest.mark.parametrize("kind", ["int32", "float32", "float64"])
def test_step_host_matches_step_for_every_action_encoding(kind):
    """gmpe_step_host (stage, copy up, decode, step, copy down, slice by slice) against gmpe_step
    on a twin engine: same bytes in the pinned mirrors, for index and one-hot action blocks."""
    import ctypes as C

    from gmpe_b200 import buffers
    from gmpe_b200.engine import BatchedGraphMPE

    B, N, A = 5000, 3, 5
    cfg = make_config(default_args(scenario_name="navigation_graph", episode_length=4, collaborative=True), B, seed=9)
    a, b = BatchedGraphMPE(cfg, 0), BatchedGraphMPE(cfg, 0)
    a.reset(), b.reset()
    names = ("obs", "node_obs", "adj", "reward", "done")
    pinned = {k: torch.empty_like(b.out[k], device="cpu").pin_memory() for k in names}
    host = buffers.make_struct(abi.GmpeStepOut, pinned)
    rs = np.random.RandomState(2)
    for t in range(7):  # through two auto-resets
        idx = rs.randint(A, size=(B, N)).astype(np.int32)
        if kind == "int32":
            src = torch.from_numpy(idx)
            dev = torch.empty((B, N), dtype=torch.int32, device="cuda")
        else:
            dt = torch.float32 if kind == "float32" else torch.float64
            src = torch.from_numpy(np.eye(A)[idx]).to(dt)
            dev = torch.empty((B, N, A), dtype=dt, device="cuda")
        pin = torch.empty_like(src).pin_memory()
        a.step(torch.as_tensor(idx, device="cuda"))
        b.step_host(src, pin, dev, host, slices=3)
        b.host_wait()
        for k in names:
            assert torch.equal(a.out[k].cpu(), pinned[k]), f"t={t} {k}"
    # error path: a NULL staging block is refused with the library's message, nothing is launched
    rc = b.lib.gmpe_step_host(b._h, C.byref(b._st), src.data_ptr(), 4, None, dev.data_ptr(), b.action_idx.data_ptr(),
                              C.byref(b._out), C.byref(host), 2, None)
    assert rc == abi.E_NULL and b"NULL" in b.lib.gmpe_last_error(b._h)
    a.close(), b.close()
