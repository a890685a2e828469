"""CPU-only tests of the host-side logic: config mapping, action decoding front half,
spaces, world sharding, the gloo (world_size 2) statistics reduction and the lazy info view."""
import os
import subprocess
import sys

import numpy as np
import pytest

from gmpe_b200 import abi, config, default_args, make_config
from gmpe_b200.dist import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_constants_follow_the_reference_config_module():
    c = make_config(default_args(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi",
                                 world_size=3, goal_rew=30), 8)
    assert c.total_actions == 25 and c.dt == 1.0
    assert c.v_min == pytest.approx(60 * 0.514444 * 0.001) and c.v_max == pytest.approx(175 * 0.514444 * 0.001)
    assert c.separation_dist == pytest.approx(1500 * 0.0003048)
    assert c.coordination_range == pytest.approx(3 * 1.60934)
    assert c.goal_dist_thresh == 0.35 and c.sensitivity == 5.0
    assert c.physics == abi.PHYS_KINEMATIC
    n = make_config(default_args(), 8)
    assert n.scenario == abi.SCEN_NAV_GRAPH and n.physics == abi.PHYS_FORCE and n.total_actions == 5
    assert n.contact_force == 3e2 and n.contact_margin == 2e-2 and n.separation_dist == 0.5


def test_config_rejects_what_the_reference_cannot_run():
    with pytest.raises(ValueError):  # obstacles crash the july graph builder
        make_config(default_args(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi",
                                 num_obstacles=2), 2)
    with pytest.raises(ValueError):
        make_config(default_args(scenario_name="simple_spread"), 2)
    with pytest.raises(NotImplementedError):
        make_config(default_args(dynamics_type="bicycle"), 2)
    with pytest.raises(ValueError):
        make_config(default_args(num_agents=5, num_landmarks=3), 2)
    with pytest.raises(NotImplementedError):
        make_config(default_args(use_safety_filter=True), 2)
    # the rot-inv family is served for 2 <= N = L <= 16 (a graph row of E = 2N entities in one 32-bit lane mask)
    make_config(default_args(scenario_name="three_phase_graph", dynamics_type="air_taxi", num_agents=16, num_landmarks=16), 2)
    with pytest.raises(ValueError, match="rot-inv"):
        make_config(default_args(scenario_name="three_phase_graph", dynamics_type="air_taxi", num_agents=17, num_landmarks=17), 2)


def test_actions_to_index_mirrors_set_action_front_half():
    from gmpe_b200.environment import actions_to_index

    a = actions_to_index([np.eye(25)[7], np.eye(25)[24], np.eye(25)[0]], 25)
    assert a.tolist() == [7, 24, 0]
    assert actions_to_index([[3], [1]], 5, discrete_action_input=True).tolist() == [3, 1]
    with pytest.raises(AssertionError):  # the reference asserts the whole action was consumed
        actions_to_index([np.eye(6)[1]], 5)


def test_spaces_have_reference_shapes():
    from gmpe_b200.environment import build_spaces

    args = default_args(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi",
                        world_size=3)
    sp = build_spaces(args, make_config(args, 1))
    assert sp["action_space"][0].n == 25 and sp["action_space"][0].__class__.__name__ == "Discrete"
    assert sp["observation_space"][0].shape == (19,) and sp["share_observation_space"][0].shape == (57,)
    assert sp["node_observation_space"][0].shape == (6, 8) and sp["adj_observation_space"][0].shape == (6, 6)
    assert sp["edge_observation_space"][0].shape == (1,) and sp["share_agent_id_observation_space"][0].shape == (3,)
    assert sp["observation_space"][0].__class__.__name__ == "Box"


def test_shard_partitions_worlds_contiguously():
    for total, ws in [(10, 3), (65536 * 8, 8), (7, 8), (0, 2)]:
        blocks = [shard(total, r, ws) for r in range(ws)]
        assert sum(c for _, c in blocks) == total
        off = 0
        for o, c in blocks:
            assert o == off
            off += c
    with pytest.raises(ValueError):
        shard(4, 2, 2)


def test_lazy_infos_walks_like_process_infos():
    from gmpe_b200.env_wrappers import LazyInfos

    info = np.arange(2 * 3 * 16, dtype=np.float32).reshape(2, 3, 16)
    rew = np.ones((2, 3), dtype=np.float32) * 7
    infos = LazyInfos(info, rew)
    assert len(infos) == 2
    for e, env_info in enumerate(infos):
        for a in range(3):
            d = env_info[a]
            assert "individual_reward" in d.keys() and d["individual_reward"] == 7
            assert d["Dist_to_goal"] == info[e, a, 0] and d["Min_time_to_goal"] == info[e, a, 15]
            assert set(d.to_dict()) == {"individual_reward", *abi.INFO_KEYS}
    # rot-inv family: 'Phase_reached' sits before 'Min_time_to_goal' (three_phase_graph.py:788-811)
    ph = np.array([[0, 1, 2], [2, 1, 0]], dtype=np.uint8)
    infos = LazyInfos(info, rew, phase_reached=ph)
    d = infos[1][0]
    assert list(d.keys())[-2:] == ["Phase_reached", "Min_time_to_goal"] and len(d) == 18
    assert d["Phase_reached"] == 2 and infos[0][1]["Phase_reached"] == 1 and d["Min_time_to_goal"] == info[1, 0, 15]
    assert "Phase_reached" not in LazyInfos(info, rew)[0][0]


GLOO_WORKER = r'''
import os, sys, torch
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from gmpe_b200 import dist as gd
rank, local, world = gd.init_from_env("gloo")
off, cnt = gd.shard(1000, rank, world)
stats = torch.tensor([float(cnt), float(off), 1.0], dtype=torch.float64)
gd.reduce_stats(stats)
t = gd.max_over_ranks(1.0 + rank)
gd.barrier()
if rank == 0:
    print("RESULT", stats.tolist(), t, flush=True)
dist.destroy_process_group()
'''


def test_gloo_world_size_2_stats_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    port = 29500 + (os.getpid() % 2000)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
    assert "[1000.0, 500.0, 2.0]" in line and line.rstrip().endswith("2.0")
