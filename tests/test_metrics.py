"""env metrics as tensor reductions, pinned against the
reference's Runner.process_infos + log_env run from their own source
(tests/golden/process_infos_n3.npz, tests/golden/make_rollout_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from gmpe_b200 import metrics

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "process_infos_n3.npz")


def test_process_infos_matches_the_reference_means():
    z = np.load(GOLD)
    want = json.loads(str(z["logged_json"]))
    got = metrics.process_infos(torch.as_tensor(z["info"]), torch.as_tensor(z["indiv_reward"]), int(z["episode_length"]),
                                float(z["dt"]), reduce=False)
    assert set(got) == set(want)
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-12 * max(1.0, abs(v)), (k, got[k], v)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    z = np.load(GOLD)
    info, rew = torch.as_tensor(z["info"]), torch.as_tensor(z["indiv_reward"])
    lo, hi = (0, 4) if rank == 0 else (4, info.shape[0])  # uneven shards
    got = metrics.process_infos(info[lo:hi], rew[lo:hi], int(z["episode_length"]), float(z["dt"]))
    q.put((rank, got))
    dist.destroy_process_group()


def test_process_infos_reduces_across_ranks_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    want = json.loads(str(np.load(GOLD)["logged_json"]))
    for r in (0, 1):
        for k, v in want.items():
            assert abs(res[r][k] - v) <= 1e-12 * max(1.0, abs(v)), (r, k)


@pytest.mark.gpu
def test_process_infos_on_the_device_matches_the_lazy_dict_walk():
    from gmpe_b200 import default_args
    from gmpe_b200.env_wrappers import GraphSubprocVecEnv

    args = default_args(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", world_size=3,
                        episode_length=40, goal_rew=30, collaborative=False)
    venv = GraphSubprocVecEnv(args=args, num_envs=64, device=0, seed=2)
    venv.reset()
    rs = np.random.RandomState(0)
    for t in range(30):
        out = venv.step(np.eye(25)[rs.randint(25, size=(64, 3))], 0)
    infos = out[6]
    eng = venv.engine
    got = metrics.process_infos(eng.out["info"], eng.out["indiv_reward"], args.episode_length, eng.cfg.dt, reduce=False)
    # the reference's walk over infos[env][agent][key] (base_runner.py:208-247) on the lazy view
    for a in range(3):
        for key, name in metrics.LOG_NAMES.items():
            vals = []
            for info in infos:
                v = info[a][key]
                if key == "Time_req_to_goal" and v == -1:
                    v = args.episode_length * eng.cfg.dt
                vals.append(float(v))
            assert abs(got[f"agent{a}/{name}"] - np.mean(vals)) <= 1e-6 * max(1.0, abs(np.mean(vals))), (a, key)
    venv.close()
