"""Stage-1 FORCE physics pinned against the live classic-MPE reference
(onpolicy/envs/mpe/core.py:183-335 driven through environment.py + simple_spread.py;
tests/golden/classic_spread_*.npz, made by tests/golden/make_golden.py:run_classic_trace).

The traces contain agent-agent contacts (122 / 179 penetrating pair-steps).  Only the
agent (x, y, vx, vy) trajectory is compared: simple_spread's reward / observation belong to a
scenario outside the path; the navigation scenario supplies the surrounding step here with
the classic parameterisation (size 0.15, contact 1e2 / 1e-3, d_min = size_a + size_b, no speed
limit -- GMPE_FLAG_CLASSIC_CONTACT)."""
import glob
import os

import numpy as np
import pytest

from gmpe_b200 import abi, default_args, make_config
from oracle.oracle import OracleEnv

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "classic_spread_*.npz")))


def _cfg(z, precision):
    N, L = int(z["num_agents"]), int(z["num_landmarks"])
    args = default_args(scenario_name="navigation_graph", num_agents=N, num_landmarks=L, episode_length=int(z["T"]) + 5,
                        total_actions=5, collaborative=True, max_speed=None)
    c = make_config(args, 1, precision=precision, auto_reset=False, classic_contact=True)
    c.entity_size = float(z["size"])
    assert c.dt == float(z["dt"]) and c.damping == float(z["damping"])
    assert c.contact_force == float(z["contact_force"]) and c.contact_margin == float(z["contact_margin"])
    assert c.max_speed == 0.0
    return c


def _reset_state(z, real):
    N, L = int(z["num_agents"]), int(z["num_landmarks"])
    return {"agent_state": z["init_agent_state"][None].astype(real), "agent_aux": np.zeros((1, N, 2), real),
            "world_f": np.zeros((1, abi.WORLD_F), real), "landmarks": z["init_landmarks"][None].astype(real),
            "obstacles": np.zeros((1, 0, 2), real), "walls": np.zeros((1, 0, abi.WALL_F), real)}


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p))
def test_oracle_force_step_matches_classic_mpe(path):
    z = np.load(path)
    env = OracleEnv(_cfg(z, "fp64"))
    env.reset(reset_state=_reset_state(z, np.float64))
    worst = 0.0
    for t in range(int(z["T"])):
        env.step(z["actions"][t][None])
        worst = max(worst, float(np.abs(env.state["agent_state"][0] - z["state"][t]).max()))
    assert worst <= 1e-10, worst  # free-running over the whole episode, through every contact


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p))
def test_cuda_force_step_matches_classic_mpe(path):
    import torch

    from gmpe_b200.engine import BatchedGraphMPE

    z = np.load(path)
    T = int(z["T"])
    # fp64: teacher-forced per step to 1e-12 over the whole trace; free-running to 1e-8 over the
    # first 40 steps (the stiff contacts amplify last-ulp differences -- FMA contraction, CUDA vs
    # glibc exp/log1p -- by ~2x per contact step: 3e-9 after 50 steps, 1e-6 after 80)
    for teacher in (True, False):
        eng = BatchedGraphMPE(_cfg(z, "fp64"), 0)
        eng.reset(reset_state=_reset_state(z, np.float64))
        prev = z["init_agent_state"]
        for t in range(T if teacher else 40):
            if teacher:
                eng.state["agent_state"][0] = torch.as_tensor(prev, dtype=torch.float64)
            eng.step(torch.as_tensor(z["actions"][t][None], dtype=torch.int32))
            d = np.abs(eng.state["agent_state"][0].cpu().numpy() - z["state"][t]).max()
            assert d <= (1e-12 if teacher else 1e-8), (teacher, t, d)
            prev = z["state"][t]
        eng.close()
    # fp32: teacher-forced per step, 1e-5 relative (north_star)
    eng = BatchedGraphMPE(_cfg(z, "fp32"), 0)
    eng.reset(reset_state=_reset_state(z, np.float32))
    prev = z["init_agent_state"]
    for t in range(T):
        eng.state["agent_state"][0] = torch.as_tensor(prev, dtype=torch.float32)
        eng.step(torch.as_tensor(z["actions"][t][None], dtype=torch.int32))
        got = eng.state["agent_state"][0].cpu().numpy().astype(np.float64)
        want = z["state"][t]
        assert (np.abs(got - want) <= 1e-5 * np.maximum(1.0, np.abs(want))).all(), (t, np.abs(got - want).max())
        prev = want
    eng.close()
