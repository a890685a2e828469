"""Stage-1 FORCE physics, `multiagent` parameterisation, pinned against the reference's OWN
functions (multiagent/core.py:766-964: apply_action_force, apply_environment_force,
get_entity_collision_force with its COLLISION_DISTANCE / 3e2 / 2e-2 contact and `status`
suppression, get_wall_collision_force with the end-cap rounding, integrate_state), driven
unmodified in the classic step order on a double-integrator world with obstacles, four wall
segments and (two traces) an agent that is done from mid-episode on
(tests/golden/multiagent_force_*.npz, made by tests/golden/make_golden.py:run_multiagent_force_trace).

This is what BASELINE config 4 (navigation_graph with 8 obstacles + 4 walls) rests on.  Only the
agent (x, y, vx, vy) trajectory is compared -- reward / observation of the navigation scenario are
defined by this build.

One deliberate deviation shows up and is kept out of the comparison: a done agent is FROZEN by the
build (like the live step's update_agent_state, core.py:823-825), whereas these force functions keep
integrating it -- the reference suppresses only the agent-agent force on it (core.py:899-900), an
obstacle or a wall still pushes it (2.6e-3 per step in trace s52).  From `done_at` on the done agent's
state is therefore taken from the trace every step (the others must feel it where the reference had
it) and the comparison covers the other agents."""
import glob
import os

import numpy as np
import pytest

from gmpe_b200 import abi, default_args, make_config
from oracle.oracle import OracleEnv

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multiagent_force_*.npz")))


def _cfg(z, precision):
    N, L, O, W = (int(z[k]) for k in ("num_agents", "num_landmarks", "num_obstacles", "num_walls"))
    args = default_args(scenario_name="navigation_graph", num_agents=N, num_landmarks=L, num_obstacles=O, num_walls=W,
                        episode_length=int(z["T"]) + 5, total_actions=5, collaborative=True, use_dones=False,
                        max_speed=float(z["max_speed"]), world_size=4)  # world_size only sizes the Philox spawn box (unused: states are injected)
    c = make_config(args, 1, precision=precision, auto_reset=False, wall_point_quirk=bool(int(z["wall_points"])))
    # every constant of the force path is the reference's own (read back from the live World / config)
    assert c.entity_size == float(z["size"]) and c.dt == float(z["dt"]) and c.damping == float(z["damping"])
    assert c.contact_force == float(z["contact_force"]) and c.contact_margin == float(z["contact_margin"])
    assert c.wall_contact_force == float(z["wall_contact_force"]) and c.wall_contact_margin == float(z["wall_contact_margin"])
    assert c.wall_width == float(z["wall_width"]) and c.separation_dist == float(z["collision_distance"])
    assert c.max_speed == float(z["max_speed"])
    return c


def _reset_state(z, real):
    N = int(z["num_agents"])
    return {"agent_state": z["init_agent_state"][None].astype(real), "agent_aux": np.zeros((1, N, 2), real),
            "world_f": np.zeros((1, abi.WORLD_F), real), "landmarks": z["init_landmarks"][None].astype(real),
            "obstacles": z["obstacles"][None].astype(real), "walls": z["walls"][None].astype(real)}


def test_traces_exist_and_exercise_every_contact():
    assert len(GOLDEN) == 3
    for p in GOLDEN:
        z = np.load(p)
        assert np.isfinite(z["state"]).all()


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p))
@pytest.mark.parametrize("teacher", [True, False], ids=["teacher_forced", "free_running"])
def test_oracle_force_step_matches_multiagent_core(path, teacher):
    """Teacher-forced from the trace every step: 1e-13 over the whole episode (measured: <= 7e-16,
    i.e. the restatement IS the reference's arithmetic).  Free-running: 1e-10 over the first 55
    steps -- the contacts are stiff (3e2 / 2e-2) and amplify a last-ulp difference by ~1.2x per step
    (1e-16 at step 50 -> 3e-6 at step 150 in trace s51), so a bound over the whole episode would
    test chaos, not the restatement."""
    z = np.load(path)
    env = OracleEnv(_cfg(z, "fp64"))
    env.reset(reset_state=_reset_state(z, np.float64))
    done_at = int(z["done_at"])
    worst = 0.0
    prev = z["init_agent_state"]
    for t in range(int(z["T"]) if teacher else 55):
        live = slice(None)
        if teacher:
            env.state["agent_state"][0] = prev
        if 0 <= done_at <= t:  # agent.status = True (what reward() does on goal reach, july:1188-1190)
            env.state["agent_flags"][0, 0, abi.AF_STATUS] = 1
            env.state["agent_state"][0, 0] = (z["state"][t - 1][0] if t > done_at else
                                              np.r_[env.state["agent_state"][0, 0, :2], 0.0, 0.0])  # reset_velocity()
            live = slice(1, None)
        env.step(z["actions"][t][None])
        worst = max(worst, float(np.abs(env.state["agent_state"][0, live] - z["state"][t][live]).max()))
        prev = z["state"][t]
    assert worst <= (1e-13 if teacher else 1e-10), worst


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p))
def test_cuda_force_step_matches_multiagent_core(path):
    import torch

    from gmpe_b200.engine import BatchedGraphMPE

    z = np.load(path)
    T, done_at = int(z["T"]), int(z["done_at"])
    for precision, real, tol in (("fp64", np.float64, 1e-12), ("fp32", np.float32, 1e-5)):
        eng = BatchedGraphMPE(_cfg(z, precision), 0)
        eng.reset(reset_state=_reset_state(z, real))
        prev = z["init_agent_state"]
        for t in range(T):  # teacher-forced per step (the stiff contacts amplify last-ulp differences when free-running)
            live = slice(None)
            if 0 <= done_at <= t:
                eng.state["agent_flags"][0, 0, abi.AF_STATUS] = 1
                live = slice(1, None)
                if t == done_at:
                    prev = prev.copy()
                    prev[0, 2:] = 0.0
            eng.state["agent_state"][0] = torch.as_tensor(prev, dtype=eng.real)
            eng.step(torch.as_tensor(z["actions"][t][None], dtype=torch.int32))
            got = eng.state["agent_state"][0].cpu().numpy().astype(np.float64)[live]
            want = z["state"][t][live]
            assert (np.abs(got - want) <= tol * np.maximum(1.0, np.abs(want))).all(), (precision, t, np.abs(got - want).max())
            prev = z["state"][t]
        eng.close()
