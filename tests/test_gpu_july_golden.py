"""GPU parity: the CUDA step path (through the C ABI) against the golden traces of the
live reference.  fp64 = strict-parity instantiation (discrete outputs exact, reals to
1e-7: closed-form integrator vs the reference's RK45); fp32 = production instantiation
(per-step teacher-forced: 1e-5 relative as BASELINE.json's north_star states; free-running
200-step drift bound 5e-5 absolute)."""
import numpy as np
import pytest
import torch

from golden_util import Trace, traces
from gmpe_b200 import abi

pytestmark = pytest.mark.gpu
ALL = traces("july_*.npz") + traces("three_phase_*.npz") + traces("two_phase_*.npz") + traces("corridor_rot_inv_*.npz")


def _close(a, b, atol, rtol, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), f"{what}: max |diff| {np.abs(a - b).max():.3e} (atol {atol}, rtol {rtol})"


def run_trace(tr: Trace, precision: str, teacher_force: bool, atol: float, rtol: float):
    from gmpe_b200.engine import BatchedGraphMPE

    eng = BatchedGraphMPE(tr.cfg(precision=precision), 0)
    z = tr.z
    N = tr.N
    for ep in range(tr.episodes):
        eng.state["agent_flags"][0, :, abi.AF_PREV_PHASE] = torch.as_tensor(z["init_prev_phase"][ep], dtype=torch.uint8)
        out = eng.reset(reset_state=tr.reset_state(ep))
        _close(out["obs"][0].cpu(), z["reset_obs"][ep], atol, rtol, "reset obs")
        _close(out["node_obs"][0].cpu(), z["reset_node_obs"][ep], atol, rtol, "reset node_obs")
        _close(out["adj"][0].cpu(), z["reset_adj"][ep], atol, rtol, "reset adj")
        for t in tr.step_rows(ep):
            if teacher_force and t > ep * tr.T:
                eng.state["agent_state"][0] = torch.as_tensor(z["state"][t - 1], dtype=eng.real)
                eng.state["agent_aux"][0, :, 0] = torch.as_tensor(z["p_dist"][t - 1], dtype=eng.real)
            th = torch.as_tensor(z["theta_draw"][t][None], dtype=eng.real)
            out = eng.step(torch.as_tensor(z["actions"][t][None], dtype=torch.int32), theta_draw=th)
            tag = f"{tr.name} t={t}"
            st = eng.state["agent_state"][0].cpu().numpy().astype(np.float64)
            gs = z["state"][t]
            _close(st[:, [0, 1, 3]], gs[:, [0, 1, 3]], atol, rtol, tag + " state")
            dth = (st[:, 2] - gs[:, 2] + np.pi) % (2 * np.pi) - np.pi  # fp32 keeps theta wrapped
            assert np.abs(dth).max() <= 10 * atol, tag + " theta"
            _close(out["obs"][0].cpu(), z["obs"][t], atol, rtol, tag + " obs")
            _close(out["node_obs"][0].cpu(), z["node_obs"][t], atol, rtol, tag + " node_obs")
            _close(out["adj"][0].cpu(), z["adj"][t], atol, rtol, tag + " adj")
            _close(out["reward"][0].cpu(), z["reward"][t], atol, rtol, tag + " reward")
            _close(out["edge_attr"][0].cpu(), z["edge_attr"][t], atol, rtol, tag + " edge_attr")
            _close(out["gnn_edge_attr"][0].cpu(), z["gnn_edge_attr"][t], atol, rtol, tag + " gnn_edge_attr")
            # info: integer-valued columns exact, ratios to tolerance (the mean / (std + 1e-4) columns,
            # july:800-823, reach 3e4 when every agent holds the same value: the kernels' means are
            # exact for that reason, see mean_div in gmpe_device.cuh)
            _close(out["info"][0].cpu(), z["info"][t], max(atol, 1e-6), max(rtol, 1e-6), tag + " info")
            # discrete outputs: exact
            assert (out["done"][0].cpu().numpy().astype(bool) == z["done"][t]).all(), tag + " done"
            fl = eng.state["agent_flags"][0].cpu().numpy()
            assert (fl[:, abi.AF_STATUS].astype(bool) == z["status"][t]).all(), tag + " status"
            assert (fl[:, abi.AF_PREV_PHASE] == z["prev_phase"][t]).all(), tag + " prev_phase"
            assert (fl[:, abi.AF_PHASE_REACHED] == z["phase_reached"][t]).all(), tag + " phase_reached"
            assert (fl[:, abi.AF_GOAL_REACHED].astype(int) - 1 == z["goal_reached"][t]).all(), tag + " goal_reached"
            assert (out["obs"][0][:, -1].cpu().numpy() == z["obs"][t][:, -1]).all(), tag + " phase"
            if tr.name.startswith(("three_phase", "two_phase", "corridor_rot_inv")):
                assert (eng.state["agent_ext"][0, :, 0].cpu().numpy() == z["cooldown"][t]).all(), tag + " cooldown"
                assert (out["phase_reached"][0].cpu().numpy() == z["phase_reached"][t]).all(), tag + " info Phase_reached"
                if tr.name.startswith("corridor_rot_inv"):  # prev_proj: float32 bits in the second slot
                    pj = eng.state["agent_ext"][0, :, 1].cpu().numpy().view(np.float32)
                    assert np.abs(pj - z["prev_proj"][t]).max() <= max(atol, 1e-6), tag + " prev_proj"
            assert (eng.state["agent_cnt"][0].cpu().numpy() == z["cnt"][t]).all(), tag + " counters"
            assert int(out["num_edges"][0]) == int(z["num_edges"][t]), tag + " num_edges"
            assert (out["edge_index"][0].cpu().numpy() == z["edge_index"][t]).all(), tag + " edge_index"
            assert int(out["gnn_num_edges"][0]) == int(z["gnn_num_edges"][t]), tag + " gnn_num_edges"
            assert (out["gnn_edge_index"][0].cpu().numpy() == z["gnn_edge_index"][t]).all(), tag + " gnn_edge_index"
            assert ((out["adj"][0].cpu().numpy() == 0) == (z["adj"][t] == 0)).all(), tag + " adj mask"
    eng.close()


@pytest.mark.parametrize("path", ALL, ids=lambda p: p.split("/")[-1])
def test_fp64_free_running_bit_exact_graph(path):
    # three_phase emits float32 features (one float32 ulp of an O(1) value) and passes the
    # integrator's heading difference on to the reward through its heading-error terms
    atol = 1e-6 if ("three_phase" in path or "two_phase" in path or "corridor_rot_inv" in path) else 2e-7
    run_trace(Trace(path), "fp64", teacher_force=False, atol=atol, rtol=0.0)


@pytest.mark.parametrize("path", ALL, ids=lambda p: p.split("/")[-1])
def test_fp32_per_step_within_1e5(path):
    run_trace(Trace(path), "fp32", teacher_force=True, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("path", ALL, ids=lambda p: p.split("/")[-1])
def test_fp32_free_running_drift_bound(path):
    run_trace(Trace(path), "fp32", teacher_force=False, atol=5e-5, rtol=1e-5)
