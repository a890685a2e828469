"""Reset parity: the placement of `reset_world / random_scenario /
setup_tube_params` (july:339-613, three_phase_graph.py:416-587) and the landmark formations
(utils.py:77-262), against the LIVE reference (tests/golden/reset_parity.npz, made by
tests/golden/make_reset_golden.py).

1. Exact: the reset is fed the reference's own np.random stream (`GmpeStepInject.reset_draws`; the
   draws of a reset seeded with s are RandomState(s).random_sample(K)) and must place every agent,
   heading, landmark and the tube frame where the reference did -- oracle 1e-12 on the CPU, CUDA
   fp64 1e-12 / fp32 1e-6 on the GPU, both kernel mappings.
2. Statistical: the production reset draws from Philox, not MT19937; its marginals (tube angle and
   length, every agent's x / y / heading) are compared with 2,000 reference resets by a two-sample
   Kolmogorov-Smirnov test."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest

from gmpe_b200 import abi, make_config
from oracle.oracle import OracleEnv

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reset_parity.npz"))
CFGS = json.loads(str(Z["cfg_json"]))
K = 96


def _cfg(name, B, precision="fp64", seed=3):
    return make_config(SimpleNamespace(**CFGS[name]), B, precision=precision, seed=seed, auto_reset=True)


def _draws(name):
    return np.stack([np.random.RandomState(int(s)).random_sample(K) for s in Z[f"{name}/seeds"]])


def _check(name, st, tol, wrap=False):
    want_as = Z[f"{name}/agent_state"]
    got_as = np.asarray(st["agent_state"], dtype=np.float64)
    d = np.abs(got_as - want_as)
    if wrap:  # the fp32 instantiation keeps headings wrapped to [-pi, pi)
        dth = np.abs((got_as[..., 2] - want_as[..., 2] + np.pi) % (2 * np.pi) - np.pi)
        d[..., 2] = dth
    assert d.max() <= tol, (name, "agent_state", d.max())
    assert np.abs(np.asarray(st["landmarks"], dtype=np.float64) - Z[f"{name}/landmarks"]).max() <= tol, (name, "landmarks")
    assert np.abs(np.asarray(st["world_f"], dtype=np.float64)[:, :12] - Z[f"{name}/tube"]).max() <= tol, (name, "tube")
    gmt = Z[f"{name}/goal_min_time"]
    assert np.abs(np.asarray(st["agent_aux"], dtype=np.float64)[..., 1] - gmt).max() <= tol * np.maximum(1, np.abs(gmt)).max(), (name, "min_time")


@pytest.mark.parametrize("name", list(CFGS))
def test_oracle_reset_on_the_reference_random_stream(name):
    R = len(Z[f"{name}/seeds"])
    assert int(Z[f"{name}/draws_used"].max()) < K
    env = OracleEnv(_cfg(name, R))
    env.reset(reset_draws=_draws(name))
    _check(name, env.state, 1e-12)


def _ks(a, b):
    """two-sample Kolmogorov-Smirnov statistic"""
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    return float(np.abs(np.searchsorted(a, allv, side="right") / a.size - np.searchsorted(b, allv, side="right") / b.size).max())


def _ks_check(name, st):
    ref_as, ref_tube = Z[f"{name}/ks_agent_state"].astype(np.float64), Z[f"{name}/ks_tube"].astype(np.float64)
    got_as, got_wf = np.asarray(st["agent_state"], dtype=np.float64), np.asarray(st["world_f"], dtype=np.float64)
    n, m = ref_as.shape[0], got_as.shape[0]
    crit = 1.95 * np.sqrt((n + m) / (n * m))  # alpha = 0.001
    cols = {"tube angle": (ref_tube[:, 0], got_wf[:, abi.WF_ANGLE]), "tube L": (ref_tube[:, 1], got_wf[:, abi.WF_L])}
    for i in range(ref_as.shape[1]):
        for c, cn in enumerate("xy"):
            cols[f"agent {i} {cn}"] = (ref_as[:, i, c], got_as[:, i, c])
        cols[f"agent {i} heading"] = (ref_as[:, i, 2] % (2 * np.pi), got_as[:, i, 2] % (2 * np.pi))
    worst = {}
    for k, (a, b) in cols.items():
        if np.ptp(a) < 1e-9:  # a constant of the scenario (fixed tube length): must simply agree
            assert np.abs(b - a[0]).max() <= 1e-6, k
            continue
        worst[k] = _ks(a, b)
    bad = {k: v for k, v in worst.items() if v > crit}
    assert not bad, (name, crit, bad)
    return max(worst.values()), crit


@pytest.mark.parametrize("name", [n for n in CFGS if f"{n}/ks_tube" in Z.files])
def test_oracle_philox_reset_has_the_reference_distribution(name):
    env = OracleEnv(_cfg(name, 16384, seed=11), nthreads=8)
    env.reset()
    _ks_check(name, env.state)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CFGS))
def test_cuda_reset_on_the_reference_random_stream(name):
    from gmpe_b200.engine import BatchedGraphMPE

    R = len(Z[f"{name}/seeds"])
    for precision, tol in (("fp64", 1e-12), ("fp32", 1e-6)):
        eng = BatchedGraphMPE(_cfg(name, R, precision), 0)
        eng.reset(reset_draws=_draws(name))
        _check(name, {k: v.cpu().numpy() for k, v in eng.state.items()}, tol, wrap=precision == "fp32")
        assert eng.placement_overflows == 0
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in CFGS if f"{n}/ks_tube" in Z.files])
def test_cuda_philox_reset_has_the_reference_distribution(name):
    from gmpe_b200.engine import BatchedGraphMPE

    eng = BatchedGraphMPE(_cfg(name, 65536, "fp32", seed=11), 0)
    eng.reset()
    _ks_check(name, {k: v.cpu().numpy() for k, v in eng.state.items()})
    eng.close()
