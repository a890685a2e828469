"""Out-of-bounds check of our own (compute-sanitizer is closed on this GPU pool): every state
and output array handed to the C ABI is a window inside a larger sentinel-filled allocation;
after resets, steps, masked resets and sliced steps on every kernel mapping the margins must
be untouched and the results must equal those of a plain engine."""
import ctypes as C

import numpy as np
import pytest
import torch

from gmpe_b200 import abi, buffers, default_args, make_config

pytestmark = pytest.mark.gpu
PAD = 4096  # bytes of sentinel on each side (a multiple of 256: windows stay aligned)

CASES = {
    "nav_n3": dict(scenario_name="navigation_graph", episode_length=4, collaborative=True),
    "july_n3": dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", world_size=3,
                    episode_length=4, goal_rew=30, collaborative=False),
    "nav_n5_walls": dict(scenario_name="navigation_graph", num_agents=5, num_landmarks=6, num_obstacles=3, num_walls=4,
                         world_size=2, episode_length=4),
    # GMPE_MAX_AGENTS agents and GMPE_MAX_ENTITIES entities
    "nav_max_entities": dict(scenario_name="navigation_graph", num_agents=64, num_landmarks=64, num_obstacles=28, num_walls=4,
                             world_size=14, episode_length=3),
    "three_phase_n3": dict(scenario_name="three_phase_graph", dynamics_type="air_taxi", world_size=3, episode_length=4,
                           goal_rew=30, collaborative=False),
    "july_n64": dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", num_agents=64,
                     num_landmarks=64, world_size=12, episode_length=3, goal_rew=30, collaborative=False),
}


def _window(shape, dtype):
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    raw = torch.full((n + 2 * PAD,), 0x5A, dtype=torch.uint8, device="cuda")
    win = raw[PAD:PAD + n].view(dtype).view(shape)
    win.zero_()
    return raw, win, n


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_kernels_never_write_outside_their_arrays(case, precision):
    from gmpe_b200.engine import BatchedGraphMPE

    B = 70 if CASES[case].get("num_agents", 3) < 64 else 5
    if precision == "fp64" and case == "nav_max_entities":
        B = 3
    cfg = make_config(default_args(**CASES[case]), B, precision=precision, seed=3, gnn_edges=True)
    ref = BatchedGraphMPE(cfg, 0)
    lib = ref.lib
    real = ref.real
    kinds = {"real": real, "i32": torch.int32, "u8": torch.uint8, "f32": torch.float32}
    sh = buffers.shapes(cfg)
    raws, wins, sizes = {}, {}, {}
    for n in buffers.STATE_FIELDS + buffers.OUT_FIELDS + buffers.MASK_FIELDS:
        raws[n], wins[n], sizes[n] = _window(sh[n][0], kinds[sh[n][1]])
    wins["agent_cnt"][..., abi.AC_TIME_REQ_I:] = -1
    st = buffers.make_struct(abi.GmpeState, {n: wins[n] for n in buffers.STATE_FIELDS})
    out = buffers.make_struct(abi.GmpeStepOut, {n: wins[n] for n in buffers.OUT_FIELDS + buffers.MASK_FIELDS})
    h = C.c_void_p()
    assert lib.gmpe_create(C.byref(cfg), 0, C.byref(h)) == 0
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    inj = abi.GmpeStepInject()
    assert lib.gmpe_reset(h, C.byref(st), None, C.byref(inj), C.byref(out), s) == 0
    ref.reset()
    g = torch.Generator().manual_seed(0)
    for t in range(6):
        a = torch.randint(0, cfg.total_actions, (B, cfg.num_agents), generator=g, dtype=torch.int32).cuda()
        if t == 3:   # the same step issued as two slices
            b0 = B // 3
            assert lib.gmpe_step_range(h, C.byref(st), a.data_ptr(), C.byref(inj), C.byref(out), 0, b0, s) == 0
            assert lib.gmpe_step_range(h, C.byref(st), a.data_ptr(), C.byref(inj), C.byref(out), b0, B - b0, s) == 0
        else:
            assert lib.gmpe_step(h, C.byref(st), a.data_ptr(), C.byref(inj), C.byref(out), s) == 0
        ref.step(a)
    m = torch.zeros(B, dtype=torch.uint8, device="cuda")
    m[::3] = 1
    assert lib.gmpe_reset(h, C.byref(st), m.data_ptr(), C.byref(inj), C.byref(out), s) == 0
    ref.reset(env_mask=m)
    torch.cuda.synchronize()
    for n, raw in raws.items():
        assert (raw[:PAD] == 0x5A).all() and (raw[PAD + sizes[n]:] == 0x5A).all(), f"{n}: wrote outside the array"
    for n in buffers.STATE_FIELDS:
        assert torch.equal(wins[n], ref.state[n]) or torch.equal(torch.nan_to_num(wins[n]), torch.nan_to_num(ref.state[n])), n
    for n in ref.out:
        a_, b_ = wins[n], ref.out[n]
        assert torch.equal(torch.nan_to_num(a_.float()), torch.nan_to_num(b_.float())), n
    lib.gmpe_destroy(h)
    ref.close()
