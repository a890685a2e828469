"""SURVEY 8f rank 2 in the layout the consumer takes: `gmpe_gnn_flatten` against process_adj
(onpolicy/algorithms/utils/gnn_new.py:328-358) run on the (B*N, E, E) stack GNNBase.forward builds
(:492-500).  `process_adj_ref` restates that function for the GPU box (no /root/reference there);
the CPU test checks the restatement against the reference's own source where it is present."""
import ast
import os
import textwrap

import numpy as np
import pytest
import torch

from gmpe_b200 import default_args, make_config


def process_adj_ref(adj: torch.Tensor, max_edge_dist: float):
    """gnn_new.py:328-358, batched branch."""
    mask = ((adj < max_edge_dist) & (adj > 0)).float()
    adj = adj * mask
    bs, n, _ = adj.shape
    idx = adj.nonzero(as_tuple=False)
    attr = adj[idx[:, 0], idx[:, 1], idx[:, 2]]
    batch = idx[:, 0] * n
    return torch.stack([batch + idx[:, 1], batch + idx[:, 2]], dim=0), attr.unsqueeze(1)


def test_restatement_equals_the_reference_source():
    path = "/root/reference/onpolicy/algorithms/utils/gnn_new.py"
    if not os.path.exists(path):
        pytest.skip("reference not present (GPU box)")
    from typing import Tuple

    from torch import Tensor

    tree = ast.parse(open(path).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "process_adj")
    fn.decorator_list = []
    ns = {"torch": torch, "Tensor": Tensor, "Tuple": Tuple}
    exec(textwrap.dedent(ast.unparse(fn)), ns)
    g = torch.Generator().manual_seed(3)
    adj = torch.rand((37, 6, 6), generator=g) * 2.0
    adj[:, torch.arange(6), torch.arange(6)] = 0
    adj[5] = 0  # a graph without edges
    ei, ea = ns["process_adj"](adj, 1.0)
    ei2, ea2 = process_adj_ref(adj, 1.0)
    assert torch.equal(ei, ei2) and torch.equal(ea, ea2) and ei.dtype == torch.int64


CASES = {
    "july_n3": (dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", world_size=3,
                     episode_length=12, goal_rew=30, collaborative=False), 3000, 1.0),
    "nav_n3_dense": (dict(scenario_name="navigation_graph", episode_length=10, max_edge_dist=5.0), 2500, 5.0),
    "nav_n16_obst_walls": (dict(scenario_name="navigation_graph", num_agents=16, num_landmarks=16, num_obstacles=8, num_walls=4,
                                world_size=4, episode_length=8, max_edge_dist=1.5), 70, 1.5),
    "july_n64": (dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", num_agents=64,
                      num_landmarks=64, world_size=12, episode_length=5, goal_rew=30, collaborative=False), 9, 1.0),
    "three_phase_n5": (dict(scenario_name="three_phase_graph", dynamics_type="air_taxi", world_size=3, num_agents=5,
                            num_landmarks=5, episode_length=12, goal_rew=20, collaborative=False), 1200, 1.0),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_flat_edges_equal_process_adj_on_the_stacked_adjacency(case):
    from gmpe_b200.engine import BatchedGraphMPE

    kw, B, thr = CASES[case]
    cfg = make_config(default_args(**kw), B, precision="fp32", seed=9, gnn_edges=True)
    assert cfg.gnn_max_edge_dist == thr
    eng = BatchedGraphMPE(cfg, 0)
    eng.reset()
    N, E = cfg.num_agents, eng.E
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(14):  # through auto-resets and done agents (masked rows / columns)
        a = torch.randint(0, cfg.total_actions, (B, N), generator=g, device="cuda", dtype=torch.int32)
        out = eng.step(a)
        if t % 3 and t != 13:
            continue
        ei, ea = eng.gnn_flat_edges()
        stack = out["adj"].unsqueeze(1).expand(B, N, E, E).reshape(B * N, E, E)  # what the runner hands to GNNBase.forward
        want_ei, want_ea = process_adj_ref(stack, thr)
        assert ei.dtype == torch.int64 and tuple(ei.shape) == tuple(want_ei.shape), (t, ei.shape, want_ei.shape)
        assert torch.equal(ei, want_ei), t
        assert torch.equal(ea, want_ea), t
        assert int(out["gnn_num_edges"].sum()) * N == ei.shape[1]
    # a capacity that is too small is reported, not overrun
    if ei.shape[1] > 8:
        eng._flat = None
        with pytest.raises(Exception):
            eng.gnn_flat_edges(capacity=ei.shape[1] - 1)
    eng.close()
