"""TEST INFRASTRUCTURE -- a stand-in for `gmpe_b200.engine.BatchedGraphMPE` backed by the CPU
oracle, so that the vec-env façade (`env_wrappers.py`) can be driven on a box without a GPU by
tests/test_runner_in_the_loop.py.  Same attributes and calls the façade uses, same arena layout;
never imported by the package."""
from __future__ import annotations

import numpy as np
import torch

from gmpe_b200 import abi, buffers
from oracle.oracle import OracleEnv


class OracleEngine:
    def __init__(self, cfg: abi.GmpeConfig, device=None):
        self.cfg = cfg.copy()
        self.device = torch.device("cpu")
        self.real = torch.float64 if cfg.precision == abi.F64 else torch.float32
        self.B, self.N = cfg.num_worlds, cfg.num_agents
        self.layout = abi.layout_for(cfg)
        self.E = self.layout.num_entities
        self.orc = OracleEnv(cfg, use_rk45=False)
        kinds = {"real": self.real, "i32": torch.int32, "u8": torch.uint8}
        sh = buffers.shapes(cfg)
        names = buffers.wanted_outputs(cfg)
        offs, off = {}, 0
        for n in names:  # the engine's arena layout (engine.py): 256-byte aligned slabs in this order
            nbytes = int(np.prod(sh[n][0])) * torch.empty((), dtype=kinds[sh[n][1]]).element_size()
            offs[n] = (off, nbytes)
            off = (off + nbytes + 255) // 256 * 256
        self.out_arena = torch.zeros(max(off, 256), dtype=torch.uint8)
        self.out_slabs = offs
        self.out = {n: self.out_arena[o:o + nb].view(kinds[sh[n][1]]).view(sh[n][0]) for n, (o, nb) in offs.items()}
        host_names = [n for n in names if n in ("obs", "node_obs", "adj", "reward", "done", "indiv_reward", "info", "phase_reached")]
        self.host_prefix_bytes = max(offs[n][0] + offs[n][1] for n in host_names)
        self.action_idx = torch.zeros((self.B, self.N), dtype=torch.int32)
        self.launch_count = 0
        self.steps = []  # (action indices) of every step, for the replay check of the test

    def _publish(self, o):
        for n, t in self.out.items():
            t.copy_(torch.from_numpy(np.ascontiguousarray(o[n])).to(t.dtype))
        return self.out

    @property
    def state(self):
        return {k: torch.from_numpy(v) for k, v in self.orc.state.items()}

    def reset(self, env_mask=None, **kw):
        return self._publish(self.orc.reset(env_mask))

    def step(self, action_idx, **kw):
        a = action_idx.cpu().numpy().astype(np.int32)
        self.steps.append(a.copy())
        self.launch_count += 1
        return self._publish(self.orc.step(a))

    def decode_onehot(self, onehot, out=None):
        out = self.action_idx if out is None else out
        out.copy_(onehot.argmax(dim=-1).to(torch.int32))  # first maximum wins, like np.argmax
        return out

    def host_wait(self):
        pass

    def close(self):
        pass
