"""BatchedGraphMPE -- one GPU-resident batch of B independent GraphMPE worlds.

Owns the device tensors (state + per-step outputs, all caller-owned as far as the C ABI
is concerned) and drives the CUDA library through ``include/gmpe.h``.  torch is used for
device memory and streams only; every number is produced by the hand-written kernels in
``csrc/``.  There is no CPU fallback: constructing an engine without the built library or
without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi, buffers


class BatchedGraphMPE:
    def __init__(self, cfg: abi.GmpeConfig, device: int | str | torch.device = 0):
        if not torch.cuda.is_available():
            raise abi.GmpeError("BatchedGraphMPE needs a CUDA device (no CPU fallback for the step path)")
        self.lib = abi.load_library()
        self.cfg = cfg.copy()
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise abi.GmpeError("device must be a CUDA device")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device(f"cuda:{dev_index}")
        self.real = torch.float64 if cfg.precision == abi.F64 else torch.float32
        self.B, self.N = cfg.num_worlds, cfg.num_agents
        self.layout = abi.layout_for(cfg)
        self.E = self.layout.num_entities
        h = C.c_void_p()
        rc = self.lib.gmpe_create(C.byref(self.cfg), dev_index, C.byref(h))
        if rc != abi.OK:
            raise abi.GmpeError(f"gmpe_create failed ({rc}): {self.lib.gmpe_last_error(None).decode()}")
        self._h = h
        kinds = {"real": self.real, "i32": torch.int32, "u8": torch.uint8}
        sh = buffers.shapes(cfg)
        with torch.cuda.device(self.device):
            self.state = {n: torch.zeros(sh[n][0], dtype=kinds[sh[n][1]], device=self.device) for n in buffers.STATE_FIELDS}
            self.state["agent_cnt"][..., abi.AC_TIME_REQ_I:] = -1
            # All outputs are views of ONE device arena (256-byte aligned slabs) so that a host
            # consumer can fetch everything a step returns with a single device->host copy.
            names = buffers.wanted_outputs(cfg)
            offs, off = {}, 0
            for n in names:
                nbytes = int(np.prod(sh[n][0])) * torch.empty((), dtype=kinds[sh[n][1]]).element_size()
                offs[n] = (off, nbytes)
                off = (off + nbytes + 255) // 256 * 256
            self.out_arena = torch.zeros(max(off, 256), dtype=torch.uint8, device=self.device)
            self.out_slabs = offs
            self.out = {n: self.out_arena[o:o + nb].view(kinds[sh[n][1]]).view(sh[n][0]) for n, (o, nb) in offs.items()}
            #: bytes from the arena start that hold what MultiAgentGraphEnv.step returns
            #: (obs, node_obs, adj, reward, done[, indiv_reward, info[, phase_reached]]); edges come after
            host_names = [n for n in names if n in ("obs", "node_obs", "adj", "reward", "done", "indiv_reward", "info", "phase_reached")]
            self.host_prefix_bytes = max(offs[n][0] + offs[n][1] for n in host_names)
            self.action_idx = torch.zeros((self.B, self.N), dtype=torch.int32, device=self.device)
        self._st = buffers.make_struct(abi.GmpeState, self.state)
        self._out = buffers.make_struct(abi.GmpeStepOut, self.out)
        self._null_inject = abi.GmpeStepInject()

    # -- plumbing -------------------------------------------------------------------
    def _check(self, rc: int, what: str) -> None:
        if rc != abi.OK:
            raise abi.GmpeError(f"{what} failed ({rc}): {self.lib.gmpe_last_error(self._h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.gmpe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_seed(self, seed: int, world_offset: int | None = None) -> None:
        """env.seed(seed) of the reference (environment.py:219-223): re-keys the Philox streams."""
        if world_offset is None:
            world_offset = self.cfg.world_offset
        self.cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.cfg.world_offset = int(world_offset)
        self._check(self.lib.gmpe_set_seed(self._h, self.cfg.seed, self.cfg.world_offset), "gmpe_set_seed")

    @property
    def launch_count(self) -> int:
        return int(self.lib.gmpe_launch_count(self._h))

    @property
    def kernel_name(self) -> str:
        """Which step kernel gmpe_create picked for this configuration."""
        return self.lib.gmpe_kernel_name(self._h).decode()

    @property
    def placement_overflows(self) -> int:
        """Worlds whose spawn placement gave up a rejection loop so far (synchronises the device)."""
        n = C.c_int64(0)
        self._check(self.lib.gmpe_placement_overflows(self._h, C.byref(n), 0), "gmpe_placement_overflows")
        return int(n.value)

    def _inject(self, theta_draw, reset_state, reset_draws=None):
        if theta_draw is None and reset_state is None and reset_draws is None:
            return self._null_inject, None
        if theta_draw is not None:
            theta_draw = torch.as_tensor(theta_draw, dtype=self.real, device=self.device).contiguous()
        if reset_state is not None:
            rs = {}
            for k, v in reset_state.items():
                t = torch.as_tensor(v, device=self.device)
                rs[k] = t.to(self.real if t.is_floating_point() else t.dtype).contiguous()
            reset_state = rs
        if reset_draws is not None:
            reset_draws = torch.as_tensor(reset_draws, dtype=torch.float64, device=self.device).contiguous()
            if reset_draws.ndim != 2 or reset_draws.shape[0] != self.B:
                raise ValueError(f"reset_draws must have shape ({self.B}, K)")
        return buffers.make_inject(theta_draw, reset_state, reset_draws)

    # -- hot path ---------------------------------------------------------------------
    def reset(self, env_mask=None, reset_state: dict | None = None, out_struct=None, reset_draws=None) -> dict:
        """MultiAgentGraphEnv.reset for every masked world (environment.py:1066-1081).
        ``reset_draws``: float64 (B, K) uniform [0,1) numbers consumed instead of the Philox stream
        (the reference's own np.random stream in the reset-parity tests)."""
        inj, keep = self._inject(None, reset_state, reset_draws)
        m = None
        if env_mask is not None:
            m = torch.as_tensor(env_mask, device=self.device).to(torch.uint8).contiguous()
        rc = self.lib.gmpe_reset(self._h, C.byref(self._st), m.data_ptr() if m is not None else None,
                                 C.byref(inj), C.byref(self._out if out_struct is None else out_struct), self._stream())
        self._check(rc, "gmpe_reset")
        self._keep = (keep, m)
        return self.out

    def redirect(self, **tensors) -> "abi.GmpeStepOut":
        """A GmpeStepOut that writes the named outputs into caller tensors (same shape / dtype,
        contiguous, on this device) instead of the engine's arena -- e.g. slot t+1 of a device
        rollout buffer, so that a step lands where the learner reads it with no copy at all.
        Pass the result as ``out_struct=`` to :meth:`step` / :meth:`reset`."""
        merged = dict(self.out)
        sh = buffers.shapes(self.cfg)
        kinds = {"real": self.real, "i32": torch.int32, "u8": torch.uint8, "f32": torch.float32}
        for k, t in tensors.items():
            shape, dtype = tuple(sh[k][0]), kinds[sh[k][1]]
            if tuple(t.shape) != shape or t.dtype != dtype or t.device != self.device or not t.is_contiguous():
                raise ValueError(f"redirect({k}): need a contiguous {shape} {dtype} tensor on {self.device}")
            merged[k] = t
        s = buffers.make_struct(abi.GmpeStepOut, merged)
        s._keepalive = merged
        return s

    def step(self, action_idx: torch.Tensor, theta_draw=None, reset_state: dict | None = None, out_struct=None) -> dict:
        """One env step for the whole batch. ``action_idx``: int32 (B, N) device tensor."""
        if action_idx.dtype != torch.int32 or action_idx.device != self.device or not action_idx.is_contiguous():
            action_idx = action_idx.to(device=self.device, dtype=torch.int32).contiguous()
        if tuple(action_idx.shape) != (self.B, self.N):
            raise ValueError(f"action_idx must have shape {(self.B, self.N)}, got {tuple(action_idx.shape)}")
        inj, keep = self._inject(theta_draw, reset_state)
        rc = self.lib.gmpe_step(self._h, C.byref(self._st), action_idx.data_ptr(), C.byref(inj),
                                C.byref(self._out if out_struct is None else out_struct), self._stream())
        self._check(rc, "gmpe_step")
        self._keep = (keep, action_idx)
        return self.out

    def step_range(self, action_idx: torch.Tensor, begin: int, count: int, stream: torch.cuda.Stream | None = None,
                   onehot: torch.Tensor | None = None) -> None:
        """The step for worlds ``[begin, begin + count)`` on ``stream`` (whole-batch tensors;
        ``onehot`` -- the (B, N, A) device block of the runner's one-hot actions -- is decoded
        into ``action_idx`` first when given).  Slices are independent: see gmpe_step_range."""
        s = C.c_void_p((stream or torch.cuda.current_stream(self.device)).cuda_stream)
        if onehot is not None:
            rc = self.lib.gmpe_decode_onehot_range(self._h, onehot.data_ptr(), onehot.element_size(), action_idx.data_ptr(),
                                                   begin, count, s)
            self._check(rc, "gmpe_decode_onehot_range")
        rc = self.lib.gmpe_step_range(self._h, C.byref(self._st), action_idx.data_ptr(), C.byref(self._null_inject),
                                      C.byref(self._out), begin, count, s)
        self._check(rc, "gmpe_step_range")

    def step_host(self, actions_host: torch.Tensor, actions_pinned: torch.Tensor, actions_dev: torch.Tensor,
                  host_out: "abi.GmpeStepOut", slices: int) -> None:
        """One host-facing step (gmpe_step_host): ``actions_host`` -- a CPU tensor, int32 (B, N) indices
        or float32 / float64 (B, N, A) one-hot -- is staged, sent up, decoded and stepped slice by slice,
        and every output that ``host_out`` names is copied back into its pinned tensor; copies and
        kernels of neighbouring slices overlap.  Returns once enqueued: call :meth:`host_wait`."""
        kind = 0 if actions_host.dtype == torch.int32 else actions_host.element_size()
        rc = self.lib.gmpe_step_host(self._h, C.byref(self._st), actions_host.data_ptr(), kind, actions_pinned.data_ptr(),
                                     actions_dev.data_ptr(), self.action_idx.data_ptr(), C.byref(self._out), C.byref(host_out),
                                     int(slices), self._stream())
        self._check(rc, "gmpe_step_host")
        self._keep = (actions_host, actions_pinned, actions_dev, host_out)

    def host_wait(self) -> None:
        self._check(self.lib.gmpe_host_wait(self._h), "gmpe_host_wait")

    def decode_onehot(self, onehot: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """argmax of the runner's one-hot actions (environment.py:431,444) on device."""
        A = self.cfg.total_actions
        if tuple(onehot.shape) != (self.B, self.N, A):
            raise ValueError(f"one-hot actions must have shape {(self.B, self.N, A)}, got {tuple(onehot.shape)}")
        if onehot.dtype not in (torch.float32, torch.float64):
            onehot = onehot.to(torch.float32)
        onehot = onehot.to(self.device).contiguous()
        out = self.action_idx if out is None else out
        rc = self.lib.gmpe_decode_onehot(self._h, onehot.data_ptr(), onehot.element_size(), out.data_ptr(), self._stream())
        self._check(rc, "gmpe_decode_onehot")
        self._keep = (onehot,)
        return out

    def gnn_flat_edges(self, capacity: int | None = None, out_struct=None):
        """The consumer-side edge list of the last step in the layout ``GNNBase.forward`` consumes
        (gnn_new.py:492-500 -> process_adj :328-358 on the (B*N, E, E) stack): ``edge_index`` int64
        (2, nnz) with node ids ``(world * N + agent) * E + entity`` in (graph, row, col) order and
        ``edge_attr`` (nnz, 1).  Built on the device from the step kernel's per-world list (scan +
        gather); reading ``nnz`` is the host synchronisation, as in ``Tensor.nonzero`` (with
        ``capacity=None`` the buffers are sized from the sum of the per-world counts first).
        Needs ``gnn_edges=True``.  The results are views of buffers the engine reuses."""
        if "gnn_edge_index" not in self.out:
            raise abi.GmpeError("gnn_flat_edges needs the consumer-side edge list: make_config(..., gnn_edges=True)")
        src = self.out if out_struct is None else getattr(out_struct, "_keepalive", self.out)
        if capacity is None:  # exact: the per-world counts are on the device already (one small reduction + read)
            capacity = max(1, int(src["gnn_num_edges"].sum(dtype=torch.int64).item()) * self.N)
        capacity = int(capacity)
        buf = getattr(self, "_flat", None)
        if buf is None or buf[0].shape[1] < capacity:
            with torch.cuda.device(self.device):
                n_scratch = int(self.lib.gmpe_gnn_flat_scratch_elems(self._h))
                buf = (torch.empty((2, capacity), dtype=torch.int64, device=self.device),
                       torch.empty((capacity,), dtype=self.real, device=self.device),
                       torch.zeros((1,), dtype=torch.int64, device=self.device),
                       torch.empty((n_scratch,), dtype=torch.int64, device=self.device))
            self._flat = buf
        ei, ea, total, scratch = buf
        rc = self.lib.gmpe_gnn_flatten(self._h, C.byref(self._out if out_struct is None else out_struct), ei.data_ptr(),
                                       ea.data_ptr(), ei.shape[1], total.data_ptr(), scratch.data_ptr(), self._stream())
        self._check(rc, "gmpe_gnn_flatten")
        nnz = int(total.item())
        if nnz > ei.shape[1]:
            raise abi.GmpeError(f"gnn_flat_edges: {nnz} edges do not fit the capacity of {ei.shape[1]}")
        return ei[:, :nnz], ea[:nnz].unsqueeze(1)

    # -- checkpoint / teacher forcing ---------------------------------
    def state_dict(self) -> dict:
        return {k: v.clone() for k, v in self.state.items()}

    def load_state_dict(self, sd: dict) -> None:
        for k, v in sd.items():
            t = torch.as_tensor(v, device=self.device)
            self.state[k].copy_(t.to(self.state[k].dtype).reshape(self.state[k].shape))

    # -- roofline bookkeeping -------------------------------------------------------------
    def algorithmic_bytes_per_env_step(self) -> int:
        """Bytes one env-step must move if every logical array is touched exactly once
        (after SURVEY.md section 8d, extended by the info arrays this build keeps)."""
        c = self.cfg
        R = self.layout.real_bytes
        N, L, O, W, E = c.num_agents, c.num_landmarks, c.num_obstacles, c.num_walls, self.E
        F, cap = self.layout.obs_dim, self.layout.edge_cap
        b = 2 * (4 * R * N)            # agent_state r+w
        b += 2 * (2 * R * N)           # agent_aux r+w
        b += 2 * (4 * N)               # agent_flags r+w
        b += 4 * N                     # action index in
        b += R * N + N                 # reward + done out
        b += 2 * 4                     # step counter r+w
        if c.scenario != abi.SCEN_NAV_GRAPH:
            b += R * abi.WORLD_F       # tube frame in (navigation_graph has none: world_f is not read there)
        b += 2 * R * (L + O) + abi.WALL_F * R * W
        b += R * F * N + R * self.layout.node_dim * N * E + R * E * E
        if c.flags & abi.FLAG_INFO:
            b += 2 * (4 * abi.AGENT_CNT * N) + R * abi.INFO_K * N
        if c.flags & abi.FLAG_ENV_EDGES:
            b += (8 + R) * cap + 4
        if c.flags & abi.FLAG_GNN_EDGES:
            b += (8 + R) * cap + 4
        return int(b)
