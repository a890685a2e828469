"""Vectorised env wrappers with the reference's interface
(onpolicy/envs/env_wrappers.py:29-140, 851-1037), backed by ONE GPU-resident batch.

``GraphSubprocVecEnv(env_fns)`` / ``GraphDummyVecEnv(env_fns)`` take the same list of env
constructors the reference's ``make_train_env`` builds (onpolicy/scripts/train_mpe.py:21-43).
Only ``env_fns[0]`` is called (it yields the args, spaces and seed); ``len(env_fns)`` is the
number of worlds B.  Instead of B processes exchanging pickles over pipes, ``step`` is: one
host->device copy of the action block, two kernel launches (one-hot argmax, fused step incl.
auto-reset), one device->host copy of the whole output arena into pinned memory.

Returned arrays follow the reference's shapes:
  reset -> obs (B,N,F), ag_ids (B,N,1), node_obs (B,N,E,8), adj (B,N,E,E)
  step  -> (..., rews (B,N) or (B,N,1) when collaborative, dones (B,N) bool, infos)
           GraphDummyVecEnv appends ``reset_count`` as an 8th value.
Differences, all documented in INTEGRATION.md: arrays are float32 (the only consumer casts to
float32, onpolicy/utils/graph_buffer.py:84-113); they are views of a double-buffered pinned
arena, valid until the step after next; ``adj`` is a stride-0 broadcast of the single
per-world matrix the reference aliases N times; ``infos`` is a lazy sequence.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import os

import numpy as np
import torch

from . import abi, buffers
from .config import make_config
from .engine import BatchedGraphMPE
from .environment import build_spaces


class ShareVecEnv(ABC):  # env_wrappers.py:29-140
    closed = False
    viewer = None
    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, num_envs, observation_space, share_observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.share_observation_space = share_observation_space
        self.action_space = action_space

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_async(self, actions):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    def close_extras(self):
        pass

    def close(self):
        if self.closed:
            return
        self.close_extras()
        self.closed = True

    def step(self, actions, num_current_episode: Optional[int] = None):
        self.step_async(actions, num_current_episode)
        return self.step_wait()

    def render(self, mode="human"):
        raise NotImplementedError("rendering is out of scope of this build")

    @property
    def unwrapped(self):
        return self


class _AgentInfo:
    """dict-like view of one (env, agent) row of the info tensor (``phase``: the rot-inv family's
    'Phase_reached' entry, three_phase_graph.py:806; None for the scenarios without that key)."""
    __slots__ = ("_row", "_rew", "_phase")

    def __init__(self, row, rew, phase=None):
        self._row, self._rew, self._phase = row, rew, phase

    def keys(self):
        return ("individual_reward",) + (abi.INFO_KEYS if self._phase is None else abi.INFO_KEYS_ROT_INV)

    def __contains__(self, k):
        return k in self.keys()

    def __getitem__(self, k):
        if k == "individual_reward":
            return self._rew
        if k == "Phase_reached" and self._phase is not None:
            return self._phase
        return self._row[abi.INFO_KEYS.index(k)]

    def get(self, k, default=None):
        return self[k] if k in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def to_dict(self):
        return dict(self.items())


class LazyInfos:
    """``infos[env][agent][key]`` as GMPERunner.process_infos walks it
    (onpolicy/runner/shared/base_runner.py:194-284) without building B*N python dicts.

    With ``fetch`` the info rows are still on the GPU: the runner reads infos only when it logs
    (graph_mpe_runner.py:179-195), so the 20 % of a step's device->host bytes they account for
    are copied on first access instead of every step.  They must be read before the env steps
    again (the device rows are overwritten); a later access raises instead of returning the
    wrong step's values."""

    def __init__(self, info: np.ndarray = None, indiv_reward: np.ndarray = None, fetch=None, phase_reached: np.ndarray = None):
        self._info, self._rew, self._fetch, self._phase = info, indiv_reward, fetch, phase_reached

    def _load(self):
        if self._info is None:
            self._info, self._rew, self._phase = self._fetch()
            self._fetch = None

    @property
    def phase_reached(self):
        """(B, N) uint8 'Phase_reached' of the rot-inv family, None for the other scenarios."""
        self._load()
        return self._phase

    @property
    def info(self) -> np.ndarray:
        self._load()
        return self._info

    @property
    def rew(self) -> np.ndarray:
        self._load()
        return self._rew

    def __len__(self):
        return self.info.shape[0]

    def __getitem__(self, e):
        if isinstance(e, slice):
            return [self[i] for i in range(*e.indices(len(self)))]
        if e < 0:
            e += len(self)
        ph = self.phase_reached
        return [_AgentInfo(self.info[e, a], self.rew[e, a], None if ph is None else ph[e, a]) for a in range(self.info.shape[1])]

    def __iter__(self):
        return (self[e] for e in range(len(self)))


class _BatchedVecEnv(ShareVecEnv):
    returns_reset_count = False

    def __init__(self, env_fns=None, spaces=None, *, args=None, num_envs=None, device=0, precision="fp32",
                 seed=None, world_offset=0, output="numpy", info=True, env_edges=False, gnn_edges=False,
                 pipeline_chunks=None, lazy_info=True):
        if args is None:
            if not env_fns:
                raise ValueError("need env_fns (reference calling convention) or args=/num_envs=")
            env0 = env_fns[0]()
            args = env0.world.all_args
            if seed is None:
                seed = getattr(env0, "_seed", None)
            env0.close()
            num_envs = len(env_fns)
        if seed is None:
            seed = getattr(args, "seed", 1)
        self.all_args = args
        self.cfg = make_config(args, num_envs, precision=precision, seed=seed, world_offset=world_offset,
                               auto_reset=True, env_edges=env_edges, gnn_edges=gnn_edges, info=info)
        self.engine = BatchedGraphMPE(self.cfg, device)
        sp = build_spaces(args, self.cfg)
        ShareVecEnv.__init__(self, num_envs, sp["observation_space"], sp["share_observation_space"], sp["action_space"])
        for k in ("node_observation_space", "adj_observation_space", "edge_observation_space",
                  "agent_id_observation_space", "share_agent_id_observation_space"):
            setattr(self, k, sp[k])
        self.output = output
        self.waiting = False
        self.closed = False
        B, N, A = num_envs, self.cfg.num_agents, self.cfg.total_actions
        self.num_agents = N
        self._A = A
        self._ag_ids = np.broadcast_to(np.arange(N, dtype=np.int64).reshape(1, N, 1), (B, N, 1))
        dev = self.engine.device
        # pinned staging: action block in (both element widths the runner may hand over) and
        # a double-buffered arena out
        self._act_pin = {4: torch.empty((B, N, A), dtype=torch.float32).pin_memory(),
                         8: torch.empty((B, N, A), dtype=torch.float64).pin_memory()}
        self._act_dev = {4: torch.empty((B, N, A), dtype=torch.float32, device=dev),
                         8: torch.empty((B, N, A), dtype=torch.float64, device=dev)}
        self._idx_pin = torch.empty((B, N), dtype=torch.int32).pin_memory()
        nbytes = self.engine.host_prefix_bytes
        slabs = self.engine.out_slabs
        self._lazy_info = bool(lazy_info) and output == "numpy" and "info" in slabs
        #: bytes of the arena a step always brings to the host (obs, node_obs, adj, reward, done)
        self._core_bytes = slabs["done"][0] + slabs["done"][1] if self._lazy_info else nbytes
        self._gen = 0
        self._arena = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._flip = 0
        self.h2d_bytes_last = 0
        self.d2h_bytes_last = nbytes
        # Host-facing steps are PCIe-bound (one-hot block in, every tensor out): 7.9 MB up and 52.9 MB
        # down per step at 65,536 navigation worlds.  Measured on the B200 boxes
        # (profiles/r2_e2e_timeline.txt, profiles/r2_e2e_sweep.txt): copies in opposite directions that
        # overlap slow each other down badly (device->host 50 -> 19 GB/s, host->device 8 GB/s), and
        # short per-slice copies reach 33-40 GB/s -- so the default is the plain sequence: one
        # threaded staging copy into pinned memory, one upload, decode + step, ONE download of the
        # arena prefix (1.32 ms against a 1.02 ms copy-only ceiling).  pipeline_chunks > 1 routes the
        # step through gmpe_step_host instead (staging / upload / kernels pipelined in that many blocks
        # by the library's own helper threads; what a non-Python host would call).
        if pipeline_chunks is None:
            pipeline_chunks = 1
        # device-resident outputs have nothing to copy back: one launch, and no host copies in
        # flight that a later launch could overwrite
        self._chunks = max(1, int(pipeline_chunks)) if output == "numpy" else 1
        #: recorded behind every non-blocking upload from a pinned action block; the block is not
        #: rewritten before the copy has executed (with output="torch" nothing else makes the host wait)
        self._h2d_done = torch.cuda.Event()
        self._h2d_pending = False
        self._host_names = [n for n in ("obs", "node_obs", "adj", "reward", "done", "indiv_reward", "info", "phase_reached")
                            if n in self.engine.out_slabs and not (self._lazy_info and n in ("indiv_reward", "info", "phase_reached"))]
        self._host_t = [self._host_tensors(a) for a in self._arena]
        #: numpy views of the two pinned arenas, built once (a step hands one of them out)
        self._views = {a.data_ptr(): self._host_views(a) for a in self._arena}
        # the same pinned views as a GmpeStepOut of host pointers, for gmpe_step_host
        self._host_struct = [buffers.make_struct(abi.GmpeStepOut, t) for t in self._host_t]
        self._pending = None

    # -- host views of the pinned arena ---------------------------------------------------
    def _host_views(self, arena: torch.Tensor) -> dict:
        np_arena = arena.numpy()
        eng = self.engine
        real = np.float64 if self.cfg.precision == abi.F64 else np.float32
        kinds = {"obs": real, "node_obs": real, "adj": real, "reward": real, "indiv_reward": real, "info": real,
                 "done": np.uint8, "phase_reached": np.uint8}
        out = {}
        for n, (off, nb) in eng.out_slabs.items():
            if n in kinds and off + nb <= np_arena.size:
                out[n] = np_arena[off:off + nb].view(kinds[n]).reshape(tuple(eng.out[n].shape))
        return out

    def _host_tensors(self, arena: torch.Tensor) -> dict:
        """torch views of a pinned arena, one per host-visible output (same offsets as the device arena)."""
        eng = self.engine
        out = {}
        for n in self._host_names:
            off, nb = eng.out_slabs[n]
            ref = eng.out[n]
            out[n] = arena[off:off + nb].view(ref.dtype).view(ref.shape)
        return out

    def _step_pipelined(self, src: torch.Tensor, pinned_actions: torch.Tensor, dev_actions: torch.Tensor):
        """The batch through ONE library call (gmpe_step_host): the caller's action block (``src``,
        pageable) is staged into pinned memory, sent up and stepped block by block, then every
        host-visible field is copied back.  The copies of the results are in flight on return."""
        self.engine.step_host(src, pinned_actions, dev_actions, self._host_struct[self._flip], self._chunks)
        self._pending = self._arena[self._flip]
        self._flip ^= 1

    def _pinned_reusable(self):
        """Block until the last upload from the pinned action blocks has executed."""
        if self._h2d_pending:
            self._h2d_done.synchronize()
            self._h2d_pending = False

    def _uploaded(self):
        self._h2d_done.record(torch.cuda.current_stream(self.engine.device))
        self._h2d_pending = True

    def _drain(self):
        """Finish a pipelined host step whose results nobody collected (reset() after step_async)."""
        if self._pending is not None:
            self.engine.host_wait()
            self._pending = None
        self.waiting = False

    def _fetch(self):
        if self.output == "torch":
            return dict(self.engine.out)
        if self._pending is not None:  # a pipelined step already has its copies in flight
            arena, self._pending = self._pending, None
            self.engine.host_wait()
            self._last_arena = arena
            return self._views[arena.data_ptr()]
        arena = self._arena[self._flip]
        self._flip ^= 1
        n = self._core_bytes
        arena[:n].copy_(self.engine.out_arena[:n], non_blocking=True)
        torch.cuda.current_stream(self.engine.device).synchronize()
        self._last_arena = arena
        return self._views[arena.data_ptr()]

    def _info_fetcher(self, arena: torch.Tensor):
        """Deferred device->host copy of this step's info rows (LazyInfos)."""
        gen, eng = self._gen, self.engine

        def fetch():
            if gen != self._gen:
                raise RuntimeError("infos of an earlier step were read after the env stepped again; read them before "
                                   "the next step() or construct the vec env with lazy_info=False")
            lo, hi = self._core_bytes, arena.numel()
            arena[lo:hi].copy_(eng.out_arena[lo:hi], non_blocking=True)
            torch.cuda.current_stream(eng.device).synchronize()
            self.d2h_bytes_last += hi - lo
            v = self._views[arena.data_ptr()]
            return v["info"], v["indiv_reward"], v.get("phase_reached")
        return fetch

    def _pack_obs(self, o):
        B, N, E = self.num_envs, self.num_agents, self.engine.E
        if self.output == "torch":
            adj = o["adj"].unsqueeze(1).expand(B, N, E, E)
            ids = torch.as_tensor(np.ascontiguousarray(self._ag_ids), device=self.engine.device)
            return o["obs"], ids, o["node_obs"], adj
        adj = np.broadcast_to(o["adj"][:, None], (B, N, E, E))
        return o["obs"], self._ag_ids, o["node_obs"], adj

    # -- reference interface ------------------------------------------------------------
    def reset(self, num_current_episode: int = 0):
        self._drain()   # a step still in flight must not have its arena handed out as the reset's
        self._gen += 1  # infos of earlier steps can no longer be fetched
        self.engine.reset()
        return self._pack_obs(self._fetch())

    def step_async(self, actions, num_current_episode=None):
        eng = self.engine
        self._gen += 1
        self.d2h_bytes_last = self._core_bytes
        B, N, A = self.num_envs, self.num_agents, self._A
        if isinstance(actions, torch.Tensor) and actions.is_cuda:
            if actions.dtype in (torch.int32, torch.int64) and tuple(actions.shape) == (B, N):
                idx = actions.to(torch.int32)
                self.h2d_bytes_last = 0
            else:
                idx = eng.decode_onehot(actions)
                self.h2d_bytes_last = 0
        else:
            a = np.asarray(actions)
            if a.shape == (B, N) and np.issubdtype(a.dtype, np.integer):  # index fast path
                self._pinned_reusable()
                self._idx_pin.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)))
                self.h2d_bytes_last = self._idx_pin.numel() * 4
                eng.action_idx.copy_(self._idx_pin, non_blocking=True)
                self._uploaded()
                idx = eng.action_idx
            else:
                if a.shape != (B, N, A):
                    raise ValueError(f"actions must have shape {(B, N, A)} (one-hot) or {(B, N)} (int), got {a.shape}")
                w = 8 if a.dtype == np.float64 else 4
                if a.dtype not in (np.float64, np.float32):
                    a = a.astype(np.float32)
                src = torch.from_numpy(np.ascontiguousarray(a))
                self.h2d_bytes_last = a.size * w
                if self._chunks > 1:
                    self._step_pipelined(src, self._act_pin[w], self._act_dev[w])
                    self.waiting = True
                    return
                # torch's threaded copy fills the pinned block ~15x faster than a numpy assignment
                self._pinned_reusable()
                self._act_pin[w].copy_(src)
                self._act_dev[w].copy_(self._act_pin[w], non_blocking=True)
                self._uploaded()
                idx = eng.decode_onehot(self._act_dev[w])
        eng.step(idx)
        self.waiting = True

    def step_wait(self):
        o = self._fetch()
        self.waiting = False
        obs, ids, node_obs, adj = self._pack_obs(o)
        collab = bool(self.cfg.flags & abi.FLAG_COLLABORATIVE)
        if self.output == "torch":
            rews = o["reward"].unsqueeze(-1) if collab else o["reward"]
            dones = o["done"].bool()
            infos = None
            if "info" in o:
                infos = {"info": o["info"], "individual_reward": o["indiv_reward"]}
                if "phase_reached" in o:
                    infos["phase_reached"] = o["phase_reached"]
        else:
            rews = o["reward"][..., None] if collab else o["reward"]  # [[sum]]*N -> (B,N,1), environment.py:1059
            dones = o["done"].view(np.bool_)
            if self._lazy_info:
                infos = LazyInfos(fetch=self._info_fetcher(self._last_arena))
            else:
                infos = LazyInfos(o["info"], o["indiv_reward"], phase_reached=o.get("phase_reached")) if "info" in o else tuple([{}] * 0)
        if self.returns_reset_count:
            # env_wrappers.py:925-936: reset_count = 1 when any env auto-reset this step
            if self.output == "torch":
                rc = int(bool(dones.all(dim=1).any()))
            else:
                rc = int(bool(dones.all(axis=1).any()))
            return obs, ids, node_obs, adj, rews, dones, infos, rc
        return obs, ids, node_obs, adj, rews, dones, infos

    def close(self):
        if self.closed:
            return
        self.engine.close()
        self.closed = True

    # -- device-side statistics for the once-per-rollout reduction (dist.py) ---------------
    def episode_stats(self) -> torch.Tensor:
        """[sum reward, #agent collisions, #obstacle collisions, #agents done, #agents] of the
        current step as a float64 device vector (reduced across GPUs by dist.reduce_stats).
        The collision counts are the episode totals so far as this step's info rows report them
        (july:806-829 Num_agent_collisions / Num_obst_collisions): a world that auto-reset this
        step still shows the totals of the episode it just finished, whereas the state counters
        have already been cleared for the new one.  Needs info=True."""
        o = self.engine.out
        if "info" not in o:
            raise RuntimeError("episode_stats needs the info rows: construct the vec env with info=True")
        info = o["info"].to(torch.float64)
        key = "indiv_reward" if "indiv_reward" in o else "reward"
        return torch.stack([o[key].to(torch.float64).sum(), info[..., abi.INFO_KEYS.index("Num_agent_collisions")].sum(),
                            info[..., abi.INFO_KEYS.index("Num_obst_collisions")].sum(), o["done"].to(torch.float64).sum(),
                            torch.tensor(float(self.num_envs * self.num_agents), dtype=torch.float64,
                                         device=self.engine.device)])


class GraphSubprocVecEnv(_BatchedVecEnv):
    """env_wrappers.py:959-1037 -- step_wait returns 7 values."""
    returns_reset_count = False


class GraphDummyVecEnv(_BatchedVecEnv):
    """env_wrappers.py:903-956 -- step_wait returns 8 values (adds reset_count)."""
    returns_reset_count = True
