"""DeviceGraphRollout -- the env-facing half of the reference's rollout storage, on the GPU.

SURVEY.md section 8(f) rank 1: with a GPU-resident env the host runner becomes the bottleneck
at once.  The reference copies every step's tensors four times on the host
(``GraphSubprocVecEnv.step_wait`` np.stack, ``GMPERunner.insert`` reshapes,
``GraphReplayBuffer.insert`` ``.copy()`` into fp32 storage, ``collect`` np.concatenate;
onpolicy/envs/env_wrappers.py:996-1004, onpolicy/runner/shared/graph_mpe_runner.py:384-427,
onpolicy/utils/graph_buffer.py:168-252, graph_mpe_runner.py:343-356).  Here the step kernel
writes obs / node_obs / adj / reward / done **directly into slot t+1** of device tensors that
have the buffer's layout (the C ABI takes output pointers, so this is a pointer swap, not a
copy), and the runner's mask bookkeeping (masks / active_masks) comes out of the same kernel,
so one launch is all a rollout step costs.

Fields mirror ``GraphReplayBuffer`` (graph_buffer.py:84-165), env-produced ones only:
``obs, share_obs, node_obs, adj, agent_id, share_agent_id, rewards, masks, active_masks,
bad_masks, available_actions``.  What the policy produces (values, log-probs, rnn states)
belongs to the learner and is not stored here.  Replicated tensors are stride-0 views:
``adj`` is one (E, E) matrix per world (the reference returns the same ndarray for all N
agents) and ``share_obs`` is the world's concatenated obs row.
"""
from __future__ import annotations

import torch

from . import abi
from .engine import BatchedGraphMPE


class DeviceGraphRollout:
    def __init__(self, engine: BatchedGraphMPE, episode_length: int, use_centralized_V: bool = True):
        self.engine, self.T = engine, int(episode_length)
        cfg = engine.cfg
        B, N, E = engine.B, engine.N, engine.E
        F = engine.layout.obs_dim
        dev, real = engine.device, engine.real
        T1 = self.T + 1
        self.use_centralized_V = bool(use_centralized_V)
        self.collaborative = bool(cfg.flags & abi.FLAG_COLLABORATIVE)
        # storage the step kernel writes into (slot t+1 of each)
        self._obs = torch.zeros((T1, B, N, F), dtype=real, device=dev)
        self._node_obs = torch.zeros((T1, B, N, E, engine.layout.node_dim), dtype=real, device=dev)
        self._adj = torch.zeros((T1, B, E, E), dtype=real, device=dev)
        self._reward = torch.zeros((self.T, B, N), dtype=real, device=dev)
        self._done = torch.zeros((self.T, B, N), dtype=torch.uint8, device=dev)
        # runner bookkeeping (graph_mpe_runner.py:384-405)
        self.masks = torch.ones((T1, B, N, 1), dtype=torch.float32, device=dev)
        self.active_masks = torch.ones_like(self.masks)
        self.bad_masks = torch.ones_like(self.masks)
        self.available_actions = torch.ones((T1, B, N, cfg.total_actions), dtype=torch.float32, device=dev)
        ids = torch.arange(N, dtype=torch.int32, device=dev)
        self.agent_id = ids.view(1, 1, N, 1).expand(T1, B, N, 1)
        self.share_agent_id = (ids.view(1, 1, 1, N).expand(T1, B, N, N) if use_centralized_V else self.agent_id)
        # masks / active_masks are written by the step kernel too (GmpeStepOut.mask / active_mask)
        self._structs = [engine.redirect(obs=self._obs[t + 1], node_obs=self._node_obs[t + 1], adj=self._adj[t + 1],
                                         reward=self._reward[t], done=self._done[t], mask=self.masks[t + 1].view(B, N),
                                         active_mask=self.active_masks[t + 1].view(B, N)) for t in range(self.T)]
        self._reset_struct = engine.redirect(obs=self._obs[0], node_obs=self._node_obs[0], adj=self._adj[0])
        self.step = 0

    # -- buffer views in the reference's shapes ----------------------------------------------
    @property
    def obs(self):            # (T+1, B, N, F)
        return self._obs

    @property
    def share_obs(self):      # (T+1, B, N, N*F) | obs  (graph_mpe_runner.py:407-422)
        if not self.use_centralized_V:
            return self._obs
        T1, B, N, F = self._obs.shape
        return self._obs.view(T1, B, 1, N * F).expand(T1, B, N, N * F)

    @property
    def node_obs(self):       # (T+1, B, N, E, 8)
        return self._node_obs

    @property
    def adj(self):            # (T+1, B, N, E, E), stride 0 over N
        T1, B, E, _ = self._adj.shape
        return self._adj.view(T1, B, 1, E, E).expand(T1, B, self.engine.N, E, E)

    @property
    def rewards(self):        # (T, B, N, 1)  (runner: rewards[:, :, np.newaxis])
        return self._reward.unsqueeze(-1)

    @property
    def dones(self):          # (T, B, N) bool
        return self._done.bool()

    # -- runner protocol -------------------------------------------------------------------------
    def warmup(self):
        """GMPERunner.warmup (graph_mpe_runner.py:213-239): reset the envs into slot 0."""
        self.engine.reset(out_struct=self._reset_struct)
        self.step = 0
        return self._obs[0], self.agent_id[0], self._node_obs[0], self.adj[0]

    def insert_env_step(self, action_idx: torch.Tensor):
        """envs.step + the env-produced part of GMPERunner.insert / GraphReplayBuffer.insert
        (graph_mpe_runner.py:82-113, 384-427; graph_buffer.py:222-250) for step ``self.step``.
        ``action_idx``: int32 (B, N) on the device (the runner's argmax'd actions)."""
        t = self.step
        self.engine.step(action_idx, out_struct=self._structs[t])
        self.step = (t + 1) % self.T
        return self._obs[t + 1], self.agent_id[t + 1], self._node_obs[t + 1], self.adj[t + 1], \
            self.rewards[t], self._done[t]

    def after_update(self):
        """GraphReplayBuffer.after_update (graph_buffer.py:254-287): last slot -> slot 0."""
        for buf in (self._obs, self._node_obs, self._adj, self.masks, self.active_masks, self.bad_masks,
                    self.available_actions):
            buf[0].copy_(buf[-1])
