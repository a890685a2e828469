"""TEST INFRASTRUCTURE: numpy restatement of the env-facing part of the reference's rollout
glue -- GMPERunner.warmup / insert (onpolicy/runner/shared/graph_mpe_runner.py:213-239,
384-427) writing into GraphReplayBuffer (onpolicy/utils/graph_buffer.py:84-165, 222-287).
Pinned against tests/golden/rollout_glue_n3.npz (made from the unmodified reference code by
tests/golden/make_rollout_golden.py) in tests/test_rollout_glue.py."""
import numpy as np


class RolloutRef:
    def __init__(self, T, B, N, E, F, A, use_centralized_V=True):
        self.T, self.B, self.N, self.cv = T, B, N, use_centralized_V
        f32 = np.float32
        self.obs = np.zeros((T + 1, B, N, F), f32)
        self.share_obs = np.zeros((T + 1, B, N, N * F if use_centralized_V else F), f32)
        self.node_obs = np.zeros((T + 1, B, N, E, 8), f32)
        self.adj = np.zeros((T + 1, B, N, E, E), f32)
        self.agent_id = np.zeros((T + 1, B, N, 1), np.int32)
        self.share_agent_id = np.zeros((T + 1, B, N, N if use_centralized_V else 1), np.int32)
        self.rewards = np.zeros((T, B, N, 1), f32)
        self.masks = np.ones((T + 1, B, N, 1), f32)
        self.active_masks = np.ones_like(self.masks)
        self.bad_masks = np.ones_like(self.masks)
        self.available_actions = np.ones((T + 1, B, N, A), f32)
        self.step = 0

    def _share(self, obs, ids):
        if not self.cv:
            return obs, ids
        so = np.expand_dims(obs.reshape(self.B, -1), 1).repeat(self.N, axis=1)      # runner :407-415
        si = np.expand_dims(ids.reshape(self.B, -1), 1).repeat(self.N, axis=1)      # runner :416-419
        return so, si

    def warmup(self, obs, ids, node_obs, adj):                                        # runner :213-239
        so, si = self._share(obs, ids)
        self.share_obs[0], self.obs[0], self.node_obs[0], self.adj[0] = so, obs, node_obs, adj
        self.agent_id[0], self.share_agent_id[0] = ids, si

    def insert(self, obs, ids, node_obs, adj, rewards, dones):                        # runner :384-427
        dones_env = np.all(dones, axis=1)
        masks = np.ones((self.B, self.N, 1), np.float32)
        masks[dones] = 0
        active = np.ones((self.B, self.N, 1), np.float32)
        active[dones] = 0
        active[dones_env] = 1
        so, si = self._share(obs, ids)
        t = self.step                                                                  # buffer :222-250
        self.share_obs[t + 1], self.obs[t + 1], self.node_obs[t + 1], self.adj[t + 1] = so, obs, node_obs, adj
        self.agent_id[t + 1], self.share_agent_id[t + 1] = ids, si
        self.rewards[t] = rewards
        self.masks[t + 1], self.active_masks[t + 1] = masks, active
        self.step = (t + 1) % self.T

    def after_update(self):                                                            # buffer :254-287
        for k in ("share_obs", "obs", "node_obs", "adj", "agent_id", "share_agent_id", "masks", "active_masks",
                  "bad_masks", "available_actions"):
            getattr(self, k)[0] = getattr(self, k)[-1].copy()
