"""Environment metrics as device reductions.

The reference walks ``infos[env][agent][key]`` in Python for every key
(``Runner.process_infos``, onpolicy/runner/shared/base_runner.py:194-284) and logs
``np.mean`` of each list (``log_env``, :292-302).  With the info rows already on the GPU as a
``(B, N, 16)`` tensor that is one masked column mean per (agent, key): a ``(N, 17)`` result,
reduced across ranks with a single sum all-reduce (NCCL on GPUs, gloo in the CPU tests).
torch does the reduction -- this is bookkeeping off the step path, run once per log interval.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import abi

#: log name suffix per column, in the order of base_runner.py:262-279 (column 0 = individual_reward)
LOG_NAMES = {
    "individual_reward": "individual_rewards", "Time_req_to_goal": "time_to_goal", "Min_time_to_goal": "min_time_to_goal",
    "Dist_to_goal": "dist_to_goal", "Num_agent_collisions": "num_agent_collisions",
    "Num_obst_collisions": "num_obstacle_collisions", "Distance_mean": "distance_mean",
    "Distance_variance": "distance_variance", "Mean_by_variance": "mean_variance", "Dists_traveled": "dists_traveled",
    "Time_taken": "time_taken", "Time_mean": "time_mean", "Time_stddev": "time_variance",
    "Time_mean_by_stddev": "time_mn_by_stddev", "Conformance": "conformance_percentages", "Delta_spacing": "delta_space",
    "Spacing_violations": "spacing_violations",
}
COLUMNS = ("individual_reward",) + abi.INFO_KEYS


def info_sums(info: torch.Tensor, indiv_reward: torch.Tensor, episode_length: int, dt: float) -> torch.Tensor:
    """Per-(agent, key) sums over the worlds of this rank plus the world count: float64
    ``(N * 17 + 1,)``.  ``Time_req_to_goal == -1`` counts as ``episode_length * dt``
    (base_runner.py:214-217)."""
    x = torch.cat([indiv_reward.unsqueeze(-1), info], dim=-1).to(torch.float64)  # (B, N, 17)
    j = 1 + abi.INFO_KEYS.index("Time_req_to_goal")
    t = x[..., j]
    x[..., j] = torch.where(t == -1, torch.full_like(t, float(episode_length) * float(dt)), t)
    sums = x.sum(dim=0).reshape(-1)
    return torch.cat([sums, torch.tensor([float(x.shape[0])], dtype=torch.float64, device=x.device)])


def process_infos(info: torch.Tensor, indiv_reward: torch.Tensor, episode_length: int, dt: float, reduce: bool = True) -> dict:
    """``{f'agent{a}/{name}': mean over all worlds}`` -- what ``log_env(process_infos(infos))``
    hands to the logger (keys with no data, e.g. ``formation_dist``, are absent there too)."""
    N = info.shape[1]
    s = info_sums(info, indiv_reward, episode_length, dt)
    if reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    means = (s[:-1] / s[-1]).reshape(N, len(COLUMNS)).cpu()
    return {f"agent{a}/{LOG_NAMES[k]}": float(means[a, c]) for a in range(N) for c, k in enumerate(COLUMNS)}
