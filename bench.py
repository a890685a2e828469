#!/usr/bin/env python
"""bench.py -- agent-steps/s of the GraphMPE step path (env step + graph obs).

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host cores)

For N > 1 the driver launches it under torchrun (one rank per GPU, NCCL).  One JSON line is
printed by rank 0.  Workload = BASELINE.json configs[1]: navigation_graph, 3 agents / 3
landmarks, 65,536 worlds per GPU (weak scaling: the worlds of a rank are independent, there
is no per-step collective; one NCCL all-reduce of the episode statistics per rollout).

A "step" is one pass of the hot path over the batch: fused kernel = _set_action + World.step
+ observation / reward / graph_observation / done / info for every agent of every world,
including auto-reset.  `value` times it with inputs resident in HBM (CUDA events around each
step on the launching stream, L2 flushed before every timed step); `e2e` times the same
metric through the reference-facing GraphSubprocVecEnv.step with HOST buffers (the one-hot
float64 action block the runner builds going in, every returned tensor coming back to pinned
host memory), copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1] (the configuration the metric is quoted on)
    "navigation_graph_n3_b65536": dict(
        args=dict(scenario_name="navigation_graph", dynamics_type="double_integrator", num_agents=3,
                  num_landmarks=3, num_obstacles=0, num_walls=0, world_size=2, episode_length=25,
                  total_actions=5, collaborative=True, use_dones=False, goal_rew=5, collision_rew=5, max_speed=2),
        worlds_per_gpu=65536),
    # BASELINE.json configs[2]
    "july_air_taxi_n3_b131072": dict(
        args=dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi",
                  num_agents=3, num_landmarks=3, world_size=3, episode_length=200, goal_rew=30, collision_rew=5,
                  formation_rew=1, collaborative=False, use_dones=False, max_speed=2),
        worlds_per_gpu=131072),
    # the rotation-invariant family (eval_scripts/commands_test_run.txt:37)
    "three_phase_air_taxi_n3_b131072": dict(
        args=dict(scenario_name="three_phase_graph", dynamics_type="air_taxi", num_agents=3, num_landmarks=3,
                  world_size=3, episode_length=75, goal_rew=30, collision_rew=5, formation_rew=1,
                  collaborative=False, use_dones=False, max_speed=2),
        worlds_per_gpu=131072),
    # ... at the shapes of the reference's eval commands (eval_scripts/commands_test_run.txt:7,30)
    "two_phase_air_taxi_n5": dict(
        args=dict(scenario_name="two_phase_graph", dynamics_type="air_taxi", num_agents=5, num_landmarks=5,
                  world_size=3, episode_length=100, goal_rew=20, collision_rew=20, collaborative=False, use_dones=False,
                  max_speed=2),
        worlds_per_gpu=65536),
    "three_phase_air_taxi_n10": dict(
        args=dict(scenario_name="three_phase_graph", dynamics_type="air_taxi", num_agents=10, num_landmarks=10,
                  world_size=4, episode_length=150, goal_rew=20, collision_rew=20, collaborative=False, use_dones=False,
                  max_speed=2),
        worlds_per_gpu=32768),
    # BASELINE.json configs[3]
    "navigation_graph_n16_o8_w4": dict(
        args=dict(scenario_name="navigation_graph", dynamics_type="double_integrator", num_agents=16,
                  num_landmarks=16, num_obstacles=8, num_walls=4, world_size=4, episode_length=25,
                  total_actions=5, collaborative=True, max_speed=2),
        worlds_per_gpu=16384),
    # BASELINE.json configs[4]
    "july_air_taxi_n64": dict(
        args=dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi",
                  num_agents=64, num_landmarks=64, world_size=12, episode_length=200, goal_rew=30,
                  collision_rew=5, formation_rew=1, collaborative=False, max_speed=2),
        worlds_per_gpu=4096),
}
DEFAULT_WORKLOAD = "navigation_graph_n3_b65536"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region.  NVML in-process (one
    cheap library call per sample); falls back to forking nvidia-smi only where pynvml is missing.
    A 10 Hz `nvidia-smi` fork per rank inside the event pairs of a 30 us kernel was the largest
    source of per-step jitter at 8 ranks in round 1."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.samples, self._stop, self._t = gpu_index, [], threading.Event(), None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
        r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        act = lambda k: "Active" if r & bits[k] else "Not Active"  # noqa: E731
        return [str(sm), str(self._max), act("hw_slowdown"), act("hw_thermal_slowdown"), act("sw_thermal_slowdown"),
                act("sw_power_cap")]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    if len(f) >= 6:
                        self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.05 if self._nvml is not None else 0.25)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(s[2 + j].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def pin_rank(local: int, n_local: int):
    """Give every rank of the box its own block of host cores (its staging helpers and copy threads
    then never migrate onto another rank's cores).  Returns the cores or None."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        per = len(allowed) // max(1, n_local)
        if n_local <= 1 or per < 1:
            return None
        mine = allowed[local * per:(local + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


def time_steps_on_device(torch, stream, flush, step_fn, K, gate=True):
    """K steps, each between its own pair of CUDA events on the launching stream, L2 flushed before
    every one.  With `gate` the device is first given a backlog of flush kernels as long as the host
    needs to enqueue the K iterations: the events then bracket the kernel alone -- a host thread
    that is late with a launch (Python, 8 ranks sharing the cores) otherwise shows up as device
    idle time INSIDE the event pair of a 30 us kernel.  Returns per-step ms."""
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    if gate:
        for _ in range(min(K, 400)):
            flush.zero_()
    for t in range(K):
        flush.zero_()  # evict the previous step's lines: every timed step starts with a cold L2
        starts[t].record(stream)
        step_fn(t)
        stops[t].record(stream)
    torch.cuda.synchronize()
    return np.array([starts[t].elapsed_time(stops[t]) for t in range(K)])


def gather_rank_steps(torch, per_step_ms, dev, world):
    """(world, K) per-step ms of every rank on rank 0 (identity for one rank)."""
    import torch.distributed as dist

    t = torch.tensor(per_step_ms, dtype=torch.float64, device=dev)
    if world > 1 and dist.is_initialized():
        out = torch.empty((world, t.numel()), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy()
    return t.cpu().numpy()[None]


def survey_bytes_per_env_step(N, L, O, W, F_obs, F_node=8):
    """SURVEY.md section 8(d)'s literal byte formula (no info rows, 48-byte world frame)."""
    E = N + L + O + W
    return 53 * N + 8 * (L + O) + 20 * W + 52 + 4 * F_obs * N + 4 * F_node * N * E + 4 * E * E + 12 * E * (E - 1)


def cpu_leg(cfg_args, worlds, steps, warmup, nthreads, max_seconds=25.0):
    """Time the oracle port (oracle/gmpe_oracle.c: fp64, RK45 where the reference uses it)
    on the host cores.  Returns (agent_steps_per_s, steps_done, ms_per_step)."""
    from gmpe_b200 import default_args, make_config
    from oracle.oracle import OracleEnv

    args = default_args(**cfg_args)
    cfg = make_config(args, worlds, precision="fp64", seed=1, env_edges=True, info=True)
    env = OracleEnv(cfg, use_rk45=True, nthreads=nthreads)
    env.reset()
    rs = np.random.RandomState(0)
    acts = rs.randint(cfg.total_actions, size=(8, worlds, cfg.num_agents)).astype(np.int32)
    for t in range(warmup):
        env.step(acts[t % 8])
    t0 = time.perf_counter()
    done = 0
    for t in range(steps):
        env.step(acts[t % 8])
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return worlds * cfg.num_agents * done / dt, done, dt / done * 1e3


def workload_config(opts, wl, args, Bg, world):
    """The `config` object both arms print (the reference arm runs on this arm's config)."""
    return {"workload": opts.workload, "scenario": args.scenario_name, "worlds_per_gpu": Bg,
            "worlds_total": Bg * world, "num_agents": args.num_agents, "num_landmarks": args.num_landmarks,
            "num_obstacles": args.num_obstacles, "num_walls": args.num_walls,
            "episode_length": args.episode_length, "actions": "seeded uniform random discrete, resident in HBM",
            "l2": "flushed before every timed step (256 MiB memset outside the event pair); "
                  "per-step CUDA events on the launching stream",
            "outputs": "obs,node_obs,adj,reward,done,info,env edges", "parallelism": f"dp{world} (sharded worlds)"}


def run_reference(opts):
    """--impl reference: the reference's own algorithm for this path on the box's host cores.
    The reference is pure Python and cannot travel to the GPU box, so this arm runs the C
    port of it (oracle/, pinned to the reference's golden traces) with every host thread, on
    the configuration of the other arm: worlds_per_gpu x gpus worlds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[opts.workload]
    Bg = opts.worlds or wl["worlds_per_gpu"]
    gpus = max(1, int(os.environ.get("WORLD_SIZE", opts.gpus)))
    worlds = Bg * gpus
    cores = os.cpu_count() or 1
    value, done, ms = cpu_leg(wl["args"], worlds, opts.steps, max(opts.warmup, 1), cores, max_seconds=120.0)
    from gmpe_b200 import default_args
    cfgd = workload_config(opts, wl, default_args(**wl["args"]), Bg, gpus)
    cfgd["actions"] = "seeded uniform random discrete (host)"
    cfgd["l2"] = "n/a (host cores)"
    cfgd["parallelism"] = f"{cores} pthreads over worlds"
    line = {
        "impl": "reference", "metric": "agent-steps/sec (env step + graph obs)", "value": value,
        "unit": "agent-steps/s", "n_gpus": opts.gpus, "steps": done, "warmup": opts.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfgd,
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{done} steps of the full {worlds}-world batch, {cores} pthreads, C port of the "
                                   "reference step (oracle/gmpe_oracle.c)"},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def short_line(torch, gdist, name, rank, world, dev, flush, K=20, W=3):
    """value / roofline fraction of a secondary workload (BASELINE configs 3-5), same timing rules."""
    from gmpe_b200 import default_args, make_config
    from gmpe_b200.engine import BatchedGraphMPE

    wl = WORKLOADS[name]
    args = default_args(**wl["args"])
    Bg, N = wl["worlds_per_gpu"], args.num_agents
    cfg = make_config(args, Bg, precision="fp32", seed=1, world_offset=rank * Bg, auto_reset=True,
                      env_edges=True, gnn_edges=False, info=True)
    eng = BatchedGraphMPE(cfg, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    acts = torch.randint(0, cfg.total_actions, (K + W, Bg, N), generator=gen, device=dev, dtype=torch.int32)
    eng.reset()
    for t in range(W):
        eng.step(acts[t])
    torch.cuda.synchronize()
    gdist.barrier()
    ms = time_steps_on_device(torch, torch.cuda.current_stream(dev), flush, lambda t: eng.step(acts[W + t]), K)
    total = gdist.max_over_ranks(float(ms.sum()), dev)
    abytes = eng.algorithmic_bytes_per_env_step() * Bg
    peak, _ = load_peaks()
    out = {"value": world * Bg * N * K / (total * 1e-3), "unit": "agent-steps/s", "ms_per_step": total / K,
           "worlds_per_gpu": Bg, "num_agents": N, "steps": K, "kernel": eng.kernel_name,
           "roofline_frac": abytes / (float(ms.mean()) * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": abytes}
    eng.close()
    del eng, acts
    torch.cuda.empty_cache()
    return out


def pcie_ceiling(torch, gdist, dev, d2h_bytes, h2d_bytes, n=20):
    """What the box can move: every rank copies the e2e step's byte counts between pinned host
    memory and its GPU, both directions at once, all ranks concurrently; seconds per step (max over
    ranks of the mean).  `e2e` can at best reach this."""
    pin_o = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    pin_i = torch.empty(max(h2d_bytes, 16), dtype=torch.uint8).pin_memory()
    d_o = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    d_i = torch.empty(max(h2d_bytes, 16), dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def once():
        with torch.cuda.stream(s1):
            pin_o.copy_(d_o, non_blocking=True)
        with torch.cuda.stream(s2):
            d_i.copy_(pin_i, non_blocking=True)
        s1.synchronize()
        s2.synchronize()

    for _ in range(3):
        once()
    gdist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        once()
    dt = (time.perf_counter() - t0) / n
    return gdist.max_over_ranks(dt, dev)


def run_ours(opts):
    import torch

    from gmpe_b200 import default_args, dist as gdist, make_config
    from gmpe_b200.engine import BatchedGraphMPE
    from gmpe_b200.env_wrappers import GraphSubprocVecEnv

    n_local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    cores = pin_rank(int(os.environ.get("LOCAL_RANK", "0")), n_local)
    if cores:  # torch's intra-op pool (the threaded staging copy) must not oversubscribe the rank's cores
        torch.set_num_threads(len(cores))
    rank, local, world = gdist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    wl = WORKLOADS[opts.workload]
    args = default_args(**wl["args"])
    Bg = opts.worlds or wl["worlds_per_gpu"]
    N = args.num_agents
    offset = rank * Bg  # weak scaling: fixed worlds per GPU, world ids are global
    cfg = make_config(args, Bg, precision="fp32", seed=1, world_offset=offset, auto_reset=True,
                      env_edges=True, gnn_edges=False, info=True)
    eng = BatchedGraphMPE(cfg, dev)
    K, W = opts.steps, max(opts.warmup, 3)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    acts = torch.randint(0, cfg.total_actions, (K + W, Bg, N), generator=gen, device=dev, dtype=torch.int32)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    eng.reset()
    for t in range(W):
        eng.step(acts[t])
    torch.cuda.synchronize()
    launches0 = eng.launch_count
    stream = torch.cuda.current_stream(dev)
    gdist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        per_step_ms = time_steps_on_device(torch, stream, flush, lambda t: eng.step(acts[W + t]), K)
        launches = eng.launch_count - launches0  # kernels of this build launched inside the timed region
        stats = eng_stats(eng, torch)
        gdist.reduce_stats(stats)  # the once-per-rollout NCCL reduction (no per-step collective)
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
        # keep the sampler alive for at least a second of load so the clocks line is meaningful
        t_end = time.perf_counter() + 1.0
        tt = 0
        while time.perf_counter() < t_end:
            eng.step(acts[W + (tt % K)])
            tt += 1
        torch.cuda.synchronize()
    gdist.barrier()
    total_ms = gdist.max_over_ranks(float(per_step_ms.sum()), dev)
    ms_per_step = total_ms / K
    value = world * Bg * N * K / (total_ms * 1e-3)
    rank_steps = gather_rank_steps(torch, per_step_ms, dev, world) * 1e3  # (world, K) us

    # hot-L2 rollout mode (no flush, K steps back to back) -- reported, not the headline
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(K):
        eng.step(acts[W + t])
    e1.record(stream)
    torch.cuda.synchronize()
    hot_ms = gdist.max_over_ranks(e0.elapsed_time(e1), dev) / K

    # roofline of the dominant (only) kernel of the step
    abytes = eng.algorithmic_bytes_per_env_step() * Bg
    sbytes = survey_bytes_per_env_step(N, args.num_landmarks, args.num_obstacles, args.num_walls, eng.layout.obs_dim,
                                       eng.layout.node_dim) * Bg
    kernel_name = eng.kernel_name
    peak, peak_src = load_peaks()
    kernel_ms = float(per_step_ms.mean())
    achieved = abytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(opts.workload)
        except Exception:
            traffic = None
    # context for the fraction: a plain device copy moving the SAME number of bytes (half read,
    # half written), timed exactly like a step.  MEASURED_PEAKS.json was taken on 1 GiB copies;
    # a launch of ~100 MB cannot reach that figure (launch ramp + DRAM pipeline fill).
    half = max(abytes // 2 // 16 * 16, 16)
    csrc = torch.empty(half, dtype=torch.uint8, device=dev)
    cdst = torch.empty(half, dtype=torch.uint8, device=dev)
    cts = []
    for t in range(13):
        flush.zero_()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        cdst.copy_(csrc)
        a1.record(stream)
        torch.cuda.synchronize()
        if t >= 3:
            cts.append(a0.elapsed_time(a1))
    copy_ms = float(np.median(cts))
    copy_gbs = 2 * half / (copy_ms * 1e-3) / 1e9
    del csrc, cdst

    # the rollout glue either side of the path: steps written in place into a
    # device rollout buffer with the runner's mask bookkeeping, actions as device int32 indices
    from gmpe_b200.rollout import DeviceGraphRollout

    slot_bytes = 4 * Bg * (N * eng.layout.obs_dim + N * eng.E * 8 + eng.E * eng.E + 4 * N)
    Tr = int(min(args.episode_length, max(2, (4 << 30) // slot_bytes)))  # rollout depth bounded to ~4 GiB of HBM
    roll = DeviceGraphRollout(eng, Tr)
    roll.warmup()
    for t in range(min(W, Tr)):
        roll.insert_env_step(acts[t])
    torch.cuda.synchronize()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(stream)
    for t in range(K):
        roll.insert_env_step(acts[W + t])
    r1.record(stream)
    torch.cuda.synchronize()
    roll_ms = gdist.max_over_ranks(r0.elapsed_time(r1), dev) / K
    del roll

    # e2e through the reference-facing vec env with host buffers
    eng.close()
    del eng
    torch.cuda.empty_cache()
    venv = GraphSubprocVecEnv(args=args, num_envs=Bg, device=dev, seed=1, world_offset=offset, info=True)
    venv.reset()
    A = cfg.total_actions
    Ke = max(3, min(K, opts.e2e_steps))
    rs = np.random.RandomState(5 + rank)
    onehots = [np.eye(A)[rs.randint(A, size=(Bg, N))] for _ in range(4)]  # float64, as graph_mpe_runner.py:375-377
    for t in range(3):
        venv.step(onehots[t % 4], 0)
    gdist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sink = 0.0
    d2h_total = 0
    for t in range(Ke):
        o = venv.step(onehots[t % 4], 0)
        sink += float(o[4].reshape(-1)[0])  # touch the host result
        if (t + 1) % 25 == 0:               # the runner reads infos when it logs (graph_mpe_runner.py:179-195)
            sink += float(o[6][0][0]["individual_reward"])
        d2h_total += venv.d2h_bytes_last
    torch.cuda.synchronize()
    e2e_rank_s = time.perf_counter() - t0
    e2e_s = gdist.max_over_ranks(e2e_rank_s, dev)
    e2e_value = world * Bg * N * Ke / e2e_s
    h2d, d2h = venv.h2d_bytes_last, d2h_total // Ke
    core_d2h = venv._core_bytes
    venv.close()
    del venv
    ceil_s = pcie_ceiling(torch, gdist, dev, int(core_d2h), int(h2d))
    ceil_value = world * Bg * N / ceil_s

    extra = {}
    if opts.extra:
        for name in ("july_air_taxi_n3_b131072", "navigation_graph_n16_o8_w4", "july_air_taxi_n64"):
            if name != opts.workload:
                extra[name] = short_line(torch, gdist, name, rank, world, dev, flush)

    if rank != 0:
        return
    cpu = None
    if world == 1 and not opts.no_cpu:
        ncpu = os.cpu_count() or 1
        # bounded sample: the whole batch for up to 400 steps or 10 s of wall time on every host
        # core (~10-30 core-seconds on the navigation workload), whichever comes first
        sample_worlds = Bg
        v, done, ms = cpu_leg(wl["args"], sample_worlds, 400, 2, ncpu, max_seconds=10.0)
        cpu = {"value": v, "unit": "agent-steps/s", "cores": ncpu, "kind": "port",
               "sample": f"{done} steps x {sample_worlds} worlds of the same workload, {ncpu} pthreads, "
                         "oracle/gmpe_oracle.c (fp64 C port of the reference step); the Python reference itself "
                         "runs ~960 agent-steps/s per process (profiles/r1_python_reference_cpu.txt)"}
    per_rank = [{"rank": r, "mean_us": float(x.mean()), "median_us": float(np.median(x)), "max_us": float(x.max()),
                 "p95_us": float(np.percentile(x, 95))} for r, x in enumerate(rank_steps)]
    slow = max(per_rank, key=lambda d: d["mean_us"])
    cfgd = workload_config(opts, wl, args, Bg, world)
    cfgd["l2"] += "; the device is given a backlog of flush kernels before the first timed step so that the host's " \
                  "launch latency stays outside the event pairs"
    line = {
        "metric": "agent-steps/sec (env step + graph obs)", "value": value, "unit": "agent-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfgd,
        "env_steps_per_s": value / N,
        "per_rank_step_us": per_rank, "slowest_rank": slow["rank"],
        "host_cores_of_rank0": cores,
        "hot_l2": {"ms_per_step": hot_ms, "value": world * Bg * N / (hot_ms * 1e-3),
                   "note": "K steps back to back without the flush (rollout mode)"},
        "device_rollout": {"ms_per_step": roll_ms, "value": world * Bg * N / (roll_ms * 1e-3),
                           "rollout_steps": Tr,
                           "note": "DeviceGraphRollout.insert_env_step: step outputs land in slot t+1 of the device "
                                   "rollout buffer (no copy) + masks / active_masks, K steps back to back"},
        "wall_s_timed_region": t_wall,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture of this kernel "
                                                           "and workload, profiles/r2x_ncu_full_*; not re-measured by this run)",
                     "peak_source": peak_src, "kernel": kernel_name + "<float>",
                     "algorithmic_bytes_per_launch": abytes, "kernel_ms": kernel_ms,
                     "frac_survey_formula": sbytes / (kernel_ms * 1e-3) / 1e9 / peak,
                     "survey_bytes_per_launch": sbytes,
                     "copy_same_bytes": {"ms": copy_ms, "gbs": copy_gbs, "frac_of_it": achieved / copy_gbs,
                                         "note": "torch copy_ of the same byte count, same flush + event timing"}},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": Ke, "ms_per_step": e2e_s / Ke * 1e3,
                "pcie_ceiling": {"value": ceil_value, "unit": "agent-steps/s", "ms_per_step": ceil_s * 1e3,
                                 "frac_of_it": e2e_value / ceil_value,
                                 "note": "every rank moving the same byte counts (down and up at once) with plain "
                                         "pinned copies, all ranks concurrently: what the box's PCIe / host memory "
                                         "can do for this API at this rank count"},
                "api": "GraphSubprocVecEnv.step(one-hot float64 host actions) -> host numpy views (pinned) of obs, "
                       "node_obs, adj, rewards, dones every step; info rows are copied when read (every 25th step "
                       "here, as the runner's logging does); one staging copy + one upload, decode + step, one download "
                       "(directions never overlap: profiles/r2_e2e_timeline.txt)"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "rollout_stats": [float(x) for x in stats.cpu().tolist()],
        "extra": extra,
    }
    print(json.dumps(line), flush=True)


def eng_stats(eng, torch):
    from gmpe_b200 import abi

    cnt = eng.state["agent_cnt"].to(torch.float64)
    return torch.stack([eng.out["reward"].to(torch.float64).sum(), cnt[..., abi.AC_AGENT_COLL].sum(),
                        cnt[..., abi.AC_OBST_COLL].sum(), eng.out["done"].to(torch.float64).sum()])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--worlds", type=int, default=0, help="worlds per GPU (default: the workload's)")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", dest="extra", action="store_false",
                    help="skip the short lines of BASELINE configs 3-5 under the `extra` key")
    opts = ap.parse_args()
    if opts.impl == "reference":
        run_reference(opts)
    else:
        run_ours(opts)
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
