#!/usr/bin/env python
"""Timeline of one host-facing GraphSubprocVecEnv.step (one-hot float64 actions in, every tensor
out to pinned host memory): where inside the ~1.3 ms do the slices' copies and kernels sit?
Run under gpurun."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmpe_b200  # noqa: E402,F401
from gmpe_b200 import default_args  # noqa: E402
from gmpe_b200.env_wrappers import GraphSubprocVecEnv  # noqa: E402

B, N, A = 65536, 3, 5
args = default_args(scenario_name="navigation_graph", num_agents=3, num_landmarks=3, episode_length=25, collaborative=True)
venv = GraphSubprocVecEnv(args=args, num_envs=B, device=0, seed=1)
venv.reset()
a64 = np.eye(A)[np.random.randint(A, size=(B, N))]
for _ in range(5):
    venv.step(a64, 0)

# instrument: wrap the three streams' work with events
marks = []
orig = venv._step_pipelined


def timed_pipelined(pinned, dev, onehot, src=None):
    eng = venv.engine
    cur = torch.cuda.current_stream(eng.device)
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record(cur)
    marks.clear()
    marks.append(("start", t0, time.perf_counter()))
    venv._s_in.wait_stream(cur)
    host = venv._host_t[venv._flip]
    arena = venv._arena[venv._flip]
    venv._flip ^= 1
    for k, (b0, b1) in enumerate(venv._bounds()):
        if src is not None:
            pinned[b0:b1].copy_(src[b0:b1])
        marks.append((f"cpu staged {k}", None, time.perf_counter()))
        with torch.cuda.stream(venv._s_in):
            dev[b0:b1].copy_(pinned[b0:b1], non_blocking=True)
            ev_in = torch.cuda.Event(enable_timing=True)
            ev_in.record(venv._s_in)
        marks.append((f"h2d done {k}", ev_in, time.perf_counter()))
        cur.wait_event(ev_in)
        eng.step_range(eng.action_idx if onehot else dev, b0, b1 - b0, stream=cur, onehot=dev if onehot else None)
        ev_k = torch.cuda.Event(enable_timing=True)
        ev_k.record(cur)
        marks.append((f"kernel done {k}", ev_k, time.perf_counter()))
        venv._s_out.wait_event(ev_k)
        with torch.cuda.stream(venv._s_out):
            for n in venv._host_names:
                host[n][b0:b1].copy_(eng.out[n][b0:b1], non_blocking=True)
            ev_o = torch.cuda.Event(enable_timing=True)
            ev_o.record(venv._s_out)
        marks.append((f"d2h done {k}", ev_o, time.perf_counter()))
    venv._pending = arena


venv._step_pipelined = timed_pipelined
for rep in range(3):
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    venv.step(a64, 0)
    w1 = time.perf_counter()
    t0 = marks[0][1]
    c0 = marks[0][2]
    print(f"--- step {rep}: wall {1e3 * (w1 - w0):.3f} ms (step() entered {1e3 * (c0 - w0):.3f} ms before the first enqueue)")
    for name, ev, cpu in marks[1:]:
        g = f"gpu {t0.elapsed_time(ev):7.3f} ms" if ev is not None else " " * 14
        print(f"  {name:16s} cpu {1e3 * (cpu - c0):7.3f} ms   {g}")
    print(f"  step() returned  cpu {1e3 * (w1 - c0):7.3f} ms")
