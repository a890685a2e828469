#!/usr/bin/env python
"""Where one host-facing GraphSubprocVecEnv.step spends its time (run under gpurun)."""
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
dev = torch.device("cuda:0")


def t(fn, n=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for nbytes in (8 << 20, 66 << 20):
    pin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h2d = t(lambda: d.copy_(pin, non_blocking=True))
    d2h = t(lambda: pin.copy_(d, non_blocking=True))
    print(f"{nbytes >> 20} MiB pinned: H2D {h2d:.3f} ms = {nbytes / h2d / 1e6:.1f} GB/s   D2H {d2h:.3f} ms = {nbytes / d2h / 1e6:.1f} GB/s")
a64 = np.eye(A)[np.random.randint(A, size=(B, N))]
pin = torch.empty((B, N, A), dtype=torch.float64).pin_memory()
print(f"numpy copy of the one-hot block into pinned memory: {t(lambda: pin.numpy().__setitem__(Ellipsis, a64)):.3f} ms")
src = torch.from_numpy(a64)
print(f"torch copy_ of the same: {t(lambda: pin.copy_(src)):.3f} ms")
print(f"np.argmax on the host: {t(lambda: a64.argmax(-1), 5):.3f} ms")
venv = GraphSubprocVecEnv(args=args, num_envs=B, device=dev, seed=1, info=True)
venv.reset()
idx = np.random.randint(A, size=(B, N)).astype(np.int32)
print(f"venv.step(one-hot float64): {t(lambda: venv.step(a64, 0)):.3f} ms")
print(f"venv.step(int32 indices):   {t(lambda: venv.step(idx, 0)):.3f} ms")
a32 = a64.astype(np.float32)
print(f"venv.step(one-hot float32): {t(lambda: venv.step(a32, 0)):.3f} ms")
print("d2h bytes", venv.d2h_bytes_last, "h2d bytes", venv.h2d_bytes_last)
