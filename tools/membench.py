#!/usr/bin/env python
"""What the memory system of this GPU does for traffic shaped like one step of the small-world
workload (about 19 MB read, 101 MB written), timed like bench.py times a step (256 MiB L2
flush before every iteration, CUDA events around the kernel).  Run under gpurun."""
import torch

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=30, do_flush=True):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


MB = 1 << 20
for mb_w, mb_r in ((101, 19), (120, 0), (60, 60), (240, 0), (120, 120), (1000, 0), (512, 512)):
    w = torch.empty(mb_w * MB // 4, dtype=torch.float32, device=dev)
    r = torch.empty(max(mb_r, 1) * MB // 4, dtype=torch.float32, device=dev)
    res = []
    for fl in (True, False):
        if mb_r == 0:
            t = timeit(lambda: w.fill_(1.0), do_flush=fl)
        elif mb_r == mb_w:
            t = timeit(lambda: w.copy_(r), do_flush=fl)
        else:
            acc = torch.empty((), dtype=torch.float32, device=dev)

            def f():
                w.fill_(1.0)
                torch.sum(r, dim=(0,), out=acc)
            t = timeit(f, do_flush=fl)
        res.append(t)
    tot = mb_w + mb_r
    print(f"write {mb_w:5d} MiB + read {mb_r:4d} MiB: cold {res[0]:8.1f} us = {tot * MB / res[0] / 1e6:7.1f} TB/s   "
          f"hot {res[1]:8.1f} us = {tot * MB / res[1] / 1e6:7.1f} TB/s", flush=True)
