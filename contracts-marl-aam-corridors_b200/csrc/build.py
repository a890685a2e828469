"""Build libgmpe_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python contracts-marl-aam-corridors_b200/csrc/build.py [--force] [--verbose]

Links the CUDA runtime statically so the library carries no dependency on torch's
bundled libcudart; the host passes raw device pointers and a cudaStream_t.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libgmpe_b200.so")
SOURCES = ["gmpe_api.cu"]
DEPS = ["gmpe_api.cu", "gmpe_step.cuh", "gmpe_small.cuh", "gmpe_coop.cuh", "gmpe_device.cuh",
        os.path.join("..", "..", "include", "gmpe.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines: tuple = ()) -> str:
    """defines: extra -D flags (tuning variants built by tools/sweep.py go to `out`)."""
    global OUT
    if out is not None:
        OUT_local = out
    else:
        OUT_local = OUT
        if not force and not is_stale():
            return OUT
    cmd = [
        nvcc_path(), "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
        "-Xptxas", "-v" if verbose else "-O3",
        "-o", OUT_local,
    ] + [f"-D{d}" for d in defines] + [os.path.join(HERE, s) for s in SOURCES]
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, cwd=HERE, env=env, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stderr[-4000:])
    return OUT_local


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
