"""Build libgmpe_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python contracts-marl-aam-corridors_b200/csrc/build.py [--force] [--verbose]

Links the CUDA runtime statically so the library carries no dependency on torch's
bundled libcudart; the host passes raw device pointers and a cudaStream_t.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libgmpe_b200.so")
# one translation unit for the host side + tile kernel, one per precision for
# the lane-per-agent kernel instantiations: compiled in parallel, linked into one library
SOURCES = ["gmpe_api.cu", "gmpe_coop_f32.cu", "gmpe_coop_f64.cu", "gmpe_coop_wide_f32.cu", "gmpe_coop_wide_f64.cu"]
DEPS = SOURCES + ["gmpe_step.cuh", "gmpe_place.cuh", "gmpe_coop.cuh", "gmpe_device.cuh", "gmpe_flat.cuh", "gmpe_coop_table.h",
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
    nvcc = nvcc_path()
    common = [nvcc, "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3"] + [f"-D{d}" for d in defines]
    if os.path.exists("/usr/bin/g++"):
        common[1:1] = ["-ccbin", "/usr/bin/g++"]
    tag = hashlib.sha1((OUT_local + "|" + ",".join(defines)).encode()).hexdigest()[:10]
    objdir = os.path.join(HERE, "..", "..", "build", "obj_" + tag)
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        res = subprocess.run(common + ["-c", os.path.join(HERE, src), "-o", obj], cwd=HERE, capture_output=True, text=True)
        return src, obj, res

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        done = list(pool.map(compile_one, SOURCES))
    for src, obj, res in done:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stderr[-4000:])
    link = [nvcc] + (["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []) + \
           ["-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT_local] + [o for _, o, _ in done]
    res = subprocess.run(link, cwd=HERE, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stderr[-4000:])
    return OUT_local


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
