"""TEST INFRASTRUCTURE -- ctypes wrapper of oracle/_build/libgmpe_oracle.so (the CPU
restatement of the reference step path, oracle/gmpe_oracle.c).

May only be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import gmpe_b200
from gmpe_b200 import abi, buffers

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libgmpe_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "gmpe_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "gmpe.h")
    stale = (not os.path.exists(LIB)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(LIB) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.gmpe_oracle_step.argtypes = [C.POINTER(abi.GmpeConfig), C.POINTER(abi.GmpeState), C.c_void_p,
                                       C.POINTER(abi.GmpeStepInject), C.POINTER(abi.GmpeStepOut), C.c_int, C.c_int]
        L.gmpe_oracle_reset.argtypes = [C.POINTER(abi.GmpeConfig), C.POINTER(abi.GmpeState), C.c_void_p,
                                        C.POINTER(abi.GmpeStepInject), C.POINTER(abi.GmpeStepOut), C.c_int]
        L.gmpe_oracle_decode_onehot.argtypes = [C.POINTER(abi.GmpeConfig), C.c_void_p, C.c_void_p]
        L.gmpe_oracle_integrate.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double]
        L.gmpe_oracle_uniform01.argtypes = [C.c_uint64, C.c_int64, C.c_uint32, C.c_uint32]
        L.gmpe_oracle_uniform01.restype = C.c_double
        L.gmpe_oracle_overflows.argtypes = [C.c_int]
        L.gmpe_oracle_overflows.restype = C.c_longlong
        _lib = L
    return _lib


class OracleEnv:
    """Batch of worlds stepped by the CPU oracle (always float64)."""

    def __init__(self, cfg: abi.GmpeConfig, use_rk45: bool = True, nthreads: int = 1):
        self.cfg = cfg.copy()
        self.cfg.precision = abi.F64
        self.use_rk45 = int(use_rk45)
        self.nthreads = int(nthreads)
        self.state = buffers.alloc_numpy(self.cfg, buffers.STATE_FIELDS)
        self.state["agent_cnt"][..., abi.AC_TIME_REQ_I:] = -1
        self.out = buffers.alloc_numpy(self.cfg, buffers.OUT_FIELDS)
        self._st = buffers.make_struct(abi.GmpeState, self.state)
        self._out = buffers.make_struct(abi.GmpeStepOut, self.out)

    def reset(self, env_mask=None, reset_state: dict | None = None, reset_draws=None):
        if reset_draws is not None:
            reset_draws = np.ascontiguousarray(reset_draws, dtype=np.float64)
            assert reset_draws.ndim == 2 and reset_draws.shape[0] == self.cfg.num_worlds
        inj, keep = buffers.make_inject(None, reset_state, reset_draws)
        m = None
        if env_mask is not None:
            m = np.ascontiguousarray(env_mask, dtype=np.uint8)
        rc = lib().gmpe_oracle_reset(C.byref(self.cfg), C.byref(self._st), m.ctypes.data if m is not None else None,
                                     C.byref(inj), C.byref(self._out), self.nthreads)
        assert rc == 0, rc
        return self.out

    def step(self, action_idx, theta_draw=None, reset_state: dict | None = None):
        a = np.ascontiguousarray(action_idx, dtype=np.int32)
        assert a.shape == (self.cfg.num_worlds, self.cfg.num_agents)
        th = None if theta_draw is None else np.ascontiguousarray(theta_draw, dtype=np.float64)
        inj, keep = buffers.make_inject(th, reset_state)
        rc = lib().gmpe_oracle_step(C.byref(self.cfg), C.byref(self._st), a.ctypes.data, C.byref(inj),
                                    C.byref(self._out), self.use_rk45, self.nthreads)
        assert rc == 0, rc
        return self.out


def integrate(dyn: int, use_rk45: bool, y, u, dt: float):
    y = np.array(y, dtype=np.float64)
    u = np.array(u, dtype=np.float64)
    lib().gmpe_oracle_integrate(dyn, int(use_rk45), y.ctypes.data, u.ctypes.data, dt)
    return y


def overflows(clear: bool = False) -> int:
    """Worlds whose placement gave up a rejection loop since the last clear (process-wide)."""
    return int(lib().gmpe_oracle_overflows(int(clear)))


def uniform01(seed, world, episode, idx) -> float:
    return lib().gmpe_oracle_uniform01(seed, world, episode, idx)
