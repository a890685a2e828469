"""gmpe_b200 -- B200-native batched simulator for the GraphMPE step path of
Jaroan/Contracts-MARL-AAM-Corridors (directory ``contracts-marl-aam-corridors_b200/``;
import it as ``gmpe_b200`` through the alias module at the repository root).

Layers (bottom up):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI of include/gmpe.h
  abi / buffers    ctypes mirror of the header; array shape table
  config           reference constants and the argparse-Namespace -> GmpeConfig mapping
  engine           BatchedGraphMPE: device tensors + calls through the C ABI
  scenario / environment / env_wrappers
                   mirrors of the reference plugin surfaces
  dist             world sharding over GPUs + the once-per-rollout NCCL stats reduction
"""
from . import abi, buffers, config  # noqa: F401
from .config import default_args, make_config  # noqa: F401

__all__ = ["abi", "buffers", "config", "default_args", "make_config"]
