"""dqn-hfo_b200 — B200-native replacement for the hot path of mhauskn/dqn-hfo's dqn::DQN.

The product is csrc/ (hand-written sm_100a CUDA behind the C-ABI of include/dqn_b200.h) plus the
C++ mirror of the reference class in host/.  This Python package is only the ctypes binding the
tests and bench.py drive the C-ABI with; it never computes anything itself and fails loudly when
libdqn_b200.so is missing (there is no CPU fallback).

The directory name contains a hyphen, so import it through `__graft_entry__.load_package()`,
which registers it as module `dqn_hfo_b200`.
"""
from .binding import (DQNB, Config, GEMM_SIMT_FP32, GEMM_TCGEN05_3XTF32, ACTOR, CRITIC, ACTOR_TARGET,
                      CRITIC_TARGET, build_library, comm_unique_id, gemm_test, lib, lib_path)

__all__ = ["DQNB", "Config", "GEMM_SIMT_FP32", "GEMM_TCGEN05_3XTF32", "ACTOR", "CRITIC", "ACTOR_TARGET",
           "CRITIC_TARGET", "build_library", "comm_unique_id", "gemm_test", "lib", "lib_path"]
