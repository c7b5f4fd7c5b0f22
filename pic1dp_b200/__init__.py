"""pic1dp_b200 -- B200-native (sm_100a) implementation of the PIC1D-PETSc per-timestep hot path.

Only what the path needs: csrc/ (CUDA kernels + the C ABI of include/pic1dp_gpu.h), a ctypes binding and the
host-side mirror of the reference's module interface.  Importing the package does not need a GPU; creating a
handle does, and fails loudly without one (no CPU fallback).
"""
from ._capi import (ARITH_STRICT, ARITH_TOLERANCE, DEPOSIT_AUTO, DEPOSIT_FIXED, DEPOSIT_GLOBAL_RED, DEPOSIT_SMEM_ATOMIC,
                    DEPOSIT_WARP_PRIVATE, FIELD_SEQUENTIAL, FIELD_TREE, Counters, Params)
from .host import Pic1dGpu, Pic1dpError, Pic1dpModules, default_params, petsc_decide

__all__ = [
    "Params", "Counters", "Pic1dGpu", "Pic1dpModules", "Pic1dpError", "default_params", "petsc_decide",
    "DEPOSIT_AUTO", "DEPOSIT_SMEM_ATOMIC", "DEPOSIT_GLOBAL_RED", "DEPOSIT_WARP_PRIVATE", "DEPOSIT_FIXED",
    "ARITH_STRICT", "ARITH_TOLERANCE", "FIELD_TREE",
    "FIELD_SEQUENTIAL",
]
