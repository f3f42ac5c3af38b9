"""B200-native batch optimiser for arm_slam_calib factor graphs (C-ABI library + thin ctypes host layer).

The product is `libasc_b200.so` (include/asc.h).  This package only loads it and mirrors the
reference's optimiser surface for tests and bench.py; it contains no CPU fallback.
"""
from ._ctypes_defs import (ASC_BATCH_DOGLEG, ASC_BATCH_LM, ASC_ERR_CUDA, ASC_ERR_INDETERMINATE, ASC_ERR_INVALID, ASC_OK,
                           AscIterLog, AscLinearStats, AscParams, AscSimConfig)
from .capi import (AscError, BatchOptimizer, IndeterminantLinearSystemException, Problem, default_params, hostlib, lib,
                   optimization_mode_from_string, shard_by_landmark)
from .sim import chain_generate, chain_mico, make_sim_problem, with_drift, with_velocity

__all__ = [n for n in dir() if not n.startswith("_")]
