"""cloud_transformers_b200 -- B200 (sm_100a) native Splat / Slice hot path of Cloud Transformers.

Public surface (mirrors the reference's layers/cloud_transform.py):
    DifferentiablePositions, Splat, Slice, DifferentiableGridModule, GradientBalancing, balance_op
plus `functional` (autograd Functions over the C ABI) and `config` (mode / fused switches).
The CUDA library is loaded lazily on first use and there is no CPU fallback.
"""
from .cloud_transform import (DifferentiableGridModule, DifferentiablePositions, Splat, Slice,
                              GradientBalancing, balance_op)
from .functional import config
from . import functional

__all__ = ["DifferentiableGridModule", "DifferentiablePositions", "Splat", "Slice", "GradientBalancing",
           "balance_op", "config", "functional"]
