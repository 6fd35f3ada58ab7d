"""Drop-in replacement for the reference's layers/cloud_transform.py.

Copy this file over `layers/cloud_transform.py` of a cloud_transformers checkout (or put this
directory's parent ahead of it on sys.path) and make `cloud_transformers_b200`
importable: model_zoo/* and train_*.py then run unchanged on the B200 kernels.
"""
from cloud_transformers_b200.cloud_transform import (  # noqa: F401
    GradientBalancing, balance_op, DifferentiableGridModule, DifferentiablePositions, Splat, Slice)
