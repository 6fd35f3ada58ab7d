"""so3_exponential_map(log_rot[H,3]) -> [H,3,3] (Rodrigues), as used by layers/utils.py:29,56."""
from cloud_transformers_b200.so3 import so3_exponential_map  # noqa: F401
