from . import so3  # noqa: F401
