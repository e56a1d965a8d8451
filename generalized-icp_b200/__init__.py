"""B200-native GICP registration engine (hot path of msi-se/generalized-icp).

Nothing CUDA-related happens at import: the reference's robot demo imports the
module in a parent process and first calls ``gicp`` in a forked worker
(robot-visualization.py:151-166,199), so the engine is created lazily inside
the call.  ``apply_transformation`` is pure numpy (it runs in the UI process).
"""
from .compat import apply_transformation, gicp  # noqa: F401

__all__ = ["gicp", "apply_transformation", "GicpEngine", "GicpParams"]


def __getattr__(name):
    if name in ("GicpEngine", "GicpParams", "RegistrationResult"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
