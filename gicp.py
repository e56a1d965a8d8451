"""Module-name-compatible stand-in for the reference's ``python-implementation/gicp.py``:
``from gicp import gicp, apply_transformation`` (visualization.py:7, robot-visualization.py:6)
resolves here when the repo root is first on sys.path.  Nothing CUDA-related runs at import."""
from generalized_icp_b200.compat import apply_transformation, gicp, gicp_extended  # noqa: F401
