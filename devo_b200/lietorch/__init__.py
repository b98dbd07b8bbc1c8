"""`lietorch` operator API (devo/lietorch/__init__.py:1-2) on top of the B200 backend."""
from .groups import LieGroupParameter, SO3, RxSO3, SE3, Sim3, cat, stack  # noqa: F401

__all__ = ["LieGroupParameter", "SO3", "RxSO3", "SE3", "Sim3", "cat", "stack"]
