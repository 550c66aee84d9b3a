from .engine import DriftSim, make_args

__all__ = ["DriftSim", "make_args"]
