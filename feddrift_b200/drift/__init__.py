"""Concept-drift FL: state machines (softcluster / states), evaluation, and the FedML-compatible ``fedavg_ens`` API."""
from .softcluster import SoftClusterState, parse_algo_arg
from .states import AdaState, DriftSurfState, KueState, MultiModelAccState
from .evaluator import Evaluator

__all__ = ["SoftClusterState", "parse_algo_arg", "AdaState", "DriftSurfState", "KueState", "MultiModelAccState", "Evaluator"]
