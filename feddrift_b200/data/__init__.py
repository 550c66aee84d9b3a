from . import changepoints
from .drift import (DEFAULT_DELTAS, DriftData, generate_drift_data, load_all_data, load_partition_data,
                    select_iterations)

__all__ = ["changepoints", "DEFAULT_DELTAS", "DriftData", "generate_drift_data", "load_all_data",
           "load_partition_data", "select_iterations"]
