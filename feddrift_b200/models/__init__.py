from .basic import FeedForwardNN, LogisticRegression
from .utils import create_model, reinitialize, flat_spec, flat_size, flatten_state_dict, unflatten_to_state_dict

__all__ = ["FeedForwardNN", "LogisticRegression", "create_model", "reinitialize", "flat_spec", "flat_size",
           "flatten_state_dict", "unflatten_to_state_dict"]
