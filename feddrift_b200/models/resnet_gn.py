"""GroupNorm ResNets (parity: ``fedml_api/model/cv/resnet_gn.py``) — thin re-export of the shared implementation."""
from .resnet import resnet18, resnet34, resnet50, resnet101, resnet152  # noqa: F401
