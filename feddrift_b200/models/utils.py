"""Model factory / re-initialisation / flat-parameter helpers.

Parity: ``fedml_api/model/utils.py:7-24`` — ``reinitialize`` re-seeds with ``torch_seed`` and calls
``reset_parameters`` on the direct children, so *every* re-initialised model gets identical weights
(FedDrift relies on it).  ``create_model`` mirrors ``main_fedavg.py:207-224``.
No pretrained ImageNet weights exist offline: resnet18/densenet121 are random-init (BASELINE.json
prescribes random-init weights).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import torch
from torch import nn

torch_seed = 42  # default value, may be reset by the experiment main (dummy_arg)


def reinitialize(model: nn.Module, seed: int = None) -> nn.Module:
    torch.manual_seed(torch_seed if seed is None else seed)
    direct = [m for m in model.children() if hasattr(m, "reset_parameters")]
    if direct or not any(True for _ in model.children()):
        for layer in direct:
            layer.reset_parameters()
        if hasattr(model, "reset_parameters") and not direct:
            model.reset_parameters()
    else:  # deep nets (resnet/densenet): the reference reloads pretrained weights; offline we re-run every init
        for layer in model.modules():
            if layer is not model and hasattr(layer, "reset_parameters"):
                layer.reset_parameters()
    return model


def create_model(model_name: str, output_dim: int, feature_dim: int = None, **kw) -> nn.Module:
    from . import zoo
    model = zoo.build(model_name, output_dim, feature_dim, **kw)
    reinitialize(model)
    return model


# ----------------------------------------------------------------------------- flat views
def flat_spec(module_or_sd) -> List[Tuple[str, torch.Size, torch.dtype, int, int]]:
    """(key, shape, dtype, offset, numel) for every state_dict entry, in state_dict order."""
    sd = module_or_sd.state_dict() if isinstance(module_or_sd, nn.Module) else module_or_sd
    out, off = [], 0
    for k, v in sd.items():
        n = v.numel()
        out.append((k, v.shape, v.dtype, off, n))
        off += n
    return out


def flat_size(module_or_sd) -> int:
    spec = flat_spec(module_or_sd)
    return spec[-1][3] + spec[-1][4] if spec else 0


def flatten_state_dict(sd: Dict[str, torch.Tensor], out: torch.Tensor = None) -> torch.Tensor:
    parts = [v.reshape(-1).to(torch.float32) for v in sd.values()]
    flat = torch.cat(parts) if parts else torch.zeros(0)
    if out is not None:
        out.copy_(flat)
        return out
    return flat


def unflatten_to_state_dict(flat: torch.Tensor, spec) -> "OrderedDict[str, torch.Tensor]":
    sd = OrderedDict()
    for k, shape, dtype, off, n in spec:
        sd[k] = flat[off:off + n].reshape(shape).to(dtype)
    return sd


def weight_param_mask(spec) -> torch.Tensor:
    """1 for learnable weights, 0 for BN running stats / counters (robust aggregation excludes them)."""
    from ..core.robustness import is_weight_param
    total = spec[-1][3] + spec[-1][4] if spec else 0
    mask = torch.zeros(total, dtype=torch.bool)
    for k, _, _, off, n in spec:
        if is_weight_param(k):
            mask[off:off + n] = True
    return mask


def count_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())
