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
_ALIGN_MIN_NUMEL = 256   # tensors at least this big start on a 16-byte boundary of the flat row
_ALIGN_FLOATS = 4


def flat_spec(module_or_sd) -> List[Tuple[str, torch.Size, torch.dtype, int, int]]:
    """(key, shape, dtype, offset, numel) for every state_dict entry, in state_dict order.

    Big tensors (conv / linear weights) are placed at offsets that are multiples of 4 floats so that the views the
    modules train through are 16-byte aligned (cuDNN's NHWC tensor-core kernels, TMA and 128-bit loads require it);
    the gaps are zero and ride along in every row operation.  Small-MLP layouts (every tensor < 256 elements) stay
    densely packed — the register-resident MLP kernels index W1 | b1 | W2 | b2 contiguously."""
    sd = module_or_sd.state_dict() if isinstance(module_or_sd, nn.Module) else module_or_sd
    out, off = [], 0
    for k, v in sd.items():
        n = v.numel()
        if n >= _ALIGN_MIN_NUMEL:
            off = (off + _ALIGN_FLOATS - 1) // _ALIGN_FLOATS * _ALIGN_FLOATS
        out.append((k, v.shape, v.dtype, off, n))
        off += n
    return out


def ohwi_stored(shape) -> bool:
    """Which 4-D tensors live in the flat rows in (O, kh, kw, I) order: the convolution weights the implicit-GEMM tensor-core
    kernels can take (channel counts multiples of 32, square filter > 1×1).  A pure function of the shape, so every flatten /
    unflatten of a state_dict agrees.  (1×1 filters have the same memory order either way; stems and depthwise filters stay
    in logical order and on the library path.)"""
    return len(shape) == 4 and shape[0] % 32 == 0 and shape[1] % 32 == 0 and shape[2] == shape[3] and shape[2] > 1


def flat_view(t: torch.Tensor) -> torch.Tensor:
    """The tensor's elements in FLAT-ROW ORDER.  Tensor-core-eligible convolution weights ``[O, I, kh, kw]`` (:func:`ohwi_stored`)
    are stored in the row in (O, kh, kw, I) order — the ``channels_last`` memory format, which is the K-major operand layout of the implicit-GEMM
    convolution kernels: the forward / data-gradient weight operand is then a plain bf16 cast of the row segment and the
    weight-gradient GEMM reduce-adds straight into the flat gradient row (no per-step transposing pack, no un-permute).
    Everything else is stored in logical (row-major) order.  Row operations (aggregation, optimizer, norms, distances) are
    elementwise over rows and therefore layout-agnostic."""
    return t.permute(0, 2, 3, 1).reshape(-1) if ohwi_stored(t.shape) else t.reshape(-1)


def flat_size(module_or_sd) -> int:
    """Row length: end of the last tensor, rounded up to 4 floats for models with aligned (big) tensors so that every
    row of a dense ``[C, M, P]`` arena starts on a 16-byte boundary too."""
    spec = flat_spec(module_or_sd)
    if not spec:
        return 0
    end = spec[-1][3] + spec[-1][4]
    if any(n >= _ALIGN_MIN_NUMEL for _, _, _, _, n in spec):
        end = (end + _ALIGN_FLOATS - 1) // _ALIGN_FLOATS * _ALIGN_FLOATS
    return end


def flatten_state_dict(sd: Dict[str, torch.Tensor], out: torch.Tensor = None) -> torch.Tensor:
    """state_dict → flat fp32 row laid out by ``flat_spec(sd)`` (alignment gaps are zero)."""
    spec = flat_spec(sd)
    total = flat_size(sd)
    vals = list(sd.values())
    dev = vals[0].device if vals else "cpu"
    dense = all(spec[i][3] + spec[i][4] == spec[i + 1][3] for i in range(len(spec) - 1))
    if dense and (not spec or spec[-1][3] + spec[-1][4] == total):   # dense layout: one cat
        flat = torch.cat([flat_view(v).to(torch.float32) for v in vals]) if vals else torch.zeros(0)
    else:
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        for (_, _, _, off, n), v in zip(spec, vals):
            flat[off:off + n] = flat_view(v).to(torch.float32)
    if out is not None:
        out.copy_(flat)
        return out
    return flat


def unflatten_to_state_dict(flat: torch.Tensor, spec) -> "OrderedDict[str, torch.Tensor]":
    sd = OrderedDict()
    for k, shape, dtype, off, n in spec:
        seg = flat[off:off + n]
        if ohwi_stored(shape):
            # stored (O, kh, kw, I): a logical [O, I, kh, kw] VIEW with channels_last strides (see flat_view); such tensors must
            # be consumed by TcConv2d, whose CPU path canonicalises memory formats (oneDNN's convolution backward corrupts the
            # heap for channels_last inputs of 1×1 stride-2 filters) — ModelBank converts its template accordingly
            v = seg.reshape(shape[0], shape[2], shape[3], shape[1]).permute(0, 3, 1, 2)
        else:
            v = seg.reshape(shape)
        sd[k] = v.to(dtype)
    return sd


def weight_param_mask(spec) -> torch.Tensor:
    """1 for learnable weights, 0 for BN running stats / counters (robust aggregation excludes them)."""
    from ..core.robustness import is_weight_param
    total = spec[-1][3] + spec[-1][4] if spec else 0
    if any(n >= _ALIGN_MIN_NUMEL for _, _, _, _, n in spec):
        total = (total + _ALIGN_FLOATS - 1) // _ALIGN_FLOATS * _ALIGN_FLOATS
    mask = torch.zeros(total, dtype=torch.bool)
    for k, _, _, off, n in spec:
        if is_weight_param(k):
            mask[off:off + n] = True
    return mask


def count_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())
