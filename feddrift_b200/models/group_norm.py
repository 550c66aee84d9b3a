"""GroupNorm2d / GroupNorm3d (K16).

Parity: ``fedml_api/model/cv/group_normalization.py:7-118`` (the reference implements GN by reshaping into
``F.batch_norm``).  On CUDA both directions are the fused sm_100a kernels of ``csrc/misc.cu`` through ``ops.group_norm``
(forward: one CTA per (sample, group), two passes over the contiguous slab, saves mean / rstd; backward: one CTA per
(sample, group) producing dx and per-sample dγ / dβ partials); CPU tensors use ``F.group_norm``.  Parameter names
(``weight``, ``bias``) match.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


class _GroupNorm(nn.Module):
    def __init__(self, num_features: int, num_groups: int = 32, eps: float = 1e-5, affine: bool = True,
                 track_running_stats: bool = False):
        super().__init__()
        if num_features % num_groups != 0:
            raise ValueError("num_features must be divisible by num_groups")
        self.num_features, self.num_groups, self.eps, self.affine = num_features, num_groups, eps, affine
        self.track_running_stats = track_running_stats
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def reset_parameters(self) -> None:
        if self.affine:
            nn.init.ones_(self.weight)
            nn.init.zeros_(self.bias)

    def _check_input_dim(self, x):
        raise NotImplementedError

    def forward(self, x):
        self._check_input_dim(x)
        if x.is_cuda:
            return ops.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)
        return F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)

    def extra_repr(self) -> str:
        return f"{self.num_features}, groups={self.num_groups}, eps={self.eps}, affine={self.affine}"


class GroupNorm2d(_GroupNorm):
    def _check_input_dim(self, x):
        if x.dim() != 4:
            raise ValueError(f"expected 4D input (got {x.dim()}D input)")


class GroupNorm3d(_GroupNorm):
    def _check_input_dim(self, x):
        if x.dim() != 5:
            raise ValueError(f"expected 5D input (got {x.dim()}D input)")
