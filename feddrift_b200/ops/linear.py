"""``TcLinear`` — nn.Linear-compatible layer whose CUDA path is the hand-written
tcgen05/TMEM/TMA GEMM (``csrc/gemm_tc.cu``) with fused bias(+ReLU) epilogue.

* forward  : Y = act(X · Wᵀ + b)         (A = X [B,K], B = W [N,K], both K-major)
* backward : dX = dY · W, dW = dYᵀ · X   (same kernel with MN-major operand
             descriptors: the row-major dY / W / X tensors are consumed as they
             are, no transpose kernels), db = column sum.

Compute dtype on CUDA is bf16 with fp32 accumulation in TMEM (master weights
stay fp32 in the parameter arena).  CPU / tiny shapes use ``F.linear`` in fp32.
State-dict keys (``weight``, ``bias``) equal ``nn.Linear``'s.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _ext

TC_CALLS = 0      # number of forward passes that went through the tcgen05 kernel (tests assert the path is live)
_MIN_TC_DIM = 64  # below this a GEMM is launch-bound; SIMT/cuBLASLt-free eager is fine


def _tc_eligible(x: torch.Tensor, weight: torch.Tensor) -> bool:
    if not (x.is_cuda and weight.is_cuda):
        return False
    n, k = weight.shape
    rows = x.numel() // max(k, 1)
    # TMA needs a 16-byte row pitch (K % 8 for bf16); K / M / N tails are zero-filled / clipped by the tensor maps
    return k % 8 == 0 and n % 8 == 0 and k >= _MIN_TC_DIM and n >= 16 and rows >= 16


class _TcLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu: bool):
        global TC_CALLS
        TC_CALLS += 1
        ext = _ext.load(required=True)
        x2 = x.reshape(-1, x.shape[-1])
        xb = x2.to(torch.bfloat16).contiguous()
        wb = weight.to(torch.bfloat16).contiguous()
        y = ext.gemm_tn_bias_act(xb, wb, bias if bias is not None else None, bool(relu), True)  # fp32 out
        ctx.save_for_backward(xb, wb, y if relu else None)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return y.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        ext = _ext.load(required=True)
        xb, wb, y = ctx.saved_tensors
        g = gy.reshape(-1, gy.shape[-1])
        if ctx.relu:
            g = g * (y > 0)
        gb = g.to(torch.bfloat16).contiguous()
        gx = gw = gbias = None
        if ctx.needs_input_grad[0]:
            # dX[B,K] = dY[B,N] · W[N,K]: W is consumed as an MN-major B operand ([reduction N, K contiguous]) — no transpose
            gx = ext.gemm_bias_act(gb, wb, False, True, None, False, True).reshape(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            # dW[N,K] = dYᵀ · X: both operands MN-major ([reduction B, N] and [reduction B, K]); any batch size
            gw = ext.gemm_bias_act(gb, xb, True, True, None, False, True)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gbias = g.sum(0)
        return gx, gw, gbias, None


class TcLinear(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, activation: str = "none"):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.activation = activation
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        self.reset_parameters()

    def reset_parameters(self) -> None:  # identical init law to nn.Linear
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features) if self.in_features > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        relu = self.activation == "relu"
        if _tc_eligible(x, self.weight) and _ext.available() and hasattr(_ext.load(), "gemm_tn_bias_act"):
            return _TcLinearFn.apply(x, self.weight, self.bias, relu)
        y = F.linear(x, self.weight, self.bias)
        return F.relu(y) if relu else y

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, activation={self.activation}"
