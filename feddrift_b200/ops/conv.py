"""``TcConv2d`` — nn.Conv2d-compatible layer whose CUDA path is im2col (fused bf16 cast, ``csrc/conv_im2col.cu``) +
the hand-written tcgen05 GEMM with fused bias(+ReLU) epilogue (``csrc/gemm_tc.cu``).

* forward : cols = im2col(x) [B·Ho·Wo, Cin·kh·kw] bf16;  y = act(cols · Wᵀ + b) written as NHWC and returned as an
            NCHW *view* in channels_last memory format (no transpose kernel);
* backward: dcols = dy · W → dx = col2im(dcols) (gather form, no atomics);  dW = dyᵀ · cols;  db = Σ dy.  Both GEMMs
            use the kernel's MN-major operand descriptors, so no tensor is transposed in memory.

bf16 operands, fp32 accumulation in TMEM, fp32 master weights.  Ineligible shapes (groups/dilation ≠ 1, reduction
length Cin·kh·kw not a multiple of 8 or < 64 — e.g. a 1-channel stem) and CPU tensors use ``F.conv2d``.  State-dict
keys and the init law equal ``nn.Conv2d``'s.  Reference: cuDNN fp32 ``nn.Conv2d`` (``fedml_api/model/cv/cnn.py:110-117``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _ext

TC_CONV_CALLS = 0


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


class _TcConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, relu: bool):
        global TC_CONV_CALLS
        TC_CONV_CALLS += 1
        ext = _ext.load(required=True)
        B, C, H, W = x.shape
        Co, _, kh, kw = weight.shape
        Ho, Wo = (H + 2 * padding[0] - kh) // stride[0] + 1, (W + 2 * padding[1] - kw) // stride[1] + 1
        cols = ext.im2col_bf16(x.float(), kh, kw, stride[0], stride[1], padding[0], padding[1])
        wb = weight.reshape(Co, -1).to(torch.bfloat16).contiguous()
        y = ext.gemm_tn_bias_act(cols, wb, bias, bool(relu), True)           # [B·Ho·Wo, Co] fp32 == NHWC
        ctx.save_for_backward(cols, wb, y if relu else None)
        ctx.geom = (B, C, H, W, kh, kw, stride, padding, Ho, Wo, Co)
        ctx.relu, ctx.has_bias = relu, bias is not None
        return y.view(B, Ho, Wo, Co).permute(0, 3, 1, 2)                     # NCHW view, channels_last strides

    @staticmethod
    def backward(ctx, gy):
        ext = _ext.load(required=True)
        cols, wb, y = ctx.saved_tensors
        B, C, H, W, kh, kw, stride, padding, Ho, Wo, Co = ctx.geom
        g = gy.permute(0, 2, 3, 1).reshape(B * Ho * Wo, Co)                  # free when gy is channels_last
        if ctx.relu:
            g = g * (y > 0)
        gb = g.to(torch.bfloat16).contiguous()
        gx = gw = gbias = None
        if ctx.needs_input_grad[0]:
            dcols = ext.gemm_bias_act(gb, wb, False, True, None, False, True)               # dy · W  → [BHW, K] fp32
            gx = ext.col2im(dcols, B, C, H, W, kh, kw, stride[0], stride[1], padding[0], padding[1])
        if ctx.needs_input_grad[1]:
            gw = ext.gemm_bias_act(gb, cols, True, True, None, False, True)                 # dyᵀ · cols, no transposes
            gw = gw.view(Co, C, kh, kw)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gbias = g.sum(0)
        return gx, gw, gbias, None, None, None


class TcConv2d(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, padding=0, dilation=1, groups: int = 1,
                 bias: bool = True, activation: str = "none"):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _pair(kernel_size), _pair(stride), _pair(padding)
        self.dilation, self.groups, self.activation = _pair(dilation), groups, activation
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self) -> None:  # identical init law to nn.Conv2d
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.weight.shape[1] * self.kernel_size[0] * self.kernel_size[1]
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def _eligible(self, x: torch.Tensor) -> bool:
        k = self.weight[0].numel()
        return (x.is_cuda and x.dim() == 4 and self.groups == 1 and self.dilation == (1, 1) and k % 8 == 0 and k >= 64
                and self.out_channels % 8 == 0 and self.out_channels >= 16 and _ext.available())

    def forward(self, x):
        relu = self.activation == "relu"
        if self._eligible(x):
            return _TcConvFn.apply(x, self.weight, self.bias, self.stride, self.padding, relu)
        y = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        return F.relu(y) if relu else y

    def extra_repr(self) -> str:
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, activation={self.activation}")
