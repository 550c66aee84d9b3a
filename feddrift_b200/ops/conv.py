"""``TcConv2d`` — nn.Conv2d-compatible layer whose CUDA path is the hand-written IMPLICIT-GEMM convolution on tcgen05
(``csrc/conv_igemm.cu``): forward, data gradient and weight gradient, NHWC activations end to end.

* forward : the producer warps gather each output pixel's (tap, channel) slice straight from the NHWC input into the
            tensor-core operand tile (no im2col matrix in memory), bias(+ReLU) fused in the epilogue; the result is
            returned as an NCHW *view* with channels_last strides, so a chain of convolutions never transposes;
* dgrad   : the same kernel gathers from dY with the transposed weight pack (strided convolutions skip non-integer taps);
* wgrad   : a GEMM whose reduction runs over the output pixels, both operands MN-major straight from the NHWC tensors,
            split over pixel ranges with coalesced ``red.global.add`` into the fp32 gradient.

fp32 activations / master weights, bf16 tensor-core operands, fp32 accumulation in TMEM.  Shapes the kernels do not cover
(groups / dilation ≠ 1, Cin or Cout not a multiple of 32 — e.g. the 1- or 3-channel stems) and CPU
tensors use ``F.conv2d``.  ``FDB_CONV_IM2COL=1`` selects the round-1 explicit-im2col + GEMM formulation (kept for A/B
measurements).  State-dict keys and the init law equal ``nn.Conv2d``'s.  Reference: cuDNN fp32 ``nn.Conv2d``
(``fedml_api/model/cv/cnn.py:110-117``).
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import _ext

TC_CONV_CALLS = 0
IGEMM_CALLS = {"fwd": 0, "dgrad": 0, "wgrad": 0}


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    """Contiguous NHWC view of an NCHW-logical tensor (free when it already is channels_last)."""
    return t.float().contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)


def igemm_eligible(cin: int, cout: int, stride, dilation, groups: int) -> bool:
    # forward needs Cin % 16 / Cout % 32; the data gradient swaps the roles → both multiples of 32
    return groups == 1 and tuple(dilation) == (1, 1) and stride[0] == stride[1] and cin % 32 == 0 and cout % 32 == 0


class _ConvIgemmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, relu: bool):
        global TC_CONV_CALLS
        TC_CONV_CALLS += 1
        IGEMM_CALLS["fwd"] += 1
        ext = _ext.load(required=True)
        xh = _nhwc(x)
        wq, wq_t = ext.conv_pack_weights(weight.detach().float().contiguous())     # both tensor-core packs, one launch
        y = ext.conv_igemm_fwd(xh, wq, bias.detach() if bias is not None else None, stride[0], padding[0], padding[1], bool(relu))
        ctx.save_for_backward(xh, wq_t, y if relu else None)
        ctx.weight_ref = weight if isinstance(weight, torch.nn.Parameter) else None
        ctx.geom = (stride, padding, tuple(weight.shape))
        ctx.relu, ctx.has_bias = relu, bias is not None
        return y.permute(0, 3, 1, 2)                                          # NCHW view, channels_last strides

    @staticmethod
    def backward(ctx, gy):
        ext = _ext.load(required=True)
        xh, wq_t, y = ctx.saved_tensors
        stride, padding, (Co, Ci, kh, kw) = ctx.geom
        g = _nhwc(gy)
        if ctx.relu:
            g = g * (y > 0)
        gx = gw = gbias = None
        if ctx.needs_input_grad[0]:
            IGEMM_CALLS["dgrad"] += 1
            gx = ext.conv_igemm_dgrad(g, wq_t, xh.shape[1], xh.shape[2], stride[0], padding[0], padding[1]).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            IGEMM_CALLS["wgrad"] += 1
            wp = ctx.weight_ref
            acc = wp.grad if (wp is not None and wp.grad is not None and wp.grad.is_contiguous() and wp.grad.dtype == torch.float32
                              and wp.grad.shape == wp.shape and wp.grad.is_cuda) else None
            # with a preallocated .grad (the federated executor binds every parameter's .grad to its slice of the flat, zeroed
            # gradient row) the kernel adds straight into it: no zero-fill launch, no AccumulateGrad add launch
            gw = ext.conv_igemm_wgrad(xh, g, kh, kw, stride[0], padding[0], padding[1], acc)           # OIHW
            if acc is not None:
                gw = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gbias = g.sum((0, 1, 2))
        return gx, gw, gbias, None, None, None


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

    def _use_igemm(self, x: torch.Tensor) -> bool:
        if not (x.is_cuda and x.dim() == 4 and _ext.available() and os.environ.get("FDB_CONV_IM2COL") != "1"
                and os.environ.get("FDB_NO_TC_CONV") != "1" and hasattr(_ext.load(), "conv_igemm_fwd")
                and igemm_eligible(self.in_channels, self.out_channels, self.stride, self.dilation, self.groups)):
            return False
        # FDB_CONV_POLICY: "igemm" = every eligible layer on the hand-written kernels (default); "auto" = leave layers with
        # fewer than 2048 output pixels to the library (a 128-pixel-row tile grid cannot fill 148 SMs there; measured
        # in profiles/conv_probe_r2.jsonl the deep 7×7 / 4×4 stages are 1.5–2.6× slower than cuDNN)
        if os.environ.get("FDB_CONV_POLICY", "igemm") == "auto":
            ho = (x.shape[2] + 2 * self.padding[0] - self.kernel_size[0]) // self.stride[0] + 1
            wo = (x.shape[3] + 2 * self.padding[1] - self.kernel_size[1]) // self.stride[1] + 1
            return x.shape[0] * ho * wo >= 2048
        return True

    def forward(self, x):
        relu = self.activation == "relu"
        if self._use_igemm(x):
            return _ConvIgemmFn.apply(x, self.weight, self.bias, self.stride, self.padding, relu)
        if self._eligible(x) and os.environ.get("FDB_CONV_IM2COL") == "1":
            return _TcConvFn.apply(x, self.weight, self.bias, self.stride, self.padding, relu)
        y = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        return F.relu(y) if relu else y

    def extra_repr(self) -> str:
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, activation={self.activation}")


def convert_convs_(module: nn.Module) -> nn.Module:
    """Replace every eligible ``nn.Conv2d`` of ``module`` (in place) by a :class:`TcConv2d` sharing the same parameters —
    state-dict keys, shapes and values are unchanged.  Used for the torchvision / model-zoo networks so that their 3×3 and
    1×1 body convolutions run on the implicit-GEMM tcgen05 kernels (stems with 1 or 3 input channels stay library convs)."""
    for name, child in list(module.named_children()):
        if type(child) is nn.Conv2d and child.padding_mode == "zeros" and not isinstance(child.padding, str) and \
                igemm_eligible(child.in_channels, child.out_channels, _pair(child.stride), _pair(child.dilation), child.groups):
            tc = TcConv2d(child.in_channels, child.out_channels, child.kernel_size, child.stride, child.padding, child.dilation,
                          child.groups, bias=child.bias is not None)
            tc.weight = child.weight
            tc.bias = child.bias
            setattr(module, name, tc)
        else:
            convert_convs_(child)
    return module
