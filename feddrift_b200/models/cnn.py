"""MNIST/EMNIST CNNs.

Parity: ``fedml_api/model/cv/cnn.py:5-68`` (CNN_OriginalFedAvg, 1 663 370 params)
and ``:71-136`` (CNN_DropOut, 1 199 882 params: conv3×3(1→32) → conv3×3(32→64)
with NO ReLU between the convs, maxpool2, dropout .25, fc 9216→128 ReLU,
dropout .5, fc →10, **Softmax before CrossEntropy**; flat 784 input reshaped).
The fully-connected layers are :class:`~feddrift_b200.ops.linear.TcLinear` (tcgen05 GEMM with fused bias/ReLU
epilogue on sm_100a, plain ``F.linear`` on CPU).  The 32→64 convolutions are
:class:`~feddrift_b200.ops.conv.TcConv2d`: implicit-GEMM forward / wgrad on the software-gather tcgen05 kernels
(``csrc/conv_igemm.cu``; Cin = 32) and the data gradient on the TMA-im2col GEMM mode (``csrc/gemm_tc.cu``; Cout = 64);
``FDB_NO_TC_CONV=1`` routes them to the library conv for A/B measurements.  The 1-channel stem convolution (9 or 25
multiply-adds per output) stays a library call.
"""
from __future__ import annotations

from torch import nn

import os

from ..ops.conv import TcConv2d
from ..ops.linear import TcLinear


def _conv(*a, **k):
    return TcConv2d(*a, **k)


class CNN_OriginalFedAvg(nn.Module):
    def __init__(self, only_digits: bool = True):
        super().__init__()
        self.only_digits = only_digits
        self.conv2d_1 = nn.Conv2d(1, 32, kernel_size=5, padding=2)
        self.max_pooling = nn.MaxPool2d(2, stride=2)
        self.conv2d_2 = _conv(32, 64, kernel_size=5, padding=2)
        self.flatten = nn.Flatten()
        self.linear_1 = TcLinear(3136, 512, activation="relu")
        self.linear_2 = TcLinear(512, 10 if only_digits else 62)
        self.softmax = nn.Softmax(dim=1)

    @staticmethod
    def stack_input(x):                     # [npairs, B, 784 | 28, 28] → [npairs, B, 1, 28, 28] (sim/stacked.py)
        return x.reshape(x.shape[0], x.shape[1], 1, 28, 28)

    def forward(self, x):
        if x.dim() == 2:
            x = x.reshape(x.shape[0], 28, 28)
        x = x.unsqueeze(1) if x.dim() == 3 else x
        x = self.max_pooling(self.conv2d_1(x))
        x = self.max_pooling(self.conv2d_2(x))
        x = self.linear_1(self.flatten(x))
        return self.softmax(self.linear_2(x))


class CNN_DropOut(nn.Module):
    def __init__(self, only_digits: bool = True):
        super().__init__()
        self.conv2d_1 = nn.Conv2d(1, 32, kernel_size=3)
        self.max_pooling = nn.MaxPool2d(2, stride=2)
        self.conv2d_2 = _conv(32, 64, kernel_size=3)
        self.dropout_1 = nn.Dropout(0.25)
        self.flatten = nn.Flatten()
        self.linear_1 = TcLinear(9216, 128, activation="relu")
        self.dropout_2 = nn.Dropout(0.5)
        self.linear_2 = TcLinear(128, 10 if only_digits else 62)
        self.softmax = nn.Softmax(dim=1)

    @staticmethod
    def stack_input(x):                     # [npairs, B, 784 | 28, 28] → [npairs, B, 1, 28, 28] (sim/stacked.py)
        return x.reshape(x.shape[0], x.shape[1], 1, 28, 28)

    def forward(self, x):
        x = x.reshape(x.shape[0], -1, 28, 28)     # 1 channel — or one per stacked (client, model) pair
        x = self.conv2d_2(self.conv2d_1(x))
        x = self.dropout_1(self.max_pooling(x))
        x = self.linear_1(self.flatten(x))
        x = self.linear_2(self.dropout_2(x))
        return self.softmax(x)
