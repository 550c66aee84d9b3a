"""ResNet family: CIFAR ResNet-56/110 (bottleneck [6,6,6] / [12,12,12]), GroupNorm ResNet-18/34/50/101/152, and the
FedGKT client/server split (ResNet-8 edge model that also returns its stem feature map, ResNet-55 server model that
consumes feature maps).

Parity: ``fedml_api/model/cv/resnet.py:113-246``, ``resnet_gn.py``, ``resnet56_gkt/{resnet_client,resnet_server}.py``
(SURVEY §2.5).  One parametrised implementation (block type, depth list, stem, norm factory, input = image | feature
map) instead of four near-identical files; state-dict keys follow the torchvision convention
(``conv1 / bn1 / layerK.i.convJ / fc``) like the reference.  The classifier head is a :class:`TcLinear`.
"""
from __future__ import annotations

from typing import Callable, List

import torch
from torch import nn

from ..ops.conv import TcConv2d
from ..ops.linear import TcLinear
from .group_norm import GroupNorm2d


def conv3x3(i, o, stride=1):
    return TcConv2d(i, o, 3, stride, 1, bias=False)   # implicit-GEMM tcgen05 conv; library conv when ineligible / on CPU


def conv1x1(i, o, stride=1):
    return TcConv2d(i, o, 1, stride, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm=nn.BatchNorm2d):
        super().__init__()
        self.conv1, self.bn1 = conv3x3(inplanes, planes, stride), norm(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2, self.bn2 = conv3x3(planes, planes), norm(planes)
        self.downsample, self.stride = downsample, stride

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + idt)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm=nn.BatchNorm2d):
        super().__init__()
        self.conv1, self.bn1 = conv1x1(inplanes, planes), norm(planes)
        self.conv2, self.bn2 = conv3x3(planes, planes, stride), norm(planes)
        self.conv3, self.bn3 = conv1x1(planes, planes * 4), norm(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample, self.stride = downsample, stride

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + idt)


class ResNet(nn.Module):
    """``stem``: 'cifar' (3×3, 16 ch), 'imagenet' (7×7/2 + maxpool, 64 ch) or 'features' (input already is the
    16-channel stem feature map — FedGKT server).  ``KD=True`` returns ``(pooled_features, logits)``."""

    def __init__(self, block, layers: List[int], num_classes=10, stem="cifar", widths=(16, 32, 64), norm: Callable = None,
                 KD=False, zero_init_residual=False, return_stem_features=False):
        super().__init__()
        self.norm = norm if norm is not None else nn.BatchNorm2d
        self.KD, self.stem, self.return_stem_features = KD, stem, return_stem_features
        self.inplanes = widths[0] if stem != "imagenet" else 64
        if stem == "cifar":
            self.conv1, self.bn1 = conv3x3(3, self.inplanes), self.norm(self.inplanes)
        elif stem == "imagenet":
            self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
            self.bn1 = self.norm(64)
            self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.relu = nn.ReLU(inplace=True)
        strides = [1] + [2] * (len(layers) - 1)
        for i, (w, n, s) in enumerate(zip(widths, layers, strides), start=1):
            setattr(self, f"layer{i}", self._make_layer(block, w, n, s))
        self.num_layers = len(layers)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = TcLinear(widths[len(layers) - 1] * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, TcConv2d)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm, GroupNorm2d)) and getattr(m, "weight", None) is not None:
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)
                elif isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)

    def _make_layer(self, block, planes, blocks, stride):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride), self.norm(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, down, self.norm)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes, norm=self.norm) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        feats = None
        if self.stem != "features":
            x = self.relu(self.bn1(self.conv1(x)))
            if self.stem == "imagenet":
                x = self.maxpool(x)
            feats = x
        for i in range(1, self.num_layers + 1):
            x = getattr(self, f"layer{i}")(x)
        x_f = torch.flatten(self.avgpool(x), 1)
        logits = self.fc(x_f)
        if self.return_stem_features:
            return logits, feats
        return (x_f, logits) if self.KD else logits


def resnet56(class_num, **kw):
    return ResNet(Bottleneck, [6, 6, 6], class_num, **kw)


def resnet110(class_num, **kw):
    return ResNet(Bottleneck, [12, 12, 12], class_num, **kw)


# ---- FedGKT split (resnet56_gkt): the edge holds stem + one small stage, the server the deep remainder -----------
def resnet8_56(c, **kw):
    """Client model: returns ``(logits, stem_features [B,16,32,32])`` (``resnet_client.py:189-204``)."""
    return ResNet(Bottleneck, [2], c, widths=(16,), return_stem_features=True, **kw)


def resnet56_server(c, **kw):
    """Server model on feature maps: layers 1-3 + fc (``resnet_server.py:185-197``)."""
    return ResNet(Bottleneck, [6, 6, 6], c, stem="features", **kw)


# ---- GroupNorm ResNets (resnet_gn.py) ------------------------------------------------------------------------------
def _gn(groups_per_channel_div: int):
    def make(ch):
        g = max(1, ch // max(1, groups_per_channel_div))
        while ch % g != 0:
            g -= 1
        return GroupNorm2d(ch, g)
    return make


def _gn_resnet(block, layers, num_classes, group_norm, **kw):
    norm = _gn(group_norm) if group_norm and group_norm > 0 else nn.BatchNorm2d
    return ResNet(block, layers, num_classes, stem="imagenet", widths=(64, 128, 256, 512), norm=norm, **kw)


def resnet18(num_classes=1000, group_norm=2, **kw):
    return _gn_resnet(BasicBlock, [2, 2, 2, 2], num_classes, group_norm, **kw)


def resnet34(num_classes=1000, group_norm=2, **kw):
    return _gn_resnet(BasicBlock, [3, 4, 6, 3], num_classes, group_norm, **kw)


def resnet50(num_classes=1000, group_norm=2, **kw):
    return _gn_resnet(Bottleneck, [3, 4, 6, 3], num_classes, group_norm, **kw)


def resnet101(num_classes=1000, group_norm=2, **kw):
    return _gn_resnet(Bottleneck, [3, 4, 23, 3], num_classes, group_norm, **kw)


def resnet152(num_classes=1000, group_norm=2, **kw):
    return _gn_resnet(Bottleneck, [3, 8, 36, 3], num_classes, group_norm, **kw)
