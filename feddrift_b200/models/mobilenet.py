"""MobileNet-v1 (depthwise-separable convolutions) for 32×32 inputs (parity: ``fedml_api/model/cv/mobilenet.py:15-209``)."""
from __future__ import annotations

from torch import nn

from ..ops.linear import TcLinear


class DepthSeperabelConv2d(nn.Module):
    def __init__(self, i, o, kernel_size, **kw):
        super().__init__()
        self.depthwise = nn.Sequential(nn.Conv2d(i, i, kernel_size, groups=i, **kw), nn.BatchNorm2d(i), nn.ReLU(inplace=True))
        self.pointwise = nn.Sequential(nn.Conv2d(i, o, 1), nn.BatchNorm2d(o), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.pointwise(self.depthwise(x))


class BasicConv2d(nn.Module):
    def __init__(self, i, o, kernel_size, **kw):
        super().__init__()
        self.conv, self.bn, self.relu = nn.Conv2d(i, o, kernel_size, **kw), nn.BatchNorm2d(o), nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu(self.bn(self.conv(x)))


class MobileNet(nn.Module):
    def __init__(self, width_multiplier=1, class_num=100):
        super().__init__()
        a = width_multiplier
        c = lambda n: int(n * a)  # noqa: E731
        ds = DepthSeperabelConv2d
        self.stem = nn.Sequential(BasicConv2d(3, c(32), 3, padding=1, bias=False), ds(c(32), c(64), 3, padding=1, bias=False))
        self.conv1 = nn.Sequential(ds(c(64), c(128), 3, stride=2, padding=1, bias=False), ds(c(128), c(128), 3, padding=1, bias=False))
        self.conv2 = nn.Sequential(ds(c(128), c(256), 3, stride=2, padding=1, bias=False), ds(c(256), c(256), 3, padding=1, bias=False))
        self.conv3 = nn.Sequential(ds(c(256), c(512), 3, stride=2, padding=1, bias=False),
                                   *[ds(c(512), c(512), 3, padding=1, bias=False) for _ in range(5)])
        self.conv4 = nn.Sequential(ds(c(512), c(1024), 3, stride=2, padding=1, bias=False), ds(c(1024), c(1024), 3, padding=1, bias=False))
        self.fc = TcLinear(c(1024), class_num)
        self.avg = nn.AdaptiveAvgPool2d(1)

    def forward(self, x):
        x = self.conv4(self.conv3(self.conv2(self.conv1(self.stem(x)))))
        return self.fc(self.avg(x).flatten(1))


def mobilenet(alpha=1, class_num=100):
    return MobileNet(alpha, class_num)
