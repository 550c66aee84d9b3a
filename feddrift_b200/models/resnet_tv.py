"""torchvision architectures used by the fMoW / CIFAR drift configs (``main_fedavg.py:219-223``): ResNet-18 and
DenseNet-121.  The reference loads ImageNet weights (``pretrained=True``) and "re-initialises" by reloading them
(``model/utils.py:10-18``); there is no network here, so both are random-init (BASELINE.json prescribes that)."""
from __future__ import annotations

import torchvision
from torch import nn


def resnet18(num_classes: int = 1000, small_input: bool = False) -> nn.Module:
    m = torchvision.models.resnet18(weights=None, num_classes=num_classes)
    if small_input:  # CIFAR-sized inputs: 3×3 stem, no max-pool
        m.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        m.maxpool = nn.Identity()
    from ..ops.conv import convert_convs_
    return convert_convs_(m)     # body convolutions → implicit-GEMM tcgen05 kernels (same state-dict keys)


def densenet121(num_classes: int = 1000) -> nn.Module:
    from ..ops.conv import convert_convs_
    return convert_convs_(torchvision.models.densenet121(weights=None, num_classes=num_classes))
