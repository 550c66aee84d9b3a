"""Name → model registry (the union of every ``create_model`` in ``fedml_experiments/**/main_*.py``)."""
from __future__ import annotations

from torch import nn


def build(name: str, output_dim: int, feature_dim: int = None, **kw) -> nn.Module:
    name = name.lower()
    if name == "lr":
        from .basic import LogisticRegression
        return LogisticRegression(feature_dim, output_dim)
    if name == "fnn":
        from .basic import FeedForwardNN
        return FeedForwardNN(feature_dim, output_dim, kw.get("hidden_dim", feature_dim * 2))
    if name == "cnn":
        from .cnn import CNN_DropOut
        return CNN_DropOut(only_digits=(output_dim == 10))
    if name in ("cnn_fedavg", "cnn_original"):
        from .cnn import CNN_OriginalFedAvg
        return CNN_OriginalFedAvg(only_digits=(output_dim == 10))
    if name == "rnn":
        from .rnn import RNN_OriginalFedAvg
        return RNN_OriginalFedAvg(**{k: v for k, v in kw.items() if k in ("embedding_dim", "vocab_size",
                                                                          "hidden_size", "per_position")})
    if name in ("rnn_stackoverflow", "lstm_nwp"):
        from .rnn import RNN_StackOverFlow
        return RNN_StackOverFlow()
    if name in ("resnet", "resnet18"):
        from .resnet_tv import resnet18
        return resnet18(num_classes=kw.get("num_classes", 1000 if kw.get("keep_imagenet_head") else output_dim),
                        small_input=kw.get("small_input", False))
    if name in ("densenet", "densenet121"):
        from .resnet_tv import densenet121
        return densenet121(num_classes=kw.get("num_classes", output_dim))
    if name == "resnet56":
        from .resnet import resnet56
        return resnet56(class_num=output_dim)
    if name == "resnet110":
        from .resnet import resnet110
        return resnet110(class_num=output_dim)
    if name in ("resnet18_gn", "resnet_gn"):
        from .resnet_gn import resnet18 as rgn18
        return rgn18(num_classes=output_dim, group_norm=kw.get("group_norm", 2))
    if name == "mobilenet":
        from .mobilenet import mobilenet
        return mobilenet(class_num=output_dim)
    if name == "vfl_feature":
        from .vfl import VFLFeatureExtractor
        return VFLFeatureExtractor(feature_dim, output_dim)
    raise ValueError(f"unknown model {name!r}")
