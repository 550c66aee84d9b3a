"""Vertical-FL party models (parity: ``fedml_api/model/finance/*``): torch feature extractor /
classifier and the numpy-in/out ``DenseModel`` / ``LocalModel`` with embedded SGD(momentum 0.9, wd 0.01)
(``vfl_models_standalone.py:6-72``)."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


class VFLFeatureExtractor(nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.classifier = nn.Sequential(nn.Linear(input_dim, output_dim), nn.LeakyReLU())
        self.output_dim = output_dim

    def forward(self, x):
        return self.classifier(x)

    def get_output_dim(self):
        return self.output_dim


class VFLClassifier(nn.Module):
    def __init__(self, input_dim, output_dim, bias=True):
        super().__init__()
        self.classifier = nn.Sequential(nn.Linear(input_dim, output_dim, bias=bias))

    def forward(self, x):
        return self.classifier(x)


class _NumpyModel(nn.Module):
    """numpy ↔ torch bridge with an embedded optimizer: ``forward(np) -> np``, ``backward(x_np, grad_np)`` applies
    the upstream gradient and takes one SGD step; ``DenseModel.backward`` also returns dL/dx for the layer below."""

    def __init__(self, learning_rate, device="cpu"):
        super().__init__()
        self.lr, self.dev = learning_rate, torch.device(device)

    def _opt(self):
        self.optimizer = torch.optim.SGD(self.parameters(), momentum=0.9, weight_decay=0.01, lr=self.lr)

    def forward(self, x):
        x = torch.as_tensor(np.asarray(x), dtype=torch.float32, device=self.dev)
        with torch.no_grad():
            return self.net(x).cpu().numpy()

    def _backward(self, x, grads, need_input_grad):
        x = torch.as_tensor(np.asarray(x), dtype=torch.float32, device=self.dev).requires_grad_(need_input_grad)
        g = torch.as_tensor(np.asarray(grads), dtype=torch.float32, device=self.dev)
        out = self.net(x)
        self.optimizer.zero_grad()
        out.backward(g)
        gx = x.grad.cpu().numpy() if need_input_grad else None
        self.optimizer.step()
        return gx


class DenseModel(_NumpyModel):
    def __init__(self, input_dim, output_dim, learning_rate=0.01, bias=True, device="cpu"):
        super().__init__(learning_rate, device)
        self.net = nn.Sequential(nn.Linear(input_dim, output_dim, bias=bias)).to(self.dev)
        self._opt()

    def backward(self, x, grads):
        return self._backward(x, grads, True)


class LocalModel(_NumpyModel):
    def __init__(self, input_dim, output_dim, learning_rate, device="cpu"):
        super().__init__(learning_rate, device)
        self.net = nn.Sequential(nn.Linear(input_dim, output_dim), nn.LeakyReLU()).to(self.dev)
        self.output_dim = output_dim
        self._opt()

    def backward(self, x, grads):
        self._backward(x, grads, False)

    def get_output_dim(self):
        return self.output_dim
