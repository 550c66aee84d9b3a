"""Small tabular models of the drift experiments.

Parity: ``fedml_api/model/linear/lr.py:4-11`` (sigmoid *before* CrossEntropy —
the double squashing is reproduced on purpose) and
``fedml_api/model/fnn/fnn.py:4-15`` (Linear-ReLU-Linear; ``hidden = 2·in`` is
chosen by the caller, ``main_fedavg.py:215``).  State-dict keys match the
reference so ``model_params.pt`` files interoperate.

Both expose ``mlp_spec()`` so the engine can route them to the fused
``fed_round_small`` kernel (all clients × models of a GPU in one persistent
kernel) instead of eager module calls.
"""
from __future__ import annotations

import torch
from torch import nn


class LogisticRegression(nn.Module):
    def __init__(self, input_dim: int, output_dim: int):
        super().__init__()
        self.linear = nn.Linear(input_dim, output_dim)

    def forward(self, x):
        return torch.sigmoid(self.linear(x))

    def mlp_spec(self):
        return {"kind": "lr", "in": self.linear.in_features, "hidden": 0, "out": self.linear.out_features}


class FeedForwardNN(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, hidden_dim: int):
        super().__init__()
        if input_dim >= 64:
            # fnn-MNIST (784→1568→10): first layer on the tcgen05 GEMM with the ReLU fused into its epilogue
            from ..ops.linear import TcLinear
            self.fc1 = TcLinear(input_dim, hidden_dim, activation="relu")
            self.relu = nn.Identity()
        else:
            self.fc1 = nn.Linear(input_dim, hidden_dim)
            self.relu = nn.ReLU()
        self.fc2 = nn.Linear(hidden_dim, output_dim)

    def forward(self, x):
        return self.fc2(self.relu(self.fc1(x)))

    def mlp_spec(self):
        return {"kind": "fnn", "in": self.fc1.in_features, "hidden": self.fc1.out_features,
                "out": self.fc2.out_features}
