"""DARTS search space for FedNAS: mixed-op cells with architecture parameters α, genotype derivation, the discrete
evaluation network, the GDAS (Gumbel-softmax) search variant and the bi-level ``Architect`` incl. MiLeNAS' mixed-level
``step_v2``.

Parity: ``fedml_api/model/cv/darts/{operations,genotypes,model_search,model_search_gdas,model,architect,utils}.py``
(2 186 LoC in 11 files): 8 primitives per edge, ``steps = 4`` intermediate nodes → 14 edges
per cell, reductions at ``layers//3`` and ``2·layers//3``, ``α_normal, α_reduce ∈ ℝ^{14×8}`` initialised
``1e-3·randn``.  Design difference: α are registered ``nn.Parameter``s (prefix ``alphas_``) so they live in the SAME
flat arena row as the weights — the federated average of weights *and* α is then one K1 launch; the weight optimizer
simply skips them (``weight_parameters()``).
"""
from __future__ import annotations

from collections import namedtuple
from typing import List

import torch
import torch.nn.functional as F
from torch import nn

from ..ops.linear import TcLinear

Genotype = namedtuple("Genotype", "normal normal_concat reduce reduce_concat")

PRIMITIVES = ["none", "max_pool_3x3", "avg_pool_3x3", "skip_connect", "sep_conv_3x3", "sep_conv_5x5", "dil_conv_3x3",
              "dil_conv_5x5"]

# Published architecture constants (``darts/genotypes.py:16-93``): NASNet-A, AmoebaNet-A, DARTS first/second order, and
# the FedNAS search result.
NASNet = Genotype(
    normal=[("sep_conv_5x5", 1), ("sep_conv_3x3", 0), ("sep_conv_5x5", 0), ("sep_conv_3x3", 0), ("avg_pool_3x3", 1),
            ("skip_connect", 0), ("avg_pool_3x3", 0), ("avg_pool_3x3", 0), ("sep_conv_3x3", 1), ("skip_connect", 1)],
    normal_concat=[2, 3, 4, 5, 6],
    reduce=[("sep_conv_5x5", 1), ("sep_conv_7x7", 0), ("max_pool_3x3", 1), ("sep_conv_7x7", 0), ("avg_pool_3x3", 1),
            ("sep_conv_5x5", 0), ("skip_connect", 3), ("avg_pool_3x3", 2), ("sep_conv_3x3", 2), ("max_pool_3x3", 1)],
    reduce_concat=[4, 5, 6])
AmoebaNet = Genotype(
    normal=[("avg_pool_3x3", 0), ("max_pool_3x3", 1), ("sep_conv_3x3", 0), ("sep_conv_5x5", 2), ("sep_conv_3x3", 0),
            ("avg_pool_3x3", 3), ("sep_conv_3x3", 1), ("skip_connect", 1), ("skip_connect", 0), ("avg_pool_3x3", 1)],
    normal_concat=[4, 5, 6],
    reduce=[("avg_pool_3x3", 0), ("sep_conv_3x3", 1), ("max_pool_3x3", 0), ("sep_conv_7x7", 2), ("sep_conv_7x7", 0),
            ("avg_pool_3x3", 1), ("max_pool_3x3", 0), ("max_pool_3x3", 1), ("conv_7x1_1x7", 0), ("sep_conv_3x3", 5)],
    reduce_concat=[3, 4, 6])
DARTS_V1 = Genotype(
    normal=[("sep_conv_3x3", 1), ("sep_conv_3x3", 0), ("skip_connect", 0), ("sep_conv_3x3", 1), ("skip_connect", 0),
            ("sep_conv_3x3", 1), ("sep_conv_3x3", 0), ("skip_connect", 2)], normal_concat=[2, 3, 4, 5],
    reduce=[("max_pool_3x3", 0), ("max_pool_3x3", 1), ("skip_connect", 2), ("max_pool_3x3", 0), ("max_pool_3x3", 0),
            ("skip_connect", 2), ("skip_connect", 2), ("avg_pool_3x3", 0)], reduce_concat=[2, 3, 4, 5])
DARTS_V2 = Genotype(
    normal=[("sep_conv_3x3", 0), ("sep_conv_3x3", 1), ("sep_conv_3x3", 0), ("sep_conv_3x3", 1), ("sep_conv_3x3", 1),
            ("skip_connect", 0), ("skip_connect", 0), ("dil_conv_3x3", 2)], normal_concat=[2, 3, 4, 5],
    reduce=[("max_pool_3x3", 0), ("max_pool_3x3", 1), ("skip_connect", 2), ("max_pool_3x3", 1), ("max_pool_3x3", 0),
            ("skip_connect", 2), ("skip_connect", 2), ("max_pool_3x3", 1)], reduce_concat=[2, 3, 4, 5])
DARTS = DARTS_V2
FedNAS_V1 = Genotype(
    normal=[("sep_conv_3x3", 1), ("sep_conv_3x3", 0), ("sep_conv_3x3", 2), ("sep_conv_5x5", 0), ("sep_conv_3x3", 1),
            ("sep_conv_5x5", 3), ("dil_conv_5x5", 3), ("sep_conv_3x3", 4)], normal_concat=range(2, 6),
    reduce=[("max_pool_3x3", 0), ("skip_connect", 1), ("max_pool_3x3", 0), ("max_pool_3x3", 2), ("max_pool_3x3", 0),
            ("dil_conv_5x5", 1), ("max_pool_3x3", 0), ("dil_conv_5x5", 2)], reduce_concat=range(2, 6))


class Zero(nn.Module):
    def __init__(self, stride):
        super().__init__()
        self.stride = stride

    def forward(self, x):
        return x.mul(0.) if self.stride == 1 else x[:, :, ::self.stride, ::self.stride].mul(0.)


class FactorizedReduce(nn.Module):
    def __init__(self, C_in, C_out, affine=True):
        super().__init__()
        self.relu = nn.ReLU(inplace=False)
        self.conv_1 = nn.Conv2d(C_in, C_out // 2, 1, 2, 0, bias=False)
        self.conv_2 = nn.Conv2d(C_in, C_out // 2, 1, 2, 0, bias=False)
        self.bn = nn.BatchNorm2d(C_out, affine=affine)

    def forward(self, x):
        x = self.relu(x)
        return self.bn(torch.cat([self.conv_1(x), self.conv_2(x[:, :, 1:, 1:])], dim=1))


class ReLUConvBN(nn.Sequential):
    def __init__(self, C_in, C_out, k, stride, pad, affine=True):
        super().__init__(nn.ReLU(inplace=False), nn.Conv2d(C_in, C_out, k, stride, pad, bias=False),
                         nn.BatchNorm2d(C_out, affine=affine))


class DilConv(nn.Sequential):
    def __init__(self, C_in, C_out, k, stride, pad, dil, affine=True):
        super().__init__(nn.ReLU(inplace=False), nn.Conv2d(C_in, C_in, k, stride, pad, dilation=dil, groups=C_in, bias=False),
                         nn.Conv2d(C_in, C_out, 1, bias=False), nn.BatchNorm2d(C_out, affine=affine))


class SepConv(nn.Sequential):
    def __init__(self, C_in, C_out, k, stride, pad, affine=True):
        super().__init__(nn.ReLU(inplace=False), nn.Conv2d(C_in, C_in, k, stride, pad, groups=C_in, bias=False),
                         nn.Conv2d(C_in, C_in, 1, bias=False), nn.BatchNorm2d(C_in, affine=affine), nn.ReLU(inplace=False),
                         nn.Conv2d(C_in, C_in, k, 1, pad, groups=C_in, bias=False), nn.Conv2d(C_in, C_out, 1, bias=False),
                         nn.BatchNorm2d(C_out, affine=affine))


OPS = {
    "none": lambda C, s, a: Zero(s),
    "avg_pool_3x3": lambda C, s, a: nn.AvgPool2d(3, s, 1, count_include_pad=False),
    "max_pool_3x3": lambda C, s, a: nn.MaxPool2d(3, s, 1),
    "skip_connect": lambda C, s, a: nn.Identity() if s == 1 else FactorizedReduce(C, C, a),
    "sep_conv_3x3": lambda C, s, a: SepConv(C, C, 3, s, 1, a),
    "sep_conv_5x5": lambda C, s, a: SepConv(C, C, 5, s, 2, a),
    "dil_conv_3x3": lambda C, s, a: DilConv(C, C, 3, s, 2, 2, a),
    "dil_conv_5x5": lambda C, s, a: DilConv(C, C, 5, s, 4, 2, a),
    # only used by the NASNet / AmoebaNet evaluation genotypes (``operations.py:13-19``)
    "sep_conv_7x7": lambda C, s, a: SepConv(C, C, 7, s, 3, a),
    "conv_7x1_1x7": lambda C, s, a: nn.Sequential(
        nn.ReLU(inplace=False), nn.Conv2d(C, C, (1, 7), stride=(1, s), padding=(0, 3), bias=False),
        nn.Conv2d(C, C, (7, 1), stride=(s, 1), padding=(3, 0), bias=False), nn.BatchNorm2d(C, affine=a)),
}


class MixedOp(nn.Module):
    def __init__(self, C, stride):
        super().__init__()
        self._ops = nn.ModuleList()
        for prim in PRIMITIVES:
            op = OPS[prim](C, stride, False)
            if "pool" in prim:
                op = nn.Sequential(op, nn.BatchNorm2d(C, affine=False))
            self._ops.append(op)

    def forward(self, x, weights):
        return sum(w * op(x) for w, op in zip(weights, self._ops))


class SearchCell(nn.Module):
    def __init__(self, steps, multiplier, C_pp, C_p, C, reduction, reduction_prev):
        super().__init__()
        self.reduction, self._steps, self._multiplier = reduction, steps, multiplier
        self.preprocess0 = FactorizedReduce(C_pp, C, False) if reduction_prev else ReLUConvBN(C_pp, C, 1, 1, 0, False)
        self.preprocess1 = ReLUConvBN(C_p, C, 1, 1, 0, False)
        self._ops = nn.ModuleList()
        for i in range(steps):
            for j in range(2 + i):
                self._ops.append(MixedOp(C, 2 if reduction and j < 2 else 1))

    def forward(self, s0, s1, weights):
        states = [self.preprocess0(s0), self.preprocess1(s1)]
        off = 0
        for _ in range(self._steps):
            states.append(sum(self._ops[off + j](h, weights[off + j]) for j, h in enumerate(states)))
            off += len(states) - 1
        return torch.cat(states[-self._multiplier:], dim=1)


class Network(nn.Module):
    """DARTS search network (``model_search.py:172-259``)."""

    def __init__(self, C, num_classes, layers, criterion=None, steps=4, multiplier=4, stem_multiplier=3):
        super().__init__()
        self._C, self._num_classes, self._layers, self._steps, self._multiplier = C, num_classes, layers, steps, multiplier
        self._criterion = criterion if criterion is not None else nn.CrossEntropyLoss()
        C_curr = stem_multiplier * C
        self.stem = nn.Sequential(nn.Conv2d(3, C_curr, 3, padding=1, bias=False), nn.BatchNorm2d(C_curr))
        C_pp, C_p, C_curr = C_curr, C_curr, C
        self.cells = nn.ModuleList()
        reduction_prev = False
        for i in range(layers):
            reduction = layers >= 3 and i in (layers // 3, 2 * layers // 3)
            if reduction:
                C_curr *= 2
            cell = SearchCell(steps, multiplier, C_pp, C_p, C_curr, reduction, reduction_prev)
            reduction_prev = reduction
            self.cells.append(cell)
            C_pp, C_p = C_p, multiplier * C_curr
        self.global_pooling = nn.AdaptiveAvgPool2d(1)
        self.classifier = TcLinear(C_p, num_classes)
        k = sum(2 + i for i in range(steps))
        self.alphas_normal = nn.Parameter(1e-3 * torch.randn(k, len(PRIMITIVES)))
        self.alphas_reduce = nn.Parameter(1e-3 * torch.randn(k, len(PRIMITIVES)))

    def reset_parameters(self):
        with torch.no_grad():
            self.alphas_normal.copy_(1e-3 * torch.randn_like(self.alphas_normal))
            self.alphas_reduce.copy_(1e-3 * torch.randn_like(self.alphas_reduce))

    def arch_parameters(self) -> List[nn.Parameter]:
        return [self.alphas_normal, self.alphas_reduce]

    def weight_parameters(self) -> List[nn.Parameter]:
        return [p for n, p in self.named_parameters() if not n.startswith("alphas_")]

    def _edge_weights(self, alphas):
        return F.softmax(alphas, dim=-1)

    def forward(self, x):
        s0 = s1 = self.stem(x)
        for cell in self.cells:
            w = self._edge_weights(self.alphas_reduce if cell.reduction else self.alphas_normal)
            s0, s1 = s1, cell(s0, s1, w)
        return self.classifier(self.global_pooling(s1).flatten(1))

    def _loss(self, x, target):
        return self._criterion(self(x), target)

    def new(self):
        m = type(self)(self._C, self._num_classes, self._layers, self._criterion, self._steps, self._multiplier)
        for a, b in zip(m.arch_parameters(), self.arch_parameters()):
            a.data.copy_(b.data)
        return m

    def genotype(self) -> Genotype:
        def parse(weights):
            gene, n, start = [], 2, 0
            none = PRIMITIVES.index("none")
            for i in range(self._steps):
                W = weights[start:start + n]
                edges = sorted(range(i + 2), key=lambda e: -max(W[e][k] for k in range(len(W[e])) if k != none))[:2]
                for j in edges:
                    kb = max((k for k in range(len(W[j])) if k != none), key=lambda k: W[j][k])
                    gene.append((PRIMITIVES[kb], j))
                start += n
                n += 1
            return gene
        concat = range(2 + self._steps - self._multiplier, self._steps + 2)
        return Genotype(parse(F.softmax(self.alphas_normal, -1).tolist()), concat,
                        parse(F.softmax(self.alphas_reduce, -1).tolist()), concat)


class Network_GumbelSoftmax(Network):
    """GDAS search: hard Gumbel-softmax sample per edge (``model_search_gdas.py``)."""

    tau = 5.0

    def set_tau(self, tau):
        self.tau = tau

    def _edge_weights(self, alphas):
        return F.gumbel_softmax(alphas, tau=self.tau, hard=True, dim=-1) if self.training else F.softmax(alphas, -1)


class EvalCell(nn.Module):
    def __init__(self, genotype, C_pp, C_p, C, reduction, reduction_prev):
        super().__init__()
        self.preprocess0 = FactorizedReduce(C_pp, C) if reduction_prev else ReLUConvBN(C_pp, C, 1, 1, 0)
        self.preprocess1 = ReLUConvBN(C_p, C, 1, 1, 0)
        ops_idx = genotype.reduce if reduction else genotype.normal
        self._concat = list(genotype.reduce_concat if reduction else genotype.normal_concat)
        self.multiplier, self.reduction = len(self._concat), reduction
        self._ops = nn.ModuleList(OPS[name](C, 2 if reduction and idx < 2 else 1, True) for name, idx in ops_idx)
        self._indices = [idx for _, idx in ops_idx]

    def forward(self, s0, s1, drop_prob: float = 0.0):
        states = [self.preprocess0(s0), self.preprocess1(s1)]
        for i in range(len(self._ops) // 2):
            op1, op2 = self._ops[2 * i], self._ops[2 * i + 1]
            h1, h2 = op1(states[self._indices[2 * i]]), op2(states[self._indices[2 * i + 1]])
            if self.training and drop_prob > 0.0:   # ``model.py:52-57``: identity edges are never dropped
                if not isinstance(op1, nn.Identity):
                    h1 = drop_path(h1, drop_prob)
                if not isinstance(op2, nn.Identity):
                    h2 = drop_path(h2, drop_prob)
            states.append(h1 + h2)
        return torch.cat([states[i] for i in self._concat], dim=1)


def drop_path(x: torch.Tensor, drop_prob: float) -> torch.Tensor:
    """Per-sample stochastic depth (``darts/utils.py:82-88``)."""
    if drop_prob > 0.0:
        keep = 1.0 - drop_prob
        mask = torch.empty(x.size(0), 1, 1, 1, device=x.device, dtype=x.dtype).bernoulli_(keep)
        x = x / keep * mask
    return x


def count_parameters_in_MB(model: nn.Module) -> float:
    """``darts/utils.py:62-63``: parameters excluding the auxiliary head, in units of 1e6."""
    return sum(p.numel() for n, p in model.named_parameters() if "auxiliary" not in n) / 1e6


def genotype_to_dot(genotype: Genotype, which: str = "normal") -> str:
    """Graphviz DOT text of a cell (``darts/visualize.py:6-44`` renders the same graph through the graphviz package,
    which this image does not ship; ``dot -Tpdf`` on the returned text gives the identical picture)."""
    ops = genotype.normal if which == "normal" else genotype.reduce
    lines = ["digraph cell {", "  rankdir=LR;", '  node [shape=rect, style=filled, fillcolor=lightblue];',
             '  "c_{k-2}" [fillcolor=darkseagreen2]; "c_{k-1}" [fillcolor=darkseagreen2];']
    steps = len(ops) // 2
    for i in range(steps):
        for k in (2 * i, 2 * i + 1):
            name, j = ops[k]
            src = "c_{k-2}" if j == 0 else "c_{k-1}" if j == 1 else str(j - 2)
            lines.append(f'  "{src}" -> "{i}" [label="{name}"];')
    lines.append('  "c_{k}" [fillcolor=palegoldenrod];')
    lines += [f'  "{i}" -> "c_{{k}}";' for i in range(steps)]
    lines.append("}")
    return "\n".join(lines)


class NetworkCIFAR(nn.Module):
    """Discrete evaluation network built from a genotype (``model.py``); auxiliary head optional."""

    def __init__(self, C, num_classes, layers, auxiliary, genotype, stem_multiplier=3):
        super().__init__()
        self._layers, self._auxiliary, self.drop_path_prob = layers, auxiliary, 0.0
        C_curr = stem_multiplier * C
        self.stem = nn.Sequential(nn.Conv2d(3, C_curr, 3, padding=1, bias=False), nn.BatchNorm2d(C_curr))
        C_pp, C_p, C_curr = C_curr, C_curr, C
        self.cells = nn.ModuleList()
        reduction_prev, C_aux = False, None
        for i in range(layers):
            reduction = layers >= 3 and i in (layers // 3, 2 * layers // 3)
            if reduction:
                C_curr *= 2
            cell = EvalCell(genotype, C_pp, C_p, C_curr, reduction, reduction_prev)
            reduction_prev = reduction
            self.cells.append(cell)
            C_pp, C_p = C_p, cell.multiplier * C_curr
            if i == 2 * layers // 3:
                C_aux = C_p
        if auxiliary and C_aux is not None:
            self.auxiliary_head = nn.Sequential(nn.ReLU(inplace=True), nn.AdaptiveAvgPool2d(2), nn.Conv2d(C_aux, 128, 1, bias=False),
                                                nn.BatchNorm2d(128), nn.ReLU(inplace=True), nn.Flatten(), TcLinear(128 * 4, num_classes))
        self.global_pooling = nn.AdaptiveAvgPool2d(1)
        self.classifier = TcLinear(C_p, num_classes)

    def forward(self, x):
        logits_aux = None
        s0 = s1 = self.stem(x)
        for i, cell in enumerate(self.cells):
            s0, s1 = s1, cell(s0, s1, self.drop_path_prob)
            if i == 2 * self._layers // 3 and self._auxiliary and self.training:
                logits_aux = self.auxiliary_head(s1)
        return self.classifier(self.global_pooling(s1).flatten(1)), logits_aux


class NetworkImageNet(nn.Module):
    """ImageNet-sized evaluation network (``model.py:161-217``): two stride-2 stems (224² → 28²), cells as in
    ``NetworkCIFAR`` with ``reduction_prev = True`` for the first cell, 7×7 average pool, optional auxiliary head."""

    def __init__(self, C, num_classes, layers, auxiliary, genotype):
        super().__init__()
        self._layers, self._auxiliary, self.drop_path_prob = layers, auxiliary, 0.0
        self.stem0 = nn.Sequential(nn.Conv2d(3, C // 2, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(C // 2),
                                   nn.ReLU(inplace=True), nn.Conv2d(C // 2, C, 3, stride=2, padding=1, bias=False),
                                   nn.BatchNorm2d(C))
        self.stem1 = nn.Sequential(nn.ReLU(inplace=True), nn.Conv2d(C, C, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(C))
        C_pp, C_p, C_curr = C, C, C
        self.cells = nn.ModuleList()
        reduction_prev, C_aux = True, None
        for i in range(layers):
            reduction = layers >= 3 and i in (layers // 3, 2 * layers // 3)
            if reduction:
                C_curr *= 2
            cell = EvalCell(genotype, C_pp, C_p, C_curr, reduction, reduction_prev)
            reduction_prev = reduction
            self.cells.append(cell)
            C_pp, C_p = C_p, cell.multiplier * C_curr
            if i == 2 * layers // 3:
                C_aux = C_p
        if auxiliary and C_aux is not None:
            self.auxiliary_head = nn.Sequential(nn.ReLU(inplace=True), nn.AvgPool2d(5, stride=2, padding=0, count_include_pad=False),
                                                nn.Conv2d(C_aux, 128, 1, bias=False), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
                                                nn.Conv2d(128, 768, 2, bias=False), nn.ReLU(inplace=True), nn.Flatten(),
                                                TcLinear(768, num_classes))
        self.global_pooling = nn.AvgPool2d(7)
        self.classifier = TcLinear(C_p, num_classes)

    def forward(self, x):
        logits_aux = None
        s0 = self.stem0(x)
        s1 = self.stem1(s0)
        for i, cell in enumerate(self.cells):
            s0, s1 = s1, cell(s0, s1, self.drop_path_prob)
            if i == 2 * self._layers // 3 and self._auxiliary and self.training:
                logits_aux = self.auxiliary_head(s1)
        return self.classifier(self.global_pooling(s1).flatten(1)), logits_aux


class Architect:
    """α optimiser (``architect.py:16-99``): Adam(lr=arch_lr, betas=(.5,.999), wd=arch_wd) on the arch parameters.
    ``step`` = first-order DARTS (∇α L_val); ``step_v2`` = MiLeNAS mixed-level ``∇α L_val + λ_train·∇α L_train``."""

    def __init__(self, model: Network, criterion, args, device=None):
        self.model, self.criterion, self.args = model, criterion, args
        self.lambda_train = float(getattr(args, "lambda_train_regularizer", 1.0))
        self.lambda_valid = float(getattr(args, "lambda_valid_regularizer", 1.0))
        self.optimizer = torch.optim.Adam(model.arch_parameters(), lr=getattr(args, "arch_learning_rate", 3e-4),
                                          betas=(0.5, 0.999), weight_decay=getattr(args, "arch_weight_decay", 1e-3))

    def step(self, input_valid, target_valid):
        self.optimizer.zero_grad()
        loss = self.criterion(self.model(input_valid), target_valid)
        grads = torch.autograd.grad(loss, self.model.arch_parameters())
        for p, g in zip(self.model.arch_parameters(), grads):
            p.grad = g
        self.optimizer.step()
        return float(loss)

    def step_v2(self, input_train, target_train, input_valid, target_valid, lambda_train=None, lambda_valid=None):
        lt = self.lambda_train if lambda_train is None else lambda_train
        lv = self.lambda_valid if lambda_valid is None else lambda_valid
        self.optimizer.zero_grad()
        g_val = torch.autograd.grad(self.criterion(self.model(input_valid), target_valid), self.model.arch_parameters())
        g_tr = torch.autograd.grad(self.criterion(self.model(input_train), target_train), self.model.arch_parameters())
        for p, gv, gt in zip(self.model.arch_parameters(), g_val, g_tr):
            p.grad = lv * gv + lt * gt
        self.optimizer.step()
