"""Gossip topologies (mixing matrices) for decentralized FL.

Parity: ``fedml_core/distributed/topology/{base,symmetric,asymmetric}_topology_manager.py``.
The reference builds ring ∪ Watts–Strogatz(k, p=0) graphs through networkx
(``nx.to_numpy_matrix`` — removed in networkx ≥ 3).  With rewiring probability
0 a Watts–Strogatz graph is exactly the ring lattice where node i links to
i±1 … i±⌊k/2⌋, so the adjacency is built directly and vectorised here; the
mixing matrix is also exported as a torch tensor for the on-device gossip
kernel (``ops.gossip_mix``).
"""
from __future__ import annotations

import abc
from typing import List

import numpy as np


def ring_lattice(n: int, k: int) -> np.ndarray:
    """Adjacency of Watts–Strogatz(n, k, p=0): each node joined to its k//2 nearest on each side."""
    adj = np.zeros((n, n), dtype=np.float32)
    if n <= 1:
        return adj
    idx = np.arange(n)
    for d in range(1, k // 2 + 1):
        adj[idx, (idx + d) % n] = 1
        adj[idx, (idx - d) % n] = 1
    np.fill_diagonal(adj, 0)
    return adj


class BaseTopologyManager(abc.ABC):
    n: int
    topology: np.ndarray

    @abc.abstractmethod
    def generate_topology(self) -> None:
        ...

    @abc.abstractmethod
    def get_in_neighbor_weights(self, node_index: int):
        ...

    @abc.abstractmethod
    def get_out_neighbor_weights(self, node_index: int):
        ...

    def get_in_neighbor_idx_list(self, node_index: int) -> List[int]:
        w = np.asarray(self.get_in_neighbor_weights(node_index))
        return [int(i) for i in np.nonzero(w > 0)[0] if i != node_index]

    def get_out_neighbor_idx_list(self, node_index: int) -> List[int]:
        w = np.asarray(self.get_out_neighbor_weights(node_index))
        return [int(i) for i in np.nonzero(w > 0)[0] if i != node_index]

    def mixing_matrix(self):
        import torch
        return torch.from_numpy(np.asarray(self.topology, dtype=np.float32))


class SymmetricTopologyManager(BaseTopologyManager):
    """ring ∪ ring-lattice(neighbor_num) with self loops, rows normalised to sum 1."""

    def __init__(self, n: int, neighbor_num: int = 2):
        self.n = n
        self.neighbor_num = neighbor_num
        self.topology = np.zeros((0, 0), dtype=np.float32)

    def generate_topology(self) -> None:
        adj = np.maximum(ring_lattice(self.n, 2), ring_lattice(self.n, int(self.neighbor_num)))
        np.fill_diagonal(adj, 1)
        self.topology = adj / adj.sum(axis=1, keepdims=True)

    def get_in_neighbor_weights(self, node_index: int):
        return [] if node_index >= self.n else self.topology[node_index]

    def get_out_neighbor_weights(self, node_index: int):
        return [] if node_index >= self.n else self.topology[node_index]


class AsymmetricTopologyManager(BaseTopologyManager):
    """Symmetric base + random directed extra links (never both directions of a pair), row-normalised.
    In-weights are the column of the matrix (asymmetric_topology_manager.py:76-82)."""

    def __init__(self, n: int, undirected_neighbor_num: int = 3, out_directed_neighbor: int = 3, rng=None):
        self.n = n
        self.undirected_neighbor_num = undirected_neighbor_num
        self.out_directed_neighbor = out_directed_neighbor
        self.rng = rng if rng is not None else np.random
        self.topology = np.zeros((0, 0), dtype=np.float32)

    def generate_topology(self) -> None:
        n = self.n
        adj = np.maximum(ring_lattice(n, 2), ring_lattice(n, int(self.undirected_neighbor_num)))
        np.fill_diagonal(adj, 1)
        taken = set()
        for i in range(n):
            zeros = np.nonzero(adj[i] == 0)[0]
            coin = self.rng.randint(2, size=len(zeros))
            for j, pick in zip(zeros, coin):
                if pick == 1 and (j * n + i) not in taken:
                    adj[i, j] = 1
                    taken.add(i * n + j)
        self.topology = adj / adj.sum(axis=1, keepdims=True)

    def get_in_neighbor_weights(self, node_index: int):
        return [] if node_index >= self.n else self.topology[:, node_index]

    def get_out_neighbor_weights(self, node_index: int):
        return [] if node_index >= self.n else self.topology[node_index]
