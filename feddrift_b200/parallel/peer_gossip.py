"""Multi-GPU decentralized mixing (K12): ``x_i ← Σ_j W_ij x_j`` with the neighbour exchange done inside the kernel over
NVLink peer memory (``csrc/gossip_peer.cu``) — one rank = one node of the topology.

    g = PeerGossip(P, device)              # collective: symmetric double buffer + epoch flags
    g.x.copy_(x0)                          # this node's vector (view of the current buffer)
    g.step(W[rank])                        # one DSGD mixing step (row of the mixing matrix); no NCCL, no host sync
    g.pushsum_step(W[rank])                # PushSum: ω rides along as element P; ``g.debiased()`` returns x/ω

Parity: ``standalone/decentralized/client_dsgd.py:88-102``, ``client_pushsum.py:104-129`` and the distributed template
``decentralized_framework/decentralized_worker_manager.py:29-46`` (one pickled MPI message per neighbour per step).
With one process it degenerates to the identity-weighted local update (``x ← W_ii·x``).
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist

from ..ops import _ext
from .arena import padded


class PeerGossip:
    def __init__(self, P: int, device, timeout_ms: int = 5000):
        self.P = P
        self.Pp = padded(P + 1, 4)   # +1: PushSum weight ω lives at index P
        self.device = torch.device(device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.timeout_ms = timeout_ms
        flag_f = 32
        if self.world == 1:
            self.buf = torch.zeros(2 * self.Pp + flag_f, dtype=torch.float32, device=self.device)
            base = [self.buf.data_ptr()]
            self.hdl = None
        else:
            import torch.distributed._symmetric_memory as symm_mem
            self.buf = symm_mem.empty(2 * self.Pp + flag_f, dtype=torch.float32, device=self.device)
            self.buf.zero_()
            self.hdl = symm_mem.rendezvous(self.buf, group=dist.group.WORLD.group_name)
            base = [int(p) for p in self.hdl.buffer_ptrs]
            torch.cuda.synchronize()
            dist.barrier()
        self.x_ptrs = base
        self.flag_ptrs = [p + 4 * 2 * self.Pp for p in base]
        self.grid_sync = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch, self.grid_base = 0, 0
        self._bufs = self.buf[: 2 * self.Pp].view(2, self.Pp)
        self._bufs[:, P] = 1.0   # ω = 1

    @property
    def x(self) -> torch.Tensor:
        """This node's current vector (a view — write the local SGD update into it in place)."""
        return self._bufs[self.epoch & 1, : self.P]

    @property
    def omega(self) -> torch.Tensor:
        return self._bufs[self.epoch & 1, self.P]

    def debiased(self) -> torch.Tensor:
        return self.x / self.omega

    def step(self, w_row: Sequence[float]) -> torch.Tensor:
        """One mixing step with this rank's row of the (row-stochastic) mixing matrix."""
        w = [float(v) for v in w_row]
        assert len(w) == self.world
        self.epoch += 1
        grid = _ext.load(required=True).gossip_mix_peer(self.x_ptrs, self.flag_ptrs, w, self.Pp, self.world, self.rank, self.grid_sync,
                                                        self.grid_base, self.epoch, self.timeout_ms, self.error_flag)
        self.grid_base += int(grid)
        return self.x

    pushsum_step = step   # ω is part of the mixed vector: the same kernel implements PushSum (column-stochastic W)

    def check(self) -> None:
        code = int(self.error_flag.item())
        if code:
            raise RuntimeError(f"peer gossip timed out (code {code})")
