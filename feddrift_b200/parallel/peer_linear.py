"""K2 — consumer-pull model broadcast fused into the first GEMM.

The reference ships every model's full ``state_dict`` to every rank every round
(``FedAvgEnsServerManager.py:28-30,66-67`` → ``mpi_send_thread.py:27``) and the client ``load_state_dict``s it before
its first layer runs (``FedAvgEnsTrainerSoftCluster.py:53-57``).  Here the bf16 weight matrices live ONCE, in the owner
GPU's symmetric-memory arena; a consumer GPU runs ``Y = act(X·Wᵀ + b)`` with the tcgen05 GEMM whose TMA producer
addresses the owner's copy directly (peer-mapped pointer in the tensor map), so the weights cross NVLink tile by tile
inside the GEMM's pipeline and never exist as a local tensor.  No NCCL call on the hot path.

    store = PeerWeights({"fc1": (1568, 784), "fc2": (16, 1568)}, device)   # collective (rendezvous)
    store.publish("fc1", W)            # owner writes its bf16 copy
    store.fence()                      # symmetric-memory barrier (cold: once per model update)
    y = store.linear(x, "fc1", owner=0, bias=b, relu=True)
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import _ext


class PeerWeights:
    def __init__(self, shapes: Dict[str, Tuple[int, int]], device):
        self.device = torch.device(device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.shapes = dict(shapes)
        self.offsets: Dict[str, int] = {}
        off = 0
        for name, (n, k) in self.shapes.items():
            if k % 8:
                raise ValueError(f"{name}: K={k} must be a multiple of 8 (16-byte TMA row pitch)")
            self.offsets[name] = off
            off += (n * k + 127) // 128 * 128  # 256-byte aligned slabs
        self.numel = max(off, 128)
        if self.world == 1:
            self.buf = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.device)
            self.ptrs = [self.buf.data_ptr()]
            self.hdl = None
        else:
            import torch.distributed._symmetric_memory as symm_mem
            self.buf = symm_mem.empty(self.numel, dtype=torch.bfloat16, device=self.device)
            self.buf.zero_()
            self.hdl = symm_mem.rendezvous(self.buf, group=dist.group.WORLD.group_name)
            self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
            torch.cuda.synchronize()
            dist.barrier()

    def view(self, name: str) -> torch.Tensor:
        n, k = self.shapes[name]
        o = self.offsets[name]
        return self.buf[o:o + n * k].view(n, k)

    def publish(self, name: str, weight: torch.Tensor) -> None:
        """Owner side: cast-copy the fp32 master weights into this rank's bf16 slab."""
        self.view(name).copy_(weight)

    def fence(self) -> None:
        """Make published weights visible to peers (device-side symmetric-memory barrier; no host sync)."""
        if self.hdl is not None:
            self.hdl.barrier()

    def ptr(self, name: str, owner: int) -> int:
        return self.ptrs[owner] + 2 * self.offsets[name]

    def linear(self, x: torch.Tensor, name: str, owner: int, bias: Optional[torch.Tensor] = None, relu: bool = False,
               out_fp32: bool = True) -> torch.Tensor:
        n, k = self.shapes[name]
        xb = x.reshape(-1, k).to(torch.bfloat16).contiguous()
        y = _ext.load(required=True).gemm_tn_bias_act_peer(xb, self.ptr(name, owner), n, bias, bool(relu), bool(out_fp32))
        return y.reshape(*x.shape[:-1], n)
