"""Multi-GPU per-cluster FedAvg aggregation + model broadcast in ONE kernel over NVLink peer memory
(``csrc/aggregate_peer.cu``): clients are sharded over the ranks, every rank ends with identical cluster models.

    agg = PeerAggregator(num_models, P, device)          # collective: allocates + rendezvous symmetric buffers
    theta = agg.aggregate(client_rows [C_local, M, P], n [C_local, M])   # -> this rank's [M, P] view of the models

No NCCL call on this path: the rendezvous (cold, once) uses ``torch.distributed._symmetric_memory``; the hot path is a
cooperative kernel doing peer loads/stores + ``st.release.sys`` / ``ld.acquire.sys`` epoch flags.  With one process
(``world == 1``) it degenerates to the single-GPU K1 kernel.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import ops
from ..ops import _ext
from .arena import padded


class PeerAggregator:
    def __init__(self, num_models: int, P: int, device, theta_init: Optional[torch.Tensor] = None):
        self.M, self.P = num_models, P
        self.Pp = padded(P, 4)
        self.device = torch.device(device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if self.world == 1:
            self.theta_full = torch.zeros(num_models, self.Pp, device=self.device)
            self.nvls = False
        else:
            import torch.distributed._symmetric_memory as symm_mem
            mp = num_models * self.Pp
            tot_f = (self.world * num_models + 31) // 32 * 32
            flag_f = max((3 * self.world + 31) // 32 * 32, 32)
            self.buf = symm_mem.empty(2 * mp + tot_f + flag_f, dtype=torch.float32, device=self.device)
            self.buf.zero_()
            self.hdl = symm_mem.rendezvous(self.buf, group=dist.group.WORLD.group_name)
            base = [int(p) for p in self.hdl.buffer_ptrs]
            self.part_ptrs = base
            self.theta_ptrs = [p + 4 * mp for p in base]
            self.tot_ptrs = [p + 4 * 2 * mp for p in base]
            self.flag_ptrs = [p + 4 * (2 * mp + tot_f) for p in base]
            # NVLS: if the allocation has a multicast mapping, phase 2 uses multimem.ld_reduce / multimem.st (in-switch
            # reduction + replication) instead of W peer loads + W peer stores per element
            import os
            mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
            # measured (profiles/peer_agg_r1.jsonl): in-switch reduction wins from 4 GPUs up (0.35 vs 0.41 ms at N = 8 for
            # ResNet-18 × 2 clusters) and loses at N = 2 (0.55 vs 0.45 ms), where a peer load is a single hop anyway
            force = os.environ.get("FDB_NVLS")
            self.nvls = bool(mc) and (force == "1" or (force != "0" and self.world >= 4))
            self.mc_part, self.mc_theta = (mc, mc + 4 * mp) if self.nvls else (0, 0)
            self.theta_full = self.buf[mp:2 * mp].view(num_models, self.Pp)
            self.grid_sync = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.error_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.epoch, self.grid_base = 0, 0
            torch.cuda.synchronize()
            dist.barrier()
        self.theta = self.theta_full[:, :P]
        if theta_init is not None:
            self.theta.copy_(theta_init)

    def aggregate(self, client_rows: torch.Tensor, n: torch.Tensor) -> torch.Tensor:
        """``client_rows`` [C_local, M, P(p)] fp32, ``n`` [C_local, M] weights (0 = did not train)."""
        if self.world == 1:
            ops.cluster_aggregate_(self.theta, client_rows[..., : self.P], n)
            return self.theta
        cp = client_rows
        if cp.shape[2] != self.Pp:  # pad rows to a multiple of 4 floats for 128-bit peer accesses
            cp = torch.nn.functional.pad(cp, (0, self.Pp - cp.shape[2]))
        self.epoch += 1
        grid = _ext.load(required=True).fedavg_reduce_apply_peer(
            cp.contiguous(), n.float().contiguous(), self.Pp, self.Pp, self.world, self.rank, self.part_ptrs, self.theta_ptrs,
            self.tot_ptrs, self.flag_ptrs, self.grid_sync, self.epoch, self.grid_base, 5000, self.error_flag,
            self.mc_part, self.mc_theta)
        self.grid_base += 2 * int(grid)
        return self.theta

    def check(self) -> None:
        if self.world > 1 and int(self.error_flag.item()) != 0:
            raise RuntimeError(f"peer aggregation timed out (code {int(self.error_flag.item())})")
