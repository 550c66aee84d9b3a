"""Multi-GPU per-cluster FedAvg aggregation + model broadcast in ONE kernel over NVLink peer memory
(``csrc/aggregate_peer.cu``): clients are sharded over the ranks, every rank ends with identical cluster models.

    agg = PeerAggregator(num_models, P, device)          # collective: allocates + rendezvous symmetric buffers
    theta = agg.aggregate(client_rows [C_local, M, P], n [C_local, M])   # -> this rank's [M, P] view of the models

No NCCL call on this path: the rendezvous (cold, once) uses ``torch.distributed._symmetric_memory``; the hot path is a
cooperative kernel whose producer CTAs stream the local client rows (read IN PLACE from the client arena through a row
list — no gather / pad copy) while its consumer CTAs reduce-scatter + normalise + all-gather every ~1 MB chunk over
NVLink as soon as all ranks have flagged it (``st.release.sys`` / ``ld.acquire.sys`` per-chunk epoch words; NVLS
``multimem`` when available).  ``theta_full`` lives in the symmetric buffer, so a ``ModelBank`` can be re-bound onto it
(``ModelBank.rebind_storage``) and the aggregated models need no copy either.  With one process (``world == 1``) it
degenerates to the single-GPU K1 kernel.
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
            self.max_chunks = 2048                                   # per-chunk epoch words: slot 0 totals, 1 final, 2+k chunk k
            flag_f = ((2 + self.max_chunks) * self.world + 31) // 32 * 32
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
            self.chunk_done = torch.zeros(self.max_chunks, dtype=torch.int32, device=self.device)
            self.launches = 0
            self.error_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.epoch, self.grid_base = 0, 0
            torch.cuda.synchronize()
            dist.barrier()
        self.theta = self.theta_full[:, :P]
        if theta_init is not None:
            self.theta.copy_(theta_init)

    def aggregate(self, client_rows: torch.Tensor, n: torch.Tensor, cidx: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``client_rows`` [C, M, P(p)] fp32 and ``n`` [C_local, M] weights (0 = did not train).  Without ``cidx`` the first
        ``C_local`` rows of ``client_rows`` are this rank's clients; with ``cidx`` (int32 ``[C_local]``) ``client_rows`` is the
        WHOLE client arena and the kernel reads rows ``cidx[c]`` in place."""
        if self.world == 1:
            rows = client_rows if cidx is None else client_rows.index_select(0, cidx.long())
            ops.cluster_aggregate_(self.theta, rows[..., : self.P], n)
            return self.theta
        cp = client_rows
        if cp.shape[2] != self.Pp:  # rows not padded to a multiple of 4 floats: one padded copy (128-bit peer accesses need it)
            if cidx is not None:
                cp, cidx = cp.index_select(0, cidx.long()), None
            cp = torch.nn.functional.pad(cp, (0, self.Pp - cp.shape[2]))
        self.epoch += 1
        grid = _ext.load(required=True).fedavg_reduce_apply_peer(
            cp.contiguous(), cidx, n.float().contiguous(), self.Pp, self.Pp, self.world, self.rank, self.part_ptrs, self.theta_ptrs,
            self.tot_ptrs, self.flag_ptrs, self.grid_sync, self.chunk_done, self.launches, self.epoch, self.grid_base,
            int(__import__("os").environ.get("FDB_SPIN_TIMEOUT_MS", 60000)), self.error_flag, self.mc_part, self.mc_theta)
        self.grid_base += int(grid)
        self.launches += 1
        return self.theta

    def check(self) -> None:
        if self.world > 1 and int(self.error_flag.item()) != 0:
            raise RuntimeError(f"peer aggregation timed out (code {int(self.error_flag.item())})")
