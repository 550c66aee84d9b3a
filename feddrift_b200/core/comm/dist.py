"""``torch.distributed`` point-to-point transport (gloo on CPU, NCCL on GPU).

Replaces the reference's mpi4py pickle p2p (``mpi_send_thread.py:20-29``,
``mpi_receive_thread.py:20-28``).  Differences by design:

* no polling sleeps — the send thread blocks on a queue, the receive thread
  blocks in ``dist.recv``;
* tensors inside the payload are shipped as ONE flattened buffer per dtype
  (header = pickled skeleton with tensor placeholders), so a state_dict costs
  one bulk transfer instead of a pickle of many small storages;
* ``finish`` is a cooperative stop (a STOP frame to self), never ``MPI_Abort``.
"""
from __future__ import annotations

import logging
import pickle
import queue
import threading
from typing import Any, Dict, List, Tuple

import torch
import torch.distributed as dist

from ..message import Message
from .base import BaseCommunicationManager

_TAG_HDR, _TAG_META, _TAG_DATA = 11, 12, 13


class _TensorSlot:
    __slots__ = ("idx",)

    def __init__(self, idx: int):
        self.idx = idx


def _extract(obj: Any, bag: List[torch.Tensor]) -> Any:
    if isinstance(obj, torch.Tensor):
        bag.append(obj.detach())
        return _TensorSlot(len(bag) - 1)
    if isinstance(obj, dict):
        return {k: _extract(v, bag) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_extract(v, bag) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_extract(v, bag) for v in obj)
    return obj


def _restore(obj: Any, bag: List[torch.Tensor]) -> Any:
    if isinstance(obj, _TensorSlot):
        return bag[obj.idx]
    if isinstance(obj, dict):
        return {k: _restore(v, bag) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_restore(v, bag) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_restore(v, bag) for v in obj)
    return obj


def pack_payload(params: Dict[str, Any]) -> Tuple[bytes, torch.Tensor]:
    """-> (pickled skeleton+tensor metadata, one flat uint8 buffer of all tensor bytes)."""
    bag: List[torch.Tensor] = []
    skeleton = _extract(params, bag)
    metas, chunks = [], []
    for t in bag:
        tc = t.contiguous().cpu()
        metas.append((tc.dtype, tuple(tc.shape)))
        chunks.append(tc.reshape(-1).view(torch.uint8))
    flat = torch.cat(chunks) if chunks else torch.empty(0, dtype=torch.uint8)
    return pickle.dumps((skeleton, metas)), flat


def unpack_payload(header: bytes, flat: torch.Tensor) -> Dict[str, Any]:
    skeleton, metas = pickle.loads(header)
    bag, off = [], 0
    for dtype, shape in metas:
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        bag.append(flat[off:off + nbytes].clone().view(dtype).reshape(shape))
        off += nbytes
    return _restore(skeleton, bag)


class _Endpoint:
    """ONE send thread + ONE receive thread per process and process group.  Managers come and go (a new
    Server/ClientManager pair per time step) but the blocking any-source ``dist.recv`` must have a single owner,
    otherwise a stale receiver of a finished manager would swallow the next time step's messages."""

    _instances: Dict[int, "_Endpoint"] = {}

    def __init__(self, rank: int, size: int, group, device):
        # The control plane relies on any-source ``dist.recv`` and message tags — ProcessGroupNCCL supports neither (the
        # receive thread would raise and the manager would block forever).  On an NCCL default group the endpoint opens a
        # gloo SIDE group for control messages; tensors travel as ``DeviceRef`` handles / peer memory, not through here.
        backend = dist.get_backend(group) if group is not None else dist.get_backend()
        if "gloo" not in str(backend).lower():
            if group is not None:
                raise RuntimeError(f"DistCommunicationManager needs a gloo process group for its control messages, got '{backend}'")
            group = dist.new_group(backend="gloo")
            device = torch.device("cpu")
        self.rank, self.size, self.group, self.device = rank, size, group, device
        self.q_send: "queue.Queue" = queue.Queue()
        self.q_recv: "queue.Queue" = queue.Queue()
        self.threads = [threading.Thread(target=self._send_loop, daemon=True, name=f"fdb-send{rank}"),
                        threading.Thread(target=self._recv_loop, daemon=True, name=f"fdb-recv{rank}")]
        for t in self.threads:
            t.start()

    @classmethod
    def get(cls, rank, size, group, device) -> "_Endpoint":
        key = id(group) if group is not None else 0
        ep = cls._instances.get(key)
        if ep is None:
            ep = cls._instances[key] = _Endpoint(rank, size, group, device)
        return ep

    # wire: [hdr_len, data_len] int64 -> header bytes -> data bytes
    def _send_one(self, dst: int, header: bytes, flat: torch.Tensor) -> None:
        hdr = torch.frombuffer(bytearray(header), dtype=torch.uint8)
        lens = torch.tensor([hdr.numel(), flat.numel()], dtype=torch.int64)
        dist.send(lens.to(self.device), dst, group=self.group, tag=_TAG_HDR)
        dist.send(hdr.to(self.device), dst, group=self.group, tag=_TAG_META)
        if flat.numel():
            dist.send(flat.to(self.device), dst, group=self.group, tag=_TAG_DATA)

    def _send_loop(self) -> None:
        while True:
            item = self.q_send.get()
            try:
                if item is None:
                    return
                self._send_one(*item)
            except Exception as exc:  # group torn down (or a transport error: say so instead of dying silently)
                if dist.is_initialized():
                    logging.error("feddrift_b200 send thread of rank %d stopped: %r", self.rank, exc)
                return
            finally:
                self.q_send.task_done()

    def _recv_loop(self) -> None:
        while True:
            lens = torch.zeros(2, dtype=torch.int64, device=self.device)
            try:
                src = dist.recv(lens, group=self.group, tag=_TAG_HDR)
                hlen, dlen = int(lens[0]), int(lens[1])
                if hlen == 0:  # shutdown frame
                    return
                hdr = torch.empty(hlen, dtype=torch.uint8, device=self.device)
                dist.recv(hdr, src=src, group=self.group, tag=_TAG_META)
                flat = torch.empty(dlen, dtype=torch.uint8, device=self.device)
                if dlen:
                    dist.recv(flat, src=src, group=self.group, tag=_TAG_DATA)
            except Exception as exc:  # process group torn down while blocked (or a transport error)
                if dist.is_initialized():
                    logging.error("feddrift_b200 receive thread of rank %d stopped: %r", self.rank, exc)
                    self.q_recv.put(None)   # unblock the dispatch loop instead of hanging forever
                return
            self.q_recv.put(Message().init(unpack_payload(hdr.cpu().numpy().tobytes(), flat.cpu())))

    def shutdown(self) -> None:
        """Orderly teardown: everybody drains, then each rank wakes its right neighbour's receiver with an empty
        frame, so no thread is left blocked inside gloo at interpreter exit."""
        self.q_send.join()
        dist.barrier(group=self.group)
        if self.size > 1:
            nxt = (self.rank + 1) % self.size
            dist.send(torch.zeros(2, dtype=torch.int64, device=self.device), nxt, group=self.group, tag=_TAG_HDR)
        self.q_send.put(None)
        for t in self.threads:
            t.join(10)


class DistCommunicationManager(BaseCommunicationManager):
    def __init__(self, rank: int, size: int, group=None, device: str = "cpu"):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (FedML_init does it)")
        self.rank, self.size, self.group = rank, size, group
        self.device = torch.device(device)
        self.ep = _Endpoint.get(rank, size, group, self.device)
        self.is_running = True

    def send_message(self, msg: Message) -> None:
        header, flat = pack_payload(msg.to_string())
        self.ep.q_send.put((int(msg.get_receiver_id()), header, flat))

    def post_local(self, msg: Message) -> None:
        """Deliver ``msg`` to this rank's own dispatch loop (thread-safe; used by the round watchdog timer)."""
        self.ep.q_recv.put(msg)

    def handle_receive_message(self) -> None:
        while self.is_running:
            msg = self.ep.q_recv.get()
            if msg is None:
                break
            self.notify(msg)

    def stop_receive_message(self) -> None:
        """Cooperative stop: drain our outgoing queue, then unblock OUR dispatch loop (the endpoint lives on)."""
        self.flush()
        self.is_running = False
        self.ep.q_recv.put(None)

    def flush(self) -> None:
        """Block until every queued outgoing message has been handed to the transport."""
        self.ep.q_send.join()


def shutdown_transport(destroy_group: bool = True) -> None:
    """Tear down every endpoint of this process (call once at the very end of a distributed run)."""
    for ep in list(_Endpoint._instances.values()):
        ep.shutdown()
    _Endpoint._instances.clear()
    if destroy_group and dist.is_initialized():
        dist.destroy_process_group()
