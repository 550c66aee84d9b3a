"""ServerManager / ClientManager — the message-handler programming model.

Parity: ``fedml_core/distributed/server/server_manager.py:11-57`` and
``fedml_core/distributed/client/client_manager.py:12-64``: a manager owns one
communication backend, registers itself as observer, maps ``msg_type`` to a
callback, ``run()`` = register handlers + blocking dispatch loop.

Backends (``backend=``):

* ``"INPROC"`` / ``"STREAM"`` — all ranks in this process (``comm`` is a
  :class:`~feddrift_b200.core.comm.inproc.World`).  STREAM is the B200 default:
  ``model_params`` payloads are :class:`DeviceRef` handles into the parameter
  arena, and delivery order is the CUDA stream order.
* ``"MPI"`` / ``"GLOO"`` / ``"NCCL"`` / ``"DIST"`` — one process per rank over
  ``torch.distributed`` p2p (the name MPI is accepted for drop-in compatibility;
  mpi4py itself is never used).
* ``"MQTT"`` — JSON pub/sub (in-process broker unless paho + host are given).

``finish()`` is a cooperative stop instead of the reference's
``MPI.COMM_WORLD.Abort()`` (``server_manager.py:54-57``), so several time
steps can run inside one process (the reference relaunches ``mpirun`` per
time step for exactly that reason — ``run_fedavg_distributed_pytorch.sh:49-84``).
"""
from __future__ import annotations

import logging
from abc import abstractmethod
from typing import Optional, Callable, Dict

from .comm.base import BaseCommunicationManager
from .comm.inproc import InProcCommunicationManager, World
from .comm.mqtt import MqttCommManager
from .message import Message
from .observer import Observer

_DIST_NAMES = {"MPI", "GLOO", "NCCL", "DIST"}


def _make_backend(comm, rank: int, size: int, backend: str, node_type: str) -> BaseCommunicationManager:
    backend = backend.upper()
    if backend in ("INPROC", "STREAM"):
        if not isinstance(comm, World):
            raise TypeError("INPROC/STREAM backends need comm=World(size)")
        return InProcCommunicationManager(comm, rank)
    if backend == "MQTT":
        broker = comm if comm is not None and hasattr(comm, "publish") else None
        return MqttCommManager(client_id=rank, client_num=size - 1, broker=broker)
    if backend in _DIST_NAMES:
        from .comm.dist import DistCommunicationManager
        group = getattr(comm, "group", None)
        device = getattr(comm, "device", "cpu")
        return DistCommunicationManager(rank, size, group=group, device=device)
    raise ValueError(f"unknown backend {backend!r}")


class _Manager(Observer):
    node_type = "client"

    def __init__(self, args, comm=None, rank: int = 0, size: int = 0, backend: str = "MPI"):
        self.args = args
        self.size = size
        self.rank = rank
        self.backend = backend
        self.com_manager = _make_backend(comm, rank, size, backend, self.node_type)
        self.com_manager.add_observer(self)
        self.message_handler_dict: Dict[object, Callable] = {}
        self.finished = False

    def run(self) -> None:
        self.register_message_receive_handlers()
        self.com_manager.handle_receive_message()

    def get_sender_id(self) -> int:
        return self.rank

    def receive_message(self, msg_type, msg_params) -> None:
        try:
            handler = self.message_handler_dict[msg_type]
        except KeyError:
            # JSON transports stringify nothing here, but be lenient with int/str msg types
            handler = self.message_handler_dict[type(next(iter(self.message_handler_dict)))(msg_type)]
        handler(msg_params)

    def send_message(self, message: Message) -> None:
        self.com_manager.send_message(message)

    @abstractmethod
    def register_message_receive_handlers(self) -> None:
        ...

    def register_message_receive_handler(self, msg_type, handler_callback_func) -> None:
        self.message_handler_dict[msg_type] = handler_callback_func

    def post_local(self, message: Message) -> bool:
        """Queue a message for this manager's own handlers (timers, watchdogs); False if the transport cannot."""
        fn = getattr(self.com_manager, "post_local", None)
        if fn is None:
            return False
        fn(message)
        return True

    def finish(self) -> None:
        logging.info("__finish %s rank %d", self.node_type, self.rank)
        self.finished = True
        self.com_manager.stop_receive_message()


class RoundWatchdog:
    """Straggler / failure tolerance for synchronous FL rounds (the reference has none: a missing client blocks the round
    forever, ``check_whether_all_receive``).  The server arms it when it broadcasts round r; if the round is still open
    after ``timeout_s`` (threaded transports: a ``threading.Timer`` posts a local ``MSG_TYPE_ROUND_TIMEOUT``; the
    deterministic INPROC event loop: as soon as the fabric is quiescent, i.e. nothing more can arrive) and at least
    ``min_workers`` uploads are in, the round is closed with the uploads that arrived — missing workers contribute
    weight 0 and receive the next broadcast like everybody else; uploads tagged with an older round are dropped.

    Enabled by ``args.round_timeout_s > 0`` (``args.min_workers_per_round`` defaults to 1)."""

    MSG_TYPE_ROUND_TIMEOUT = 9

    def __init__(self, manager: "_Manager", timeout_s: float, min_workers: int = 1):
        self.manager, self.timeout_s, self.min_workers = manager, float(timeout_s or 0.0), max(int(min_workers or 1), 1)
        self.round_open: Optional[int] = None
        self._timer = None
        self.timeouts = 0

    @property
    def enabled(self) -> bool:
        return self.timeout_s > 0

    def arm(self, round_idx: int) -> None:
        self.cancel()
        self.round_open = round_idx
        if not self.enabled or self.manager.backend in ("INPROC", "STREAM"):
            return   # INPROC: the quiescence hook plays the timer's role
        import threading
        self._timer = threading.Timer(self.timeout_s, self._fire, args=(round_idx,))
        self._timer.daemon = True
        self._timer.start()

    def _fire(self, round_idx: int) -> None:
        msg = Message(self.MSG_TYPE_ROUND_TIMEOUT, self.manager.rank, self.manager.rank)
        msg.add_params("round_idx", round_idx)
        self.manager.post_local(msg)

    def fire_if_open(self) -> bool:
        """INPROC quiescence hook: post the timeout for the open round (once)."""
        if not self.enabled or self.round_open is None:
            return False
        r, self.round_open = self.round_open, None
        self._fire(r)
        return True

    def cancel(self) -> None:
        if self._timer is not None:
            self._timer.cancel()
            self._timer = None
        self.round_open = None


class ServerManager(_Manager):
    node_type = "server"


class ClientManager(_Manager):
    node_type = "client"

    def send_message(self, message: Message) -> None:
        # the reference re-wraps into a fresh Message (client_manager.py:44-52); keep the
        # observable effect (header keys first, then payload) without the copy semantics mattering
        msg = Message(message.get_type(), message.get_sender_id(), message.get_receiver_id())
        for key, value in message.get_params().items():
            msg.add(key, value)
        self.com_manager.send_message(msg)
