"""In-process transport: every logical rank lives in one Python process.

The reference has no fake/mock transport; its only path is
mpi4py p2p with two helper threads and 0.3 s polling sleeps
(``fedml_core/distributed/communication/mpi/com_manager.py:71-79``).  This
backend replaces that with mailboxes that block on a condition variable (no
polling) and supports two execution modes:

* ``World.run(managers)`` — deterministic single-threaded event loop (used by
  tests and by the STREAM scheduler, where a "send" of model parameters is a
  stream-ordered device op rather than a pickle);
* threaded — each manager's blocking ``run()`` executes in its own thread.
"""
from __future__ import annotations

import queue
import threading
from collections import deque
from typing import Deque, Dict, List, Optional

from ..message import Message
from .base import BaseCommunicationManager

_STOP = object()


class World:
    """Shared mailbox fabric for ``size`` logical ranks."""

    def __init__(self, size: int):
        self.size = size
        self.mailboxes: List["queue.Queue"] = [queue.Queue() for _ in range(size)]
        self.managers: Dict[int, "InProcCommunicationManager"] = {}
        self.order: Deque[int] = deque()  # global FIFO of receiver ranks (deterministic replay)
        self.lock = threading.Lock()
        self.delivered = 0
        self.barrier_obj = threading.Barrier(size) if size > 0 else None

    def post(self, msg: Message) -> None:
        dst = int(msg.get_receiver_id())
        if not 0 <= dst < self.size:
            raise ValueError(f"receiver {dst} outside world of size {self.size}")
        with self.lock:
            self.order.append(dst)
        self.mailboxes[dst].put(msg)

    # -- deterministic single-threaded event loop ------------------------------
    def run(self, managers=None, max_messages: Optional[int] = None) -> int:
        """Dispatch messages in global FIFO order until every rank stopped or
        the fabric is quiescent.  Returns the number of messages delivered."""
        if managers is not None:
            for m in managers:
                m.register_message_receive_handlers()
        n = 0
        while True:
            with self.lock:
                idle = not self.order
            if idle:
                # quiescent fabric = "virtual timeout": nothing more can arrive, so let the observers that registered an
                # `on_quiescent` hook (the server's round watchdog) decide whether to move on without the missing peers
                fired = False
                for cm in list(self.managers.values()):
                    for ob in list(getattr(cm, "_observers", [])):
                        hook = getattr(ob, "on_quiescent", None)
                        if cm.is_running and hook is not None and hook():
                            fired = True
                if not fired:
                    break
                continue
            with self.lock:
                if not self.order:
                    continue
                dst = self.order.popleft()
            msg = _STOP
            try:
                while msg is _STOP:  # stale stop sentinels of a finished manager must not shift the FIFO
                    msg = self.mailboxes[dst].get_nowait()
            except queue.Empty:
                continue
            cm = self.managers.get(dst)
            if cm is None or not cm.is_running:
                continue
            cm.notify(msg)
            n += 1
            self.delivered += 1
            if max_messages is not None and n >= max_messages:
                break
        return n

    def run_threads(self, managers, timeout: Optional[float] = None) -> None:
        threads = [threading.Thread(target=m.run, name=f"rank{m.rank}", daemon=True) for m in managers]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout)


class InProcCommunicationManager(BaseCommunicationManager):
    def __init__(self, world: World, rank: int):
        super().__init__()
        self.world = world
        self.rank = rank
        self.is_running = True
        world.managers[rank] = self
        box, keep = world.mailboxes[rank], []
        while True:  # a new manager on this rank starts with a mailbox free of old stop sentinels
            try:
                item = box.get_nowait()
            except queue.Empty:
                break
            if item is not _STOP:
                keep.append(item)
        for item in keep:
            box.put(item)

    def send_message(self, msg: Message) -> None:
        self.world.post(msg)

    def post_local(self, msg: Message) -> None:
        """Deliver ``msg`` to this rank's own dispatch loop (used by the round watchdog)."""
        self.world.post(msg)

    def handle_receive_message(self) -> None:
        """Blocking receive loop (threaded mode).  No sleeps: the mailbox blocks."""
        box = self.world.mailboxes[self.rank]
        while self.is_running:
            msg = box.get()
            if msg is _STOP:
                break
            with self.world.lock:
                try:
                    self.world.order.remove(self.rank)
                except ValueError:
                    pass
            self.notify(msg)
            self.world.delivered += 1

    def stop_receive_message(self) -> None:
        self.is_running = False
        self.world.mailboxes[self.rank].put(_STOP)
