"""MQTT-style pub/sub transport with JSON payloads.

Parity: ``fedml_core/distributed/communication/mqtt/mqtt_comm_manager.py:14-126``
— server (id 0) subscribes ``<topic><cid>`` for each client and publishes to
``<topic>0_<cid>``; clients do the reverse; payload = ``Message.to_json()`` so
tensors travel as nested lists (the ``is_mobile`` wire form).

Transports: with ``host``/``port`` the manager speaks REAL MQTT 3.1.1 over TCP
through the dependency-free client of :mod:`.mqtt_wire` (CONNECT / SUBSCRIBE /
PUBLISH QoS-0 / PINGREQ framing, interoperable with mosquitto / EMQX and with
the embedded :class:`~.mqtt_wire.MqttBroker`); if ``paho`` happens to be
importable it is used instead.  Without a host, an in-process
:class:`LocalBroker` with the same topic semantics is used (single-process
simulations and tests).
"""
from __future__ import annotations

import queue
import threading
from collections import defaultdict
from typing import Callable, Dict, List, Optional

from ..message import Message
from .base import BaseCommunicationManager


class LocalBroker:
    """Minimal topic broker: exact-match topics, QoS-0, synchronous fan-out."""

    _default: Optional["LocalBroker"] = None

    def __init__(self) -> None:
        self._subs: Dict[str, List[Callable[[str, str], None]]] = defaultdict(list)
        self._lock = threading.Lock()
        self.published = 0

    @classmethod
    def default(cls) -> "LocalBroker":
        if cls._default is None:
            cls._default = LocalBroker()
        return cls._default

    def subscribe(self, topic: str, callback: Callable[[str, str], None]) -> None:
        with self._lock:
            self._subs[topic].append(callback)

    def unsubscribe(self, topic: str, callback) -> None:
        with self._lock:
            if callback in self._subs.get(topic, []):
                self._subs[topic].remove(callback)

    def publish(self, topic: str, payload: str) -> int:
        with self._lock:
            subs = list(self._subs.get(topic, []))
            self.published += 1
        for cb in subs:
            cb(topic, payload)
        return len(subs)


class MqttCommManager(BaseCommunicationManager):
    def __init__(self, host: Optional[str] = None, port: int = 1883, topic: str = "fedml", client_id: int = 0,
                 client_num: int = 0, broker: Optional[LocalBroker] = None):
        super().__init__()
        self._topic = topic
        self._client_id = int(client_id)
        self.client_num = client_num
        self._inbox: "queue.Queue" = queue.Queue()
        self.is_running = True
        self._paho = None
        self._wire = None
        if broker is None and host is not None:
            try:  # pragma: no cover - paho is not in this image
                import paho.mqtt.client as mqtt
                self._paho = mqtt.Client(client_id=str(self._client_id))
                self._paho.on_message = lambda c, u, m: self._on_message(m.topic, str(m.payload, encoding="utf-8"))
                self._paho.connect(host, port, 60)
                self._paho.loop_start()
            except ImportError:
                self._paho = None
            if self._paho is None:
                # the real protocol without paho: MQTT 3.1.1 over a TCP socket (mqtt_wire.MqttClient)
                from .mqtt_wire import MqttClient
                self._wire = MqttClient(f"{topic}-{self._client_id}",
                                        on_message=lambda t, p: self._on_message(t, p.decode("utf-8")))
                self._wire.connect(host, port)
        self._broker = broker if broker is not None else (None if (self._paho or self._wire) else LocalBroker.default())
        for t in self._rx_topics():
            if self._paho is not None:  # pragma: no cover
                self._paho.subscribe(t, 0)
            elif self._wire is not None:
                self._wire.subscribe(t, 0)
            else:
                self._broker.subscribe(t, self._on_message)

    @property
    def client_id(self) -> int:
        return self._client_id

    @property
    def topic(self) -> str:
        return self._topic

    def _rx_topics(self) -> List[str]:
        if self._client_id == 0:  # server listens to every client's uplink topic
            return [f"{self._topic}{cid}" for cid in range(1, self.client_num + 1)]
        return [f"{self._topic}0_{self._client_id}"]

    def _tx_topic(self, msg: Message) -> str:
        if self._client_id == 0:
            return f"{self._topic}0_{msg.get_receiver_id()}"
        return f"{self._topic}{self._client_id}"

    def _on_message(self, topic: str, payload: str) -> None:
        self._inbox.put(payload)

    def send_message(self, msg: Message) -> None:
        payload = msg.to_json()
        if self._paho is not None:  # pragma: no cover
            self._paho.publish(self._tx_topic(msg), payload=payload)
        elif self._wire is not None:
            self._wire.publish(self._tx_topic(msg), payload)
        else:
            self._broker.publish(self._tx_topic(msg), payload)

    def poll(self, max_messages: Optional[int] = None) -> int:
        """Dispatch pending messages without blocking (single-threaded use)."""
        n = 0
        while max_messages is None or n < max_messages:
            try:
                payload = self._inbox.get_nowait()
            except queue.Empty:
                break
            if payload is None:
                break
            self.notify(Message().init_from_json_string(payload))
            n += 1
        return n

    def handle_receive_message(self) -> None:
        while self.is_running:
            payload = self._inbox.get()
            if payload is None:
                break
            self.notify(Message().init_from_json_string(payload))

    def stop_receive_message(self) -> None:
        self.is_running = False
        self._inbox.put(None)
        if self._paho is not None:  # pragma: no cover
            self._paho.loop_stop()
            self._paho.disconnect()
        if self._wire is not None:
            self._wire.disconnect()
