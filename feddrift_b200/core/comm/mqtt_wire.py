"""MQTT 3.1.1 on the wire — a dependency-free client and a small embedded broker (TCP sockets, QoS 0).

The reference's mobile path talks to a public broker through paho-mqtt (``mqtt_comm_manager.py:21-25,47-70``: connect,
subscribe ``fedml<cid>`` / ``fedml0_<cid>``, publish JSON, ``loop_forever``).  paho is not in this image and there is no
network, so this module implements the protocol itself (OASIS MQTT 3.1.1 §2–§3): fixed header + variable-length
"remaining length", CONNECT / CONNACK, SUBSCRIBE / SUBACK, UNSUBSCRIBE / UNSUBACK, PUBLISH (QoS 0, also accepts QoS 1 and
answers PUBACK), PINGREQ / PINGRESP, DISCONNECT.  :class:`MqttClient` interoperates with any standard broker
(mosquitto, EMQX); :class:`MqttBroker` is a loopback broker for single-box deployments and tests — real paho clients can
connect to it.  ``MqttCommManager(host=..., port=...)`` uses the client; ``MqttBroker.start()`` provides the endpoint.
"""
from __future__ import annotations

import socket
import struct
import threading
from collections import defaultdict
from typing import Callable, Dict, List, Optional, Set, Tuple

CONNECT, CONNACK, PUBLISH, PUBACK, SUBSCRIBE, SUBACK, UNSUBSCRIBE, UNSUBACK, PINGREQ, PINGRESP, DISCONNECT = \
    1, 2, 3, 4, 8, 9, 10, 11, 12, 13, 14


# ------------------------------------------------------------------------------------------------ encoding helpers
def encode_remaining_length(n: int) -> bytes:
    if n < 0 or n > 268_435_455:
        raise ValueError("MQTT remaining length out of range")
    out = bytearray()
    while True:
        d, n = n % 128, n // 128
        out.append(d | (0x80 if n else 0))
        if not n:
            return bytes(out)


def encode_string(s: str) -> bytes:
    b = s.encode("utf-8")
    return struct.pack("!H", len(b)) + b


def packet(ptype: int, flags: int, body: bytes) -> bytes:
    return bytes([(ptype << 4) | (flags & 0x0F)]) + encode_remaining_length(len(body)) + body


def connect_packet(client_id: str, keepalive: int = 60, clean_session: bool = True) -> bytes:
    body = encode_string("MQTT") + bytes([4, 0x02 if clean_session else 0x00]) + struct.pack("!H", keepalive) + encode_string(client_id)
    return packet(CONNECT, 0, body)


def publish_packet(topic: str, payload: bytes, qos: int = 0, packet_id: int = 0, retain: bool = False) -> bytes:
    body = encode_string(topic) + (struct.pack("!H", packet_id) if qos else b"") + payload
    return packet(PUBLISH, (qos << 1) | (1 if retain else 0), body)


def subscribe_packet(packet_id: int, topics: List[Tuple[str, int]]) -> bytes:
    body = struct.pack("!H", packet_id) + b"".join(encode_string(t) + bytes([q]) for t, q in topics)
    return packet(SUBSCRIBE, 0x02, body)


def unsubscribe_packet(packet_id: int, topics: List[str]) -> bytes:
    return packet(UNSUBSCRIBE, 0x02, struct.pack("!H", packet_id) + b"".join(encode_string(t) for t in topics))


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("MQTT peer closed the connection")
        buf += chunk
    return bytes(buf)


def read_packet(sock: socket.socket) -> Tuple[int, int, bytes]:
    """-> (packet type, flags, body)."""
    h = _recv_exact(sock, 1)[0]
    mult, n = 1, 0
    for _ in range(4):
        d = _recv_exact(sock, 1)[0]
        n += (d & 0x7F) * mult
        if not d & 0x80:
            break
        mult *= 128
    else:
        raise ValueError("malformed MQTT remaining length")
    return h >> 4, h & 0x0F, _recv_exact(sock, n) if n else b""


def parse_publish(flags: int, body: bytes) -> Tuple[str, bytes, int, int]:
    """-> (topic, payload, qos, packet id)."""
    tl = struct.unpack("!H", body[:2])[0]
    topic = body[2:2 + tl].decode("utf-8")
    qos, off, pid = (flags >> 1) & 3, 2 + tl, 0
    if qos:
        pid = struct.unpack("!H", body[off:off + 2])[0]
        off += 2
    return topic, body[off:], qos, pid


def topic_matches(filt: str, topic: str) -> bool:
    """MQTT topic filter semantics (``+`` single level, ``#`` multi level)."""
    f, t = filt.split("/"), topic.split("/")
    for i, seg in enumerate(f):
        if seg == "#":
            return True
        if i >= len(t) or (seg != "+" and seg != t[i]):
            return False
    return len(f) == len(t)


# ------------------------------------------------------------------------------------------------ client
class MqttClient:
    """Blocking-socket MQTT 3.1.1 client with a reader thread (the subset paho's ``Client`` exposes to FedML)."""

    def __init__(self, client_id: str, on_message: Optional[Callable[[str, bytes], None]] = None, keepalive: int = 60):
        self.client_id, self.on_message, self.keepalive = client_id, on_message, keepalive
        self._sock: Optional[socket.socket] = None
        self._wlock = threading.Lock()
        self._pid = 0
        self._acks: Dict[int, threading.Event] = {}
        self._reader: Optional[threading.Thread] = None
        self._pinger: Optional[threading.Thread] = None
        self._stop = threading.Event()
        self.connected = False

    def _next_pid(self) -> int:
        self._pid = self._pid % 65535 + 1
        return self._pid

    def connect(self, host: str, port: int = 1883, timeout: float = 10.0) -> None:
        self._sock = socket.create_connection((host, port), timeout=timeout)
        self._sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        self._sock.sendall(connect_packet(self.client_id, self.keepalive))
        ptype, _, body = read_packet(self._sock)
        if ptype != CONNACK or len(body) != 2 or body[1] != 0:
            raise ConnectionError(f"MQTT CONNACK refused (type {ptype}, body {body!r})")
        self._sock.settimeout(None)
        self.connected = True
        self._reader = threading.Thread(target=self._read_loop, daemon=True, name=f"mqtt-{self.client_id}")
        self._reader.start()
        if self.keepalive > 0:
            self._pinger = threading.Thread(target=self._ping_loop, daemon=True, name=f"mqtt-ping-{self.client_id}")
            self._pinger.start()

    def _send(self, data: bytes) -> None:
        with self._wlock:
            self._sock.sendall(data)

    def _ping_loop(self) -> None:
        while not self._stop.wait(self.keepalive / 2.0):
            try:
                self._send(packet(PINGREQ, 0, b""))
            except OSError:
                return

    def _read_loop(self) -> None:
        try:
            while not self._stop.is_set():
                ptype, flags, body = read_packet(self._sock)
                if ptype == PUBLISH:
                    topic, payload, qos, pid = parse_publish(flags, body)
                    if qos == 1:
                        self._send(packet(PUBACK, 0, struct.pack("!H", pid)))
                    if self.on_message is not None:
                        self.on_message(topic, payload)
                elif ptype in (SUBACK, UNSUBACK, PUBACK):
                    ev = self._acks.pop(struct.unpack("!H", body[:2])[0], None)
                    if ev is not None:
                        ev.set()
                # PINGRESP and anything else: nothing to do
        except (ConnectionError, OSError, ValueError):
            self.connected = False

    def subscribe(self, topic: str, qos: int = 0, timeout: float = 10.0) -> None:
        pid, ev = self._next_pid(), threading.Event()
        self._acks[pid] = ev
        self._send(subscribe_packet(pid, [(topic, qos)]))
        if not ev.wait(timeout):
            raise TimeoutError(f"no SUBACK for {topic}")

    def unsubscribe(self, topic: str, timeout: float = 10.0) -> None:
        pid, ev = self._next_pid(), threading.Event()
        self._acks[pid] = ev
        self._send(unsubscribe_packet(pid, [topic]))
        ev.wait(timeout)

    def publish(self, topic: str, payload, qos: int = 0) -> None:
        data = payload.encode("utf-8") if isinstance(payload, str) else bytes(payload)
        self._send(publish_packet(topic, data, qos=0))

    def disconnect(self) -> None:
        self._stop.set()
        try:
            self._send(packet(DISCONNECT, 0, b""))
            self._sock.shutdown(socket.SHUT_RDWR)
        except OSError:
            pass
        finally:
            try:
                self._sock.close()
            except OSError:
                pass
            self.connected = False


# ------------------------------------------------------------------------------------------------ embedded broker
class MqttBroker:
    """Loopback MQTT 3.1.1 broker: one thread per connection, topic filters with wildcards, QoS-0 delivery."""

    def __init__(self, host: str = "127.0.0.1", port: int = 0):
        self._srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._srv.bind((host, port))
        self._srv.listen(64)
        self.host, self.port = self._srv.getsockname()
        self._subs: Dict[socket.socket, Set[str]] = defaultdict(set)
        self._wlocks: Dict[socket.socket, threading.Lock] = {}
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self.published = 0
        self._thread: Optional[threading.Thread] = None

    def start(self) -> "MqttBroker":
        self._thread = threading.Thread(target=self._accept_loop, daemon=True, name="mqtt-broker")
        self._thread.start()
        return self

    def _accept_loop(self) -> None:
        while not self._stop.is_set():
            try:
                conn, _ = self._srv.accept()
            except OSError:
                return
            conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            threading.Thread(target=self._serve, args=(conn,), daemon=True).start()

    def _send(self, conn: socket.socket, data: bytes) -> None:
        lock = self._wlocks.get(conn)
        if lock is None:
            return
        with lock:
            try:
                conn.sendall(data)
            except OSError:
                pass

    def _serve(self, conn: socket.socket) -> None:
        try:
            ptype, _, body = read_packet(conn)
            if ptype != CONNECT or body[2:6] != b"MQTT" or body[6] != 4:
                conn.sendall(packet(CONNACK, 0, bytes([0, 1])))     # unacceptable protocol version
                return
            with self._lock:
                self._wlocks[conn] = threading.Lock()
            self._send(conn, packet(CONNACK, 0, bytes([0, 0])))
            while not self._stop.is_set():
                ptype, flags, body = read_packet(conn)
                if ptype == PUBLISH:
                    topic, payload, qos, pid = parse_publish(flags, body)
                    if qos == 1:
                        self._send(conn, packet(PUBACK, 0, struct.pack("!H", pid)))
                    out = publish_packet(topic, payload, qos=0)
                    with self._lock:
                        self.published += 1
                        targets = [c for c, fs in self._subs.items() if any(topic_matches(f, topic) for f in fs)]
                    for c in targets:
                        self._send(c, out)
                elif ptype == SUBSCRIBE:
                    pid, off, codes = struct.unpack("!H", body[:2])[0], 2, bytearray()
                    while off < len(body):
                        tl = struct.unpack("!H", body[off:off + 2])[0]
                        filt = body[off + 2:off + 2 + tl].decode("utf-8")
                        off += 2 + tl + 1
                        with self._lock:
                            self._subs[conn].add(filt)
                        codes.append(0)                                # granted QoS 0
                    self._send(conn, packet(SUBACK, 0, struct.pack("!H", pid) + bytes(codes)))
                elif ptype == UNSUBSCRIBE:
                    pid, off = struct.unpack("!H", body[:2])[0], 2
                    while off < len(body):
                        tl = struct.unpack("!H", body[off:off + 2])[0]
                        with self._lock:
                            self._subs[conn].discard(body[off + 2:off + 2 + tl].decode("utf-8"))
                        off += 2 + tl
                    self._send(conn, packet(UNSUBACK, 0, struct.pack("!H", pid)))
                elif ptype == PINGREQ:
                    self._send(conn, packet(PINGRESP, 0, b""))
                elif ptype == DISCONNECT:
                    return
        except (ConnectionError, OSError, ValueError, IndexError, struct.error):
            pass
        finally:
            with self._lock:
                self._subs.pop(conn, None)
                self._wlocks.pop(conn, None)
            try:
                conn.close()
            except OSError:
                pass

    def stop(self) -> None:
        self._stop.set()
        try:
            self._srv.close()
        except OSError:
            pass
        with self._lock:
            conns = list(self._wlocks)
        for c in conns:
            try:
                c.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass
