"""Real MQTT 3.1.1 framing: golden packets from the OASIS spec, client <-> embedded broker over TCP, comm-manager round trip."""
import threading
import time

from feddrift_b200.core.comm import mqtt_wire as mw
from feddrift_b200.core.comm.mqtt import MqttCommManager
from feddrift_b200.core.message import Message


def test_golden_packets():
    # remaining-length examples of MQTT 3.1.1 §2.2.3
    assert mw.encode_remaining_length(0) == b"\x00"
    assert mw.encode_remaining_length(127) == b"\x7f"
    assert mw.encode_remaining_length(128) == b"\x80\x01"
    assert mw.encode_remaining_length(16383) == b"\xff\x7f"
    assert mw.encode_remaining_length(16384) == b"\x80\x80\x01"
    assert mw.encode_remaining_length(268435455) == b"\xff\xff\xff\x7f"
    # CONNECT with protocol name "MQTT", level 4, clean session, keepalive 60, client id "c1"
    assert mw.connect_packet("c1", 60) == bytes([0x10, 14, 0, 4]) + b"MQTT" + bytes([4, 2, 0, 60, 0, 2]) + b"c1"
    # PUBLISH QoS 0 topic "a/b" payload "hi"
    assert mw.publish_packet("a/b", b"hi") == bytes([0x30, 7, 0, 3]) + b"a/b" + b"hi"
    # SUBSCRIBE packet id 10, topic "fedml0_1" QoS 0 (fixed-header flags 0b0010)
    assert mw.subscribe_packet(10, [("fedml0_1", 0)]) == bytes([0x82, 13, 0, 10, 0, 8]) + b"fedml0_1" + b"\x00"
    assert mw.packet(mw.PINGREQ, 0, b"") == b"\xc0\x00" and mw.packet(mw.DISCONNECT, 0, b"") == b"\xe0\x00"
    assert mw.topic_matches("a/+/c", "a/b/c") and mw.topic_matches("a/#", "a/b/c") and not mw.topic_matches("a/+", "a/b/c")


def test_client_broker_pubsub_over_tcp():
    broker = mw.MqttBroker().start()
    got, ev = [], threading.Event()
    sub = mw.MqttClient("sub", on_message=lambda t, p: (got.append((t, p)), ev.set()))
    pub = mw.MqttClient("pub")
    try:
        sub.connect(broker.host, broker.port)
        pub.connect(broker.host, broker.port)
        sub.subscribe("fedml0_3")
        big = ("x" * 70000).encode()          # > 16383 bytes: 3-byte remaining length
        pub.publish("fedml0_3", big)
        assert ev.wait(5.0)
        assert got[0] == ("fedml0_3", big)
        ev.clear()
        pub.publish("other", b"ignored")
        pub.publish("fedml0_3", "second")
        assert ev.wait(5.0) and got[1] == ("fedml0_3", b"second")
    finally:
        sub.disconnect(); pub.disconnect(); broker.stop()


def test_comm_manager_round_trip_over_real_mqtt():
    """Server (id 0) <-> client (id 1) with the reference's topic scheme over real MQTT frames."""
    broker = mw.MqttBroker().start()
    try:
        server = MqttCommManager(host=broker.host, port=broker.port, topic="fedml", client_id=0, client_num=1)
        client = MqttCommManager(host=broker.host, port=broker.port, topic="fedml", client_id=1, client_num=1)
        seen = {}

        class Obs:
            def __init__(self, name):
                self.name = name

            def receive_message(self, msg_type, msg):
                seen[self.name] = (msg_type, msg.get("model_params"))

        server.add_observer(Obs("server")); client.add_observer(Obs("client"))
        m = Message(2, 0, 1)
        m.add_params("model_params", {"w": [[1.0, 2.0], [3.0, 4.0]]})
        server.send_message(m)
        up = Message(3, 1, 0)
        up.add_params("model_params", {"w": [[0.5]]})
        client.send_message(up)
        t0 = time.time()
        while (server.poll() + client.poll() >= 0) and len(seen) < 2 and time.time() - t0 < 5:
            time.sleep(0.01)
        assert seen["client"] == (2, {"w": [[1.0, 2.0], [3.0, 4.0]]})
        assert seen["server"] == (3, {"w": [[0.5]]})
        server.stop_receive_message(); client.stop_receive_message()
    finally:
        broker.stop()
