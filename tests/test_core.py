"""L1 core runtime: Message / Observer / managers / transports / topology / robustness (CPU)."""
import numpy as np
import torch

from feddrift_b200.core import (AsymmetricTopologyManager, ClientManager, LocalBroker, Message, MqttCommManager,
                                RobustAggregator, ServerManager, SymmetricTopologyManager, World, vectorize_weight)
from feddrift_b200.core.comm.dist import pack_payload, unpack_payload


def test_message_roundtrip_json_and_payload_packing():
    m = Message(3, 1, 0)
    m.add_params("model_params", {"w": torch.arange(6.).reshape(2, 3)})
    m.add_params("n", np.int64(7))
    back = Message().init_from_json_string(m.to_json())
    assert back.get_type() == 3 and back.get_sender_id() == 1 and back.get("model_params")["w"] == [[0, 1, 2], [3, 4, 5]]
    hdr, flat = pack_payload({"a": [torch.ones(3), {"b": torch.arange(4, dtype=torch.int32)}], "k": "v"})
    out = unpack_payload(hdr, flat)
    assert torch.equal(out["a"][0], torch.ones(3)) and out["a"][1]["b"].dtype == torch.int32 and out["k"] == "v"


class _SumServer(ServerManager):
    """base_framework template: clients send a scalar, server sums and broadcasts (central_manager.py:25-44)."""

    def __init__(self, args, comm, size, rounds):
        super().__init__(args, comm, 0, size, "INPROC")
        self.rounds, self.round, self.got, self.totals = rounds, 0, {}, []

    def register_message_receive_handlers(self):
        self.register_message_receive_handler("c2s", self.on_value)

    def start(self):
        for r in range(1, self.size):
            self.send_message(Message("s2c", 0, r))

    def on_value(self, msg):
        self.got[msg.get_sender_id()] = msg.get("value")
        if len(self.got) == self.size - 1:
            self.totals.append(sum(self.got.values()))
            self.got = {}
            self.round += 1
            if self.round == self.rounds:
                return self.finish()
            self.start()


class _SumClient(ClientManager):
    def register_message_receive_handlers(self):
        self.register_message_receive_handler("s2c", self.on_go)

    def on_go(self, msg):
        out = Message("c2s", self.rank, 0)
        out.add_params("value", self.rank)
        self.send_message(out)


def test_inproc_star_framework_deterministic_and_threaded():
    for threaded in (False, True):
        world = World(5)
        server = _SumServer(None, world, 5, rounds=3)
        clients = [_SumClient(None, world, r, 5, "INPROC") for r in range(1, 5)]
        if threaded:
            import threading
            ths = [threading.Thread(target=m.run, daemon=True) for m in [server] + clients]
            for t in ths:
                t.start()
            server.start()
            ths[0].join(10)
            for c in clients:
                c.finish()
        else:
            for m in [server] + clients:
                m.register_message_receive_handlers()
            server.start()
            world.run()
        assert server.totals == [10, 10, 10]


def test_mqtt_topics_and_json_wire():
    broker = LocalBroker()
    server = MqttCommManager(client_id=0, client_num=2, broker=broker)
    c1 = MqttCommManager(client_id=1, client_num=2, broker=broker)
    got = []

    class Obs:
        def receive_message(self, t, m):
            got.append((t, m.get("x")))
    c1.add_observer(Obs())
    server.add_observer(Obs())
    msg = Message(2, 0, 1)
    msg.add_params("x", torch.tensor([1.0, 2.0]))
    server.send_message(msg)            # published on fedml0_1
    up = Message(3, 1, 0)
    up.add_params("x", 5)
    c1.send_message(up)                 # published on fedml1
    assert c1.poll() == 1 and server.poll() == 1
    assert got == [(2, [1.0, 2.0]), (3, 5)]


def test_topologies_are_row_stochastic_and_ring_connected():
    s = SymmetricTopologyManager(8, 4)
    s.generate_topology()
    assert np.allclose(s.topology.sum(1), 1) and np.allclose(s.topology, s.topology.T)
    assert s.get_in_neighbor_idx_list(0) == [1, 2, 6, 7]
    a = AsymmetricTopologyManager(8, 4, 2, rng=np.random.RandomState(0))
    a.generate_topology()
    assert np.allclose(a.topology.sum(1), 1)
    for i in range(8):
        assert (i + 1) % 8 in a.get_out_neighbor_idx_list(i)
    assert len(a.get_in_neighbor_weights(3)) == 8


def test_robust_aggregator_clips_only_weight_params():
    class A:
        defense_type, norm_bound, stddev = "norm_diff_clipping", 1.0, 0.1
    ra = RobustAggregator(A())
    g = {"w": torch.zeros(4), "bn.running_mean": torch.zeros(2)}
    l = {"w": torch.full((4,), 3.0), "bn.running_mean": torch.ones(2)}
    out = ra.norm_diff_clipping(l, g)
    assert abs(out["w"].norm().item() - 1.0) < 1e-6 and torch.equal(out["bn.running_mean"], torch.ones(2))
    assert vectorize_weight(l).numel() == 4
    rows = torch.stack([torch.full((6,), 3.0), torch.full((6,), 0.1)])
    ra.clip_flat(rows, torch.zeros(6))
    assert abs(rows[0].norm().item() - 1.0) < 1e-5 and torch.allclose(rows[1], torch.full((6,), 0.1))
    assert ra.add_noise(torch.zeros(1000)).std().item() > 0.05
