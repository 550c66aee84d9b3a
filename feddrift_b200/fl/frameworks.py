"""The two FedML algorithm templates, used as comm-layer smoke tests.

* ``base_framework`` (``fedml_api/distributed/base_framework/*``): clients send a scalar, the central worker sums
  and broadcasts, for ``comm_round`` rounds.
* ``decentralized_framework`` (``.../decentralized_framework/*``): every rank sends to the out-neighbours of a
  ``SymmetricTopologyManager(n, 2)`` ring and waits for all in-neighbours, for ``comm_round`` rounds.
"""
from __future__ import annotations

from typing import Dict, List

from ..core.managers import ClientManager, ServerManager
from ..core.message import Message
from ..core.topology import SymmetricTopologyManager

MSG_S2C_INIT, MSG_S2C_INFO, MSG_C2S_INFO, MSG_P2P = "s2c_init", "s2c_info", "c2s_info", "p2p_result"


class BaseCentralWorker:
    def __init__(self, client_num, args):
        self.client_num, self.args = client_num, args
        self.client_local_result_list: Dict[int, float] = {}
        self.flag = {i: False for i in range(client_num)}

    def add_client_local_result(self, index, result):
        self.client_local_result_list[index] = result
        self.flag[index] = True

    def check_whether_all_receive(self):
        if not all(self.flag.values()):
            return False
        self.flag = {i: False for i in range(self.client_num)}
        return True

    def aggregate(self):
        return sum(self.client_local_result_list.values())


class BaseClientWorker:
    def __init__(self, client_index):
        self.client_index, self.updated_information = client_index, 0

    def update(self, updated_information):
        self.updated_information = updated_information

    def train(self):
        return self.client_index  # the template's "local computation"


class BaseCentralManager(ServerManager):
    def __init__(self, args, comm, rank, size, aggregator, backend="INPROC"):
        super().__init__(args, comm, rank, size, backend)
        self.aggregator, self.round_num, self.round_idx, self.history = aggregator, args.comm_round, 0, []

    def send_init_msg(self):
        for pid in range(1, self.size):
            self.send_message(Message(MSG_S2C_INIT, self.get_sender_id(), pid))

    def register_message_receive_handlers(self):
        self.register_message_receive_handler(MSG_C2S_INFO, self.handle_message_receive_model_from_client)

    def handle_message_receive_model_from_client(self, msg):
        self.aggregator.add_client_local_result(msg.get_sender_id() - 1, msg.get("client_local_result"))
        if not self.aggregator.check_whether_all_receive():
            return
        g = self.aggregator.aggregate()
        self.history.append(g)
        self.round_idx += 1
        if self.round_idx == self.round_num:
            return self.finish()
        for rid in range(1, self.size):
            m = Message(MSG_S2C_INFO, self.get_sender_id(), rid)
            m.add_params("global_result", g)
            self.send_message(m)


class BaseClientManager(ClientManager):
    def __init__(self, args, comm, rank, size, trainer, backend="INPROC"):
        super().__init__(args, comm, rank, size, backend)
        self.trainer, self.num_rounds, self.round_idx = trainer, args.comm_round, 0

    def register_message_receive_handlers(self):
        self.register_message_receive_handler(MSG_S2C_INIT, self.handle_message_init)
        self.register_message_receive_handler(MSG_S2C_INFO, self.handle_message_receive_model_from_server)

    def handle_message_init(self, msg):
        self.trainer.update(0)
        self._train()

    def handle_message_receive_model_from_server(self, msg):
        self.trainer.update(msg.get("global_result"))
        self.round_idx += 1
        self._train()
        if self.round_idx == self.num_rounds - 1:
            self.finish()

    def _train(self):
        m = Message(MSG_C2S_INFO, self.get_sender_id(), 0)
        m.add_params("client_local_result", self.trainer.train())
        self.send_message(m)


def FedML_Base_distributed(process_id, worker_number, comm, args):
    """INPROC: builds the star and runs it; returns the central manager (its ``history`` = per-round sums)."""
    world = comm.world if hasattr(comm, "world") else comm
    srv = BaseCentralManager(args, world, 0, worker_number, BaseCentralWorker(worker_number - 1, args))
    cls = [BaseClientManager(args, world, r, worker_number, BaseClientWorker(r - 1)) for r in range(1, worker_number)]
    for m in [srv] + cls:
        m.register_message_receive_handlers()
    srv.send_init_msg()
    world.run()
    return srv


class DecentralizedWorker:
    def __init__(self, worker_index, topology_manager):
        self.worker_index = worker_index
        self.in_neighbor_idx_list = topology_manager.get_in_neighbor_idx_list(worker_index)
        self.worker_result_dict: Dict[int, float] = {}
        self.flag = {n: False for n in self.in_neighbor_idx_list}

    def add_result(self, worker_index, updated_information):
        self.worker_result_dict[worker_index] = updated_information
        self.flag[worker_index] = True

    def check_whether_all_receive(self):
        if not all(self.flag.values()):
            return False
        self.flag = {n: False for n in self.in_neighbor_idx_list}
        return True

    def train(self):
        self.worker_result_dict.clear()
        return 0


class DecentralizedWorkerManager(ClientManager):
    def __init__(self, args, comm, rank, size, trainer, topology_manager, backend="INPROC"):
        super().__init__(args, comm, rank, size, backend)
        self.worker, self.topology_manager = trainer, topology_manager
        self.num_rounds, self.round_idx, self.completed = args.comm_round, 0, []

    def start_training(self):
        self._send_to_neighbors(self.worker.train())

    def register_message_receive_handlers(self):
        self.register_message_receive_handler(MSG_P2P, self.handle_msg_from_neighbor)

    def handle_msg_from_neighbor(self, msg):
        self.worker.add_result(msg.get_sender_id(), msg.get("result"))
        if not self.worker.check_whether_all_receive():
            return
        self.completed.append(self.round_idx)
        self.round_idx += 1
        if self.round_idx == self.num_rounds:
            return self.finish()
        self._send_to_neighbors(self.worker.train())

    def _send_to_neighbors(self, value):
        for n in self.topology_manager.get_out_neighbor_idx_list(self.rank):
            m = Message(MSG_P2P, self.get_sender_id(), n)
            m.add_params("result", value)
            self.send_message(m)


def FedML_Decentralized_Demo_distributed(process_id, worker_number, comm, args) -> List[DecentralizedWorkerManager]:
    world = comm.world if hasattr(comm, "world") else comm
    tpmgr = SymmetricTopologyManager(worker_number, 2)
    tpmgr.generate_topology()
    mgrs = [DecentralizedWorkerManager(args, world, r, worker_number, DecentralizedWorker(r, tpmgr), tpmgr)
            for r in range(worker_number)]
    for m in mgrs:
        m.register_message_receive_handlers()
    for m in mgrs:
        m.start_training()
    world.run()
    return mgrs
