"""Mobile / IoT serving façade: device registration over HTTP + FedAvg over MQTT-style JSON pub/sub.

Parity: ``fedml_mobile/server/executor/{app.py,mobile_client_simulator.py,log.py,conf/*}``:
``POST /api/register?device_id=…`` → ``{errno, executorId, executorTopic, client_id, training_task_args{…}}`` and the
MQTT topic scheme (server publishes ``fedml0_<cid>``, subscribes ``fedml<cid>``; payload = JSON with tensors as nested
lists — ``is_mobile = 1``).  Flask / gunicorn / paho are not in this image, so the HTTP endpoint is a stdlib
``ThreadingHTTPServer`` and the broker is the in-process :class:`LocalBroker` (a real paho client is used when
importable and a host is given).  The Android/Java SDK is a non-goal; this is the server side + a Python client
simulator speaking the same protocol.
"""
from __future__ import annotations

import copy
import json
import logging
import logging.handlers
import os
import threading
import urllib.parse
import urllib.request
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Dict, Optional

from ..core.comm.mqtt import LocalBroker
from ..fl.fedavg import FedAVGAggregator, FedAvgClientManager, FedAvgServerManager, FedAVGTrainer

TASK_KEYS = ("dataset", "data_dir", "partition_method", "partition_alpha", "model", "client_num_per_round", "comm_round",
             "epochs", "lr", "wd", "batch_size", "frequency_of_the_test", "is_mobile")


def make_logger(name: str = "fedml_mobile", path: Optional[str] = None, level=logging.INFO) -> logging.Logger:
    """Rotating-file + console logger (parity: ``log.py:15-60``)."""
    log = logging.getLogger(name)
    log.setLevel(level)
    if not log.handlers:
        fmt = logging.Formatter("%(asctime)s %(levelname)s %(filename)s[%(lineno)d] %(message)s")
        sh = logging.StreamHandler()
        sh.setFormatter(fmt)
        log.addHandler(sh)
        if path:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            fh = logging.handlers.RotatingFileHandler(path, maxBytes=10 << 20, backupCount=5)
            fh.setFormatter(fmt)
            log.addHandler(fh)
    return log


def load_conf(path: str) -> Dict:
    """YAML configuration (parity: ``conf/conf.py`` + ``conf.yaml``)."""
    import yaml
    with open(path) as fh:
        return yaml.safe_load(fh) or {}


class DeviceRegistry:
    def __init__(self, args):
        self.args, self.map, self.lock = args, {}, threading.Lock()

    def register(self, device_id: str) -> Dict:
        with self.lock:
            if device_id not in self.map:
                self.map[device_id] = len(self.map) + 1
            cid = self.map[device_id]
        task = {k: getattr(self.args, k, None) for k in TASK_KEYS}
        return {"errno": 0, "executorId": "executorId", "executorTopic": "executorTopic", "client_id": cid,
                "training_task_args": task}


def make_http_server(registry: DeviceRegistry, host: str = "127.0.0.1", port: int = 5000) -> ThreadingHTTPServer:
    class Handler(BaseHTTPRequestHandler):
        def do_POST(self):  # noqa: N802
            url = urllib.parse.urlparse(self.path)
            if url.path != "/api/register":
                self.send_error(404)
                return
            q = urllib.parse.parse_qs(url.query)
            if "device_id" not in q:
                length = int(self.headers.get("Content-Length", 0) or 0)
                q = urllib.parse.parse_qs(self.rfile.read(length).decode()) if length else q
            if "device_id" not in q:
                self.send_error(400, "device_id required")
                return
            body = json.dumps(registry.register(q["device_id"][0])).encode()
            self.send_response(200)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, *a):  # quiet
            pass

    return ThreadingHTTPServer((host, port), Handler)


def register(url: str, device_id: str) -> Dict:
    """Client side of the registration (parity: ``mobile_client_simulator.py:36-49``)."""
    req = urllib.request.Request(f"{url}/api/register?device_id={urllib.parse.quote(device_id)}", method="POST", data=b"")
    with urllib.request.urlopen(req, timeout=10) as r:
        return json.loads(r.read().decode())


class MobileFedAvgServer:
    """HTTP registration endpoint + MQTT FedAvg server manager (``app.py:225-232``)."""

    def __init__(self, args, dataset, model, device="cpu", host="127.0.0.1", port=0, broker: Optional[LocalBroker] = None):
        [train_num, test_num, train_global, test_global, local_num, train_local, test_local, class_num] = dataset[:8]
        args.is_mobile = 1
        self.args, self.broker = args, broker or LocalBroker()
        self.registry = DeviceRegistry(args)
        self.http = make_http_server(self.registry, host, port)
        self.url = f"http://{self.http.server_address[0]}:{self.http.server_address[1]}"
        self.aggregator = FedAVGAggregator(train_global, test_global, train_num, train_local, test_local, local_num,
                                           args.client_num_per_round, device, model, args)
        self.manager = FedAvgServerManager(args, self.aggregator, self.broker, 0, args.client_num_per_round + 1, backend="MQTT")
        self.manager.register_message_receive_handlers()
        self._thread = threading.Thread(target=self.http.serve_forever, daemon=True)

    def start_http(self):
        self._thread.start()
        return self.url

    def stop(self):
        self.http.shutdown()
        self.http.server_close()


class MobileClientSimulator:
    """A phone: registers over HTTP, then runs ``FedAVGTrainer`` + ``FedAvgClientManager(backend='MQTT')``."""

    def __init__(self, device_id: str, server_url: str, dataset, model, broker: LocalBroker, device="cpu"):
        from types import SimpleNamespace
        info = register(server_url, device_id)
        self.client_id = int(info["client_id"])
        a = SimpleNamespace(**info["training_task_args"])
        a.client_optimizer = getattr(a, "client_optimizer", "sgd")
        a.client_num_in_total = getattr(a, "client_num_in_total", a.client_num_per_round)
        a.dummy_arg = 0
        self.args = a
        [train_num, _, _, _, local_num, train_local, _, _] = dataset[:8]
        trainer = FedAVGTrainer(self.client_id - 1, train_local, local_num, train_num, device, copy.deepcopy(model), a)
        self.manager = FedAvgClientManager(a, trainer, broker, self.client_id, a.client_num_per_round + 1, backend="MQTT")
        self.manager.register_message_receive_handlers()


def run_mobile_federation(server: MobileFedAvgServer, clients, max_spins: int = 100000) -> int:
    """Single-threaded pump of the MQTT mailboxes until the server finishes its rounds."""
    server.manager.send_init_msg()
    spins = 0
    while not server.manager.finished and spins < max_spins:
        moved = sum(c.manager.com_manager.poll() for c in clients) + server.manager.com_manager.poll()
        spins += 1
        if moved == 0 and not server.manager.finished:
            break
    return server.manager.round_idx
