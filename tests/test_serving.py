from types import SimpleNamespace

import numpy as np

from feddrift_b200.core.comm.mqtt import LocalBroker
from feddrift_b200.data.drift import generate_drift_data, load_partition_data
from feddrift_b200.models import create_model
from feddrift_b200.serving.mobile import MobileClientSimulator, MobileFedAvgServer, register, run_mobile_federation
from feddrift_b200.utils.metrics import MetricsSink, set_sink


def test_mobile_register_and_mqtt_fedavg_rounds():
    sink = set_sink(MetricsSink())
    d = generate_drift_data("sine", 1, 3, 40, 0.0, 1, np.zeros((2, 3), dtype=np.int64))
    ds = list(load_partition_data(d, 20, 0, "win-1", rng=np.random.RandomState(0))[1:])
    args = SimpleNamespace(dataset="sine", data_dir="", partition_method="homo", partition_alpha=0.5, model="fnn",
                           client_num_per_round=3, client_num_in_total=3, comm_round=3, epochs=2, lr=0.05, wd=0.0, batch_size=20,
                           frequency_of_the_test=1, is_mobile=1, client_optimizer="sgd", report_client=0, ci=0, dummy_arg=0)
    broker = LocalBroker()
    model = create_model("fnn", 2, 2)
    server = MobileFedAvgServer(args, ds, model, broker=broker)
    url = server.start_http()
    try:
        r1, r2 = register(url, "phone-A"), register(url, "phone-A")
        assert r1["errno"] == 0 and r1["client_id"] == r2["client_id"] == 1
        assert r1["training_task_args"]["comm_round"] == 3 and r1["training_task_args"]["is_mobile"] == 1
        clients = [MobileClientSimulator(f"phone-{n}", url, ds, model, broker) for n in "ABC"]
        assert sorted(c.client_id for c in clients) == [1, 2, 3]
        rounds = run_mobile_federation(server, clients)
    finally:
        server.stop()
    assert rounds == 3 and len(sink.series("Test/Acc")) == 3
    assert broker.published >= 3 * 3 * 2   # JSON messages really went through the topic fabric
