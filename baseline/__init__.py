"""Reference arm (unmodified microsoft/FedDrift under shims) + the benchmark config shared by both arms."""


def headline_config(n_gpus: int) -> dict:
    """The ``config`` dict of the bench JSON line — key- and value-identical in both arms (BASELINE.json headline)."""
    return {"model": "FeedForwardNN(3,6,2) SEA-4", "clients": 10, "global_batch": 1000, "seq_len": None,
            "local_steps": 5, "batch_size": 500, "samples_per_client_per_step": 100, "model_slots": 4, "lr": 0.01,
            "optimizer": "adam-amsgrad", "algo": "softcluster H_A_C_1_10_0 (FedDrift), change points A",
            "time_step": 5, "eval": "train+test of every client every round", "parallelism": f"fl-clients-over-{n_gpus}gpu",
            "l2": "inputs re-sent every round; device-timed arm also writes a 256 MiB L2 flush between timed rounds"}
