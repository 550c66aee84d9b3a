"""One MPI-rank worth of the UNMODIFIED reference (``fedml_experiments/distributed/fedavg_cont_ens/main_fedavg.py``)
executed with ``runpy`` under the shims.  Rank 0 gets a timing probe: ``test_on_all_clients`` is the last thing the
server does in a round (``FedAvgEnsServerManager.py:46-47``), so wrapping it (instrumentation only — the reference
code itself is untouched) yields an end-of-round timestamp after a ``torch.cuda.synchronize``."""
import json
import os
import runpy
import sys
import time


def main():
    ref_root = os.environ["FDB_REF_ROOT"]
    exp_dir = os.path.join(ref_root, "fedml_experiments", "distributed", os.environ.get("FDB_REF_EXP", "fedavg_cont_ens"))
    os.chdir(exp_dir)
    rank = int(os.environ.get("RANK", "0"))
    out_path = os.environ.get("FDB_REF_TIMING", "")
    if rank == 0 and out_path:
        import importlib
        import torch
        stamps = []

        def wrap(cls):
            orig = cls.test_on_all_clients

            def timed(self, round_idx):
                r = orig(self, round_idx)
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                stamps.append(time.perf_counter())
                with open(out_path, "w") as fh:
                    json.dump({"round_end": stamps}, fh)
                return r

            cls.test_on_all_clients = timed

        for name in ("SoftCluster", "Aue", "AuePc", "Kue", "DriftSurf", "MultiModelAcc", "ClusterFL", "Ada", "Vanilla"):
            try:   # whichever aggregator the run's algorithm selects (FedAvgEnsAPI.py:95-141)
                mod = importlib.import_module(f"fedml_api.distributed.fedavg_ens.FedAvgEnsAggregator{name}")
                wrap(getattr(mod, f"FedAvgEnsAggregator{name}"))
            except Exception:
                pass
        try:   # fedavg_cont_one (single-model window baselines) uses the plain FedAvg aggregator
            mod = importlib.import_module("fedml_api.distributed.fedavg.FedAVGAggregator")
            wrap(mod.FedAVGAggregator)
        except Exception:
            pass
    metrics_path = os.environ.get("FDB_REF_METRICS", "")
    if rank == 0 and metrics_path:
        # observation only: mirror what the reference sends to wandb.log (Train/Acc, Test/Acc, … keyed by round) into a
        # JSON-lines file so that accuracy trajectories can be compared (tools/e2e_parity.py)
        import wandb

        def install():
            orig_log = wandb.log

            def log(data=None, *a, **k):
                try:
                    row = {str(kk): (float(v) if isinstance(v, (int, float)) else str(v)) for kk, v in dict(data or {}).items()}
                    with open(metrics_path, "a") as fh:
                        fh.write(json.dumps(row) + "\n")
                except Exception:
                    pass
                return orig_log(data, *a, **k)

            wandb.log = log

        orig_init = wandb.init

        def init(*a, **k):   # wandb.init rebinds wandb.log to the run's method: wrap again afterwards
            run = orig_init(*a, **k)
            install()
            return run

        wandb.init = init
        install()
    if os.environ.get("FDB_REF_NOSLEEP") == "1":
        # "sleep-removed" variant of the baseline: the reference's event loop sleeps 0.3 s after every
        # dispatched message (com_manager.py:71-79) and its send thread sleeps 0.3 s when idle (mpi_send_thread.py:28-29).
        # Only those two modules get a `time` whose sleep() yields for 50 us instead; the sources stay untouched.
        import importlib
        import types
        quick = types.SimpleNamespace(**{k: getattr(time, k) for k in dir(time) if not k.startswith("_")})
        _real_sleep = time.sleep
        quick.sleep = lambda s: _real_sleep(min(s, 5e-5))
        for name in ("fedml_core.distributed.communication.mpi.com_manager",
                     "fedml_core.distributed.communication.mpi.mpi_send_thread"):
            importlib.import_module(name).time = quick
    script = os.environ.get("FDB_REF_SCRIPT", "main_fedavg.py")
    sys.argv = [script] + sys.argv[1:]
    runpy.run_path(os.path.join(exp_dir, script), run_name="__main__")


if __name__ == "__main__":
    main()
