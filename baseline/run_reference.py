"""Reference arm of bench.py: run the unmodified FedDrift reference (installed in ``baseline/_ref``) on the headline
config through its own stock path — ``prepare_data.py`` then ``main_fedavg.py`` once per time step with
``WORKER_NUM + 1`` ranks (what ``run_fedavg_distributed_pytorch.sh:55-84`` does with mpirun) — and report FL
rounds/sec of the timed time step.  MPI is provided by ``baseline/shims/mpi4py`` (torch.distributed/gloo p2p).
"""
from __future__ import annotations

import json
import os
import shutil
import socket
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SHIMS = os.path.join(HERE, "shims")
EXP = os.path.join(REF, "fedml_experiments", "distributed", "fedavg_cont_ens")
CLIENTS = 10
BENCH_TIME_STEP = 5


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


_ENV_DROP_PREFIXES = ("TORCHELASTIC_", "TORCH_NCCL_", "NCCL_ASYNC", "GROUP_", "ROLE_", "PET_")
_ENV_DROP = ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "MASTER_PORT", "MASTER_ADDR", "OMP_NUM_THREADS")


def _env(extra=None):
    """Child environment for one reference rank.  Everything a launcher (torchrun / torch.distributed.run) exports is
    dropped — in particular ``TORCHELASTIC_USE_AGENT_STORE``, which would turn rank 0 of the reference's own 11-rank gloo
    job into a TCPStore *client* of the outer agent's store and hang the rendezvous — the reference job gets its own
    MASTER_PORT / RANK / WORLD_SIZE from ``_run_time_step``."""
    env = {k: v for k, v in os.environ.items() if k not in _ENV_DROP and not k.startswith(_ENV_DROP_PREFIXES)}
    env["PYTHONPATH"] = os.pathsep.join([SHIMS, REF, env.get("PYTHONPATH", "")])
    env.update({"FDB_REF_SHIMS": "1", "FDB_REF_ROOT": REF, "WANDB_MODE": "disabled", "WANDB_SILENT": "true",
                "MASTER_ADDR": "127.0.0.1", "OMP_NUM_THREADS": "1"})
    if extra:
        env.update(extra)
    return env


ALGO = ("softcluster", "H_A_C_1_10_0")   # (concept_drift_algo, concept_drift_algo_arg); tools/e2e_parity.py overrides it
CHANGE_POINTS = "A"
CONT_ONE = None   # set to a --retrain_data value ("win-1", "all", …) to run fedavg_cont_one (single-model baselines) instead


def _common_flags(gpus: int, rounds: int, it: int, total_iter: int):
    if CONT_ONE is not None:   # fedavg_cont_one/main_fedavg.py: same driver, no drift-algorithm flags
        return ["--gpu_server_num", "1", "--gpu_num_per_server", str(max(gpus, 1)), "--model", "fnn", "--dataset", "sea",
                "--data_dir", "./../../../data/", "--noise_prob", "0", "--client_num_in_total", str(CLIENTS),
                "--client_num_per_round", str(CLIENTS), "--comm_round", str(rounds), "--epochs", "5", "--batch_size", "500",
                "--lr", "0.01", "--ci", "0", "--total_train_iteration", str(total_iter), "--curr_train_iteration", str(it),
                "--reset_models", "0", "--drift_together", "0", "--report_client", "1", "--retrain_data", CONT_ONE,
                "--time_stretch", "1", "--dummy_arg", "0", "--change_points", CHANGE_POINTS]
    return ["--gpu_server_num", "1", "--gpu_num_per_server", str(max(gpus, 1)), "--model", "fnn", "--dataset", "sea",
            "--data_dir", "./../../../data/", "--noise_prob", "0", "--client_num_in_total", str(CLIENTS),
            "--client_num_per_round", str(CLIENTS), "--comm_round", str(rounds), "--epochs", "5", "--batch_size", "500",
            "--lr", "0.01", "--ci", "0", "--total_train_iteration", str(total_iter), "--curr_train_iteration", str(it),
            "--concept_num", "4", "--reset_models", "0", "--drift_together", "0", "--report_client", "1",
            "--retrain_data", "win-1", "--concept_drift_algo", ALGO[0], "--concept_drift_algo_arg", ALGO[1],
            "--time_stretch", "1", "--dummy_arg", "0", "--change_points", CHANGE_POINTS]


def _run_time_step(gpus: int, rounds: int, it: int, total_iter: int, timing_path: str, timeout_s: float):
    """Spawn WORKER_NUM+1 ranks (rank 0 = server).  The job ends when the server calls MPI Abort."""
    port = _free_port()
    world = CLIENTS + 1
    procs = []
    if os.path.exists(timing_path):
        os.remove(timing_path)
    for rank in range(world):
        env = _env({"RANK": str(rank), "WORLD_SIZE": str(world), "MASTER_PORT": str(port),
                    "FDB_REF_TIMING": timing_path if rank == 0 else "",
                    "FDB_REF_EXP": "fedavg_cont_ens" if CONT_ONE is None else "fedavg_cont_one"})
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "ref_rank.py")] +
                                      _common_flags(gpus, rounds, it, total_iter), env=env,
                                      cwd=EXP if CONT_ONE is None else EXP.replace("fedavg_cont_ens", "fedavg_cont_one"),
                                      stdout=subprocess.DEVNULL, stderr=subprocess.PIPE if rank == 0 else subprocess.DEVNULL))
    t0 = time.time()
    err = b""
    try:
        _, err = procs[0].communicate(timeout=timeout_s)
    except subprocess.TimeoutExpired:
        pass
    finally:
        for p in procs:  # exact PIDs we started; never pattern kills
            if p.poll() is None:
                p.kill()
        for p in procs:
            try:
                p.wait(timeout=10)
            except Exception:
                pass
    stamps = []
    if os.path.exists(timing_path):
        with open(timing_path) as fh:
            stamps = json.load(fh)["round_end"]
    return stamps, time.time() - t0, err.decode(errors="replace")[-2000:] if err else ""


def main(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:  # under torchrun only rank 0 drives the reference job (it spawns its own 11 ranks)
        return
    if not os.path.isdir(os.path.join(REF, "fedml_api")):
        from baseline import install_reference
        if install_reference.main() != 0 or not os.path.isdir(os.path.join(REF, "fedml_api")):
            print(json.dumps({"impl": "reference", "unavailable": "reference not installed in baseline/_ref "
                              "(pip install of /root/reference fails: no setup.py/pyproject.toml; see DESIGN.md)"}))
            return
    K, W = int(args.steps), max(int(args.warmup), 1)
    budget_s = float(os.environ.get("FDB_REF_MAX_SECONDS", "240"))
    per_round_est = 0.3 * CLIENTS + 1.0  # ≥ 0.3 s sleep per ingested upload (com_manager.py:71-79)
    K_eff = K
    if (W + K) * per_round_est > budget_s:
        K_eff = max(3, int(budget_s / per_round_est) - W)
    total_iter = 10
    # stale state from a previous run must not leak in
    for f in ("model_params.pt", "sc_state.pkl", "output.log", "model_params.pt.t5", "sc_state.pkl.t5"):
        p = os.path.join(EXP, f)
        if os.path.exists(p):
            os.remove(p)
    # 1) stock data preparation
    prep = subprocess.run([sys.executable, os.path.join(HERE, "ref_rank.py"), "--dataset", "sea", "--data_dir",
                           "./../../../data/", "--sample_num", "100", "--noise_prob", "0", "--partition_method", "homo",
                           "--client_num_in_total", str(CLIENTS), "--client_num_per_round", str(CLIENTS), "--batch_size",
                           "500", "--train_iteration", str(total_iter), "--drift_together", "0", "--time_stretch", "1",
                           "--change_points", "A"], env=_env({"FDB_REF_SCRIPT": "prepare_data.py", "RANK": "0",
                                                              "WORLD_SIZE": "1"}), cwd=EXP, capture_output=True, text=True)
    if prep.returncode != 0:
        print(json.dumps({"impl": "reference", "unavailable": "prepare_data.py failed: " + prep.stderr[-300:].replace("\n", " ")}))
        return
    timing = os.path.join(HERE, "_ref_timing.json")
    # 2) untimed: time steps 0..4 with one round each so that step 5 sees real cluster state
    for it in range(BENCH_TIME_STEP):
        stamps, wall, err = _run_time_step(args.gpus, 1, it, total_iter, timing, 300)
        if not stamps:
            print(json.dumps({"impl": "reference", "unavailable": f"time step {it} produced no round: {err[-300:]}".replace("\n", " ")}))
            return
    # 3) timed time step: W warm-up rounds then K timed rounds — the reference AS SHIPPED (0.3 s polling sleeps included)
    def timed(k_rounds, nosleep):
        os.environ["FDB_REF_NOSLEEP"] = "1" if nosleep else "0"
        # the timed step mutates sc_state.pkl / model_params.pt: run every variant from the same saved state
        for f in ("model_params.pt", "sc_state.pkl"):
            src, bak = os.path.join(EXP, f), os.path.join(EXP, f + ".t5")
            if os.path.exists(bak):
                shutil.copyfile(bak, src)
            elif os.path.exists(src):
                shutil.copyfile(src, bak)
        est = (0.02 if nosleep else per_round_est)
        stamps, wall, err = _run_time_step(args.gpus, W + k_rounds, BENCH_TIME_STEP, total_iter, timing,
                                           (W + k_rounds) * est * 2 + 120)
        if len(stamps) < W + 2:
            return None, f"timed step finished {len(stamps)} rounds: {err[-300:]}".replace("\n", " ")
        done = min(len(stamps) - W, k_rounds)
        elapsed = stamps[W + done - 1] - stamps[W - 1]
        return (done, elapsed), ""

    no_sleep_only = bool(getattr(args, "no_sleep", False))
    res = None
    if not no_sleep_only:
        res, why = timed(K_eff, nosleep=False)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": why}))
            return
    # sleep-removed variant (BASELINE.md "reported both as-is and with the 0.3 s sleeps removed"): same stock code path,
    # only time.sleep inside com_manager / mpi_send_thread shortened by the shim (baseline/ref_rank.py)
    ns, ns_why = timed(K, nosleep=True)
    os.environ.pop("FDB_REF_NOSLEEP", None)
    no_sleep = ({"value": ns[0] / ns[1], "unit": "rounds/s", "steps": ns[0], "ms_per_step": 1e3 * ns[1] / ns[0],
                 "how": "time.sleep(0.3) in com_manager.py:79 and mpi_send_thread.py:29 replaced by a 50 us yield via the shim"}
                if ns else {"unavailable": ns_why})
    if no_sleep_only:
        if ns is None:
            print(json.dumps({"impl": "reference", "unavailable": ns_why}))
            return
        res = ns
    done, elapsed = res
    value = done / elapsed
    from baseline import headline_config
    out = {"metric": "fl_rounds_per_sec", "value": value, "unit": "rounds/s", "n_gpus": int(args.gpus), "steps": done,
           "warmup": W, "ms_per_step": 1e3 * elapsed / done, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
           "config": headline_config(int(args.gpus)),
           "arm_details": {"parallelism": f"{CLIENTS}+1 ranks round-robin on {args.gpus} GPU(s) (init_training_device)",
                           "transport": "mpi4py shim over torch.distributed gloo p2p (pickled CPU state_dicts)",
                           "timing": "server wall clock at end of test_on_all_clients after cuda synchronize",
                           "requested_steps": K, "sleeps": "removed" if no_sleep_only else "as shipped"},
           # the reference's own round already moves every model host<->device through pickled CPU state_dicts, so its
           # end-to-end number IS its round rate; bytes are the pickled payloads per round (M state_dicts to and from N ranks)
           "e2e": {"value": value, "unit": "rounds/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
           "no_sleep": no_sleep,
           "gpu_launches": None, "impl": "reference"}
    try:   # lets the other arm quote the sleep-removed baseline measured on the SAME box (informational extra key)
        with open(os.path.join(HERE, "_ref_last.json"), "w") as fh:
            json.dump({"n_gpus": int(args.gpus), "value": value, "no_sleep": no_sleep}, fh)
    except OSError:
        pass
    print(json.dumps(out))
