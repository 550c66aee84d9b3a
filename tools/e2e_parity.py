"""End-to-end trajectory parity on IDENTICAL data: the unmodified reference pipeline (prepare_data.py + main_fedavg.py with
10+1 ranks over the gloo MPI shim, CPU or GPU) vs the feddrift_b200 device engine fed with the CSV files the reference
generated.  Batch sampling RNGs differ, so the comparison is statistical: per time step, the final Train/Acc and
Test/Acc of both runs.

    python tools/e2e_parity.py --steps 3 --rounds 10        # ≈ 1 min per reference time step (0.3 s polling sleeps)
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import run_reference as rr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3, help="time steps 0..steps-1 are trained")
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--algo", type=str, default="softcluster")
    ap.add_argument("--algo_arg", type=str, default="H_A_C_1_10_0")
    ap.add_argument("--strict_ref", type=int, default=0, help="1: also reproduce the reference's known quirks (e.g. AUE weight shift)")
    ap.add_argument("--change_points", type=str, default="A")
    ap.add_argument("--cont_one", type=str, default="", help="run fedavg_cont_one with this --retrain_data (win-1, win-2, all, …)")
    a = ap.parse_args()
    rr.ALGO = (a.algo, a.algo_arg)
    rr.CHANGE_POINTS = a.change_points
    if a.cont_one:
        rr.CONT_ONE = a.cont_one
        a.algo, a.algo_arg = a.cont_one, ""
    if not os.path.isdir(os.path.join(rr.REF, "fedml_api")):
        from baseline import install_reference
        assert install_reference.main() == 0
    total_iter = 10
    for f in ("model_params.pt", "sc_state.pkl", "output.log", "ds_state.pkl", "kue_state.pkl", "ada_state.pkl", "mm_state.pkl"):
        for d in (rr.EXP, rr.EXP.replace("fedavg_cont_ens", "fedavg_cont_one")):
            p = os.path.join(d, f)
            if os.path.exists(p):
                os.remove(p)
    prep = subprocess.run([sys.executable, os.path.join(rr.HERE, "ref_rank.py"), "--dataset", "sea", "--data_dir", "./../../../data/",
                           "--sample_num", "100", "--noise_prob", "0", "--partition_method", "homo", "--client_num_in_total",
                           str(rr.CLIENTS), "--client_num_per_round", str(rr.CLIENTS), "--batch_size", "500", "--train_iteration",
                           str(total_iter), "--drift_together", "0", "--time_stretch", "1", "--change_points", a.change_points],
                          env=rr._env({"FDB_REF_SCRIPT": "prepare_data.py", "RANK": "0", "WORLD_SIZE": "1"}), cwd=rr.EXP,
                          capture_output=True, text=True)
    assert prep.returncode == 0, prep.stderr[-500:]
    metrics = os.path.join(rr.HERE, "_ref_metrics.jsonl")
    if os.path.exists(metrics):
        os.remove(metrics)
    os.environ["FDB_REF_METRICS"] = metrics
    ref_hist = []
    for it in range(a.steps):
        n0 = sum(1 for _ in open(metrics)) if os.path.exists(metrics) else 0
        stamps, wall, err = rr._run_time_step(0, a.rounds, it, total_iter, os.path.join(rr.HERE, "_ref_timing.json"),
                                              a.rounds * 8 + 180)
        rows = [json.loads(l) for l in open(metrics)][n0:] if os.path.exists(metrics) else []
        last = {}
        for r in rows:
            for k in ("Train/Acc", "Test/Acc"):
                if k in r:
                    last[k] = r[k]
        ref_hist.append({"iteration": it, "rounds_done": len(stamps), **last})
        print("reference", ref_hist[-1], flush=True)

    # ours, on the reference's own CSV files
    import numpy as np
    import torch
    from feddrift_b200.data import changepoints
    from feddrift_b200.data.drift import DriftData
    from feddrift_b200.sim import DriftSim, make_args
    from feddrift_b200.utils.metrics import MetricsSink
    data_dir = os.path.join(rr.REF, "data", "sea")
    cand = [d for d, _, fs in os.walk(data_dir) if any(f.startswith("client_0_iter_0") for f in fs)]
    assert cand, f"no generated CSVs under {data_dir}"
    cp = changepoints.named(a.change_points)
    data = DriftData.from_csv_dir(cand[0], "sea", rr.CLIENTS, total_iter + 1, 2, cp)
    args = make_args(dataset="sea", model="fnn", client_num_in_total=rr.CLIENTS, client_num_per_round=rr.CLIENTS,
                     comm_round=a.rounds, epochs=5, batch_size=500, lr=0.01, total_train_iteration=total_iter, concept_num=4,
                     concept_drift_algo=a.algo, concept_drift_algo_arg=a.algo_arg, change_points=a.change_points, sample_num=100,
                     strict_ref=a.strict_ref)
    sim = DriftSim(args, data=data, device="cuda" if torch.cuda.is_available() else "cpu", sink=MetricsSink())
    ours = []
    for it in range(a.steps):
        out = sim.run_time_step(it, rounds=a.rounds)
        ours.append({"iteration": it, "Train/Acc": out["train_acc"], "Test/Acc": out["test_acc"]})
        print("ours     ", ours[-1], flush=True)
    print("\n| time step | reference Train/Acc | ours Train/Acc | reference Test/Acc | ours Test/Acc |\n|---|---:|---:|---:|---:|")
    for r, o in zip(ref_hist, ours):
        print(f"| {r['iteration']} | {r.get('Train/Acc', float('nan')):.4f} | {o['Train/Acc']:.4f} | "
              f"{r.get('Test/Acc', float('nan')):.4f} | {o['Test/Acc']:.4f} |")
    d = [abs(r.get("Test/Acc", np.nan) - o["Test/Acc"]) for r, o in zip(ref_hist, ours)]
    print("max |Δ Test/Acc| =", max(d))


if __name__ == "__main__":
    main()
