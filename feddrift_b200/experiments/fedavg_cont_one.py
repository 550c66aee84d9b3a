"""``fedml_experiments/distributed/fedavg_cont_one`` — the single-model continual baselines of the README
(``win-1``, ``win-2``, ``all``, ``weight-linear``, ``weight-exp``): the same driver as ``fedavg_cont_ens`` with the
window selector passed as ``--retrain_data`` (it becomes the planner ``WindowAlgo``)."""
from __future__ import annotations

import sys

from .fedavg_cont_ens import main as _main


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    retrain = "win-1"
    if "--retrain_data" in argv:
        retrain = argv[argv.index("--retrain_data") + 1]
    if "--concept_drift_algo" in argv:
        i = argv.index("--concept_drift_algo")
        del argv[i:i + 2]
    return _main(argv + ["--concept_drift_algo", retrain, "--concept_num", "1"])


if __name__ == "__main__":
    main()
