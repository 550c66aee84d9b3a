#!/bin/bash
# Positional-argument compatible front end of the reference's
# fedml_experiments/distributed/fedavg_cont_ens/run_fedavg_distributed_pytorch.sh (24 args, README.md:46):
#   sh run_fedavg_distributed.sh 10 10 1 4 fnn homo 200 5 500 0.01 sea ./data/ 100 0 0 10 4 0 0 softcluster H_A_C_1_10_0 1 0 A
# ONE process runs data preparation and every time step (no mpirun-per-time-step loop).
CLIENT_NUM=$1; WORKER_NUM=$2; SERVER_NUM=$3; GPU_NUM_PER_SERVER=$4; MODEL=$5; DISTRIBUTION=$6; ROUND=$7; EPOCH=$8
BATCH_SIZE=$9; LR=${10}; DATASET=${11}; DATA_DIR=${12}; SAMPLE_NUM=${13}; NOISE_PROB=${14}; CI=${15}; TRAIN_ITER=${16}
CONCEPT_NUM=${17}; RESET_MODELS=${18}; DRIFT_TOGETHER=${19}; CL_ALGO=${20}; CL_ALGO_ARG=${21}; TIME_STRETCH=${22}
DUMMY_ARG=${23}; CHANGE_POINTS=${24}
export PYTHONPATH="$(cd "$(dirname "$0")/../.." && pwd):$PYTHONPATH"
exec python -m feddrift_b200.experiments.fedavg_cont_ens \
  --gpu_server_num "$SERVER_NUM" --gpu_num_per_server "$GPU_NUM_PER_SERVER" --model "$MODEL" --dataset "$DATASET" \
  --noise_prob "$NOISE_PROB" --client_num_in_total "$CLIENT_NUM" --client_num_per_round "$WORKER_NUM" \
  --comm_round "$ROUND" --epochs "$EPOCH" --batch_size "$BATCH_SIZE" --lr "$LR" --ci "$CI" \
  --total_train_iteration "$TRAIN_ITER" --concept_num "$CONCEPT_NUM" --reset_models "$RESET_MODELS" \
  --drift_together "$DRIFT_TOGETHER" --report_client 1 --retrain_data win-1 --concept_drift_algo "$CL_ALGO" \
  --concept_drift_algo_arg "$CL_ALGO_ARG" --time_stretch "$TIME_STRETCH" --dummy_arg "$DUMMY_ARG" \
  --sample_num "$SAMPLE_NUM" --change_points "${CHANGE_POINTS:-rand}" "${@:25}"
