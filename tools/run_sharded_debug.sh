#!/bin/bash
# Runs the 2-GPU sharded generic-executor worker of tests/test_gpu_multi.py directly with the fault handler on and
# keeps the full log (pytest only shows the last 2000 characters).
python - <<PY
import runpy
open("/tmp/gw.py", "w").write(runpy.run_path("tests/test_gpu_multi.py")["GENERIC_WORKER"])
PY
FDB_ROOT=$PWD PYTHONFAULTHANDLER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 /tmp/gw.py > gpurun_out/sharded.log 2>&1
grep -n "Fatal\|File \"\|rank\|Error\|line " gpurun_out/sharded.log | grep -v "site-packages/torch/distributed" | head -60
