#!/bin/bash
# One 8-GPU box: headline bench at N=4 and N=8 (N=1,2 are measured on smaller boxes).  Every step is bounded.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
for N in 4 8; do
  timeout 150 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_v2_n$N.log 2>gpurun_out/bench_v2_n$N.err
done
tail -n 1 gpurun_out/bench_v2_n4.log gpurun_out/bench_v2_n8.log | cut -c1-900
tail -n 4 gpurun_out/bench_v2_n4.err gpurun_out/bench_v2_n8.err | cut -c1-300
timeout 120 python -m pytest tests/test_gpu_multi.py tests/test_gpu_small_round.py -m gpu -q -x 2>&1 | tail -5
