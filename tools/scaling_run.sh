#!/bin/bash
# One 8-GPU box: headline bench at N=1,2,4,8 plus the peer-memory kernel benches at N=2,4,8.  Every step is bounded.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err
for N in 2 4 8; do
  timeout 300 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n$N.log 2>gpurun_out/bench_n$N.err
  timeout 300 $TR --nproc-per-node $N --master-port $((29700+N)) tools/peer_agg_bench.py > gpurun_out/peer_agg_n$N.log 2>gpurun_out/peer_agg_n$N.err
  timeout 300 $TR --nproc-per-node $N --master-port $((29800+N)) tools/peer_ops_bench.py > gpurun_out/peer_ops_n$N.log 2>gpurun_out/peer_ops_n$N.err
done
tail -n 2 gpurun_out/bench_n*.log | cut -c1-600
cat gpurun_out/peer_agg_n*.log gpurun_out/peer_ops_n*.log | cut -c1-400
tail -n 3 gpurun_out/*.err | cut -c1-300
