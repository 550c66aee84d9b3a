#!/bin/bash
# Race / memory checking of the native kernels on a GPU box (the reference has no sanitizer story).
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh'
# memcheck over the streaming / evaluation / optimizer kernel tests, memcheck + racecheck over the fused round kernel
# (clusters, DSMEM, named barriers) and the eval-matrix kernel; reports land in gpurun_out/sanitize_*.log and the
# summaries are copied to profiles/sanitizer/README.md.  (The tcgen05 / TMA GEMM tests are left out: compute-sanitizer
# serialises them to minutes per launch.)
set -u
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_kernels.py -q -m gpu \
    -k "cluster_aggregate or robust_clip or weighted or gossip or adam or eval" > gpurun_out/sanitize_memcheck.log 2>&1
echo "kernels memcheck: $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck.log | tail -1)"
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 3 python -m pytest tests/test_gpu_small_round.py -q -m gpu \
      -k "matches_reference and cfg0 or eval_matrix" > gpurun_out/sanitize_round_$tool.log 2>&1
  echo "round kernel $tool: $(grep 'SUMMARY' gpurun_out/sanitize_round_$tool.log | tail -1)"
done
