#!/bin/bash
# Race / memory checking of the native kernels on a GPU box (the reference has no sanitizer story).
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh'
# memcheck + racecheck (shared-memory hazards of the fused round kernel, the GEMM pipeline, the reductions) and
# synccheck (barrier misuse) over the GPU unit tests; reports land in gpurun_out/sanitizer_*.log.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 \
      python -m pytest tests/test_gpu_small_round.py tests/test_gpu_kernels.py -x -q -k "not tcgen05 and not tclinear" \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitizer_$tool.log) summaries: $(grep 'ERROR SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
