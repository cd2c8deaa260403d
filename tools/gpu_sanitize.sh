#!/bin/bash
# compute-sanitizer over small parity cases: memcheck on all kernels, racecheck + synccheck on the shared-memory ones
set -u
OUT=gpurun_out; mkdir -p $OUT
T="tests/test_gpu_paths.py::test_all_paths_bitwise_identical tests/test_gpu_sweep_parity.py::test_ragged_and_tiny_grids tests/test_gpu_time_series.py"
for tool in memcheck racecheck synccheck; do
  echo "## compute-sanitizer --tool $tool" | tee -a $OUT/sanitize.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T -m gpu -x -q 2>&1 | tail -6 | tee -a $OUT/sanitize.log
  echo "exit code: ${PIPESTATUS[0]}" | tee -a $OUT/sanitize.log
done
