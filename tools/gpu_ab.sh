#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; LOG=$OUT/ab.log; : > $LOG
timeout 600 python -m pytest tests/test_gpu_patch.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10 python -m pytest "tests/test_gpu_patch.py::test_patch_form_matches_stream_form_on_cartesian_grids[16-True-21-64]" "tests/test_gpu_patch.py::test_patch_form_matches_stream_form_on_cartesian_grids[10-True-16-8]" -m gpu -x -q 2>&1 | grep -E "RACECHECK|passed|failed|hazard detected" | sort | uniq -c | head -8 | tee -a $LOG
for sh in 0 0 2 8; do
  EXTRA=""; [ "$sh" != "0" ] && EXTRA="--emulate-shard $sh"
  echo "## shard=$sh" | tee -a $LOG
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA 2>&1 | tail -1 | cut -c1-300 | tee -a $LOG
done
