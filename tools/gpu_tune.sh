#!/bin/bash
# tuning sweep of the stream kernel's launch geometry: lines of "threads blocks_per_sm groups stages"
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sweep_parity.py tests/test_gpu_paths.py -m gpu -q -x 2>&1 | tail -5
: > $OUT/tune.jsonl
while read -r t b g st; do
  [ -z "$t" ] && continue
  SSW_STREAM_THREADS=$t SSW_STREAM_BPS=$b SSW_STREAM_GROUPS=$g SSW_STREAM_STAGES=$st \
    timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} 2>&1 | tail -2 | tee -a $OUT/tune.jsonl
done < ${1:-tools/tune_configs.txt}
