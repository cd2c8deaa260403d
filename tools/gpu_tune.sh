#!/bin/bash
# tuning sweep of the stream kernel's launch geometry: lines of "threads blocks_per_sm groups stages"
set -u
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/tune.jsonl
while read -r t b g st; do
  [ -z "$t" ] && continue
  echo "## threads=$t bps=$b groups=$g stages=$st" | tee -a $OUT/tune.jsonl
  SSW_STREAM_THREADS=$t SSW_STREAM_BPS=$b SSW_STREAM_GROUPS=$g SSW_STREAM_STAGES=$st \
    timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} 2>&1 | tail -2 | tee -a $OUT/tune.jsonl
done < ${1:-tools/tune_configs.txt}
echo "## profile hook (default geometry)" | tee -a $OUT/tune.jsonl
SSW_STREAM_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} 2>&1 | tail -4 | tee -a $OUT/tune.jsonl
