#!/bin/bash
# tuning sweep of the stream kernel: lines of "threads blocks_per_sm groups stages [balanced] [interleave] [poll_ns]"
set -u
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/tune.jsonl
if [ "${RUN_TESTS:-1}" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/tune_tests.log
fi
while read -r t b g st bal il pns; do
  [ -z "$t" ] && continue
  echo "## threads=$t bps=$b groups=$g stages=$st balanced=${bal:-1} interleave=${il:-1} poll_ns=${pns:-20}" | tee -a $OUT/tune.jsonl
  export SSW_STREAM_THREADS=$t SSW_STREAM_BPS=$b SSW_STREAM_GROUPS=$g SSW_STREAM_STAGES=$st SSW_STREAM_BALANCED=${bal:-1} SSW_STREAM_INTERLEAVE=${il:-1} SSW_STREAM_POLL_NS=${pns:-20}
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} 2>&1 | tail -2 | tee -a $OUT/tune.jsonl
  if [ "${PROFILE_EACH:-0}" = "1" ]; then
    SSW_STREAM_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} 2>&1 | grep "stream profile" | tail -1 | tee -a $OUT/tune.jsonl
  fi
done < ${1:-tools/tune_configs.txt}
