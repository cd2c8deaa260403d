#!/bin/bash
# tuning sweep of the patch-ordered sweep: lines of "patch_cells kd threads stages [emulate_shard]"
set -u
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/tune_patch.jsonl
: > $LOG
if [ "${RUN_TESTS:-0}" = "1" ]; then
  timeout 1200 python -m pytest ${TESTS:-tests} -m gpu -q -x 2>&1 | tail -8 | tee $OUT/tune_patch_tests.log
fi
while read -r pc kd th st sh; do
  [ -z "${pc:-}" ] && continue
  echo "## cells=$pc kd=$kd threads=$th stages=$st shard=${sh:-0}" | tee -a $LOG
  export SSW_PATCH_CELLS=$pc SSW_PATCH_KD=$kd SSW_PATCH_THREADS=$th SSW_PATCH_STAGES=$st
  EXTRA=""; [ "${sh:-0}" != "0" ] && EXTRA="--emulate-shard $sh"
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('ms_per_step','all_cells_sweep_ms','sweep_ms','chemistry_ms','macro_tiles','patch_levels','mean_xhii','all_cells_form')})
except Exception as e: print('failed', e)" | tee -a $LOG
  if [ "${PROFILE_EACH:-1}" = "1" ]; then
    SSW_STREAM_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA ${BENCH_ARGS:-} 2>&1 | grep "profile\]" | tail -1 | tee -a $LOG
  fi
done < ${1:-tools/tune_patch_configs.txt}
