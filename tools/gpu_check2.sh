#!/bin/bash
# full GPU suite, default bench line, shard emulation with the default settings
set -u
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/check2.log; : > $LOG
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee -a $LOG
timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/check2_bench.json 2> $OUT/check2_bench.err
tail -3 $OUT/check2_bench.err | tee -a $LOG
cat $OUT/check2_bench.json | tee -a $LOG
for W in 2 4 8; do
  echo "## emulate shard $W" | tee -a $LOG
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --emulate-shard $W 2>&1 | tail -1 | cut -c1-600 | tee -a $LOG
done
echo "## stream form, 1 GPU" | tee -a $LOG
SSW_PATCH=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-600 | tee -a $LOG
