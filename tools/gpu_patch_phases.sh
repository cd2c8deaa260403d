#!/bin/bash
# where does a tile's time go?  per-phase cycle counters of thread 0 (SSW_STREAM_PROFILE) + two what-if switches
set -u
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/patch_phases.log; : > $LOG
for cfg in "512 6 128 0" "64 11 128 0" "512 6 128 8"; do
  set -- $cfg
  export SSW_PATCH_CELLS=$1 SSW_PATCH_KD=$2 SSW_PATCH_THREADS=$3
  EXTRA=""; [ "$4" != "0" ] && EXTRA="--emulate-shard $4"
  for exp in 0 1 2 3; do
    echo "## cells=$1 kd=$2 threads=$3 shard=$4 exp=$exp" | tee -a $LOG
    SSW_PATCH_EXP=$exp SSW_STREAM_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA 2>&1 | grep "^\[patch" | tail -2 | tee -a $LOG
  done
done
