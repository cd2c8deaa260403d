#!/bin/bash
# ncu evidence for the patch-ordered sweep: launch list of the bench command + full captures at 84 and 10 directions
set -u
TAG=${1:-r1g}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    > $OUT/${TAG}_ncu_bench.log 2>&1
tail -2 $OUT/${TAG}_ncu_bench.log
bash tools/gpu_ncu.sh ${TAG}_patch84 patch_sweep_kernel 2
BENCH_ARGS="--emulate-shard 8" bash tools/gpu_ncu.sh ${TAG}_patch10 patch_sweep_kernel 2
ls -la $OUT | grep ${TAG}
