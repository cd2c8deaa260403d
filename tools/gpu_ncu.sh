#!/bin/bash
# one ncu --set full capture of a kernel inside a short bench run.  usage: tools/gpu_ncu.sh <tag> <kernel-regex> [skip]
set -u
TAG=$1; KREGEX=$2; SKIP=${3:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${SKIP} -c 1 \
    -f -o $OUT/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} \
    > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log
