#!/bin/bash
# Runs on the GPU box (under gpurun): full GPU test-suite, a bench line, the ncu launch list of
# the same bench command and one --set full capture of the dominant sweep kernel.
# usage: tools/gpu_profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r1}
KREGEX=${2:-sweep_stream_kernel}
OUT=gpurun_out
mkdir -p $OUT
python -m subsweep_b200.build >/dev/null
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps 4 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 2 -c 1 \
    -f -o $OUT/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_tests.log
cat $OUT/${TAG}_bench.json
ls -la $OUT
