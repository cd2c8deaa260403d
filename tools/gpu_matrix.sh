#!/bin/bash
# the other BASELINE.json configs on one GPU (parity-test cases, not bench lines): one JSON line each
set -u
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/matrix.jsonl
run() {
  echo "## $*" | tee -a $OUT/matrix.jsonl
  timeout ${T:-600} python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2> $OUT/matrix.err | tail -1 | tee -a $OUT/matrix.jsonl
  tail -2 $OUT/matrix.err
}
if [ "${VORONOI:-0}" = "1" ]; then T=1500 run --grid voronoi; fi   # Qhull on 2 M points: minutes of host time
run --workload front --dirs 21
for d in 1 16 32 64; do run --dirs $d; done
if [ "${BIG:-0}" = "1" ]; then T=1500 run --n 256; fi
