#!/bin/bash
# multi-GPU run under gpurun --gpus N: bench at N ranks (torchrun) and, if N matches, the reference arm
set -u
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 4 --warmup 3 ${BENCH_ARGS:-} > $OUT/scale_n$N.json 2> $OUT/scale_n$N.err
tail -3 $OUT/scale_n$N.err
cat $OUT/scale_n$N.json
