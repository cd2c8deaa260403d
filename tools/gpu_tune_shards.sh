#!/bin/bash
# stream-kernel geometry at the per-GPU work of 2- and 4-rank jobs (one GPU, emulated shard)
set -u
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/tune_shards.jsonl
for w in 2 4; do
  for cfg in "256 4 2" "512 2 1" "512 2 2" "256 4 1"; do
    set -- $cfg
    echo "## shard=1/$w threads=$1 bps=$2 groups=$3" | tee -a $OUT/tune_shards.jsonl
    SSW_STREAM_THREADS=$1 SSW_STREAM_BPS=$2 SSW_STREAM_GROUPS=$3 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --emulate-shard $w 2>&1 | tail -1 | tee -a $OUT/tune_shards.jsonl
  done
done
