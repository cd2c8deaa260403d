#!/bin/bash
# quick GPU check: parity tests + one bench line (no profiler)
set -u
TAG=${1:-chk}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > $OUT/${TAG}_tests.log
tail -15 $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps 4 --warmup 3 ${BENCH_ARGS:-} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    print(d["timing"])
except Exception as e: print("no bench line", e)
PY
