#!/bin/bash
# under gpurun --gpus 8: the bench at N = 8, 4, 2 ranks (torchrun), one line each
set -u
OUT=gpurun_out; mkdir -p $OUT
for N in ${NS:-8 4 2}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) \
      bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > $OUT/scale_n$N.json 2> $OUT/scale_n$N.err
  tail -2 $OUT/scale_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/scale_n$N.json").read().strip().splitlines()[-1])
    t=d["timing"]
    print("N=$N", "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "all-cells sweep ms %.3f" % d["roofline"]["ms_per_launch"],
          "sweep %.3f chem %.3f allreduce %.3f levels %.3f (ms/step)" % tuple(t[k] / d["steps"] for k in ("sweep_ms", "chemistry_ms", "allreduce_ms", "update_levels_ms")), t["checksum"])
except Exception as e: print("N=$N failed", e)
PY
done
