#!/bin/bash
# first contact of the patch-ordered sweep with the GPU: memcheck on one small case, the patch tests, the
# whole GPU suite, then bench lines (patch form, per-block cycle profile, shard emulation)
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "## memcheck" | tee $OUT/patch_first.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_patch.py::test_patch_form_matches_stream_form_on_cartesian_grids" -m gpu -x -q 2>&1 | tail -25 | tee -a $OUT/patch_first.log
echo "## patch tests" | tee -a $OUT/patch_first.log
timeout 900 python -m pytest tests/test_gpu_patch.py tests/test_gpu_paths.py -m gpu -x -q 2>&1 | tail -25 | tee -a $OUT/patch_first.log
echo "## bench (patch)" | tee -a $OUT/patch_first.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/patch_bench.json 2> $OUT/patch_bench.err
tail -5 $OUT/patch_bench.err | tee -a $OUT/patch_first.log
python - <<PY | tee -a $OUT/patch_first.log
import json
try:
    d=json.load(open("$OUT/patch_bench.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    print(d["timing"])
except Exception as e: print("no bench line", e)
PY
echo "## profile" | tee -a $OUT/patch_first.log
SSW_STREAM_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep "profile" | tail -2 | tee -a $OUT/patch_first.log
for W in 2 4 8; do
  echo "## emulate shard $W" | tee -a $OUT/patch_first.log
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --emulate-shard $W 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['ms_per_launch'], d['timing']['sweep_ms'], d['timing']['chemistry_ms'])" | tee -a $OUT/patch_first.log
done
echo "## full GPU suite" | tee -a $OUT/patch_first.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee -a $OUT/patch_first.log
