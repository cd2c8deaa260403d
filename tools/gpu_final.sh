#!/bin/bash
# end-of-round evidence on one GPU: sanitizer on small patch cases, full GPU suite, bench line, ncu launch list + full captures
set -u
TAG=${1:-r1h}
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/${TAG}_final.log; : > $LOG
T="tests/test_gpu_patch.py::test_patch_form_matches_stream_form_on_cartesian_grids tests/test_gpu_paths.py::test_all_paths_bitwise_identical"
for tool in memcheck racecheck synccheck; do
  echo "## compute-sanitizer --tool $tool" | tee -a $LOG
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T -m gpu -x -q 2>&1 | tail -4 | tee -a $LOG
  echo "exit code: ${PIPESTATUS[0]}" | tee -a $LOG
done
echo "## full GPU suite" | tee -a $LOG
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee -a $LOG
echo "## bench" | tee -a $LOG
timeout 900 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -2 $OUT/${TAG}_bench.err | tee -a $LOG
cat $OUT/${TAG}_bench.json | tee -a $LOG
echo "## reference arm" | tee -a $LOG
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_ref.json 2>> $LOG
cat $OUT/${TAG}_ref.json | cut -c1-300 | tee -a $LOG
echo "## ncu" | tee -a $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_bench.log 2>&1
bash tools/gpu_ncu.sh ${TAG}_patch84 patch_sweep_kernel 2
BENCH_ARGS="--emulate-shard 8" bash tools/gpu_ncu.sh ${TAG}_patch10 patch_sweep_kernel 2
# the all-cells chemistry launch of the last step: 8 chemistry launches per step, 5 steps before it (4 spin-up... counted from the launch list)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chemistry_kernel -s 32 -c 1 -f -o $OUT/${TAG}_chem_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_chem.log 2>&1
for W in 2 4 8; do
  echo "## emulate shard $W" | tee -a $LOG
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --emulate-shard $W 2>&1 | tail -1 | cut -c1-420 | tee -a $LOG
done
echo "## stream form" | tee -a $LOG
SSW_PATCH=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-420 | tee -a $LOG
ls -la $OUT | grep ${TAG} | tee -a $LOG
