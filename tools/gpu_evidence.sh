# Round evidence on one B200: GPU tests, bench lines, ncu launch lists and --set full captures.
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh r2f'
# Everything lands under gpurun_out/<tag>_*; tools/ncu_summary.py turns the .ncu-rep files into the text kept in profiles/.
TAG=${1:-r2f}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/$TAG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > ${O}_tests.log
  tail -3 ${O}_tests.log
fi
# bench lines (the driver's command first)
timeout 600 python bench.py > ${O}_bench_n1.json 2> ${O}_bench_n1.err
cut -c1-400 ${O}_bench_n1.json
timeout 600 python bench.py --grid voronoi --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_voronoi_n1.json 2> ${O}_bench_voronoi_n1.err
cut -c1-400 ${O}_bench_voronoi_n1.json
timeout 300 python bench.py --workload front --dirs 21 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > ${O}_front.json 2> ${O}_front.err
timeout 300 python bench.py --workload front --front-scale 0.03 --dirs 21 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > ${O}_front003.json 2> ${O}_front003.err
cut -c1-700 ${O}_front.json ${O}_front003.json
timeout 300 python bench.py --emulate-shard 8 --steps 20 --warmup 4 --no-e2e --no-cpu-baseline > ${O}_shard8.json 2> ${O}_shard8.err
cut -c1-700 ${O}_shard8.json
# launch lists (time only, one pass per kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_shard8.csv \
  python bench.py --emulate-shard 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_ncu_shard8.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_voronoi.csv \
  python bench.py --grid voronoi --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_ncu_voronoi.log 2>&1
# full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:patch_sweep_kernel --launch-skip 5 --launch-count 1 -o ${O}_patch84 -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_ncu_patch84.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel --launch-skip 5 --launch-count 1 -o ${O}_walk_voronoi -f \
  python bench.py --grid voronoi --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_ncu_walk.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chemistry_kernel --launch-skip 40 --launch-count 2 -o ${O}_chem_front -f \
  python bench.py --workload front --front-scale 0.03 --dirs 21 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > ${O}_ncu_chem.log 2>&1
ls -la gpurun_out | grep $TAG
