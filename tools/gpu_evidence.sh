# Round evidence on one B200: GPU tests, smoke, bench lines, ncu launch list and --set full captures.
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh r2z'
# Everything lands under gpurun_out/<tag>_*; tools/ncu_summary.py / tools/launch_summary.py turn the reports into the text
# kept in profiles/.
TAG=${1:-r2z}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/$TAG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > ${O}_tests.log
tail -3 ${O}_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
# bench lines (the driver's command first)
timeout 600 python bench.py > ${O}_bench_n1.json 2> ${O}_bench_n1.err
cut -c1-300 ${O}_bench_n1.json
timeout 600 python bench.py --grid voronoi --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_voronoi_n1.json 2> ${O}_bench_voronoi_n1.err
cut -c1-300 ${O}_bench_voronoi_n1.json
timeout 300 python bench.py --workload front --dirs 21 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > ${O}_front.json 2> ${O}_front.err
timeout 300 python bench.py --workload front --front-scale 0.03 --dirs 21 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > ${O}_front003.json 2> ${O}_front003.err
cut -c1-500 ${O}_front.json ${O}_front003.json
# launch list of the bench command (time only, one pass per kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_ncu_bench.log 2>&1
# full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:patch_sweep_kernel --launch-skip 5 --launch-count 1 -o ${O}_patch84 -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_ncu_patch84.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_stream_kernel --launch-skip 5 --launch-count 1 -o ${O}_stream_voronoi -f \
  python bench.py --grid voronoi --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_ncu_stream_voronoi.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chemistry_kernel --launch-skip 40 --launch-count 2 -o ${O}_chem_front -f \
  python bench.py --workload front --front-scale 0.03 --dirs 21 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > ${O}_ncu_chem.log 2>&1
ls -la gpurun_out | grep $TAG
