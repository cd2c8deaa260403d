set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rA 2>&1 | tail -150 > gpurun_out/r2a_tests.log
tail -40 gpurun_out/r2a_tests.log
rm -f gpurun_out/r2a_tune.jsonl
timeout 900 python tools/tune_stream.py --n 128 --grid voronoi --out gpurun_out/r2a_tune.jsonl \
  SSW_STREAM_SOLO=0 SSW_STREAM_SOLO=0,SSW_STREAM_GROUPS=4 SSW_STREAM_SOLO=0,SSW_STREAM_GROUPS=8 \
  - SSW_SOLO_TILE=1024 SSW_SOLO_TILE=256,SSW_SOLO_STAGES=8 SSW_SOLO_THREADS=512,SSW_SOLO_TILE=512 SSW_SOLO_THREADS=512,SSW_SOLO_TILE=256,SSW_SOLO_STAGES=8 2>&1 | tail -30
timeout 600 python tools/tune_stream.py --n 32 --grid voronoi --out gpurun_out/r2a_tune.jsonl \
  SSW_STREAM_SOLO=0 - SSW_SOLO_THREADS=512,SSW_SOLO_TILE=256 SSW_SOLO_THREADS=256,SSW_SOLO_TILE=128 SSW_SOLO_THREADS=256,SSW_SOLO_TILE=64,SSW_SOLO_STAGES=8 2>&1 | tail -12
timeout 600 python tools/tune_stream.py --n 128 --grid cartesian --out gpurun_out/r2a_tune.jsonl \
  - SSW_PATCH=0 SSW_PATCH=0,SSW_SOLO_TILE=1024 SSW_PATCH=0,SSW_STREAM_SOLO=0 2>&1 | tail -12
timeout 600 python bench.py --workload front --dirs 21 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_front.json 2> gpurun_out/r2a_front.err
cat gpurun_out/r2a_front.json
timeout 600 python bench.py --workload front --front-scale 0.03 --dirs 21 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_front003.json 2> gpurun_out/r2a_front003.err
cat gpurun_out/r2a_front003.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chemistry_kernel --launch-skip 24 --launch-count 6 -o gpurun_out/r2a_chem_front -f \
  python bench.py --workload front --front-scale 0.03 --dirs 21 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_ncu_chem.log 2>&1
tail -5 gpurun_out/r2a_ncu_chem.log
ls -la gpurun_out | tail -20
