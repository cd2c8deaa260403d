cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -8
rm -f gpurun_out/r2m_tune.jsonl
timeout 400 python tools/tune_stream.py --n 128 --grid voronoi --out gpurun_out/r2m_tune.jsonl - SSW_WALK_GROUPS=1 SSW_WALK_GROUPS=4,SSW_WALK_THREADS=128 SSW_WALK=0,SSW_STREAM_GROUPS=4 2>&1 | tail -4 | cut -c1-300
timeout 300 python tools/tune_stream.py --n 32 --grid voronoi --out gpurun_out/r2m_tune.jsonl - SSW_WALK_GROUPS=1 SSW_WALK_GROUPS=4,SSW_WALK_THREADS=128 SSW_WALK_GROUPS=1,SSW_WALK_THREADS=128 SSW_WALK=0 2>&1 | tail -5 | cut -c1-300
timeout 300 python tools/tune_stream.py --n 128 --grid cartesian --out gpurun_out/r2m_tune.jsonl SSW_PATCH=0 2>&1 | tail -1 | cut -c1-300
