cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -6
rm -f gpurun_out/r2o_tune.jsonl
timeout 400 python tools/tune_stream.py --n 128 --grid voronoi --out gpurun_out/r2o_tune.jsonl - 2>&1 | tail -1 | cut -c1-300
timeout 300 python tools/tune_stream.py --n 64 --grid voronoi --out gpurun_out/r2o_tune.jsonl - SSW_WALK=0 SSW_WALK=1 2>&1 | tail -3 | cut -c1-300
