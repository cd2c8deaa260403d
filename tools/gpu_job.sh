cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 120 -x 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_n2_peers.json 2> gpurun_out/r2q_n2_peers.err
tail -c 400 gpurun_out/r2q_n2_peers.err; cat gpurun_out/r2q_n2_peers.json | cut -c1-300
