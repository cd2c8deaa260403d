cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/r2m_bench_n2.json 2>gpurun_out/r2m_bench_n2.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r2m_bench_n2.json').read().strip().splitlines()[-1]); b=r["timing"]["breakdown"]
print("N2 ms/step %.3f value %.3e e2e %.3e kernel %.3f ms | breakdown: step %.3f sweep %.3f chem %.3f wait %.3f lvl %s" % (r["ms_per_step"], r["value"], r["e2e"]["value"], r["roofline"]["ms_per_launch"], b["ms_per_step"], b["sweep_ms"], b["chemistry_ms"], b["exchange_wait_ms"], [round(x,3) for x in b["sweep_level_ms"]]), r["timing"]["checksum"])
PY
tail -2 gpurun_out/r2m_bench_n2.err
