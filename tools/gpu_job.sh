cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -6
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r2k_bench_n1.json').read().strip().splitlines()[-1]); b=r["timing"]["breakdown"]
print("N1 ms/step %.3f value %.3e e2e %.3e frac %.3f kernel %.3f ms | breakdown: step %.3f sweep %.3f chem %.3f lvl %s" % (r["ms_per_step"], r["value"], r["e2e"]["value"], r["roofline"]["frac"], r["roofline"]["ms_per_launch"], b["ms_per_step"], b["sweep_ms"], b["chemistry_ms"], [round(x,3) for x in b["sweep_level_ms"]]), r["timing"]["checksum"])
PY
tail -3 gpurun_out/r2k_bench_n1.err
