cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r2n_vor_tune.jsonl
FMT='
import sys, json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d = json.loads(l)
    print(d["config"], "ERR " + d["error"][:150] if "error" in d else "all-cells %.3f ms  step %.3f  xhii %.12e" % (d["all_cells_sweep_ms"], d["step_ms"], d["mean_xhii"]))
'
timeout 400 python tools/tune_stream.py --n 128 --grid voronoi --calls 6 --timed 2 --out gpurun_out/r2n_vor_tune.jsonl \
  - SSW_STREAM_STAGES=2 SSW_STREAM_STAGES=2,SSW_STREAM_BPS=6 SSW_STREAM_THREADS=512 SSW_STREAM_GROUPS=6 2>&1 | python -c "$FMT"
timeout 200 python bench.py --n 32 --grid voronoi --levels 1 --steps 20 --warmup 4 > gpurun_out/r2n_config0_voronoi32.json 2> gpurun_out/r2n_config0.err
cut -c1-250 gpurun_out/r2n_config0_voronoi32.json
rm -f gpurun_out/r2n_config3_dirs.jsonl
for d in 1 16 32 64; do
timeout 200 python bench.py --dirs $d --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r2n_config3_dirs.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r2n_config3_dirs.jsonl'):
    r=json.loads(l); print("dirs: ms/step %.3f all-cells %.3f value %.3e" % (r["ms_per_step"], r["all_cells_sweep_ms"], r["value"]), r["all_cells_form"][:70])
PY
