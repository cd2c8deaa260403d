cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_sweep_parity.py tests/test_gpu_paths.py -m gpu -q --timeout 300 2>&1 | tail -3
for e in "" "SSW_FUSED_SMALL=0"; do
env $e timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=r['timing']; k=r['steps']
print('N1 $e ms/step %.3f sweep %.3f chem %.3f lvl %s launches/step %.1f' % (r['ms_per_step'], t['sweep_ms']/k, t['chemistry_ms']/k, [round(x/k,3) for x in t['sweep_level_ms']], r['gpu_launches']/k), t['checksum'])"
env $e timeout 300 python bench.py --emulate-shard 8 --steps 20 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shard8 $e', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('ms_per_step','all_cells_sweep_ms','sweep_ms','chemistry_ms','sweep_level_ms')})"
done
