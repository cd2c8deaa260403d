#!/usr/bin/env python
"""GPU experiment driver: one workload, many kernel configurations (environment tunables of stream.cuh / patch.cuh).

    python tools/tune_stream.py --n 128 --grid voronoi --out gpurun_out/tune.jsonl CONFIG [CONFIG ...]

CONFIG is a comma-separated list of NAME=VALUE environment settings ("SSW_STREAM_SOLO=1,SSW_SOLO_TILE=256"); "-" is
the default configuration.  The grid is built once; every configuration gets a fresh Sweep, `--calls` run_sweeps calls
(the first ones unlock the timestep levels and warm up), and reports the all-cells sweep kernel time of the last
`--timed` calls.  Profiling aid, not a bench line."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--grid", default="voronoi")
    ap.add_argument("--dirs", type=int, default=84)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--workload", default="box")
    ap.add_argument("--calls", type=int, default=6)
    ap.add_argument("--timed", type=int, default=3)
    ap.add_argument("--shard", type=int, default=0, help="run rank 0's direction shard of a W-rank job (no-op all-reduce)")
    ap.add_argument("--out", default="gpurun_out/tune.jsonl")
    ap.add_argument("configs", nargs="+")
    args = ap.parse_args()
    from subsweep_b200 import Sweep, build
    build.build()
    t0 = time.time()
    params, g, f = bench.build_workload(args.n, args.grid, args.dirs, args.levels, args.workload)
    b_alg, f_up = bench.algorithmic_bytes_per_update(g, Sweep.__init__.__globals__["Directions"].from_spec(args.dirs).xyz)
    print(f"grid built in {time.time() - t0:.1f} s: {g.n_cells} cells, F_up {f_up:.2f}, B_alg {b_alg:.1f}", flush=True)
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6650.0) if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    base_env = {k: v for k, v in os.environ.items() if k.startswith("SSW_")}
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    for cfg in args.configs:
        for k in [k for k in os.environ if k.startswith("SSW_")]:
            del os.environ[k]
        os.environ.update(base_env)
        if cfg != "-":
            for kv in cfg.split(","):
                k, v = kv.split("=")
                os.environ[k] = v
        rec = {"config": cfg, "n": args.n, "grid": args.grid, "dirs": args.dirs, "shard": args.shard}
        try:
            kw = {}
            if args.shard:
                kw = dict(rank=0, world_size=args.shard, allreduce=lambda ptr, n, stream: None)
            t0 = time.time()
            s = Sweep(params, g, **f, **kw)
            for _ in range(args.calls - args.timed):
                s.run_sweeps()
            rec["setup_s"] = time.time() - t0
            s.reset_timings()
            for _ in range(args.timed):
                s.run_sweeps()
            tim = s.timings()
            lvl = int(np.argmax(tim["kernel_level_tasks"]))
            k_ms = tim["kernel_level_ms"][lvl] / max(1, tim["kernel_level_launches"][lvl])
            tasks = tim["kernel_level_tasks"][lvl] / max(1, tim["kernel_level_launches"][lvl])
            rec.update(all_cells_sweep_ms=k_ms, step_ms=tim["step_ms"] / args.timed, sweep_ms=tim["sweep_ms"] / args.timed,
                       chemistry_ms=tim["chemistry_ms"] / args.timed, schedule_ms=tim["schedule_ms"] / args.timed,
                       roofline_frac=(b_alg * tasks / (k_ms * 1e-3) / 1e9) / peak if k_ms > 0 else None,
                       wavefront_levels=s.stat("wavefront_levels"), macro_tiles=s.stat("patch_macro_tiles"),
                       patch_note=s.patch_note(), mean_xhii=float(s.read("ionized_hydrogen_fraction").mean()),
                       mean_T=float(s.read("temperature").mean()))
            s.close()
        except Exception as exc:   # noqa: BLE001 - an experiment that fails is a result
            rec["error"] = repr(exc)
        print(json.dumps(rec), flush=True)
        with open(args.out, "a") as fh:
            fh.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
