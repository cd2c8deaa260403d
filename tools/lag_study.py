#!/usr/bin/env python
"""CPU study (oracle only, no GPU): how far do the two definitions of `periodic_source` reads drift apart on a realistic
periodic Voronoi box?  HEAP = the reference's single-rank BinaryHeap task order (src/sweep/task.rs:25-35), LAGGED = the
order-independent definition the CUDA path implements.

    python tools/lag_study.py [--n 32] [--myr 1 2 4] [--source 1e52 1e54]

Workload: bench.py's headline box (--grid cartesian) or its Voronoi variant (default) at n^3 cells (log-normal density, point sources at the density
peaks, 84 directions, 4 timestep levels).  Writes one JSON line per (source strength, time) to stdout."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench   # noqa: E402
import oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--grid", choices=("voronoi", "cartesian"), default="voronoi")
    ap.add_argument("--dirs", type=int, default=84)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--myr", type=float, nargs="+", default=[1.0, 2.0, 4.0])
    ap.add_argument("--source", type=float, nargs="+", default=[1e52, 1e54])
    args = ap.parse_args()
    params, g, f = bench.build_workload(args.n, args.grid, args.dirs, args.levels)
    for strength in args.source:
        fields = dict(f)
        fields["source"] = np.where(f["source"] > 0, strength, 0.0)
        a = oracle.OracleSweep(params, g, **fields, periodic_mode=oracle.PERIODIC_HEAP)
        b = oracle.OracleSweep(params, g, **fields, periodic_mode=oracle.PERIODIC_LAGGED)
        t_sim, t0 = 0.0, time.time()
        for target in sorted(args.myr):
            while t_sim < target * 3.15576e13 * (1 - 1e-12):
                ea, eb = a.run_sweeps(), b.run_sweeps()
                assert ea == eb
                t_sim += ea
            xa, xb = a.read("ionized_hydrogen_fraction"), b.read("ionized_hydrogen_fraction")
            ta, tb = a.read("temperature"), b.read("temperature")
            rel = np.abs(xa - xb) / np.maximum(xb, 1e-300)
            big = xb > 1e-3                      # cells that are noticeably ionized
            print(json.dumps({
                "grid": args.grid, "cells": int(g.n_cells), "directions": args.dirs, "levels": args.levels, "source_per_s": strength,
                "time_myr": t_sim / 3.15576e13, "tasks": int(a.stat("tasks_solved")),
                "nonlagged_periodic_reads": int(a.stat("nonlagged_periodic_reads")),
                "levels_equal": bool(np.array_equal(a.levels(), b.levels())),
                "mean_xhii_heap": float(xa.mean()), "mean_xhii_lagged": float(xb.mean()),
                "mean_xhii_rel_diff": float(abs(xa.mean() - xb.mean()) / xb.mean()),
                "xhii_max_abs_diff": float(np.abs(xa - xb).max()),
                "xhii_rel_diff_median": float(np.median(rel)), "xhii_rel_diff_p99": float(np.quantile(rel, 0.99)),
                "xhii_rel_diff_max": float(rel.max()),
                "cells_xhii_gt_1e-3": int(big.sum()),
                "xhii_rel_diff_max_where_gt_1e-3": float(rel[big].max()) if big.any() else None,
                "temperature_rel_diff_max": float((np.abs(ta - tb) / tb).max()),
                "wall_s": time.time() - t0}), flush=True)


if __name__ == "__main__":
    main()
