#!/usr/bin/env python
"""bench.py -- throughput of the sweep + chemistry hot path on B200 (driver contract).

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement on the host cores

A *step* is one ``Sweep::run_sweeps`` call (src/sweep/mod.rs:258-272): every single sweep of the
reference's level order, chemistry after each, and the timestep-level update.  The metric is
BASELINE.json's: cell-direction updates per second, where one update is one solved task (cell,
direction) and a step's updates are the sum over its single sweeps of (#active cells x D).

Workload (``config.workload``): BASELINE.json configs[1] -- synthetic 128^3-cell periodic box,
log-normal density, 64 point sources, 84 directions, 4 timestep levels (SURVEY.md section 8d,
config 2, Cartesian variant; ``--grid voronoi`` tiles a periodic Voronoi block instead).
For N > 1 the directions are sharded over the ranks (strong scaling: the total work is fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "cell_direction_updates_per_s"
UNIT = "updates/s"


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def build_front_workload(n: int, n_dirs: int, n_levels: int, cell_scale: float = 1.0):
    """chemistry-stiff ionization front.  n^3 Cartesian box, non-periodic,
    dense neutral slab x in [0.25, 0.75] (n_H = 1 cm^-3, elsewhere 1e-4), one 5e54 /s source at the centre of
    the -x face."""
    from subsweep_b200 import SweepParameters, grid as G
    from subsweep_b200 import units as U

    cell = cell_scale * 10.0 * U.MEGAPARSEC / 128.0   # cell_scale < 1: the same photons meet fewer atoms, the front crosses the slab
    g = G.cartesian((n, n, n), cell * n, periodic=False)
    ix = np.arange(g.n_cells) // (n * n)
    slab = (ix >= n // 4) & (ix < 3 * n // 4)
    rho = np.where(slab, 1.0, 1e-4) * U.PER_CUBIC_CENTIMETER * U.PROTON_MASS
    src = np.zeros(g.n_cells)
    src[(0 * n + n // 2) * n + n // 2] = 5e54
    fields = dict(density=rho, ionized_hydrogen_fraction=np.full(g.n_cells, 1e-10), temperature=np.full(g.n_cells, 100.0),
                  source=src)
    params = SweepParameters(directions=n_dirs, num_timestep_levels=n_levels, periodic=False,
                             max_timestep=1.0 * U.MEGAYEARS, significant_rate_threshold=1e-5,
                             timestep_safety_factor=0.1, chemistry_timestep_safety_factor=0.1, prevent_cooling=True)
    return params, g, fields


def build_workload(n: int, grid_kind: str, n_dirs: int, n_levels: int, workload: str = "box", front_scale: float = 1.0):
    """SURVEY.md section 8d config 2 at n^3 cells (cell size fixed at 10 Mpc / 128)."""
    from subsweep_b200 import SweepParameters, grid as G
    from subsweep_b200 import units as U

    if workload == "front":
        return build_front_workload(n, n_dirs, n_levels, front_scale)
    cell = 10.0 * U.MEGAPARSEC / 128.0
    box = cell * n
    if grid_kind == "cartesian":
        g = G.cartesian((n, n, n), box, periodic=True)
    elif grid_kind == "voronoi":
        unit_n = 16 if n % 16 == 0 else n
        rng = np.random.default_rng(1338)
        ijk = np.stack(np.meshgrid(*(np.arange(unit_n),) * 3, indexing="ij"), axis=-1).reshape(-1, 3)
        pts = (ijk + 0.5 + 0.35 * rng.uniform(-1, 1, size=ijk.shape)) * cell
        unit = G.voronoi(pts, cell * unit_n, periodic=True)
        reps = n // unit_n
        g = G.tile_periodic(unit, (reps, reps, reps)) if reps > 1 else unit
    else:
        raise ValueError(grid_kind)
    N = g.n_cells
    mean_rho = 1e-3 * U.PER_CUBIC_CENTIMETER * U.PROTON_MASS
    rho = G.lognormal_density((n, n, n), mean_rho, sigma_g=1.0, smooth_cells=4.0, seed=2024)
    if grid_kind == "voronoi":
        # sample the field at the generator positions
        idx = np.clip((g.positions / cell).astype(np.int64), 0, n - 1)
        rho = rho.reshape(n, n, n)[idx[:, 0], idx[:, 1], idx[:, 2]]
    n_src = max(1, int(round(64 * (n / 128.0) ** 3)))
    src = np.zeros(N)
    src[np.argsort(rho)[-n_src:]] = 1e52
    fields = dict(density=np.ascontiguousarray(rho), ionized_hydrogen_fraction=np.full(N, 1e-10),
                  temperature=np.full(N, 100.0), source=src)
    params = SweepParameters(directions=n_dirs, num_timestep_levels=n_levels, periodic=True,
                             max_timestep=1.0 * U.MEGAYEARS, significant_rate_threshold=1e-5,
                             timestep_safety_factor=0.1, chemistry_timestep_safety_factor=0.1,
                             prevent_cooling=True)
    return params, g, fields


def algorithmic_bytes_per_update(g, dirs) -> tuple[float, float]:
    """B_alg = 20 * F_up + 24: per flux-carrying upwind face 8 B neighbour
    flux + 4 B index + 8 B geometric share; per task 8 B source/periodic + 8 B attenuation input +
    8 B outgoing flux write."""
    sample = dirs[:: max(1, len(dirs) // 12)]
    f_up = g.mean_upwind_faces(sample)
    return 20.0 * f_up + 24.0, f_up


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU DURING the timed region: NVML in a
    background thread every few ms (the timed region is tens of ms), nvidia-smi as fall-back."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, device_index: int, period_s: float = 0.004):
        self.device_index = device_index
        self.period_s = period_s
        self.samples: list[tuple[float, float, int]] = []
        self.sm_max = None
        self._stop = threading.Event()
        self.thread = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = device_index
            if visible:
                try:
                    index = int(visible.split(",")[device_index])
                except (ValueError, IndexError):
                    index = device_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:
            power = float("nan")
        try:
            reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            try:
                reasons = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                reasons = 0
        self.samples.append((sm, power, reasons))

    def _pump(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            self._stop.wait(self.period_s)

    def start(self):
        if self.nvml is None or self.handle is None:
            return
        self._stop.clear()
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            self.thread = None
        if self.nvml is None or not self.samples:
            return self._smi_once()
        sm = [s[0] for s in self.samples]
        power = [s[1] for s in self.samples if s[1] == s[1]]
        bits = 0
        for s in self.samples:
            bits |= s[2]
        reasons = [name for name, bit in self.REASONS if bits & bit]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": self.sm_max, "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": reasons, "source": "nvml, sampled during the timed region"}

    def _smi_once(self) -> dict:
        query = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={query}", "--format=csv,noheader,nounits", "-i",
                                  str(self.device_index)], capture_output=True, text=True, timeout=10).stdout.strip()
            parts = [p.strip() for p in out.split(",")]
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            reasons = [n for n, v in zip(names, parts[3:7]) if v.lower().startswith("active")]
            return {"sm_mhz": float(parts[0]), "sm_max_mhz": float(parts[1]), "power_w_max": float(parts[2]), "samples": 1,
                    "reasons": reasons, "source": "nvidia-smi, one sample right after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}


# ---------------------------------------------------------------------------------------------
# CPU arms (the oracle; the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_run(n: int, grid_kind: str, n_dirs: int, n_levels: int, steps: int, warmup: int, threads: int, workload: str = "box",
            front_scale: float = 1.0, spin_up: bool = True):
    """The CPU restatement of the reference algorithm (oracle/, kind "port"): the workload at n^3 cells on `threads`
    host threads (direction shards; one thread = the reference's single-rank task queue).  Returns (updates/s, ms per
    step, sample text, updates)."""
    import oracle
    params, g, f = build_workload(n, grid_kind, n_dirs, n_levels, workload, front_scale)
    s = oracle.OracleSweep(params, g, **f, periodic_mode=oracle.PERIODIC_LAGGED)
    for _ in range(n_levels if spin_up else 0):   # spin-up: unlock all timestep levels (same as the GPU arm)
        s.run_sweeps_threads(threads)
    for _ in range(warmup):
        s.run_sweeps_threads(threads)
    t0_tasks = s.stat("tasks_solved")
    t0 = time.perf_counter()
    for _ in range(steps):
        s.run_sweeps_threads(threads)
    dt = time.perf_counter() - t0
    tasks = s.stat("tasks_solved") - t0_tasks
    sample = (f"{n}^3-cell box of the same workload (cell size, density field statistics, source density, "
              f"{n_dirs} directions, {n_levels} levels), {warmup} warm-up + {steps} timed run_sweeps calls on {threads} thread(s)"
              + ("" if spin_up else ", no level spin-up: the first calls, one all-cells sweep each"))
    return tasks / dt, dt / steps * 1e3, sample, tasks


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    # the same configuration as the GPU arm (args.n cells per dimension) unless --cpu-n bounds the sample further
    n = args.cpu_n or args.n
    value, ms, sample, _ = cpu_run(n, args.grid, args.dirs, args.levels, args.steps, args.warmup, threads, args.workload,
                                   args.front_scale)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, note=None if n == args.n else "CPU arm runs a bounded sample: " + sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, note: str | None = None) -> dict:
    if args.workload == "front":
        text = (f"{args.n}^3-cell Cartesian box, non-periodic, dense neutral slab (n_H = 1 cm^-3) with a 5e54/s source at the "
                f"-x face centre, {args.dirs} directions, {args.levels} timestep levels, max_timestep 1 Myr (BASELINE.json configs[3])")
    else:
        text = (f"{args.n}^3-cell periodic box ({args.grid}), log-normal density, "
                f"{max(1, int(round(64 * (args.n / 128.0) ** 3)))} point sources of 1e52/s, {args.dirs} directions, "
                f"{args.levels} timestep levels, max_timestep 1 Myr (BASELINE.json configs[1])")
    cfg = {
        "workload": text,
        "cells": args.n ** 3, "directions": args.dirs, "timestep_levels": args.levels, "grid": args.grid,
        "parallelism": (f"direction sharding x{args.gpus}, " + ("NCCL hooks (SSW_BENCH_HOOKS)" if os.environ.get("SSW_BENCH_HOOKS")
                                                                  else "peer-mapped exchange over NVLink") if args.gpus > 1 else "single GPU"),
        "l2": "working set (per-direction flux state + level sets, > 2 GB) is far larger than the 126 MB L2; no flush needed",
        "step": "one Sweep::run_sweeps call (all single sweeps of the level order + chemistry + level update)",
    }
    if note:
        cfg["note"] = note
    return cfg


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def pinned(n: int) -> np.ndarray:
    import torch
    return torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()


def attempts_histogram(sweep) -> dict:
    """Substep attempts of every cell's last chemistry update, in octave bins."""
    a = sweep.chem_attempts().astype(np.int64)
    edges = [0, 1, 2, 3, 5, 9, 17, 33, 65, 129, 257, 513, 1025, 65536]
    counts, _ = np.histogram(a, bins=edges)
    labels = ["0 (never updated)", "1", "2", "3-4", "5-8", "9-16", "17-32", "33-64", "65-128", "129-256", "257-512", "513-1024", ">1024"]
    return {"bins": labels, "cells": [int(c) for c in counts], "max": int(a.max())}


def all_cells_form(sweep) -> str:
    """Which compiled form the all-cells sweep runs in and the kernel that carries it."""
    if sweep.stat("patch_macro_tiles"):
        return "patch_sweep_kernel (macro-tile dataflow, %d macro-tiles in %d dependent levels)" % (
            sweep.stat("patch_macro_tiles"), sweep.stat("patch_levels"))
    why = sweep.patch_note() or "patch form off"
    if sweep.stat("walk_window"):
        return "walk_kernel (one block per direction, %d-slot shared-memory window, %.1f %% of the upwind entries read from it; %s)" % (
            sweep.stat("walk_window"), sweep.stat("walk_near_permille") / 10.0, why)
    return "sweep_stream_kernel (level-barrier stream; %s)" % why


def run_b200(args) -> None:
    import torch
    import torch.distributed as dist
    from subsweep_b200 import Sweep, build as libbuild
    from subsweep_b200.distributed import attach_peers, init_from_env, make_allreduce, make_collectives

    rank, world, local_rank = init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if rank == 0:
        libbuild.build()
    if world > 1:
        dist.barrier()

    params, g, fields = build_workload(args.n, args.grid, args.dirs, args.levels, args.workload, args.front_scale)
    # N > 1: the ranks exchange through peer-mapped arenas over NVLink (csrc/peer.cuh); torch.distributed only carries
    # the 64-byte IPC handles and the bench's own barriers.  SSW_BENCH_HOOKS=1 runs the older NCCL-hook path instead.
    use_hooks = world > 1 and bool(os.environ.get("SSW_BENCH_HOOKS"))
    allreduce = make_allreduce(device) if use_hooks else None
    collectives = make_collectives(device) if use_hooks and not os.environ.get("SSW_BENCH_REPLICATED_CHEMISTRY") else None
    shard_rank, shard_world = rank, world
    if args.emulate_shard and world == 1:
        # profiling aid: one GPU runs rank 0's direction shard of a W-rank job with a no-op all-reduce, to tune the
        # kernels at the per-GPU work of a multi-GPU run without holding W GPUs (results are NOT a bench line)
        shard_rank, shard_world = 0, args.emulate_shard
        allreduce = lambda ptr, n, stream: None   # noqa: E731
        args.no_e2e = True
    sweep = Sweep(params, g, **fields, device_id=local_rank, rank=shard_rank, world_size=shard_world, allreduce=allreduce,
                  collectives=collectives)
    if world > 1 and not use_hooks:
        attach_peers(sweep)
    N = g.n_cells
    b_alg, f_up = algorithmic_bytes_per_update(g, sweep.directions.xyz)

    def sync_all():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # spin-up (workload set-up, not warm-up): the reference unlocks one timestep level per call
    # (timestep_state.rs:37-48), so the first n_levels calls are partial steps; the benchmark
    # measures full steady-state steps ("1 Myr + steady-state steps")
    # the timed region runs at timing level 0 (production: four CUDA event records per step -- the step and the
    # all-cells sweep kernel, which the roofline needs); the per-phase breakdown is measured in a separate pass below
    sweep.set_timing_level(0)
    for _ in range(args.levels):
        sweep.run_sweeps()
    for _ in range(args.warmup):
        sweep.run_sweeps()

    # ---- timed region: K steps, inputs resident in HBM -------------------------------------------
    sampler = ClockSampler(local_rank)
    sync_all()
    sweep.reset_timings()
    tasks0, launches0 = sweep.stat("tasks_solved"), sweep.stat("kernel_launches")
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sweep.run_sweeps()
    sync_all()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    tim = sweep.timings()
    my_tasks = sweep.stat("tasks_solved") - tasks0
    launches = sweep.stat("kernel_launches") - launches0
    # device time of the K steps: CUDA events on the library's stream (ssw_timings.step_ms)
    step_ms = torch.tensor([tim["step_ms"], wall * 1e3], dtype=torch.float64, device=device)
    tot = torch.tensor([float(my_tasks), float(launches)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms = (float(v) for v in step_ms.tolist())
    total_tasks, total_launches = (float(v) for v in tot.tolist())
    value = total_tasks / (dev_ms * 1e-3)

    # ---- per-phase breakdown (diagnostics, NOT the timed region): a few more steps at timing level 1 -------
    bsteps = max(1, min(args.steps, 8))
    sweep.set_timing_level(1)
    sweep.reset_timings()
    for _ in range(bsteps):
        sweep.run_sweeps()
    sync_all()
    tbd = sweep.timings()
    sweep.set_timing_level(0)
    breakdown = {
        "steps": bsteps,
        "note": "separate pass after the timed region with the library's per-phase timers on (ssw_set_timing_level 1); per step",
        "ms_per_step": tbd["step_ms"] / bsteps, "sweep_ms": tbd["sweep_ms"] / bsteps, "chemistry_ms": tbd["chemistry_ms"] / bsteps,
        "update_levels_ms": tbd["update_levels_ms"] / bsteps, "schedule_ms": tbd["schedule_ms"] / bsteps,
        "exchange_wait_ms": tbd["allreduce_ms"] / bsteps, "sweep_kernel_ms": tbd["sweep_kernel_ms"] / bsteps,
        "sweep_level_ms": [v / bsteps for v in tbd["sweep_level_ms"][:args.levels]],
        "kernel_level_ms": [v / bsteps for v in tbd["kernel_level_ms"][:args.levels]],
    }

    # ---- end to end: host buffers in, host buffers out, every step --------------------------------
    if args.no_e2e:
        if rank == 0:
            lvl0 = int(np.argmax(tim["kernel_level_tasks"]))
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": dev_ms / args.steps,
                              "all_cells_sweep_ms": tim["kernel_level_ms"][lvl0] / max(1, tim["kernel_level_launches"][lvl0]),
                              "sweep_ms": breakdown["sweep_ms"], "chemistry_ms": breakdown["chemistry_ms"],
                              "schedule_ms": breakdown["schedule_ms"], "update_levels_ms": breakdown["update_levels_ms"],
                              "allreduce_ms": breakdown["exchange_wait_ms"], "sweep_level_ms": breakdown["sweep_level_ms"],
                              "ms_per_step_with_phase_timers": breakdown["ms_per_step"],
                              "level_counts": [int(v) for v in sweep.level_counts()],
                              "chem_attempts": sweep.stat("chem_attempts"), "chem_cells": sweep.stat("chem_cells"),
                              "chem_max_depth": sweep.stat("chem_max_depth"), "schedule_builds": sweep.stat("schedule_builds"),
                              "all_cells_form": all_cells_form(sweep),
                              "chem_attempts_histogram": attempts_histogram(sweep),
                              "macro_tiles": sweep.stat("patch_macro_tiles"), "patch_levels": sweep.stat("patch_levels"),
                              "patch_phases": sweep.stat("patch_phases"),
                              "mean_xhii": float(sweep.read("ionized_hydrogen_fraction").mean()),
                              "env": {k: v for k, v in os.environ.items() if k.startswith("SSW_")},
                              "note": "profiling run (--no-e2e): not a bench line"}), flush=True)
        return
    src_host = pinned(N)
    src_host[:] = fields["source"]
    outs = {k: pinned(N) for k in ("ionized_hydrogen_fraction", "temperature", "timestep", "photon_rate", "ionization_time")}
    phase = {"set_inputs": 0.0, "run_sweeps": 0.0, "write_back": 0.0}
    e_tasks0, t0 = 0, 0.0
    # one untimed pass first: the photon_rate read-back of a sharded job is the first N-length all-reduce NCCL sees
    # (lazy channel set-up, tens of ms at 8 ranks)
    for it in range(args.steps + 1):
        if it == 1:
            sync_all()
            phase = {k: 0.0 for k in phase}
            e_tasks0 = sweep.stat("tasks_solved")
            t0 = time.perf_counter()
        ta = time.perf_counter()
        sweep.set_inputs(source=src_host)        # the Source component of this step (H2D), on every rank
        tb = time.perf_counter()
        sweep.run_sweeps()
        tc = time.perf_counter()
        # run_sweep_system write-back (D2H), mod.rs:718-738: five copies queued, one synchronisation.  All ranks hold the
        # same cell state; the rank that owns the output (0) reads it, the others only join the photon_rate all-reduce
        for k, buf in outs.items():
            if rank == 0:
                sweep.read_begin(k, buf)
            elif k == "photon_rate":
                sweep.read_begin(k, None)
        sweep.sync()
        td = time.perf_counter()
        phase["set_inputs"] += tb - ta
        phase["run_sweeps"] += tc - tb
        phase["write_back"] += td - tc
    sync_all()
    e_wall = time.perf_counter() - t0
    e = torch.tensor([e_wall], dtype=torch.float64, device=device)
    et = torch.tensor([float(sweep.stat("tasks_solved") - e_tasks0)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        dist.all_reduce(et, op=dist.ReduceOp.SUM)
    e2e_value = float(et.item()) / float(e.item())

    if rank != 0:
        return

    # ---- roofline of the dominant kernel: the all-cells sweep (current level 0) --------------------
    peaks = {}
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peaks = json.loads(peaks_file.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    lvl = int(np.argmax(tim["kernel_level_tasks"]))
    k_ms = tim["kernel_level_ms"][lvl]
    k_tasks = tim["kernel_level_tasks"][lvl]
    k_launches = max(1, tim["kernel_level_launches"][lvl])
    achieved = (b_alg * k_tasks / (k_ms * 1e-3)) / 1e9 if k_ms > 0 else 0.0
    # measured DRAM traffic of the same kernel on the same workload, from the committed ncu --set full capture
    traffic, traffic_source = None, None
    tfile = ROOT / "profiles" / "roofline_traffic.json"
    if tfile.exists() and world == 1 and (args.n, args.grid, args.dirs, args.workload) == (128, "cartesian", 84, "box") \
            and sweep.stat("patch_macro_tiles") > 0:
        t = json.loads(tfile.read_text())
        traffic, traffic_source = t["dram_bytes_per_launch"], t["source"]
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_nominal": 8000.0, "frac_nominal": achieved / 8000.0, "traffic": traffic,
        "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_source,
        "algorithmic_bytes_per_launch": b_alg * k_tasks / k_launches,
        "kernel": all_cells_form(sweep) + " of the all-cells single sweep (timestep level %d)" % lvl,
        "algorithmic_bytes_per_update": b_alg, "mean_upwind_faces": f_up,
        "updates_per_launch": k_tasks / k_launches, "ms_per_launch": k_ms / k_launches, "peak_source": peak_kind,
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        # bounded sample (about 30 s of CPU work): the same configuration, one warm-up and two timed steps on all host
        # threads; plus the single-core number (one thread = the reference on one rank) on a 64^3 box of the workload
        v, _, sample, _ = cpu_run(args.cpu_n or args.n, args.grid, args.dirs, args.levels, 2, 1, threads, args.workload,
                                  args.front_scale)
        v1, _, sample1, _ = cpu_run(min(64, args.n), args.grid, args.dirs, args.levels, 2, 0, 1, args.workload,
                                    args.front_scale, spin_up=False)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                        "single_core": {"value": v1, "unit": UNIT, "cores": 1, "sample": sample1}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * N * world,
                "d2h_bytes_per_step": 8 * N * len(outs),
                "rank0_wall_ms_per_step": {k: 1e3 * v / args.steps for k, v in phase.items()},
                "note": "every rank uploads the step's Source component; rank 0 reads back the five result components "
                        "(a cell's state lives on its owner rank and is pulled at read-back; the other ranks only join the photon_rate exchange)"},
        "gpu_launches": int(total_launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "timing": {"device_ms_total": dev_ms, "wall_ms_total": wall_ms, "updates_total": total_tasks,
                   "all_cells_sweep_kernel_ms_total": tim["kernel_level_ms"][lvl],
                   "all_cells_sweep_launches": tim["kernel_level_launches"][lvl],
                   "breakdown": breakdown,
                   "level_counts": [int(v) for v in sweep.level_counts()],
                   "chem_attempts": sweep.stat("chem_attempts"), "chem_cells": sweep.stat("chem_cells"),
                   "chem_max_depth": sweep.stat("chem_max_depth"), "wavefront_levels": sweep.stat("wavefront_levels"),
                   "chem_attempts_histogram": attempts_histogram(sweep),
                   "schedule_builds": sweep.stat("schedule_builds"), "schedule_replays": sweep.stat("schedule_replays"),
                   "patch_macro_tiles": sweep.stat("patch_macro_tiles"), "patch_levels": sweep.stat("patch_levels"),
                   "patch_note": sweep.patch_note(),
                   "checksum": {"mean_xhii": float(outs["ionized_hydrogen_fraction"].mean()),
                                "mean_temperature": float(outs["temperature"].mean())}},
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--n", "--cells-per-dim", dest="n", type=int, default=128,
                    help="cells per dimension (use the long form under torchrun, whose own parser chokes on --n)")
    ap.add_argument("--grid", choices=("cartesian", "voronoi"), default="cartesian")
    ap.add_argument("--dirs", type=int, default=84)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--workload", choices=("box", "front"), default="box",
                    help="box: BASELINE.json configs[1] (default); front: configs[3], the chemistry-stiff ionization front")
    ap.add_argument("--front-scale", type=float, default=1.0,
                    help="front workload: cell size factor (1 = SURVEY.md config 4; 0.03: the front crosses into the slab)")
    ap.add_argument("--cpu-n", type=int, default=0,
                    help="cells per dimension of the CPU arm (default 0: the same configuration as the GPU arm, --n)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emulate-shard", type=int, default=0, metavar="W",
                    help="profiling only: run rank 0's direction shard of a W-rank job on one GPU (no-op all-reduce)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: stop after the device-timed region")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3   # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
