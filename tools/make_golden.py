#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ (committed, small).

The reference itself (Rust nightly + MPI + HDF5) can neither be built nor imported in this image
(SURVEY.md section 8c), so the vectors come from two sources:

* ``reference_known_answers.json`` -- every known answer the reference's OWN tests pin for this
  path, transcribed as data with the file:line they come from (level rule, sweep order, warm-up,
  the two production-like chemistry inputs that must terminate).
* ``sweep_*.npz`` / ``chemistry_cells.npz`` -- outputs of the CPU oracle (oracle/oracle.c, the
  restatement of the reference algorithm) on small seeded problems, stored together with the
  complete inputs (flat grid, fields, parameters) so the fixtures do not depend on Qhull or numpy
  random streams staying stable.  They pin the oracle against regressions and give the CUDA
  path a fixed target; they are NOT outputs of the Rust binary ("parity unpinned", oracle.h).

usage: python tools/make_golden.py        (rewrites tests/golden/)
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
OUT = ROOT / "tests" / "golden"

CASES = {
    # name: (grid kind, n, periodic, n_dirs, n_levels, steps)
    "sweep_cartesian_periodic": ("cartesian", 7, True, 21, 3, 6),
    "sweep_cartesian_open": ("cartesian", 6, False, 84, 1, 2),
    "sweep_voronoi_open": ("voronoi", 5, False, 84, 2, 5),
    "sweep_voronoi_periodic": ("voronoi", 5, True, 16, 4, 7),
}
FIELDS = ("ionized_hydrogen_fraction", "temperature", "timestep", "change_timescale", "previous_rate", "photon_rate")


def sweep_case(name, kind, n, periodic, n_dirs, n_levels, steps):
    import oracle
    from helpers import make_problem
    params, g, f = make_problem(kind, n, periodic, n_dirs=n_dirs, n_levels=n_levels, max_timestep_myr=0.25)
    mode = oracle.PERIODIC_LAGGED if periodic else oracle.PERIODIC_HEAP
    s = oracle.OracleSweep(params, g, **f, periodic_mode=mode)
    out = dict(
        face_offsets=g.face_offsets.astype(np.uint64), face_area=g.face_area, face_normal=g.face_normal,
        face_neighbour=g.face_neighbour.astype(np.int32), face_kind=g.face_kind.astype(np.uint8),
        cell_size=g.cell_size, cell_volume=g.cell_volume,
        density=f["density"], xhii0=f["ionized_hydrogen_fraction"], temperature0=f["temperature"], source=f["source"],
        params=np.array([n_dirs, n_levels, int(periodic), steps], dtype=np.int64),
        fparams=np.array([params.max_timestep, params.significant_rate_threshold, params.timestep_safety_factor,
                          params.chemistry_timestep_safety_factor, float(params.prevent_cooling)]),
        periodic_mode=np.array([mode]),
    )
    out["wavefront_levels_dir0"] = s.wavefront_levels(n_levels - 1, 0)   # initial state: every cell active
    elapsed = []
    for step in range(steps):
        elapsed.append(s.run_sweeps())
        if step == 0:
            for k in FIELDS:
                out["step1_" + k] = s.read(k)
            out["step1_levels"] = s.levels()
            out["step1_outgoing"] = s.dir_state("outgoing")
    for k in FIELDS:
        out["final_" + k] = s.read(k)
    out["final_levels"] = s.levels()
    out["final_level_counts"] = s.level_counts()
    out["final_outgoing"] = s.dir_state("outgoing")
    out["final_incoming"] = s.dir_state("incoming")
    out["final_periodic"] = s.dir_state("periodic")
    out["elapsed"] = np.array(elapsed)
    out["stats"] = np.array([s.stat("tasks_solved"), s.stat("single_sweeps"), s.stat("chem_cells")], dtype=np.uint64)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, "cells", g.n_cells, "faces", len(g.face_area), "bytes", (OUT / f"{name}.npz").stat().st_size)


def chemistry_case():
    import oracle
    from subsweep_b200 import units as U
    rng = np.random.default_rng(20261017)
    n = 768
    x = np.where(rng.random(n) < 0.3, 1e-10, rng.uniform(0, 1, n))
    x[rng.random(n) < 0.1] = 1.0 - 1e-10
    T = 10.0 ** rng.uniform(1, 7, n)
    rho = 10.0 ** rng.uniform(-6, 1, n) * U.PER_CUBIC_CENTIMETER * U.PROTON_MASS
    length = 10.0 ** rng.uniform(-1, 2, n) * U.KILOPARSEC
    vol = length ** 3 * rng.uniform(0.5, 4.0, n)
    rate = np.where(rng.random(n) < 0.3, 0.0, 10.0 ** rng.uniform(30, 56, n))
    dt = 10.0 ** rng.uniform(-3, 1, n) * U.MEGAYEARS
    out = dict(xhii=x, temperature=T, density=rho, volume=vol, length=length, rate=rate, timestep=dt)
    for pc in (0, 1):
        r = oracle.chemistry(x, T, rho, vol, length, rate, dt, scale_factor=0.5, safety=0.1, prevent_cooling=bool(pc))
        for k, v in r.items():
            out[f"pc{pc}_{k}"] = v
    np.savez_compressed(OUT / "chemistry_cells.npz", **out)
    print("chemistry_cells", n, "bytes", (OUT / "chemistry_cells.npz").stat().st_size)


REFERENCE_KNOWN_ANSWERS = {
    "_comment": "Known answers pinned by the reference's own unit tests for the sweep + chemistry path; "
                "paths relative to the reference repository root.",
    "compute_timestep_level": {
        "source": "src/sweep/timestep_level.rs:66-89",
        "max_timestep_s": 1.0,
        "cases": [[1, 1.0, 0], [2, 1.0, 0], [1, 0.001, 0], [2, 0.001, 1], [3, 0.001, 2], [2, 0.500001, 1],
                  [2, 0.499999, 1], [3, 0.499999, 2], [5, 100.0, 0], [5, 0.0, 4]],
        "columns": ["max_num_levels", "desired_timestep_s", "level"],
    },
    "iter_levels_in_sweep_order": {
        "source": "src/sweep/timestep_state.rs:116-136",
        "num_levels": 5,
        "by_lowest_allowed": {"4": [4], "3": [3, 4], "2": [2, 4, 3, 4], "1": [1, 4, 3, 4, 2, 4, 3, 4],
                              "0": [0, 4, 3, 4, 2, 4, 3, 4, 1, 4, 3, 4, 2, 4, 3, 4]},
    },
    "lowest_allowed_warm_up": {
        "source": "src/sweep/timestep_state.rs:98-114",
        "num_levels": 5,
        "lowest_allowed_before_each_of_7_calls": [4, 4, 3, 2, 1, 0, 0],
    },
    "chemistry_must_terminate": {
        "source": "src/chemistry/hydrogen_only/mod.rs:948-984",
        "temperature_K": 1791871.5383082589, "density_g_per_cm3": 1.5411844211187435e-26,
        "volume_m3": 8.873284571355481e60, "length_kpc": 6.709257125565072, "rate_per_s": 4.661030976656667e44,
        "scale_factor": 8.35028211377591, "timestep_Myr": 1.0, "timestep_safety_factor": 0.1,
        "ionized_hydrogen_fraction": [1.0, 0.0],
    },
    "constants": {
        "source": "src/units/mod.rs:21,48,61,101-108",
        "year_s": 3.15576e7, "parsec_m": 3.0857e16, "boltzmann": 1.380649e-23, "proton_mass_kg": 1.67262192369e-27,
        "gamma": 5.0 / 3.0, "sigma_cm2": 2.9580524545305314e-18, "photon_energy_eV": 18.028356312818811,
        "rydberg_eV": 13.65693, "eV_J": 1.602176634e-19,
    },
}


def main():
    import oracle
    oracle.build()
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "reference_known_answers.json").write_text(json.dumps(REFERENCE_KNOWN_ANSWERS, indent=1) + "\n")
    for name, spec in CASES.items():
        sweep_case(name, *spec)
    chemistry_case()


if __name__ == "__main__":
    main()
