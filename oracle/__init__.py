"""ctypes binding of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/oracle.h.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.  The product
(subsweep_b200/) never does.  PARITY UNPINNED: the reference cannot be built here and pins no
fluxes / xHII / T itself; see oracle.h.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"

PERIODIC_HEAP, PERIODIC_LAGGED = 0, 1

FIELDS = {
    "ionized_hydrogen_fraction": 0, "temperature": 1, "timestep": 2, "photon_rate": 3,
    "change_timescale": 4, "photoionization_rate": 5, "heating_rate": 6, "recombination_rate": 7,
    "collisional_ionization_rate": 8, "previous_rate": 9, "density": 10, "source": 11,
}
STATS = {"tasks_solved": 0, "nonlagged_periodic_reads": 1, "chem_attempts": 2, "chem_max_depth": 3,
         "chem_failures": 4, "single_sweeps": 5, "chem_cells": 6}
FITS = {"alpha_b": 0, "dalpha_b": 1, "recomb_cool": 2, "drecomb_cool": 3, "coll_ion": 4, "dcoll_ion": 5,
        "coll_ion_cool": 6, "dcoll_ion_cool": 7, "coll_exc_cool": 8, "dcoll_exc_cool": 9, "brems": 10,
        "dbrems": 11, "compton": 12, "dcompton": 13, "cooling": 14, "dcooling": 15}
CONSTS = {"proton_mass": 0, "boltzmann": 1, "gamma": 2, "sigma": 3, "photon_energy": 4, "rydberg": 5,
          "year": 6, "megayear": 7, "parsec": 8, "kiloparsec": 9}

dp = C.POINTER(C.c_double)


class Params(C.Structure):
    _fields_ = [
        ("n_dirs", C.c_int32), ("dirs_xyz", dp), ("n_levels", C.c_int32), ("max_timestep", C.c_double),
        ("timestep_safety_factor", C.c_double), ("chemistry_timestep_safety_factor", C.c_double),
        ("significant_rate_threshold", C.c_double), ("prevent_cooling", C.c_int32),
        ("scale_factor", C.c_double), ("check_deadlock", C.c_int32), ("periodic_mode", C.c_int32),
        ("dir_begin", C.c_int32), ("dir_end", C.c_int32),
    ]


class Grid(C.Structure):
    _fields_ = [
        ("n_cells", C.c_uint64), ("face_offsets", C.POINTER(C.c_uint64)), ("face_area", dp),
        ("face_normal", dp), ("face_neighbour", C.POINTER(C.c_int32)), ("face_kind", C.POINTER(C.c_uint8)),
        ("cell_size", dp), ("cell_volume", dp),
    ]


class Solver(C.Structure):
    _fields_ = [
        ("xhii", C.c_double), ("temperature", C.c_double), ("density", C.c_double), ("volume", C.c_double),
        ("length", C.c_double), ("rate", C.c_double), ("scale_factor", C.c_double), ("has_floor", C.c_int32),
        ("floor_temperature", C.c_double), ("floor_xhii", C.c_double),
    ]


class ChemResult(C.Structure):
    _fields_ = [("timescale", C.c_double), ("process", C.c_int32), ("failed", C.c_int32),
                ("attempts", C.c_uint64), ("max_depth", C.c_int32)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)

_lib = None


def build(force: bool = False) -> Path:
    src = [HERE / "oracle.c", HERE / "oracle.h", HERE / "Makefile"]
    if force or not LIB_PATH.exists() or any(p.stat().st_mtime > LIB_PATH.stat().st_mtime for p in src):
        res = subprocess.run(["make", "-C", str(HERE), "-B", "liboracle.so"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    lib = C.CDLL(str(LIB_PATH))
    H = C.c_void_p
    lib.orc_create.restype = H
    lib.orc_create.argtypes = [C.POINTER(Params), C.POINTER(Grid), dp, dp, dp, dp]
    lib.orc_destroy.argtypes = [H]
    lib.orc_destroy.restype = None
    lib.orc_set_allreduce.argtypes = [H, ALLREDUCE_FN, C.c_void_p]
    lib.orc_set_allreduce.restype = None
    lib.orc_run_sweeps.restype = C.c_double
    lib.orc_run_sweeps.argtypes = [H]
    lib.orc_run_sweeps_threads.restype = C.c_double
    lib.orc_run_sweeps_threads.argtypes = [H, C.c_int]
    lib.orc_single_sweep.argtypes = [H, C.c_int]
    lib.orc_single_sweep.restype = None
    lib.orc_set_levels.argtypes = [H, C.POINTER(C.c_uint8)]
    lib.orc_set_levels.restype = None
    lib.orc_set_change_timescale.argtypes = [H, dp]
    lib.orc_set_change_timescale.restype = None
    lib.orc_update_timestep_levels.argtypes = [H]
    lib.orc_update_timestep_levels.restype = None
    lib.orc_wavefront_levels.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    lib.orc_wavefront_levels.restype = None
    lib.orc_read.argtypes = [H, C.c_int, dp]
    lib.orc_read.restype = C.c_int
    lib.orc_read_levels.argtypes = [H, C.POINTER(C.c_uint8)]
    lib.orc_read_levels.restype = None
    lib.orc_level_counts.argtypes = [H, C.POINTER(C.c_uint64)]
    lib.orc_level_counts.restype = None
    lib.orc_read_dir_state.argtypes = [H, C.c_int, dp]
    lib.orc_read_dir_state.restype = None
    lib.orc_lowest_allowed_level.argtypes = [H]
    lib.orc_lowest_allowed_level.restype = C.c_int
    lib.orc_stat.argtypes = [H, C.c_int]
    lib.orc_stat.restype = C.c_uint64
    lib.orc_perform_timestep.argtypes = [C.POINTER(Solver), C.c_double, C.c_double, C.POINTER(ChemResult)]
    lib.orc_perform_timestep.restype = None
    lib.orc_fit.argtypes = [C.POINTER(Solver), C.c_int]
    lib.orc_fit.restype = C.c_double
    lib.orc_photoheating_rate.argtypes = [C.POINTER(Solver), C.c_double]
    lib.orc_photoheating_rate.restype = C.c_double
    lib.orc_photoionization_rate.argtypes = [C.POINTER(Solver), C.c_double]
    lib.orc_photoionization_rate.restype = C.c_double
    lib.orc_level_from_timesteps.argtypes = [C.c_int, C.c_double, C.c_double]
    lib.orc_level_from_timesteps.restype = C.c_int
    lib.orc_levels_in_sweep_order.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    lib.orc_levels_in_sweep_order.restype = C.c_int
    lib.orc_heap_pop_order.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]
    lib.orc_heap_pop_order.restype = None
    lib.orc_const.argtypes = [C.c_int]
    lib.orc_const.restype = C.c_double
    _lib = lib
    return lib


def const(name: str) -> float:
    return load().orc_const(CONSTS[name])


def _d(a):
    return a.ctypes.data_as(dp)


class OracleSweep:
    """``Sweep<HydrogenOnly>`` on the CPU; same surface as subsweep_b200.sweep.Sweep."""

    def __init__(self, parameters, grid, density, ionized_hydrogen_fraction, temperature, source,
                 scale_factor: float = 1.0, periodic_mode: int = PERIODIC_LAGGED, rank: int = 0,
                 world_size: int = 1, allreduce=None, directions=None):
        from subsweep_b200.sweep import Directions, direction_shard  # data + pure helper only
        self.lib = load()
        self.parameters = parameters
        self.directions = directions if directions is not None else Directions.from_spec(parameters.directions)
        D = len(self.directions)
        self.dir_begin, self.dir_end = direction_shard(D, world_size, rank)
        N = grid.n_cells
        self.n_cells = N
        p = Params()
        p.n_dirs = D
        self._dirs = np.ascontiguousarray(self.directions.xyz, dtype=np.float64)
        p.dirs_xyz = _d(self._dirs)
        p.n_levels = parameters.num_timestep_levels
        p.max_timestep = parameters.max_timestep
        p.timestep_safety_factor = parameters.timestep_safety_factor
        p.chemistry_timestep_safety_factor = parameters.chemistry_timestep_safety_factor
        p.significant_rate_threshold = parameters.significant_rate_threshold
        p.prevent_cooling = int(parameters.prevent_cooling)
        p.scale_factor = scale_factor
        p.check_deadlock = int(parameters.check_deadlock)
        p.periodic_mode = periodic_mode
        p.dir_begin, p.dir_end = (self.dir_begin, self.dir_end) if world_size > 1 else (0, 0)
        g = Grid()
        g.n_cells = N
        keep = (np.ascontiguousarray(grid.face_offsets, dtype=np.uint64),
                np.ascontiguousarray(grid.face_area, dtype=np.float64),
                np.ascontiguousarray(grid.face_normal, dtype=np.float64),
                np.ascontiguousarray(grid.face_neighbour, dtype=np.int32),
                np.ascontiguousarray(grid.face_kind, dtype=np.uint8),
                np.ascontiguousarray(grid.cell_size, dtype=np.float64),
                np.ascontiguousarray(grid.cell_volume, dtype=np.float64))
        g.face_offsets = keep[0].ctypes.data_as(C.POINTER(C.c_uint64))
        g.face_area = _d(keep[1])
        g.face_normal = _d(keep[2])
        g.face_neighbour = keep[3].ctypes.data_as(C.POINTER(C.c_int32))
        g.face_kind = keep[4].ctypes.data_as(C.POINTER(C.c_uint8))
        g.cell_size = _d(keep[5])
        g.cell_volume = _d(keep[6])
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (density, ionized_hydrogen_fraction, temperature, source)]
        self._h = self.lib.orc_create(C.byref(p), C.byref(g), *(_d(a) for a in arrs))
        self._cb = None
        if world_size > 1:
            if allreduce is None:
                raise ValueError("world_size > 1 needs an allreduce callable")

            def trampoline(_ctx, buf, n):
                try:
                    allreduce(int(buf), int(n), None)
                    return 0
                except Exception as exc:
                    import sys
                    print(f"oracle: allreduce hook failed: {exc!r}", file=sys.stderr)
                    return -1
            self._cb = ALLREDUCE_FN(trampoline)
            self.lib.orc_set_allreduce(self._h, self._cb, None)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_sweeps(self) -> float:
        return self.lib.orc_run_sweeps(self._h)

    def run_sweeps_threads(self, n_threads: int) -> float:
        return self.lib.orc_run_sweeps_threads(self._h, n_threads)

    def single_sweep(self, level: int) -> None:
        self.lib.orc_single_sweep(self._h, level)

    def update_timestep_levels(self) -> None:
        self.lib.orc_update_timestep_levels(self._h)

    def read(self, name: str) -> np.ndarray:
        out = np.empty(self.n_cells)
        rc = self.lib.orc_read(self._h, FIELDS[name], _d(out))
        assert rc == 0
        return out

    def levels(self) -> np.ndarray:
        out = np.empty(self.n_cells, dtype=np.uint8)
        self.lib.orc_read_levels(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def level_counts(self) -> np.ndarray:
        out = np.zeros(self.parameters.num_timestep_levels, dtype=np.uint64)
        self.lib.orc_level_counts(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out

    def lowest_allowed_level(self) -> int:
        return self.lib.orc_lowest_allowed_level(self._h)

    def set_levels(self, levels) -> None:
        lv = np.ascontiguousarray(levels, dtype=np.uint8)
        self.lib.orc_set_levels(self._h, lv.ctypes.data_as(C.POINTER(C.c_uint8)))

    def set_change_timescale(self, tau) -> None:
        t = np.ascontiguousarray(tau, dtype=np.float64)
        self.lib.orc_set_change_timescale(self._h, _d(t))

    def dir_state(self, which: str) -> np.ndarray:
        idx = {"incoming": 0, "outgoing": 1, "periodic": 2}[which]
        out = np.empty((self.n_cells, self.dir_end - self.dir_begin))
        self.lib.orc_read_dir_state(self._h, idx, _d(out))
        return out

    def wavefront_levels(self, level: int, direction: int) -> np.ndarray:
        out = np.empty(self.n_cells, dtype=np.int32)
        self.lib.orc_wavefront_levels(self._h, level, direction, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def stat(self, name: str) -> int:
        return int(self.lib.orc_stat(self._h, STATS[name]))


def chemistry(xhii, temperature, density, volume, length, rate, timestep, scale_factor=1.0, safety=0.1,
              prevent_cooling=False):
    """HydrogenOnly::update_abundances on arrays of independent cells.  Returns a dict of arrays."""
    lib = load()
    arrs = np.broadcast_arrays(*(np.asarray(a, dtype=np.float64) for a in
                                 (xhii, temperature, density, volume, length, rate, timestep)))
    n = arrs[0].size
    out = {k: np.empty(n) for k in ("xhii", "temperature", "timescale")}
    out["process"] = np.empty(n, dtype=np.int32)
    out["depth"] = np.empty(n, dtype=np.int32)
    out["attempts"] = np.empty(n, dtype=np.uint64)
    flat = [a.ravel() for a in arrs]
    for i in range(n):
        s = Solver(flat[0][i], flat[1][i], flat[2][i], flat[3][i], flat[4][i], flat[5][i], scale_factor,
                   int(prevent_cooling), flat[1][i], flat[0][i])
        r = ChemResult()
        lib.orc_perform_timestep(C.byref(s), flat[6][i], safety, C.byref(r))
        out["xhii"][i] = s.xhii
        out["temperature"][i] = s.temperature
        out["timescale"][i] = r.timescale
        out["process"][i] = -1 if r.failed else r.process
        out["depth"][i] = r.max_depth
        out["attempts"][i] = r.attempts
    return out
