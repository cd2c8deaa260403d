"""ctypes binding of include/subsweep_b200.h -- the Python twin of the Rust FFI crate in INTEGRATION.md.

Nothing in here computes: every call goes to libsubsweep_b200.so (CUDA, sm_100a).  If the
library is missing the import of :func:`load` raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libsubsweep_b200.so"

SSW_OK, SSW_E_INVALID, SSW_E_CUDA, SSW_E_DEADLOCK, SSW_E_NOMEM, SSW_E_COMM = 0, -1, -2, -3, -4, -5
FACE_LOCAL, FACE_BOUNDARY, FACE_LOCAL_PERIODIC = 0, 1, 2
FLAG_NO_SCHEDULE_CACHE, FLAG_NO_COMPILED_PATH, FLAG_NO_PATCH_PATH, FLAG_SHARED_DEVICE = 1, 2, 4, 8

FIELDS = {
    "ionized_hydrogen_fraction": 0, "temperature": 1, "timestep": 2, "photon_rate": 3,
    "change_timescale": 4, "photoionization_rate": 5, "heating_rate": 6, "recombination_rate": 7,
    "collisional_ionization_rate": 8, "previous_rate": 9, "density": 10, "source": 11,
    "ionization_time": 12,
}
STATS = {
    "tasks_solved": 0, "single_sweeps": 1, "chem_cells": 2, "chem_failures": 3, "schedule_builds": 4,
    "schedule_replays": 5, "kernel_launches": 6, "wavefront_levels": 7, "chem_attempts": 8,
    "chem_max_depth": 9, "patch_macro_tiles": 10, "patch_levels": 11, "patch_phases": 12,
    "walk_window": 13, "walk_near_permille": 14,
}

c_double_p = C.POINTER(C.c_double)


class Params(C.Structure):
    _fields_ = [
        ("n_dirs", C.c_int32), ("dirs_xyz", c_double_p), ("n_levels", C.c_int32),
        ("max_timestep_s", C.c_double), ("timestep_safety_factor", C.c_double),
        ("chemistry_timestep_safety_factor", C.c_double),
        ("significant_rate_threshold_per_s", C.c_double), ("prevent_cooling", C.c_int32),
        ("scale_factor", C.c_double), ("check_deadlock", C.c_int32), ("device_id", C.c_int32),
        ("rank", C.c_int32), ("world_size", C.c_int32), ("flags", C.c_uint32),
    ]


class Grid(C.Structure):
    _fields_ = [
        ("n_cells", C.c_uint64), ("face_offsets", C.POINTER(C.c_uint64)), ("face_area", c_double_p),
        ("face_normal", c_double_p), ("face_neighbour", C.POINTER(C.c_int32)),
        ("face_kind", C.POINTER(C.c_uint8)), ("cell_size", c_double_p), ("cell_volume", c_double_p),
    ]


class Timings(C.Structure):
    _fields_ = [
        ("sweep_ms", C.c_double), ("chemistry_ms", C.c_double), ("update_levels_ms", C.c_double),
        ("schedule_ms", C.c_double), ("allreduce_ms", C.c_double), ("sweep_kernel_ms", C.c_double),
        ("sweep_kernel_launches", C.c_uint64), ("sweep_kernel_tasks", C.c_uint64),
        ("sweep_level_ms", C.c_double * 32), ("step_ms", C.c_double), ("steps", C.c_uint64),
        ("kernel_level_ms", C.c_double * 32), ("kernel_level_tasks", C.c_uint64 * 32),
        ("kernel_level_launches", C.c_uint64 * 32),
    ]

    def as_dict(self):
        d = {}
        for k, t in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


class TimeSeries(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "hydrogen_ionization_mass_average", "hydrogen_ionization_volume_average", "temperature_mass_average",
        "temperature_volume_average", "photoionization_rate_volume_average",
        "weighted_photoionization_rate_volume_average", "total_mass", "total_volume")]


PEER_HANDLE_BYTES = 64
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)
COLLECTIVE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p)
COLL_REDUCE_SCATTER, COLL_ALL_GATHER = 1, 2

# every symbol include/subsweep_b200.h declares: name -> (restype, argtypes)
H = C.c_void_p
SYMBOLS = {
    "ssw_create": (C.c_int, [C.POINTER(Params), C.POINTER(Grid), c_double_p, c_double_p, c_double_p, c_double_p, C.POINTER(H)]),
    "ssw_destroy": (None, [H]),
    "ssw_set_allreduce": (C.c_int, [H, ALLREDUCE_FN, C.c_void_p]),
    "ssw_set_collectives": (C.c_int, [H, COLLECTIVE_FN, C.c_void_p]),
    "ssw_set_directions": (C.c_int, [H, c_double_p]),
    "ssw_peer_arena": (C.c_int, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "ssw_peer_export": (C.c_int, [H, C.c_void_p]),
    "ssw_peer_attach_ipc": (C.c_int, [H, C.c_void_p]),
    "ssw_peer_attach": (C.c_int, [H, C.POINTER(C.c_void_p)]),
    "ssw_set_cell_positions": (C.c_int, [H, c_double_p]),
    "ssw_patch_note": (C.c_char_p, [H]),
    "ssw_run_sweeps": (C.c_int, [H, c_double_p]),
    "ssw_set_inputs": (C.c_int, [H, c_double_p, c_double_p]),
    "ssw_read": (C.c_int, [H, C.c_int, c_double_p]),
    "ssw_read_begin": (C.c_int, [H, C.c_int, c_double_p]),
    "ssw_sync": (C.c_int, [H]),
    "ssw_time_series_compute": (C.c_int, [H, c_double_p, C.c_int32, C.POINTER(TimeSeries)]),
    "ssw_read_levels": (C.c_int, [H, C.POINTER(C.c_uint8)]),
    "ssw_read_chem_attempts": (C.c_int, [H, C.POINTER(C.c_uint16)]),
    "ssw_level_counts": (C.c_int, [H, C.POINTER(C.c_uint64)]),
    "ssw_lowest_allowed_level": (C.c_int, [H, C.POINTER(C.c_int32)]),
    "ssw_single_sweep": (C.c_int, [H, C.c_int32]),
    "ssw_set_levels": (C.c_int, [H, C.POINTER(C.c_uint8)]),
    "ssw_set_change_timescale": (C.c_int, [H, c_double_p]),
    "ssw_update_timestep_levels": (C.c_int, [H]),
    "ssw_read_dir_state": (C.c_int, [H, C.c_int32, c_double_p]),
    "ssw_read_wavefront_levels": (C.c_int, [H, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "ssw_get_stat": (C.c_int, [H, C.c_int, C.POINTER(C.c_uint64)]),
    "ssw_get_timings": (C.c_int, [H, C.POINTER(Timings)]),
    "ssw_reset_timings": (C.c_int, [H]),
    "ssw_set_timing_level": (C.c_int, [H, C.c_int32]),
    "ssw_direction_shard": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ssw_patch_lattice": (C.c_int32, [c_double_p, C.c_uint64, C.c_int32, C.POINTER(C.c_uint32)]),
    "ssw_direction_groups": (C.c_int32, [c_double_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "ssw_patch_levels": (C.c_int32, [C.POINTER(C.c_uint32), C.c_int32, C.c_int32, C.POINTER(C.c_uint32)]),
    "ssw_level_from_timesteps": (C.c_int32, [C.c_int32, C.c_double, C.c_double]),
    "ssw_levels_in_sweep_order": (C.c_int32, [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32]),
    "ssw_chemistry_batch": (C.c_int, [C.c_int32, C.c_uint64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                     c_double_p, c_double_p, C.c_double, C.c_double, C.c_int32, c_double_p,
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "ssw_last_error": (C.c_char_p, []),
    "ssw_abi_version": (C.c_int32, []),
}

_lib = None


class SubsweepError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libsubsweep_b200 error {code}: {message}")
        self.code = code


def load(path: str | Path | None = None) -> C.CDLL:
    """dlopen the CUDA library and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise FileNotFoundError(
            f"{p} is not built (run `python -m subsweep_b200.build`); there is no CPU fallback for the sweep")
    lib = C.CDLL(str(p))
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if path is None:
        _lib = lib
    return lib


def check(lib, rc: int) -> None:
    if rc != SSW_OK:
        raise SubsweepError(rc, (lib.ssw_last_error() or b"").decode())


def dptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_double_p)
