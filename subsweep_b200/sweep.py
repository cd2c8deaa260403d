"""Host-side mirror of the reference's sweep plugin surface, on top of the C ABI.

Names follow the reference: ``SweepParameters`` (src/sweep/parameters.rs:8-46, the ``sweep:`` YAML
section), ``Directions`` (src/sweep/direction/mod.rs:36-109), ``Sweep`` with ``run_sweeps``
(src/sweep/mod.rs:172-272) and the ``init_sweep_system`` / ``run_sweep_system`` pair
(:634-739) operating on a dict of per-particle component arrays.

Everything numerical happens in libsubsweep_b200.so (CUDA); this file only marshals arrays.
"""
from __future__ import annotations

import ctypes as C
import json
import sys
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, Optional, Sequence, Union

import numpy as np

from . import capi
from .grid import FlatGrid
from .units import parse_quantity

_DATA = Path(__file__).resolve().parent / "data" / "direction_bins.json"


class Directions:
    """Direction bins: hard-coded tables for 1/16/21/32/64/84 (not re-normalised) or explicit,
    normalised lists (src/sweep/direction/mod.rs:58-109)."""

    def __init__(self, xyz: np.ndarray):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)

    @classmethod
    def from_num(cls, num: int) -> "Directions":
        tables = json.loads(_DATA.read_text())
        if str(num) not in tables:
            raise NotImplementedError(f"no direction table for {num} directions")  # unimplemented!() :67
        return cls(np.array(tables[str(num)], dtype=np.float64))

    @classmethod
    def explicit(cls, vectors: Sequence[Sequence[float]]) -> "Directions":
        v = np.array(vectors, dtype=np.float64).reshape(-1, 3)
        length = np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2])
        return cls(v * (1.0 / length)[:, None])

    @classmethod
    def from_spec(cls, spec: Union[int, Sequence]) -> "Directions":
        return cls.from_num(spec) if isinstance(spec, int) else cls.explicit(spec)

    def __len__(self) -> int:
        return len(self.xyz)


@dataclass
class SweepParameters:
    """The ``sweep:`` parameter section, same keys and defaults as src/sweep/parameters.rs."""
    directions: Union[int, Sequence] = 84
    num_timestep_levels: int = 1
    periodic: bool = False
    max_timestep: float = 1.0                      # s
    rotate_directions: bool = False
    significant_rate_threshold: float = 0.0        # 1/s
    timestep_safety_factor: float = 0.1
    chemistry_timestep_safety_factor: float = 0.1
    check_deadlock: bool = False
    prevent_cooling: bool = True
    num_tasks_to_solve_before_send_receive: int = 10000   # accepted, unused (no MPI messages)

    _KEYS = ("directions", "num_timestep_levels", "periodic", "max_timestep", "rotate_directions",
             "significant_rate_threshold", "timestep_safety_factor", "chemistry_timestep_safety_factor",
             "check_deadlock", "prevent_cooling", "num_tasks_to_solve_before_send_receive")

    @classmethod
    def from_dict(cls, section: dict) -> "SweepParameters":
        unknown = set(section) - set(cls._KEYS)
        if unknown:   # serde deny_unknown_fields
            raise ValueError(f"unknown field(s) in sweep section: {sorted(unknown)}")
        for required in ("directions", "num_timestep_levels", "periodic", "max_timestep"):
            if required not in section:
                raise ValueError(f"missing field `{required}` in sweep section")
        kw = dict(section)
        for q in ("max_timestep", "significant_rate_threshold", "timestep_safety_factor",
                  "chemistry_timestep_safety_factor"):
            if q in kw:
                kw[q] = parse_quantity(kw[q])
        return cls(**kw)

    @classmethod
    def from_yaml(cls, text: str) -> "SweepParameters":
        import yaml
        return cls.from_dict(yaml.safe_load(text)["sweep"])


def rotation_matrix(axis, angle: float) -> np.ndarray:
    """get_rotation_matrix (src/sweep/direction/mod.rs:113-134): axis-angle (Rodrigues) rotation."""
    x, y, z = axis
    c, s = np.cos(angle), np.sin(angle)
    return np.array([
        [c + x * x * (1.0 - c), x * y * (1.0 - c) - z * s, x * z * (1.0 - c) + y * s],
        [y * x * (1.0 - c) + z * s, c + y * y * (1.0 - c), y * z * (1.0 - c) - x * s],
        [z * x * (1.0 - c) - y * s, z * y * (1.0 - c) + x * s, c + z * z * (1.0 - c)]])


def random_rotation_matrix(rng: np.random.Generator) -> np.ndarray:
    """get_random_rotation_matrix (:136-148): uniform axis on the sphere, uniform angle -- the same construction, drawn
    from a numpy Generator (the reference seeds a Rust StdRng with 1337, whose stream cannot be reproduced here)."""
    phi = rng.uniform(0.0, 2.0 * np.pi)
    theta = np.arccos(2.0 * rng.uniform(0.0, 1.0) - 1.0)
    axis = (np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta))
    return rotation_matrix(axis, rng.uniform(0.0, 2.0 * np.pi))


AllReduce = Callable[[int, int, Optional[int]], None]   # (pointer, n_doubles, stream) -> in-place sum
# (op, pointer, n_doubles_per_rank, stream): in-place reduce-scatter (op 1) / all-gather (op 2) over world_size chunks
Collectives = Callable[[int, int, int, Optional[int]], None]


def direction_shard(n_dirs: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous direction shard [begin, end) of ``rank`` (ssw_direction_shard)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    return n_dirs * rank // world_size, n_dirs * (rank + 1) // world_size


def allreduce_trampoline(fn: AllReduce):
    """C callback around a Python all-reduce.  An exception must not propagate through C: it is reported on stderr and
    turned into a non-zero return, which the library maps to SSW_E_COMM."""
    def trampoline(_ctx, buf, n, stream):
        try:
            fn(int(buf), int(n), int(stream) if stream else None)
            return 0
        except Exception as exc:   # noqa: BLE001 - reported through the C error path
            print(f"subsweep_b200: allreduce hook failed: {exc!r}", file=sys.stderr)
            return -1
    return capi.ALLREDUCE_FN(trampoline)


def collective_trampoline(fn: Collectives):
    """C callback around a Python reduce-scatter / all-gather (same error contract as allreduce_trampoline)."""
    def trampoline(_ctx, op, buf, n, stream):
        try:
            fn(int(op), int(buf), int(n), int(stream) if stream else None)
            return 0
        except Exception as exc:   # noqa: BLE001 - reported through the C error path
            print(f"subsweep_b200: collective hook failed: {exc!r}", file=sys.stderr)
            return -1
    return capi.COLLECTIVE_FN(trampoline)


class Sweep:
    """``Sweep<HydrogenOnly>`` behind the C ABI (src/sweep/mod.rs:172-272)."""

    def __init__(self, parameters: SweepParameters, grid: FlatGrid, density, ionized_hydrogen_fraction,
                 temperature, source, scale_factor: float = 1.0, device_id: int = 0, rank: int = 0,
                 world_size: int = 1, allreduce: Optional[AllReduce] = None, flags: int = 0, lib=None,
                 positions="grid", collectives: Optional[Collectives] = None):
        if parameters.rotate_directions and world_size > 1:
            raise NotImplementedError("rotate_directions is not available under direction sharding")
        self.lib = lib if lib is not None else capi.load()
        self.parameters = parameters
        self.grid = grid
        self.directions = Directions.from_spec(parameters.directions)
        self.rank, self.world_size = rank, world_size
        self.dir_begin, self.dir_end = direction_shard(len(self.directions), world_size, rank)
        N = grid.n_cells
        self.n_cells = N
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (density, ionized_hydrogen_fraction, temperature, source)]
        for a in arrs:
            if a.shape != (N,):
                raise ValueError("per-cell arrays must have shape (n_cells,)")
        p = capi.Params()
        p.n_dirs = len(self.directions)
        p.dirs_xyz = capi.dptr(self.directions.xyz)
        p.n_levels = parameters.num_timestep_levels
        p.max_timestep_s = parameters.max_timestep
        p.timestep_safety_factor = parameters.timestep_safety_factor
        p.chemistry_timestep_safety_factor = parameters.chemistry_timestep_safety_factor
        p.significant_rate_threshold_per_s = parameters.significant_rate_threshold
        p.prevent_cooling = int(parameters.prevent_cooling)
        p.scale_factor = scale_factor
        p.check_deadlock = int(parameters.check_deadlock)
        p.device_id = device_id
        p.rank, p.world_size = rank, world_size
        p.flags = flags
        g = capi.Grid()
        g.n_cells = N
        self._keep = (np.ascontiguousarray(grid.face_offsets, dtype=np.uint64),
                      np.ascontiguousarray(grid.face_area, dtype=np.float64),
                      np.ascontiguousarray(grid.face_normal, dtype=np.float64),
                      np.ascontiguousarray(grid.face_neighbour, dtype=np.int32),
                      np.ascontiguousarray(grid.face_kind, dtype=np.uint8),
                      np.ascontiguousarray(grid.cell_size, dtype=np.float64),
                      np.ascontiguousarray(grid.cell_volume, dtype=np.float64))
        fo, fa, fn, fnb, fk, cs, cv = self._keep
        g.face_offsets = fo.ctypes.data_as(C.POINTER(C.c_uint64))
        g.face_area = capi.dptr(fa)
        g.face_normal = capi.dptr(fn)
        g.face_neighbour = fnb.ctypes.data_as(C.POINTER(C.c_int32))
        g.face_kind = fk.ctypes.data_as(C.POINTER(C.c_uint8))
        g.cell_size = capi.dptr(cs)
        g.cell_volume = capi.dptr(cv)
        self._h = C.c_void_p()
        self._check(self.lib.ssw_create(C.byref(p), C.byref(g), *(capi.dptr(a) for a in arrs), C.byref(self._h)))
        self._cb = None
        self._coll_cb = None
        # the Position component (optional): lets the library run the all-cells sweep patch by patch
        if isinstance(positions, str):
            positions = getattr(grid, "positions", None) if positions == "grid" else None
        if positions is not None:
            pos = np.ascontiguousarray(positions, dtype=np.float64)
            if pos.shape != (N, 3):
                raise ValueError("positions must have shape (n_cells, 3)")
            self._check(self.lib.ssw_set_cell_positions(self._h, capi.dptr(pos)))
        # world_size > 1: either attach the peers' arenas (peer_attach / peer_attach_ipc, the NVLink path) or give hooks
        if world_size > 1 and allreduce is not None:
            self.set_allreduce(allreduce)
            if collectives is not None:
                self.set_collectives(collectives)

    # -- plumbing ----------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise capi.SubsweepError(rc, (self.lib.ssw_last_error() or b"").decode())

    def set_allreduce(self, fn: AllReduce) -> None:
        self._cb = allreduce_trampoline(fn)
        self._check(self.lib.ssw_set_allreduce(self._h, self._cb, None))

    def set_collectives(self, fn: Collectives) -> None:
        """Reduce-scatter / all-gather hook (ssw_set_collectives): chemistry sliced by cells instead of replicated."""
        self._coll_cb = collective_trampoline(fn)
        self._check(self.lib.ssw_set_collectives(self._h, self._coll_cb, None))

    # -- peer-mapped direction sharding (include/subsweep_b200.h, "direction sharding without hooks") -------------
    def peer_arena(self) -> tuple[int, int]:
        """(device pointer, bytes) of this rank's arena."""
        base, nbytes = C.c_void_p(), C.c_uint64()
        self._check(self.lib.ssw_peer_arena(self._h, C.byref(base), C.byref(nbytes)))
        return int(base.value), int(nbytes.value)

    def peer_export(self) -> bytes:
        """The arena's CUDA IPC handle (64 bytes) for the other processes of the box."""
        buf = C.create_string_buffer(capi.PEER_HANDLE_BYTES)
        self._check(self.lib.ssw_peer_export(self._h, buf))
        return buf.raw

    def peer_attach_ipc(self, handles: Sequence[bytes]) -> None:
        """Attach the arenas of all ranks from their IPC handles (rank order; the own entry is ignored)."""
        blob = b"".join(handles)
        if len(blob) != capi.PEER_HANDLE_BYTES * self.world_size:
            raise ValueError("need one 64-byte handle per rank")
        self._check(self.lib.ssw_peer_attach_ipc(self._h, C.c_char_p(blob)))

    def peer_attach(self, bases: Sequence[int]) -> None:
        """Attach the arenas of all ranks by device pointer (handles of one process)."""
        arr = (C.c_void_p * self.world_size)(*[C.c_void_p(b) for b in bases])
        self._check(self.lib.ssw_peer_attach(self._h, arr))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self.lib.ssw_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the hot path --------------------------------------------------------------------------
    def run_sweeps(self) -> float:
        """Sweep::run_sweeps: one full step; returns the elapsed simulation time in s."""
        t = C.c_double()
        self._check(self.lib.ssw_run_sweeps(self._h, C.byref(t)))
        return t.value

    def single_sweep(self, level: int) -> None:
        self._check(self.lib.ssw_single_sweep(self._h, level))

    def update_timestep_levels(self) -> None:
        self._check(self.lib.ssw_update_timestep_levels(self._h))

    def set_directions(self, xyz) -> None:
        """rotate_directions_system (src/sweep/direction/mod.rs:158-174): continue with another direction set of the same
        size; the flux state follows the best aligned old direction (ssw_set_directions)."""
        v = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        if v.shape != self.directions.xyz.shape:
            raise ValueError("the new direction set must have the same size")
        self._check(self.lib.ssw_set_directions(self._h, capi.dptr(v)))
        self.directions = Directions(v)

    def set_inputs(self, density=None, source=None) -> None:
        d = None if density is None else np.ascontiguousarray(density, dtype=np.float64)
        s = None if source is None else np.ascontiguousarray(source, dtype=np.float64)
        self._check(self.lib.ssw_set_inputs(self._h, capi.dptr(d) if d is not None else None,
                                            capi.dptr(s) if s is not None else None))

    # -- read-back -------------------------------------------------------------------------------
    def read(self, field_name: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.n_cells)
        self._check(self.lib.ssw_read(self._h, capi.FIELDS[field_name], capi.dptr(out)))
        return out

    def read_begin(self, field_name: str, out: Optional[np.ndarray]) -> None:
        """Queue the read-back of a field (``out`` pinned host memory; None on a worker rank); valid after sync()."""
        self._check(self.lib.ssw_read_begin(self._h, capi.FIELDS[field_name], None if out is None else capi.dptr(out)))

    def sync(self) -> None:
        self._check(self.lib.ssw_sync(self._h))

    def read_as_worker(self, field_name: str) -> None:
        """Worker rank of a sharded job: take part in the field's collective, copy nothing (ssw_read with NULL)."""
        self._check(self.lib.ssw_read(self._h, capi.FIELDS[field_name], None))

    def time_series(self, mass=None, with_rates: bool = False) -> dict:
        """compute_time_series_system (src/sweep/time_series.rs:61-155), reduced on the device; keys are the
        reference's time-series names."""
        ts = capi.TimeSeries()
        m = None if mass is None else np.ascontiguousarray(mass, dtype=np.float64)
        self._check(self.lib.ssw_time_series_compute(self._h, capi.dptr(m) if m is not None else None,
                                                     int(with_rates), C.byref(ts)))
        return {k: getattr(ts, k) for k, _ in capi.TimeSeries._fields_}

    def num_particles_at_timestep_levels(self) -> list:
        """num_particles_at_timestep_levels_system (time_series.rs:167-188): cumulative counts per level."""
        counts = self.level_counts()
        return [{"level": l, "num": int(counts[l]), "timestep": self.parameters.max_timestep * 0.5 ** l}
                for l in range(self.parameters.num_timestep_levels)]

    def levels(self) -> np.ndarray:
        out = np.empty(self.n_cells, dtype=np.uint8)
        self._check(self.lib.ssw_read_levels(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def chem_attempts(self) -> np.ndarray:
        """Substep attempts of every cell's last chemistry update (ssw_read_chem_attempts; 0 = never updated)."""
        out = np.empty(self.n_cells, dtype=np.uint16)
        self._check(self.lib.ssw_read_chem_attempts(self._h, out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    def level_counts(self) -> np.ndarray:
        out = np.zeros(self.parameters.num_timestep_levels, dtype=np.uint64)
        self._check(self.lib.ssw_level_counts(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def lowest_allowed_level(self) -> int:
        v = C.c_int32()
        self._check(self.lib.ssw_lowest_allowed_level(self._h, C.byref(v)))
        return v.value

    def set_levels(self, levels) -> None:
        lv = np.ascontiguousarray(levels, dtype=np.uint8)
        self._check(self.lib.ssw_set_levels(self._h, lv.ctypes.data_as(C.POINTER(C.c_uint8))))

    def set_change_timescale(self, tau) -> None:
        t = np.ascontiguousarray(tau, dtype=np.float64)
        self._check(self.lib.ssw_set_change_timescale(self._h, capi.dptr(t)))

    def dir_state(self, which: str) -> np.ndarray:
        idx = {"incoming": 0, "outgoing": 1, "periodic": 2}[which]
        out = np.empty((self.n_cells, self.dir_end - self.dir_begin))
        self._check(self.lib.ssw_read_dir_state(self._h, idx, capi.dptr(out)))
        return out

    def wavefront_levels(self, level: int, direction: int) -> np.ndarray:
        out = np.empty(self.n_cells, dtype=np.int32)
        self._check(self.lib.ssw_read_wavefront_levels(self._h, level, direction,
                                                       out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def patch_note(self) -> str:
        """Why the patch-ordered all-cells sweep is not in use ('' = in use, or not tried yet)."""
        return (self.lib.ssw_patch_note(self._h) or b"").decode()

    def stat(self, name: str) -> int:
        v = C.c_uint64()
        self._check(self.lib.ssw_get_stat(self._h, capi.STATS[name], C.byref(v)))
        return v.value

    def timings(self) -> dict:
        t = capi.Timings()
        self._check(self.lib.ssw_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    def reset_timings(self) -> None:
        self._check(self.lib.ssw_reset_timings(self._h))

    def set_timing_level(self, level: int) -> None:
        """0: whole steps and the all-cells sweep kernel only; 1: every phase of every single sweep (default)."""
        self._check(self.lib.ssw_set_timing_level(self._h, int(level)))


# ---------------------------------------------------------------------------------------------
# the two bevy systems of SweepPlugin, on a dict of per-particle component arrays
# ---------------------------------------------------------------------------------------------
@dataclass
class SweepPlugin:
    """``SweepPlugin`` (src/sweep/mod.rs:111-169): init in StartupStages::InitSweep, one
    ``run_sweep_system`` per update.  ``components`` maps the reference's snapshot field names
    (src/components.rs:14-83) to numpy arrays in ParticleId.index order."""
    parameters: SweepParameters
    scale_factor: float = 1.0
    device_id: int = 0
    rank: int = 0
    world_size: int = 1
    allreduce: Optional[AllReduce] = None
    collectives: Optional[Collectives] = None
    solver: Optional[Sweep] = field(default=None, init=False)
    directions_rng: np.random.Generator = field(default_factory=lambda: np.random.default_rng(1337), init=False)  # DIRECTIONS_RNG_SEED
    is_first_time: bool = field(default=True, init=False)
    simulation_time: float = field(default=0.0, init=False)

    def init_sweep_system(self, grid: FlatGrid, components: dict) -> None:
        self.solver = Sweep(self.parameters, grid, components["density"],
                            components["ionized_hydrogen_fraction"], components["temperature"],
                            components["source"], scale_factor=self.scale_factor,
                            device_id=self.device_id, rank=self.rank, world_size=self.world_size,
                            allreduce=self.allreduce, collectives=self.collectives,
                            positions=components.get("position", "grid"))
        n = grid.n_cells
        components.setdefault("photon_rate", np.zeros(n))
        components.setdefault("timestep", np.zeros(n))
        components.setdefault("ionization_time", np.full(n, np.inf))   # IonizationTime::default()

    def run_sweep_system(self, components: dict) -> None:
        # the first call is a no-op so that the initial conditions get written (mod.rs:711-714)
        if self.is_first_time:
            self.is_first_time = False
            return
        s = self.solver
        if self.parameters.rotate_directions:   # rotate_directions_system runs in front of the sweep (mod.rs:143-149)
            m = random_rotation_matrix(self.directions_rng)
            s.set_directions(s.directions.xyz @ m.T)
        self.simulation_time += s.run_sweeps()
        s.read("ionized_hydrogen_fraction", components["ionized_hydrogen_fraction"])
        s.read("temperature", components["temperature"])
        s.read("timestep", components["timestep"])
        s.read("photon_rate", components["photon_rate"])
        s.read("ionization_time", components["ionization_time"])
