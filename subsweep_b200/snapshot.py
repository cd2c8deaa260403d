"""Snapshot output of the per-particle components.

The reference writes one 1-D f64 dataset per component into ``<output_dir>/snapshots/<NNN>/<file>.hdf5``
(src/io/output/mod.rs:142-172, 265-343): dataset name = ``Named::name()`` of the component
(src/components.rs:14-83), six integer attributes with the exponents of the dimension
(``scaling_length``, ``scaling_time``, ``scaling_mass``, ``scaling_temperature``, ``scaling_h``, ``scaling_a``,
src/io/output/mod.rs:42-48) and ``scale_factor_si`` = the factor to SI of the stored numbers (1.0: everything is stored
in SI base units, src/units/dimension.rs:28-46).  Static components (position, density, mass, source) are written into
the first snapshot only (src/io/output/timer.rs:43-58, src/components.rs:103-119).  File attributes: ``time`` and, with
cosmology, ``scale_factor``, ``redshift``, ``little_h`` (src/simulation_plugin/mod.rs:78-102).

HDF5 is not available in this image, so the same content is stored as one ``.npy`` file per dataset plus a JSON side
car holding the attributes -- names, values, units and directory layout exactly as above -- and ``to_hdf5`` converts
such a directory into the reference's single-file form wherever ``h5py`` exists.  Host-side post-processing of the
arrays ``ssw_read`` returns; nothing here touches the device.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Iterable, Mapping, Optional, Union

import numpy as np

SCALE_FACTOR_IDENTIFIER = "scale_factor_si"
DIMENSION_IDENTIFIERS = ("scaling_length", "scaling_time", "scaling_mass", "scaling_temperature", "scaling_h", "scaling_a")


@dataclass(frozen=True)
class Dimension:
    """Exponents of the reference's `Dimension` struct (src/units/dimension.rs:8-16)."""
    length: int = 0
    time: int = 0
    mass: int = 0
    temperature: int = 0
    h: int = 0
    a: int = 0

    def base_conversion_factor(self) -> float:
        return 1.0          # LENGTH/TIME/MASS/TEMPERATURE_TO_SI are all 1.0 (src/units/dimension.rs:3-6, 32-46)

    def attrs(self) -> dict:
        out = {SCALE_FACTOR_IDENTIFIER: self.base_conversion_factor()}
        out.update(zip(DIMENSION_IDENTIFIERS, (self.length, self.time, self.mass, self.temperature, self.h, self.a)))
        return out


_RATE = Dimension(time=-1)
# component name -> (dimension, static, columns); src/components.rs:14-83 (names), :103-119 (dimension, static)
COMPONENTS = {
    "position": (Dimension(length=1), True, 3),
    "density": (Dimension(mass=1, length=-3), True, 1),
    "mass": (Dimension(mass=1), True, 1),
    "source": (_RATE, True, 1),
    "ionized_hydrogen_fraction": (Dimension(), False, 1),
    "temperature": (Dimension(temperature=1), False, 1),
    "photon_rate": (_RATE, False, 1),
    "photoionization_rate": (_RATE, False, 1),
    "recombination_rate": (_RATE, False, 1),
    "collisional_ionization_rate": (_RATE, False, 1),
    "heating_rate": (Dimension(mass=1, length=-1, time=-3), False, 1),     # Energy / (Volume3D * Time)
    "timestep": (Dimension(time=1), False, 1),
    "ionization_time": (Dimension(time=1), False, 1),
}


@dataclass
class OutputParameters:
    """The `output:` section as far as snapshots go (src/io/output/parameters.rs:14-95, same defaults)."""
    output_dir: Union[str, Path] = "output"
    snapshots_dir: Union[str, Path] = "snapshots"
    fields: Union[str, Iterable[str]] = "all"        # 'all' or the names of the fields to write
    snapshot_padding: int = 3
    num_output_files: int = 1
    time_between_snapshots: float = 0.0              # seconds; 0 = every step
    time_first_snapshot: Optional[float] = None

    def snapshot_dir(self) -> Path:
        return Path(self.output_dir) / Path(self.snapshots_dir)

    def is_desired_field(self, name: str) -> bool:
        return self.fields == "all" or name in self.fields

    @classmethod
    def from_dict(cls, section: Mapping) -> "OutputParameters":
        known = {"output_dir", "snapshots_dir", "fields", "snapshot_padding", "num_output_files",
                 "time_between_snapshots", "time_first_snapshot"}
        # the section has more keys (time series, performance file, handle_existing_output ...): not snapshot business
        kw = {k: v for k, v in section.items() if k in known}
        return cls(**kw)


class SnapshotWriter:
    """Timer + writer: `OutputPlugin` for the per-particle components (src/io/output/timer.rs, mod.rs:151-343)."""

    def __init__(self, parameters: OutputParameters):
        if parameters.num_output_files < 1:
            raise ValueError("num_output_files must be >= 1")
        self.parameters = parameters
        self.snapshot_num = 0
        self.next_output_time = parameters.time_first_snapshot if parameters.time_first_snapshot is not None else 0.0

    # Timer::run_criterion (timer.rs:30-41)
    def should_write(self, simulation_time: float, simulation_finished: bool = False) -> bool:
        return simulation_finished or simulation_time >= self.next_output_time

    def snapshot_path(self) -> Path:
        return self.parameters.snapshot_dir() / f"{self.snapshot_num:0{self.parameters.snapshot_padding}d}"

    def write(self, components: Mapping[str, np.ndarray], simulation_time: float,
              file_attributes: Optional[Mapping[str, float]] = None) -> Path:
        """Write one snapshot of `components` (name -> array in SI, particle order) and advance the timer."""
        p = self.parameters
        directory = self.snapshot_path()
        directory.mkdir(parents=True, exist_ok=True)
        first = self.snapshot_num == 0
        n_particles = None
        datasets = {}
        for name, data in components.items():
            if name not in COMPONENTS:
                raise KeyError(f"unknown component {name!r}")
            dim, static, cols = COMPONENTS[name]
            if not p.is_desired_field(name) or (static and not first):      # timer.rs:43-58
                continue
            arr = np.ascontiguousarray(data, dtype=np.float64)
            want = (arr.shape[0],) if cols == 1 else (arr.shape[0], cols)
            if arr.shape != want:
                raise ValueError(f"{name}: expected shape (N,{'' if cols == 1 else cols}), got {arr.shape}")
            if n_particles is None:
                n_particles = arr.shape[0]
            elif arr.shape[0] != n_particles:
                raise ValueError(f"{name}: {arr.shape[0]} particles, other components have {n_particles}")
            datasets[name] = (arr, dim)
        n_particles = n_particles or 0
        # particles are dealt to the files in contiguous regions: total / n each, the last file takes the remainder
        # (get_output_rank_assignment, src/io/file_distribution.rs:105-116); file names are zero-padded to
        # floor(log10(num_output_files)) + 1 digits (src/io/output/mod.rs:157-168)
        n_files = p.num_output_files
        pad = int(np.floor(np.log10(n_files))) + 1
        share = n_particles // n_files
        start = 0
        attrs = {"time": float(simulation_time)}
        attrs.update({k: float(v) for k, v in (file_attributes or {}).items()})
        for fi in range(n_files):
            count = share if fi + 1 < n_files else n_particles - share * (n_files - 1)
            fdir = directory / f"{fi:0{pad}d}"
            fdir.mkdir(exist_ok=True)
            meta = {"format": "subsweep-b200 snapshot v1 (npy per dataset; see subsweep_b200/snapshot.py)",
                    "attributes": attrs, "region": [start, start + count], "datasets": {}}
            for name, (arr, dim) in datasets.items():
                np.save(fdir / f"{name}.npy", arr[start:start + count])
                meta["datasets"][name] = dim.attrs()
            (fdir / "attributes.json").write_text(json.dumps(meta, indent=1))
            start += count
        self.snapshot_num += 1                                              # Timer::update_system (timer.rs:60-63)
        self.next_output_time += p.time_between_snapshots
        return directory


def read_snapshot(directory: Union[str, Path]) -> tuple[dict, dict, dict]:
    """Read a snapshot directory back: (datasets, dataset attributes, file attributes), files concatenated."""
    directory = Path(directory)
    parts = sorted(d for d in directory.iterdir() if d.is_dir())
    data, dattrs, fattrs = {}, {}, {}
    for d in parts:
        meta = json.loads((d / "attributes.json").read_text())
        fattrs = meta["attributes"]
        for name, a in meta["datasets"].items():
            data.setdefault(name, []).append(np.load(d / f"{name}.npy"))
            dattrs[name] = a
    return {k: np.concatenate(v) for k, v in data.items()}, dattrs, fattrs


def to_hdf5(directory: Union[str, Path]) -> list:
    """Convert a snapshot directory into the reference's files ``<dir>/<i>.hdf5`` (needs h5py)."""
    import h5py   # not available in the build image; the converter runs wherever the reference's tooling does
    directory = Path(directory)
    out = []
    for d in sorted(x for x in directory.iterdir() if x.is_dir()):
        meta = json.loads((d / "attributes.json").read_text())
        path = directory / f"{d.name}.hdf5"
        with h5py.File(path, "w") as f:
            for k, v in meta["attributes"].items():
                f.attrs[k] = np.float64(v)
            for name, a in meta["datasets"].items():
                ds = f.create_dataset(name, data=np.load(d / f"{name}.npy"))
                ds.attrs[SCALE_FACTOR_IDENTIFIER] = np.float64(a[SCALE_FACTOR_IDENTIFIER])
                for key in DIMENSION_IDENTIFIERS:
                    ds.attrs[key] = np.int32(a[key])
        out.append(path)
    return out


def remap_from(positions, temperature, ionized_hydrogen_fraction, old_positions, old_temperature,
               old_ionized_hydrogen_fraction, box_size=None):
    """``remap_abundances_and_energies_system`` (src/arepo_postprocess/remap.rs:380-429): initialise a run from the last
    snapshot of an earlier one.  Every particle looks up the old particle nearest to its position (periodic box if
    ``box_size`` is given) and keeps the larger temperature and the larger ionized fraction (``remap_from``,
    remap.rs:381-384).  Host-side preprocessing in front of ``Sweep``; returns the new (temperature, fraction)."""
    from scipy.spatial import cKDTree

    old = np.asarray(old_positions, dtype=np.float64)
    new = np.asarray(positions, dtype=np.float64)
    if box_size is not None:
        old, new = np.mod(old, box_size), np.mod(new, box_size)
    tree = cKDTree(old, boxsize=box_size)
    _, idx = tree.query(new)
    return (np.maximum(np.asarray(temperature, dtype=np.float64), np.asarray(old_temperature, dtype=np.float64)[idx]),
            np.maximum(np.asarray(ionized_hydrogen_fraction, dtype=np.float64),
                       np.asarray(old_ionized_hydrogen_fraction, dtype=np.float64)[idx]))
