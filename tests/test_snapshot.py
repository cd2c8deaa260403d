"""Snapshot output: dataset names, dimension attributes, static components, file layout of the
reference's OutputPlugin (src/io/output/mod.rs, src/components.rs), in the container format available here."""
import json

import numpy as np
import pytest

from subsweep_b200 import snapshot as S
from subsweep_b200 import units as U


def components(n, seed=0):
    rng = np.random.default_rng(seed)
    return {
        "position": rng.uniform(0, 1, (n, 3)), "density": rng.uniform(1, 2, n), "source": np.zeros(n),
        "ionized_hydrogen_fraction": rng.uniform(0, 1, n), "temperature": rng.uniform(10, 1e4, n),
        "photon_rate": rng.uniform(0, 1e50, n), "timestep": rng.uniform(1, 2, n) * U.MEGAYEARS,
        "ionization_time": np.full(n, np.inf),
    }


def test_component_table_matches_the_reference():
    # names: src/components.rs:14-83; dimensions and static flags: :103-119 with src/units/mod.rs:15-99
    assert set(S.COMPONENTS) == {"position", "density", "mass", "ionized_hydrogen_fraction", "temperature", "source",
                                 "photon_rate", "photoionization_rate", "recombination_rate",
                                 "collisional_ionization_rate", "heating_rate", "timestep", "ionization_time"}
    static = {k for k, (_, st, _) in S.COMPONENTS.items() if st}
    assert static == {"position", "density", "source", "mass"}
    a = S.COMPONENTS["heating_rate"][0].attrs()      # Energy / (Volume3D * Time) = kg m^-1 s^-3
    assert (a["scaling_mass"], a["scaling_length"], a["scaling_time"], a["scaling_temperature"]) == (1, -1, -3, 0)
    a = S.COMPONENTS["density"][0].attrs()
    assert (a["scaling_mass"], a["scaling_length"], a["scale_factor_si"]) == (1, -3, 1.0)
    assert S.COMPONENTS["photon_rate"][0].attrs()["scaling_time"] == -1
    assert list(S.COMPONENTS["temperature"][0].attrs()) == ["scale_factor_si", "scaling_length", "scaling_time", "scaling_mass",
                                                            "scaling_temperature", "scaling_h", "scaling_a"]


def test_snapshots_layout_static_fields_and_round_trip(tmp_path):
    p = S.OutputParameters(output_dir=tmp_path / "output", time_between_snapshots=1.0 * U.MEGAYEARS)
    w = S.SnapshotWriter(p)
    c = components(1000)
    assert w.should_write(0.0)
    d0 = w.write(c, 0.0, {"scale_factor": 1.0})
    assert d0 == tmp_path / "output" / "snapshots" / "000" and (d0 / "0" / "attributes.json").exists()
    assert not w.should_write(0.5 * U.MEGAYEARS) and w.should_write(1.0 * U.MEGAYEARS)
    assert w.should_write(0.1, simulation_finished=True)                       # StopSimulationEvent forces a snapshot
    c["ionized_hydrogen_fraction"] = c["ionized_hydrogen_fraction"] * 0.5
    d1 = w.write(c, 1.0 * U.MEGAYEARS)
    assert d1.name == "001"
    data0, attrs0, file0 = S.read_snapshot(d0)
    data1, _, file1 = S.read_snapshot(d1)
    assert set(data0) == set(c)                                                # first snapshot: everything
    assert set(data1) == set(c) - {"position", "density", "source"}           # later ones: dynamic components only
    for k in data1:
        assert np.array_equal(data1[k], c[k])
    assert file0 == {"time": 0.0, "scale_factor": 1.0} and file1 == {"time": 1.0 * U.MEGAYEARS}
    assert attrs0["timestep"]["scaling_time"] == 1 and attrs0["position"]["scaling_length"] == 1
    assert np.isposinf(data0["ionization_time"]).all()                         # IonizationTime::default() survives the format


def test_fields_subset_and_several_output_files(tmp_path):
    p = S.OutputParameters(output_dir=tmp_path, fields=["temperature", "density"], num_output_files=3, snapshot_padding=2)
    w = S.SnapshotWriter(p)
    c = components(10)
    d = w.write(c, 3.0)
    assert d.name == "00" and sorted(x.name for x in d.iterdir()) == ["0", "1", "2"]
    regions = [json.loads((d / str(i) / "attributes.json").read_text())["region"] for i in range(3)]
    assert regions == [[0, 3], [3, 6], [6, 10]]                                 # total / n each, the last takes the rest
    data, _, _ = S.read_snapshot(d)
    assert set(data) == {"temperature", "density"} and np.array_equal(data["temperature"], c["temperature"])
    with pytest.raises(KeyError):
        w.write({"bogus": np.zeros(3)}, 0.0)
    with pytest.raises(ValueError):
        w.write({"temperature": np.zeros((3, 2))}, 0.0)


def test_output_section_parsing():
    p = S.OutputParameters.from_dict({"output_dir": "out", "time_between_snapshots": 5.0, "snapshot_padding": 4,
                                      "time_series_dir": "ts", "handle_existing_output": "delete"})
    assert p.snapshot_dir().as_posix() == "out/snapshots" and p.snapshot_padding == 4 and p.is_desired_field("anything")
