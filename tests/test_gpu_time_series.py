"""The step after the path: time series (src/sweep/time_series.rs:61-188) reduced on the device,
against the reference's formulas evaluated on the oracle's state."""
import math

import numpy as np
import pytest

import oracle
from helpers import make_problem
from subsweep_b200 import Sweep

pytestmark = pytest.mark.gpu


def reference_time_series(ref, g, density, mass=None):
    """compute_time_series_system, time_series.rs:61-155 (exact sums, the GPU reduces in another order)."""
    x, T = ref.read("ionized_hydrogen_fraction"), ref.read("temperature")
    vol = g.cell_volume
    m = density * vol if mass is None else mass
    gamma = ref.read("photoionization_rate")
    fs = math.fsum
    return {
        "hydrogen_ionization_mass_average": fs(m * x) / fs(m),
        "hydrogen_ionization_volume_average": fs(vol * x) / fs(vol),
        "temperature_mass_average": fs(T * m) / fs(m),
        "temperature_volume_average": fs(T * vol) / fs(vol),
        "photoionization_rate_volume_average": fs(gamma * vol) / fs(vol),
        "weighted_photoionization_rate_volume_average": fs(gamma * x * vol) / fs(vol),
    }


@pytest.mark.parametrize("kind,n,periodic", [("cartesian", 11, True), ("voronoi", 8, False)])
def test_time_series_matches_reference_formulas(cuda_lib, kind, n, periodic):
    params, g, f = make_problem(kind, n, periodic, n_dirs=21, n_levels=3, max_timestep_myr=0.25)
    mode = oracle.PERIODIC_LAGGED if periodic else oracle.PERIODIC_HEAP
    got, ref = Sweep(params, g, **f), oracle.OracleSweep(params, g, **f, periodic_mode=mode)
    for _ in range(5):
        got.run_sweeps()
        ref.run_sweeps()
    want = reference_time_series(ref, g, f["density"])
    ts = got.time_series(with_rates=True)
    for k, v in want.items():
        assert ts[k] == pytest.approx(v, rel=1e-6), k
    # against the library's own per-cell read-back the reduction itself is accurate to round-off
    x, T = got.read("ionized_hydrogen_fraction"), got.read("temperature")
    vol, m = g.cell_volume, f["density"] * g.cell_volume
    assert ts["hydrogen_ionization_mass_average"] == pytest.approx(math.fsum(m * x) / math.fsum(m), rel=1e-13)
    assert ts["temperature_volume_average"] == pytest.approx(math.fsum(T * vol) / math.fsum(vol), rel=1e-13)
    assert ts["total_volume"] == pytest.approx(math.fsum(vol), rel=1e-13)
    # an explicit Mass component and no rates
    rng = np.random.default_rng(1)
    mass = m * rng.uniform(0.5, 2.0, g.n_cells)
    ts2 = got.time_series(mass=mass)
    assert ts2["hydrogen_ionization_mass_average"] == pytest.approx(math.fsum(mass * x) / math.fsum(mass), rel=1e-13)
    assert math.isnan(ts2["photoionization_rate_volume_average"])
    # num_particles_at_timestep_levels_system, time_series.rs:167-188
    levels = got.num_particles_at_timestep_levels()
    assert [e["num"] for e in levels] == [int(v) for v in ref.level_counts()]
    assert levels[1]["timestep"] == params.max_timestep / 2
