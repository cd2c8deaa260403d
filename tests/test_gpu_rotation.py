"""Direction rotation (src/sweep/direction/mod.rs:137-205, `sweep.rotate_directions`) through ssw_set_directions."""
import numpy as np
import pytest

from helpers import assert_close, make_problem
from subsweep_b200 import Sweep, SweepPlugin
from subsweep_b200.sweep import Directions, random_rotation_matrix

pytestmark = pytest.mark.gpu

FIELDS = ("ionized_hydrogen_fraction", "temperature", "change_timescale", "photon_rate")


@pytest.mark.parametrize("kind,n,periodic", [("cartesian", 10, True), ("voronoi", 8, False)])
def test_permuting_the_directions_changes_nothing_physical(cuda_lib, kind, n, periodic):
    """A 'rotation' that only renumbers the direction bins: every new direction is an old one, the remapped flux state
    is the old state under new labels, and the run continues as if nothing had happened (the per-cell rate is folded
    over the directions in another order: round-off)."""
    params, g, f = make_problem(kind, n, periodic, n_dirs=21, n_levels=2, max_timestep_myr=0.25)
    a, b = Sweep(params, g, **f), Sweep(params, g, **f)
    for _ in range(3):
        a.run_sweeps()
        b.run_sweeps()
    perm = np.random.default_rng(3).permutation(21)
    out_before = b.dir_state("outgoing")
    b.set_directions(b.directions.xyz[perm])
    assert_close(b.dir_state("outgoing"), out_before[:, perm], 1e-13, floor=1e-300, what="remapped outgoing rates")
    for _ in range(3):
        a.run_sweeps()
        b.run_sweeps()
    assert np.array_equal(a.levels(), b.levels())
    for k in FIELDS:
        v = a.read(k)
        assert_close(b.read(k), v, 1e-10, floor=1e-7 * np.nanmax(np.abs(v)), what=k)
    assert_close(b.dir_state("outgoing"), a.dir_state("outgoing")[:, perm], 1e-10, floor=1e-7 * a.dir_state("outgoing").max())


def test_rotated_run_stays_physical_and_close(cuda_lib):
    """`rotate_directions: true` through the plugin surface: a random rotation before every sweep.  The result is a
    different discretisation of the same problem: same photon budget to within the angular resolution."""
    import dataclasses
    params, g, f = make_problem("cartesian", 12, True, n_dirs=84, n_levels=2, max_timestep_myr=0.25)
    fixed = {k: v.copy() for k, v in f.items()}
    turned = {k: v.copy() for k, v in f.items()}
    p_fixed, p_turned = SweepPlugin(params), SweepPlugin(dataclasses.replace(params, rotate_directions=True))
    p_fixed.init_sweep_system(g, fixed)
    p_turned.init_sweep_system(g, turned)
    for _ in range(7):
        p_fixed.run_sweep_system(fixed)
        p_turned.run_sweep_system(turned)
    d = p_turned.solver.directions.xyz
    assert np.allclose(np.linalg.norm(d, axis=1), np.linalg.norm(Directions.from_num(84).xyz, axis=1), rtol=1e-12)
    assert not np.allclose(d, Directions.from_num(84).xyz)
    x_f, x_t = fixed["ionized_hydrogen_fraction"], turned["ionized_hydrogen_fraction"]
    assert np.all((x_t >= 1e-10) & (x_t <= 1.0)) and np.all(np.isfinite(turned["temperature"]))
    assert abs(x_t.mean() - x_f.mean()) < 0.2 * x_f.mean()
    assert p_turned.simulation_time == p_fixed.simulation_time
