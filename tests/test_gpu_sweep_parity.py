"""The CUDA sweep + chemistry against the CPU oracle on identical grids and inputs, through the C ABI.

Tolerances (north_star): ionized fraction, temperature and rates within relative 1e-9 after one
step and 1e-6 after a full run; timestep levels / active sets bit-exact.  On periodic grids the
oracle runs in LAGGED mode -- the order-independent definition of the reference's periodic lag
(DESIGN.md section 4); on non-periodic grids HEAP (the reference's exact task order) is used.
"""
import numpy as np
import pytest

import oracle
from helpers import assert_close, make_problem
from subsweep_b200 import Sweep

pytestmark = pytest.mark.gpu

CELL_FIELDS = ("ionized_hydrogen_fraction", "temperature", "timestep", "change_timescale", "previous_rate", "photon_rate")


def compare(got, ref, rtol, dir_states=True):
    for name in CELL_FIELDS:
        a, b = got.read(name), ref.read(name)
        floor = 1e-7 * np.nanmax(np.abs(b)) if name in ("previous_rate", "photon_rate") else 0.0
        assert_close(a, b, rtol, floor=floor, what=name)
    if dir_states:
        for which in ("outgoing", "incoming", "periodic"):
            a, b = got.dir_state(which), ref.dir_state(which)
            assert_close(a, b, rtol, floor=1e-7 * max(np.abs(b).max(), 1e-300), what=which)
    assert np.array_equal(got.levels(), ref.levels())
    assert np.array_equal(got.level_counts(), ref.level_counts())
    assert got.lowest_allowed_level() == ref.lowest_allowed_level()


def pair(params, g, f, periodic, **kw):
    mode = oracle.PERIODIC_LAGGED if periodic else oracle.PERIODIC_HEAP
    return Sweep(params, g, **f, **kw), oracle.OracleSweep(params, g, **f, periodic_mode=mode)


@pytest.mark.parametrize("kind,n,periodic,n_dirs", [
    ("cartesian", 12, False, 84), ("cartesian", 12, True, 84), ("cartesian", 10, True, 21),
    ("voronoi", 9, False, 84), ("voronoi", 9, True, 84), ("jittered", 8, True, 16),
    ("cartesian", 9, False, 1), ("cartesian", 9, True, 1),
])
def test_one_step_single_level(cuda_lib, kind, n, periodic, n_dirs):
    params, g, f = make_problem(kind, n, periodic, n_dirs=n_dirs, n_levels=1)
    got, ref = pair(params, g, f, periodic)
    assert got.run_sweeps() == ref.run_sweeps()
    compare(got, ref, 1e-9)
    assert got.stat("tasks_solved") == ref.stat("tasks_solved") == g.n_cells * n_dirs


@pytest.mark.parametrize("kind,n,periodic,n_levels", [
    ("cartesian", 12, True, 4), ("cartesian", 11, False, 3), ("voronoi", 9, True, 4), ("voronoi", 8, False, 2),
])
def test_one_myr_with_timestep_levels(cuda_lib, kind, n, periodic, n_levels):
    """Warm-up of the levels plus steady-state steps: 1 + 1 + 2 + 4 + ... single sweeps."""
    params, g, f = make_problem(kind, n, periodic, n_dirs=21, n_levels=n_levels, max_timestep_myr=0.25)
    got, ref = pair(params, g, f, periodic)
    elapsed_g = elapsed_r = 0.0
    for step in range(n_levels + 3):
        elapsed_g += got.run_sweeps()
        elapsed_r += ref.run_sweeps()
        if step == 0:
            compare(got, ref, 1e-9)
        assert np.array_equal(got.levels(), ref.levels()), f"active sets differ after step {step}"
    assert elapsed_g == elapsed_r
    compare(got, ref, 1e-6)
    assert got.stat("tasks_solved") == ref.stat("tasks_solved")
    assert got.stat("single_sweeps") == ref.stat("single_sweeps")
    # the sub-level sweeps really ran on partial active sets
    assert ref.level_counts()[0] == g.n_cells


def test_bench_sweep_configuration(cuda_lib):
    """benches/sweep/main.rs:36-100 re-parameterised: random points, non-periodic Voronoi, 84
    directions, 3 levels, zero sources, threshold 0, the test components of
    initialize_sweep_test_components_system (src/sweep/mod.rs:783-797)."""
    from subsweep_b200 import SweepParameters, grid as G
    rng = np.random.default_rng(1338)
    pts = rng.uniform(0.0, 1e5, size=(500, 3))
    g = G.voronoi(pts, 1e5, periodic=False)
    params = SweepParameters(directions=84, num_timestep_levels=3, periodic=False, max_timestep=1e-3,
                             significant_rate_threshold=0.0, prevent_cooling=False)
    N = g.n_cells
    f = dict(density=np.full(N, 1e-10 / 1e-6), ionized_hydrogen_fraction=np.full(N, 1e-10),
             temperature=np.full(N, 1000.0), source=np.zeros(N))
    got, ref = pair(params, g, f, False)
    for _ in range(10):   # run_sim: 10 updates
        assert got.run_sweeps() == ref.run_sweeps()
    compare(got, ref, 1e-9)
    # rate 0 with threshold 0: relative change is NaN -> 1/eps, every cell sits at the highest level
    assert np.all(got.levels() == 2)


def test_ragged_and_tiny_grids(cuda_lib):
    for shape_n, periodic, n_dirs in ((1, False, 16), (1, True, 16), (2, True, 21), (3, False, 84)):
        params, g, f = make_problem("cartesian", shape_n, periodic, n_dirs=n_dirs, n_levels=2, n_sources=1)
        got, ref = pair(params, g, f, periodic)
        for _ in range(3):
            got.run_sweeps()
            ref.run_sweeps()
        compare(got, ref, 1e-9)


def test_optional_chemistry_outputs(cuda_lib):
    params, g, f = make_problem("voronoi", 8, True, n_dirs=21, n_levels=2)
    got, ref = pair(params, g, f, True)
    for _ in range(3):
        got.run_sweeps()
        ref.run_sweeps()
    for name in ("photoionization_rate", "heating_rate", "recombination_rate", "collisional_ionization_rate"):
        a, b = got.read(name), ref.read(name)
        assert_close(a, b, 1e-6, floor=1e-9 * np.abs(b).max(), what=name)


def test_threshold_cuts_flux(cuda_lib):
    """incoming < significant_rate_threshold -> outgoing = 0 (hydrogen_only/mod.rs:81-82)."""
    params, g, f = make_problem("cartesian", 10, False, n_dirs=21, n_levels=1, threshold=1e40, nh_cm3=1e-3)
    got, ref = pair(params, g, f, False)
    got.run_sweeps()
    ref.run_sweeps()
    compare(got, ref, 1e-9)
    out = got.dir_state("outgoing")
    assert np.count_nonzero(out == 0.0) > 0.5 * out.size


def test_dependency_cycle_is_reported(cuda_lib):
    """A cycle among active Local faces would hang the reference's solve(); the library returns
    SSW_E_DEADLOCK instead (include/subsweep_b200.h)."""
    from subsweep_b200 import capi
    params, g, f = make_problem("cartesian", 4, True, n_dirs=1, n_levels=1)
    import dataclasses
    bad = dataclasses.replace(g, face_kind=np.where(g.face_kind == 2, 0, g.face_kind).astype(np.uint8))
    s = Sweep(params, bad, **f)
    with pytest.raises(capi.SubsweepError) as e:
        s.run_sweeps()
    assert e.value.code == capi.SSW_E_DEADLOCK


def test_invalid_grid_is_rejected(cuda_lib):
    from subsweep_b200 import capi
    import dataclasses
    params, g, f = make_problem("cartesian", 4, False, n_dirs=1)
    nb = g.face_neighbour.copy()
    nb[7] = g.n_cells + 5
    with pytest.raises(capi.SubsweepError) as e:
        Sweep(params, dataclasses.replace(g, face_neighbour=nb), **f)
    assert e.value.code == capi.SSW_E_INVALID
    nrm = g.face_normal.copy()
    nrm[1] = [0.0, 1.0, 0.0]   # no longer the negative of its reverse face
    with pytest.raises(capi.SubsweepError) as e:
        Sweep(params, dataclasses.replace(g, face_normal=nrm), **f)
    assert e.value.code == capi.SSW_E_INVALID
