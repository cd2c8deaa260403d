"""The CUDA path against the CPU oracle AT BASELINE.json's own configurations.

north_star: "Results must match the reference's own CPU sweep+chemistry on identical grids and inputs.  Wavefront
levels and active sets must be bit-exact.  Ionized fraction, temperature and rates must agree within relative 1e-9
after one step and 1e-6 after a full 1 Myr run."  The CPU side is oracle/ (the restatement of
src/sweep/mod.rs:258-289, 549-589), run on all host threads as direction shards (``run_sweeps_threads``: the same
task-queue algorithm per direction; directions never interact inside a sweep).

* config 2-C  128^3 Cartesian periodic box, log-normal density, 64 sources, 84 directions, 4 levels, to 1 Myr
              (compared with the oracle's lagged periodic reads; on this workload the reference's heap order reads the
              same values -- zero non-lagged reads, profiles/lag_study_cartesian32.jsonl) -- every
              cell compared;
* config 4    chemistry-stiff ionization front (dense neutral slab, 5e54 /s source, 21 directions, 4 levels) at 48^3:
              deep substepping on both sides, levels bit-exact at every step, substep-path flips counted;
* config 1    benches/sweep: the FULL 32^3 random-point Voronoi grid, non-periodic, 84 directions, 1 and 3 levels,
              against the reference's exact single-rank task order (HEAP).
"""
import os

import numpy as np
import pytest

import bench
import oracle
from helpers import assert_close
from subsweep_b200 import Sweep, SweepParameters, grid as G

pytestmark = pytest.mark.gpu

CELL_FIELDS = ("ionized_hydrogen_fraction", "temperature", "timestep", "change_timescale", "previous_rate", "photon_rate")
THREADS = len(os.sched_getaffinity(0))


def compare_cells(got, ref, rtol, what=""):
    for name in CELL_FIELDS:
        a, b = got.read(name), ref.read(name)
        # rates span > 40 decades across the box; below 1e-12 of the largest one they are compared absolutely
        floor = 1e-12 * np.nanmax(np.abs(b)) if name in ("previous_rate", "photon_rate") else 0.0
        assert_close(a, b, rtol, floor=floor, what=f"{what} {name}")


def assert_levels_equal(got, ref, what):
    lg, lr = got.levels(), ref.levels()
    n_diff = int(np.count_nonzero(lg != lr))
    assert n_diff == 0, f"{what}: {n_diff} of {lg.size} cells sit at a different timestep level"
    assert np.array_equal(got.level_counts(), ref.level_counts()), what
    assert got.lowest_allowed_level() == ref.lowest_allowed_level(), what


def test_config2_cartesian_128_to_one_myr(cuda_lib):
    """128^3 x 84 directions x 4 levels: 1e-9 after the first step, 1e-6 after the four calls that reach 1 Myr."""
    params, g, f = bench.build_workload(128, "cartesian", 84, 4)
    got = Sweep(params, g, **f)
    ref = oracle.OracleSweep(params, g, **f, periodic_mode=oracle.PERIODIC_LAGGED)
    elapsed_g = elapsed_r = 0.0
    for step in range(4):
        elapsed_g += got.run_sweeps()
        elapsed_r += ref.run_sweeps_threads(THREADS)
        assert_levels_equal(got, ref, f"step {step}")
        if step == 0:
            compare_cells(got, ref, 1e-9, "after one step:")
    assert elapsed_g == elapsed_r == params.max_timestep          # 1/8 + 1/8 + 1/4 + 1/2 Myr
    compare_cells(got, ref, 1e-6, "after 1 Myr:")
    assert got.stat("tasks_solved") == ref.stat("tasks_solved")
    assert got.stat("single_sweeps") == ref.stat("single_sweeps")
    assert got.stat("patch_macro_tiles") > 0, got.patch_note()   # the shipped form of the headline kernel ran
    x = got.read("ionized_hydrogen_fraction")
    assert x.max() > 1e-4                                         # the sources do ionize their surroundings
    got.close()
    ref.close()


@pytest.mark.parametrize("cell_scale", [1.0, 0.03])
def test_config4_ionization_front(cuda_lib, cell_scale):
    """Dense neutral slab + 5e54 /s source (BASELINE.json configs[3]) at 48^3.  cell_scale 1 is the config as SURVEY.md
    section 8d states it (78 kpc cells: the front stays in the thin gas in front of the slab, substep depth 41);
    cell_scale 0.03 shrinks the cells so that the same photons cross into the slab (up to ~8e6 substep attempts per
    step, depth 56).  Levels must be bit-exact at every step; a cell whose substep decision flipped (last-bit libm vs
    libdevice difference at the `relative change > safety` comparison) shows up as a different attempt count --
    counted and bounded."""
    params, g, f = bench.build_front_workload(48, 21, 4, cell_scale=cell_scale)
    got = Sweep(params, g, **f)
    ref = oracle.OracleSweep(params, g, **f, periodic_mode=oracle.PERIODIC_HEAP)
    flips = []
    for step in range(12):
        a0, b0 = got.stat("chem_attempts"), ref.stat("chem_attempts")
        got.run_sweeps()
        ref.run_sweeps_threads(THREADS)
        flips.append((got.stat("chem_attempts") - a0) - (ref.stat("chem_attempts") - b0))
        assert_levels_equal(got, ref, f"step {step}")
        compare_cells(got, ref, 1e-9 if step == 0 else 1e-6, f"step {step}:")
    assert got.stat("chem_max_depth") >= 20 and ref.stat("chem_max_depth") >= 20
    assert got.stat("chem_max_depth") == ref.stat("chem_max_depth")
    assert got.stat("chem_failures") == ref.stat("chem_failures") == 0
    attempts = ref.stat("chem_attempts")
    print(f"front 48^3 (cell scale {cell_scale}): {attempts} oracle substep attempts over 12 steps, max depth "
          f"{ref.stat('chem_max_depth')}, attempt-count difference per step {flips}")
    assert sum(abs(v) for v in flips) <= 1e-6 * attempts
    x = got.read("ionized_hydrogen_fraction").reshape(48, 48, 48)
    assert x.max() > 0.5 and x[36:].min() < 1e-3       # ionized around the source, neutral behind the slab
    if cell_scale < 1.0:
        assert x[12:36].max() > 0.3                    # the front is inside the slab
    got.close()
    ref.close()


@pytest.fixture(scope="module")
def voronoi_32():
    rng = np.random.default_rng(1338)          # benches/sweep/main.rs:116-126 (seed constant; numpy stream)
    pts = rng.uniform(0.0, 1e5, size=(32 ** 3, 3))
    return G.voronoi(pts, 1e5, periodic=False)


@pytest.mark.parametrize("n_levels,source", [(1, 0.0), (1, 1e50), (3, 1e50)])
def test_config1_voronoi_32_vs_heap_order(cuda_lib, voronoi_32, n_levels, source):
    """benches/sweep (src/sweep/mod.rs:783-797 components): 32^3 random points, non-periodic Voronoi, 84 directions.
    No periodic faces -> the oracle follows the reference's exact BinaryHeap task order."""
    g = voronoi_32
    N = g.n_cells
    params = SweepParameters(directions=84, num_timestep_levels=n_levels, periodic=False, max_timestep=1e-3,
                             significant_rate_threshold=0.0, timestep_safety_factor=0.1,
                             chemistry_timestep_safety_factor=0.1, prevent_cooling=False)
    src = np.zeros(N)
    if source:
        centre = np.argmin(((g.positions - 5e4) ** 2).sum(axis=1))
        src[centre] = source
    f = dict(density=np.full(N, 1e-10 / 1e-6), ionized_hydrogen_fraction=np.full(N, 1e-10),
             temperature=np.full(N, 1000.0), source=src)
    got = Sweep(params, g, **f)
    ref = oracle.OracleSweep(params, g, **f, periodic_mode=oracle.PERIODIC_HEAP)
    steps = 1 if n_levels == 1 else 5
    for step in range(steps):
        assert got.run_sweeps() == ref.run_sweeps_threads(THREADS)
        assert_levels_equal(got, ref, f"step {step}")
        if step == 0:
            compare_cells(got, ref, 1e-9, "after one step:")
            a, b = got.dir_state("outgoing"), ref.dir_state("outgoing")
            assert_close(a, b, 1e-9, floor=1e-12 * max(np.abs(b).max(), 1e-300), what="outgoing")
    compare_cells(got, ref, 1e-6, "final:")
    assert got.stat("tasks_solved") == ref.stat("tasks_solved")
    # wavefront level sets of the all-cells sweep, bit-exact (three directions)
    if n_levels == 1 and not source:
        for d in (0, 40, 83):
            assert np.array_equal(got.wavefront_levels(0, d), ref.wavefront_levels(0, d)), d
    got.close()
    ref.close()
