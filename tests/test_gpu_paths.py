"""Every way the library can run a single sweep must give bit-identical results: fused
build+solve, cached replay of the level sets, the compiled slot-ordered path, and rebuilding the
level sets every time (SSW_FLAG_NO_SCHEDULE_CACHE)."""
import numpy as np
import pytest

from helpers import assert_close, make_problem
from subsweep_b200 import Sweep, capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,n,periodic", [("cartesian", 12, True), ("voronoi", 9, True), ("voronoi", 9, False)])
def test_all_paths_bitwise_identical(cuda_lib, kind, n, periodic):
    params, g, f = make_problem(kind, n, periodic, n_dirs=21, n_levels=3, max_timestep_myr=0.25)
    variants = {"default": 0, "no_cache": capi.FLAG_NO_SCHEDULE_CACHE, "no_compiled": capi.FLAG_NO_COMPILED_PATH,
                "no_patch": capi.FLAG_NO_PATCH_PATH}
    results = {}
    for name, flags in variants.items():
        s = Sweep(params, g, **f, flags=flags)
        for _ in range(6):
            s.run_sweeps()
        results[name] = {k: s.read(k) for k in ("ionized_hydrogen_fraction", "temperature", "change_timescale", "photon_rate")}
        results[name]["levels"] = s.levels()
        results[name]["outgoing"] = s.dir_state("outgoing")
        if name == "default":
            assert s.stat("schedule_replays") > 0
        if name == "no_cache":
            assert s.stat("schedule_replays") == 0
    # fused build+solve and cached replay do the same arithmetic per task: bit-identical
    for k, v in results["no_cache"].items():
        assert np.array_equal(v, results["no_compiled"][k], equal_nan=True), k
    # the compiled paths (patch-ordered dataflow where the grid admits it, level-barrier stream otherwise) fold
    # 1 / sum_downwind(A n.d) into their precomputed shares: round-off only
    for name in ("default", "no_patch"):
        for k, v in results[name].items():
            if k == "levels":
                assert np.array_equal(v, results["no_compiled"][k])
            else:
                assert_close(v, results["no_compiled"][k], 1e-11, floor=1e-7 * np.nanmax(np.abs(v)), what=k)


def test_sweep_plugin_surface(cuda_lib):
    """init_sweep_system / run_sweep_system on component arrays (src/sweep/mod.rs:634-739)."""
    from subsweep_b200 import SweepPlugin
    params, g, f = make_problem("cartesian", 8, True, n_dirs=16, n_levels=2)
    comps = {k: v.copy() for k, v in f.items()}
    plugin = SweepPlugin(params)
    plugin.init_sweep_system(g, comps)
    x0 = comps["ionized_hydrogen_fraction"].copy()
    plugin.run_sweep_system(comps)            # first call: no-op so the ICs get written (:711-714)
    assert np.array_equal(comps["ionized_hydrogen_fraction"], x0) and plugin.simulation_time == 0.0
    plugin.run_sweep_system(comps)
    assert plugin.simulation_time == params.max_timestep / 2.0      # warm-up: lowest allowed level 1
    assert comps["ionized_hydrogen_fraction"].max() > x0.max()
    assert comps["photon_rate"].max() > 0.0
    for _ in range(12):
        plugin.run_sweep_system(comps)
    ionized = comps["ionized_hydrogen_fraction"] > 0.5
    assert ionized.any()
    assert np.all(np.isfinite(comps["ionization_time"][ionized])) and np.all(np.isposinf(comps["ionization_time"][~ionized]))


def test_queued_read_back_equals_blocking_read(cuda_lib):
    """ssw_read_begin + ssw_sync (the five-component write-back with one synchronisation) against ssw_read."""
    params, g, f = make_problem("cartesian", 8, True, n_dirs=16, n_levels=2)
    s = Sweep(params, g, **f)
    for _ in range(3):
        s.run_sweeps()
    names = ("ionized_hydrogen_fraction", "temperature", "timestep", "photon_rate", "ionization_time")
    queued = {k: np.full(g.n_cells, np.nan) for k in names}
    for k in names:
        s.read_begin(k, queued[k])
    s.sync()
    for k in names:
        assert np.array_equal(queued[k], s.read(k), equal_nan=True), k


@pytest.mark.parametrize("kind,n,periodic,n_dirs", [("voronoi", 10, True, 84), ("voronoi", 10, False, 32), ("cartesian", 12, True, 84),
                                                    ("jittered", 9, True, 64)])
def test_walk_form_matches_stream_form(cuda_lib, monkeypatch, kind, n, periodic, n_dirs):
    """The two compiled forms of the all-cells sweep on grids without a patch form: one block per direction walking its
    wavefront with the recent rates in a shared-memory window (walk.cuh, SSW_WALK=1) against the level-barrier stream
    (stream.cuh, SSW_WALK=0).  Per task the arithmetic is the same: outgoing rates bit-identical; the per-cell rate is
    folded over directions in a different order (round-off)."""
    params, g, f = make_problem(kind, n, periodic, n_dirs=n_dirs, n_levels=3, max_timestep_myr=0.25)
    res = {}
    for form in ("0", "1"):
        monkeypatch.setenv("SSW_WALK", form)
        s = Sweep(params, g, **f, flags=capi.FLAG_NO_PATCH_PATH)
        s.run_sweeps()
        res[form, "first"] = s.dir_state("outgoing")
        for _ in range(5):
            s.run_sweeps()
        res[form] = {k: s.read(k) for k in ("ionized_hydrogen_fraction", "temperature", "change_timescale", "photon_rate")}
        res[form]["levels"] = s.levels()
        res[form]["outgoing"] = s.dir_state("outgoing")
        s.close()
    assert np.array_equal(res["0", "first"], res["1", "first"])          # same inputs: bit-identical outgoing rates
    assert np.array_equal(res["0"]["levels"], res["1"]["levels"])
    for k, v in res["0"].items():
        if k != "levels":
            assert_close(res["1"][k], v, 1e-11, floor=1e-7 * np.nanmax(np.abs(v)), what=k)


def test_timing_level_zero_times_steps_and_all_cells_kernels_only(cuda_lib):
    """ssw_set_timing_level(0): no per-phase event records; the step and the all-cells sweep kernel are still timed, and
    the results do not depend on the level."""
    params, g, f = make_problem("cartesian", 10, True, n_dirs=21, n_levels=2, max_timestep_myr=0.25)
    a, b = Sweep(params, g, **f), Sweep(params, g, **f)
    b.set_timing_level(0)
    for _ in range(4):
        a.run_sweeps()
        b.run_sweeps()
    ta, tb = a.timings(), b.timings()
    assert ta["sweep_ms"] > 0 and ta["chemistry_ms"] > 0 and ta["step_ms"] > 0
    assert tb["sweep_ms"] == 0 and tb["chemistry_ms"] == 0 and tb["update_levels_ms"] == 0
    assert tb["step_ms"] > 0 and tb["steps"] == 4 and tb["kernel_level_ms"][0] > 0
    assert tb["kernel_level_tasks"] == ta["kernel_level_tasks"]
    for k in ("ionized_hydrogen_fraction", "temperature", "photon_rate"):
        assert np.array_equal(a.read(k), b.read(k))
    with pytest.raises(Exception):
        b.set_timing_level(7)


def test_chem_attempts_read_back(cuda_lib):
    """ssw_read_chem_attempts: the substep attempts of every cell's last chemistry update.  After one step at one timestep
    level every cell was updated exactly once, so the per-cell counts add up to the library's attempt counter."""
    params, g, f = make_problem("cartesian", 10, False, n_dirs=21, n_levels=1, source_rate=1e54, max_timestep_myr=1.0)
    s = Sweep(params, g, **f)
    assert not s.chem_attempts().any()
    s.run_sweeps()
    a = s.chem_attempts()
    assert a.dtype == np.uint16 and a.shape == (g.n_cells,) and a.min() >= 1
    assert int(a.astype(np.int64).sum()) == s.stat("chem_attempts")
    assert a.max() > 1   # the cells next to the sources substep
