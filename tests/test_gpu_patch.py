"""The patch-ordered all-cells sweep (patch.cuh: macro-tiles with point-to-point done flags, sub-levels walked
in shared memory) against the level-barrier stream, the generic CSR kernels and the CPU oracle.

Per task the patch kernel does the arithmetic of the stream kernel bit for bit (same shares, same summation
order); only the per-cell rate sums are grouped differently (direction groups by octant).  After ONE all-cells
sweep from the same state the outgoing rates must therefore be bit-identical, everything else agrees to
round-off."""
import numpy as np
import pytest

import oracle
from helpers import assert_close, make_problem
from subsweep_b200 import Sweep, capi

pytestmark = pytest.mark.gpu

FIELDS = ("ionized_hydrogen_fraction", "temperature", "change_timescale", "photon_rate", "previous_rate")


def run(params, g, f, flags, steps):
    s = Sweep(params, g, **f, flags=flags)
    for _ in range(steps):
        s.run_sweeps()
    out = {k: s.read(k) for k in FIELDS}
    out["levels"] = s.levels()
    for which in ("outgoing", "incoming", "periodic"):
        out[which] = s.dir_state(which)
    out["macro_tiles"] = s.stat("patch_macro_tiles")
    out["patch_levels"] = s.stat("patch_levels")
    out["wavefront_levels"] = s.stat("wavefront_levels")
    out["note"] = s.patch_note()
    out["phases"] = s.stat("patch_phases")
    s.close()
    return out


@pytest.mark.parametrize("n,periodic,n_dirs,patch_cells", [
    (12, True, 84, 512), (12, False, 84, 512), (16, True, 21, 64), (13, True, 84, 27), (9, False, 1, 27), (10, True, 16, 8),
])
def test_patch_form_matches_stream_form_on_cartesian_grids(cuda_lib, monkeypatch, n, periodic, n_dirs, patch_cells):
    monkeypatch.setenv("SSW_PATCH_CELLS", str(patch_cells))
    params, g, f = make_problem("cartesian", n, periodic, n_dirs=n_dirs, n_levels=1)
    # two steps: the first all-cells sweep is the fused build, the second one runs the compiled form
    a = run(params, g, f, 0, 2)
    b = run(params, g, f, capi.FLAG_NO_PATCH_PATH, 2)
    assert a["macro_tiles"] > 0 and a["note"] == "", a["note"]
    assert b["macro_tiles"] == 0
    assert a["patch_levels"] <= a["wavefront_levels"]
    assert np.array_equal(a["outgoing"], b["outgoing"])          # same arithmetic per task
    assert np.array_equal(a["levels"], b["levels"])
    for k in FIELDS + ("incoming", "periodic"):
        assert_close(a[k], b[k], 1e-12, floor=1e-9 * max(np.nanmax(np.abs(b[k])), 1e-300), what=k)


@pytest.mark.parametrize("threads,kd", [(32, 1), (32, 3), (64, 2), (256, 11)])
def test_patch_kernel_variants_are_bit_identical(cuda_lib, monkeypatch, threads, kd):
    """Block size and direction-group width of the patch kernel are tunables (SSW_PATCH_THREADS / SSW_PATCH_KD; 32 threads:
    one warp per macro-tile, warp barriers only).  Per task the arithmetic does not depend on them."""
    params, g, f = make_problem("cartesian", 12, True, n_dirs=84, n_levels=1)
    monkeypatch.setenv("SSW_PATCH_CELLS", "64")
    a = run(params, g, f, 0, 2)
    monkeypatch.setenv("SSW_PATCH_THREADS", str(threads))
    monkeypatch.setenv("SSW_PATCH_KD", str(kd))
    b = run(params, g, f, 0, 2)
    assert a["macro_tiles"] > 0 and b["macro_tiles"] > 0 and b["note"] == "", b["note"]
    assert np.array_equal(a["outgoing"], b["outgoing"])
    for k in FIELDS + ("incoming", "periodic"):
        assert_close(a[k], b[k], 1e-12, floor=1e-9 * max(np.nanmax(np.abs(b[k])), 1e-300), what=k)


@pytest.mark.parametrize("kind,n,periodic", [("cartesian", 12, True), ("cartesian", 11, False), ("voronoi", 9, True),
                                             ("voronoi", 9, False), ("jittered", 8, True)])
def test_patch_default_against_oracle_with_timestep_levels(cuda_lib, monkeypatch, kind, n, periodic):
    monkeypatch.setenv("SSW_PATCH_CELLS", "64")
    params, g, f = make_problem(kind, n, periodic, n_dirs=21, n_levels=3, max_timestep_myr=0.25)
    mode = oracle.PERIODIC_LAGGED if periodic else oracle.PERIODIC_HEAP
    got, ref = Sweep(params, g, **f), oracle.OracleSweep(params, g, **f, periodic_mode=mode)
    for _ in range(6):
        assert got.run_sweeps() == ref.run_sweeps()
    if kind == "cartesian":
        assert got.stat("patch_macro_tiles") > 0, got.patch_note()
    else:   # a Voronoi grid may or may not admit the patch form; either way the answer must not change
        assert got.stat("patch_macro_tiles") > 0 or got.patch_note() != ""
    for name in ("ionized_hydrogen_fraction", "temperature", "change_timescale"):
        assert_close(got.read(name), ref.read(name), 1e-6, what=name)
    b = ref.read("photon_rate")
    assert_close(got.read("photon_rate"), b, 1e-6, floor=1e-7 * np.nanmax(np.abs(b)), what="photon_rate")
    assert np.array_equal(got.levels(), ref.levels())
    assert got.stat("tasks_solved") == ref.stat("tasks_solved")


@pytest.mark.parametrize("kind,n,periodic,n_dirs,patch_cells", [
    ("voronoi", 9, True, 84, 27), ("voronoi", 9, False, 21, 64), ("jittered", 8, True, 16, 27), ("voronoi", 10, True, 21, 125),
])
def test_patch_form_with_phases_matches_stream_form_on_voronoi_grids(cuda_lib, monkeypatch, kind, n, periodic, n_dirs, patch_cells):
    """Voronoi grids have cyclic patch graphs; the macro-tiles are then split into phases (patch.cuh, p_phase_kernel).
    Same per-task arithmetic as the stream form: outgoing rates bit-identical after a sweep from the same state."""
    monkeypatch.setenv("SSW_PATCH_CELLS", str(patch_cells))
    monkeypatch.setenv("SSW_PATCH_PHASES", "1")
    params, g, f = make_problem(kind, n, periodic, n_dirs=n_dirs, n_levels=1)
    a = run(params, g, f, 0, 2)
    b = run(params, g, f, capi.FLAG_NO_PATCH_PATH, 2)
    assert a["macro_tiles"] > 0 and a["note"] == "", a["note"]
    assert b["macro_tiles"] == 0
    assert np.array_equal(a["outgoing"], b["outgoing"])
    assert np.array_equal(a["levels"], b["levels"])
    for k in FIELDS + ("incoming", "periodic"):
        assert_close(a[k], b[k], 1e-12, floor=1e-9 * max(np.nanmax(np.abs(b[k])), 1e-300), what=k)
