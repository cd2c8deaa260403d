"""Parity at BASELINE.json's full size (128^3 cells x 84 directions), where the CPU oracle would take
minutes: size-independent properties instead of a cell-by-cell oracle comparison.

* wavefront level sets have a closed form on a Cartesian grid: level(c, d) = sum over the axes of the
  cell index counted from the side the direction enters (periodic wrap faces carry no dependency,
  src/sweep/mod.rs:346-386) -- bit-exact;
* the two independent implementations of the sweep (the TMA-streamed compiled schedule and the
  generic CSR gather kernels) must agree to round-off after several steps with timestep levels;
* the reference's invariants of a finished sweep: every task solved exactly once
  (debug_assert in src/sweep/mod.rs:286-288), cumulative level counts consistent.
"""
import numpy as np
import pytest

from helpers import assert_close
from subsweep_b200 import Sweep, SweepParameters, capi, grid as G
from subsweep_b200 import units as U

pytestmark = pytest.mark.gpu

N = 128


@pytest.fixture(scope="module")
def box():
    cell = 10.0 * U.MEGAPARSEC / 128.0
    g = G.cartesian((N, N, N), cell * N, periodic=True)
    rho = G.lognormal_density((N, N, N), 1e-3 * U.PER_CUBIC_CENTIMETER * U.PROTON_MASS, sigma_g=1.0, smooth_cells=4.0, seed=2024)
    src = np.zeros(g.n_cells)
    src[np.argsort(rho)[-64:]] = 1e52
    f = dict(density=np.ascontiguousarray(rho), ionized_hydrogen_fraction=np.full(g.n_cells, 1e-10),
             temperature=np.full(g.n_cells, 100.0), source=src)
    params = SweepParameters(directions=84, num_timestep_levels=3, periodic=True, max_timestep=1.0 * U.MEGAYEARS,
                             significant_rate_threshold=1e-5)
    return params, g, f


def test_wavefront_levels_closed_form(cuda_lib, box):
    params, g, f = box
    s = Sweep(params, g, **f)
    i, j, k = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    for d in (0, 41, 83):
        dx, dy, dz = s.directions.xyz[d]
        expect = ((i if dx > 0 else N - 1 - i) if dx != 0 else 0) + ((j if dy > 0 else N - 1 - j) if dy != 0 else 0) \
            + ((k if dz > 0 else N - 1 - k) if dz != 0 else 0)
        got = s.wavefront_levels(params.num_timestep_levels - 1, d).reshape(N, N, N)
        assert np.array_equal(got, expect), d
    s.close()


def test_compiled_and_generic_paths_agree_at_full_size(cuda_lib, box, monkeypatch):
    params, g, f = box
    results = {}
    for name, flags in (("compiled", 0), ("stream", capi.FLAG_NO_PATCH_PATH), ("generic", capi.FLAG_NO_COMPILED_PATH)):
        s = Sweep(params, g, **f, flags=flags)
        for _ in range(5):
            s.run_sweeps()
        results[name] = {k: s.read(k) for k in ("ionized_hydrogen_fraction", "temperature", "change_timescale", "photon_rate")}
        results[name]["levels"] = s.levels()
        results[name]["counts"] = s.level_counts()
        # every task of every single sweep solved exactly once: 1 + 1 + 2 all-cells sweeps + sub-level sweeps
        counts = s.level_counts()
        assert counts[0] == g.n_cells and np.all(np.diff(counts.astype(np.int64)) <= 0)
        assert s.stat("tasks_solved") >= 5 * g.n_cells * 84
        assert s.stat("chem_failures") == 0
        # the default is the patch-ordered dataflow: 46 dependent macro-tile levels instead of 382 wavefront levels
        assert (s.stat("patch_macro_tiles") > 0) == (name == "compiled"), s.patch_note()
        if name == "compiled":
            assert s.stat("patch_levels") == 76     # 26 patches of 4-5 cells per axis: 3 * 25 + 1
        s.close()
    b = results["generic"]
    for a in (results["compiled"], results["stream"]):
        assert np.array_equal(a["levels"], b["levels"]) and np.array_equal(a["counts"], b["counts"])
        for k in ("ionized_hydrogen_fraction", "temperature", "change_timescale", "photon_rate"):
            assert_close(a[k], b[k], 1e-10, floor=1e-7 * np.nanmax(np.abs(b[k])), what=k)
    a = results["compiled"]
    x = a["ionized_hydrogen_fraction"]
    assert 1e-10 <= x.min() and x.max() <= 1.0 - 1e-10 and x.max() > 1e-4     # the sources start to ionize their cells (78 kpc cells: slowly)
