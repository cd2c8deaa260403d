"""Timestep levels and active sets on the device must be bit-exact given identical
change_timescale inputs (north_star; src/sweep/mod.rs:576-589, timestep_level.rs:27-36)."""
import numpy as np
import pytest

import oracle
from helpers import make_problem
from subsweep_b200 import Sweep

pytestmark = pytest.mark.gpu


def taus(n, seed, max_dt):
    rng = np.random.default_rng(seed)
    t = max_dt * 10.0 ** rng.uniform(-4, 3, n)
    special = np.array([0.0, np.inf, np.nan, -1.0, 1e-320, max_dt, max_dt * 10, max_dt / 0.1])
    t[:len(special)] = special
    # exact powers of two of the ratio max_dt / (0.1 * tau)
    k = np.arange(-6, 10)
    t[20:20 + len(k)] = max_dt / 0.1 / 2.0 ** k
    return t


@pytest.mark.parametrize("n_levels", [1, 2, 4, 7])
def test_level_update_bit_exact(cuda_lib, n_levels):
    params, g, f = make_problem(n=8, n_dirs=1, n_levels=n_levels, n_sources=0)
    tau = taus(g.n_cells, n_levels, params.max_timestep)
    ref = oracle.OracleSweep(params, g, **f)
    got = Sweep(params, g, **f)
    ref.set_change_timescale(tau)
    got.set_change_timescale(tau)
    ref.update_timestep_levels()
    got.update_timestep_levels()
    assert np.array_equal(got.levels(), ref.levels())
    assert np.array_equal(got.level_counts(), ref.level_counts())
    # levels never fall below the lowest allowed level (timestep_state.rs:54-64)
    assert got.levels().min() >= got.lowest_allowed_level() == n_levels - 1


def test_level_counts_after_set_levels(cuda_lib):
    params, g, f = make_problem(n=9, n_dirs=1, n_levels=4, n_sources=0)
    rng = np.random.default_rng(0)
    lv = rng.integers(0, 4, g.n_cells).astype(np.uint8)
    got = Sweep(params, g, **f)
    got.set_levels(lv)
    expect = np.array([(lv >= l).sum() for l in range(4)], dtype=np.uint64)
    assert np.array_equal(got.level_counts(), expect)
    assert np.array_equal(got.levels(), lv)
