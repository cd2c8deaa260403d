"""Wavefront level sets built on the device (Kahn peeling of init_counts / handle_local_neighbour,
src/sweep/mod.rs:346-398, 487-503) must be bit-exact against the oracle."""
import numpy as np
import pytest

import oracle
from helpers import make_problem
from subsweep_b200 import Sweep

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,n,periodic", [("cartesian", 9, False), ("cartesian", 8, True),
                                             ("voronoi", 8, True), ("voronoi", 8, False)])
def test_wavefront_levels_all_active(cuda_lib, kind, n, periodic):
    params, g, f = make_problem(kind, n, periodic, n_dirs=84, n_levels=1)
    ref = oracle.OracleSweep(params, g, **f)
    got = Sweep(params, g, **f)
    for d in (0, 1, 17, 40, 83):
        a, b = got.wavefront_levels(0, d), ref.wavefront_levels(0, d)
        assert np.array_equal(a, b), (kind, d)
        assert a.min() == 0


def test_wavefront_levels_axis_aligned(cuda_lib):
    """d = (1,0,0): lateral faces have n.d = +0.0 and are neither upwind nor downwind
    (strict comparisons; reference test sweep_along_grid_axes...)."""
    params, g, f = make_problem("cartesian", 7, False, n_dirs=1, n_levels=1)
    ref = oracle.OracleSweep(params, g, **f)
    got = Sweep(params, g, **f)
    a = got.wavefront_levels(0, 0)
    assert np.array_equal(a, ref.wavefront_levels(0, 0))
    assert np.array_equal(a.reshape(7, 7, 7), np.broadcast_to(np.arange(7)[:, None, None], (7, 7, 7)))


@pytest.mark.parametrize("kind,n,periodic", [("cartesian", 10, True), ("voronoi", 8, True)])
def test_wavefront_levels_partial_active_sets(cuda_lib, kind, n, periodic):
    params, g, f = make_problem(kind, n, periodic, n_dirs=21, n_levels=4)
    rng = np.random.default_rng(4)
    lv = rng.choice(4, size=g.n_cells, p=[0.55, 0.2, 0.15, 0.1]).astype(np.uint8)
    ref = oracle.OracleSweep(params, g, **f)
    got = Sweep(params, g, **f)
    ref.set_levels(lv)
    got.set_levels(lv)
    for cur in (0, 1, 2, 3):
        for d in (0, 5, 20):
            a, b = got.wavefront_levels(cur, d), ref.wavefront_levels(cur, d)
            assert np.array_equal(a, b), (cur, d)
            assert np.array_equal(a >= 0, lv >= cur)    # the active set itself
