"""The CUDA path against the PYTHON restatement of the reference (tests/test_oracle_double_entry.py: run_sweeps typed from
the reference source, Rust's BinaryHeap included) -- a checker that shares no code with oracle/."""
import numpy as np
import pytest

from helpers import assert_close, make_problem
from subsweep_b200 import Sweep
from test_oracle_double_entry import PySweep

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,n,periodic,n_dirs,n_levels", [("jittered", 4, False, 16, 3), ("voronoi", 3, True, 16, 2)])
def test_cuda_matches_the_python_restatement(cuda_lib, kind, n, periodic, n_dirs, n_levels):
    params, g, f = make_problem(kind, n, periodic, n_dirs=n_dirs, n_levels=n_levels, source_rate=3e51, max_timestep_myr=0.5)
    # periodic grids: the order-independent (lagged) definition of the periodic reads, DESIGN.md section 4
    ref = PySweep(params, g, **f, lagged=periodic)
    got = Sweep(params, g, **f)
    for step in range(n_levels + 2):
        assert got.run_sweeps() == ref.run_sweeps()
        assert np.array_equal(got.levels(), ref.level), step
        for name, want in (("ionized_hydrogen_fraction", ref.x), ("temperature", ref.T), ("timestep", ref.ts),
                           ("change_timescale", ref.tau), ("previous_rate", ref.prev)):
            floor = 1e-12 * np.nanmax(np.abs(want)) if name == "previous_rate" else 0.0
            assert_close(got.read(name), want, 1e-8, floor=floor, what=f"{name} step {step}")
        want = ref.out
        assert_close(got.dir_state("outgoing"), want, 1e-8, floor=1e-12 * max(np.abs(want).max(), 1e-300), what=f"outgoing step {step}")
