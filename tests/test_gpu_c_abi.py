"""include/subsweep_b200.h compiled by a C compiler and linked with the library: tests/c_abi_smoke.c drives
create -> run -> read through the header itself; its result is checked against the oracle on the same grid."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def build_c_smoke() -> Path:
    import __graft_entry__ as entry
    return entry.build_c_abi_smoke()


def test_header_compiles_as_c_and_links():
    """No GPU needed: the header is valid C and every symbol the program uses resolves against the library."""
    exe = build_c_smoke()
    assert exe.exists()
    out = subprocess.run(["nm", "-u", str(exe)], capture_output=True, text=True).stdout
    assert "ssw_create" in out and "ssw_run_sweeps" in out and "ssw_read" in out


@pytest.mark.gpu
def test_c_program_runs_and_matches_the_oracle(cuda_lib):
    exe = build_c_smoke()
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    tag, mean, tasks = res.stdout.split()
    assert tag == "ok"
    # the same problem through the oracle
    import oracle
    from subsweep_b200 import SweepParameters, grid as G
    h = 3.0857e19 * 5.0
    g = G.cartesian((4, 4, 4), 4 * h, periodic=True)
    N = g.n_cells
    src = np.zeros(N)
    src[(1 * 4 + 2) * 4 + 3] = 1e51
    dirs = [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0.6, 0.48, 0.64], [-0.6, -0.48, -0.64]]
    params = SweepParameters(directions=dirs, num_timestep_levels=2, periodic=True, max_timestep=3.15576e12,
                             significant_rate_threshold=1e-5)
    ref = oracle.OracleSweep(params, g, np.full(N, 1e-4 * 1e6 * 1.67262192369e-27), np.full(N, 1e-10), np.full(N, 100.0), src,
                             periodic_mode=oracle.PERIODIC_LAGGED)
    for _ in range(3):
        ref.run_sweeps()
    assert abs(float(mean) - ref.read("ionized_hydrogen_fraction").mean()) <= 1e-9 * float(mean)
    assert int(tasks) == ref.stat("tasks_solved")
