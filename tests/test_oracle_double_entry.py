"""Double entry for the sweep: single_sweep typed a second time in plain Python straight from the reference
(src/sweep/mod.rs:346-513 scatter form, site.rs:49-56, hydrogen_only/mod.rs:72-88, update_chemistry's rate fold
:554-558) and compared with the C restatement (oracle/) that the GPU parity tests trust.  Small non-periodic grids:
without periodic faces the result does not depend on the order in which ready tasks are taken (only the rounding of
the `+=` chains does), so the Python side uses a plain work list where the C side follows the BinaryHeap order."""
import math

import numpy as np
import pytest

import oracle
from helpers import make_problem
from subsweep_b200 import units as U
from subsweep_b200.sweep import Directions

SIGMA = 2.9580524545305314e-18 * 0.01 * 0.01   # src/units/mod.rs: number weighted cross section, cm^2 -> m^2


def py_single_sweep(g, dirs, rho, x, src, threshold):
    N, D = g.n_cells, len(dirs)
    off = g.face_offsets.astype(np.int64)
    inc = np.zeros((N, D))
    out = np.zeros((N, D))
    rate = np.zeros(N)
    for d in range(D):
        dx, dy, dz = dirs[d]
        dot = [g.face_normal[f, 0] * dx + g.face_normal[f, 1] * dy + g.face_normal[f, 2] * dz for f in range(g.n_faces)]
        miss = [sum(1 for f in range(off[c], off[c + 1]) if dot[f] < 0.0 and g.face_kind[f] == 0) for c in range(N)]   # :346-386
        todo = [c for c in range(N) if miss[c] == 0]                                                                    # :388-398
        solved = 0
        while todo:
            c = todo.pop()
            solved += 1
            inc[c, d] = max(inc[c, d], 0.0)                                                      # :418
            total = (inc[c, d] + src[c] / D) + 0.0                                               # site.rs:53-56 (no periodic faces)
            nhi = rho[c] / U.PROTON_MASS * (1.0 - x[c])
            o = 0.0 if total < threshold else total * math.exp(-nhi * SIGMA * g.cell_size[c])  # hydrogen_only/mod.rs:78-86
            delta = o - out[c, d]
            out[c, d] = o
            ttot = 0.0
            for f in range(off[c], off[c + 1]):                                                  # :453-461
                if dot[f] > 0.0:
                    ttot += g.face_area[f] * dot[f]
            for f in range(off[c], off[c + 1]):
                if dot[f] > 0.0:
                    share = delta * ((g.face_area[f] * dot[f]) / ttot)
                    if g.face_kind[f] == 0:                                                       # handle_local_neighbour :487-503
                        nb = g.face_neighbour[f]
                        inc[nb, d] += share
                        miss[nb] -= 1
                        if miss[nb] == 0:
                            todo.append(nb)
        assert solved == N   # every task ran exactly once
    for c in range(N):                                                                           # rate fold :554-558
        acc = 0.0
        for d in range(D):
            acc = acc + ((inc[c, d] + src[c] / D) + 0.0)
        rate[c] = acc
    return out, inc, rate


@pytest.mark.parametrize("kind,n,n_dirs", [("cartesian", 5, 16), ("voronoi", 4, 21), ("jittered", 4, 16)])
def test_python_single_sweep_equals_the_c_oracle(kind, n, n_dirs):
    params, g, f = make_problem(kind, n, False, n_dirs=n_dirs, n_levels=1, source_rate=1e52)
    dirs = Directions.from_num(n_dirs).xyz
    out, inc, rate = py_single_sweep(g, dirs, f["density"], f["ionized_hydrogen_fraction"], f["source"],
                                     params.significant_rate_threshold)
    s = oracle.OracleSweep(params, g, **f, periodic_mode=oracle.PERIODIC_HEAP)
    s.single_sweep(0)
    for name, mine in (("outgoing", out), ("incoming", inc)):
        ref = s.dir_state(name)
        assert ref.shape == mine.shape
        scale = np.abs(ref).max()
        assert scale > 0
        np.testing.assert_allclose(mine, ref, rtol=1e-11, atol=1e-13 * scale, err_msg=name)
    ref_rate = s.read("previous_rate")
    np.testing.assert_allclose(rate, ref_rate, rtol=1e-11, atol=1e-13 * np.abs(ref_rate).max())
