"""Flat-grid producers follow the reference's conventions (cartesian.rs, voronoi/cell.rs)."""
import numpy as np
import pytest

from subsweep_b200 import Directions, grid as G


def test_cartesian_conventions():
    g = G.cartesian((3, 4, 5), (3.0, 4.0, 5.0), periodic=False)
    assert g.n_cells == 60 and g.n_faces == 360
    # index = (x*ny + y)*nz + z ; faces -x,+x,-y,+y,-z,+z
    c = (1 * 4 + 2) * 5 + 3
    nb = g.face_neighbour[6 * c:6 * c + 6].tolist()
    assert nb == [(0 * 4 + 2) * 5 + 3, (2 * 4 + 2) * 5 + 3, (1 * 4 + 1) * 5 + 3, (1 * 4 + 3) * 5 + 3, c - 1, c + 1]
    assert np.allclose(g.face_normal[6 * c:6 * c + 6], [[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]])
    assert np.all(g.face_area == 1.0) and np.all(g.cell_volume == 1.0) and np.all(g.cell_size == 1.0)
    # corner cell: three boundary faces
    assert (g.face_kind[:6] == [G.FACE_BOUNDARY, G.FACE_LOCAL, G.FACE_BOUNDARY, G.FACE_LOCAL, G.FACE_BOUNDARY, G.FACE_LOCAL]).all()
    assert (g.face_neighbour[:6][[0, 2, 4]] == -1).all()


def test_cartesian_periodic_wrap():
    g = G.cartesian((4, 4, 4), 8.0, periodic=True)
    assert np.all(g.face_kind[:6] == [G.FACE_LOCAL_PERIODIC, G.FACE_LOCAL, G.FACE_LOCAL_PERIODIC, G.FACE_LOCAL, G.FACE_LOCAL_PERIODIC, G.FACE_LOCAL])
    assert g.face_neighbour[0] == (3 * 4 + 0) * 4 + 0
    # normals of wrap faces still point outwards (unwrapped neighbour position)
    assert np.allclose(g.face_normal[0], [-1, 0, 0])
    assert np.count_nonzero(g.face_kind == G.FACE_LOCAL_PERIODIC) == 6 * 16


@pytest.mark.parametrize("periodic", [True, False])
def test_voronoi_contract(periodic):
    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 2.0, size=(300, 3))
    g = G.voronoi(pts, 2.0, periodic)
    assert np.isclose(g.cell_volume.sum(), 8.0, rtol=1e-12)       # pyramids tile the box
    assert np.allclose(g.cell_size, np.cbrt(3 * g.cell_volume / (4 * np.pi)))
    assert np.allclose(np.linalg.norm(g.face_normal, axis=1), 1.0, atol=1e-14)
    fo = g.face_offsets.astype(np.int64)
    cell = np.repeat(np.arange(g.n_cells), np.diff(fo))
    # closed cells: sum of area * normal vanishes
    closure = np.zeros((g.n_cells, 3))
    np.add.at(closure, cell, g.face_area[:, None] * g.face_normal)
    assert np.abs(closure).max() < 1e-9 * g.face_area.max()
    # every non-boundary face has a reverse face with the opposite normal and the same area
    for f in range(g.n_faces):
        if g.face_kind[f] == G.FACE_BOUNDARY:
            assert g.face_neighbour[f] == -1
            continue
        nb = g.face_neighbour[f]
        cand = [h for h in range(fo[nb], fo[nb + 1]) if g.face_neighbour[h] == cell[f] and g.face_kind[h] == g.face_kind[f]]
        assert any(np.allclose(g.face_normal[h], -g.face_normal[f], atol=1e-12) and
                   np.isclose(g.face_area[h], g.face_area[f], rtol=1e-12) for h in cand)
    if not periodic:
        assert not np.any(g.face_kind == G.FACE_LOCAL_PERIODIC)
    d = Directions.from_num(84).xyz
    assert (6.5 if periodic else 5.0) < g.mean_upwind_faces(d) < 9.0   # 300 cells: many wrap faces


def test_tile_periodic():
    rng = np.random.default_rng(9)
    unit = G.voronoi(rng.uniform(0, 1.0, size=(120, 3)), 1.0, periodic=True)
    big = G.tile_periodic(unit, (2, 2, 2))
    assert big.n_cells == 8 * unit.n_cells and big.n_faces == 8 * unit.n_faces
    assert np.isclose(big.cell_volume.sum(), 8.0)
    # compare with tessellating the replicated point set directly
    pts = np.concatenate([unit.positions + np.array([bx, by, bz]) for bx in range(2) for by in range(2) for bz in range(2)])
    direct = G.voronoi(pts, 2.0, periodic=True)
    assert np.allclose(np.sort(direct.cell_volume), np.sort(big.cell_volume), rtol=1e-9)
    fo_b, fo_d = big.face_offsets.astype(np.int64), direct.face_offsets.astype(np.int64)
    for c in range(0, big.n_cells, 37):
        sb = sorted(zip(big.face_neighbour[fo_b[c]:fo_b[c + 1]].tolist(), big.face_kind[fo_b[c]:fo_b[c + 1]].tolist()))
        sd = sorted(zip(direct.face_neighbour[fo_d[c]:fo_d[c + 1]].tolist(), direct.face_kind[fo_d[c]:fo_d[c + 1]].tolist()))
        assert sb == sd
