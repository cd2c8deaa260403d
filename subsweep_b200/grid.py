"""Flat (CSR) sweep grids -- host-side preprocessing.

The reference builds a per-particle ``Cell { neighbours: Vec<(Face, ParticleType)>, size, volume }``
component (src/sweep/grid/cell.rs:92-133) either from its Voronoi constructor
(src/voronoi/constructor/mod.rs:138-169) or from the Cartesian test grid
(src/sweep/grid/cartesian.rs).  The CUDA library takes the same information flattened:
``face_offsets[N+1]`` and, per face, area, outward unit normal, neighbour index and kind.

Grid construction stays on the host (north_star: "Voronoi construction and grid read-in stay
host-side preprocessing"); this module provides the two producers the tests and the benchmark
need, following the reference's conventions so that oracle and CUDA library see exactly what
the reference's solver would see.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FACE_LOCAL, FACE_BOUNDARY, FACE_LOCAL_PERIODIC = 0, 1, 2


@dataclass
class FlatGrid:
    face_offsets: np.ndarray   # uint64 [N+1]
    face_area: np.ndarray      # float64 [F]
    face_normal: np.ndarray    # float64 [F,3]
    face_neighbour: np.ndarray # int32 [F], -1 = boundary
    face_kind: np.ndarray      # uint8 [F]
    cell_size: np.ndarray      # float64 [N]
    cell_volume: np.ndarray    # float64 [N]
    positions: np.ndarray      # float64 [N,3] generator points (not used by the solver)
    box: np.ndarray = field(default_factory=lambda: np.ones(3))

    @property
    def n_cells(self) -> int:
        return len(self.cell_size)

    @property
    def n_faces(self) -> int:
        return len(self.face_area)

    def validate(self) -> None:
        N = self.n_cells
        assert self.face_offsets.dtype == np.uint64 and len(self.face_offsets) == N + 1
        assert self.face_offsets[0] == 0 and self.face_offsets[-1] == self.n_faces
        assert self.face_normal.shape == (self.n_faces, 3)
        assert self.face_neighbour.dtype == np.int32 and self.face_kind.dtype == np.uint8
        nb = self.face_neighbour
        assert np.all((nb >= 0) | (self.face_kind == FACE_BOUNDARY))
        assert np.all(nb < N)

    def faces_per_cell(self) -> np.ndarray:
        return np.diff(self.face_offsets.astype(np.int64))

    def mean_upwind_faces(self, dirs: np.ndarray) -> float:
        """F_up: mean number of flux-carrying (Local or periodic, n.d < 0) faces per (cell, dir)."""
        carrying = self.face_kind != FACE_BOUNDARY
        n = self.face_normal[carrying]
        total = 0
        for d in np.asarray(dirs, dtype=np.float64):
            total += int(np.count_nonzero(n @ d < 0.0))
        return total / (self.n_cells * len(dirs))


def _normalize(v: np.ndarray) -> np.ndarray:
    # glam DVec3::normalize: v * (1 / sqrt(x*x + y*y + z*z))
    length = np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1] + v[..., 2] * v[..., 2])
    return v * (1.0 / length)[..., None]


def cartesian(shape, box_size, periodic: bool) -> FlatGrid:
    """Regular grid following src/sweep/grid/cartesian.rs.

    Cell index = (x * ny + y) * nz + z (iter_all_contained, :182-194); faces in the order
    -x, +x, -y, +y, -z, +z (:205-213); position = side * i / n (:161-178); normal =
    normalize(neighbour_pos - pos) with the *unwrapped* neighbour position (:279-283);
    area = h^2, size = h, volume = h^3 (:56-76); out-of-box neighbours are LocalPeriodic when
    ``periodic`` else Boundary (:315-348).
    """
    nx, ny, nz = (int(s) for s in shape)
    box = np.broadcast_to(np.asarray(box_size, dtype=np.float64), (3,)).copy()
    h = box[0] / nx
    assert np.allclose(box / np.array([nx, ny, nz]), h), "cartesian.rs uses one cell size"
    N = nx * ny * nz
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ipos = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.int64)
    n_arr = np.array([nx, ny, nz], dtype=np.int64)

    def to_pos(ip):
        return box[None, :] * ip.astype(np.float64) / n_arr[None, :].astype(np.float64)

    pos = to_pos(ipos)
    offsets = np.array([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]], dtype=np.int64)
    F = 6 * N
    normal = np.empty((N, 6, 3))
    nb = np.empty((N, 6), dtype=np.int32)
    kind = np.empty((N, 6), dtype=np.uint8)
    for k, o in enumerate(offsets):
        nip = ipos + o[None, :]
        normal[:, k, :] = _normalize(to_pos(nip) - pos)
        outside = np.any((nip < 0) | (nip >= n_arr[None, :]), axis=1)
        w = np.mod(nip, n_arr[None, :])
        idx = (w[:, 0] * ny + w[:, 1]) * nz + w[:, 2]
        if periodic:
            nb[:, k] = idx
            kind[:, k] = np.where(outside, FACE_LOCAL_PERIODIC, FACE_LOCAL)
        else:
            nb[:, k] = np.where(outside, -1, idx)
            kind[:, k] = np.where(outside, FACE_BOUNDARY, FACE_LOCAL)
    g = FlatGrid(
        face_offsets=(np.arange(N + 1, dtype=np.uint64) * np.uint64(6)),
        face_area=np.full(F, h ** 2),
        face_normal=np.ascontiguousarray(normal.reshape(F, 3)),
        face_neighbour=np.ascontiguousarray(nb.reshape(F)),
        face_kind=np.ascontiguousarray(kind.reshape(F)),
        cell_size=np.full(N, h),
        cell_volume=np.full(N, h ** 3),
        positions=pos + 0.0,
        box=box,
    )
    g.validate()
    return g


def _polygon_area(pts: np.ndarray) -> float:
    # src/voronoi/primitives/polygon3d.rs:10-16: fan from points[0] over periodic windows
    r = pts[0]
    a = r[None, :] - pts
    b = r[None, :] - np.roll(pts, -1, axis=0)
    return float(np.sum(0.5 * np.linalg.norm(np.cross(a, b), axis=1)))


def voronoi(points: np.ndarray, box_size, periodic: bool, pad_cells: float = 3.0) -> FlatGrid:
    """Periodic Voronoi tessellation of ``points`` in the box [0, L)^3 via scipy/Qhull.

    Follows the reference's output contract (src/voronoi/cell.rs:63-71, 160-243): face normal =
    normalize(p2 - p1) between the two generators; face area = fan-triangulated polygon area;
    volume = sum of the pyramids (generator, face); size = (3 V / 4 pi)^(1/3).  Like the
    reference's constructor the box is always tessellated periodically (images are imported,
    src/voronoi/constructor/parallel/mod.rs:63-79) and ``periodic = False`` only relabels the wrap
    faces as Boundary (map_ptype, src/voronoi/constructor/mod.rs:125-135).
    """
    from scipy.spatial import Voronoi

    pts = np.asarray(points, dtype=np.float64)
    N = len(pts)
    box = np.broadcast_to(np.asarray(box_size, dtype=np.float64), (3,)).copy()
    pad = pad_cells * (np.prod(box) / N) ** (1.0 / 3.0)
    pad = float(min(pad, 0.999 * box.min()))
    all_pts = [pts]
    origin = [np.arange(N)]
    for sx in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sz in (-1, 0, 1):
                if sx == sy == sz == 0:
                    continue
                shifted = pts + np.array([sx, sy, sz]) * box
                keep = np.all((shifted > -pad) & (shifted < box + pad), axis=1)
                all_pts.append(shifted[keep])
                origin.append(np.nonzero(keep)[0])
    all_pts = np.concatenate(all_pts)
    origin = np.concatenate(origin)
    vor = Voronoi(all_pts)
    verts = vor.vertices
    rp = vor.ridge_points
    # per cell face lists
    faces = [[] for _ in range(N)]
    for (a, b), rv in zip(rp, vor.ridge_vertices):
        if a >= N and b >= N:
            continue
        if -1 in rv:
            raise RuntimeError("unbounded ridge on a primary cell: increase pad_cells")
        poly = verts[np.asarray(rv)]
        area = _polygon_area(poly)
        for p1, p2 in ((a, b), (b, a)):
            if p1 >= N:
                continue
            nrm = _normalize((all_pts[p2] - all_pts[p1])[None, :])[0]
            height = abs(float(nrm @ (all_pts[p1] - poly[0])))
            vol = 1.0 / 3.0 * area * height
            is_image = p2 >= N
            faces[p1].append((area, nrm, int(origin[p2]), is_image, vol))
    counts = np.array([len(f) for f in faces], dtype=np.int64)
    offs = np.zeros(N + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(counts)
    F = int(offs[-1])
    area = np.empty(F)
    normal = np.empty((F, 3))
    nb = np.empty(F, dtype=np.int32)
    kind = np.empty(F, dtype=np.uint8)
    volume = np.zeros(N)
    k = 0
    for c in range(N):
        # deterministic face order: by neighbour index, images after primaries
        for (a, nrm, j, is_image, vol) in sorted(faces[c], key=lambda t: (t[3], t[2], tuple(t[1]))):
            area[k] = a
            normal[k] = nrm
            if is_image:
                nb[k] = j if periodic else -1
                kind[k] = FACE_LOCAL_PERIODIC if periodic else FACE_BOUNDARY
            else:
                nb[k] = j
                kind[k] = FACE_LOCAL
            volume[c] += vol
            k += 1
    size = np.cbrt(3.0 * volume / (4.0 * np.pi))
    g = FlatGrid(offs, area, normal, nb, kind, size, volume, pts.copy(), box)
    g.validate()
    return g


def tile_periodic(unit: FlatGrid, reps) -> FlatGrid:
    """Repeat a *periodic* unit grid reps = (rx, ry, rz) times into one larger periodic grid.

    Wrap faces of the unit block become Local faces to the adjacent block (or stay
    LocalPeriodic across the outer boundary).  Used to build 128^3-cell Voronoi boxes out of
    one small Qhull tessellation.  Cell index = block * N_unit + local index.
    """
    rx, ry, rz = (int(r) for r in reps)
    Nu = unit.n_cells
    box = unit.box
    fo = unit.face_offsets.astype(np.int64)
    cell_of_face = np.repeat(np.arange(Nu), np.diff(fo))
    # which wrap does a periodic face cross?  neighbour image position = own pos + normal * dist;
    # decide by comparing the generator displacement with the box
    p = unit.positions
    nbp = p[np.clip(unit.face_neighbour, 0, None)]
    disp = nbp - p[cell_of_face]
    # a periodic face points towards the image: pick the shift s in {-1,0,1}^3 that makes
    # (disp + s*box) most parallel to the normal
    shifts = np.array([[sx, sy, sz] for sx in (-1, 0, 1) for sy in (-1, 0, 1) for sz in (-1, 0, 1)])
    is_per = unit.face_kind == FACE_LOCAL_PERIODIC
    wrap = np.zeros((unit.n_faces, 3), dtype=np.int64)
    if np.any(is_per):
        cand = disp[is_per][:, None, :] + shifts[None, :, :] * box[None, None, :]
        cand_n = cand / np.linalg.norm(cand, axis=2, keepdims=True)
        score = np.einsum("fsk,fk->fs", cand_n, unit.face_normal[is_per])
        score[:, 13] = -2.0  # s = 0 is not a wrap
        wrap[is_per] = shifts[np.argmax(score, axis=1)]
    blocks = [(bx, by, bz) for bx in range(rx) for by in range(ry) for bz in range(rz)]
    B = len(blocks)
    r_arr = np.array([rx, ry, rz])
    offs = np.concatenate([[0], np.cumsum(np.tile(np.diff(fo), B))]).astype(np.uint64)
    area = np.tile(unit.face_area, B)
    normal = np.tile(unit.face_normal, (B, 1))
    size = np.tile(unit.cell_size, B)
    volume = np.tile(unit.cell_volume, B)
    nb = np.empty(B * unit.n_faces, dtype=np.int32)
    kind = np.empty(B * unit.n_faces, dtype=np.uint8)
    pos = np.empty((B * Nu, 3))
    for bi, b in enumerate(blocks):
        b = np.array(b)
        tb = b[None, :] + wrap                       # target block per face
        crosses_outer = np.any((tb < 0) | (tb >= r_arr[None, :]), axis=1)
        tbw = np.mod(tb, r_arr[None, :])
        tbi = (tbw[:, 0] * ry + tbw[:, 1]) * rz + tbw[:, 2]
        sl = slice(bi * unit.n_faces, (bi + 1) * unit.n_faces)
        local_nb = unit.face_neighbour.astype(np.int64)
        nb[sl] = np.where(unit.face_kind == FACE_BOUNDARY, -1, tbi * Nu + local_nb)
        k = unit.face_kind.copy()
        k[is_per & ~crosses_outer] = FACE_LOCAL
        kind[sl] = k
        pos[bi * Nu:(bi + 1) * Nu] = p + b[None, :] * box[None, :]
    g = FlatGrid(offs, area, normal, nb, kind, size, volume, pos, box * r_arr)
    g.validate()
    return g


def relabel_nonperiodic(g: FlatGrid) -> FlatGrid:
    """map_ptype (src/voronoi/constructor/mod.rs:125-135): wrap faces become Boundary."""
    kind = g.face_kind.copy()
    nb = g.face_neighbour.copy()
    per = kind == FACE_LOCAL_PERIODIC
    kind[per] = FACE_BOUNDARY
    nb[per] = -1
    return FlatGrid(g.face_offsets, g.face_area, g.face_normal, nb, kind, g.cell_size, g.cell_volume,
                    g.positions, g.box)


def lognormal_density(shape, mean_density: float, sigma_g: float = 1.0, smooth_cells: float = 4.0,
                      seed: int = 2024) -> np.ndarray:
    """rho = mean * exp(sigma_g * g - sigma_g^2 / 2), g a unit Gaussian random field smoothed over
    ``smooth_cells`` cells (periodic, FFT).  SURVEY.md section 8d, config 2."""
    rng = np.random.default_rng(seed)
    white = rng.standard_normal(shape)
    k = [np.fft.fftfreq(n) * 2.0 * np.pi for n in shape]
    kx, ky, kz = np.meshgrid(*k, indexing="ij")
    filt = np.exp(-0.5 * (kx ** 2 + ky ** 2 + kz ** 2) * smooth_cells ** 2)
    g = np.real(np.fft.ifftn(np.fft.fftn(white) * filt))
    g = (g - g.mean()) / g.std()
    return (mean_density * np.exp(sigma_g * g - 0.5 * sigma_g ** 2)).ravel()
