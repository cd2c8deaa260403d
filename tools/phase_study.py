#!/usr/bin/env python
"""Offline study (CPU, numpy) for DESIGN.md section 5.3: how many phases / dependent macro-tile levels does the patch
form with phases need on a Voronoi grid, for different patch orders pi and patch sizes?  Pure graph arithmetic on the
flat grid; no GPU.  usage: python tools/phase_study.py [cells_per_dim] [patch_cells] [lattice jitter; default: Poisson points]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from subsweep_b200 import Directions, capi, grid as G   # noqa: E402


def longest_path_levels(src, dst, n):
    """level[v] = 1 + max level of its predecessors (0 if none), by relaxation (the graph is a DAG)."""
    lvl = np.zeros(n, dtype=np.int64)
    while True:
        new = lvl.copy()
        np.maximum.at(new, dst, lvl[src] + 1)
        if np.array_equal(new, lvl):
            return lvl
        lvl = new


def study(n=14, patch_cells=64, n_dirs=84, seed=1338, jitter=None):
    rng = np.random.default_rng(seed)
    box = 1.0
    if jitter is None:        # Poisson-Voronoi
        pts = rng.uniform(0.0, box, size=(n ** 3, 3))
    else:                     # jittered lattice (the Voronoi variant of the headline box, SURVEY.md 8d config 2: jitter 0.35)
        h = box / n
        i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        pts = (np.stack([i, j, k], axis=-1).reshape(-1, 3) + 0.5 + jitter * rng.uniform(-1.0, 1.0, size=(n ** 3, 3))) * h
    g = G.voronoi(pts, box, periodic=True)
    N = g.n_cells
    lib = capi.load()
    patch = np.empty(N, dtype=np.uint32)
    pos = np.ascontiguousarray(g.positions, dtype=np.float64)
    P = lib.ssw_patch_lattice(capi.dptr(pos), N, patch_cells, patch.ctypes.data_as(C.POINTER(C.c_uint32)))
    patch = patch.astype(np.int64)
    centre = np.zeros((P, 3))
    np.add.at(centre, patch, pos)
    centre /= np.maximum(np.bincount(patch, minlength=P), 1)[:, None]
    dirs = Directions.from_num(n_dirs).xyz
    off = np.asarray(g.face_offsets, dtype=np.int64)
    cell_of_face = np.repeat(np.arange(N), np.diff(off))
    local = np.asarray(g.face_kind) == 0
    nb = np.asarray(g.face_neighbour, dtype=np.int64)
    normal = np.asarray(g.face_normal).reshape(-1, 3)
    rows = []
    for d in range(0, n_dirs, max(1, n_dirs // 12)):          # a sample of the directions, each its own group (kd = 1)
        nd = normal @ dirs[d]
        up = local & (nd < 0.0)
        src, dst = nb[up], cell_of_face[up]                   # src is upwind of dst
        wl = longest_path_levels(src, dst, N)
        res = {"dir": d, "wavefront_levels": int(wl.max()) + 1}
        orders = {
            "lattice": None,
            "projection": np.argsort(np.argsort(centre @ dirs[d], kind="stable")),
            "mean_level": np.argsort(np.argsort(np.bincount(patch, weights=wl, minlength=P) / np.maximum(np.bincount(patch, minlength=P), 1), kind="stable")),
        }
        # lattice order as in patch.cuh: diagonal of the octant, needs the lattice coordinates -> approximate by projection on the octant diagonal
        diag = np.sign(dirs[d]) + (dirs[d] == 0)
        orders["lattice"] = np.argsort(np.argsort(centre @ diag, kind="stable"))
        for name, pi in orders.items():
            back = (pi[patch[src]] > pi[patch[dst]]).astype(np.int64)
            phase = np.zeros(N, dtype=np.int64)
            while True:                                        # phase[dst] = max(phase[src] + back), DAG relaxation
                new = phase.copy()
                np.maximum.at(new, dst, phase[src] + back)
                if np.array_equal(new, phase):
                    break
                phase = new
            n_phase = int(phase.max()) + 1
            mt = patch * n_phase + phase                      # macro-tile = (patch, phase)
            uniq, mt_id = np.unique(mt, return_inverse=True)
            e = np.unique(np.stack([mt_id[src], mt_id[dst]], axis=1), axis=0)
            e = e[e[:, 0] != e[:, 1]]
            # chain the phases of one patch
            pp = uniq // n_phase
            chain = np.flatnonzero(pp[1:] == pp[:-1])
            e = np.concatenate([e, np.stack([chain, chain + 1], axis=1)])
            ml = longest_path_levels(e[:, 0], e[:, 1], len(uniq))
            res[name] = (n_phase, len(uniq), int(ml.max()) + 1)
        rows.append(res)
    return N, P, rows


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    pc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    jit = float(sys.argv[3]) if len(sys.argv) > 3 else None
    N, P, rows = study(n, pc, jitter=jit)
    print(f"{N} Voronoi cells, {P} patches of ~{pc} cells; per direction: wavefront levels | (phases, macro-tiles, macro-tile levels) per patch order")
    for r in rows:
        print(r)
