"""Shared problem builders for the parity tests (seeded, small enough for the CPU oracle)."""
from __future__ import annotations

import numpy as np

from subsweep_b200 import SweepParameters, grid as gridmod
from subsweep_b200 import units as U


def make_grid(kind: str, n: int, periodic: bool, box: float, seed: int = 1338):
    if kind == "cartesian":
        return gridmod.cartesian((n, n, n), box, periodic)
    if kind == "voronoi":
        rng = np.random.default_rng(seed)
        pts = rng.uniform(0.0, box, size=(n ** 3, 3))
        return gridmod.voronoi(pts, box, periodic)
    if kind == "jittered":
        rng = np.random.default_rng(seed)
        h = box / n
        ijk = np.stack(np.meshgrid(*(np.arange(n),) * 3, indexing="ij"), axis=-1).reshape(-1, 3)
        pts = (ijk + 0.5 + 0.35 * rng.uniform(-1, 1, size=ijk.shape)) * h
        return gridmod.voronoi(pts, box, periodic)
    raise ValueError(kind)


def make_problem(kind="cartesian", n=10, periodic=False, n_dirs=21, n_levels=1, n_sources=2,
                 source_rate=1e51, nh_cm3=1e-4, cell_kpc=5.0, max_timestep_myr=0.1, threshold=1e-5,
                 prevent_cooling=True, seed=7, lognormal=True, x0=1e-10, t0=100.0):
    """A small reionisation-like box: log-normal density, a few point sources in the densest cells."""
    box = n * cell_kpc * U.KILOPARSEC
    g = make_grid(kind, n, periodic, box, seed=1338 + seed)
    N = g.n_cells
    rng = np.random.default_rng(seed)
    mean_rho = nh_cm3 * U.PER_CUBIC_CENTIMETER * U.PROTON_MASS
    if lognormal:
        rho = mean_rho * np.exp(0.8 * rng.standard_normal(N) - 0.32)
    else:
        rho = np.full(N, mean_rho)
    x = np.full(N, x0)
    T = np.full(N, t0)
    src = np.zeros(N)
    if n_sources > 0:
        idx = np.argsort(rho)[-n_sources:]
        src[idx] = source_rate * (1.0 + 0.5 * np.arange(n_sources))
    params = SweepParameters(
        directions=n_dirs, num_timestep_levels=n_levels, periodic=periodic,
        max_timestep=max_timestep_myr * U.MEGAYEARS, significant_rate_threshold=threshold,
        timestep_safety_factor=0.1, chemistry_timestep_safety_factor=0.1, prevent_cooling=prevent_cooling)
    return params, g, dict(density=rho, ionized_hydrogen_fraction=x, temperature=T, source=src)


def rel_err(a: np.ndarray, b: np.ndarray, floor: float = 0.0) -> float:
    """max |a-b| / max(|a|, |b|, floor) over the arrays (0/0 counts as 0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(den > 0, np.abs(a - b) / den, 0.0)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    r = np.where(both_inf, 0.0, r)
    both_nan = np.isnan(a) & np.isnan(b)
    r = np.where(both_nan, 0.0, r)
    r = np.where(np.isnan(r), np.inf, r)
    return float(r.max()) if r.size else 0.0


def assert_close(a, b, rtol, floor=0.0, what=""):
    err = rel_err(a, b, floor)
    assert err <= rtol, f"{what}: max relative error {err:.3e} > {rtol:.1e}"
