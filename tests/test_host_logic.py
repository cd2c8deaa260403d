"""Host-side logic that needs no GPU: parameter section, directions, sharding, scheduling rules
exported by the C ABI (pure host functions of libsubsweep_b200.so)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from subsweep_b200 import Directions, SweepParameters, capi, direction_shard
from subsweep_b200 import units as U

ROOT = Path(__file__).resolve().parent.parent


def test_sweep_parameters_defaults_and_units():
    p = SweepParameters.from_yaml("""
sweep:
  directions: 84
  num_timestep_levels: 4
  periodic: true
  max_timestep: 1 Myr
  significant_rate_threshold: 1.0e-5 s^-1
""")
    assert p.directions == 84 and p.num_timestep_levels == 4 and p.periodic
    assert p.max_timestep == 1e6 * 3.15576e7
    assert p.significant_rate_threshold == 1e-5
    # defaults of src/sweep/parameters.rs:64-78
    assert p.timestep_safety_factor == 0.1 and p.chemistry_timestep_safety_factor == 0.1
    assert p.prevent_cooling is True and p.rotate_directions is False and p.check_deadlock is False
    assert p.num_tasks_to_solve_before_send_receive == 10000


def test_sweep_parameters_deny_unknown_fields():
    with pytest.raises(ValueError, match="unknown field"):
        SweepParameters.from_dict(dict(directions=1, num_timestep_levels=1, periodic=False,
                                       max_timestep="1 s", bogus=3))
    with pytest.raises(ValueError, match="missing field"):
        SweepParameters.from_dict(dict(directions=1, num_timestep_levels=1, periodic=False))


def test_direction_tables():
    for n in (1, 16, 21, 32, 64, 84):
        d = Directions.from_num(n)
        assert d.xyz.shape == (n, 3)
        # the tables are 6-digit literals and NOT re-normalised (direction/mod.rs:58-75)
        assert np.all(np.abs(np.linalg.norm(d.xyz, axis=1) - 1.0) < 2e-6)
    assert Directions.from_num(1).xyz.tolist() == [[1.0, 0.0, 0.0]]
    assert Directions.from_num(16).xyz[0].tolist() == [-0.887773, 0.0580969, -0.456601]
    with pytest.raises(NotImplementedError):
        Directions.from_num(17)
    e = Directions.explicit([[2.0, 0.0, 0.0], [1.0, 1.0, 0.0]])
    assert np.allclose(np.linalg.norm(e.xyz, axis=1), 1.0, atol=1e-15)


@pytest.mark.parametrize("D", [1, 16, 21, 84])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_direction_shard_partitions(D, world):
    lib = capi.load()
    covered = []
    for r in range(world):
        b, e = direction_shard(D, world, r)
        cb, ce = C.c_int32(), C.c_int32()
        assert lib.ssw_direction_shard(D, world, r, C.byref(cb), C.byref(ce)) == 0
        assert (cb.value, ce.value) == (b, e)
        covered += list(range(b, e))
    assert covered == list(range(D))
    sizes = [direction_shard(D, world, r)[1] - direction_shard(D, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


# src/sweep/timestep_level.rs:66-89 against the library's host-side rule
@pytest.mark.parametrize("max_num_levels,secs_desired,result", [
    (1, 1.0, 0), (2, 1.0, 0), (1, 0.001, 0), (2, 0.001, 1), (3, 0.001, 2), (2, 0.500001, 1),
    (2, 0.499999, 1), (3, 0.499999, 2), (5, 100.0, 0), (5, 0.0, 4),
])
def test_library_level_rule(max_num_levels, secs_desired, result):
    assert capi.load().ssw_level_from_timesteps(max_num_levels, 1.0, secs_desired) == result


def test_library_level_rule_matches_oracle_on_random_input():
    import oracle
    lib, olib = capi.load(), oracle.load()
    rng = np.random.default_rng(3)
    vals = np.concatenate([10.0 ** rng.uniform(-30, 30, 20000), [0.0, np.inf, np.nan, -1.0, 1e-320],
                           2.0 ** np.arange(-40, 40, dtype=np.float64)])
    for L in (1, 2, 4, 7):
        for v in vals:
            assert lib.ssw_level_from_timesteps(L, 3.15576e13, float(v)) == olib.orc_level_from_timesteps(L, 3.15576e13, float(v))


# src/sweep/timestep_state.rs:116-136
def test_library_sweep_order():
    lib = capi.load()

    def order(L, lowest):
        out = (C.c_int32 * 64)()
        n = lib.ssw_levels_in_sweep_order(L, lowest, out, 64)
        return list(out[:n])
    assert order(5, 4) == [4]
    assert order(5, 3) == [3, 4]
    assert order(5, 2) == [2, 4, 3, 4]
    assert order(5, 1) == [1, 4, 3, 4, 2, 4, 3, 4]
    assert order(5, 0) == [0, 4, 3, 4, 2, 4, 3, 4, 1, 4, 3, 4, 2, 4, 3, 4]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "subsweep_b200.h").read_text()
    declared = set(re.findall(r"\b(ssw_[a-z_0-9]+)\s*\(", header))
    declared -= {"ssw_allreduce_fn"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    lib = C.CDLL(str(capi.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert capi.load().ssw_abi_version() == 1


def test_product_never_touches_the_oracle():
    """No import, link, dlopen or call of anything under oracle/ from the product."""
    pat = re.compile(r"(^|\s)(import|from)\s+oracle\b|liboracle|oracle\.h|\borc_[a-z]|oracle/")
    files = list((ROOT / "subsweep_b200").rglob("*.py")) + list((ROOT / "subsweep_b200" / "csrc").glob("*.cu*"))
    files += list((ROOT / "include").glob("*.h"))
    assert files
    for path in files:
        m = pat.search(path.read_text())
        assert m is None, (path, m.group(0))


def test_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from helpers import make_problem
    from subsweep_b200 import Sweep
    params, g, f = make_problem(n=3, n_dirs=1)
    with pytest.raises(capi.SubsweepError) as e:
        Sweep(params, g, **f)
    assert e.value.code == capi.SSW_E_CUDA and "no CPU fallback" in str(e.value)


def test_units():
    assert U.parse_quantity("1 Myr") == 1e6 * 3.15576e7
    assert U.parse_quantity("2.5 kpc") == 2.5 * 1000 * 3.0857e16
    assert U.parse_quantity(0.1) == 0.1
    with pytest.raises(ValueError):
        U.parse_quantity("1 parsecs")


# ---- host-side pieces of the patch-ordered sweep (patch.cuh / sweep.cu) ----------------
def _patch_lattice(pos, target):
    lib = capi.load()
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    out = np.empty(len(pos), dtype=np.uint32)
    n = lib.ssw_patch_lattice(capi.dptr(pos), len(pos), target, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return n, out


@pytest.mark.parametrize("n,target,per_axis,cells", [(16, 64, 4, 64), (16, 512, 2, 512), (128, 512, 16, 512), (12, 8, 6, 8)])
def test_patch_lattice_aligns_with_a_cartesian_grid(n, target, per_axis, cells):
    """Cell centres of a Cartesian grid, boxes of `target` cells: the lattice must coincide with whole blocks of cells."""
    h = 3.7
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    pos = np.stack([(i + 0.5) * h, (j + 0.5) * h, (k + 0.5) * h], axis=-1).reshape(-1, 3)
    n_patches, patch = _patch_lattice(pos, target)
    assert n_patches == per_axis ** 3
    w = n // per_axis
    expect = ((i // w) * per_axis + (j // w)) * per_axis + (k // w)
    assert np.array_equal(patch.reshape(n, n, n), expect)
    assert np.all(np.bincount(patch, minlength=n_patches) == cells)


def test_patch_lattice_on_scattered_points_and_degenerate_axes():
    rng = np.random.default_rng(7)
    pos = rng.uniform(0.0, 1.0, size=(20000, 3))
    n_patches, patch = _patch_lattice(pos, 125)
    counts = np.bincount(patch, minlength=n_patches)
    assert n_patches > 1 and counts.sum() == len(pos) and counts.max() <= 1024
    assert 60 < counts.mean() < 250                       # about the requested size
    # neighbouring points land in the same or an adjacent box: boxes are spatially compact
    flat = pos.copy()
    flat[:, 2] = 0.25                                     # a plane: one box along the degenerate axis
    n2, patch2 = _patch_lattice(flat, 100)
    assert n2 > 1 and np.bincount(patch2, minlength=n2).max() <= 1024
    n1, patch1 = _patch_lattice(np.zeros((1, 3)), 512)    # a single cell
    assert n1 == 1 and patch1[0] == 0


@pytest.mark.parametrize("kd", [1, 3, 6, 11, 32])
def test_direction_groups_follow_the_octants(kd):
    lib = capi.load()
    d = Directions.from_num(84).xyz
    grp = np.empty(84, dtype=np.int32)
    n_groups = lib.ssw_direction_groups(capi.dptr(d), 84, kd, grp.ctypes.data_as(C.POINTER(C.c_int32)))
    assert n_groups >= 1 and grp.min() == 0 and grp.max() == n_groups - 1
    sign = np.sign(d).astype(np.int64)
    for g in range(n_groups):
        members = np.flatnonzero(grp == g)
        assert 1 <= len(members) <= kd
        assert len({tuple(s) for s in sign[members]}) == 1          # one sign pattern per group: one dependency DAG
    # as few groups as the octants allow
    classes = {}
    for s in map(tuple, sign):
        classes[s] = classes.get(s, 0) + 1
    assert n_groups == sum(-(-c // kd) for c in classes.values())


def test_patch_levels_kahn_over_the_quotient_graph():
    lib = capi.load()
    n, empty = 4, 0xFFFFFFFF
    P = n ** 3
    dep = np.full((2, P, 64), empty, dtype=np.uint32)   # 64 slots per macro-tile (kMaxPatchDeps)
    idx = lambda i, j, k: (i * n + j) * n + k   # noqa: E731
    for i in range(n):
        for j in range(n):
            for k in range(n):
                up0 = [idx(a, b, c) for a, b, c in ((i - 1, j, k), (i, j - 1, k), (i, j, k - 1)) if min(a, b, c) >= 0]
                up1 = [idx(a, b, c) for a, b, c in ((i + 1, j, k), (i, j, k - 1)) if 0 <= a < n and c >= 0]
                dep[0, idx(i, j, k), :len(up0)] = up0       # octant (+,+,+)
                dep[1, idx(i, j, k), :len(up1)] = up1       # a direction with d_y = 0, d_x < 0
    lvl = np.empty((2, P), dtype=np.uint32)
    n_levels = lib.ssw_patch_levels(dep.ctypes.data_as(C.POINTER(C.c_uint32)), 2, P, lvl.ctypes.data_as(C.POINTER(C.c_uint32)))
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    assert n_levels == 3 * (n - 1) + 1
    assert np.array_equal(lvl[0].reshape(n, n, n), i + j + k)
    assert np.array_equal(lvl[1].reshape(n, n, n), (n - 1 - i) + k)
    # two patches that are upwind of each other (a jagged Voronoi patch boundary): no patch form
    dep[1, idx(0, 0, 0), 0] = idx(0, 0, 1)
    dep[1, idx(0, 0, 1), :2] = [idx(1, 0, 1), idx(0, 0, 0)]
    assert lib.ssw_patch_levels(dep.ctypes.data_as(C.POINTER(C.c_uint32)), 2, P, lvl.ctypes.data_as(C.POINTER(C.c_uint32))) == capi.SSW_E_DEADLOCK


def test_failing_collective_hooks_return_an_error_code(capsys):
    """A raising Python hook must come back to C as a non-zero return (the library maps it to SSW_E_COMM), never as
    an exception swallowed by ctypes with a 0 = success return."""
    from subsweep_b200.sweep import allreduce_trampoline, collective_trampoline

    def boom(*_a):
        raise RuntimeError("link down")

    cb = allreduce_trampoline(boom)
    assert cb(None, None, 4, None) == -1
    cc = collective_trampoline(boom)
    assert cc(None, 1, None, 4, None) == -1
    err = capsys.readouterr().err
    assert "allreduce hook failed" in err and "collective hook failed" in err
    ok = allreduce_trampoline(lambda *a: None)
    assert ok(None, 8, 4, None) == 0


def test_random_rotation_is_a_rotation():
    """get_random_rotation_matrix (src/sweep/direction/mod.rs:113-148) in the Python mirror."""
    from subsweep_b200.sweep import random_rotation_matrix, rotation_matrix
    m = random_rotation_matrix(np.random.default_rng(1337))
    assert np.allclose(m @ m.T, np.eye(3), atol=1e-14) and abs(np.linalg.det(m) - 1.0) < 1e-14
    z90 = rotation_matrix((0.0, 0.0, 1.0), np.pi / 2)
    assert np.allclose(z90 @ np.array([1.0, 0.0, 0.0]), [0.0, 1.0, 0.0], atol=1e-15)


def test_remap_from_takes_the_nearest_old_particle_and_the_larger_value():
    """remap_abundances_and_energies_system (src/arepo_postprocess/remap.rs:380-429)."""
    from subsweep_b200.snapshot import remap_from
    old_pos = np.array([[0.1, 0.1, 0.1], [0.9, 0.9, 0.9], [0.5, 0.5, 0.5]])
    new_pos = np.array([[0.12, 0.1, 0.1], [0.52, 0.5, 0.49], [0.02, 0.98, 0.95]])
    T, x = remap_from(new_pos, [100.0, 5e4, 100.0], [1e-10, 0.9, 1e-10], old_pos, [2e4, 100.0, 1e4], [0.5, 1e-10, 0.3], box_size=1.0)
    assert T.tolist() == [2e4, 5e4, 100.0] and x.tolist() == [0.5, 0.9, 1e-10]       # third: nearest through the wrap is old #1
    T2, _ = remap_from(new_pos, [100.0] * 3, [1e-10] * 3, old_pos, [2e4, 100.0, 1e4], [0.5, 1e-10, 0.3])
    assert T2.tolist() == [2e4, 1e4, 1e4]                                             # no wrap: old #2 is nearer


def test_rust_sys_crate_declares_every_symbol():
    """crates/subsweep_b200_sys/src/lib.rs is a transcription of the header: same functions, same enum values."""
    header = (ROOT / "include" / "subsweep_b200.h").read_text()
    crate = (ROOT / "crates" / "subsweep_b200_sys" / "src" / "lib.rs").read_text()
    declared = set(re.findall(r"\b(ssw_[a-z_0-9]+)\s*\(", header)) - {"ssw_allreduce_fn"}
    bound = set(re.findall(r"pub fn (ssw_[a-z_0-9]+)\s*\(", crate))
    assert declared == bound, declared ^ bound
    for name, value in re.findall(r"SSW_F_([A-Z_]+) = (\d+)", header):
        assert re.search(rf"\b{name} = {value},", crate), name
    for name, value in re.findall(r"SSW_STAT_([A-Z_]+) = (\d+)", header):
        assert re.search(rf"\b{name} = {value},", crate), name
