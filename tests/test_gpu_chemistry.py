"""CUDA chemistry kernel vs the oracle: HydrogenOnly::update_abundances
(src/chemistry/hydrogen_only/mod.rs:90-119) on independent cells, through ssw_chemistry_batch."""
import ctypes as C

import numpy as np
import pytest

import oracle
from helpers import assert_close
from subsweep_b200 import capi
from subsweep_b200 import units as U

pytestmark = pytest.mark.gpu

RTOL = 1e-9   # north_star: ionized fraction, temperature and rates within relative 1e-9 after one step


def gpu_chemistry(lib, x, T, rho, vol, length, rate, dt, scale_factor, safety, prevent_cooling):
    n = len(x)
    x, T = x.copy(), T.copy()
    ts = np.empty(n)
    proc = np.empty(n, dtype=np.int32)
    depth = np.empty(n, dtype=np.int32)
    att = np.empty(n, dtype=np.uint64)
    rc = lib.ssw_chemistry_batch(0, n, capi.dptr(x), capi.dptr(T), capi.dptr(rho), capi.dptr(vol), capi.dptr(length),
                                 capi.dptr(rate), capi.dptr(dt), scale_factor, safety, int(prevent_cooling),
                                 capi.dptr(ts), proc.ctypes.data_as(C.POINTER(C.c_int32)),
                                 depth.ctypes.data_as(C.POINTER(C.c_int32)), att.ctypes.data_as(C.POINTER(C.c_uint64)))
    capi.check(lib, rc)
    return dict(xhii=x, temperature=T, timescale=ts, process=proc, depth=depth, attempts=att)


def random_cells(n, seed):
    rng = np.random.default_rng(seed)
    x = np.where(rng.random(n) < 0.3, 1e-10, rng.uniform(0, 1, n))
    x[rng.random(n) < 0.1] = 1.0 - 1e-10
    T = 10.0 ** rng.uniform(1, 7, n)
    rho = 10.0 ** rng.uniform(-6, 1, n) * U.PER_CUBIC_CENTIMETER * U.PROTON_MASS
    length = 10.0 ** rng.uniform(-1, 2, n) * U.KILOPARSEC
    vol = length ** 3 * rng.uniform(0.5, 4.0, n)
    rate = np.where(rng.random(n) < 0.3, 0.0, 10.0 ** rng.uniform(30, 56, n))
    dt = 10.0 ** rng.uniform(-3, 1, n) * U.MEGAYEARS
    return x, T, rho, vol, length, rate, dt


def conditioning(cells, ref):
    """Cells where the reference's own formulas lose more than 1e-10 of relative precision:
    absorbed fraction 1 - exp(-tau) with tau << 1 (hydrogen_only/mod.rs:315-316) amplifies a
    1-ulp difference of exp() to eps/tau, and the neutral fraction 1 - x with x -> 1 amplifies a
    1-ulp difference of x to eps/(1-x).  glibc's and CUDA's exp/pow legitimately differ in the
    last bit (the Rust reference calls the platform libm), so 1e-9 is only meaningful where these
    amplifications stay below it."""
    x0, _, rho, _, length, _, _ = cells
    sigma = oracle.const("sigma")
    nh = rho / U.PROTON_MASS
    neutral = np.minimum(1.0 - x0, 1.0 - ref["xhii"])
    tau = nh * sigma * length * neutral
    return (tau >= 1e-4) & (neutral >= 1e-4)


@pytest.mark.parametrize("prevent_cooling", [False, True])
@pytest.mark.parametrize("scale_factor", [1.0, 0.125])
def test_chemistry_matches_oracle(cuda_lib, prevent_cooling, scale_factor):
    cells = random_cells(3000, seed=11 + int(prevent_cooling))
    ref = oracle.chemistry(*cells, scale_factor=scale_factor, safety=0.1, prevent_cooling=prevent_cooling)
    got = gpu_chemistry(cuda_lib, *cells, scale_factor, 0.1, prevent_cooling)
    same_path = (got["attempts"] == ref["attempts"]) & (got["depth"] == ref["depth"])
    # a different substep path means a threshold comparison flipped on a last-bit difference of
    # exp/pow between CUDA and glibc; that must stay a rare event (hard part 3)
    assert same_path.mean() > 0.995, f"{(~same_path).sum()} of {len(same_path)} cells took another substep path"
    well = conditioning(cells, ref)
    assert well.mean() > 0.5
    strict = same_path & well
    loose = same_path & ~well
    for k in ("xhii", "temperature", "timescale"):
        err = np.abs(got[k] - ref[k]) / np.maximum(np.abs(ref[k]), 1e-300)
        err = np.where((got[k] == ref[k]) | (np.isinf(got[k]) & np.isinf(ref[k])), 0.0, err)
        worst = np.argsort(err * strict)[-3:]
        detail = [(int(i), float(err[i]), int(ref["attempts"][i])) for i in worst]
        assert err[strict].max() <= RTOL, f"{k}: well-conditioned cells off by {err[strict].max():.3e}: {detail}"
        # the recommended timescale divides by |dx/x| whose numerator c - x (c + d) cancels: it inherits
        # the amplified last-bit differences once more, so it gets a looser bound than the state
        loose_tol = 1e-3 if k == "timescale" else 1e-5
        assert err[loose].max(initial=0.0) <= loose_tol, f"{k}: ill-conditioned cells off by {err[loose].max():.3e}"
    assert np.array_equal(got["process"][same_path], ref["process"][same_path])
    # cells on another path still agree to the accuracy of the integrator
    for k in ("xhii", "temperature"):
        assert_close(got[k][~same_path], ref[k][~same_path], 5e-2, what=k + " (other path)")


# the reference's two production-like inputs (hydrogen_only/mod.rs:948-984)
@pytest.mark.parametrize("x0,attempts,depth", [(1.0, 1, 0), (0.0, 741, 16)])
def test_production_like_inputs(cuda_lib, x0, attempts, depth):
    one = lambda v: np.array([v], dtype=np.float64)
    args = (one(x0), one(1791871.5383082589), one(1.5411844211187435e-26 * 1e-3 / 1e-6), one(8.873284571355481e60),
            one(6.709257125565072 * U.KILOPARSEC), one(4.661030976656667e44), one(1.0 * U.MEGAYEARS))
    ref = oracle.chemistry(*args, scale_factor=8.35028211377591, safety=0.1, prevent_cooling=False)
    got = gpu_chemistry(cuda_lib, *args, 8.35028211377591, 0.1, False)
    assert ref["attempts"][0] == attempts and ref["depth"][0] == depth
    assert got["attempts"][0] == attempts and got["depth"][0] == depth
    for k in ("xhii", "temperature", "timescale"):
        assert_close(got[k], ref[k], RTOL, what=k)


def test_convergence_failure_is_not_an_error(cuda_lib):
    """depth > 100 -> pessimistic timescale dt/10, state keeps the progress made (mod.rs:426-441)."""
    one = lambda v: np.array([v], dtype=np.float64)
    # an absurd rate forces |dx/x| > safety at every depth down to 2^-100 dt
    args = (one(1e-10), one(100.0), one(1e-3 * U.PER_CUBIC_CENTIMETER * U.PROTON_MASS), one((U.KILOPARSEC) ** 3),
            one(U.KILOPARSEC), one(1e120), one(1.0 * U.MEGAYEARS))
    ref = oracle.chemistry(*args, safety=0.1)
    got = gpu_chemistry(cuda_lib, *args, 1.0, 0.1, False)
    assert ref["process"][0] == -1 and ref["depth"][0] == 100 and ref["attempts"][0] == 101   # the case does fail to converge
    assert got["process"][0] == -1 and got["depth"][0] == 100 and got["attempts"][0] == 101
    assert got["timescale"][0] == ref["timescale"][0] == U.MEGAYEARS / 10.0
    for k in ("xhii", "temperature"):
        assert_close(got[k], ref[k], RTOL, what=k)
