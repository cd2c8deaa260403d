"""Pins the oracle against every known answer the reference's own tests hold for this path
(SURVEY.md section 8c).  The reference pins no fluxes / xHII / T, so parity for those stays
"unpinned"; these are the scheduling rules, the derivative-consistency checks and the two
production-like chemistry inputs that must terminate."""
import ctypes as C

import numpy as np
import pytest

import oracle


# src/sweep/timestep_level.rs:66-89, compute_timestep_level
@pytest.mark.parametrize("max_num_levels,secs_desired,result", [
    (1, 1.0, 0), (2, 1.0, 0), (1, 0.001, 0), (2, 0.001, 1), (3, 0.001, 2), (2, 0.500001, 1),
    (2, 0.499999, 1), (3, 0.499999, 2), (5, 100.0, 0), (5, 0.0, 4),
])
def test_compute_timestep_level(max_num_levels, secs_desired, result):
    assert oracle.load().orc_level_from_timesteps(max_num_levels, 1.0, secs_desired) == result


def test_level_rule_saturating_cast():
    lib = oracle.load()
    # Rust `as usize`: NaN -> 0, negative -> 0, +inf -> usize::MAX (then clamped)
    assert lib.orc_level_from_timesteps(4, 1.0, float("nan")) == 0
    assert lib.orc_level_from_timesteps(4, 1.0, float("inf")) == 0
    assert lib.orc_level_from_timesteps(4, 1.0, -1.0) == 0
    assert lib.orc_level_from_timesteps(4, 1.0, 1e-300) == 3


def _order(n_levels, lowest):
    out = (C.c_int * 4096)()
    n = oracle.load().orc_levels_in_sweep_order(n_levels, lowest, out, 4096)
    return list(out[:n])


# src/sweep/timestep_state.rs:116-136, iter_levels_in_sweep_order_advances_properly
def test_iter_levels_in_sweep_order_advances_properly():
    assert _order(5, 4) == [4]
    assert _order(5, 3) == [3, 4]
    assert _order(5, 2) == [2, 4, 3, 4]
    assert _order(5, 1) == [1, 4, 3, 4, 2, 4, 3, 4]
    assert _order(5, 0) == [0, 4, 3, 4, 2, 4, 3, 4, 1, 4, 3, 4, 2, 4, 3, 4]


# src/sweep/timestep_state.rs:98-114, lowest_allowed: the warm-up of the lowest allowed level
def test_lowest_allowed_warm_up():
    from helpers import make_problem
    params, g, f = make_problem(n=3, n_dirs=1, n_levels=5, n_sources=0)
    s = oracle.OracleSweep(params, g, **f)
    seen = [s.lowest_allowed_level()]
    for _ in range(6):
        s.run_sweeps()
        seen.append(s.lowest_allowed_level())
    assert seen == [4, 4, 3, 2, 1, 0, 0]


def _solver(T):
    return oracle.Solver(0.0, T, 0.0, 0.0, 0.0, 0.0, 1.0, 0, 0.0, 0.0)


# src/chemistry/hydrogen_only/mod.rs:491-591: analytic derivative vs finite difference, 10 %
@pytest.mark.parametrize("fn,dfn", [
    ("alpha_b", "dalpha_b"), ("coll_ion", "dcoll_ion"), ("coll_ion_cool", "dcoll_ion_cool"),
    ("coll_exc_cool", "dcoll_exc_cool"), ("recomb_cool", "drecomb_cool"), ("brems", "dbrems"),
    ("compton", "dcompton"),
])
def test_numerical_derivative(fn, dfn):
    lib = oracle.load()
    epsilon, delta = 1e-1, 1e-6
    for T in [1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7]:
        s = _solver(T)
        analytical = lib.orc_fit(C.byref(s), oracle.FITS[dfn])
        v1 = lib.orc_fit(C.byref(s), oracle.FITS[fn])
        s.temperature += delta
        v2 = lib.orc_fit(C.byref(s), oracle.FITS[fn])
        numerical = (v2 - v1) / delta
        assert abs(analytical - numerical) / (abs(analytical) + abs(numerical) + 1e-50) < epsilon, (fn, T)


# src/chemistry/hydrogen_only/mod.rs:948-984: fully_ionized_solver / fully_neutral_solver
@pytest.mark.parametrize("x0", [1.0, 0.0])
def test_production_like_inputs_terminate(x0):
    lib = oracle.load()
    s = oracle.Solver(x0, 1791871.5383082589, 0.000000000000000000000000015411844211187435 * 1e-3 / 1e-6,
                      8873284571355481000000000000000000000000000000000000000000000.0,
                      6.709257125565072 * oracle.const("kiloparsec"),
                      466103097665666700000000000000000000000000000.0, 8.35028211377591, 0, 0.0, 0.0)
    r = oracle.ChemResult()
    lib.orc_perform_timestep(C.byref(s), 1.0 * oracle.const("megayear"), 0.1, C.byref(r))
    assert not r.failed
    assert 1e-10 <= s.xhii <= 1.0 - 1e-10 and s.temperature > 0 and np.isfinite(r.timescale)
    assert r.attempts >= 1 and r.max_depth <= 100


def test_unit_constants():
    # src/units/mod.rs:21,48,101-108
    assert oracle.const("proton_mass") == 1.67262192369e-27
    assert oracle.const("boltzmann") == 1.380649e-23
    assert oracle.const("gamma") == 5.0 / 3.0
    assert oracle.const("year") == 3.15576e7
    assert oracle.const("parsec") == 3.0857e16
    assert oracle.const("sigma") == pytest.approx(2.9580524545305314e-22, rel=1e-15)
    assert oracle.const("photon_energy") == pytest.approx(18.028356312818811 * 1.602176634e-19, rel=1e-15)


def _py_heap_pop_order(keys):
    """Independent transcription of Rust's BinaryHeap on (key, id) pairs."""
    data = [(k, i) for i, k in enumerate(keys)]

    def sift_down_range(pos, end):
        elem = data[pos]
        child = 2 * pos + 1
        while child <= max(end - 2, 0) and end >= 2:
            if data[child][0] <= data[child + 1][0]:
                child += 1
            if elem[0] >= data[child][0]:
                data[pos] = elem
                return
            data[pos] = data[child]
            pos = child
            child = 2 * pos + 1
        if child == end - 1 and elem[0] < data[child][0]:
            data[pos] = data[child]
            pos = child
        data[pos] = elem

    def sift_up(start, pos):
        elem = data[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if elem[0] <= data[parent][0]:
                break
            data[pos] = data[parent]
            pos = parent
        data[pos] = elem

    n = len(data) // 2
    while n > 0:
        n -= 1
        sift_down_range(n, len(data))
    order = []
    while data:
        item = data.pop()
        if data:
            item, data[0] = data[0], item
            end = len(data)
            pos = 0
            elem = data[0]
            child = 1
            while end >= 2 and child <= end - 2:
                if data[child][0] <= data[child + 1][0]:
                    child += 1
                data[pos] = data[child]
                pos = child
                child = 2 * pos + 1
            if child == end - 1:
                data[pos] = data[child]
                pos = child
            data[pos] = elem
            sift_up(0, pos)
        order.append(item[1])
    return order


@pytest.mark.parametrize("n,nkeys", [(1, 1), (2, 2), (7, 3), (64, 4), (1000, 21), (4097, 84)])
def test_binary_heap_pop_order(n, nkeys):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, nkeys, size=n).astype(np.uint32)
    order = np.empty(n, dtype=np.uint32)
    oracle.load().orc_heap_pop_order(keys.ctypes.data_as(C.POINTER(C.c_uint32)), n,
                                     order.ctypes.data_as(C.POINTER(C.c_uint32)))
    popped = keys[order]
    assert sorted(order.tolist()) == list(range(n))          # a permutation
    assert np.all(np.diff(popped.astype(np.int64)) <= 0)     # max-heap on the direction index
    assert order.tolist() == _py_heap_pop_order(keys.tolist())


# ---------------------------------------------------------------------------------------------
# Double entry: the rate fits typed a second time, in Python, straight from the reference's expressions
# (src/chemistry/hydrogen_only/mod.rs:161-263), against the C restatement -- a transcription slip in either shows.
# ---------------------------------------------------------------------------------------------
def _fits_second_entry(t):
    import math
    cm3_per_s, ergs_cm3_per_s = 1e-6, 1e-7 * 1e-6
    lam = 315614.0 / t
    fit = math.sqrt(t) / (1.0 + math.sqrt(t / 1e5)) * math.exp(-157809.1 / t)                       # :161-164
    c1, c2 = 1.0 / 1e5, 157809.1
    dfit = (math.exp(-c2 / t) * (c1 * c2 * t + 0.5 * math.sqrt(c1 * t) * (2.0 * c2 + t))) / \
           (math.sqrt(t ** 3) * math.sqrt(c1 * t) * (math.sqrt(c1 * t) + 1.0) ** 2)                 # :166-173
    p = (lam / 2.74) ** 0.407
    q = (315614.0 / (2.25 * t)) ** 0.376
    return {
        "alpha_b": 2.753e-14 * lam ** 1.5 / (1.0 + p) ** 2.242 * cm3_per_s,                          # :175-180
        "dalpha_b": 2.753e-14 * (-math.sqrt(lam) * (p + 1.0) ** (-2.242 - 1.0) * (0.407 * 2.242 * p - 1.5 * p - 1.5))
                    * cm3_per_s * (-315614.0 / t ** 2),                                            # :182-194
        "recomb_cool": 3.435e-30 * t * lam ** 1.97 / (1.0 + (lam / 2.25) ** 0.376) ** 3.72 * ergs_cm3_per_s,   # :196-202
        "drecomb_cool": 3.435e-30 * ((1.0 + q) ** (-1.0 - 3.72) * (1.0 - 1.97 + (1.0 - 1.97 + 0.376 * 3.72) * q)
                                     * (315614.0 / t) ** 1.97) * ergs_cm3_per_s,                   # :204-216
        "coll_ion": 5.85e-11 * fit * cm3_per_s, "dcoll_ion": 5.85e-11 * dfit * cm3_per_s,             # :218-225
        "coll_ion_cool": 1.27e-21 * fit * ergs_cm3_per_s, "dcoll_ion_cool": 1.27e-21 * dfit * ergs_cm3_per_s,   # :227-235
        "coll_exc_cool": 7.5e-19 / (1.0 + math.sqrt(t / 1e5)) * math.exp(-118348.0 / t) * ergs_cm3_per_s,     # :237-242
        "brems": 1.42e-27 * math.sqrt(t) * ergs_cm3_per_s, "dbrems": 1.42e-27 / (2.0 * math.sqrt(t)) * ergs_cm3_per_s,   # :255-263
    }


@pytest.mark.parametrize("T", [1e1, 3e2, 1e4, 2.5e4, 1e5, 3e6, 1e7])
def test_rate_fits_double_entry(T):
    lib = oracle.load()
    s = _solver(T)
    for name, want in _fits_second_entry(T).items():
        got = lib.orc_fit(C.byref(s), oracle.FITS[name])
        assert got == pytest.approx(want, rel=1e-12, abs=1e-300), (name, T)


def test_case_b_recombination_matches_the_literature_value():
    """alpha_B(1e4 K) = 2.59e-13 cm^3/s (Hui & Gnedin 1997, the fit the reference uses via Rosdahl et al. 2013)."""
    lib = oracle.load()
    s = _solver(1e4)
    assert lib.orc_fit(C.byref(s), oracle.FITS["alpha_b"]) == pytest.approx(2.59e-13 * 1e-6, rel=5e-3)
