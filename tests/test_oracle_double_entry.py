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


# ---------------------------------------------------------------------------------------------
# Chemistry: Solver::perform_timestep typed a second time, RECURSIVELY as the reference writes it
# (src/chemistry/hydrogen_only/mod.rs:139-461), in numpy float64 scalars (IEEE division by zero, NaN, inf like Rust).
# ---------------------------------------------------------------------------------------------
F = np.float64
INV_EPS = F(1.0) / F(np.finfo(np.float64).eps)
KB, MP, GAMMA = F(1.380649e-23), F(1.67262192369e-27), F(5.0) / F(3.0)
EV = F(1.602176634e-19)
E_PHOTON, RYDBERG = F(18.028356312818811) * EV, F(13.65693) * EV
CM3_PER_S, ERGS_CM3_PER_S, ERGS_PER_S = F(1e-6), F(1e-7) * F(1e-6), F(1e-7)
SIG = F(2.9580524545305314e-18) * (F(0.01) * F(0.01))


def _exp(v): return F(math.exp(v))      # glibc, like the C restatement (numpy's own SIMD exp may differ in the last bit)
def _sqrt(v): return F(math.sqrt(v))
def _pow(b, e): return F(math.pow(b, e))


class PySolver:
    def __init__(self, x, t, rho, vol, length, rate, a, prevent_cooling):
        self.x, self.t, self.rho, self.vol, self.len, self.rate, self.a = map(F, (x, t, rho, vol, length, rate, a))
        self.floor = (self.t, self.x) if prevent_cooling else None
        self.attempts, self.max_depth = 0, 0

    # :139-159
    def nh(self): return self.rho / MP
    def nh1(self): return self.nh() * self.x
    def nh0(self): return self.nh() * (F(1.0) - self.x)
    def ne(self): return self.nh1()
    def mu(self): return F(1.0) / (self.x + F(1.0))

    def fit(self):                                                                     # :161-164
        t = self.t
        return _sqrt(t) / (F(1.0) + _sqrt(t / F(1e5))) * _exp(F(-157809.1) / t)

    def dfit(self):                                                                    # :166-173
        c1, c2, t = F(1.0) / F(1e5), F(157809.1), self.t
        return (_exp(-c2 / t) * (c1 * c2 * t + F(0.5) * _sqrt(c1 * t) * (F(2.0) * c2 + t))) / \
               (_sqrt(t * t * t) * _sqrt(c1 * t) * ((_sqrt(c1 * t) + F(1.0)) * (_sqrt(c1 * t) + F(1.0))))

    def alpha(self):                                                                   # :175-180
        lam = F(315614.0) / self.t
        return F(2.753e-14) * _pow(lam, F(1.5)) / _pow(F(1.0) + _pow(lam / F(2.74), F(0.407)), F(2.242)) * CM3_PER_S

    def dalpha(self):                                                                  # :182-194
        lam = F(315614.0) / self.t
        dlam = F(-315614.0) / (self.t * self.t)
        c1, c2, c3 = F(1.0) / F(2.74), F(0.407), F(2.242)
        p = _pow(c1 * lam, c2)
        d = -_sqrt(lam) * _pow(p + F(1.0), -c3 - F(1.0)) * (c2 * c3 * p - F(1.5) * p - F(1.5))
        return (F(2.753e-14) * d) * CM3_PER_S * dlam

    def rec_cool(self):                                                                # :196-202
        lam = F(315614.0) / self.t
        return (F(3.435e-30) * self.t * _pow(lam, F(1.97)) / _pow(F(1.0) + _pow(lam / F(2.25), F(0.376)), F(3.72))) * ERGS_CM3_PER_S

    def drec_cool(self):                                                               # :204-216
        c1, c2, c3, c4, c5, t = F(315614.0), F(1.97), F(0.376), F(3.72), F(2.25), self.t
        p = _pow(c1 / (c5 * t), c3)
        der = _pow(F(1.0) + p, F(-1.0) - c4) * (F(1.0) - F(1.0) * c2 + (F(1.0) - F(1.0) * c2 + c3 * c4) * p) * _pow(c1 / t, c2)
        return (F(3.435e-30) * der) * ERGS_CM3_PER_S

    def beta(self): return (F(5.85e-11) * self.fit()) * CM3_PER_S                      # :218-225
    def dbeta(self): return (F(5.85e-11) * self.dfit()) * CM3_PER_S
    def ion_cool(self): return (F(1.27e-21) * self.fit()) * ERGS_CM3_PER_S             # :227-235
    def dion_cool(self): return (F(1.27e-21) * self.dfit()) * ERGS_CM3_PER_S

    def exc_cool(self):                                                                # :237-242
        t = self.t
        return (F(7.5e-19) / (F(1.0) + _sqrt(t / F(1e5))) * _exp(F(-118348.0) / t)) * ERGS_CM3_PER_S

    def dexc_cool(self):                                                               # :244-253
        t, c1, c2, c3 = self.t, F(7.5e-19), F(118348.0), F(1.0) / F(1e5)
        s = _sqrt(c3 * t)
        return ((c1 * _exp(-c2 / t) * (c2 * c3 * t - F(0.5) * c3 * (t * t) + c2 * s)) /
                ((t * t) * s * ((F(1.0) + s) * (F(1.0) + s)))) * ERGS_CM3_PER_S

    def brems(self): return (F(1.42e-27) * _sqrt(self.t)) * ERGS_CM3_PER_S           # :255-263
    def dbrems(self): return (F(1.42e-27) / (F(2.0) * _sqrt(self.t))) * ERGS_CM3_PER_S

    def _x4(self):
        x = F(2.727) / self.a
        return (x * x) * (x * x)

    def compton(self): return (F(1.017e-37) * self._x4() * (self.t - F(2.727) / self.a)) * ERGS_PER_S   # :265-273
    def dcompton(self): return (F(1.017e-37) * self._x4()) * ERGS_PER_S

    def cooling(self):                                                                 # :275-287
        ne, n0, n1 = self.ne(), self.nh0(), self.nh1()
        return (self.exc_cool() + self.ion_cool()) * ne * n0 + self.rec_cool() * ne * n1 + self.brems() * ne * n1 + self.compton() * ne

    def dcooling(self):                                                                # :289-302
        ne, n0, n1 = self.ne(), self.nh0(), self.nh1()
        return (self.dexc_cool() + self.dion_cool()) * ne * n0 + self.drec_cool() * ne * n1 + self.dbrems() * ne * n1 + self.dcompton() * ne

    def newly_ionized(self, h):                                                        # :312-319
        return (h * self.rate) * (F(1.0) - _exp(-self.nh0() * SIG * self.len))

    def heating(self, h):                                                              # :321-325
        return self.newly_ionized(h) / self.vol * (E_PHOTON - RYDBERG) / h

    def photoionization(self, h):                                                      # :327-332
        return self.newly_ionized(h) / (self.nh0() * self.vol) / h

    def dtemperature(self, h):                                                         # :304-310
        k = (GAMMA - F(1.0)) * MP / (self.rho * KB)
        lam = self.heating(h) - self.cooling()
        dlam = -self.dcooling()
        mu = self.mu()
        return k * mu * lam * h / (F(1.0) - k * mu * dlam * h)

    def dxhii(self, h):                                                                # :334-354
        nh, ne = self.nh(), self.ne()
        alpha, dalpha, beta, dbeta = self.alpha(), self.dalpha(), self.beta(), self.dbeta()
        c = beta * ne + self.photoionization(h)
        mu = self.mu()
        d = alpha * ne
        dcdx = nh * beta - ne * self.t * mu * F(1.0) * dbeta
        dddx = nh * alpha - ne * self.t * mu * F(1.0) * dalpha
        j = dcdx - (c + d) - self.x * (dcdx + dddx)
        return h * (c - self.x * (c + d)) / (F(1.0) - j * h)

    def clamp(self):                                                                   # :356-369
        lo = self.floor[1] if self.floor else F(1e-10)
        self.x = min(max(self.x, lo), F(1.0) - F(1e-10))
        if self.floor and self.t < self.floor[0]:
            self.t = self.floor[0]

    @staticmethod
    def update(value, change, max_allowed, h):                                         # :444-461 (f64::min ignores a NaN)
        rel = abs(change / value)
        rel = INV_EPS if rel != rel else min(rel, INV_EPS)
        if rel > max_allowed:
            return None
        return value + change, h * (max_allowed / rel)

    def try_update(self, h, safety):                                                   # :371-392
        r = self.update(self.t, self.dtemperature(h), safety, h)
        if r is None:
            return None
        self.t, t_rec = r
        r = self.update(self.x, self.dxhii(h), safety, h)
        if r is None:
            return None
        self.x, x_rec = r
        self.clamp()
        return t_rec if t_rec < x_rec else x_rec

    def internal(self, h, safety, depth):                                              # :394-424
        self.clamp()
        saved = (self.t, self.x)
        if depth > 100:
            raise OverflowError("TimestepConvergenceFailed")
        self.attempts += 1
        self.max_depth = max(self.max_depth, depth)
        r = self.try_update(h, safety)
        if r is None:
            self.t, self.x = saved
            self.internal(h / F(2.0), safety, depth + 1)
            return self.internal(h / F(2.0), safety, depth + 1)
        return r

    def perform(self, h, safety):                                                      # :426-441
        try:
            return self.internal(F(h), F(safety), 0), False
        except OverflowError:
            return F(h) / F(10.0), True


def test_python_recursive_chemistry_equals_the_c_oracle():
    z = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "chemistry_cells.npz")
    cols = [z[k] for k in ("xhii", "temperature", "density", "volume", "length", "rate", "timestep")]
    # every eighth cell plus the heaviest substeppers of the set
    heavy = np.argsort(z["pc1_attempts"])[-12:]
    pick = np.unique(np.concatenate([np.arange(0, len(cols[0]), 8), heavy]))
    assert z["pc1_attempts"][pick].max() > 50
    with np.errstate(all="ignore"):
        for pc in (False, True):
            ref = oracle.chemistry(*(c[pick] for c in cols), scale_factor=0.5, safety=0.1, prevent_cooling=pc)
            for j, i in enumerate(pick):
                s = PySolver(*(c[i] for c in cols[:6]), 0.5, pc)
                timescale, failed = s.perform(cols[6][i], 0.1)
                assert s.attempts == ref["attempts"][j] and s.max_depth == ref["depth"][j], (i, pc)
                assert failed == (ref["process"][j] == -1)
                for got, want in ((s.x, ref["xhii"][j]), (s.t, ref["temperature"][j]), (timescale, ref["timescale"][j])):
                    assert got == pytest.approx(want, rel=1e-12, abs=0.0), (i, pc)


# ---------------------------------------------------------------------------------------------
# The whole of Sweep::run_sweeps, a second time: timestep levels, partial active sets, periodic faces read in the
# order Rust's BinaryHeap pops the tasks (src/sweep/mod.rs:258-589, task.rs:25-35, active_list.rs, timestep_state.rs).
# ---------------------------------------------------------------------------------------------
class PyBinaryHeap:
    """std::collections::BinaryHeap of (key, payload) compared by key only."""

    def __init__(self, items):
        self.data = list(items)
        n = len(self.data) // 2
        while n > 0:
            n -= 1
            self._sift_down_range(n, len(self.data))

    def _sift_down_range(self, pos, end):
        d = self.data
        elem = d[pos]
        child = 2 * pos + 1
        while end >= 2 and child <= end - 2:
            if d[child][0] <= d[child + 1][0]:
                child += 1
            if elem[0] >= d[child][0]:
                d[pos] = elem
                return
            d[pos] = d[child]
            pos = child
            child = 2 * pos + 1
        if child == end - 1 and elem[0] < d[child][0]:
            d[pos] = d[child]
            pos = child
        d[pos] = elem

    def _sift_up(self, start, pos):
        d = self.data
        elem = d[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if elem[0] <= d[parent][0]:
                break
            d[pos] = d[parent]
            pos = parent
        d[pos] = elem

    def push(self, item):
        self.data.append(item)
        self._sift_up(0, len(self.data) - 1)

    def pop(self):
        d = self.data
        if not d:
            return None
        item = d.pop()
        if d:
            item, d[0] = d[0], item
            end, pos, elem, child = len(d), 0, d[0], 1     # sift_down_to_bottom(0)
            while end >= 2 and child <= end - 2:
                if d[child][0] <= d[child + 1][0]:
                    child += 1
                d[pos] = d[child]
                pos = child
                child = 2 * pos + 1
            if child == end - 1:
                d[pos] = d[child]
                pos = child
            d[pos] = elem
            self._sift_up(0, pos)
        return item


class PySweep:
    def __init__(self, params, g, density, ionized_hydrogen_fraction, temperature, source, scale_factor=1.0, lagged=False):
        # lagged: a task reads periodic_source as it was when the single sweep started (the order-independent definition
        # the CUDA path implements) instead of whatever has accumulated when the heap pops it
        self.g, self.p, self.a, self.lagged = g, params, scale_factor, lagged
        self.dirs = Directions.from_spec(params.directions).xyz
        self.N, self.D, self.L = g.n_cells, len(self.dirs), params.num_timestep_levels
        self.off = g.face_offsets.astype(np.int64)
        self.rho, self.src = density.astype(np.float64), source.astype(np.float64)
        self.x, self.T = ionized_hydrogen_fraction.astype(np.float64).copy(), temperature.astype(np.float64).copy()
        self.inc, self.out, self.per = (np.zeros((self.N, self.D)) for _ in range(3))
        self.prev, self.tau, self.ts = np.zeros(self.N), np.zeros(self.N), np.zeros(self.N)
        self.level = np.full(self.N, self.L - 1, dtype=np.int64)
        self.lowest, self.first_done = self.L - 1, False
        self.dot = np.empty((self.D, g.n_faces))
        for d, (dx, dy, dz) in enumerate(self.dirs):       # glam DVec3::dot: x*x + y*y + z*z, left to right
            self.dot[d] = (g.face_normal[:, 0] * dx + g.face_normal[:, 1] * dy) + g.face_normal[:, 2] * dz

    def active_in_bin_order(self, cur):                    # active_list.rs:55-66, 143-153
        return [c for lvl in range(cur, self.L) for c in range(self.N) if self.level[c] == lvl]

    def single_sweep(self, cur):                           # mod.rs:274-289
        g, off, D = self.g, self.off, self.D
        act = self.active_in_bin_order(cur)
        is_active = self.level >= cur
        miss = {}
        for c in act:                                      # init_counts :346-386
            for d in range(D):
                miss[c, d] = sum(1 for f in range(off[c], off[c + 1])
                                 if self.dot[d, f] < 0.0 and g.face_kind[f] == 0 and is_active[g.face_neighbour[f]])
        heap = PyBinaryHeap([(d, c) for d in range(D) for c in act if miss[c, d] == 0])   # get_initial_tasks :388-398
        thr = self.p.significant_rate_threshold
        per_read = self.per.copy() if self.lagged else self.per
        while True:                                        # solve :291-314
            task = heap.pop()
            if task is None:
                break
            d, c = task
            self.inc[c, d] = max(self.inc[c, d], 0.0)      # make_positive :418
            total = (self.inc[c, d] + self.src[c] / D) + per_read[c, d]
            nhi = self.rho[c] / U.PROTON_MASS * (1.0 - self.x[c])
            o = 0.0 if total < thr else total * math.exp(-nhi * SIGMA * g.cell_size[c])
            delta = o - self.out[c, d]
            self.out[c, d] = o
            down = [f for f in range(off[c], off[c + 1]) if self.dot[d, f] > 0.0]
            ttot = 0.0
            for f in down:
                ttot += g.face_area[f] * self.dot[d, f]
            for f in down:
                share = delta * ((g.face_area[f] * self.dot[d, f]) / ttot)
                nb, kind = g.face_neighbour[f], g.face_kind[f]
                if kind == 0:                              # handle_local_neighbour :487-503
                    self.inc[nb, d] += share
                    if is_active[nb]:
                        miss[nb, d] -= 1
                        if miss[nb, d] == 0:
                            heap.push((d, nb))
                elif kind == 2:                            # handle_local_periodic_neighbour :505-513
                    self.per[nb, d] += share
        assert all(v == 0 for v in miss.values())
        for c in act:                                      # update_chemistry :549-574
            dt = self.p.max_timestep * 0.5 ** int(self.level[c])
            rate = 0.0
            for d in range(D):
                rate = rate + ((self.inc[c, d] + self.src[c] / D) + self.per[c, d])
            with np.errstate(all="ignore"):
                if abs(rate) < abs(thr):
                    rel = F(0.0)
                else:
                    rel = abs(F(abs(rate - self.prev[c])) / F(rate))
                    rel = INV_EPS if rel != rel else min(rel, INV_EPS)
                self.prev[c] = rate
                t_rate = F(dt) / rel
                s = PySolver(self.x[c], self.T[c], self.rho[c], g.cell_volume[c], g.cell_size[c], rate, self.a,
                             self.p.prevent_cooling)
                t_chem, _ = s.perform(dt, self.p.chemistry_timestep_safety_factor)
            self.x[c], self.T[c], self.ts[c] = s.x, s.t, t_chem
            self.tau[c] = t_rate if t_rate < t_chem else t_chem                       # Timescale::min, timescale.rs:32-38

    def run_sweeps(self):                                  # mod.rs:258-272, timestep_state.rs
        counts = [int((self.level >= l).sum()) for l in range(self.L)]
        num = self.L - self.lowest
        for i in range(2 ** (num - 1)):
            first_bit = (i & -i).bit_length() - 1 if i else num - 1
            cur = self.lowest + (num - 1 - first_bit)
            if counts[cur] > 0:
                self.single_sweep(cur)
        elapsed = self.p.max_timestep * 0.5 ** self.lowest
        if self.first_done and self.lowest > 0:
            self.lowest -= 1
        self.first_done = True
        with np.errstate(all="ignore"):
            for c in range(self.N):                        # update_timestep_levels :576-589, timestep_level.rs:27-36
                ratio = F(self.p.max_timestep) / (F(self.p.timestep_safety_factor) * F(self.tau[c]))
                lv = np.ceil(np.log2(ratio))
                lv = 0 if (lv != lv or lv < 0) else (2 ** 62 if lv == np.inf else int(lv))   # Rust `as usize` saturates
                self.level[c] = max(min(lv, self.L - 1), self.lowest)
        return elapsed


@pytest.mark.parametrize("kind,n,periodic,n_dirs,n_levels,lagged", [
    ("voronoi", 3, True, 16, 2, False), ("cartesian", 4, True, 21, 3, False), ("jittered", 4, False, 16, 3, False),
    ("voronoi", 3, True, 16, 2, True), ("jittered", 3, True, 21, 2, True), ("cartesian", 4, True, 21, 3, True)])
def test_python_run_sweeps_equals_the_c_oracle(kind, n, periodic, n_dirs, n_levels, lagged):
    params, g, f = make_problem(kind, n, periodic, n_dirs=n_dirs, n_levels=n_levels, source_rate=3e51, max_timestep_myr=0.5)
    mine = PySweep(params, g, **f, lagged=lagged)
    ref = oracle.OracleSweep(params, g, **f, periodic_mode=oracle.PERIODIC_LAGGED if lagged else oracle.PERIODIC_HEAP)
    for step in range(n_levels + 2):
        assert mine.run_sweeps() == ref.run_sweeps()
        assert np.array_equal(mine.level, ref.levels()), step
        for name, arr in (("ionized_hydrogen_fraction", mine.x), ("temperature", mine.T), ("timestep", mine.ts),
                          ("change_timescale", mine.tau), ("previous_rate", mine.prev)):
            want = ref.read(name)
            np.testing.assert_allclose(arr, want, rtol=1e-10, atol=1e-12 * np.abs(want[np.isfinite(want)]).max(), err_msg=f"{name} step {step}")
        for name, arr in (("outgoing", mine.out), ("incoming", mine.inc), ("periodic", mine.per)):
            want = ref.dir_state(name)
            np.testing.assert_allclose(arr, want, rtol=1e-10, atol=1e-12 * max(np.abs(want).max(), 1e-300), err_msg=f"{name} step {step}")
    if periodic and not lagged:
        assert ref.stat("nonlagged_periodic_reads") > 0   # the heap order did matter (on the Cartesian grid: in the partial sweeps)
