"""Direction sharding on the CUDA path: W ranks, emulated as W host threads that share one device,
must reproduce the single-rank result -- with the all-reduce hook alone (replicated chemistry) and with the
reduce-scatter / all-gather hook (chemistry sliced by cells).  The hooks here exchange the device buffers of the
W handles directly; on the GPU box they are NCCL calls (subsweep_b200/distributed.py)."""
import threading

import numpy as np
import pytest

from helpers import assert_close, make_problem
from subsweep_b200 import Sweep, capi
from subsweep_b200.distributed import tensor_from_pointer

pytestmark = pytest.mark.gpu

FIELDS = ("ionized_hydrogen_fraction", "temperature", "timestep", "change_timescale", "previous_rate", "photon_rate")


class Exchange:
    """In-process stand-in for NCCL: every 'rank' is a thread; rank 0 combines the buffers of all ranks.  The
    library keeps its direction table in per-process constant memory (one solver per process, like the reference's
    NonSend resource), so the threads take turns: `turn` is held whenever a thread is inside the library and handed
    over while it waits in a hook (SSW_FLAG_SHARED_DEVICE makes the library re-bind its table afterwards)."""

    def __init__(self, world):
        import torch
        self.torch = torch
        self.world = world
        self.barrier = threading.Barrier(world, timeout=120)
        self.turn = threading.Lock()
        self.slots = [None] * world
        self.dev = torch.device("cuda", 0)

    def _meet(self, rank, ptr, n, stream, combine):
        torch = self.torch
        if stream:
            torch.cuda.ExternalStream(stream, device=self.dev).synchronize()
        self.slots[rank] = tensor_from_pointer(ptr, n, self.dev)
        self.turn.release()
        try:
            self.barrier.wait()
            if rank == 0:
                combine(self.slots)
                torch.cuda.synchronize(self.dev)
            self.barrier.wait()
        finally:
            self.turn.acquire()

    def allreduce(self, rank):
        def fn(ptr, n, stream):
            def combine(ts):
                total = ts[0].clone()
                for t in ts[1:]:
                    total += t
                for t in ts:
                    t.copy_(total)
            self._meet(rank, ptr, n, stream, combine)
        return fn

    def collectives(self, rank):
        def fn(op, ptr, n_per, stream):
            def combine(ts):
                if op == 1:      # reduce-scatter: rank r keeps the sum of everybody's chunk r
                    total = ts[0].clone()
                    for t in ts[1:]:
                        total += t
                    for r, t in enumerate(ts):
                        t[r * n_per:(r + 1) * n_per].copy_(total[r * n_per:(r + 1) * n_per])
                else:            # all-gather: chunk r of rank r goes to everybody
                    for r, src in enumerate(ts):
                        for q, dst in enumerate(ts):
                            if q != r:
                                dst[r * n_per:(r + 1) * n_per].copy_(src[r * n_per:(r + 1) * n_per])
            self._meet(rank, ptr, n_per * self.world, stream, combine)
        return fn


def run_sharded(params, g, f, world, steps, sliced):
    ex = Exchange(world)
    out, errors = [None] * world, []

    def work(rank):
        try:
            with ex.turn:
                s = Sweep(params, g, **f, rank=rank, world_size=world, allreduce=ex.allreduce(rank),
                          collectives=ex.collectives(rank) if sliced else None, flags=capi.FLAG_SHARED_DEVICE)
            for _ in range(steps):
                with ex.turn:
                    s.run_sweeps()
            res = {}
            for k in FIELDS:
                with ex.turn:
                    res[k] = s.read(k)
            with ex.turn:
                res["levels"] = s.levels()
                res["outgoing"] = s.dir_state("outgoing")
                res["macro_tiles"] = s.stat("patch_macro_tiles")
                res["chem_cells"] = s.stat("chem_cells")
                s.close()
            out[rank] = res
        except Exception as exc:   # noqa: BLE001
            errors.append(exc)
            ex.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out


def run_peers(params, g, f, world, steps):
    """W ranks as W host threads, one handle each on the same device, exchanging through peer-mapped arenas
    (ssw_peer_arena / ssw_peer_attach): no hooks, the threads run concurrently and meet only on the device."""
    barrier = threading.Barrier(world, timeout=120)
    create = threading.Lock()
    sweeps, out, errors = [None] * world, [None] * world, []

    def work(rank):
        try:
            with create:
                sweeps[rank] = Sweep(params, g, **f, rank=rank, world_size=world, flags=capi.FLAG_SHARED_DEVICE)
            barrier.wait()
            s = sweeps[rank]
            s.peer_attach([sw.peer_arena()[0] for sw in sweeps])
            barrier.wait()
            for _ in range(steps):
                s.run_sweeps()
            res = {k: s.read(k) for k in FIELDS}          # photon_rate is a collective: same order on every rank
            res["levels"] = s.levels()
            res["outgoing"] = s.dir_state("outgoing")
            res["macro_tiles"] = s.stat("patch_macro_tiles")
            res["chem_cells"] = s.stat("chem_cells")
            res["series"] = s.time_series()
            barrier.wait()
            s.close()
            out[rank] = res
        except Exception as exc:   # noqa: BLE001
            errors.append(exc)
            barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("kind,n,periodic", [("cartesian", 12, True), ("voronoi", 8, False), ("voronoi", 8, True)])
def test_peer_mapped_shards_reproduce_the_single_rank_result(cuda_lib, monkeypatch, kind, n, periodic, world):
    """The NVLink path of direction sharding (csrc/peer.cuh): rate partials pushed into the owners' receive buffers,
    chemistry on the owner, absorption factors and timestep levels pushed to every rank."""
    monkeypatch.setenv("SSW_PATCH_CELLS", "64")
    params, g, f = make_problem(kind, n, periodic, n_dirs=21, n_levels=3, max_timestep_myr=0.25)
    steps = 5
    one = Sweep(params, g, **f)
    for _ in range(steps):
        one.run_sweeps()
    ranks = run_peers(params, g, f, world, steps)
    for r in ranks[1:]:   # every rank assembles the same cell state, bit for bit
        for k in FIELDS + ("levels",):
            assert np.array_equal(r[k], ranks[0][k], equal_nan=True), k
    assert np.array_equal(ranks[0]["levels"], one.levels())
    for k in FIELDS:
        b = one.read(k)
        floor = 1e-7 * np.nanmax(np.abs(b)) if k in ("previous_rate", "photon_rate") else 0.0
        assert_close(ranks[0][k], b, 1e-9, floor=floor, what=k)
    out = np.concatenate([r["outgoing"] for r in ranks], axis=1)
    b = one.dir_state("outgoing")
    assert_close(out, b, 1e-9, floor=1e-7 * max(np.abs(b).max(), 1e-300), what="outgoing")
    assert sum(r["chem_cells"] for r in ranks) == one.stat("chem_cells")     # every cell update ran on exactly one rank
    ts = one.time_series()
    for key, v in ranks[-1]["series"].items():
        if np.isfinite(ts[key]):
            assert abs(v - ts[key]) <= 1e-9 * abs(ts[key]), key


@pytest.mark.parametrize("world,sliced", [(2, False), (2, True), (3, True)])
@pytest.mark.parametrize("kind,n,periodic", [("cartesian", 12, True), ("voronoi", 8, False)])
def test_direction_shards_reproduce_the_single_rank_result(cuda_lib, monkeypatch, kind, n, periodic, world, sliced):
    monkeypatch.setenv("SSW_PATCH_CELLS", "64")
    params, g, f = make_problem(kind, n, periodic, n_dirs=21, n_levels=3, max_timestep_myr=0.25)
    steps = 5
    one = Sweep(params, g, **f)
    for _ in range(steps):
        one.run_sweeps()
    ranks = run_sharded(params, g, f, world, steps, sliced)
    # every rank holds the same cell state, bit for bit
    for r in ranks[1:]:
        for k in FIELDS[:5] + ("levels",):
            assert np.array_equal(r[k], ranks[0][k], equal_nan=True), k
    assert np.array_equal(ranks[0]["levels"], one.levels())
    for k in FIELDS:
        b = one.read(k)
        floor = 1e-7 * np.nanmax(np.abs(b)) if k in ("previous_rate", "photon_rate") else 0.0
        assert_close(ranks[0][k], b, 1e-9, floor=floor, what=k)
    # the per-direction state of the shards, side by side, is the single rank's
    out = np.concatenate([r["outgoing"] for r in ranks], axis=1)
    b = one.dir_state("outgoing")
    assert_close(out, b, 1e-9, floor=1e-7 * max(np.abs(b).max(), 1e-300), what="outgoing")
    if kind == "cartesian":
        assert all(r["macro_tiles"] > 0 for r in ranks)      # the shards run the patch-ordered form
    if sliced:   # all-cells sweeps update every cell on exactly one rank
        assert sum(r["chem_cells"] for r in ranks) < world * one.stat("chem_cells")
