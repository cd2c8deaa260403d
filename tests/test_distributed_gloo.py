"""Direction sharding over two ranks on the CPU (gloo): the host-side plumbing of the N > 1 path.

Two processes each own half of the directions, sweep them with the CPU oracle and sum the
per-cell partial rates through ``subsweep_b200.distributed.make_allreduce`` -- the same hook
(pointer, count) -> in-place all-reduce the CUDA library calls with a device pointer on the GPU
box (NCCL there, gloo here).  The sharded run must agree with the single-rank run: the only
difference is the summation order of the rate.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, kind: str, n_dirs: int, out_dir: str) -> None:
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import oracle
    from helpers import make_problem
    from subsweep_b200.distributed import init_from_env, make_allreduce
    r, w, _ = init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    params, g, f = make_problem(kind, 7, True, n_dirs=n_dirs, n_levels=3, max_timestep_myr=0.25)
    s = oracle.OracleSweep(params, g, **f, periodic_mode=oracle.PERIODIC_LAGGED, rank=rank, world_size=world,
                           allreduce=make_allreduce("cpu"))
    elapsed = [s.run_sweeps() for _ in range(5)]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=s.read("ionized_hydrogen_fraction"),
             T=s.read("temperature"), tau=s.read("change_timescale"), levels=s.levels(), elapsed=np.array(elapsed),
             tasks=np.array([s.stat("tasks_solved")]), shard=np.array([s.dir_begin, s.dir_end]),
             photon=s.read("photon_rate"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,n_dirs", [("cartesian", 21), ("voronoi", 16)])
def test_two_rank_direction_sharding_matches_single_rank(tmp_path, kind, n_dirs):
    import torch.multiprocessing as mp
    import oracle
    from helpers import assert_close, make_problem
    port = _free_port()
    mp.start_processes(_worker, args=(2, port, kind, n_dirs, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in (0, 1))
    # the two shards partition the directions
    assert r0["shard"][0] == 0 and r0["shard"][1] == r1["shard"][0] and r1["shard"][1] == n_dirs
    # replicated chemistry: both ranks hold bit-identical cell state after the all-reduce
    for k in ("x", "T", "tau", "levels", "elapsed"):
        assert np.array_equal(r0[k], r1[k]), k

    params, g, f = make_problem(kind, 7, True, n_dirs=n_dirs, n_levels=3, max_timestep_myr=0.25)
    ref = oracle.OracleSweep(params, g, **f, periodic_mode=oracle.PERIODIC_LAGGED)
    elapsed = [ref.run_sweeps() for _ in range(5)]
    assert np.array_equal(r0["elapsed"], np.array(elapsed))
    assert_close(r0["x"], ref.read("ionized_hydrogen_fraction"), 1e-9, what="xHII")
    assert_close(r0["T"], ref.read("temperature"), 1e-9, what="T")
    assert np.array_equal(r0["levels"], ref.levels())
    assert int(r0["tasks"][0] + r1["tasks"][0]) == ref.stat("tasks_solved")
    pr = ref.read("photon_rate")
    assert_close(r0["photon"], pr, 1e-9, floor=1e-7 * np.abs(pr).max(), what="photon_rate")
