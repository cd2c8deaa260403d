"""torch.distributed plumbing for direction sharding (one process per GPU).

The library sweeps this rank's shard of the directions and hands the per-cell partial rates to an
all-reduce hook (``ssw_set_allreduce``).  This module provides that hook on top of
``torch.distributed`` -- NCCL over NVLink on the GPU box, gloo in the CPU tests.  With one rank
there is no collective (north_star).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np


class _CudaBuffer:
    """Zero-copy view of a device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def tensor_from_pointer(ptr: int, n: int, device):
    import torch
    device = torch.device(device)
    if device.type == "cuda":
        return torch.as_tensor(_CudaBuffer(ptr, n), device=device)
    buf = (ctypes.c_double * n).from_address(ptr)
    return torch.from_numpy(np.ctypeslib.as_array(buf))


def make_allreduce(device, group=None):
    """Returns ``fn(ptr, n, stream)`` summing ``n`` f64 at ``ptr`` in place over the process group."""
    import torch
    import torch.distributed as dist

    device = torch.device(device)

    def allreduce(ptr: int, n: int, stream) -> None:
        t = tensor_from_pointer(ptr, n, device)
        if device.type == "cuda":
            # stream-ordered on the library's own stream: torch's NCCL stream waits for the work queued
            # there and the stream waits for the collective; no host synchronisation
            ext = torch.cuda.ExternalStream(stream, device=device) if stream else torch.cuda.current_stream(device)
            with torch.cuda.device(device), torch.cuda.stream(ext):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    return allreduce


def make_collectives(device, group=None):
    """Returns ``fn(op, ptr, n_per_rank, stream)``: in-place reduce-scatter (op 1) / all-gather (op 2) of
    ``world_size`` chunks of ``n_per_rank`` f64 at ``ptr`` (ssw_collective_fn), NCCL over NVLink."""
    import torch
    import torch.distributed as dist

    device = torch.device(device)
    scratch = {}

    def collective(op: int, ptr: int, n_per_rank: int, stream) -> None:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        t = tensor_from_pointer(ptr, n_per_rank * world, device)
        mine = t[rank * n_per_rank:(rank + 1) * n_per_rank]
        ext = torch.cuda.ExternalStream(stream, device=device) if stream else torch.cuda.current_stream(device)
        with torch.cuda.device(device), torch.cuda.stream(ext):
            if op == 1:
                out = scratch.get(n_per_rank)
                if out is None:
                    out = scratch[n_per_rank] = torch.empty(n_per_rank, dtype=torch.float64, device=device)
                dist.reduce_scatter_tensor(out, t, op=dist.ReduceOp.SUM, group=group)
                mine.copy_(out)
            elif op == 2:
                dist.all_gather_into_tensor(t, mine, group=group)   # in place: the input is the output's own chunk
            else:
                raise ValueError(f"unknown collective {op}")

    return collective


def attach_peers(sweep, group=None) -> None:
    """Peer-mapped direction sharding (one process per GPU of one box): exchange the ranks' 64-byte CUDA IPC handles
    with ``torch.distributed`` -- the only thing the process group is used for -- and attach them.  From then on the
    library moves rate partials, absorption factors and timestep levels itself over NVLink (csrc/peer.cuh)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    handles = [None] * world
    dist.all_gather_object(handles, sweep.peer_export(), group=group)
    sweep.peer_attach_ipc(handles)
    dist.barrier(group=group)


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank
