"""Horizontal sharding of the column-stacked state over the GPUs of one box (one process per GPU).

The reference never communicates state itself: `dss!` is a user closure and reductions are local
(src/functions.jl:99, newtons_method.jl:6-10).  This module is what a distributed state
type plugs into those hooks on B200: a slab partition of the slow horizontal axis, an NCCL communicator
attached to the library context (halo exchange inside `dss!`, all-reduced Krylov/Newton reductions) and the
reduction window that restricts norms/dots to owned rows.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L


@dataclass(frozen=True)
class SlabPartition:
    """Contiguous blocks of `ny_global/nranks` rows per rank (level-fastest layout => one contiguous slab),
    periodic ring of neighbours along the slow horizontal axis."""
    ny_global: int
    nranks: int
    rank: int

    def __post_init__(self):
        if self.ny_global % self.nranks:
            raise ValueError("ny_global must be divisible by the number of ranks")

    @property
    def ny_local(self) -> int:
        return self.ny_global // self.nranks

    @property
    def row0(self) -> int:
        return self.rank * self.ny_local

    @property
    def lo(self) -> int:  # neighbour owning the rows below mine
        return (self.rank - 1) % self.nranks

    @property
    def hi(self) -> int:
        return (self.rank + 1) % self.nranks

    def local_rows(self, halo: int = 1) -> np.ndarray:
        """Global row index of every local row (ghost rows included, periodic wrap)."""
        return np.arange(self.row0 - halo, self.row0 + self.ny_local + halo) % self.ny_global

    def owned_window(self, nx: int, nlev: int, halo: int = 1):
        """(offset, count) of the owned elements inside the local array."""
        return halo * nx * nlev, self.ny_local * nx * nlev


    def reduction_layout(self, nx: int, nlev: int, halo: int = 1, max_chunk: int = 8192):
        """(local_offset, count, global_offset, global_count, chunk) for `Context.set_reduction_layout`: the chunk is the
        largest power of two <= max_chunk that divides a row, so every slab of whole rows owns whole chunks and the
        reduction tree is the same for every rank count."""
        row = nx * nlev
        chunk = 1
        while chunk * 2 <= max_chunk and row % (chunk * 2) == 0:
            chunk *= 2
        if chunk < 4:
            raise ValueError("row size must be a multiple of 4 elements for the shard-invariant reduction layout")
        return halo * row, self.ny_local * row, self.row0 * row, self.ny_global * row, chunk


def attach_communicator(ctx, dist=None):
    """Create the NCCL communicator of the library context: rank 0 makes the unique id, torch.distributed
    (the plumbing; MPI/ClimaComms in the Julia host) broadcasts it, every rank joins."""
    if dist is None:
        import torch.distributed as dist
    import torch
    nranks, rank = dist.get_world_size(), dist.get_rank()
    if nranks == 1:
        return ctx
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        L.check(None, ctx.lib.cts_comm_get_unique_id(buf), "cts_comm_get_unique_id")
    dev = f"cuda:{ctx.device}" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    ctx.check(ctx.lib.cts_comm_init(ctx.handle, nranks, rank, raw), "cts_comm_init")
    ctx.nranks, ctx.rank = nranks, rank
    return ctx
