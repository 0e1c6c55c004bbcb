"""torch.distributed plumbing for multi-GPU contexts (one process per GPU).

torch.distributed is used only to hand the NCCL unique id to every rank and to merge
result fields for the caller; halo exchange and all-reduces inside the solvers are
issued by the library itself on its own NCCL communicator.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def make_comm(dist=None):
    """Return ``(rank, world, nccl_id_bytes)`` for ``Context(..., comm=...)``.

    ``dist``: an initialised ``torch.distributed`` module (nccl or gloo backend) or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return (0, 1, b"")
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_char * 128)()
    if rank == 0:
        L.check(L.load().cfem_nccl_unique_id(C.addressof(buf)))
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    return (rank, world, bytes(t.cpu().numpy().tobytes()))


def make_partition(dist, x, cells, method="metis"):
    """Owning rank of every node (int32, caller numbering) for ``Context(..., partition=...)``: computed on rank 0
    (METIS k-way on the nodal graph, or ``"hilbert"`` = equal ranges of the Hilbert order) and broadcast."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n = np.asarray(x).shape[0]
    if rank == 0:
        part = torch.as_tensor(L.host_partition(x, cells, world, method), device=dev)
    else:
        part = torch.empty(n, dtype=torch.int32, device=dev)
    dist.broadcast(part, src=0)
    return part.cpu().numpy()


def allgather_field(ctx, local_result, dist):
    """Merge per-rank results (valid at owned dofs) into the full field on every rank."""
    if ctx.world == 1:
        return local_result
    import torch

    owned = ctx.owned_dofs()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    full = torch.zeros(ctx.n, dtype=torch.float64, device=dev)
    full[torch.as_tensor(owned.astype(np.int64), device=dev)] = torch.as_tensor(local_result[owned], device=dev)
    dist.all_reduce(full)
    return full.cpu().numpy()
