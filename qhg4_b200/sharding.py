"""Cell-range sharding of a population over several GPUs, host side.

The grid is cut into `nranks` contiguous cell-index ranges balanced by AGENTS, not cells (the sea is empty);
rank r owns cells [begin[r], begin[r+1]).  `connect` wires a `GpuPopulation` to its peers: rank 0 creates the
NCCL id through the C ABI, `torch.distributed` (any backend; gloo is enough) only carries those 128 bytes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def partition_cells(agents_per_cell, nranks: int) -> np.ndarray:
    """Boundaries (int32, nranks+1) of contiguous cell ranges holding about the same number of agents each."""
    cnt = np.asarray(agents_per_cell, dtype=np.int64)
    ncell = len(cnt)
    cum = np.concatenate([[0], np.cumsum(cnt)])
    total = cum[-1]
    begin = np.zeros(nranks + 1, dtype=np.int32)
    begin[nranks] = ncell
    for r in range(1, nranks):
        if total > 0:
            begin[r] = int(np.searchsorted(cum, total * r / nranks, side="left"))
        else:
            begin[r] = ncell * r // nranks
        begin[r] = max(begin[r], begin[r - 1])
    return np.minimum(begin, ncell).astype(np.int32)


def owner_of(cells, begin) -> np.ndarray:
    """Rank owning each cell index."""
    return (np.searchsorted(np.asarray(begin), np.asarray(cells), side="right") - 1).astype(np.int32)


def connect(pop, begin, rank: int, nranks: int):
    """Create the NCCL communicator of `pop` (a GpuPopulation); call before add_agents."""
    import torch
    import torch.distributed as dist
    buf = (C.c_char * 128)()
    if rank == 0:
        from .capi import check
        check(pop.L.qhgb_comm_get_unique_id(buf, 128), "qhgb_comm_get_unique_id")
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
    if nranks > 1:
        dist.broadcast(t, src=0)
    uid = bytes(t.tolist())
    pop.comm_init(rank, nranks, uid, begin)
