"""Cell-range sharding of a population over several GPUs, host side.

The grid is cut into `nranks` contiguous cell-index ranges balanced by AGENTS, not cells (the sea is empty);
rank r owns cells [begin[r], begin[r+1]).  `connect` wires a `GpuPopulation` to its peers: rank 0 creates the
NCCL id through the C ABI, `torch.distributed` (any backend; gloo is enough) only carries those 128 bytes.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np


# fixed cost of an occupied cell in agent-equivalents: one step costs about a * agents + b * occupied cells on the device
# (measured on B200: 1e8 agents -> 1.90 ms, 1e7 agents -> 0.50 ms on the same 459k land cells, so b / a = 44)
CELL_COST_AGENTS = 44
# ... and an EMPTY cell of the range is not free either: both passes hand out cells, not agents (8 B200, 1e8 agents: the ranks
# whose ranges hold 80,000 more sea cells than the others need 22-28 us more per step, at 0.028 ns per agent-step: 11)
EMPTY_CELL_COST_AGENTS = 11


def partition_cells(agents_per_cell, nranks: int, cell_cost: float = CELL_COST_AGENTS, empty_cost: float = None) -> np.ndarray:
    """Boundaries (int32, nranks+1) of contiguous cell ranges of about equal step COST: the agents of a range plus
    `cell_cost` agent-equivalents for every occupied cell and `empty_cost` for every empty one (cell_cost=0: balanced by
    agents alone)."""
    cnt = np.asarray(agents_per_cell, dtype=np.int64)
    if empty_cost is None:
        empty_cost = EMPTY_CELL_COST_AGENTS if cell_cost > 0 else 0
    cnt = cnt + np.int64(cell_cost) * (cnt > 0) + np.int64(empty_cost) * (cnt == 0)
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


def connect(pop, begin, rank: int, nranks: int, p2p: bool | None = None):
    """Join `pop` (a GpuPopulation) to the sharded run; call before add_agents.

    The host side only carries small tables between the ranks (torch.distributed, any backend): the 128-byte NCCL id
    and, for the exchange over peer memory (default; QHG_P2P=0 or p2p=False keeps the NCCL calls), the CUDA IPC
    handles of every rank's exchange buffers."""
    import os
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
    if p2p is None:
        p2p = os.environ.get("QHG_P2P", "1") != "0"
    if p2p and nranks > 1:
        # every rank must end up with the same exchange: if one cannot export or map the buffers (no peer access, IPC
        # not permitted in the container), all of them keep the NCCL calls
        ok = 1
        try:
            mine = torch.tensor(list(pop.comm_p2p_handle()), dtype=torch.uint8)
        except Exception as e:  # noqa: BLE001
            print(f"[qhg4_b200] rank {rank}: no peer-memory exchange ({e}); using NCCL", file=sys.stderr)
            mine, ok = torch.zeros(128, dtype=torch.uint8), 0
        table = [torch.zeros_like(mine) for _ in range(nranks)]
        dist.all_gather(table, mine)
        flag = torch.tensor([ok])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 1:
            try:
                pop.comm_p2p_connect(b"".join(bytes(x.tolist()) for x in table))
            except Exception as e:  # noqa: BLE001
                print(f"[qhg4_b200] rank {rank}: could not map the peers' buffers ({e}); using NCCL", file=sys.stderr)
                ok = 0
        flag = torch.tensor([ok])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 0:
            pop.comm_p2p_connect(None)


def rebalance(pop, make_population, rank: int, nranks: int, workdir: str, p2p: bool | None = None):
    """Re-split a sharded run over new cell ranges when its load has shifted ("may need re-splitting after env
    events").  Collective over all ranks, between two steps:

      1. every rank dumps its state (`qhgb_dump_state`) into `workdir` (a directory all ranks see) and frees its GPU memory;
      2. the per-cell counts of all ranks give the new cost-balanced ranges (`partition_cells`);
      3. `make_population()` builds a population configured like the old one -- cells, the CURRENT environment arrays,
         attributes, priorities, seed; not yet connected -- and `prepare(new_pop)` (optional attribute of the factory) sets what
         must follow the connection (navigation tables);
      4. the new population joins the run with the new ranges and restores from ALL ranks' dumps, keeping the agents (and
         genome rows) of its own range (`qhgb_restore_state` with several files).

    Returns (new population, new boundaries).  The continued run is bit-identical to one that was never re-split: results do
    not depend on which rank owns a cell."""
    import os
    import torch
    import torch.distributed as dist
    cnt = torch.from_numpy(pop.counts().astype(np.int64))
    if nranks > 1:
        dist.all_reduce(cnt)
    begin = partition_cells(cnt.numpy(), nranks)
    path = os.path.join(workdir, f"rebalance_rank{rank}.qhgb")
    pop.dump_state(path)
    pop.close()
    if nranks > 1:
        dist.barrier()
    new = make_population()
    connect(new, begin, rank, nranks, p2p=p2p)
    prepare = getattr(make_population, "prepare", None)
    if prepare is not None:
        prepare(new)
    new.restore_state("\n".join(os.path.join(workdir, f"rebalance_rank{r}.qhgb") for r in range(nranks)))
    if nranks > 1:
        dist.barrier()
    if rank == 0:
        for r in range(nranks):
            try:
                os.unlink(os.path.join(workdir, f"rebalance_rank{r}.qhgb"))
            except OSError:
                pass
    return new, begin
