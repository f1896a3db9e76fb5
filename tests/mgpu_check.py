#!/usr/bin/env python
"""Sharded run on N GPUs against the unsharded CPU oracle, bit-exact.  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qhg4_b200 import sharding  # noqa: E402
from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_population  # noqa: E402
from qhg4_b200.params import seed_state, tut_environ_alt  # noqa: E402
from qhg4_b200.population import GpuPopulation  # noqa: E402

FIELDS = ("cell", "id", "birth", "gender", "age", "last_birth", "life")


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nbr, xyz = make_ico_grid(31)
    alt = synthetic_altitude(xyz, seed=3)
    pop = synthetic_population(300000, alt, seed=5, fertile=True)
    par, st = tut_environ_alt(45.0), seed_state(21)
    begin = sharding.partition_cells(np.bincount(pop["cell"], minlength=len(nbr)), world)
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, device=int(os.environ.get("LOCAL_RANK", 0)))
    sharding.connect(g, begin, rank, world)
    g.add_agents(pop)
    g.pre_loop()
    o = None
    if rank == 0:
        from oracle import port
        o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st)
        o.add_agents(pop)
        o.start()
    nsteps, moved = 12, 0
    for k in range(nsteps):
        g.step(float(k))
        moved += g.comm_traffic()[0]
        mine = g.agents()
        lo, hi = begin[rank], begin[rank + 1]
        assert np.all((mine["cell"] >= lo) & (mine["cell"] < hi)), "an agent sits on a rank that does not own its cell"
        parts = [None] * world
        dist.gather_object({f: mine[f] for f in FIELDS}, parts if rank == 0 else None, dst=0)
        cnts = [None] * world
        dist.gather_object(g.counts(), cnts if rank == 0 else None, dst=0)
        if rank == 0:
            o.step(float(k))
            allg = {f: np.concatenate([p[f] for p in parts]) for f in FIELDS}
            oa = o.agents()
            og, oo = np.argsort(allg["id"]), np.argsort(oa["id"])
            assert len(allg["id"]) == o.num_agents(), (k, len(allg["id"]), o.num_agents())
            for f in FIELDS:
                assert np.array_equal(allg[f][og], oa[f][oo]), (k, f)
            assert np.array_equal(np.sum(cnts, axis=0), o.counts()), k
    # the same through qhgb_run: the steps are queued on every rank without a host round trip (peer-memory exchange; the NCCL
    # exchange needs the host in every step and runs them one by one) -- still bit-exact
    nq = 6
    a0, s0, _ = g.run_totals()
    g.run(float(nsteps), nq)
    a1, s1, _ = g.run_totals()
    moved += s1 - s0
    mine = g.agents()
    parts = [None] * world
    dist.gather_object({f: mine[f] for f in FIELDS}, parts if rank == 0 else None, dst=0)
    asteps = [None] * world
    dist.gather_object(a1 - a0, asteps if rank == 0 else None, dst=0)
    if rank == 0:
        expect = 0
        for k in range(nsteps, nsteps + nq):
            expect += o.num_agents()
            o.step(float(k))
        allg = {f: np.concatenate([p[f] for p in parts]) for f in FIELDS}
        oa = o.agents()
        og, oo = np.argsort(allg["id"]), np.argsort(oa["id"])
        assert len(allg["id"]) == o.num_agents(), ("queued", len(allg["id"]), o.num_agents())
        for f in FIELDS:
            assert np.array_equal(allg[f][og], oa[f][oo]), ("queued", f)
        assert sum(asteps) == expect, (asteps, expect)
        nsteps += nq
    tot = torch.tensor([moved])
    dist.all_reduce(tot)
    if rank == 0:
        how = "peer-memory" if os.environ.get("QHG_P2P", "1") != "0" else "nccl"
        print(f"mgpu_check ok: {world} ranks, {nsteps} steps (the last {nq} queued by qhgb_run), {o.num_agents()} agents, {int(tot)} cross-rank migrations ({how} exchange), "
              f"bit-exact vs the unsharded oracle")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
