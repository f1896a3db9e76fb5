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


def gather_sorted(parts, fields):
    allg = {f: np.concatenate([p[f] for p in parts]) for f in fields}
    order = np.argsort(allg["id"])
    return {f: v[order] for f, v in allg.items()}


def main_genetic(rank, world):
    """OoANavGenPop WITH Navigate (BASELINE config #5's class) sharded over the ranks: genome rows and m_iNumBabies travel with
    the migrants, far jumps cross any number of shard boundaries, a GEO + NAV event in the middle -- agents, genomes and
    NumBabies bit-exact against the unsharded oracle after every step."""
    from qhg4_b200.icogrid import synthetic_climate
    from qhg4_b200.params import ooa_nav_gen
    nbr, xyz = make_ico_grid(15)
    alt = synthetic_altitude(xyz, seed=5)
    env = synthetic_climate(xyz, alt, seed=6)
    pop = synthetic_population(60000, alt, seed=6, fertile=True)
    rng = np.random.default_rng(3)
    land = np.flatnonzero(alt > 0)
    occupied = np.unique(pop["cell"])
    nports = 150
    ports = rng.choice(occupied[occupied > 8], nports, replace=False).astype(np.int32)
    ptr = np.arange(0, 4 * nports + 1, 4, dtype=np.int32)
    dests = rng.choice(land, 4 * nports).astype(np.int32)  # anywhere on the globe: most jumps cross a shard boundary
    dist_km = rng.uniform(100, 700, 4 * nports)
    bridges = rng.choice(occupied, (8, 2), replace=False).astype(np.int32)
    G = 256
    par = ooa_nav_gen(G, 3, 2e-3)
    par.modules["Navigate"] = {"Navigate_decay": "-0.001", "Navigate_dist0": "150.0", "Navigate_prob0": "0.1",
                               "Navigate_min_dens": "0.0", "Navigate_bridge_prob": "0.3"}
    par.prios["Navigate"] = 10
    st = seed_state(43)
    row = 2 * (G // 64)
    gen0 = rng.integers(0, 2 ** 63, size=(len(pop["id"]), row), dtype=np.int64).astype(np.uint64)
    begin = sharding.partition_cells(np.bincount(pop["cell"], minlength=len(nbr)), world)
    lo, hi = begin[rank], begin[rank + 1]
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, env=env, device=int(os.environ.get("LOCAL_RANK", 0)))
    sharding.connect(g, begin, rank, world)
    g.set_navigation(ports, ptr, dests, dist_km, bridges)
    g.add_agents(pop)
    own = (pop["cell"] >= lo) & (pop["cell"] < hi)
    g.set_genomes(gen0[own])
    g.pre_loop()
    o = None
    if rank == 0:
        from oracle import port
        o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
        o.set_navigation(ports, ptr, dests, dist_km, bridges)
        o.add_agents(pop)
        o.set_genomes(gen0)
        o.start()
    fields = FIELDS + ("genome", "nbabies")

    def compare(tag):
        mine = g.agents()
        assert np.all((mine["cell"] >= lo) & (mine["cell"] < hi)), "an agent sits on a rank that does not own its cell"
        gg, gnb = g.genomes(row)
        rec = {f: mine[f] for f in FIELDS} | {"genome": gg, "nbabies": gnb}
        parts = [None] * world
        dist.gather_object(rec, parts if rank == 0 else None, dst=0)
        cnts = [None] * world
        dist.gather_object(g.counts(), cnts if rank == 0 else None, dst=0)
        if rank == 0:
            got = gather_sorted(parts, fields)
            oa = o.agents()
            og, onb = o.genomes(row)
            oo = np.argsort(oa["id"])
            assert len(got["id"]) == o.num_agents(), (tag, len(got["id"]), o.num_agents())
            for f in FIELDS:
                assert np.array_equal(got[f], oa[f][oo]), (tag, f)
            assert np.array_equal(got["genome"], og[oo]), (tag, "genomes")
            assert np.array_equal(got["nbabies"], onb[oo]), (tag, "NumBabies")
            assert np.array_equal(np.sum(cnts, axis=0), o.counts()), tag

    nsteps, moved = 10, 0
    for k in range(nsteps):
        g.step(float(k))
        moved += g.comm_traffic()[0]
        if rank == 0:
            o.step(float(k))
        compare(k)
        if k == 4:  # the sea level rises: some bridges drown, agents of flooded cells die, the jump tables are rebuilt on the flush
            alt2 = alt - 150.0
            for q in ([g, o] if rank == 0 else [g]):
                q.set_env("Altitude", alt2)
                q.update_event(2, 5.0); q.update_event(5, 5.0); q.flush_events(5.0)
            compare("event")
    nq = 5  # queued by qhgb_run without a host round trip (peer-memory exchange)
    a0, s0, _ = g.run_totals()
    g.run(float(nsteps), nq)
    a1, s1, _ = g.run_totals()
    moved += s1 - s0
    if rank == 0:
        for k in range(nsteps, nsteps + nq):
            o.step(float(k))
    compare("queued")
    tot = torch.tensor([moved])
    dist.all_reduce(tot)
    if rank == 0:
        how = "peer-memory" if os.environ.get("QHG_P2P", "1") != "0" else "nccl"
        print(f"mgpu_check ok [genetic]: OoANavGenPop + Navigate, {world} ranks, {nsteps + nq} steps (the last {nq} queued by qhgb_run), {o.num_agents()} agents, "
              f"{int(tot)} cross-rank migrations incl. far jumps ({how} exchange), agents + genomes + NumBabies bit-exact vs the unsharded oracle")


def main_rebalance(rank, world, genetic):
    """A run that starts on a poor split (equal numbers of CELLS per rank: the sea is empty) and is re-split in the middle with
    sharding.rebalance -- every rank dumps, new cost-balanced ranges, new populations restore from all dumps -- then goes on:
    bit-exact against the unsharded oracle before and after, agents (and genomes, NumBabies)."""
    import tempfile
    from qhg4_b200.icogrid import synthetic_climate
    from qhg4_b200.params import ooa_nav_gen
    nbr, xyz = make_ico_grid(15)
    alt = synthetic_altitude(xyz, seed=5)
    env = synthetic_climate(xyz, alt, seed=6) if genetic else None
    pop = synthetic_population(60000, alt, seed=6, fertile=True)
    G = 128
    row = 2 * (G // 64)
    par = ooa_nav_gen(G, -1, 2e-3) if genetic else tut_environ_alt(45.0)
    st = seed_state(51)
    gen0 = np.random.default_rng(4).integers(0, 2 ** 63, size=(len(pop["id"]), row), dtype=np.int64).astype(np.uint64)
    begin = sharding.partition_cells(np.ones(len(nbr), np.int64), world, cell_cost=0)  # by cells: unbalanced on purpose
    dev = int(os.environ.get("LOCAL_RANK", 0))

    def make():
        return GpuPopulation.from_params(par, nbr, alt, state16=st, env=env, device=dev)

    g = make()
    sharding.connect(g, begin, rank, world)
    g.add_agents(pop)
    if genetic:
        own = (pop["cell"] >= begin[rank]) & (pop["cell"] < begin[rank + 1])
        g.set_genomes(gen0[own])
    g.pre_loop()
    o = None
    if rank == 0:
        from oracle import port
        o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
        o.add_agents(pop)
        if genetic:
            o.set_genomes(gen0)
        o.start()

    def compare(tag, lo, hi):
        mine = g.agents()
        assert np.all((mine["cell"] >= lo) & (mine["cell"] < hi)), "an agent sits on a rank that does not own its cell"
        rec = {f: mine[f] for f in FIELDS}
        if genetic:
            rec["genome"], rec["nbabies"] = g.genomes(row)
        parts = [None] * world
        dist.gather_object(rec, parts if rank == 0 else None, dst=0)
        if rank == 0:
            got = gather_sorted(parts, tuple(rec.keys()))
            oa = o.agents()
            oo = np.argsort(oa["id"])
            assert len(got["id"]) == o.num_agents(), (tag, len(got["id"]), o.num_agents())
            for f in FIELDS:
                assert np.array_equal(got[f], oa[f][oo]), (tag, f)
            if genetic:
                og, onb = o.genomes(row)
                assert np.array_equal(got["genome"], og[oo]) and np.array_equal(got["nbabies"], onb[oo]), tag

    sizes = [g.num_agents()]
    for k in range(12):
        if k == 6:  # re-split: the shared directory is rank 0's temporary directory
            box = [tempfile.mkdtemp(prefix="qhg_rebalance_")] if rank == 0 else [None]
            dist.broadcast_object_list(box, src=0)
            g, begin = sharding.rebalance(g, make, rank, world, box[0])
            sizes.append(g.num_agents())
            compare("after the re-split", begin[rank], begin[rank + 1])
        g.step(float(k))
        if rank == 0:
            o.step(float(k))
        compare(k, begin[rank], begin[rank + 1])
    per = [None] * world
    dist.gather_object(sizes, per if rank == 0 else None, dst=0)
    if rank == 0:
        before, after = [p[0] for p in per], [p[1] for p in per]
        assert max(after) - min(after) < max(before) - min(before), (before, after)
        how = "peer-memory" if os.environ.get("QHG_P2P", "1") != "0" else "nccl"
        print(f"mgpu_check ok [rebalance{', genetic' if genetic else ''}]: {world} ranks, agents per rank {before} -> {after} after sharding.rebalance at step 6, "
              f"12 steps, {o.num_agents()} agents ({how} exchange), bit-exact vs the unsharded oracle before and after")


def main_bigcell(rank, world):
    """A few cells far beyond the fast path's limits (2500 agents, hundreds of births and ranked fertile females per cell) in ONE
    rank's range: that rank's pass 1 raises the flag, it travels to every rank through the first cross-GPU barrier, all of them
    leave the step undone and redo it with the recovery kernels -- step by step and inside a run of queued steps."""
    nbr, xyz = make_ico_grid(15)
    alt = synthetic_altitude(xyz, seed=3)
    base = synthetic_population(40000, alt, seed=5, fertile=True)
    land = np.flatnonzero(alt > 0)
    hot = land[:3]                                     # the lowest land cells: rank 0's
    extra = synthetic_population(7500, alt, seed=9, fertile=True, cells=hot)
    pop = {k: np.concatenate([base[k], extra[k]]) for k in base}
    order = np.argsort(pop["cell"], kind="stable")
    pop = {k: v[order] for k, v in pop.items()}
    pop["id"] = np.arange(len(pop["id"]), dtype=np.int64)
    par, st = tut_environ_alt(4000.0), seed_state(31)
    begin = sharding.partition_cells(np.bincount(pop["cell"], minlength=len(nbr)), world)
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, device=int(os.environ.get("LOCAL_RANK", 0)), capacity_hint=200000)
    sharding.connect(g, begin, rank, world)
    g.add_agents(pop)
    g.pre_loop()
    o = None
    if rank == 0:
        from oracle import port
        o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st)
        o.add_agents(pop)
        o.start()

    def compare(tag):
        mine = g.agents()
        parts = [None] * world
        dist.gather_object({f: mine[f] for f in FIELDS}, parts if rank == 0 else None, dst=0)
        if rank == 0:
            allg = {f: np.concatenate([p[f] for p in parts]) for f in FIELDS}
            oa = o.agents()
            og, oo = np.argsort(allg["id"]), np.argsort(oa["id"])
            assert len(allg["id"]) == o.num_agents(), (tag, len(allg["id"]), o.num_agents())
            for f in FIELDS:
                assert np.array_equal(allg[f][og], oa[f][oo]), (tag, f)

    for k in range(3):
        g.step(float(k))
        if rank == 0:
            o.step(float(k))
        compare(k)
    g.run(3.0, 3)
    if rank == 0:
        for k in range(3, 6):
            o.step(float(k))
    compare("queued")
    rec = [None] * world
    dist.gather_object(g.path_counts()[2], rec if rank == 0 else None, dst=0)
    if rank == 0:
        assert min(rec) >= 4 and len(set(rec)) == 1, rec   # every rank redid the same steps
        how = "peer-memory" if os.environ.get("QHG_P2P", "1") != "0" else "nccl"
        biggest = int(o.counts().max())
        print(f"mgpu_check ok [bigcell]: {world} ranks, 6 steps (3 queued by qhgb_run), largest cell {biggest} agents, {rec[0]} steps redone with the "
              f"recovery kernels on every rank, {o.num_agents()} agents ({how} exchange), bit-exact vs the unsharded oracle")


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if len(sys.argv) > 1 and sys.argv[1] in ("genetic", "rebalance", "rebalance-genetic", "bigcell"):
        if sys.argv[1] == "bigcell":
            main_bigcell(rank, world)
        elif sys.argv[1] == "genetic":
            main_genetic(rank, world)
        else:
            main_rebalance(rank, world, sys.argv[1].endswith("genetic"))
        dist.barrier()
        dist.destroy_process_group()
        return
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
