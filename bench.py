#!/usr/bin/env python
"""bench.py -- agent-steps/s of the per-step agent update on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4|C5]

A "step" is one `PopLooper::doStep` over the whole synthetic population.  `--config` names the BASELINE.json
configuration; the default, and what the driver measures, is C4:

  C2  tut_EnvironAltPop action set (GetOld, ATanDeath, WeightedMove on SingleEvaluator[Alt], Fertility, RandomPair,
      Verhulst), 1e7 agents on the subdivision-256 icosahedral grid (655,362 cells)
  C3  OoANavGenPop without Navigate (OldAgeDeath, VerhulstVarK, NPPCapacity, MultiEvaluator[Alt+NPP], Genetics with 4096
      one-bit sites = 1 KiB of genome per agent, free recombination, mutation rate 1e-5), 1e7 agents
  C4  the C2 action set, 1e8 agents, sharded by contiguous cell ranges at N > 1 (migration over NVLink)
  C5  OoANavGenPop WITH Navigate (2000 ports x 4 destinations anywhere on the globe) and a climate + vegetation + sea-level
      event every 10 steps (environment interpolated on the device), 1e8 agents (1e7 on a single GPU: the genomes of
      1e8 agents do not fit one B200)

Rank 0 prints ONE JSON line.

* `value`    agent-steps/s with the population resident in HBM, CUDA events on the population's stream, max over ranks.
* `e2e`      the same loop through the C ABI the way the reference's host loop uses a population: per step the host passes
             the step's inputs and reads back the step's results (totals + the per-cell count array that
             PopBase::getNumAgents exposes) into page-locked host memory.
* `roofline` whole-step algorithmic bytes (66 B per live agent + 80 B per cell, + 3 genome rows per birth)
             over the summed device time of the step's kernels, against the measured HBM copy bandwidth.
* `checksum` 64-bit sum and xor of the agent ids and a hash of the per-cell counts after the last step: the same at every N.
* `cpu_baseline` / `--impl reference`: the reference's own OpenMP code (oracle/_ref) on the host cores, ON THE SAME
             CONFIGURATION (same grid, same agents, same parameters), fewer steps.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_AGENT = 66.0
ALG_BYTES_PER_CELL = 80.0
GENOME_SITES = 4096  # C3 / C5: one-bit sites per strand -> 2 x 64 words = 1 KiB per agent
EVENT_EVERY = 10     # C5: steps between two environment events

CONFIGS = {
    "C2": {"cls": "tut", "agents": 10_000_000},
    "C3": {"cls": "gen", "agents": 10_000_000, "nav": False, "events": False},
    "C4": {"cls": "tut", "agents": 100_000_000},
    "C5": {"cls": "gen", "agents": 100_000_000, "agents_1gpu": 10_000_000, "nav": True, "events": True},
}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(config):
    """DRAM bytes per launch of the two fast-path kernels from the committed `ncu --set full` capture of THIS configuration
    (profiles/traffic_r02.json, written by profiles/ncu_summary.py from the .ncu-rep).  Reported as captured -- with the
    agents of the captured launch beside it, never rescaled; null when no capture of this configuration is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r02.json")) as f:
            d = json.load(f)[config]
        return {"bytes_per_launch": {k: v["dram_bytes_per_launch"] for k, v in d["kernels"].items()},
                "agents_in_captured_launch": d["agents"], "source": d.get("source", "profiles/traffic_r02.json")}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []   # (arrival time, line)
        self.proc = None
        self.t_begin = 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def begin(self):
        """samples from here on count (nvidia-smi is started early: it needs a moment before its first line)"""
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter() + 0.03  # one more sampling interval: the line for the last steps
        time.sleep(0.05)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ta, ln in self.lines:
            if ta < self.t_begin or ta > t_end:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_world(subdiv, n_agents, seed=1):
    """the tutorial population (C2 / C4) on the icosahedral grid of the given subdivision"""
    from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_population
    from qhg4_b200.params import tut_environ_alt
    nbr, xyz = make_ico_grid(subdiv)
    alt = synthetic_altitude(xyz, seed=1)
    n_land = int((alt > 0).sum())
    K = max(1.0, round(n_agents / n_land / 0.7))  # N/K ~ 0.7 on land cells
    pop = synthetic_population(n_agents, alt, seed=seed, fertile=True)  # pairing and births active from the first step
    return nbr, alt, pop, tut_environ_alt(K), K


def config_agents(args, world):
    c = CONFIGS[args.config]
    if args.agents:
        return args.agents
    if world == 1 and "agents_1gpu" in c:
        return c["agents_1gpu"]
    return c["agents"]


def build_workload(args, world):
    """grid, environment, population and parameters of the configuration -- identical on every rank and for both arms"""
    from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_climate, synthetic_population
    from qhg4_b200.params import ooa_nav_gen
    c = CONFIGS[args.config]
    n = config_agents(args, world)
    if c["cls"] == "tut":
        nbr, alt, pop, par, K = build_world(args.subdiv, n)
        return {"nbr": nbr, "alt": alt, "pop": pop, "par": par, "K": K, "env": None, "nav": None, "row": 0, "agents": n}
    nbr, xyz = make_ico_grid(args.subdiv)
    alt = synthetic_altitude(xyz, seed=1)
    env = synthetic_climate(xyz, alt, seed=2)
    pop = synthetic_population(n, alt, seed=1, fertile=True)
    par = ooa_nav_gen(GENOME_SITES, -1, 1e-5)
    nav = None
    if c.get("nav"):
        rng = np.random.default_rng(11)
        land = np.flatnonzero(alt > 0)
        nports = min(2000, len(land) // 4)
        ports = rng.choice(land, nports, replace=False).astype(np.int32)
        nav = {"ports": ports, "ptr": np.arange(0, 4 * nports + 1, 4, dtype=np.int32),
               "dests": rng.choice(land, 4 * nports).astype(np.int32), "dist": rng.uniform(100, 700, 4 * nports)}
        par.modules["Navigate"] = {"Navigate_decay": "-0.001", "Navigate_dist0": "150.0", "Navigate_prob0": "0.1",
                                   "Navigate_min_dens": "0.0", "Navigate_bridge_prob": "0.3"}
        par.prios["Navigate"] = 10
    return {"nbr": nbr, "alt": alt, "pop": pop, "par": par, "K": None, "env": env, "nav": nav, "row": 2 * (GENOME_SITES // 64),
            "agents": n}


def synthetic_genomes(ids, row):
    """random genome rows for the founders, reproducible per agent id at any sharding: row = block[id mod 65,521]"""
    block = np.random.default_rng(5).integers(0, 2 ** 63, size=(65521, row), dtype=np.int64).astype(np.uint64)
    return block[np.asarray(ids) % 65521]


def workload_name(args, world):
    n = config_agents(args, world)
    grid = "the subdivision-256 icosahedral grid (655,362 cells)" if args.subdiv == 255 else f"eq:{args.subdiv}"
    if CONFIGS[args.config]["cls"] == "tut":
        return (f"{args.config}: tut_EnvironAltPop action set (GetOld, ATanDeath, WeightedMove+SingleEvaluator[Alt], Fertility, "
                f"RandomPair, Verhulst), {n} agents on {grid}")
    extra = (", Navigate (2000 ports x 4 destinations), GEO+CLIMATE+VEG+NAV event every %d steps (environment interpolated on the device)" % EVENT_EVERY
             if CONFIGS[args.config].get("nav") else " without Navigate")
    return (f"{args.config}: OoANavGenPop (OldAgeDeath, WeightedMove+MultiEvaluator[Alt+NPP], Fertility, RandomPair, VerhulstVarK, NPPCapacity, "
            f"Genetics {GENOME_SITES} one-bit sites, free recombination, mutation rate 1e-5){extra}, {n} agents on {grid}")


def capacity_scale(cap, n_agents):
    """NPPCap_efficiency such that N/K ~ 0.7 on the land cells at the start"""
    land = cap[cap > 0]
    return float((n_agents / max(1, len(land)) / 0.7) / land.mean()) if len(land) else 1.0


def run_reference(args, world, steps, warmup):
    """The reference's own OpenMP step loop (oracle/_ref) on the host cores, on the SAME configuration as the GPU arm."""
    from oracle import refsim
    if not refsim.available():
        return {"impl": "reference", "unavailable": "oracle/_ref/libqhgref.so was not built (needs /root/reference at build time)"}
    cores = os.cpu_count() or 1
    threads = int(os.environ.get("QHG_REF_THREADS", cores))
    w = build_workload(args, world)
    c = CONFIGS[args.config]
    kind = "reference"
    if c["cls"] == "gen" and not refsim.has_class("OoANavGenPop"):
        return {"impl": "reference", "unavailable": "OoANavGenPop is not part of this build of oracle/_ref"}
    s = refsim.RefSim(w["par"], w["nbr"], w["alt"], threads=threads, env=w["env"])
    if w["nav"]:
        s.set_navigation(w["nav"]["ports"], w["nav"]["ptr"], w["nav"]["dests"], w["nav"]["dist"], ())
    s.add_agents(w["pop"])
    if c["cls"] == "gen":
        s.set_genomes(synthetic_genomes(w["pop"]["id"], w["row"]))
    s.start()
    if c["cls"] == "gen":
        s.modify_attribute("NPPCap_efficiency", capacity_scale(s.capacities(), w["agents"]))
        s.event(4, 0.0, True)
    s.run(0.0, warmup)
    sec, asteps = s.run(float(warmup), steps)
    val = asteps / sec
    sample = (f"the whole configuration ({w['agents']} agents, {len(w['nbr'])} cells, same parameters and initial population as the GPU arm), "
              f"{warmup} warm-up + {steps} timed steps of PopLooper::doStep, {threads} OpenMP threads, stdout to /dev/null")
    line = {"metric": "agent-steps/sec", "value": val, "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args, world), "cells": len(w["nbr"]), "agents_start": w["agents"],
                       "agents_end": s.num_agents()},
            "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    s.close()
    return line


def run_ours(args, rank, world):
    from qhg4_b200 import sharding
    from qhg4_b200.population import GpuPopulation
    device = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist  # host-side plumbing only (128-byte NCCL id, IPC handles, scalar reductions)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    t_setup = time.time()
    sampler = ClockSampler(device)
    sampler.start()
    c = CONFIGS[args.config]
    genetic, events = c["cls"] == "gen", bool(c.get("events"))
    w = build_workload(args, world)
    nbr, alt, pop, par, K, env, nav, row = (w[k] for k in ("nbr", "alt", "pop", "par", "K", "env", "nav", "row"))
    n_start = w["agents"]
    ncell = len(nbr)
    load = np.bincount(pop["cell"], minlength=ncell).astype(np.float64)
    expected = load  # where the population is expected to settle (the agent buffers are sized for the larger of the two)
    if genetic and world > 1:
        # VerhulstVarK: the population settles where the carrying capacity is (NPP), away from the uniform start -- the ranges are
        # balanced for the mean of the initial and the expected load (capacities from a population without agents on this GPU)
        try:
            par_probe = par.copy()
            par_probe.prios.pop("Navigate", None)   # (its preLoop wants the Navigation group; the capacities do not)
            probe = GpuPopulation.from_params(par_probe, nbr, alt, device=device, env=env)
            probe.pre_loop()
            cap = np.maximum(probe.capacities(), 0.0)
            probe.close()
            if cap.sum() > 0:
                expected = cap * (load.sum() / cap.sum())
                load = 0.5 * load + 0.5 * expected
        except Exception as e:  # noqa: BLE001 -- the ranges then follow the initial load alone
            print(f"[bench] capacity probe failed ({e}); partition by the initial population", file=sys.stderr)
    begin = sharding.partition_cells(np.rint(load).astype(np.int64), world)
    if dist is not None:  # one table for everybody (rank 0's)
        box = [begin]
        dist.broadcast_object_list(box, src=0)
        begin = box[0]
    lo, hi = np.searchsorted(pop["cell"], [begin[rank], begin[rank + 1]])  # the population is generated binned by cell
    pop = {k: v[lo:hi] for k, v in pop.items()}
    n_own = max(int(hi - lo), int(expected[begin[rank]:begin[rank + 1]].sum()))
    g = GpuPopulation.from_params(par, nbr, alt, device=device, capacity_hint=int(n_own * 1.6) + 4096, env=env)
    if world > 1:
        sharding.connect(g, begin, rank, world)
    if nav:
        g.set_navigation(nav["ports"], nav["ptr"], nav["dests"], nav["dist"], ())
    t0 = time.time()
    g.add_agents(pop)
    if genetic:
        g.set_genomes(synthetic_genomes(pop["id"], row))
    g.pre_loop()
    if genetic:  # carrying capacities scaled to the population (N/K ~ 0.7 on land), through modifyAttributes + a VEG event
        g.modify_attributes("NPPCap_efficiency", capacity_scale(g.capacities(), n_start))
        g.update_event(4, 0.0); g.flush_events(0.0)
    if events:  # the environment drifts between two dated files: difference arrays resident on the device (AutoInterpolator)
        g.set_env_delta("Altitude", np.full(ncell, -0.5))
        g.set_env_delta("AnnualMeanTemp", np.full(ncell, -0.05))
        g.set_env_delta("BaseNPP", -0.002 * env["BaseNPP"])
    g.synchronize()
    upload_s = time.time() - t0
    del pop

    def env_event(t):  # app/Simulator.cpp:338-374: interpolate, deliver the interpolator's events, flush
        g.interpolate_env(EVENT_EVERY)
        for ev in (2, 3, 4, 5):
            g.update_event(ev, t)
        g.flush_events(t)

    state = {"t": 0.0, "k": 0}

    def advance(nsteps, queued):
        """nsteps steps from state['t'] on; C5 delivers an environment event every EVENT_EVERY steps"""
        left = nsteps
        while left > 0:
            chunk = min(left, EVENT_EVERY - state["k"] % EVENT_EVERY) if events else left
            if queued:
                g.run(state["t"], chunk)
            else:
                for i in range(chunk):
                    g.step(state["t"] + i)
            state["t"] += float(chunk); state["k"] += chunk; left -= chunk
            if events and state["k"] % EVENT_EVERY == 0:
                env_event(state["t"])

    sampler.begin()  # sampled from the warm-up on: the timed region alone can be shorter than one sample
    advance(args.warmup, False)
    g.synchronize()

    def allsum(x):
        if dist is None:
            return x
        import torch
        v = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return float(v)

    def allmax(x):
        if dist is None:
            return x
        import torch
        v = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v)

    def barrier():
        g.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- timed region 1: device-resident loop --------------------------------------------------------------
    # qhgb_run queues the K steps on the stream (no host round trip per step); the agent-step sum is kept on the device
    launches0 = g.launch_count()
    as0, sent0, _ = g.run_totals()
    barrier()
    g.event_record(0)
    advance(args.steps, True)
    g.event_record(1)
    barrier()
    as1, sent1, _ = g.run_totals()
    agent_steps, migrated = as1 - as0, sent1 - sent0
    ms = allmax(g.event_elapsed_ms(0, 1))  # device time of the K steps, max over ranks
    clocks = sampler.stop()
    launches = g.launch_count() - launches0
    agent_steps = allsum(agent_steps)
    migrated = allsum(migrated)
    value = agent_steps / (ms * 1e-3)

    # ---- per-kernel CUDA events over K steps of the same loop (kept out of region 1: the event calls cost host time) ----
    g.reset_kernel_times(True)
    prof_agent_steps, prof_births = 0, 0
    barrier()
    for _ in range(args.steps):
        prof_agent_steps += g.num_agents()
        advance(1, False)
        prof_births += g.step_stats().births
    barrier()
    ktimes = g.kernel_times()
    g.reset_kernel_times(False)
    prof_agent_steps = allsum(prof_agent_steps)
    prof_births = allsum(prof_births)
    pipeline_ms = ktimes.pop("pipeline_total", (0.0, 0))[0]  # device time of the steps, gaps between launches included

    # ---- timed region 2: end to end through the C ABI with host buffers --------------------------------------
    # the per-cell counts live in a page-locked host array that the library keeps current after every step, like the reference's
    # m_aiNumAgentsPerCell (PopBase::getNumAgentsArray hands out the pointer): the copy is queued behind the step's kernels and
    # arrives under the step's own synchronisation
    c_lo, c_hi = (begin[rank], begin[rank + 1]) if world > 1 else (0, ncell)   # a shard reads back the counts of its own cell range
    counts = g.host_array(c_hi - c_lo, np.uint64)
    g.mirror_counts(counts, c_lo, c_hi)
    e2e_steps = 0
    seen = 0
    levels = sorted(set(g.prios.values()))
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        t = state["t"]
        e2e_steps += g.num_agents()
        g.initialize_step(t)
        for lvl in levels:
            g.do_actions(lvl, t)
        g.finalize_step()            # D2H inside: the step's totals and the per-cell counts of the mirror
        st = g.step_stats()
        seen += int(counts[0]) + int(counts[-1])  # the host touches the result of every step
        state["t"] += 1.0; state["k"] += 1
        if events and state["k"] % EVENT_EVERY == 0:
            env_event(state["t"])
    barrier()
    e2e_sec = allmax(time.perf_counter() - w0)
    e2e_val = allsum(e2e_steps) / e2e_sec
    mirror_ok = bool(np.array_equal(counts, g.counts_range(c_lo, c_hi)))  # the mirrored array is the library's own read-back
    g.mirror_counts(None)

    # ---- checksum of the final state: the same whatever the number of GPUs -------------------------------------
    t1 = time.time()
    final = g.agents()
    download_s = time.time() - t1
    ids = final["id"].astype(np.uint64)
    n_final = int(allsum(len(ids)))
    id_sum, id_xor = int(ids.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(ids)) if len(ids) else 0
    cnt = g.counts().astype(np.int64)
    if dist is not None:
        import torch
        parts = [None] * world
        dist.all_gather_object(parts, (id_sum, id_xor))
        id_sum = sum(p[0] for p in parts) % (1 << 64)
        id_xor = 0
        for p in parts:
            id_xor ^= p[1]
        tc = torch.from_numpy(cnt)
        dist.all_reduce(tc)
        cnt = tc.numpy()
    checksum = {"agents": n_final, "id_sum_mod_2_64": f"{id_sum % (1 << 64):016x}", "id_xor": f"{id_xor:016x}",
                "cell_counts_sha1": hashlib.sha1(np.ascontiguousarray(cnt, np.int64).tobytes()).hexdigest()[:16],
                "steps_done": int(state["k"])}
    del final

    # ---- roofline: whole-step algorithmic bytes over the summed kernel time ------------------------------------
    peak, peak_src = measured_peak_gbs()
    wait_names = ("k_xbarrier_merge", "k_place_migrants")  # kernels that contain a cross-GPU barrier: their time is mostly waiting
    my_busy = sum(v[0] for k, v in ktimes.items() if k not in wait_names) / args.steps
    per_rank_busy = [my_busy]
    per_rank_kernels = None
    if dist is not None:
        per_rank_busy = [None] * world
        dist.all_gather_object(per_rank_busy, my_busy)
        per_rank_kernels = [None] * world  # where the ranks differ: the two passes of every rank, its agents and occupied cells
        mine = {k: round(v[0] / args.steps, 4) for k, v in ktimes.items() if k.startswith(("k_cell_decide", "k_cell_scatter"))}
        own = cnt_own = g.counts_range(int(begin[rank]), int(begin[rank + 1]))
        mine.update({"agents": int(own.sum()), "occupied_cells": int((cnt_own > 0).sum()), "cells": int(begin[rank + 1] - begin[rank])})
        dist.all_gather_object(per_rank_kernels, mine)
    kern_ms = allmax(sum(v[0] for v in ktimes.values()))  # slowest rank
    peak *= world
    genome_bytes = 3.0 * row * 8.0 * prof_births  # per birth: two parents read, one child written
    alg_bytes = ALG_BYTES_PER_AGENT * prof_agent_steps + ALG_BYTES_PER_CELL * ncell * args.steps + genome_bytes
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    top = max(ktimes.items(), key=lambda kv: kv[1][0])[0] if ktimes else None
    # the two hot kernels against their own algorithmic bytes: pass 1 reads 17 B and writes 1 B per agent,
    # pass 2 reads 18 B per agent and writes 17 B per agent that is alive afterwards (~ the same number)
    per_kernel = {}
    for kname, bpa in (("k_cell_decide", 18.0), ("k_cell_decide_genetic", 18.0), ("k_cell_decide_genetic_nav", 18.0),
                       ("k_cell_scatter", 35.0), ("k_cell_scatter_genetic", 51.0)):
        if kname in ktimes and ktimes[kname][0] > 0:
            kms = allmax(ktimes[kname][0])
            gbs = bpa * prof_agent_steps / (kms * 1e-3) / 1e9
            per_kernel[kname] = {"alg_bytes_per_agent": bpa, "ms_per_step": round(kms / args.steps, 4), "achieved": round(gbs, 1),
                                 "frac": round(gbs / peak, 4)}
    roof = {"bound": "hbm", "kernel": "whole step (all kernels of one doStep, summed device time)", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "traffic_ncu": ncu_traffic(args.config),
            "peak_source": peak_src,
            "alg_bytes_per_step": alg_bytes / args.steps, "dominant_kernel": top,
            "pipeline_ms_per_step": round(pipeline_ms / args.steps, 4), "per_kernel": per_kernel,
            "per_rank_busy_ms_per_step": [round(x, 4) for x in per_rank_busy], "per_rank": per_rank_kernels,
            "kernels_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])}}

    line = {"metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "name": args.config,
                       "cells": ncell, "agents_start": n_start, "agents_end": n_final, "verhulst_K": K,
                       "l2": "agent state per step (>2 GB at 1e8 agents) exceeds the 126 MB L2; no explicit flush",
                       "upload_s": round(upload_s, 2), "download_s": round(download_s, 2), "setup_s": round(t0 - t_setup, 2),
                       "parallelism": (f"cell-range shards x{world}, migration over " + ("peer memory (NVLink stores from the scatter kernel)"
                                       if os.environ.get("QHG_P2P", "1") != "0" else "NCCL send/recv")) if world > 1 else "single GPU",
                       "migrations_per_step": migrated / args.steps},
            "clocks": clocks, "gpu_launches": launches, "checksum": checksum,
            "e2e": {"value": e2e_val, "unit": "agent-steps/s", "h2d_bytes_per_step": 320, "d2h_bytes_per_step": 8 * int(begin[rank + 1] - begin[rank]) + 48,
                    "what": "initializeStep + doActions per level + finalizeStep through the C ABI; every step ends with the step's "
                            "totals and the per-cell count array (ulong per cell, as PopBase::getNumAgentsArray; sharded: every rank "
                            "its own cell range) in page-locked host memory (qhgb_mirror_num_agents_array: queued behind the step's "
                            "kernels, one synchronisation), and the host reads them",
                    "mirror_equals_readback": mirror_ok},
            "roofline": roof}
    g.close()
    if dist is not None:
        dist.barrier()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4", choices=sorted(CONFIGS))
    ap.add_argument("--agents", type=int, default=0, help="override the configuration's agent count (testing)")
    ap.add_argument("--subdiv", type=int, default=255)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        # the reference's CPU path on this arm's configuration; every timed step is a full doStep of the whole population
        # (about 2.7 s at 1e8 agents on 16 cores), so the step count is bounded to keep the run within a few minutes
        if rank == 0:
            line = run_reference(args, world, max(1, min(args.steps, 5)), max(1, min(args.warmup, 2)))
            print(json.dumps(line))
        return

    line = run_ours(args, rank, world)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            ref = run_reference(args, world, 2, 1)  # bounded: 1 warm-up + 2 timed steps of the same configuration
            if "cpu_baseline" in ref:
                line["cpu_baseline"] = ref["cpu_baseline"]
            else:
                line["cpu_baseline"] = {"value": None, "unit": "agent-steps/s", "cores": 0, "kind": "reference", "sample": ref.get("unavailable", "")}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
