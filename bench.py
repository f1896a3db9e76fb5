#!/usr/bin/env python
"""bench.py -- agent-steps/s of the per-step agent update on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--agents A] [--subdiv S]

A "step" is one `PopLooper::doStep` of the tutorial action set (GetOld, ATanDeath, WeightedMove on
SingleEvaluator(altitude), Fertility, RandomPair, Verhulst) over the whole synthetic population.
Workload at every N: C4 of SURVEY.md §8 -- 1e8 agents on the subdivision-256 icosahedral grid
(655,362 cells), K scaled so that N/K ~ 0.7 on land.  Rank 0 prints ONE JSON line.

* `value`    agent-steps/s with the population resident in HBM, timed with CUDA events on the population's stream.
* `e2e`      the same loop through the C ABI the way the reference's host loop uses a population: per step the
             host passes the step's inputs (time, action parameters) and reads back the step's results
             (totals + the per-cell count array that PopBase::getNumAgents exposes) into host memory.
* `roofline` whole-step algorithmic bytes (66 B per live agent + 80 B per cell) over the summed
             device time of the step's kernels, against the measured HBM copy bandwidth.
* `cpu_baseline` the reference's own OpenMP code (oracle/_ref) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
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


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_per_step(agents):
    """DRAM bytes per step of the two fast-path kernels (everything else is per-cell and tiny) from the committed
    ncu --set full capture, scaled to this run's agents."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r01.json")) as f:
            d = json.load(f)
        per_agent = (d["k_cell_decide"]["dram_bytes_per_launch"] + d["k_cell_scatter"]["dram_bytes_per_launch"]) / d["agents"]
        return per_agent * agents
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
    from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_population
    from qhg4_b200.params import tut_environ_alt
    nbr, xyz = make_ico_grid(subdiv)
    alt = synthetic_altitude(xyz, seed=1)
    n_land = int((alt > 0).sum())
    K = max(1.0, round(n_agents / n_land / 0.7))  # N/K ~ 0.7 on land cells
    pop = synthetic_population(n_agents, alt, seed=seed, fertile=True)  # pairing and births active from the first step
    return nbr, alt, pop, tut_environ_alt(K), K


def workload_name(args):
    if args.subdiv == 255:
        return (f"C4: tut_EnvironAltPop action set (GetOld, ATanDeath, WeightedMove+SingleEvaluator[Alt], Fertility, "
                f"RandomPair, Verhulst), {args.agents} agents on the subdivision-256 icosahedral grid")
    return f"tut_EnvironAltPop action set, {args.agents} agents on eq:{args.subdiv}"


def run_reference(args):
    """The reference's own OpenMP step loop (oracle/_ref) on the host cores, bounded sample of the workload."""
    from oracle import refsim
    if not refsim.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libqhgref.so was not built (needs /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    threads = int(os.environ.get("QHG_REF_THREADS", cores))
    dens = args.agents / (0.7 * (10 * (args.subdiv + 1) ** 2 + 2))
    sub = args.ref_subdiv
    ncell = 10 * (sub + 1) ** 2 + 2
    n = int(dens * 0.7 * ncell)
    nbr, alt, pop, par, K = build_world(sub, n)
    s = refsim.RefSim(par, nbr, alt, threads=threads)
    s.add_agents(pop)
    s.start()
    s.run(0.0, args.ref_warmup)
    sec, asteps = s.run(float(args.ref_warmup), args.ref_steps)
    val = asteps / sec
    sample = (f"same action set and density ({dens:.0f} agents per land cell, K={K:.0f}) on an eq:{sub} grid ({ncell} cells), "
              f"{n} agents, {args.ref_warmup} warm-up + {args.ref_steps} timed steps, {threads} OpenMP threads, stdout to /dev/null")
    line = {"metric": "agent-steps/sec", "value": val, "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": args.ref_steps,
            "warmup": args.ref_warmup, "ms_per_step": 1e3 * sec / args.ref_steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args), "sample": f"bounded sample of it: {n} agents on eq:{sub} ({ncell} cells), same density and K"},
            "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": threads, "kind": "reference", "sample": sample},
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
        import torch.distributed as dist  # host-side plumbing only (128-byte NCCL id, scalar reductions); the data
        dist.init_process_group("gloo", rank=rank, world_size=world)  # path is NCCL inside the C library
    t_setup = time.time()
    sampler = ClockSampler(device)
    sampler.start()
    nbr, alt, pop, par, K = build_world(args.subdiv, args.agents)
    ncell = len(nbr)
    begin = sharding.partition_cells(np.bincount(pop["cell"], minlength=ncell), world)
    lo, hi = np.searchsorted(pop["cell"], [begin[rank], begin[rank + 1]])  # the population is generated binned by cell
    pop = {k: v[lo:hi] for k, v in pop.items()}
    g = GpuPopulation.from_params(par, nbr, alt, device=device, capacity_hint=int((hi - lo) * 1.6) + 4096)
    if world > 1:
        sharding.connect(g, begin, rank, world)
    t0 = time.time()
    g.add_agents(pop)
    g.pre_loop()
    g.synchronize()
    upload_s = time.time() - t0
    del pop
    sampler.begin()  # sampled from the warm-up on: the timed region alone can be shorter than one sample
    t = 0.0
    for _ in range(args.warmup):
        g.step(t); t += 1.0
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
    g.run(t, args.steps); t += float(args.steps)
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
    prof_agent_steps = 0
    barrier()
    for _ in range(args.steps):
        prof_agent_steps += g.num_agents()
        g.step(t); t += 1.0
    barrier()
    ktimes = g.kernel_times()
    g.reset_kernel_times(False)
    prof_agent_steps = allsum(prof_agent_steps)
    pipeline_ms = ktimes.pop("pipeline_total", (0.0, 0))[0]  # device time of the steps, gaps between launches included

    # ---- timed region 2: end to end through the C ABI with host buffers --------------------------------------
    counts = g.host_array(ncell, np.uint64)  # page-locked: the per-step result is one DMA into host memory
    e2e_steps = 0
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_steps += g.num_agents()
        g.initialize_step(t)
        for lvl in sorted(set(g.prios.values())):
            g.do_actions(lvl, t)
        g.finalize_step()
        st = g.step_stats()          # totals of the step (D2H inside finalize_step)
        if world > 1:                # a shard reads back the counts of its own cell range
            g.counts_range(begin[rank], begin[rank + 1], counts)
        else:
            g.counts(counts)         # per-cell counts into host memory, as PopBase::getNumAgentsArray exposes them
        t += 1.0
    barrier()
    e2e_sec = allmax(time.perf_counter() - w0)
    e2e_val = allsum(e2e_steps) / e2e_sec
    t1 = time.time()
    final = g.agents()
    download_s = time.time() - t1
    n_final = int(allsum(len(final["id"])))
    del final

    # ---- roofline: whole-step algorithmic bytes over the summed kernel time ------------------------------------
    peak, peak_src = measured_peak_gbs()
    kern_ms = allmax(sum(v[0] for v in ktimes.values()))  # slowest rank
    peak *= world
    alg_bytes = ALG_BYTES_PER_AGENT * prof_agent_steps + ALG_BYTES_PER_CELL * ncell * args.steps
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    top = max(ktimes.items(), key=lambda kv: kv[1][0])[0] if ktimes else None
    # the two hot kernels against their own algorithmic bytes: pass 1 reads 17 B and writes 1 B per agent,
    # pass 2 reads 18 B per agent and writes 17 B per agent that is alive afterwards (~ the same number)
    per_kernel = {}
    for kname, bpa in (("k_cell_decide", 18.0), ("k_cell_scatter", 35.0)):
        if kname in ktimes and ktimes[kname][0] > 0:
            kms = allmax(ktimes[kname][0])
            gbs = bpa * prof_agent_steps / (kms * 1e-3) / 1e9
            per_kernel[kname] = {"alg_bytes_per_agent": bpa, "ms_per_step": round(kms / args.steps, 4), "achieved": round(gbs, 1),
                                 "frac": round(gbs / peak, 4)}
    roof = {"bound": "hbm", "kernel": "whole step (all kernels of one doStep, summed device time)", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_per_step(agent_steps / args.steps),
            "peak_source": peak_src,
            "alg_bytes_per_step": alg_bytes / args.steps, "dominant_kernel": top,
            "pipeline_ms_per_step": round(pipeline_ms / args.steps, 4), "per_kernel": per_kernel,
            "kernels_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])}}

    line = {"metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args),
                       "cells": ncell, "agents_start": args.agents, "agents_end": n_final, "verhulst_K": K,
                       "l2": "agent state per step (>2 GB at 1e8 agents) exceeds the 126 MB L2; no explicit flush",
                       "upload_s": round(upload_s, 2), "download_s": round(download_s, 2), "setup_s": round(t0 - t_setup, 2),
                       "parallelism": (f"cell-range shards x{world}, migration over " + ("peer memory (NVLink stores from the scatter kernel)"
                                       if os.environ.get("QHG_P2P", "1") != "0" else "NCCL send/recv")) if world > 1 else "single GPU",
                       "migrations_per_step": migrated / args.steps},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_val, "unit": "agent-steps/s", "h2d_bytes_per_step": 320, "d2h_bytes_per_step": 8 * int(begin[rank + 1] - begin[rank]) + 48,
                    "what": "initializeStep + doActions per level + finalizeStep through the C ABI, then totals and the per-cell count "
                            "array (ulong per cell, as PopBase::getNumAgentsArray; sharded: every rank its own cell range) copied into "
                            "page-locked host memory every step"},
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
    ap.add_argument("--agents", type=int, default=100_000_000)
    ap.add_argument("--subdiv", type=int, default=255)
    ap.add_argument("--ref-subdiv", type=int, default=80)
    ap.add_argument("--ref-steps", type=int, default=3)
    ap.add_argument("--ref-warmup", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        if rank == 0:
            args.ref_steps, args.ref_warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
            line = run_reference(args)
            if line:
                print(json.dumps(line))
        return

    line = run_ours(args, rank, world)
    if rank == 0:
        if not args.no_cpu_baseline:
            ref = run_reference(args)
            if ref:
                line["cpu_baseline"] = ref["cpu_baseline"]
        print(json.dumps(line))


if __name__ == "__main__":
    main()
