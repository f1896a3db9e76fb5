"""Throughput of the other BASELINE.json configurations on ONE B200 (device-resident steps, CUDA events):
  C2  tutorial action set, 1e7 agents on the subdivision-256 grid
  C3  genetic population (OoANavGenPop without Navigate: OldAgeDeath, VerhulstVarK, NPPCapacity, MultiEvaluator, Genetics
      with 4096 one-bit sites, free recombination, mutation rate 1e-5), 1e7 agents
  C5' dynamic environment on one GPU: the genetic population plus Navigate (2000 ports x 4 destinations), a climate + sea
      level event every 10 steps, 1e7 agents
Parity for all of them is in tests/test_parity_gpu.py; this script only times them.
    python profiles/bench_configs.py [agents] > gpurun_out/bench_configs_r01.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np

import bench
from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_climate, synthetic_population
from qhg4_b200.params import ooa_nav_gen, seed_state
from qhg4_b200.population import GpuPopulation


def timed_steps(g, steps, warmup, t0=0.0, every=None, event=None):
    t = t0
    for _ in range(warmup):
        g.step(t); t += 1.0
    g.synchronize()
    agent_steps = 0
    g.event_record(0)
    for k in range(steps):
        agent_steps += g.num_agents()
        g.step(t); t += 1.0
        if every and (k + 1) % every == 0:
            event(t)
    g.event_record(1)
    g.synchronize()
    ms = g.event_elapsed_ms(0, 1)
    return {"agent_steps_per_s": agent_steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps, "agents_end": g.num_agents(),
            "fast_path_steps": None}


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    out = {}
    # ---- C2 ----
    nbr, alt, pop, par, K = bench.build_world(255, n)
    g = GpuPopulation.from_params(par, nbr, alt, capacity_hint=int(n * 1.6))
    g.add_agents(pop); g.pre_loop()
    out["C2"] = dict(timed_steps(g, 20, 3), workload=f"tutorial action set, {n} agents, 655362 cells, K={K}")
    g.close()
    # ---- C3 ----
    _, xyz = make_ico_grid(255)
    env = synthetic_climate(xyz, alt, seed=2)
    G = 4096
    row = 2 * (G // 64)
    par3 = ooa_nav_gen(G, -1, 1e-5)
    rng = np.random.default_rng(1)
    gen0 = rng.integers(0, 2 ** 63, size=(n, row), dtype=np.int64).astype(np.uint64)
    g = GpuPopulation.from_params(par3, nbr, alt, state16=seed_state(3), env=env)
    g.add_agents(pop); g.set_genomes(gen0); g.pre_loop()
    out["C3"] = dict(timed_steps(g, 10, 3), workload=f"OoANavGenPop without Navigate, {n} agents, genome 4096 one-bit sites "
                     f"({row * 8} B per agent), free recombination, mutation rate 1e-5")
    g.close()
    # ---- C5' ----
    land = np.flatnonzero(alt > 0)
    occupied = np.unique(pop["cell"])
    ports = rng.choice(occupied, 2000, replace=False).astype(np.int32)
    ptr = np.arange(0, 4 * 2000 + 1, 4, dtype=np.int32)
    dests = rng.choice(land, 4 * 2000).astype(np.int32)
    dist = rng.uniform(100, 700, 4 * 2000)
    par5 = ooa_nav_gen(G, -1, 1e-5)
    par5.modules["Navigate"] = {"Navigate_decay": "-0.001", "Navigate_dist0": "150.0", "Navigate_prob0": "0.1",
                                "Navigate_min_dens": "0.0", "Navigate_bridge_prob": "0.3"}
    par5.prios["Navigate"] = 10
    g = GpuPopulation.from_params(par5, nbr, alt, state16=seed_state(5), env=env)
    g.set_navigation(ports, ptr, dests, dist, np.zeros((0, 2), np.int32))
    g.add_agents(pop); g.set_genomes(gen0); g.pre_loop()
    state = {"alt": alt.copy(), "k": 0}

    def event(t):  # climate, vegetation and sea-level arrays replaced, then the events the reference's Simulator delivers
        state["k"] += 1
        state["alt"] = state["alt"] - 5.0
        g.set_env("Altitude", state["alt"])
        g.set_env("AnnualMeanTemp", env["AnnualMeanTemp"] - 0.5 * state["k"])
        g.set_env("BaseNPP", env["BaseNPP"] * (1.0 - 0.02 * state["k"]))
        for ev in (2, 3, 4, 5):
            g.update_event(ev, t)
        g.flush_events(t)

    out["C5_single_gpu"] = dict(timed_steps(g, 20, 3, every=10, event=event),
                                workload=f"OoANavGenPop with Navigate (2000 ports x 4 destinations), GEO+CLIMATE+VEG+NAV event every 10 steps "
                                         f"(arrays re-uploaded from the host inside the timed region), {n} agents")
    g.close()
    # ---- C5'' : the same population, the environment interpolated on the device (AutoInterpolator::interpolate): the
    # difference arrays are resident, an environment step is three small kernels + the events, no host array traffic
    g = GpuPopulation.from_params(par5, nbr, alt, state16=seed_state(5), env=env)
    g.set_navigation(ports, ptr, dests, dist, np.zeros((0, 2), np.int32))
    g.add_agents(pop); g.set_genomes(gen0); g.pre_loop()
    g.set_env_delta("Altitude", np.full(len(alt), -5.0))
    g.set_env_delta("AnnualMeanTemp", np.full(len(alt), -0.5))
    g.set_env_delta("BaseNPP", -0.02 * env["BaseNPP"])

    def event_dev(t):
        g.interpolate_env(1)
        for ev in (2, 3, 4, 5):
            g.update_event(ev, t)
        g.flush_events(t)

    out["C5_single_gpu_interpolated"] = dict(timed_steps(g, 20, 3, every=10, event=event_dev),
                                             workload=f"as C5_single_gpu, the environment changed by qhgb_interpolate_env on the device "
                                                      f"(resident difference arrays) instead of re-uploaded arrays, {n} agents")
    g.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
