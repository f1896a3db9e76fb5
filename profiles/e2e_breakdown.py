"""Where the end-to-end step goes: wall-clock per C-ABI call of bench.py's e2e loop (1 GPU, C4 workload)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from qhg4_b200.population import GpuPopulation

nbr, alt, pop, par, K = bench.build_world(255, int(os.environ.get("AGENTS", 100_000_000)))
g = GpuPopulation.from_params(par, nbr, alt, device=0, capacity_hint=int(len(pop["id"]) * 1.6))
g.add_agents(pop); g.pre_loop(); g.synchronize()
t = 0.0
for _ in range(3):
    g.step(t); t += 1.0
g.synchronize()
counts = g.host_array(len(nbr), np.uint64)
if os.environ.get("MIRROR", "1") == "1":
    g.mirror_counts(counts)
acc = {}
def timed(name, f, *a):
    t0 = time.perf_counter(); r = f(*a); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
N = 5
w0 = time.perf_counter()
for _ in range(N):
    timed("num_agents", g.num_agents)
    timed("initialize_step", g.initialize_step, t)
    for lvl in sorted(set(g.prios.values())):
        timed("do_actions", g.do_actions, lvl, t)
    timed("finalize_step", g.finalize_step)
    timed("step_stats", g.step_stats)
    if os.environ.get("MIRROR", "1") != "1":
        timed("counts", g.counts, counts)
    t += 1.0
g.synchronize()
tot = time.perf_counter() - w0
print("total ms/step", 1e3 * tot / N, {k: round(1e3 * v / N, 3) for k, v in acc.items()})
w0 = time.perf_counter()
for _ in range(N):
    g.step(t); t += 1.0
g.synchronize()
print("step() ms/step", 1e3 * (time.perf_counter() - w0) / N)
