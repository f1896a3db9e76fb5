import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from qhg4_b200.icogrid import make_ico_grid, synthetic_climate
from qhg4_b200.params import ooa_nav_gen, seed_state
from qhg4_b200.population import GpuPopulation
n = 10_000_000
nbr, alt, pop, par, K = bench.build_world(255, n)
_, xyz = make_ico_grid(255)
env = synthetic_climate(xyz, alt, seed=2)
G = 4096; row = 2 * (G // 64)
par3 = ooa_nav_gen(G, -1, 1e-5)
gen0 = np.random.default_rng(1).integers(0, 2 ** 63, size=(n, row), dtype=np.int64).astype(np.uint64)
g = GpuPopulation.from_params(par3, nbr, alt, state16=seed_state(3), env=env)
g.add_agents(pop); g.set_genomes(gen0); g.pre_loop()
t = 0.0
for _ in range(3):
    g.step(t); t += 1
g.reset_kernel_times(True)
for _ in range(5):
    g.step(t); t += 1
kt = g.kernel_times()
print({k: round(v[0] / 5, 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])}, g.num_agents(), g.step_stats().births)
