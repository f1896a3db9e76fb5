"""A genetic population on the GENERIC device path (one thread per agent: k_actions, k_scatter, k_pair_*, k_make_offspring,
k_free_genomes) for an ncu capture of those kernels: QHG_B200_PATH=generic python profiles/prof_generic.py"""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np  # noqa: E402
import bench  # noqa: E402
from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_climate, synthetic_population  # noqa: E402
from qhg4_b200.params import ooa_nav_gen  # noqa: E402
from qhg4_b200.population import GpuPopulation  # noqa: E402

nbr, xyz = make_ico_grid(127)
alt = synthetic_altitude(xyz, seed=1)
env = synthetic_climate(xyz, alt, seed=2)
n = int(os.environ.get("AGENTS", 2_000_000))
pop = synthetic_population(n, alt, seed=1, fertile=True)
par = ooa_nav_gen(1024, -1, 1e-5)
g = GpuPopulation.from_params(par, nbr, alt, env=env)
g.add_agents(pop)
g.set_genomes(bench.synthetic_genomes(pop["id"], 2 * (1024 // 64)))
g.pre_loop()
g.modify_attributes("NPPCap_efficiency", bench.capacity_scale(g.capacities(), n))
g.update_event(4, 0.0); g.flush_events(0.0)
for k in range(4):
    g.step(float(k))
print("generic path:", g.path_counts(), g.num_agents(), "agents")
