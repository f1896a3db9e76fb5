#!/usr/bin/env python
"""A/B of the compiled shapes of k_cell_scatter on ONE population: the variants take turns step by step (the population drifts
slowly, interleaving keeps the comparison fair); prints the mean kernel time per variant and checks that the trajectory is the
one of the default variant (checksum of the final state against a second run with the default only).

    python profiles/ab_scatter.py C4 dense "384,6,4,1 192,6,16,2 ..." [steps per variant]
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    cfg, kind, variants = sys.argv[1], sys.argv[2], sys.argv[3].split()
    rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    var = "QHG_SCATTER_DENSE" if kind == "dense" else "QHG_SCATTER_SPARSE"
    import argparse
    args = argparse.Namespace(config=cfg, gpus=1, agents=int(os.environ.get("AB_AGENTS", 0)), subdiv=int(os.environ.get("AB_SUBDIV", 255)))
    from qhg4_b200.population import GpuPopulation
    w = bench.build_workload(args, 1)
    genetic = bench.CONFIGS[cfg]["cls"] == "gen"

    def fresh():
        g = GpuPopulation.from_params(w["par"], w["nbr"], w["alt"], capacity_hint=int(w["agents"] * 1.6) + 4096, env=w["env"])
        g.add_agents(w["pop"])
        if genetic:
            g.set_genomes(bench.synthetic_genomes(w["pop"]["id"], w["row"]))
        g.pre_loop()
        if genetic:
            g.modify_attributes("NPPCap_efficiency", bench.capacity_scale(g.capacities(), w["agents"]))
            g.update_event(4, 0.0); g.flush_events(0.0)
        return g

    def digest(g):
        a = g.agents()
        ids = np.sort(a["id"])
        return hashlib.sha1(ids.tobytes() + g.counts().tobytes()).hexdigest()[:16]

    g = fresh()
    t = 0.0
    for _ in range(3):
        g.step(t); t += 1.0
    acc = {v: [] for v in variants}
    other = {}
    for r in range(rounds):
        for v in variants:
            os.environ[var] = v
            g.reset_kernel_times(True)
            g.step(t); t += 1.0
            g.synchronize()
            kt = g.kernel_times()
            for k, x in kt.items():
                if k.startswith("k_cell_scatter"):
                    acc[v].append(x[0])
                elif k.startswith("k_cell_decide"):
                    other.setdefault(k, []).append(x[0])
    g.reset_kernel_times(False)
    d_mixed = digest(g)
    n_end = g.num_agents()
    g.close()
    os.environ.pop(var, None)
    g = fresh()
    g.run(0.0, int(t))
    d_default = digest(g)
    g.close()
    print(f"== {cfg} {kind}: {n_end} agents at the end, trajectory {'identical to' if d_mixed == d_default else 'DIFFERS from'} the default variant's ({d_mixed} / {d_default})")
    for v in variants:
        print(f"   {v:>14}: k_cell_scatter {np.mean(acc[v]):.4f} ms (min {np.min(acc[v]):.4f})")
    for k, x in other.items():
        print(f"   {k}: {np.mean(x):.4f} ms")


if __name__ == "__main__":
    main()
