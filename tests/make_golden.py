#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE itself (oracle/_ref/libqhgref.so, i.e. the unmodified
QHG4 sources compiled by oracle/Makefile).  Run in the build container (needs /root/reference at build
time); the committed fixtures are what travels.

    python tests/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refsim  # noqa: E402
from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_population  # noqa: E402
from qhg4_b200.params import DEFAULT_STATE, tut_environ_alt  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
POLY = "-0.1 0 0.1 0.01 1500 1.0 2000 1 3000 -9999"  # AltCapPref of tutorial_data/xmldat/tut_EnvironAlt.xml


def cases(names):
    """trajectories of the reference with one thread for the named cases of tests/golden_cases.py"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_cases as gc

    def ref_factory(par, nbr, alt_, st, env):
        return refsim.RefSim(par, nbr, alt_, threads=1, state16=st, env=env)

    for name in names:
        d = gc.build_inputs(name)
        out = gc.run_case(name, ref_factory, d)
        assert out["totals"][-1] > 0 and out["totals"][-1] != len(d["pop_id"]), name
        np.savez_compressed(os.path.join(OUT, f"case_{name}.npz"), **d, **{"ref_" + k: v for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:  # only the named cases (a new case does not have to rewrite the other fixtures)
        cases(sys.argv[1:])
        print("wrote", sys.argv[1:])
        return
    # --- utils/WELL512.cpp: default state (app/SimParams.cpp:82-87) and the thread-0 permutation (core/SPopulation.cpp:171-176)
    st = np.array(DEFAULT_STATE, np.uint32)
    st0 = np.array([st[(13 * j) % 16] for j in range(16)], np.uint32)
    np.savez(os.path.join(OUT, "well512.npz"), state=st, seq=refsim.well_sequence(st, 256),
             state_thread0=st0, seq_thread0=refsim.well_sequence(st0, 256))
    # --- utils/PolyLine.cpp through the evaluator's (float) cast
    x = np.concatenate([np.array([-500, -0.1, 0, 0.05, 0.1, 100, 750, 1500, 1750, 2000, 2000.5, 2500, 3000, 5000.0]),
                        np.linspace(-600, 3600, 211)])
    np.savez(os.path.join(OUT, "polyline.npz"), definition=POLY, x=x, y_float_cast=refsim.polyline_eval(POLY, x, True),
             y_double=refsim.polyline_eval(POLY, x, False))
    # --- deterministic sub-steps on a small grid: counts, Verhulst b/d, weights, ATanDeath p(age)
    nbr, xyz = make_ico_grid(3)
    alt = synthetic_altitude(xyz, seed=4)
    ice = (xyz[:, 2] > 0.9).astype(np.uint8)
    pop = synthetic_population(3000, alt, seed=11)
    par = tut_environ_alt(20.0)
    r = refsim.RefSim(par, nbr, alt, ice=ice, threads=1)
    r.add_agents(pop)
    r.start()
    counts0 = r.counts()
    ages = np.linspace(0, 90, 721).astype(np.float32)
    p_atan = r.atan_prob(ages)
    # --- a whole trajectory with ONE thread (bit-reproducible): per-step totals and the final agent table
    totals = []
    for k in range(12):
        r.step(float(k))
        totals.append(r.num_agents())
        if k == 0:
            b, d = r.bd()
            w = r.weights()
    fin = r.agents()
    np.savez_compressed(os.path.join(OUT, "tut_environ_alt_ico3.npz"), nbr=nbr, alt=alt, ice=ice, K=20.0,
                        **{"pop_" + k: v for k, v in pop.items()}, counts0=counts0, b_step0=b, d_step0=d, weights=w,
                        ages=ages, p_atan=p_atan, totals=np.array(totals), counts_final=r.counts(),
                        **{"fin_" + k: v for k, v in fin.items()})
    r.close()
    # --- every other population / action the oracle restates: trajectories of the reference with one thread
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_cases as gc

    def ref_factory(par, nbr, alt_, st, env):
        return refsim.RefSim(par, nbr, alt_, threads=1, state16=st, env=env)

    for name in gc.CASES:
        d = gc.build_inputs(name)
        out = gc.run_case(name, ref_factory, d)
        assert out["totals"][-1] > 0 and out["totals"][-1] != len(d["pop_id"]), name
        np.savez_compressed(os.path.join(OUT, f"case_{name}.npz"), **d, **{"ref_" + k: v for k, v in out.items()})
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
