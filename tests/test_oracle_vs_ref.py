"""Pins the CPU restatement against the reference ITSELF (oracle/_ref, built from /root/reference/QHG4 by
oracle/Makefile).  Skipped where the reference library was not built."""
import numpy as np
import pytest

from oracle import port, refsim
from qhg4_b200.icogrid import make_ico_grid, make_torus_grid, synthetic_altitude, synthetic_population
from qhg4_b200.params import seed_state, tut_environ_alt

pytestmark = pytest.mark.skipif(not refsim.available(), reason="oracle/_ref/libqhgref.so not built")

FIELDS = ("cell", "id", "birth", "gender", "age", "last_birth", "life", "slot")


@pytest.mark.parametrize("seed,K,grid", [(0, 20.0, "ico"), (7, 8.0, "ico"), (3, 40.0, "torus")])
def test_well_mode_equals_reference_one_thread(seed, K, grid):
    if grid == "ico":
        nbr, xyz = make_ico_grid(7)
        alt = synthetic_altitude(xyz, seed=seed + 1)
    else:
        nbr = make_torus_grid(20, 20)
        alt = np.full(len(nbr), 700.0)
    pop = synthetic_population(12000, alt, seed=seed + 2, fertile=bool(seed % 2))
    st = seed_state(seed)
    par = tut_environ_alt(K)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    for k in range(20):
        r.step(float(k)); o.step(float(k))
        assert r.num_agents() == o.num_agents(), k
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        assert np.array_equal(r.counts(), o.counts()), k
    rb, rd = r.bd(); ob, od = o.bd()
    assert np.array_equal(rb, ob) and np.array_equal(rd, od)
    assert np.array_equal(r.weights(), o.weights())
    r.close()


def test_geo_event_equals_reference():
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=5)
    pop = synthetic_population(10000, alt, seed=6)
    par = tut_environ_alt(20.0)
    r = refsim.RefSim(par, nbr, alt, threads=1)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    for k in range(4):
        r.step(float(k)); o.step(float(k))
    alt2 = alt - 300.0
    ice = (xyz[:, 2] > 0.7).astype(np.uint8)
    r.geo_event(alt2, ice, 4.0)
    o.set_env("Altitude", alt2); o.set_env("Ice", ice.astype(float)); o.update_event(2, 4.0)
    for k in range(4, 9):
        r.step(float(k)); o.step(float(k))
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
    assert np.array_equal(r.weights(), o.weights())
    r.close()


def test_counter_mode_statistically_equivalent_to_reference():
    """The counter-mode law (other random streams, key-rank pairing, rank-based ids) against the reference with
    several threads: total population, age and per-cell occupancy distributions, two-sample KS at p > 0.01."""
    from scipy import stats
    nbr = make_torus_grid(16, 16)
    alt = np.full(len(nbr), 900.0)
    par = tut_environ_alt(20.0)
    tot_o, tot_r, age_o, age_r, occ_o, occ_r = [], [], [], [], [], []
    for s in range(16):
        pop = synthetic_population(3000, alt, seed=50 + s, fertile=True)
        st = seed_state(500 + s)
        o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st)
        r = refsim.RefSim(par, nbr, alt, threads=2, state16=st)
        o.add_agents(pop); r.add_agents(pop)
        o.start(); r.start()
        for k in range(30):
            o.step(float(k)); r.step(float(k))
        tot_o.append(o.num_agents()); tot_r.append(r.num_agents())
        age_o.append(o.agents()["age"]); age_r.append(r.agents()["age"])
        occ_o.append(o.counts()); occ_r.append(r.counts())
        r.close()
    assert stats.ks_2samp(tot_o, tot_r).pvalue > 0.01
    assert stats.ks_2samp(np.concatenate(age_o)[::5], np.concatenate(age_r)[::5]).pvalue > 0.01
    assert stats.ks_2samp(np.concatenate(occ_o), np.concatenate(occ_r)).pvalue > 0.01


def _cap_world(seed=2):
    from qhg4_b200.icogrid import synthetic_climate
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=seed)
    return nbr, xyz, alt, synthetic_climate(xyz, alt, seed=seed + 1)


def test_cap_alt_population_equals_reference():
    """tut_EnvironCapAltPop: NPPCapacity (Miami NPP, ramps, bonuses), MultiEvaluator[NPP+Alt] (double cumulation),
    VerhulstVarK -- WELL mode against the reference with one thread, incl. a climate event + flush."""
    from qhg4_b200.params import tut_environ_cap_alt
    nbr, xyz, alt, env = _cap_world()
    pop = synthetic_population(9000, alt, seed=4, fertile=True)
    par = tut_environ_cap_alt()
    r = refsim.RefSim(par, nbr, alt, threads=1, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, env=env)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    assert np.array_equal(r.capacities(), o.capacities()) and r.capacities().max() > 10
    for k in range(8):
        r.step(float(k)); o.step(float(k))
    assert np.array_equal(r.weights(), o.weights())
    # the climate gets colder and drier: capacities change on the flush, weights do not (the evaluator is not an observer)
    env2 = dict(env, AnnualMeanTemp=env["AnnualMeanTemp"] - 6.0, AnnualRainfall=env["AnnualRainfall"] * 0.6, BaseNPP=env["BaseNPP"] * 0.7)
    for name in ("AnnualMeanTemp", "AnnualRainfall", "BaseNPP"):
        r.set_env(name, env2[name]); o.set_env(name, env2[name])
    r.event(3, 8.0, flush=False); r.event(4, 8.0, flush=True)
    o.update_event(3, 8.0); o.update_event(4, 8.0); o.flush_events(8.0)
    assert np.array_equal(r.capacities(), o.capacities())
    for k in range(8, 16):
        r.step(float(k)); o.step(float(k))
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
    rb, rd = r.bd(); ob, od = o.bd()
    assert np.array_equal(rb, ob) and np.array_equal(rd, od) and np.array_equal(r.weights(), o.weights())
    r.close()
