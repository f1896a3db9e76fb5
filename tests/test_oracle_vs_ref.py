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


@pytest.mark.parametrize("which", ["tut_SexualPop", "tut_MovePop", "tut_OldAgeDiePop"])
def test_small_tutorial_populations_equal_reference(which):
    """RandomMove (actions/RandomMove.cpp:65-100) and the other action orders of the tutorial ladder
    (tut_Sexual.xml: Fertility 2, RandomPair 3, Verhulst 6, RandomMove 7, GetOld 8, ATanDeath 10): the oracle's WELL
    mode reproduces the reference slot for slot."""
    from qhg4_b200.params import tut_move, tut_old_age_die, tut_sexual
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=4)
    pop = synthetic_population(9000, alt, seed=8, fertile=True)
    par = {"tut_SexualPop": tut_sexual(25.0, 0.2), "tut_MovePop": tut_move(0.3), "tut_OldAgeDiePop": tut_old_age_die()}[which]
    st = seed_state(13)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    fields = FIELDS if which == "tut_SexualPop" else tuple(f for f in FIELDS if f != "last_birth")
    for k in range(15):
        r.step(float(k)); o.step(float(k))
        assert r.num_agents() == o.num_agents(), k
        ra, oa = r.agents(), o.agents()
        for f in fields:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        assert np.array_equal(r.counts(), o.counts()), k
        if k == 7:
            # a GEO event that floods most of the land: these classes inherit SPopulation::updateEvent, which does nothing
            # (core/SPopulation.h:116) -- nobody drowns, unlike in tut_EnvironAltPop (populations/tut_EnvironAltPop.cpp:100-116)
            before = r.num_agents()
            alt2 = alt - 800.0
            r.geo_event(alt2, None, 8.0)
            o.set_env("Altitude", alt2); o.update_event(2, 8.0); o.flush_events(8.0)
            assert r.num_agents() == before == o.num_agents()
    assert r.num_agents() != len(pop["id"])  # something happened
    r.close()


def _lonlat(xyz):
    return np.degrees(np.arctan2(xyz[:, 1], xyz[:, 0])), np.degrees(np.arcsin(np.clip(xyz[:, 2], -1, 1)))


def test_partheno_and_static_populations_equal_reference():
    """The rest of the tutorial ladder.  tut_ParthenoPop (populations/tut_ParthenoPop.cpp): no pairing action, every female
    counts as mated (LinearBirth only tests m_iMateIndex >= 0, actions/LinearBirth.cpp:139-142; the founders' index is the
    zero of fresh memory, which oracle/ref_driver.cpp writes explicitly), newborns are forced female AFTER the gender draw
    decided their life state (:107-119).  tut_StaticPop: no actions, nothing may change."""
    from qhg4_b200.params import tut_partheno, tut_static
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=4)
    pop = synthetic_population(7000, alt, seed=8, fertile=True)
    pop["gender"][:] = 0
    st = seed_state(21)
    r = refsim.RefSim(tut_partheno(25.0, 0.2), nbr, alt, threads=1, state16=st)
    o = port.OraclePop(tut_partheno(25.0, 0.2), nbr, alt, mode=port.MODE_WELL, state16=st)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    births = 0
    for k in range(15):
        r.step(float(k)); o.step(float(k))
        assert r.num_agents() == o.num_agents(), k
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        assert np.array_equal(r.counts(), o.counts()), k
        births += o.step_stats()[0]
    assert births > 1000 and not oa["gender"].any() and set(np.unique(oa["life"])) >= {1, 5}
    r.close()
    r = refsim.RefSim(tut_static(), nbr, alt, threads=1, state16=st)
    o = port.OraclePop(tut_static(), nbr, alt, mode=port.MODE_WELL, state16=st)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    c0 = o.counts().copy()
    for k in range(3):
        r.step(float(k)); o.step(float(k))
    ra, oa = r.agents(), o.agents()
    for f in ("cell", "id", "birth", "gender", "life", "slot"):  # the bare Agent has no age / last birth
        assert np.array_equal(ra[f], oa[f]), f
    assert np.array_equal(r.counts(), c0) and np.array_equal(o.counts(), c0) and np.array_equal(oa["id"], pop["id"])
    r.close()


def test_confined_move_equals_reference():
    """ConfinedMove (actions/ConfinedMove.cpp:44-101) pinned against the reference's own action added to tut_EnvironAltPop
    (ConfProbePop in oracle/ref_driver.cpp): moves that would leave the disc around (x, y) are turned into moves to the
    cell they start from -- still counted -- so nobody ever leaves, slot for slot."""
    from qhg4_b200.params import tut_environ_alt_confined
    nbr, xyz = make_ico_grid(7)
    alt = np.minimum(np.abs(synthetic_altitude(xyz, seed=5)) + 50.0, 1400.0)  # all land: the disc's border is the only barrier
    lon, lat = _lonlat(xyz)
    env = {"Longitude": lon, "Latitude": lat}
    par = tut_environ_alt_confined(150.0, 20.0, 10.0, 5000.0)
    par.modules["WeightedMove"]["WeightedMove_prob"] = "0.4"
    st = seed_state(5)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st, env=env)
    # the population starts inside the disc (the action only filters moves, actions/ConfinedMove.cpp:86-101)
    xc = np.array([np.cos(np.radians(20)) * np.cos(np.radians(10)), np.sin(np.radians(20)) * np.cos(np.radians(10)), np.sin(np.radians(10))])
    inside = np.flatnonzero(6371.3 * np.arccos(np.clip(xyz @ xc, -1, 1)) < 4000.0)
    pop = synthetic_population(9000, alt, seed=6, fertile=True, cells=inside)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    moves = 0
    for k in range(12):
        r.step(float(k)); o.step(float(k))
        assert r.num_agents() == o.num_agents(), k
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        assert np.array_equal(r.counts(), o.counts()), k
        moves += o.step_stats()[2]
    occ = np.flatnonzero(o.counts())
    d = 6371.3 * np.arccos(np.clip(xyz[occ] @ xc, -1, 1))
    assert moves > 10000 and 4000.0 < d.max() < 5000.0  # they reached the border and stayed inside
    # the same run without the action spills over the border
    par2 = par.copy(); del par2.prios["ConfinedMove"]
    o2 = port.OraclePop(par2, nbr, alt, mode=port.MODE_WELL, state16=st, env=env)
    o2.add_agents(pop); o2.start()
    for k in range(12):
        o2.step(float(k))
    occ2 = np.flatnonzero(o2.counts())
    assert (6371.3 * np.arccos(np.clip(xyz[occ2] @ xc, -1, 1))).max() > 5000.0
    r.close()


@pytest.mark.parametrize("move_rand,sig_death", [(True, False), (False, True), (True, True)])
def test_weighted_move_rand_and_sig_death_equal_reference(move_rand, sig_death):
    """WeightedMoveRand (actions/WeightedMoveRand.cpp:43-100: the cumulated weights decide unless they are all zero, then a
    uniform pick with `wrandr`) and SigDeath (actions/SigDeath.cpp:49-90: p = scale / (1 + exp(-slope (age - max_age))), one
    draw per agent) pinned against the reference's own templates added to tut_EnvironAltPop (VarProbePop in
    oracle/ref_driver.cpp); the altitude field has plateaus above the poly-line's range, so all-zero weight rows occur."""
    from qhg4_b200.params import tut_environ_alt_variants
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=5)
    alt[(alt > 400) & (alt < 900)] = 2600.0      # a belt of cells whose own and neighbours' weights are zero
    pop = synthetic_population(9000, alt, seed=6, fertile=True, max_age=70.0)
    par = tut_environ_alt_variants(25.0, move_rand, sig_death)
    st = seed_state(12)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    moves = deaths = 0
    for k in range(12):
        r.step(float(k)); o.step(float(k))
        assert r.num_agents() == o.num_agents(), k
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        assert np.array_equal(r.counts(), o.counts()), k
        b, d, m = o.step_stats()
        moves += m; deaths += d
    w = o.weights()
    assert moves > 3000 and deaths > 500 and (w[np.unique(pop["cell"]), 6] == 0).any()
    r.close()


@pytest.mark.parametrize("cond_mode,perm_pair,move_stats", [(2, False, -1), (3, True, -1), (6, True, 2), (7, False, 0), (-1, True, 0),
                                                            (1, False, 1), (5, True, 2), (-1, False, 1)])
def test_cond_weighted_move_rand_perm_pair_and_move_stats_equal_reference(cond_mode, perm_pair, move_stats):
    """CondWeightedMove (actions/CondWeightedMove.cpp:41-86 with actions/SimpleCondition.cpp over the altitudes: greater, less,
    less-or-equal, different ...), RandPermPair (actions/RandPermPair.cpp:67-196: partial Fisher-Yates over the larger sex) and
    MoveStats (actions/MoveStats.cpp:107-285: hops / distance / time of the first, the minimal or the last arrival per cell)
    pinned against the reference's own templates added to tut_EnvironAltPop (ExtProbePop<m> in oracle/ref_driver.cpp)."""
    from qhg4_b200.params import tut_environ_alt_ext
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=5)
    ice = (xyz[:, 2] > 0.9).astype(np.float64)       # CondWeightedMove looks at the ice of the cell the agent is IN
    land = np.flatnonzero(alt > 0)
    alt[land[::3]] = 30.0                             # lowlands: what "less" (0.2 x new < 10 and < current) and "greater" need
    start = land[np.argsort(xyz[land, 0])[:60]]       # a small home range: most cells are reached during the run
    pop = synthetic_population(6000, alt, seed=6, fertile=True, cells=start)
    par = tut_environ_alt_ext(150.0, cond_mode, perm_pair, move_stats, move_prob=0.3)
    lon = np.degrees(np.arctan2(xyz[:, 1], xyz[:, 0])); lat = np.degrees(np.arcsin(xyz[:, 2]))
    env = {"Longitude": lon, "Latitude": lat}
    st = seed_state(21)
    r = refsim.RefSim(par, nbr, alt, ice=ice, threads=1, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, ice=ice, mode=port.MODE_WELL, state16=st, env=env)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    moves = births = 0
    for k in range(14):
        r.step(float(k)); o.step(float(k))
        assert r.num_agents() == o.num_agents(), k
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        assert np.array_equal(r.counts(), o.counts()), k
        if move_stats >= 0:
            for x, y, name in zip(r.move_stats(), o.move_stats(), ("hops", "dist", "time")):
                assert np.array_equal(x, y), (k, name)
        b, d, m = o.step_stats()
        moves += m; births += b
    assert moves > (100 if cond_mode in (2, 3) else 1500) and births > 300   # "greater" / "less" with the 0.2 factor allow few moves
    if move_stats >= 0:
        h, dist, tm = o.move_stats()
        assert (h > 0).sum() > 5 and (h < 0).any() and dist[h > 0].min() > 50.0
    r.close()


def test_navigate_equals_reference():
    """Navigate (actions/Navigate.cpp:94-250) pinned against the reference's own action: the reference's Navigate<T> is
    added to the reference's tut_EnvironAltPop (NavProbePop in oracle/ref_driver.cpp; the shipped populations that carry
    Navigate also need Genetics and QDF sequence I/O), the oracle runs the same action list.  Sea crossings with
    distance-dependent probabilities, the search bound `iNumDests = map key` (port cells with a small index), manual
    bridges, a GEO event that drowns a bridge end and rebuilds the tables on the flush -- slot for slot."""
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=5)
    rng = np.random.default_rng(3)
    land = np.flatnonzero(alt > 0)
    pop = synthetic_population(12000, alt, seed=6, fertile=True)
    occupied = np.unique(pop["cell"])
    ports = np.concatenate([occupied[:3], rng.choice(occupied[occupied > 8], 50, replace=False)]).astype(np.int32)  # incl. tiny indices
    ptr = np.arange(0, 4 * len(ports) + 1, 4, dtype=np.int32)
    dests = np.concatenate([rng.choice(land, 4, replace=False) for _ in ports]).astype(np.int32)
    dist = rng.uniform(100, 700, len(dests))
    bridges = rng.choice(occupied, (6, 2), replace=False).astype(np.int32)
    par = tut_environ_alt(20.0)
    par.class_name = "tut_EnvironAltNavPop"
    par.modules["Navigate"] = {"Navigate_decay": "-0.001", "Navigate_dist0": "150.0", "Navigate_prob0": "0.1",
                               "Navigate_min_dens": "0.0", "Navigate_bridge_prob": "0.3"}
    par.prios["Navigate"] = 8
    par.modules["OldAgeDeath"] = {"OAD_max_age": "60.0", "OAD_uncertainty": "0.1"}  # registered in the probe class, no <prio>: never runs
    st = seed_state(9)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st)
    for q in (r, o):
        q.set_navigation(ports, ptr, dests, dist, bridges)
        q.add_agents(pop)
    r.start(); o.start()
    moves_seen = 0
    for k in range(12):
        r.step(float(k)); o.step(float(k))
        assert r.num_agents() == o.num_agents(), k
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        assert np.array_equal(r.counts(), o.counts()), k
        moves_seen += o.step_stats()[2]
        if k == 5:  # sea level rises: the drowned die, bridges with an end under water close (rebuilt on the flush)
            alt2 = alt - 150.0
            r.geo_event(alt2, None, 6.0)
            o.set_env("Altitude", alt2); o.update_event(2, 6.0); o.flush_events(6.0)
    far = np.setdiff1d(dests, np.concatenate([occupied, nbr[occupied].ravel()]))
    assert moves_seen > 0 and (far.size == 0 or o.counts()[far].sum() > 0)  # somebody did cross
    r.close()


def test_old_age_death_equals_reference():
    """OldAgeDeath (actions/OldAgeDeath.cpp:48-67, `wrandr` draw) pinned the same way: the probe population runs it
    INSTEAD of ATanDeath (an action without a <prio> entry never runs, core/Prioritizer.cpp:19-30)."""
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=2)
    pop = synthetic_population(9000, alt, seed=3, fertile=True, max_age=75.0)
    par = tut_environ_alt(20.0)
    par.class_name = "tut_EnvironAltNavPop"
    del par.prios["ATanDeath"]
    par.modules["OldAgeDeath"] = {"OAD_max_age": "60.0", "OAD_uncertainty": "0.1"}
    par.prios["OldAgeDeath"] = 2
    par.modules["Navigate"] = {"Navigate_decay": "-0.001", "Navigate_dist0": "150.0", "Navigate_prob0": "0.1",
                               "Navigate_min_dens": "0.0", "Navigate_bridge_prob": "0.3"}  # registered, no <prio>
    st = seed_state(4)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st)
    for q in (r, o):
        q.add_agents(pop)
    r.start(); o.start()
    deaths = 0
    for k in range(10):
        r.step(float(k)); o.step(float(k))
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        deaths += o.step_stats()[1]
    assert deaths > 500 and ra["age"].max() < 67.0  # nobody outlives max_age + 1 + 0.1 * max_age
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


def test_genome_primitives_equal_reference():
    """genes/BitGeneUtils.cpp crossOver (incl. its unsorted break lists), freeReco, mutateNucs and the mutation-count
    table of utils/BinomialDist.cpp: the oracle's restatement against the reference functions, same WELL512 stream."""
    import ctypes as C
    R, O = refsim.lib(), port.lib()
    for L, pre in ((R, "qref"), (O, "qor")):
        getattr(L, pre + "_binomial_table").argtypes = [C.c_double, C.c_int, C.c_double, C.c_int, C.c_void_p]
        getattr(L, pre + "_binomial_get_n").argtypes = [C.c_double, C.c_int, C.c_double, C.c_double]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(0)
    for trial in range(120):
        G = int(rng.choice([64, 128, 200, 1000, 4096]))
        nb = (G + 63) // 64
        gin = rng.integers(0, 2 ** 63, size=2 * nb, dtype=np.int64).astype(np.uint64)
        st = seed_state(trial + 1)
        for nc in (1, 2, 3, 7, 20):
            a, b = np.zeros(2 * nb, np.uint64), np.zeros(2 * nb, np.uint64)
            R.qref_bit_crossover(p(st), p(gin), G, nc, p(a)); O.qor_bit_crossover(p(st), p(gin), G, nc, p(b))
            assert np.array_equal(a, b), (trial, nc)
        a, b = np.zeros(2 * nb, np.uint64), np.zeros(2 * nb, np.uint64)
        R.qref_bit_freereco(p(st), p(gin), nb, p(a)); O.qor_bit_freereco(p(st), p(gin), nb, p(b))
        assert np.array_equal(a, b)
        a, b = gin.copy(), gin.copy()
        R.qref_bit_mutate(p(st), p(a), 2 * G, 6); O.qor_bit_mutate(p(st), p(b), 2 * G, 6)
        assert np.array_equal(a, b) and not np.array_equal(a, gin)
    for pr, n in ((1e-5, 8192), (1e-3, 8192), (0.01, 256), (1e-4, 128)):
        ta, tb = np.zeros(128), np.zeros(128)
        na = R.qref_binomial_table(pr, n, 1e-6, 128, p(ta)); nb_ = O.qor_binomial_table(pr, n, 1e-6, 128, p(tb))
        assert na == nb_ and np.array_equal(ta, tb)
        for r in (0.0, 0.5, 0.93, 0.999, 0.9999999):
            assert R.qref_binomial_get_n(pr, n, 1e-6, r) == O.qor_binomial_get_n(pr, n, 1e-6, r)
    # BinomialDist::create(1e-5, 8192, 1e-6)
    t = np.zeros(16)
    assert O.qor_binomial_table(1e-5, 8192, 1e-6, 16, p(t)) == 5 and abs(t[0] - 0.9213453) < 1e-6
    assert [O.qor_binomial_get_n(1e-5, 8192, 1e-6, r) for r in (0.5, 0.93, 0.999, 0.9999999)] == [0, 1, 2, 4]


@pytest.mark.parametrize("bits", [1, 2])
@pytest.mark.parametrize("ncross,mut", [(-1, 1e-3), (3, 5e-3), (0, 0.0)])
def test_genetics_action_equals_reference(bits, ncross, mut):
    """The Genetics action itself (actions/Genetics.cpp:285-337: strand choice, crossover or free recombination of both
    parents, mutation count from the binomial table, mutation positions -- all drawn from the action's OWN generator) with
    1-bit (genes/BitGeneUtils.cpp) and 2-bit (genes/GeneUtils.cpp) nucleotides, pinned against the reference's own
    Genetics<T,U> added to the reference's tut_EnvironAltPop (GenProbePop<U> in oracle/ref_driver.cpp, called from
    makePopSpecificOffspring like populations/OoANavGenPop.cpp:231-245): every genome of every live agent, slot for slot."""
    from qhg4_b200.params import tut_environ_alt_genetic
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=5)
    G = 200
    row = 2 * ((G * bits + 63) // 64)
    par = tut_environ_alt_genetic(20.0, G, ncross, mut, bits)
    pop = synthetic_population(8000, alt, seed=6, fertile=True)
    gen0 = np.random.default_rng(3).integers(0, 2 ** 63, size=(8000, row), dtype=np.int64).astype(np.uint64)
    st = seed_state(9)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st)
    r.add_agents(pop); o.add_agents(pop)
    r.set_genomes(gen0); o.set_genomes(gen0)
    o.set_genetics_well(*r.genetics_well())  # built in the reference from aiSeeds[1] through MD5 digests (utils/WELLUtils.cpp:127-148)
    r.start(); o.start()
    for k in range(10):
        r.step(float(k)); o.step(float(k))
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        og, _ = o.genomes(row)
        assert np.array_equal(r.genomes(row), og), k
    born = oa["id"] >= 8000
    assert born.sum() > 1500
    if ncross == 0 and mut == 0.0:  # whole parental strands are handed down: every strand of a newborn is a founder strand
        nb = row // 2
        founders = {bytes(x) for x in gen0.reshape(-1, nb)}
        assert all(bytes(x) in founders for x in og[born].reshape(-1, nb))
    r.close()


def _genotype_samples(make_sim, nseeds, nsteps, G, bits, ncross, mut):
    """per-site allele frequencies, per-agent heterozygosity and per-strand switch counts after nsteps, over nseeds runs"""
    from qhg4_b200.params import tut_environ_alt_genetic
    nbr = make_torus_grid(12, 12)
    alt = np.full(len(nbr), 900.0)
    par = tut_environ_alt_genetic(20.0, G, ncross, mut, bits)
    nb = (G * bits + 63) // 64
    freq, het, sw, tot = [], [], [], []
    for s in range(nseeds):
        pop = synthetic_population(2500, alt, seed=70 + s, fertile=True)
        # founder strands are all-0 or all-1: every switch along a descendant's strand is a recombination break or a mutation
        hap = np.random.default_rng(900 + s).integers(0, 2, size=(2500, 2), dtype=np.uint64) * np.uint64(0xFFFFFFFFFFFFFFFF)
        gen0 = np.ascontiguousarray(np.repeat(hap, nb, axis=1))
        g = make_sim(par, nbr, alt, seed_state(700 + s), pop, gen0)
        for k in range(nsteps):
            g.step(float(k))
        rows = g.genomes(2 * nb)
        rows = rows[0] if isinstance(rows, tuple) else rows
        bitsarr = np.unpackbits(rows.view(np.uint8).reshape(len(rows), 2, nb * 8), axis=2, bitorder="little")[:, :, :G * bits]
        freq.append([bitsarr.mean()])  # one value per run: with two founder haplotypes the sites of a run move together
        het.append((bitsarr[:, 0, :] != bitsarr[:, 1, :]).sum(axis=1))
        sw.append((bitsarr[:, :, 1:] != bitsarr[:, :, :-1]).sum(axis=2).ravel())
        tot.append(len(rows))
        if hasattr(g, "close"):
            g.close()
    return np.concatenate(freq), np.concatenate(het), np.concatenate(sw), np.array(tot)


def _ref_sim(threads):
    def make(par, nbr, alt, st, pop, gen0):
        r = refsim.RefSim(par, nbr, alt, threads=threads, state16=st)
        r.add_agents(pop); r.set_genomes(gen0); r.start()
        return r
    return make


def _oracle_sim(par, nbr, alt, st, pop, gen0):
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st)
    o.add_agents(pop); o.set_genomes(gen0); o.start()
    return o


@pytest.mark.parametrize("bits,ncross", [(1, -1), (2, 2)])
def test_genotype_distributions_statistically_equivalent_to_reference(bits, ncross):
    """Genotype side of the north-star's statistical gate: the counter-mode law (other random streams for strand choice,
    recombination and mutation; key-rank pairing decides who the father is) against the reference's Genetics with two
    threads -- the allele frequency of every run after 25 steps of drift, per-agent heterozygosity, and the number of switches along
    a strand (founder strands are all-0 or all-1, so switches count recombination breaks and mutations), two-sample KS,
    p > 0.01.  The switch statistic does tell laws apart: another crossover count or mutation rate fails it."""
    from scipy import stats
    G = 96
    fo, ho, so, to = _genotype_samples(_oracle_sim, 12, 25, G, bits, ncross, 2e-3)
    fr, hr, sr, tr = _genotype_samples(_ref_sim(2), 12, 25, G, bits, ncross, 2e-3)
    ps = [stats.ks_2samp(fo, fr).pvalue, stats.ks_2samp(ho[::5], hr[::5]).pvalue, stats.ks_2samp(so[::11], sr[::11]).pvalue,
          stats.ks_2samp(to, tr).pvalue]
    assert min(ps) > 0.01, ps
    assert fo.std() > 0.002 and so.mean() > 1.0  # drift and recombination happened
    _, _, sx, _ = _genotype_samples(_oracle_sim, 12, 25, G, bits, ncross, 8e-3)  # four times the mutation rate: detected
    assert stats.ks_2samp(sx[::11], sr[::11]).pvalue < 0.01


def test_two_bit_genome_primitives_equal_reference():
    """genes/GeneUtils.cpp crossOver (breaks on nucleotide boundaries), freeReco (doubled mask bits), mutateNucs (XOR with
    01 / 10 / 11): the oracle's restatement against the reference functions on the same WELL512 stream."""
    import ctypes as C
    R, O = refsim.lib(), port.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(0)
    for trial in range(100):
        G = int(rng.choice([32, 64, 100, 500, 2048]))
        nb = (2 * G + 63) // 64
        gin = rng.integers(0, 2 ** 63, size=2 * nb, dtype=np.int64).astype(np.uint64)
        st = seed_state(trial + 1)
        for nc in (1, 2, 3, 7, 20):
            a, b = np.zeros(2 * nb, np.uint64), np.zeros(2 * nb, np.uint64)
            R.qref_gene2_crossover(p(st), p(gin), G, nc, p(a)); O.qor_gene2_crossover(p(st), p(gin), G, nc, p(b))
            assert np.array_equal(a, b), (trial, nc)
        a, b = np.zeros(2 * nb, np.uint64), np.zeros(2 * nb, np.uint64)
        R.qref_gene2_freereco(p(st), p(gin), nb, p(a)); O.qor_gene2_freereco(p(st), p(gin), nb, p(b))
        assert np.array_equal(a, b)
        a, b = gin.copy(), gin.copy()
        R.qref_gene2_mutate(p(st), p(a), 2 * G, 6); O.qor_gene2_mutate(p(st), p(b), 2 * G, 6)
        assert np.array_equal(a, b) and not np.array_equal(a, gin)


def test_genetic_population_counter_mode_properties():
    """OoANavGenPop in the oracle: order invariance with genomes, inheritance (every newborn strand is made of parental
    alleles when there is no mutation)."""
    from qhg4_b200.params import ooa_nav_gen
    nbr, xyz, alt, env = _cap_world()
    pop = synthetic_population(6000, alt, seed=4, fertile=True)
    G, row = 128, 4
    gen0 = np.random.default_rng(3).integers(0, 2 ** 63, size=(6000, row), dtype=np.int64).astype(np.uint64)
    res = []
    for perm in (np.arange(6000), np.random.default_rng(1).permutation(6000)):
        o = port.OraclePop(ooa_nav_gen(G, -1, 0.0), nbr, alt, mode=port.MODE_COUNTER, state16=seed_state(9), env=env)
        o.add_agents({k: v[perm] for k, v in pop.items()})
        o.set_genomes(gen0[perm])
        o.start()
        for k in range(8):
            o.step(float(k))
        a = o.agents(); g, nbab = o.genomes(row)
        s = np.argsort(a["id"])
        res.append((a["id"][s], a["cell"][s], g[s], nbab[s]))
    for x, y in zip(res[0], res[1]):
        assert np.array_equal(x, y)
    ids, _, g, nbab = res[0]
    assert nbab.sum() > 500 and (ids >= 6000).sum() > 500
    # with all founders carrying allele 0 at bit 0 of both strands, no descendant can carry allele 1 there
    gz = gen0.copy(); gz[:, 0] &= ~np.uint64(1); gz[:, 2] &= ~np.uint64(1)
    o = port.OraclePop(ooa_nav_gen(G, 2, 0.0), nbr, alt, mode=port.MODE_COUNTER, state16=seed_state(9), env=env)
    o.add_agents(pop); o.set_genomes(gz); o.start()
    for k in range(8):
        o.step(float(k))
    g, _ = o.genomes(row)
    assert np.all((g[:, 0] & np.uint64(1)) == 0) and np.all((g[:, 2] & np.uint64(1)) == 0)


@pytest.mark.parametrize("navigate", [False, True])
def test_ooa_nav_gen_pop_class_equals_reference(navigate):
    """The shipped class of BASELINE configs #3 / #5 as a whole: `OoANavGenPop` compiled from populations/OoANavGenPop.cpp where
    it lies (oracle/Makefile), its Climate / Vegetation / Navigation objects filled in memory, Genetics configured the way the
    QDF reader does it (oracle/ref_driver.cpp) -- against the oracle's WELL mode: action order and wiring of the constructor
    (:33-97), preLoop (:160-169), makePopSpecificOffspring with `m_aAgents[iMother].m_iNumBabies++` (:231-245), updateEvent /
    flushEvents with the MultiEvaluator registered as an observer (:59, :179-226).  Agents, genomes and NumBabies of every live
    agent, slot for slot, across a GEO + CLIMATE + VEG + NAV event."""
    from qhg4_b200.icogrid import synthetic_climate
    from qhg4_b200.params import ooa_nav_gen
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=5)
    env = synthetic_climate(xyz, alt, seed=6)
    pop = synthetic_population(8000, alt, seed=6, fertile=True)
    G = 200
    row = 2 * ((G + 63) // 64)
    par = ooa_nav_gen(G, 3, 5e-3)
    rng = np.random.default_rng(3)
    if navigate:
        par.prios["Navigate"] = 10
    occ = np.unique(pop["cell"])
    ports = np.concatenate([occ[:3], rng.choice(occ[occ > 8], 40, replace=False)]).astype(np.int32)
    ptr = np.arange(0, 4 * len(ports) + 1, 4, dtype=np.int32)
    dests = np.concatenate([rng.choice(np.flatnonzero(alt > 0), 4, replace=False) for _ in ports]).astype(np.int32)
    dist = rng.uniform(100, 700, 4 * len(ports))
    bridges = rng.choice(occ, (5, 2), replace=False).astype(np.int32)
    gen0 = rng.integers(0, 2 ** 63, size=(8000, row), dtype=np.int64).astype(np.uint64)
    st = seed_state(9)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st, env=env)
    for q in (r, o):
        q.set_navigation(ports, ptr, dests, dist, bridges)
        q.add_agents(pop)
        q.set_genomes(gen0)
    o.set_genetics_well(*r.genetics_well())
    r.start(); o.start()
    assert np.array_equal(r.capacities(), o.capacities())
    for k in range(12):
        r.step(float(k)); o.step(float(k))
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        og, onb = o.genomes(row)
        assert np.array_equal(r.genomes(row), og), k
        assert np.array_equal(r.num_babies(), onb), k
        if k == 0:
            assert np.array_equal(r.weights(), o.weights())
        if k == 5:  # the sea level rises and the climate cools: drownings, new capacities and weights, rebuilt jump tables
            alt2 = alt - 120.0
            for q in (r, o):
                q.set_env("Altitude", alt2)
                q.set_env("AnnualMeanTemp", env["AnnualMeanTemp"] - 2.0)
                q.set_env("BaseNPP", env["BaseNPP"] * 0.9)
            for ev in (2, 3, 4, 5):  # app/Simulator.cpp:728-735,372-374: updateEvent per id, then one flushEvents
                r.event(ev, 6.0, flush=(ev == 5)); o.update_event(ev, 6.0)
            o.flush_events(6.0)
            assert r.num_agents() == o.num_agents()
            assert np.array_equal(r.capacities(), o.capacities())
    assert np.array_equal(r.weights(), o.weights())
    assert (oa["id"] >= 8000).sum() > 800 and onb.sum() > 0
    r.close()


MULTI_PROBES = {"add_block": "tut_EnvironCapAltAddBlockPop", "mul": "tut_EnvironCapAltMulPop", "max": "tut_EnvironCapAltMaxPop",
                "max_block": "tut_EnvironCapAltMaxBlockPop", "min": "tut_EnvironCapAltMinPop"}


@pytest.mark.parametrize("mode", sorted(MULTI_PROBES))
def test_multi_evaluator_modes_equal_reference(mode):
    """The other five combine modes of MultiEvaluator (actions/MultiEvaluator.cpp:263-298 ADD_BLOCK, 308-342 MUL_SIMPLE,
    350-381 MAX_SIMPLE -- rows NOT cumulated --, 391-432 MAX_BLOCK, 440-478 MIN_SIMPLE) and findBlockings (:579-598) over
    non-cumulating evaluators: the reference's own MultiEvaluator in the probe class MultiProbePop<MODE> (oracle/ref_driver.cpp:
    tut_EnvironCapAltPop with its evaluator rebuilt the way populations/OoANavPop.cpp:50-62 builds the multiplicative one) against
    the oracle's WELL mode.  Weight rows and whole trajectories, incl. the quirks: the BLOCK modes yield all-zero rows until an
    event makes the evaluators recompute twice in one initialize, and an evaluator that is not recomputed contributes zeros."""
    from qhg4_b200.icogrid import synthetic_climate
    from qhg4_b200.params import tut_environ_cap_alt
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=5)
    env = synthetic_climate(xyz, alt, seed=6)
    pop = synthetic_population(8000, alt, seed=6, fertile=True)
    par = tut_environ_cap_alt()
    par.class_name = MULTI_PROBES[mode]
    par.modules["WeightedMove"]["WeightedMove_prob"] = "0.3"
    st = seed_state(19)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st, env=env)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    seen = []
    for k in range(12):
        r.step(float(k)); o.step(float(k))
        assert np.array_equal(r.weights(), o.weights(), equal_nan=True), k
        seen.append(r.weights().copy())
        ra, oa = r.agents(), o.agents()
        for f in FIELDS:
            assert np.array_equal(ra[f], oa[f]), (k, f)
        if k in (3, 7):  # k=3: only the altitude changes (GEO): the NPP evaluator is not recomputed; k=7: both
            alt2 = alt - 60.0 * (k - 1)
            for q in (r, o):
                q.set_env("Altitude", alt2)
                if k == 7:
                    q.set_env("BaseNPP", env["BaseNPP"] * 0.8)
            evs = (2,) if k == 3 else (2, 3, 4)
            for ev in evs:
                r.event(ev, float(k + 1), flush=(ev == evs[-1])); o.update_event(ev, float(k + 1))
            o.flush_events(float(k + 1))
    assert not np.array_equal(seen[0], seen[-1], equal_nan=True)   # the events did change the rows
    if mode in ("add_block", "max_block"):
        assert np.all((seen[0] == 0) | np.isinf(seen[0]))           # nothing is combined before the first event
    r.close()
