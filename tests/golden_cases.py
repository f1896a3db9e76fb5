"""Golden trajectories of every population / action the oracle restates, generated from the REFERENCE itself
(oracle/_ref: the unmodified QHG4 sources, one OpenMP thread) by tests/make_golden.py and replayed against the oracle's
WELL mode by tests/test_oracle_golden.py.  One definition serves both sides, so the fixture holds the inputs as they were
generated and the reference's outputs; the test needs neither /root/reference nor oracle/_ref."""
import numpy as np

from qhg4_b200 import params as P
from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_climate, synthetic_population

FIELDS = ("cell", "id", "birth", "gender", "age", "last_birth", "life", "slot")
NSTEPS = 10


def _nav_par():
    par = P.tut_environ_alt(20.0)
    par.class_name = "tut_EnvironAltNavPop"
    par.modules["Navigate"] = {"Navigate_decay": "-0.001", "Navigate_dist0": "150.0", "Navigate_prob0": "0.1",
                               "Navigate_min_dens": "0.0", "Navigate_bridge_prob": "0.3"}
    par.prios["Navigate"] = 8
    par.modules["OldAgeDeath"] = {"OAD_max_age": "60.0", "OAD_uncertainty": "0.1"}
    return par


def _oad_par():
    par = _nav_par()
    del par.prios["Navigate"]
    del par.prios["ATanDeath"]
    par.prios["OldAgeDeath"] = 2
    return par


def _ooa_par():
    """OoANavGenPop itself (populations/OoANavGenPop.cpp), Navigate included: the class of BASELINE configs #3 / #5"""
    par = P.ooa_nav_gen(100, 3, 5e-3)
    par.prios["Navigate"] = 10
    return par


# name -> (parameter set, needs climate arrays, needs lon/lat, navigation, genome (size, bits) or None, all agents female)
CASES = {
    "tut_sexual": (lambda: P.tut_sexual(25.0, 0.2), False, False, False, None, False),
    "tut_move": (lambda: P.tut_move(0.3), False, False, False, None, False),
    "tut_old_age_die": (lambda: P.tut_old_age_die(), False, False, False, None, False),
    "tut_partheno": (lambda: P.tut_partheno(25.0, 0.2), False, False, False, None, True),
    "tut_environ_cap_alt": (lambda: P.tut_environ_cap_alt(), True, False, False, None, False),
    "confined_move": (lambda: P.tut_environ_alt_confined(60.0, 20.0, 10.0, 6000.0), False, True, False, None, False),
    "navigate": (_nav_par, False, False, True, None, False),
    "old_age_death": (_oad_par, False, False, False, None, False),
    "move_rand_sig_death": (lambda: P.tut_environ_alt_variants(25.0, True, True), False, False, False, None, False),
    "genetics_1bit_free": (lambda: P.tut_environ_alt_genetic(20.0, 100, -1, 2e-3, 1), False, False, False, (100, 1), False),
    "genetics_2bit_cross": (lambda: P.tut_environ_alt_genetic(20.0, 100, 3, 5e-3, 2), False, False, False, (100, 2), False),
    "ooa_nav_gen": (_ooa_par, True, False, True, (100, 1), False),
    # CondWeightedMove ("different" / "less or equal" over the altitudes), RandPermPair, MoveStats (first / minimum) -- ExtProbePop<m>
    "cond_move_perm_pair_stats_first": (lambda: P.tut_environ_alt_ext(150.0, 7, True, 0, 0.3), False, True, False, None, False),
    "cond_move_stats_min": (lambda: P.tut_environ_alt_ext(150.0, 6, False, 1, 0.3), False, True, False, None, False),
    "perm_pair_stats_last": (lambda: P.tut_environ_alt_ext(150.0, -1, True, 2, 0.3), False, True, False, None, False),
}


def build_inputs(name):
    """the inputs of a case, generated (make_golden.py stores them in the fixture; the test reads them from there)"""
    par_fn, climate, lonlat, nav, genome, female = CASES[name]
    nbr, xyz = make_ico_grid(5)
    alt = synthetic_altitude(xyz, seed=4)
    if name == "move_rand_sig_death":
        alt[(alt > 400) & (alt < 900)] = 2600.0
    if name == "confined_move":
        alt = np.minimum(np.abs(alt) + 50.0, 1400.0)
    d = {"nbr": nbr, "alt": alt}
    seed = sum(map(ord, name)) % 97
    cells = None
    if "stats" in name:  # MoveStats: a home range, so that cells are reached for the first time during the run
        land = np.flatnonzero(alt > 0)
        cells = land[np.argsort(xyz[land, 0])[: len(land) // 5]]
    pop = synthetic_population(4000, alt, seed=11 + seed, fertile=True, max_age=70.0 if "death" in name else 60.0, cells=cells)
    if female:
        pop["gender"][:] = 0
    for k, v in pop.items():
        d["pop_" + k] = v
    env = {}
    if climate:
        env = synthetic_climate(xyz, alt, seed=3)
    if lonlat:
        env = {"Longitude": np.degrees(np.arctan2(xyz[:, 1], xyz[:, 0])), "Latitude": np.degrees(np.arcsin(np.clip(xyz[:, 2], -1, 1)))}
    for k, v in env.items():
        d["env_" + k] = np.asarray(v, np.float64)
    if nav:
        rng = np.random.default_rng(3)
        occ = np.unique(pop["cell"])
        ports = np.concatenate([occ[:3], rng.choice(occ[occ > 8], 30, replace=False)]).astype(np.int32)
        d["nav_ports"] = ports
        d["nav_ptr"] = np.arange(0, 4 * len(ports) + 1, 4, dtype=np.int32)
        d["nav_dests"] = np.concatenate([rng.choice(np.flatnonzero(alt > 0), 4, replace=False) for _ in ports]).astype(np.int32)
        d["nav_dist"] = rng.uniform(100, 700, 4 * len(ports))
        d["nav_bridges"] = rng.choice(occ, (4, 2), replace=False).astype(np.int32)
    if genome:
        G, bits = genome
        row = 2 * ((G * bits + 63) // 64)
        d["genomes"] = np.random.default_rng(5).integers(0, 2 ** 63, size=(4000, row), dtype=np.int64).astype(np.uint64)
    d["seed_state"] = P.seed_state(40 + seed)
    return d


def run_case(name, sim_factory, d, well_from=None):
    """drive one simulator (reference or oracle) through the case; returns its outputs"""
    par = CASES[name][0]()
    env = {k[4:]: d[k] for k in d if k.startswith("env_")}
    pop = {k[4:]: d[k] for k in d if k.startswith("pop_")}
    s = sim_factory(par, d["nbr"], d["alt"], d["seed_state"], env or None)
    if "nav_ports" in d:
        s.set_navigation(d["nav_ports"], d["nav_ptr"], d["nav_dests"], d["nav_dist"], d["nav_bridges"])
    s.add_agents(pop)
    out = {}
    if "genomes" in d:
        s.set_genomes(d["genomes"])
        if well_from is None:
            st, idx = s.genetics_well()
            out["gen_well_state"], out["gen_well_index"] = st, np.array([idx], np.uint32)
        else:
            s.set_genetics_well(well_from[0], int(well_from[1][0]))
    s.start()
    totals = []
    for k in range(NSTEPS):
        s.step(float(k))
        totals.append(s.num_agents())
    out["totals"] = np.array(totals)
    a = s.agents()
    for f in FIELDS:
        out["fin_" + f] = a[f]
    out["counts_final"] = np.asarray(s.counts())
    if "genomes" in d:
        g = s.genomes(d["genomes"].shape[1])
        out["fin_genomes"] = g[0] if isinstance(g, tuple) else g
        if par.class_name.startswith("OoANavGen"):  # m_iNumBabies, populations/OoANavGenPop.cpp:243
            out["fin_nbabies"] = g[1] if isinstance(g, tuple) else s.num_babies()
    if CASES[name][1]:
        out["capacities"] = s.capacities()
    if par.prios.get("MoveStats") is not None:  # m_aiHops, m_adDist, m_adTime (actions/MoveStats.h:49-51)
        out["move_hops"], out["move_dist"], out["move_time"] = s.move_stats()
    return out
