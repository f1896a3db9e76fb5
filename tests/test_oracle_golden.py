"""The CPU oracle (oracle/qhg_oracle.cpp) against golden vectors generated from the reference itself
(tests/make_golden.py -> tests/golden/*.npz) and against published known-answer vectors."""
import os

import numpy as np
import pytest

from oracle import port
from qhg4_b200.params import tut_environ_alt

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_well512_sequences():
    g = np.load(os.path.join(GOLD, "well512.npz"))
    assert np.array_equal(port.well_sequence(g["state"], 256), g["seq"])
    assert np.array_equal(port.well_sequence(g["state_thread0"], 256), g["seq_thread0"])
    # SURVEY.md §9.2 known answers
    assert [hex(x) for x in g["seq"][:4]] == ["0x846f945a", "0xf7691ff", "0x7880c84b", "0x8c1a003d"]
    assert [hex(x) for x in g["seq_thread0"][:4]] == ["0x5c4b0026", "0x9a92fa79", "0xf28c9caf", "0x967286a"]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in port.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in port.philox([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in port.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_polyline():
    g = np.load(os.path.join(GOLD, "polyline.npz"))
    d = str(g["definition"])
    assert np.array_equal(port.polyline_eval(d, g["x"], True), g["y_float_cast"])
    assert np.array_equal(port.polyline_eval(d, g["x"], False), g["y_double"])
    with pytest.raises(ValueError):
        port.polyline_eval("1 2 3", [0.0])


def _load_case():
    g = np.load(os.path.join(GOLD, "tut_environ_alt_ico3.npz"))
    pop = {k[4:]: g[k] for k in g.files if k.startswith("pop_")}
    return g, pop


def test_deterministic_substeps_bit_exact():
    g, pop = _load_case()
    o = port.OraclePop(tut_environ_alt(float(g["K"])), g["nbr"], g["alt"], ice=g["ice"].astype(float), mode=port.MODE_WELL)
    o.add_agents(pop)
    o.start()
    assert np.array_equal(o.counts(), g["counts0"])
    assert np.array_equal(o.atan_prob(g["ages"]), g["p_atan"])
    o.step(0.0)
    b, d = o.bd()
    assert np.array_equal(b, g["b_step0"]) and np.array_equal(d, g["d_step0"])
    assert np.array_equal(o.weights(), g["weights"])
    # Verhulst b0 0.8, d0 0.001, theta 0.1, K 20
    n = g["counts0"].astype(float)
    assert np.allclose(b, 0.8 + (0.1 - 0.8) * n / 20.0, rtol=0, atol=1e-15)


def test_trajectory_bit_exact_one_thread():
    """WELL mode = the reference with one OpenMP thread: same agents in the same slots after 12 steps."""
    g, pop = _load_case()
    o = port.OraclePop(tut_environ_alt(float(g["K"])), g["nbr"], g["alt"], ice=g["ice"].astype(float), mode=port.MODE_WELL)
    o.add_agents(pop)
    o.start()
    tot = []
    for k in range(12):
        o.step(float(k))
        tot.append(o.num_agents())
    assert tot == list(g["totals"])
    a = o.agents()
    for f in ("cell", "id", "birth", "gender", "age", "last_birth", "life", "slot"):
        assert np.array_equal(a[f], g["fin_" + f]), f
    assert np.array_equal(o.counts(), g["counts_final"])


def _golden_case_names():
    import golden_cases as gc
    return sorted(gc.CASES)


@pytest.mark.parametrize("name", _golden_case_names())
def test_golden_trajectories_of_the_reference(name):
    """Every population class and probed action (tutorial ladder, tut_EnvironCapAltPop, ConfinedMove, Navigate, OldAgeDeath,
    WeightedMoveRand + SigDeath, Genetics with 1- and 2-bit nucleotides): the oracle's WELL mode replays the trajectory the
    REFERENCE produced with one thread (tests/make_golden.py, fixtures generated from oracle/_ref) -- per-step totals, the
    final agent table slot for slot, per-cell counts, capacities and genomes.  Needs neither /root/reference nor oracle/_ref."""
    import golden_cases as gc
    path = os.path.join(GOLD, f"case_{name}.npz")
    z = np.load(path)
    d = {k: z[k] for k in z.files if not k.startswith("ref_")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}

    def oracle_factory(par, nbr, alt, st, env):
        return port.OraclePop(par, nbr, alt, mode=port.MODE_WELL, state16=st, env=env)

    well = (ref["gen_well_state"], ref["gen_well_index"]) if "gen_well_state" in ref else None
    out = gc.run_case(name, oracle_factory, d, well_from=well if well is not None else ())
    for k, v in ref.items():
        if k.startswith("gen_well"):
            continue
        if k == "fin_last_birth" and name in ("tut_move", "tut_old_age_die"):
            continue  # these classes' agent structs have no such field (populations/tut_MovePop.h, tut_OldAgeDiePop.h)
        assert np.array_equal(out[k], v), (name, k)
    assert ref["totals"][-1] > 0


def test_counter_mode_is_order_invariant():
    """counter mode must not depend on the order in which agents are stored"""
    g, pop = _load_case()
    perm = np.random.default_rng(0).permutation(len(pop["id"]))
    res = []
    for p in (pop, {k: v[perm] for k, v in pop.items()}):
        o = port.OraclePop(tut_environ_alt(float(g["K"])), g["nbr"], g["alt"], mode=port.MODE_COUNTER)
        o.add_agents(p)
        o.start()
        for k in range(10):
            o.step(float(k))
        a = o.agents()
        s = np.argsort(a["id"])
        res.append({f: a[f][s] for f in ("cell", "id", "birth", "gender", "life")})
    for f in res[0]:
        assert np.array_equal(res[0][f], res[1][f]), f


def test_env_interpolation_restates_auto_interpolator():
    """AutoInterpolator::interpolate (core/AutoInterpolator.cpp:461-483): target[i] += iSteps * diff[i] in double, product and
    sum rounded separately.  The oracle against numpy's elementwise arithmetic, and a run whose climate is interpolated step
    by step (events + flush as app/Simulator.cpp:338-374 delivers them) against the same run fed with re-uploaded arrays."""
    from oracle import port
    from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_climate, synthetic_population
    from qhg4_b200.params import seed_state, tut_environ_cap_alt
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=2)
    env = synthetic_climate(xyz, alt, seed=3)
    pop = synthetic_population(6000, alt, seed=4, fertile=True)
    rng = np.random.default_rng(1)
    delta = {"AnnualMeanTemp": rng.normal(-0.3, 0.1, len(alt)), "AnnualRainfall": rng.normal(-15.0, 3.0, len(alt)),
             "BaseNPP": rng.normal(-0.01, 0.003, len(alt)), "Altitude": rng.normal(-4.0, 1.0, len(alt))}
    a = port.OraclePop(tut_environ_cap_alt(), nbr, alt, mode=port.MODE_COUNTER, state16=seed_state(3), env=env)
    b = port.OraclePop(tut_environ_cap_alt(), nbr, alt, mode=port.MODE_COUNTER, state16=seed_state(3), env=env)
    for q in (a, b):
        q.add_agents(pop); q.start()
    for name, d in delta.items():
        a.set_env_delta(name, d)
    cur = dict(env, Altitude=alt.copy())
    for k in range(10):
        a.step(float(k)); b.step(float(k))
        steps = 1 if k % 3 else 2
        a.interpolate_env(steps)
        for name, d in delta.items():
            cur[name] = cur[name] + steps * d          # numpy: one rounded product, one rounded sum per element
            b.set_env(name, cur[name])
            assert np.array_equal(a.env_array(name), cur[name]), (k, name)
        for q in (a, b):
            for ev in (2, 3, 4):
                q.update_event(ev, float(k + 1))
            q.flush_events(float(k + 1))
        assert np.array_equal(a.capacities(), b.capacities()), k
        assert a.num_agents() == b.num_agents(), k
    assert not np.array_equal(a.env_array("Altitude"), alt) and a.num_agents() > 0
    a.set_env_delta("Altitude", None)
    before = a.env_array("Altitude").copy()
    a.interpolate_env(5)
    assert np.array_equal(a.env_array("Altitude"), before)
