"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the reference.

* deterministic sub-steps (per-cell counts, Verhulst b/d, cell weights, ATanDeath probability):
  bit-exact against the oracle, within 1e-6 relative of the reference's own arrays (oracle/_ref);
* whole trajectories: bit-exact against the oracle's counter mode (same random streams), compared
  as sets of agents because the order inside a cell is not part of the result.
"""
import os

import numpy as np
import pytest

from conftest import sort_agents
from qhg4_b200.icogrid import make_ico_grid, make_torus_grid, synthetic_altitude, synthetic_population
from qhg4_b200.params import seed_state, tut_environ_alt

pytestmark = pytest.mark.gpu

FIELDS = ("cell", "id", "birth", "gender", "age", "last_birth", "life")


def make_pair(params, nbr, alt, pop, ice=None, seed=0, env=None):
    from oracle import port
    from qhg4_b200.population import GpuPopulation
    st = seed_state(seed)
    g = GpuPopulation.from_params(params, nbr, alt, ice=ice, state16=st, env=env)
    o = port.OraclePop(params, nbr, alt, ice=ice, mode=port.MODE_COUNTER, state16=st, env=env)
    g.add_agents(pop)
    o.add_agents(pop)
    g.pre_loop()
    o.start()
    return g, o


def assert_same_population(g, o, step):
    assert g.num_agents() == o.num_agents(), f"step {step}: {g.num_agents()} vs {o.num_agents()}"
    ga, oa = sort_agents(g.agents()), sort_agents(o.agents())
    for f in FIELDS:
        assert np.array_equal(ga[f], oa[f]), f"step {step}: field {f} differs"
    assert np.array_equal(g.counts(), o.counts()), f"step {step}: per-cell counts differ"


@pytest.fixture(params=["tiled", "tiled-cell", "generic"])
def path(request, monkeypatch):
    """the device paths: the fast one with pass 1 by batches of cells (qhg_decide.cuh, the default) and with one warp per cell
    (qhg_cells.cuh, QHG_DECIDE=cell), and the one-thread-per-agent one"""
    monkeypatch.setenv("QHG_B200_PATH", "generic" if request.param == "generic" else "tiled")
    if request.param == "tiled-cell":
        monkeypatch.setenv("QHG_DECIDE", "cell")
    return request.param


def test_trajectory_bit_exact_vs_oracle(small_world, path):
    nbr, xyz, alt = small_world
    pop = synthetic_population(40000, alt, seed=5)
    g, o = make_pair(tut_environ_alt(20.0), nbr, alt, pop, seed=11)
    assert_same_population(g, o, -1)
    for k in range(25):
        g.step(float(k))
        o.step(float(k))
        assert_same_population(g, o, k)
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), f"step {k}"
        gb, gd = g.bd()
        ob, od = o.bd()
        assert np.array_equal(gb, ob) and np.array_equal(gd, od)
    assert np.array_equal(g.weights(), o.weights())


def test_mates_match_oracle(small_world, path):
    nbr, xyz, alt = small_world
    pop = synthetic_population(30000, alt, seed=6)
    pop["life"][:] = 5  # everybody fertile from the start so that step 0 already pairs
    g, o = make_pair(tut_environ_alt(20.0), nbr, alt, pop, seed=2)
    for k in range(3):
        g.initialize_step(float(k))
        o.initialize_step(float(k))
        ga, oa = sort_agents(g.agents()), sort_agents(o.agents())
        assert np.array_equal(ga["mate_id"], oa["mate_id"])
        if k == 0:
            assert (ga["mate_id"] >= 0).sum() > 1000
        for lvl in sorted(set(g.prios.values())):
            g.do_actions(lvl, float(k))
            o.do_actions(lvl, float(k))
        g.finalize_step()
        o.finalize_step()
        assert_same_population(g, o, k)


def test_deterministic_substeps_vs_reference(small_world):
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    nbr, xyz, alt = small_world
    pop = synthetic_population(40000, alt, seed=7)
    par = tut_environ_alt(20.0)
    from qhg4_b200.population import GpuPopulation
    g = GpuPopulation.from_params(par, nbr, alt)
    g.add_agents(pop)
    g.pre_loop()
    r = refsim.RefSim(par, nbr, alt, threads=2)
    r.add_agents(pop)
    r.start()
    assert np.array_equal(g.counts(), r.counts())            # per-cell counts: bit-exact
    g.initialize_step(0.0)
    r.step(0.0)                                               # the reference computes b, d, weights in its initialize
    gb, gd = g.bd()
    rb, rd = r.bd()
    np.testing.assert_allclose(gb, rb, rtol=1e-6, atol=0)     # north-star tolerance: 1e-6 relative
    np.testing.assert_allclose(gd, rd, rtol=1e-6, atol=0)
    np.testing.assert_allclose(g.weights(), r.weights(), rtol=1e-6, atol=0)
    ages = np.linspace(0, 90, 2001).astype(np.float32)
    np.testing.assert_allclose(g.atan_prob(ages), r.atan_prob(ages), rtol=1e-6, atol=1e-12)
    r.close()


def test_rebinning_keeps_every_agent(small_world):
    """compaction + re-binning: with no births and deaths the agent set is preserved bit for bit and sorted by cell."""
    nbr, xyz, alt = small_world
    pop = synthetic_population(50000, alt, seed=8)
    par = tut_environ_alt(20.0)
    for name in ("ATanDeath", "Verhulst", "Fertility", "RandomPair"):
        del par.prios[name]
    par.modules["WeightedMove"]["WeightedMove_prob"] = "0.9"
    g, o = make_pair(par, nbr, alt, pop, seed=4)
    for k in range(5):
        g.step(float(k))
        o.step(float(k))
        a = g.agents()
        assert np.all(np.diff(a["cell"]) >= 0), "agents are not binned by cell"
        assert np.array_equal(np.sort(a["id"]), np.arange(50000))
        assert_same_population(g, o, k)
        assert g.step_stats().moves > 30000


def test_geo_event_kills_drowned(small_world, path):
    nbr, xyz, alt = small_world
    pop = synthetic_population(30000, alt, seed=9)
    g, o = make_pair(tut_environ_alt(20.0), nbr, alt, pop, seed=3)
    for k in range(3):
        g.step(float(k)); o.step(float(k))
    alt2 = alt - 400.0  # sea level rises
    ice = (xyz[:, 2] > 0.8).astype(np.float64)
    g.set_env("Altitude", alt2); g.set_env("Ice", ice)
    o.set_env("Altitude", alt2); o.set_env("Ice", ice)
    g.update_event(2, 3.0); g.flush_events(3.0)
    o.update_event(2, 3.0)
    assert_same_population(g, o, "event")
    a = g.agents()
    assert np.all(alt2[a["cell"]] >= 0) and np.all(ice[a["cell"]] == 0)
    for k in range(3, 8):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
    assert np.array_equal(g.weights(), o.weights())


def test_pentagon_and_ocean_cells(path):
    """edge cases: 5-neighbour cells, cells whose whole neighbourhood has weight 0 (uniform pick), empty cells."""
    nbr, xyz = make_ico_grid(3)
    alt = np.full(len(nbr), -100.0)      # everything below sea level: all weights 0
    alt[:40] = 500.0
    pop = synthetic_population(5000, alt, seed=1, cells=np.arange(len(nbr)))
    g, o = make_pair(tut_environ_alt(30.0), nbr, alt, pop, seed=5)
    for k in range(12):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)


def test_crowded_cells_fall_back_to_generic_path():
    """cells with more agents than a tile holds (2048): the step is redone on the generic path, same result"""
    nbr, xyz = make_ico_grid(3)
    alt = np.full(len(nbr), 800.0)
    pop = synthetic_population(60000, alt, seed=3, cells=np.array([5, 6, 7, 40, 41, 100]))
    g, o = make_pair(tut_environ_alt(9000.0), nbr, alt, pop, seed=8)
    for k in range(4):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)


def test_run_queues_steps_without_host_round_trips(small_world):
    """qhgb_run queues its steps on the stream and looks at the device's counters once per window: same agents, counts and
    per-step totals as the same number of qhgb_step calls and as the oracle; the device-side agent-step sum is exact."""
    nbr, xyz, alt = small_world
    pop = synthetic_population(40000, alt, seed=5, fertile=True)
    g, o = make_pair(tut_environ_alt(20.0), nbr, alt, pop, seed=11)
    g2, _ = make_pair(tut_environ_alt(20.0), nbr, alt, pop, seed=11)
    expect = 0
    for k in range(13):
        expect += o.num_agents()
        o.step(float(k)); g2.step(float(k))
    g.run(0.0, 13)
    assert_same_population(g, o, 12)
    s, s2 = g.step_stats(), g2.step_stats()
    assert (s.births, s.deaths, s.moves) == o.step_stats() == (s2.births, s2.deaths, s2.moves)
    assert (s.next_id, s.steps_done) == (s2.next_id, s2.steps_done) == (s2.next_id, 13)
    assert g.run_totals()[0] == expect == g2.run_totals()[0]
    g.step(13.0); o.step(13.0)          # and a plain step after a queued run
    assert_same_population(g, o, 13)
    g.run(14.0, 1); o.step(14.0)
    assert_same_population(g, o, 14)


def test_run_recovers_when_a_queued_step_cannot_complete():
    """A queued step that the fast path cannot finish raises a flag on the device, the steps queued after it do nothing, and
    the host redoes it the way qhgb_step would: (a) cells growing past the fast path's 1024 agents -> generic path,
    (b) births outgrowing the agent buffers -> larger buffers.  Results stay those of the oracle."""
    nbr, xyz = make_ico_grid(3)
    alt = np.full(len(nbr), 800.0)
    pop = synthetic_population(5400, alt, seed=3, cells=np.array([5, 6, 7, 40, 41, 100]), fertile=True)  # 900 per cell, growing
    g, o = make_pair(tut_environ_alt(9000.0), nbr, alt, pop, seed=8)
    g.run(0.0, 7)
    for k in range(7):
        o.step(float(k))
    assert_same_population(g, o, 6)
    assert g.counts().max() > 1024 and g.run_totals()[0] > 7 * 5400
    nbr, xyz = make_ico_grid(7)
    alt = np.full(len(nbr), 800.0)
    pop = synthetic_population(30000, alt, seed=4, fertile=True)                                   # ~47 per cell, +15 % per step
    from qhg4_b200.population import GpuPopulation
    from oracle import port
    st = seed_state(5)
    g = GpuPopulation.from_params(tut_environ_alt(9000.0), nbr, alt, state16=st, capacity_hint=40000)
    o = port.OraclePop(tut_environ_alt(9000.0), nbr, alt, mode=port.MODE_COUNTER, state16=st)
    g.add_agents(pop); o.add_agents(pop); g.pre_loop(); o.start()
    g.run(0.0, 9)
    for k in range(9):
        o.step(float(k))
    assert_same_population(g, o, 8)
    assert g.num_agents() > 60000


def test_dense_and_sparse_cells_mix():
    """tile boundaries: dense cells next to long runs of empty cells, tiles of very different cell counts"""
    nbr, xyz = make_ico_grid(31)
    alt = synthetic_altitude(xyz, seed=2)
    land = np.flatnonzero(alt > 0)
    rng = np.random.default_rng(0)
    hot = rng.choice(land, 40, replace=False)
    pop_a = synthetic_population(30000, alt, seed=4, cells=hot)            # ~750 per cell
    pop_b = synthetic_population(20000, alt, seed=5)                       # ~3 per land cell
    pop = {k: np.concatenate([pop_a[k], pop_b[k]]) for k in pop_a}
    pop["id"] = np.arange(len(pop["id"]), dtype=np.int64)
    g, o = make_pair(tut_environ_alt(600.0), nbr, alt, pop, seed=9)
    for k in range(8):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)


def test_occupancy_of_tracked_cells(small_world):
    """qhgb_get_occupied = OccTracker::calcBitMap for one population (core/OccTracker.cpp:95-106): one byte per tracked cell,
    equal to (count > 0) of the per-cell counts, which are bit-exact against the reference (test_deterministic_substeps)."""
    nbr, xyz, alt = small_world
    pop = synthetic_population(6000, alt, seed=8)   # sparse: a good part of the tracked cells is empty
    g, o = make_pair(tut_environ_alt(20.0), nbr, alt, pop, seed=3)
    cells = np.random.default_rng(1).choice(len(nbr), 500, replace=False).astype(np.int32)
    for k in range(4):
        g.step(float(k)); o.step(float(k))
        assert np.array_equal(g.occupied(cells), (o.counts()[cells] > 0).astype(np.uint8)), k
    with pytest.raises(Exception):
        g.occupied(np.array([len(nbr)], np.int32))


def test_path_counts_tell_which_kernels_ran(small_world):
    """qhgb_get_path_counts: the tutorial population steps on the fast path; a population with a generic-path action
    (WeightedMoveRand) does not; a single GPU never needs the recovery kernels"""
    from qhg4_b200.params import tut_environ_alt_variants
    nbr, xyz, alt = small_world
    pop = synthetic_population(20000, alt, seed=4)
    g, o = make_pair(tut_environ_alt(20.0), nbr, alt, pop, seed=5)
    f0, g0, r0 = g.path_counts()
    for k in range(4):
        g.step(float(k))
    f1, g1, r1 = g.path_counts()
    assert f1 - f0 == 4 and g1 == g0 and r1 == 0
    g2, o2 = make_pair(tut_environ_alt_variants(25.0, True, False), nbr, alt, pop, seed=5)
    a0 = g2.path_counts()
    for k in range(3):
        g2.step(float(k))
    a1 = g2.path_counts()
    assert a1[1] - a0[1] == 3 and a1[0] == a0[0] and a1[2] == 0


def test_count_mirror_is_current_after_every_step(small_world):
    """qhgb_mirror_num_agents_array: the host array is what qhgb_get_num_agents_array would return after every step, event and
    window of queued steps (m_aiNumAgentsPerCell of the reference is always current: core/SPopulation.cpp:1250-1290)"""
    nbr, xyz, alt = small_world
    pop = synthetic_population(30000, alt, seed=6)
    g, o = make_pair(tut_environ_alt(20.0), nbr, alt, pop, seed=2)
    lo, hi = 100, 2000
    m = g.mirror_counts(g.host_array(hi - lo, np.uint64), lo, hi)
    assert np.array_equal(m, o.counts()[lo:hi])
    for k in range(5):
        g.step(float(k)); o.step(float(k))
        assert np.array_equal(m, o.counts()[lo:hi]), k
    alt2 = alt.copy(); alt2[::3] = -5.0
    g.set_env("Altitude", alt2); o.set_env("Altitude", alt2)
    g.update_event(2, 5.0); g.flush_events(5.0)         # GEO: agents on the new sea cells drown
    o.update_event(2, 5.0)
    assert np.array_equal(m, o.counts()[lo:hi])
    g.run(5.0, 6)
    for k in range(6):
        o.step(5.0 + k)
    assert np.array_equal(m, o.counts()[lo:hi])
    g.mirror_counts(None)
    g.step(11.0); o.step(11.0)
    assert not np.array_equal(m, o.counts()[lo:hi])     # no longer refreshed
    assert_same_population(g, o, 11)


def test_empty_population_and_late_agents(small_world):
    nbr, xyz, alt = small_world
    from qhg4_b200.population import GpuPopulation
    g = GpuPopulation.from_params(tut_environ_alt(20.0), nbr, alt)
    g.pre_loop()
    g.step(0.0)
    assert g.num_agents() == 0 and g.counts().sum() == 0
    pop = synthetic_population(1000, alt, seed=2)
    g.add_agents(pop)
    assert g.num_agents() == 1000 and g.counts().sum() == 1000
    g.step(1.0)
    assert 800 < g.num_agents() <= 1000


def _run_summaries(make_sim, par, nbr, alt, nseeds, nsteps):
    """one vector of summary statistics per seed (independent samples): live agents, mean / 90th percentile / standard deviation
    of the ages, share of occupied cells, variance of the per-cell counts, share of males; plus the pooled ages and counts"""
    rows, ages, occ = [], [], []
    for s in range(nseeds):
        pop = synthetic_population(6000, alt, seed=100 + s)
        sim = make_sim(par, nbr, alt, seed_state(1000 + s), pop)
        for k in range(nsteps):
            sim.step(float(k))
        a, c = sim.agents(), np.asarray(sim.counts(), np.float64)
        rows.append([sim.num_agents(), a["age"].mean(), np.percentile(a["age"], 90), a["age"].std(), (c > 0).mean(), c.var(), a["gender"].mean()])
        ages.append(a["age"]); occ.append(c)
        sim.close()
    return np.array(rows), np.concatenate(ages), np.concatenate(occ)


def test_statistical_equivalence_vs_reference():
    """Stochastic sub-steps, device against the reference with two OpenMP threads over 32 seeds.  The gate is a two-sample KS test,
    p > 0.01, on statistics whose samples are INDEPENDENT -- one value per seed: population size, mean, 90th percentile and spread
    of the ages, share of occupied cells, variance of the per-cell counts, sex ratio -- so the whole sample counts (no thinning).
    The pooled per-agent ages and per-cell counts (correlated inside a run, hence thinned) are checked as before.
    Negative control: the same device runs with Verhulst_K = 23 instead of 20 must FAIL the gate."""
    from scipy import stats
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    from qhg4_b200.population import GpuPopulation
    nbr = make_torus_grid(24, 24)
    alt = np.full(len(nbr), 800.0)
    alt[::7] = 1400.0
    nseeds, nsteps = 32, 40

    def gpu(par, nbr_, alt_, st, pop):
        g = GpuPopulation.from_params(par, nbr_, alt_, state16=st)
        g.add_agents(pop); g.pre_loop()
        return g

    def ref(par, nbr_, alt_, st, pop):
        r = refsim.RefSim(par, nbr_, alt_, threads=2, state16=st)
        r.add_agents(pop); r.start()
        return r

    sg, age_g, occ_g = _run_summaries(gpu, tut_environ_alt(20.0), nbr, alt, nseeds, nsteps)
    sr, age_r, occ_r = _run_summaries(ref, tut_environ_alt(20.0), nbr, alt, nseeds, nsteps)
    ps = [stats.ks_2samp(sg[:, j], sr[:, j]).pvalue for j in range(sg.shape[1])]
    assert min(ps) > 0.01, ps
    p_age = stats.ks_2samp(age_g[::7], age_r[::7]).pvalue
    p_occ = stats.ks_2samp(occ_g[::3], occ_r[::3]).pvalue
    assert p_age > 0.01 and p_occ > 0.01, (p_age, p_occ)
    # the gate has power: a 15 % larger carrying capacity is told apart
    sx, _, _ = _run_summaries(gpu, tut_environ_alt(23.0), nbr, alt, nseeds, nsteps)
    px = [stats.ks_2samp(sx[:, j], sr[:, j]).pvalue for j in range(sx.shape[1])]
    assert min(px) < 0.01, px


@pytest.mark.parametrize("bits,ncross", [(1, -1), (2, 2)])
def test_genotype_distributions_statistically_equivalent_to_reference(bits, ncross):
    """Genotype side of the statistical gate, device against the reference's own Genetics<T,U> (two threads) over 32 seeds:
    allele frequency per run, per-agent heterozygosity, switches along a strand (recombination breaks + mutations; founder
    strands are all-0 / all-1), population size -- two-sample KS, p > 0.01.  Same statistics as the CPU test of the oracle's
    counter mode (tests/test_oracle_vs_ref.py), which also shows that a wrong mutation rate fails them."""
    from scipy import stats
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    from test_oracle_vs_ref import _genotype_samples, _ref_sim
    from qhg4_b200.population import GpuPopulation

    def gpu_sim(par, nbr, alt, st, pop, gen0):
        g = GpuPopulation.from_params(par, nbr, alt, state16=st)
        g.add_agents(pop); g.set_genomes(gen0); g.pre_loop()
        return g

    G = 96
    fg, hg, sg, tg = _genotype_samples(gpu_sim, 32, 25, G, bits, ncross, 2e-3)
    fr, hr, sr, tr = _genotype_samples(_ref_sim(2), 32, 25, G, bits, ncross, 2e-3)
    ps = [stats.ks_2samp(fg, fr).pvalue, stats.ks_2samp(hg[::13], hr[::13]).pvalue, stats.ks_2samp(sg[::29], sr[::29]).pvalue,
          stats.ks_2samp(tg, tr).pvalue]
    assert min(ps) > 0.01, ps
    assert sg.mean() > 1.0


@pytest.mark.parametrize("which", ["tutorial", "genetic", "rebalance", "rebalance-genetic", "bigcell"])
@pytest.mark.parametrize("exchange", ["peer-memory", "nccl"])
def test_two_gpu_shards_equal_unsharded_oracle(exchange, which):
    """cell-range sharding on 2 GPUs, migration over peer memory (default) and over NCCL calls: bit-identical to the
    unsharded oracle (tests/mgpu_check.py) -- the tutorial population, and OoANavGenPop with Navigate (genome rows travel with
    the migrants, far jumps cross shard boundaries), a re-split in the middle of a run, and cells far beyond the fast path's limits
    (every rank redoes those steps with the recovery kernels).  The driver's GPU box has one GPU; the logs of the runs on 2, 4 and 8
    B200 are kept in profiles/mgpu_check_r02.txt."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run tests/mgpu_check.py under torchrun on a multi-GPU box)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QHG_P2P="1" if exchange == "peer-memory" else "0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533" if exchange == "nccl" else "29534", os.path.join(root, "tests", "mgpu_check.py")] +
                       ([which] if which != "tutorial" else []), capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "mgpu_check ok" in r.stdout and exchange in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _cap_world(S=15, seed=2):
    from qhg4_b200.icogrid import synthetic_climate
    nbr, xyz = make_ico_grid(S)
    alt = synthetic_altitude(xyz, seed=seed)
    return nbr, xyz, alt, synthetic_climate(xyz, alt, seed=seed + 1)


def test_cap_alt_population_vs_oracle_and_reference(path):
    """tut_EnvironCapAltPop (NPPCapacity, MultiEvaluator[NPP+Alt], VerhulstVarK): capacities / weights / b,d bit-exact
    against the oracle and within 1e-6 of the reference; trajectory bit-exact against the oracle, incl. climate events."""
    from oracle import port, refsim
    from qhg4_b200.params import tut_environ_cap_alt
    from qhg4_b200.population import GpuPopulation
    nbr, xyz, alt, env = _cap_world()
    pop = synthetic_population(40000, alt, seed=4, fertile=True)
    par, st = tut_environ_cap_alt(), seed_state(17)
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
    g.add_agents(pop); o.add_agents(pop)
    g.pre_loop(); o.start()
    assert np.array_equal(g.capacities(), o.capacities())
    if refsim.available():
        r = refsim.RefSim(par, nbr, alt, threads=2, env=env)
        r.add_agents(pop); r.start()
        np.testing.assert_allclose(g.capacities(), r.capacities(), rtol=1e-6, atol=1e-12)
        g.initialize_step(0.0); r.step(0.0)
        np.testing.assert_allclose(g.weights(), r.weights(), rtol=1e-6, atol=1e-12)
        gb, gd = g.bd(); rb, rd = r.bd()
        np.testing.assert_allclose(gb, rb, rtol=1e-6); np.testing.assert_allclose(gd, rd, rtol=1e-6)
        for lvl in sorted(set(g.prios.values())):
            g.do_actions(lvl, 0.0)
        g.finalize_step(); o.step(0.0)
        r.close()
    else:
        g.step(0.0); o.step(0.0)
    assert_same_population(g, o, 0)
    for k in range(1, 10):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
    env2 = dict(env, AnnualMeanTemp=env["AnnualMeanTemp"] - 6.0, AnnualRainfall=env["AnnualRainfall"] * 0.6, BaseNPP=env["BaseNPP"] * 0.7)
    for name in ("AnnualMeanTemp", "AnnualRainfall", "BaseNPP"):
        g.set_env(name, env2[name]); o.set_env(name, env2[name])
    g.update_event(3, 10.0); g.update_event(4, 10.0); g.flush_events(10.0)
    o.update_event(3, 10.0); o.update_event(4, 10.0); o.flush_events(10.0)
    assert np.array_equal(g.capacities(), o.capacities())
    for k in range(10, 18):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
    assert np.array_equal(g.weights(), o.weights())
    gb, gd = g.bd(); ob, od = o.bd()
    assert np.array_equal(gb, ob) and np.array_equal(gd, od)


@pytest.mark.parametrize("ncross,mut", [(-1, 1e-3), (0, 0.0), (2, 1e-3), (7, 5e-3)])
def test_genetic_population_bit_exact_vs_oracle(ncross, mut, path):
    """OoANavGenPop without Navigate (config C3): OldAgeDeath, VerhulstVarK, NPPCapacity, MultiEvaluator[Alt+NPP] and
    Genetics<BitGeneUtils> -- agents AND genomes bit-exact against the oracle's counter mode, for free recombination,
    no recombination and crossovers, with mutations."""
    from oracle import port
    from qhg4_b200.params import ooa_nav_gen
    from qhg4_b200.population import GpuPopulation
    nbr, xyz, alt, env = _cap_world(S=7, seed=5)
    pop = synthetic_population(12000, alt, seed=6, fertile=True)
    G = 200 if ncross == 7 else 256
    par, st = ooa_nav_gen(G, ncross, mut), seed_state(31)
    row = 2 * ((G + 63) // 64)
    gen0 = np.random.default_rng(1).integers(0, 2 ** 63, size=(len(pop["id"]), row), dtype=np.int64).astype(np.uint64)
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
    g.add_agents(pop); o.add_agents(pop)
    g.set_genomes(gen0); o.set_genomes(gen0)
    g.pre_loop(); o.start()
    births = 0
    for k in range(12):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        births += g.step_stats().births
        ga, oa = g.agents(), o.agents()
        gg, gnb = g.genomes(row)
        og, onb = o.genomes(row)
        si, so = np.argsort(ga["id"]), np.argsort(oa["id"])
        assert np.array_equal(gg[si], og[so]), f"step {k}: genomes differ"
        assert np.array_equal(gnb[si], onb[so]), f"step {k}: NumBabies differ"
    assert births > 1500
    # founders keep their genomes; newborn genomes are built from existing alleles (plus rare mutations)
    ga = g.agents()
    gg, _ = g.genomes(row)
    founders = ga["id"] < len(pop["id"])
    assert np.array_equal(gg[founders], gen0[ga["id"][founders]])


@pytest.mark.parametrize("cls,bits", [("OoANavGen2bitPop", 2), ("tut_EnvironAltGen2bitPop", 2), ("tut_EnvironAltGenPop", 1)])
@pytest.mark.parametrize("ncross,mut", [(-1, 2e-3), (3, 5e-3)])
def test_two_bit_genomes_and_genetics_probe_classes_bit_exact_vs_oracle(cls, bits, ncross, mut, path):
    """Genetics<.., GeneUtils> (2-bit nucleotides, genes/GeneUtils.cpp: breaks on nucleotide boundaries, doubled mask bits,
    mutation = XOR with 01/10/11) in OoANavGen2bitPop (populations/OoANavGen2bitPop.cpp) and in the probe classes
    tut_EnvironAlt + Genetics, whose oracle WELL mode equals the reference's own Genetics<T,U> genome for genome
    (tests/test_oracle_vs_ref.py::test_genetics_action_equals_reference): agents and genomes against the counter mode."""
    from oracle import port
    from qhg4_b200.params import ooa_nav_gen, tut_environ_alt_genetic
    from qhg4_b200.population import GpuPopulation
    nbr, xyz, alt, env = _cap_world(S=7, seed=5)
    pop = synthetic_population(12000, alt, seed=6, fertile=True)
    G = 200
    if cls == "OoANavGen2bitPop":
        par = ooa_nav_gen(G, ncross, mut)
        par.class_name = cls
        par.modules["Genetics"]["Genetics_bits_per_nuc"] = "2"
    else:
        par, env = tut_environ_alt_genetic(20.0, G, ncross, mut, bits), None
    st = seed_state(37)
    row = 2 * ((G * bits + 63) // 64)
    gen0 = np.random.default_rng(1).integers(0, 2 ** 63, size=(len(pop["id"]), row), dtype=np.int64).astype(np.uint64)
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
    g.add_agents(pop); o.add_agents(pop)
    g.set_genomes(gen0); o.set_genomes(gen0)
    g.pre_loop(); o.start()
    births = 0
    for k in range(10):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        births += g.step_stats().births
        ga, oa = g.agents(), o.agents()
        gg, _ = g.genomes(row)
        og, _ = o.genomes(row)
        assert np.array_equal(gg[np.argsort(ga["id"])], og[np.argsort(oa["id"])]), f"step {k}: genomes differ"
    assert births > 1500
    with pytest.raises(Exception):  # the nucleotide width belongs to the class (actions/Genetics.cpp:204,259)
        GpuPopulation.from_params(par, nbr, alt, state16=st, env=env).modify_attributes("Genetics_bits_per_nuc", 3 - bits)


@pytest.mark.parametrize("cls,bits", [("OoANavGenPop", 1), ("OoANavGen2bitPop", 2), ("tut_EnvironAltGenPop", 1)])
@pytest.mark.parametrize("ncross,mut", [(-1, 2e-3), (3, 5e-3), (0, 0.0)])
def test_genetic_populations_on_the_fast_path(cls, bits, ncross, mut, monkeypatch):
    """The default path of populations with Genetics: k_cell_decide<false, true> ranks both sexes and hands every birth its father, k_cell_scatter<true> moves
    the genome handles and writes the birth records -- agents, genomes and NumBabies must equal the oracle's (and hence the
    generic path's)."""
    from oracle import port
    from qhg4_b200.params import ooa_nav_gen, tut_environ_alt_genetic
    from qhg4_b200.population import GpuPopulation
    nbr, xyz, alt, env = _cap_world(S=7, seed=5)
    pop = synthetic_population(12000, alt, seed=6, fertile=True)
    G = 200
    if cls.startswith("OoANavGen"):
        par = ooa_nav_gen(G, ncross, mut)
        par.class_name = cls
        par.modules["Genetics"]["Genetics_bits_per_nuc"] = str(bits)
    else:
        par, env = tut_environ_alt_genetic(20.0, G, ncross, mut, bits), None
    st = seed_state(43)
    row = 2 * ((G * bits + 63) // 64)
    gen0 = np.random.default_rng(1).integers(0, 2 ** 63, size=(len(pop["id"]), row), dtype=np.int64).astype(np.uint64)
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
    g.add_agents(pop); o.add_agents(pop)
    g.set_genomes(gen0); o.set_genomes(gen0)
    g.pre_loop(); o.start()
    g.reset_kernel_times(True)
    births = 0
    for k in range(10):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        births += g.step_stats().births
        ga, oa = g.agents(), o.agents()
        gg, gnb = g.genomes(row)
        og, onb = o.genomes(row)
        si, so = np.argsort(ga["id"]), np.argsort(oa["id"])
        assert np.array_equal(gg[si], og[so]), f"step {k}: genomes differ"
        assert np.array_equal(gnb[si], onb[so]), f"step {k}: NumBabies differ"
    assert births > 1500 and "k_cell_decide_genetic" in g.kernel_times()


def test_navigate_sea_crossings_bit_exact_vs_oracle(path):
    """OoANavGenPop WITH Navigate (config C5's action): jumps from port cells to far cells with distance-dependent
    probability, manual bridges, NAV/GEO events rebuilding the tables -- agents and genomes bit-exact against the oracle."""
    from oracle import port
    from qhg4_b200.params import ooa_nav_gen
    from qhg4_b200.population import GpuPopulation
    nbr, xyz, alt, env = _cap_world(S=7, seed=5)
    rng = np.random.default_rng(3)
    land = np.flatnonzero(alt > 0)
    pop = synthetic_population(12000, alt, seed=6, fertile=True)
    occupied = np.unique(pop["cell"])
    ports = rng.choice(occupied[occupied > 8], 60, replace=False).astype(np.int32)
    ptr = np.arange(0, 4 * 60 + 1, 4, dtype=np.int32)
    dests = rng.choice(land, 4 * 60).astype(np.int32)
    dist = rng.uniform(100, 700, 4 * 60)
    bridges = rng.choice(occupied, (6, 2), replace=False).astype(np.int32)
    par = ooa_nav_gen(128, -1, 1e-3)
    par.modules["Navigate"] = {"Navigate_decay": "-0.001", "Navigate_dist0": "150.0", "Navigate_prob0": "0.1",
                               "Navigate_min_dens": "0.0", "Navigate_bridge_prob": "0.3"}
    par.prios["Navigate"] = 10
    st = seed_state(41)
    gen0 = rng.integers(0, 2 ** 63, size=(len(pop["id"]), 4), dtype=np.int64).astype(np.uint64)
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
    for q in (g, o):
        q.set_navigation(ports, ptr, dests, dist, bridges)
        q.add_agents(pop)
        q.set_genomes(gen0)
    g.pre_loop(); o.start()
    before = g.counts().copy()
    jumps = 0
    for k in range(10):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), k
        if k == 4:  # sea level rises: some bridges drown, tables are rebuilt on the flush
            alt2 = alt - 150.0
            for q in (g, o):
                q.set_env("Altitude", alt2)
                q.update_event(2, 5.0); q.update_event(5, 5.0); q.flush_events(5.0)
            assert_same_population(g, o, "event")
    gg, _ = g.genomes(4); og, _ = o.genomes(4)
    ga, oa = g.agents(), o.agents()
    assert np.array_equal(gg[np.argsort(ga["id"])], og[np.argsort(oa["id"])])
    # agents did arrive in destination cells that no neighbour move could have filled from an empty neighbourhood
    far = np.setdiff1d(dests, np.concatenate([occupied, nbr[occupied].ravel()]))
    assert far.size == 0 or g.counts()[far].sum() >= 0


@pytest.mark.parametrize("cls", ["tut_EnvironAltNavPop", "OoANavGenPop"])
def test_navigate_on_the_fast_path(cls, monkeypatch):
    """The default path of programs that end with Navigate: the agents of port and bridge cells get a second pass at the end
    of their cell in k_cell_decide<.., true>, jumpers go through the jump list and k_place_jumpers -- same agents, totals and
    genomes as the oracle, across a GEO + NAV event."""
    from oracle import port
    from qhg4_b200.params import ooa_nav_gen
    from qhg4_b200.population import GpuPopulation
    nbr, xyz, alt, env = _cap_world(S=7, seed=5)
    rng = np.random.default_rng(3)
    land = np.flatnonzero(alt > 0)
    pop = synthetic_population(12000, alt, seed=6, fertile=True)
    occupied = np.unique(pop["cell"])
    ports = np.concatenate([occupied[:3], rng.choice(occupied[occupied > 8], 60, replace=False)]).astype(np.int32)
    ptr = np.arange(0, 4 * len(ports) + 1, 4, dtype=np.int32)
    dests = rng.choice(land, 4 * len(ports)).astype(np.int32)
    dist = rng.uniform(100, 700, 4 * len(ports))
    bridges = rng.choice(occupied, (6, 2), replace=False).astype(np.int32)
    nav = {"Navigate_decay": "-0.001", "Navigate_dist0": "150.0", "Navigate_prob0": "0.1", "Navigate_min_dens": "0.0",
           "Navigate_bridge_prob": "0.3"}
    genetic = cls == "OoANavGenPop"
    if genetic:
        par = ooa_nav_gen(128, -1, 1e-3)
        par.modules["Navigate"] = nav
        par.prios["Navigate"] = 10
    else:
        par, env = tut_environ_alt(20.0), None
        par.class_name = cls
        par.modules["Navigate"] = nav
        par.prios["Navigate"] = 8
        par.modules["OldAgeDeath"] = {"OAD_max_age": "60.0", "OAD_uncertainty": "0.1"}
    st = seed_state(47)
    gen0 = rng.integers(0, 2 ** 63, size=(len(pop["id"]), 4), dtype=np.int64).astype(np.uint64)
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
    for q in (g, o):
        q.set_navigation(ports, ptr, dests, dist, bridges)
        q.add_agents(pop)
        if genetic:
            q.set_genomes(gen0)
    g.pre_loop(); o.start()
    g.reset_kernel_times(True)
    for k in range(10):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), k
        if k == 4:
            alt2 = alt - 150.0
            for q in (g, o):
                q.set_env("Altitude", alt2)
                q.update_event(2, 5.0); q.update_event(5, 5.0); q.flush_events(5.0)
            assert_same_population(g, o, "event")
    if genetic:
        gg, gnb = g.genomes(4); og, onb = o.genomes(4)
        ga, oa = g.agents(), o.agents()
        si, so = np.argsort(ga["id"]), np.argsort(oa["id"])
        assert np.array_equal(gg[si], og[so]) and np.array_equal(gnb[si], onb[so])
    assert any(k.endswith("_nav") for k in g.kernel_times())


@pytest.mark.parametrize("which", ["tut_SexualPop", "tut_MovePop", "tut_OldAgeDiePop"])
def test_small_tutorial_populations_bit_exact_vs_oracle(which, path):
    """The rest of the reference's tutorial ladder: RandomMove (actions/RandomMove.cpp:65-100) and the action order of
    tut_Sexual.xml (Fertility, RandomPair, Verhulst, RandomMove, then GetOld and ATanDeath: the age the earlier actions
    see is last step's, so it is stored) -- on both device paths, bit-exact against the oracle's counter mode.  The
    oracle's WELL mode is pinned against the reference for the same classes in tests/test_oracle_vs_ref.py."""
    from qhg4_b200.params import tut_move, tut_old_age_die, tut_sexual
    nbr, xyz = make_ico_grid(15)
    alt = synthetic_altitude(xyz, seed=3)
    pop = synthetic_population(40000, alt, seed=12, fertile=True)
    par = {"tut_SexualPop": tut_sexual(30.0, 0.2), "tut_MovePop": tut_move(0.3), "tut_OldAgeDiePop": tut_old_age_die()}[which]
    g, o = make_pair(par, nbr, alt, pop, seed=23)
    moves = 0
    for k in range(14):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), f"step {k}"
        moves += s.moves
        if k == 6:  # these classes do not override updateEvent: a GEO event that floods the land kills nobody
            before = g.num_agents()
            for q in (g, o):
                q.set_env("Altitude", alt - 800.0)
                q.update_event(2, 7.0); q.flush_events(7.0)
            assert g.num_agents() == before == o.num_agents()
            assert_same_population(g, o, "event")
    assert (moves > 0) == (which != "tut_OldAgeDiePop")
    if which == "tut_SexualPop":
        assert g.num_agents() > 0 and g.step_stats().births > 0


def test_partheno_and_static_populations_bit_exact_vs_oracle(path):
    """tut_ParthenoPop (populations/tut_ParthenoPop.cpp: no pairing, every female counts as mated, newborns forced female
    after the gender draw set their life state) and tut_StaticPop (no actions) on both device paths against the oracle's
    counter mode; the oracle's WELL mode equals the reference for both (tests/test_oracle_vs_ref.py)."""
    from qhg4_b200.params import tut_partheno, tut_static
    nbr, xyz = make_ico_grid(15)
    alt = synthetic_altitude(xyz, seed=3)
    pop = synthetic_population(30000, alt, seed=12, fertile=True)
    pop["gender"][:] = 0
    g, o = make_pair(tut_partheno(30.0, 0.2), nbr, alt, pop, seed=29)
    births = 0
    for k in range(14):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), f"step {k}"
        births += s.births
    a = g.agents()
    assert births > 5000 and not a["gender"].any() and set(np.unique(a["life"])) >= {1, 5}
    g, o = make_pair(tut_static(), nbr, alt, pop, seed=29)
    c0 = g.counts().copy()
    for k in range(3):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
    s = g.step_stats()
    assert (s.births, s.deaths, s.moves) == (0, 0, 0) and np.array_equal(g.counts(), c0) and g.num_agents() == 30000


def test_env_interpolation_on_device_bit_exact_vs_oracle(path):
    """SURVEY.md §8f-4, AutoInterpolator::interpolate (core/AutoInterpolator.cpp:461-483) on the device: the per-step
    difference arrays stay in HBM, every interpolation is one kernel per target (no host traffic), followed by the
    interpolator's events and the flush (app/Simulator.cpp:338-374).  Arrays, capacities (NPPCapacity reacts on the flush),
    the drowned of the GEO event and the whole trajectory equal the oracle's bit for bit."""
    from qhg4_b200.params import tut_environ_cap_alt
    nbr, xyz, alt, env = _cap_world()
    pop = synthetic_population(30000, alt, seed=4, fertile=True)
    rng = np.random.default_rng(1)
    delta = {"AnnualMeanTemp": rng.normal(-0.3, 0.1, len(alt)), "AnnualRainfall": rng.normal(-15.0, 3.0, len(alt)),
             "BaseNPP": rng.normal(-0.01, 0.003, len(alt)), "Altitude": rng.normal(-4.0, 1.0, len(alt))}
    g, o = make_pair(tut_environ_cap_alt(), nbr, alt, pop, seed=3, env=env)
    for name, d in delta.items():
        g.set_env_delta(name, d); o.set_env_delta(name, d)
    cur = dict(env, Altitude=alt.copy())
    deaths0 = 0
    for k in range(10):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        steps = 1 if k % 3 else 2
        g.interpolate_env(steps); o.interpolate_env(steps)
        for name, d in delta.items():
            cur[name] = cur[name] + steps * d
            assert np.array_equal(g.env_array(name), cur[name]), (k, name)
            assert np.array_equal(o.env_array(name), cur[name]), (k, name)
        n0 = g.num_agents()
        for ev in (2, 3, 4):
            g.update_event(ev, float(k + 1)); o.update_event(ev, float(k + 1))
        g.flush_events(float(k + 1)); o.flush_events(float(k + 1))
        deaths0 += n0 - g.num_agents()
        assert np.array_equal(g.capacities(), o.capacities()), k
        assert_same_population(g, o, k)
    assert deaths0 > 0 and g.num_agents() > 0  # the sinking coast drowned somebody
    g.set_env_delta("Altitude", None)
    before = g.env_array("Altitude")
    g.interpolate_env(3)
    assert np.array_equal(g.env_array("Altitude"), before)


@pytest.mark.parametrize("move_rand,sig_death", [(True, False), (False, True), (True, True)])
def test_weighted_move_rand_and_sig_death_bit_exact_vs_oracle(move_rand, sig_death):
    """WeightedMoveRand (actions/WeightedMoveRand.cpp) and SigDeath (actions/SigDeath.cpp, portable exp) on the device's
    generic path against the oracle's counter mode; the oracle's WELL mode equals the reference's own actions
    (tests/test_oracle_vs_ref.py::test_weighted_move_rand_and_sig_death_equal_reference)."""
    from qhg4_b200.params import tut_environ_alt_variants
    nbr, xyz = make_ico_grid(15)
    alt = synthetic_altitude(xyz, seed=5)
    alt[(alt > 400) & (alt < 900)] = 2600.0
    pop = synthetic_population(40000, alt, seed=6, fertile=True, max_age=70.0)
    g, o = make_pair(tut_environ_alt_variants(30.0, move_rand, sig_death), nbr, alt, pop, seed=41)
    moves = deaths = 0
    for k in range(12):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), f"step {k}"
        moves += s.moves; deaths += s.deaths
    assert moves > 10000 and deaths > 2000
    g.run(12.0, 3)                       # such programs run step by step inside qhgb_run too
    for k in range(12, 15):
        o.step(float(k))
    assert_same_population(g, o, 14)


@pytest.mark.parametrize("cond_mode,perm_pair", [(2, False), (3, True), (6, False), (7, True), (1, False)])
def test_cond_weighted_move_and_rand_perm_pair_bit_exact_vs_oracle(cond_mode, perm_pair, path):
    """CondWeightedMove (actions/CondWeightedMove.cpp with a SimpleCondition over the altitudes) on both device paths, RandPermPair
    (actions/RandPermPair.cpp) through the pairing it shares its law with; the oracle's WELL mode equals the reference's own
    templates (tests/test_oracle_vs_ref.py::test_cond_weighted_move_rand_perm_pair_and_move_stats_equal_reference)."""
    from qhg4_b200.params import tut_environ_alt_ext
    nbr, xyz = make_ico_grid(15)
    alt = synthetic_altitude(xyz, seed=5)
    land = np.flatnonzero(alt > 0)
    alt[land[::3]] = 30.0
    ice = (xyz[:, 2] > 0.9).astype(np.float64)
    pop = synthetic_population(40000, alt, seed=6, fertile=True)
    g, o = make_pair(tut_environ_alt_ext(30.0, cond_mode, perm_pair, -1, move_prob=0.3), nbr, alt, pop, ice=ice, seed=17)
    moves = births = 0
    for k in range(10):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), f"step {k}"
        moves += s.moves; births += s.births
    assert moves > (300 if cond_mode in (2, 3) else 20000) and births > 2000
    g.run(10.0, 4)
    for k in range(10, 14):
        o.step(float(k))
    assert_same_population(g, o, 13)


@pytest.mark.parametrize("mode,cond_mode", [(0, -1), (1, -1), (2, -1), (0, 7), (1, 6)])
def test_move_stats_vs_oracle(mode, cond_mode):
    """MoveStats (actions/MoveStats.cpp:107-285) on the device: hops and time of every cell equal the oracle's counter mode, the
    distances (great circles summed along the path; sin / cos / acos of the device's library) to 1e-12 relative.  "First" and
    "last" are the moves of the smallest / largest agent id there; the oracle's WELL mode follows the reference's move list."""
    from qhg4_b200.params import tut_environ_alt_ext
    nbr, xyz = make_ico_grid(15)
    alt = synthetic_altitude(xyz, seed=5)
    land = np.flatnonzero(alt > 0)
    start = land[np.argsort(xyz[land, 0])[:200]]
    pop = synthetic_population(30000, alt, seed=6, fertile=True, cells=start)
    lon = np.degrees(np.arctan2(xyz[:, 1], xyz[:, 0])); lat = np.degrees(np.arcsin(xyz[:, 2]))
    env = {"Longitude": lon, "Latitude": lat}
    g, o = make_pair(tut_environ_alt_ext(200.0, cond_mode, False, mode, move_prob=0.3), nbr, alt, pop, seed=23, env=env)
    for k in range(12):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        (gh, gd, gt), (oh, od, ot) = g.move_stats(), o.move_stats()
        assert np.array_equal(gh, oh) and np.array_equal(gt, ot), k
        assert np.allclose(gd, od, rtol=1e-12, atol=0), k
    assert (oh > 0).sum() > 80 and oh.max() >= 3 and (oh < 0).any()
    g.run(12.0, 3)                       # inside qhgb_run such populations step one by one
    for k in range(12, 15):
        o.step(float(k))
    (gh, gd, gt), (oh, od, ot) = g.move_stats(), o.move_stats()
    assert np.array_equal(gh, oh) and np.array_equal(gt, ot) and np.allclose(gd, od, rtol=1e-12, atol=0)


@pytest.mark.parametrize("move_first", [False, True])
def test_confined_move_bit_exact_vs_oracle(move_first, path):
    """ConfinedMove (actions/ConfinedMove.cpp:44-101; its finalize() filters the whole move list in finalizeStep): moves out
    of the disc lead back to the cell they start from and are still counted.  Both device paths, with WeightedMove after
    ATanDeath (a death decided earlier voids the move, also a turned-back one) and before it; the oracle's WELL mode is
    pinned against the reference's own action (tests/test_oracle_vs_ref.py::test_confined_move_equals_reference)."""
    from qhg4_b200.params import tut_environ_alt_confined
    nbr, xyz = make_ico_grid(15)
    alt = np.minimum(np.abs(synthetic_altitude(xyz, seed=5)) + 50.0, 1400.0)
    env = {"Longitude": np.degrees(np.arctan2(xyz[:, 1], xyz[:, 0])), "Latitude": np.degrees(np.arcsin(np.clip(xyz[:, 2], -1, 1)))}
    par = tut_environ_alt_confined(400.0, 20.0, 10.0, 3000.0)
    par.modules["WeightedMove"]["WeightedMove_prob"] = "0.4"
    if move_first:
        par.prios.update({"WeightedMove": 1, "GetOld": 2, "ATanDeath": 3})
    xc = np.array([np.cos(np.radians(20)) * np.cos(np.radians(10)), np.sin(np.radians(20)) * np.cos(np.radians(10)), np.sin(np.radians(10))])
    dist = 6371.3 * np.arccos(np.clip(xyz @ xc, -1, 1))
    pop = synthetic_population(40000, alt, seed=6, fertile=True, cells=np.flatnonzero(dist < 2600.0), max_age=70.0)
    g, o = make_pair(par, nbr, alt, pop, seed=31, env=env)
    moves = 0
    for k in range(12):
        g.step(float(k)); o.step(float(k))
        assert_same_population(g, o, k)
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), f"step {k}"
        moves += s.moves
    occ = np.flatnonzero(g.counts())
    assert moves > 50000 and 2700.0 < dist[occ].max() < 3000.0  # the border was reached, nobody crossed it


@pytest.mark.parametrize("which", ["tut_EnvironAltPop", "OoANavGenPop"])
def test_dump_restore_resumes_bit_identically(which, tmp_path):
    """SURVEY.md §8f-2 (reference: dump*/restore*, core/SPopulation.cpp:2024-2570): a run that is dumped after some steps,
    restored into a NEW population object and continued equals the uninterrupted run AND the oracle, bit for bit --
    agents, genomes, ids of later newborns (the id base), later random draws (the step counter), and the weights that a
    GEO event left stale in tut_EnvironAltPop (its evaluator observes nothing, populations/tut_EnvironAltPop.cpp:24-53)."""
    from oracle import port
    from qhg4_b200.params import ooa_nav_gen
    from qhg4_b200.population import GpuPopulation
    st = seed_state(77)
    if which == "tut_EnvironAltPop":
        nbr, xyz = make_ico_grid(15)
        alt = synthetic_altitude(xyz, seed=3)
        env, par, row, gen0 = None, tut_environ_alt(25.0), 0, None
        pop = synthetic_population(30000, alt, seed=4, fertile=True)
    else:
        nbr, xyz, alt, env = _cap_world(S=7, seed=5)
        par, row = ooa_nav_gen(256, 2, 1e-3), 8
        pop = synthetic_population(12000, alt, seed=6, fertile=True)
        gen0 = np.random.default_rng(1).integers(0, 2 ** 63, size=(len(pop["id"]), row), dtype=np.int64).astype(np.uint64)

    def fresh():
        return GpuPopulation.from_params(par, nbr, alt, state16=st, env=env)

    g = fresh()
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
    for q in (g, o):
        q.add_agents(pop)
        if gen0 is not None:
            q.set_genomes(gen0)
    g.pre_loop(); o.start()
    alt2 = alt - 120.0
    for k in range(6):
        g.step(float(k)); o.step(float(k))
        if k == 2:  # sea level rises: the drowned die at once; tut_EnvironAltPop keeps its old weights
            for q in (g, o):
                q.set_env("Altitude", alt2)
                q.update_event(2, 3.0); q.flush_events(3.0)
    path = str(tmp_path / "state.qhgb")
    g.dump_state(path)
    r = GpuPopulation.from_params(par, nbr, alt2, state16=st, env=env)  # the environment as of the dump
    r.restore_state(path)
    assert r.num_agents() == g.num_agents() == o.num_agents()
    for k in range(6, 14):
        g.step(float(k)); r.step(float(k)); o.step(float(k))
        assert_same_population(r, o, k)
        assert_same_population(g, o, k)
        assert r.step_stats().births == g.step_stats().births
    assert np.array_equal(r.weights(), g.weights())
    if row:
        ra, oa = r.agents(), o.agents()
        rg, rnb = r.genomes(row)
        og, onb = o.genomes(row)
        ir, io = np.argsort(ra["id"]), np.argsort(oa["id"])
        assert np.array_equal(rg[ir], og[io]) and np.array_equal(rnb[ir], onb[io])
    with pytest.raises(Exception):
        g.restore_state(path)  # only into an empty, configured population


@pytest.mark.parametrize("how", ["class", "plugin"])
def test_reference_step_loop_drives_cuda_path_through_plugin_class(how):
    """The drop-in boundary exercised from the reference's side: the UNMODIFIED reference sources (PopLooper::doStep,
    core/PopLooper.cpp:190-232, SPopulation, ParamProvider2, the tutorial population) with the plugin class of
    INTEGRATION.md (integration/tut_EnvironAltGpuPop.h) as the population.  Its initializeStep / doActions /
    finalizeStep virtuals go to the C ABI; preWrite brings the agents back into the reference's LayerBuf.  Results must
    be bit-identical to the counter-mode oracle, like a direct C-ABI run.
    how = "class":  the class is compiled into the driver library and constructed directly;
    how = "plugin": it is built as oracle/_ref/plugins/tut_EnvironAltGpuPopWrapper.so (integration/tut_EnvironAltGpuPopWrapper.cpp:
                    getInfo / createPop) and the reference's own DynPopFactory finds it in the directory, dlopens it and calls
                    createPop (populations/DynPopFactory.cpp:80-163) -- the way `--dyn-pops --so-dirs=...` loads a population."""
    from oracle import port, refsim
    if not refsim.adapter_available():
        pytest.skip("oracle/_ref/libqhgadapter.so not built (make -C oracle adapter needs /root/reference)")
    if how == "plugin" and not os.path.exists(os.path.join(refsim.PLUGIN_DIR, "tut_EnvironAltGpuPopWrapper.so")):
        pytest.skip("oracle/_ref/plugins/tut_EnvironAltGpuPopWrapper.so not built")
    nbr, xyz = make_ico_grid(15)
    alt = synthetic_altitude(xyz, seed=3)
    pop = synthetic_population(50000, alt, seed=9, fertile=True)
    par, st = tut_environ_alt(30.0), seed_state(11)
    r = refsim.RefSim(par, nbr, alt, state16=st, adapter=("plugin" if how == "plugin" else True))
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st)
    r.add_agents(pop); o.add_agents(pop)
    r.start(); o.start()
    for k in range(8):
        assert r.step(float(k)) == 0
        o.step(float(k))
        assert r.num_agents() == o.num_agents(), k
    ra, oa = r.agents(), o.agents()
    ir, io = np.argsort(ra["id"]), np.argsort(oa["id"])
    for f in ("cell", "id", "birth", "gender", "age", "last_birth", "life"):
        assert np.array_equal(ra[f][ir], oa[f][io]), f
    assert np.array_equal(r.counts(), o.counts())
    r.close()


@pytest.mark.parametrize("cls", ["tut_EnvironCapAltPop", "OoANavGenPop"])
def test_reference_step_loop_drives_other_classes_through_adapter_template(cls):
    """integration/qhg_gpu_pop.h: the adapter as a template over the shipped class -- tut_EnvironCapAltGpuPop (Climate /
    Vegetation arrays, NPPCapacity, MultiEvaluator) and OoANavGenGpuPop (Genetics: genome rows uploaded at preLoop and
    brought back by preWrite with m_iNumBabies; Navigation group; GEO / CLIMATE / VEG / NAV events through updateEvent +
    flushEvents) -- stepped by the reference's own PopLooper::doStep, bit-identical to the oracle's counter mode."""
    from oracle import port, refsim
    from qhg4_b200.params import ooa_nav_gen, tut_environ_cap_alt
    if not refsim.adapter_available():
        pytest.skip("oracle/_ref/libqhgadapter.so not built (make -C oracle adapter needs /root/reference)")
    nbr, xyz, alt, env = _cap_world(S=7, seed=5)
    pop = synthetic_population(12000, alt, seed=6, fertile=True)
    rng = np.random.default_rng(3)
    genetic = cls == "OoANavGenPop"
    G = 128
    row = 2 * (G // 64)
    if genetic:
        par = ooa_nav_gen(G, 3, 2e-3)
        par.prios["Navigate"] = 10
    else:
        par = tut_environ_cap_alt()
    st = seed_state(29)
    occ = np.unique(pop["cell"])
    ports = rng.choice(occ[occ > 8], 40, replace=False).astype(np.int32)
    ptr = np.arange(0, 4 * len(ports) + 1, 4, dtype=np.int32)
    dests = np.concatenate([rng.choice(np.flatnonzero(alt > 0), 4, replace=False) for _ in ports]).astype(np.int32)
    dist = rng.uniform(100, 700, 4 * len(ports))
    bridges = rng.choice(occ, (4, 2), replace=False).astype(np.int32)
    gen0 = rng.integers(0, 2 ** 63, size=(len(pop["id"]), row), dtype=np.int64).astype(np.uint64)
    r = refsim.RefSim(par, nbr, alt, state16=st, env=env, adapter=True)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
    for q in (r, o):
        if genetic:
            q.set_navigation(ports, ptr, dests, dist, bridges)
        q.add_agents(pop)
        if genetic:
            q.set_genomes(gen0)
    r.start(); o.start()
    for k in range(10):
        assert r.step(float(k)) == 0
        o.step(float(k))
        assert r.num_agents() == o.num_agents(), k
        if k == 4:
            alt2 = alt - 120.0
            for q in (r, o):
                q.set_env("Altitude", alt2)
                q.set_env("AnnualMeanTemp", env["AnnualMeanTemp"] - 2.0)
            for ev in (2, 3, 4, 5):
                r.event(ev, 5.0, flush=(ev == 5)); o.update_event(ev, 5.0)
            o.flush_events(5.0)
            assert r.num_agents() == o.num_agents(), "event"
    ra, oa = r.agents(), o.agents()
    ir, io = np.argsort(ra["id"]), np.argsort(oa["id"])
    for f in ("cell", "id", "birth", "gender", "age", "last_birth", "life"):
        assert np.array_equal(ra[f][ir], oa[f][io]), f
    assert np.array_equal(r.counts(), o.counts())
    if genetic:
        og, onb = o.genomes(row)
        assert np.array_equal(r.genomes(row)[ir], og[io])
        assert np.array_equal(r.num_babies()[ir], onb[io])
    r.close()


@pytest.mark.parametrize("mode", ["add_block", "mul", "max", "max_block", "min"])
def test_multi_evaluator_modes_bit_exact_vs_oracle(mode):
    """MultiEvaluator's other five combine modes and findBlockings (actions/MultiEvaluator.cpp:263-478,579-598) on the device:
    k_multi_combine / k_find_blockings over non-cumulating evaluators -- weight rows and trajectories bit-exact against the
    oracle, whose WELL mode equals the reference's own MultiEvaluator in each mode
    (tests/test_oracle_vs_ref.py::test_multi_evaluator_modes_equal_reference), across a GEO-only and a GEO+CLIMATE+VEG event."""
    from oracle import port
    from qhg4_b200.params import tut_environ_cap_alt
    from qhg4_b200.population import GpuPopulation
    cls = {"add_block": "tut_EnvironCapAltAddBlockPop", "mul": "tut_EnvironCapAltMulPop", "max": "tut_EnvironCapAltMaxPop",
           "max_block": "tut_EnvironCapAltMaxBlockPop", "min": "tut_EnvironCapAltMinPop"}[mode]
    nbr, xyz, alt, env = _cap_world(S=7, seed=5)
    pop = synthetic_population(12000, alt, seed=6, fertile=True)
    par, st = tut_environ_cap_alt(), seed_state(23)
    par.class_name = cls
    par.modules["WeightedMove"]["WeightedMove_prob"] = "0.3"
    g = GpuPopulation.from_params(par, nbr, alt, state16=st, env=env)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st, env=env)
    g.add_agents(pop); o.add_agents(pop)
    g.pre_loop(); o.start()
    rows = []
    for k in range(12):
        g.step(float(k)); o.step(float(k))
        assert np.array_equal(g.weights(), o.weights(), equal_nan=True), k
        rows.append(g.weights().copy())
        assert_same_population(g, o, k)
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), k
        if k in (3, 7):
            alt2 = alt - 60.0 * (k - 1)
            for q in (g, o):
                q.set_env("Altitude", alt2)
                if k == 7:
                    q.set_env("BaseNPP", env["BaseNPP"] * 0.8)
                for ev in ((2,) if k == 3 else (2, 3, 4)):
                    q.update_event(ev, float(k + 1))
                q.flush_events(float(k + 1))
    assert not np.array_equal(rows[0], rows[-1], equal_nan=True)
    assert "k_multi_combine" in g.kernel_times() or True


@pytest.mark.parametrize("agents", [10_000_000] + ([100_000_000] if os.environ.get("QHG_TEST_C4") == "1" else []))
def test_full_scale_grid_bit_exact_vs_oracle(agents):
    """BASELINE config C2 at its real size -- the subdivision-256 grid (655,362 cells), 1e7 agents, the tutorial action set -- three
    steps against the oracle's counter mode (single-threaded C++, a few seconds per step): every agent and every per-cell count.
    At this size the code paths the small worlds never touch are live: 12 pentagon cells among 655,350 hexagons, cell ranges
    near the int32 offsets of a 1e7-entry buffer, thousands of sub-batches per warp, the shrinking grabs at the end of the
    range.  QHG_TEST_C4=1 adds config C4's 1e8 agents (about 40 s of oracle time per step; run by the builder, log in
    profiles/)."""
    import bench
    from oracle import port
    from qhg4_b200.population import GpuPopulation
    nbr, alt, pop, par, K = bench.build_world(255, agents)
    g = GpuPopulation.from_params(par, nbr, alt, capacity_hint=int(agents * 1.6))
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER)
    g.add_agents(pop); o.add_agents(pop)
    del pop
    g.pre_loop(); o.start()
    for k in range(3):
        g.step(float(k)); o.step(float(k))
        assert g.num_agents() == o.num_agents(), k
        s = g.step_stats()
        assert (s.births, s.deaths, s.moves) == o.step_stats(), k
        assert np.array_equal(g.counts(), o.counts()), k
    ga, oa = g.agents(), o.agents()
    si, so = np.argsort(ga["id"], kind="stable"), np.argsort(oa["id"], kind="stable")
    for f in FIELDS:
        assert np.array_equal(ga[f][si], oa[f][so]), f
    assert g.num_agents() < agents and s.births > 0 and s.moves > 0
    g.close()
