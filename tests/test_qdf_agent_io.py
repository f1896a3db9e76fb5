"""QDF agent I/O: the agent dataset of a population file, through the REFERENCE's own writer and reader.

The host keeps QDF / HDF5 I/O (BASELINE.json north_star); what this project owes it is the agent table in the reference's compound
layout at the moment PopWriter asks for it (PopBase::preWrite + writeAgentDataQDF, io/PopWriter.cpp:84-118) and taking over the
agents PopReader hands in (readAgentDataQDF, io/PopReader.cpp:143-170).  HDF5 is not installed in this environment: the dozen
H5S / H5T / H5D calls of that path are backed by memory (oracle/stubs/hdf5_stubs.cpp), everything around them is the reference's
unmodified code (core/SPopulation.cpp:1356-1372 compound type, :1465-1568 writeAgentDataQDFSafe, :1689-1741 readAgentDataQDF).
"""
import numpy as np
import pytest

from oracle import port, refsim
from qhg4_b200.icogrid import make_ico_grid, synthetic_altitude, synthetic_population
from qhg4_b200.params import seed_state, tut_environ_alt

pytestmark = pytest.mark.skipif(not refsim.available() or not hasattr(refsim.lib(), "qref_qdf_write_agents"),
                                reason="oracle/_ref was not built with the QDF agent-dataset driver")

FIELDS = ("cell", "id", "birth", "gender", "age", "last_birth", "life")
# dataset member -> field of the agent tables used everywhere else in the tests
MEMBERS = {"LifeState": "life", "CellIdx": "cell", "AgentID": "id", "BirthTime": "birth", "Gender": "gender", "Age": "age",
           "LastBirth": "last_birth"}


def world():
    nbr, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz, seed=5)
    pop = synthetic_population(6000, alt, seed=6, fertile=True)
    return nbr, alt, pop


def as_table(rec):
    names = rec.dtype.names
    assert all(m in names for m in MEMBERS), names
    return {f: np.asarray(rec[m]) for m, f in MEMBERS.items()}


def by_id(t):
    o = np.argsort(t["id"], kind="stable")
    return {k: v[o] for k, v in t.items()}


def test_reference_writes_and_reads_its_agent_dataset():
    """the stand-in HDF5 is good enough for the reference itself: after steps with births and deaths the written dataset holds
    exactly the live agents (dead slots squeezed out layer by layer), in the reference's compound layout, and a fresh population
    that reads it back holds the same agents"""
    nbr, alt, pop = world()
    par, st = tut_environ_alt(25.0), seed_state(3)
    r = refsim.RefSim(par, nbr, alt, threads=1, state16=st, layer_size=1024)   # several layers, holes in most of them
    r.add_agents(pop); r.start()
    for k in range(6):
        r.step(float(k))
    rec = r.qdf_write_agents(6.0)
    assert rec.dtype.names[:6] == ("LifeState", "CellIdx", "CellID", "AgentID", "BirthTime", "Gender")  # core/SPopulation.cpp:1363-1368
    live = r.agents()
    assert len(rec) == r.num_agents() == len(live["id"]) and len(rec) != len(pop["id"])
    a, b = by_id(as_table(rec)), by_id({f: live[f] for f in FIELDS})
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), f
    r2 = refsim.RefSim(par, nbr, alt, threads=1, state16=st, layer_size=1024)
    r2.qdf_read_agents(rec)
    r2.start()
    got = by_id({f: r2.agents()[f] for f in FIELDS})
    for f in FIELDS:
        assert np.array_equal(got[f], b[f]), f
    assert np.array_equal(r2.counts(), r.counts())
    r.close(); r2.close()


def test_reference_writes_the_agent_dataset_of_ooa_nav_gen_pop():
    """the same for the class of BASELINE configurations #3 / #5: OoANavGenPop registers the same eight members (m_iNumBabies is
    not part of its dataset, populations/OoANavGenPop.cpp addPopSpecificAgentDataTypeQDF) and writes exactly its live agents"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import golden_cases as gc
    if not refsim.has_class("OoANavGenPop"):
        pytest.skip("OoANavGenPop is not part of this build of oracle/_ref")
    d = gc.build_inputs("ooa_nav_gen")
    par = gc.CASES["ooa_nav_gen"][0]()
    env = {k[4:]: d[k] for k in d if k.startswith("env_")}
    pop = {k[4:]: d[k] for k in d if k.startswith("pop_")}
    r = refsim.RefSim(par, d["nbr"], d["alt"], threads=1, state16=d["seed_state"], env=env)
    r.set_navigation(d["nav_ports"], d["nav_ptr"], d["nav_dests"], d["nav_dist"], d["nav_bridges"])
    r.add_agents(pop); r.set_genomes(d["genomes"]); r.start()
    for k in range(5):
        r.step(float(k))
    rec = r.qdf_write_agents(5.0)
    assert rec.dtype.names == ("LifeState", "CellIdx", "CellID", "AgentID", "BirthTime", "Gender", "Age", "LastBirth")
    live = r.agents()
    assert len(rec) == r.num_agents() and len(rec) != len(pop["id"])
    a, b = by_id(as_table(rec)), by_id({f: live[f] for f in FIELDS})
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), f
    r.close()


@pytest.mark.gpu
def test_gpu_population_goes_through_the_reference_qdf_writer_and_reader():
    """a population that lives on the GPU (the adapter class of INTEGRATION.md, driven by the reference's PopLooper): PopWriter's
    sequence -- preWrite, writeAgentDataQDF -- hands HDF5 exactly the oracle's agents in the reference's compound layout; a second
    GPU population reads that dataset through PopReader's sequence, and carries on bit-identically to the oracle"""
    if not refsim.adapter_available():
        pytest.skip("oracle/_ref/libqhgadapter.so was not built")
    nbr, alt, pop = world()
    par, st = tut_environ_alt(25.0), seed_state(3)
    g = refsim.RefSim(par, nbr, alt, threads=1, state16=st, adapter=True, layer_size=1024)
    o = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st)
    g.add_agents(pop); o.add_agents(pop)
    g.start(); o.start()
    for k in range(6):
        g.step(float(k)); o.step(float(k))
    rec = g.qdf_write_agents(6.0)
    assert len(rec) == o.num_agents()
    a, b = by_id(as_table(rec)), by_id({f: o.agents()[f] for f in FIELDS})
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), f
    g2 = refsim.RefSim(par, nbr, alt, threads=1, state16=st, adapter=True, layer_size=1024)
    g2.qdf_read_agents(rec)
    g2.start()
    # (the generator state of a resumed run is the step counter: core of qhgb_dump_state; here a fresh run with the same agents)
    o2 = port.OraclePop(par, nbr, alt, mode=port.MODE_COUNTER, state16=st)
    o2.add_agents({f: b[f] for f in FIELDS})
    o2.start()
    for k in range(4):
        g2.step(float(k)); o2.step(float(k))
        ga, oa = by_id({f: g2.agents()[f] for f in FIELDS}), by_id({f: o2.agents()[f] for f in FIELDS})
        for f in FIELDS:
            assert np.array_equal(ga[f], oa[f]), (k, f)
    g.close(); g2.close()
