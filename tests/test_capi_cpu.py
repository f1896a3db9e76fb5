"""The C-ABI library without a GPU: it loads, exports every symbol include/qhg_b200.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device.  Host-side helpers (parameter files, grids)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from qhg4_b200 import capi
from qhg4_b200.icogrid import make_ico_grid, make_torus_grid, num_cells, synthetic_altitude, synthetic_population
from qhg4_b200.params import PopParams, tut_environ_alt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qhg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qhgb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(capi.LIB_PATH):
        from qhg4_b200 import build
        build.build()
    lib = C.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/qhg_b200.h but not exported"
    assert set(names) == set(capi.SYMBOLS), set(names) ^ set(capi.SYMBOLS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = capi.load()
    h = C.c_void_p()
    rc = L.qhgb_create(b"tut_EnvironAltPop", 0, 42, 6, 0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in L.qhgb_last_error()


def test_null_arguments_return_minus_one():
    L = capi.load()
    assert L.qhgb_pre_loop(None) == -1
    assert L.qhgb_step(None, 0.0) == -1
    assert L.qhgb_get_num_agents_effective(None) == -1
    assert L.qhgb_version().startswith(b"qhg4_b200")


def test_param_xml_roundtrip():
    p = tut_environ_alt(20.0)
    q = PopParams.from_xml(p.to_xml())
    assert q.class_name == "tut_EnvironAltPop" and q.prios == p.prios and q.modules == p.modules
    assert q.prios["SingleEvaluator[Alt]"] == 4 and "AltCapPref" in q.modules["SingleEvaluator[Alt]"]


@pytest.mark.parametrize("S", [1, 3, 15])
def test_ico_grid_topology(S):
    nbr, xyz = make_ico_grid(S)
    n = num_cells(S)
    assert nbr.shape == (n, 6) and n == 10 * (S + 1) ** 2 + 2
    deg = (nbr >= 0).sum(1)
    assert (deg == 5).sum() == 12 and (deg == 6).sum() == n - 12
    assert np.all(deg[:12] == 5)                      # the 12 icosahedron vertices come first
    for c in range(n):                                 # ascending, -1 padded, symmetric
        row = nbr[c][nbr[c] >= 0]
        assert np.all(np.diff(row) > 0) and np.all(nbr[c][len(row):] == -1)
        assert all(c in nbr[m] for m in row)
    assert np.allclose(np.linalg.norm(xyz, axis=1), 1.0)


def test_torus_and_population_helpers():
    nbr = make_torus_grid(8, 8)
    assert nbr.shape == (64, 6) and np.all(np.diff(nbr, axis=1) > 0)
    _, xyz = make_ico_grid(7)
    alt = synthetic_altitude(xyz)
    assert 0.25 < (alt < 0).mean() < 0.35
    pop = synthetic_population(5000, alt, seed=1, fertile=True)
    assert np.all(np.diff(pop["cell"]) >= 0) and np.all(alt[pop["cell"]] > 0)
    assert set(np.unique(pop["life"])) <= {1, 5} and len(np.unique(pop["id"])) == 5000
