"""ctypes wrapper of oracle/liboracle.so (our CPU restatement, qhg_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, `__graft_entry__.smoke()` and bench.py's
cpu_baseline leg, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
MODE_WELL, MODE_COUNTER = 0, 1
_lib = None


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("qhg_oracle.cpp", "qhg_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f32, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint
        L.qor_create.restype = vp
        L.qor_create.argtypes = [C.c_char_p, i32, i32, i32]
        L.qor_destroy.argtypes = [vp]
        L.qor_set_cells.argtypes = [vp, vp, vp]
        L.qor_set_env_array.argtypes = [vp, C.c_char_p, vp, i64]
        L.qor_set_env_delta.argtypes = [vp, C.c_char_p, vp, i64]
        L.qor_interpolate_env.argtypes = [vp, i32]
        L.qor_get_env_array.argtypes = [vp, C.c_char_p, vp]
        L.qor_set_attribute.argtypes = [vp, C.c_char_p, C.c_double]
        L.qor_set_attribute_str.argtypes = [vp, C.c_char_p, C.c_char_p]
        L.qor_set_prio.argtypes = [vp, C.c_char_p, i32]
        L.qor_enable_action.argtypes = [vp, C.c_char_p, i32]
        L.qor_set_seed.argtypes = [vp, vp]
        L.qor_add_agents.argtypes = [vp, i64] + [vp] * 7
        L.qor_pre_loop.argtypes = [vp]
        L.qor_initialize_step.argtypes = [vp, f32]
        L.qor_do_actions.argtypes = [vp, u32, f32]
        L.qor_finalize_step.argtypes = [vp]
        L.qor_step.argtypes = [vp, f32]
        L.qor_update_event.argtypes = [vp, i32, f32]
        L.qor_flush_events.argtypes = [vp, f32]
        L.qor_get_num_agents_effective.restype = i64
        L.qor_get_num_agents_effective.argtypes = [vp]
        L.qor_get_num_agents_array.argtypes = [vp, vp]
        L.qor_get_agents.restype = i64
        L.qor_get_agents.argtypes = [vp, i64] + [vp] * 9
        L.qor_get_env_weights.argtypes = [vp, vp]
        L.qor_get_birth_death_probs.argtypes = [vp, vp, vp]
        L.qor_atan_death_prob.argtypes = [vp, i32, vp, vp]
        L.qor_get_capacities.argtypes = [vp, vp]
        L.qor_get_move_stats.argtypes = [vp, vp, vp, vp]
        L.qor_set_navigation.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp]
        L.qor_set_genomes.argtypes = [vp, i64, vp]
        L.qor_get_genomes.restype = i64
        L.qor_get_genomes.argtypes = [vp, i64, vp, vp]
        L.qor_get_step_stats.argtypes = [vp, vp, vp, vp]
        L.qor_set_genetics_well.argtypes = [vp, vp, u32]
        L.qor_gene2_crossover.argtypes = [vp, vp, i32, i32, vp]
        L.qor_gene2_freereco.argtypes = [vp, vp, i32, vp]
        L.qor_gene2_mutate.argtypes = [vp, vp, i32, i32]
        L.qor_get_pending_births.restype = i64
        L.qor_get_pending_births.argtypes = [vp]
        L.qor_set_birth_id_offset.argtypes = [vp, i64, i64]
        L.qor_extract_foreign.restype = i64
        L.qor_extract_foreign.argtypes = [vp, i32, i32, i64] + [vp] * 7
        L.qor_recount.argtypes = [vp]
        L.qor_get_max_id.restype = i64
        L.qor_get_max_id.argtypes = [vp]
        L.qor_set_max_id.argtypes = [vp, i64]
        L.qor_philox4x32_10.argtypes = [vp, vp, vp]
        L.qor_well_sequence.argtypes = [vp, i32, vp]
        L.qor_polyline_eval.argtypes = [C.c_char_p, i32, vp, vp, i32]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, np.uint32)
    k = np.ascontiguousarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().qor_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def well_sequence(state16, n):
    st = np.ascontiguousarray(state16, np.uint32)
    out = np.zeros(n, np.uint32)
    lib().qor_well_sequence(_p(st), n, _p(out))
    return out


def polyline_eval(defn, x, float_cast=True):
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros_like(x)
    if lib().qor_polyline_eval(defn.encode(), len(x), _p(x), _p(out), int(float_cast)) != 0:
        raise ValueError(defn)
    return out


class OraclePop:
    """Same call sequence as the product's `GpuPopulation` (and the reference's PopBase)."""

    def __init__(self, params, nbr, altitude, ice=None, mode=MODE_COUNTER, state16=None, env=None):
        from qhg4_b200.params import DEFAULT_STATE
        L = lib()
        nbr = np.ascontiguousarray(nbr, np.int32)
        self.ncells, self.max_neigh = nbr.shape
        self.h = L.qor_create(params.class_name.encode(), self.ncells, self.max_neigh, mode)
        if not self.h:
            raise RuntimeError(f"unknown population class {params.class_name}")
        L.qor_set_cells(self.h, _p(nbr), None)
        self.set_env("Altitude", altitude)
        if ice is not None:
            self.set_env("Ice", ice)
        for k, v in (env or {}).items():
            self.set_env(k, v)
        for mod, pars in params.modules.items():
            for k, v in pars.items():
                if L.qor_set_attribute_str(self.h, k.encode(), str(v).encode()) != 0:
                    raise ValueError(f"attribute {k}={v}")
        for name, pr in params.prios.items():
            if L.qor_set_prio(self.h, name.encode(), int(pr)) != 0:
                raise ValueError(f"prio for unknown action {name}")
        st = np.ascontiguousarray(DEFAULT_STATE if state16 is None else state16, np.uint32)
        L.qor_set_seed(self.h, _p(st))

    def set_env(self, name, v):
        v = np.ascontiguousarray(v, np.float64)
        assert lib().qor_set_env_array(self.h, name.encode(), _p(v), len(v)) == 0

    def set_env_delta(self, name, delta):
        d = None if delta is None else np.ascontiguousarray(delta, np.float64)
        assert lib().qor_set_env_delta(self.h, name.encode(), _p(d), 0 if d is None else len(d)) == 0, name

    def interpolate_env(self, steps=1):
        assert lib().qor_interpolate_env(self.h, int(steps)) == 0

    def env_array(self, name):
        out = np.zeros(self.ncells)
        assert lib().qor_get_env_array(self.h, name.encode(), _p(out)) == 0, name
        return out

    def add_agents(self, pop):
        n = len(pop["cell"])
        arrs = [np.ascontiguousarray(pop["cell"], np.int32), np.ascontiguousarray(pop["id"], np.int64),
                np.ascontiguousarray(pop["birth"], np.float32), np.ascontiguousarray(pop["gender"], np.uint8),
                np.ascontiguousarray(pop["age"], np.float32), np.ascontiguousarray(pop["last_birth"], np.float32),
                np.ascontiguousarray(pop["life"], np.uint32)]
        assert lib().qor_add_agents(self.h, n, *[_p(a) for a in arrs]) == 0

    def start(self):
        assert lib().qor_pre_loop(self.h) == 0

    def step(self, t):
        return lib().qor_step(self.h, float(t))

    def initialize_step(self, t):
        return lib().qor_initialize_step(self.h, float(t))

    def do_actions(self, prio, t):
        return lib().qor_do_actions(self.h, int(prio), float(t))

    def finalize_step(self):
        return lib().qor_finalize_step(self.h)

    def update_event(self, ev, t=0.0):
        return lib().qor_update_event(self.h, int(ev), float(t))

    def flush_events(self, t=0.0):
        return lib().qor_flush_events(self.h, float(t))

    def capacities(self):
        out = np.zeros(self.ncells)
        assert lib().qor_get_capacities(self.h, _p(out)) == 0
        return out

    def move_stats(self):
        """MoveStats' per-cell arrays (hops, dist, time)"""
        h, d, t = np.zeros(self.ncells, np.int32), np.zeros(self.ncells), np.zeros(self.ncells)
        assert lib().qor_get_move_stats(self.h, _p(h), _p(d), _p(t)) == 0
        return h, d, t

    def set_navigation(self, port_cell, port_ptr, dest_cell, dist, bridges=()):
        pc, pp = np.ascontiguousarray(port_cell, np.int32), np.ascontiguousarray(port_ptr, np.int32)
        dc, dd = np.ascontiguousarray(dest_cell, np.int32), np.ascontiguousarray(dist, np.float64)
        br = np.ascontiguousarray(np.asarray(bridges, np.int32).reshape(-1, 2))
        assert lib().qor_set_navigation(self.h, len(pc), _p(pc), _p(pp), _p(dc), _p(dd), len(br), _p(br) if len(br) else None) == 0

    def set_genetics_well(self, state16, index):
        st = np.ascontiguousarray(state16, np.uint32)
        assert lib().qor_set_genetics_well(self.h, _p(st), int(index)) == 0

    def set_genomes(self, genomes):
        g = np.ascontiguousarray(genomes, np.uint64)
        assert lib().qor_set_genomes(self.h, g.shape[0], _p(g)) == 0

    def genomes(self, row_words):
        n = self.num_agents()
        g = np.zeros((n, row_words), np.uint64)
        nb = np.zeros(n, np.int32)
        assert lib().qor_get_genomes(self.h, n, _p(g), _p(nb)) == n
        return g, nb

    def enable_action(self, name, on=True):
        return lib().qor_enable_action(self.h, name.encode(), int(on))

    def set_attribute(self, name, v):
        return lib().qor_set_attribute(self.h, name.encode(), float(v))

    def num_agents(self):
        return int(lib().qor_get_num_agents_effective(self.h))

    def agents(self):
        n = self.num_agents()
        out = dict(cell=np.zeros(n, np.int32), id=np.zeros(n, np.int64), birth=np.zeros(n, np.float32),
                   gender=np.zeros(n, np.uint8), age=np.zeros(n, np.float32), last_birth=np.zeros(n, np.float32),
                   life=np.zeros(n, np.uint32), mate_id=np.zeros(n, np.int64), slot=np.zeros(n, np.int32))
        k = lib().qor_get_agents(self.h, n, *[_p(out[f]) for f in
                                              ("cell", "id", "birth", "gender", "age", "last_birth", "life", "mate_id", "slot")])
        assert k == n, (k, n)
        return out

    def counts(self):
        out = np.zeros(self.ncells, np.uint64)
        lib().qor_get_num_agents_array(self.h, _p(out))
        return out

    def weights(self):
        out = np.zeros((self.ncells, self.max_neigh + 1))
        lib().qor_get_env_weights(self.h, _p(out))
        return out

    def bd(self):
        b, d = np.zeros(self.ncells), np.zeros(self.ncells)
        lib().qor_get_birth_death_probs(self.h, _p(b), _p(d))
        return b, d

    def atan_prob(self, age):
        age = np.ascontiguousarray(age, np.float32)
        p = np.zeros(len(age))
        lib().qor_atan_death_prob(self.h, len(age), _p(age), _p(p))
        return p

    def step_stats(self):
        b, d, m = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        lib().qor_get_step_stats(self.h, C.byref(b), C.byref(d), C.byref(m))
        return b.value, d.value, m.value

    # ---- sharded runs ---------------------------------------------------------------------
    def pending_births(self):
        return int(lib().qor_get_pending_births(self.h))

    def set_birth_id_offset(self, offset, total):
        lib().qor_set_birth_id_offset(self.h, int(offset), int(total))

    def extract_foreign(self, c0, c1):
        cap = self.num_agents()
        out = dict(cell=np.zeros(cap, np.int32), id=np.zeros(cap, np.int64), birth=np.zeros(cap, np.float32),
                   gender=np.zeros(cap, np.uint8), age=np.zeros(cap, np.float32), last_birth=np.zeros(cap, np.float32),
                   life=np.zeros(cap, np.uint32))
        k = lib().qor_extract_foreign(self.h, int(c0), int(c1), cap, *[_p(out[f]) for f in
                                                                          ("cell", "id", "birth", "gender", "age", "last_birth", "life")])
        return {f: v[:k] for f, v in out.items()}

    def recount(self):
        lib().qor_recount(self.h)

    def max_id(self):
        return int(lib().qor_get_max_id(self.h))

    def set_max_id(self, v):
        lib().qor_set_max_id(self.h, int(v))

    def close(self):
        if self.h:
            lib().qor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
