"""ctypes wrapper of oracle/_ref/libqhgref.so (the unmodified reference step loop).

TEST INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's cpu_baseline /
`--impl reference` leg, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libqhgref.so")
# the same driver and reference objects plus the plugin class of INTEGRATION.md (integration/tut_EnvironAltGpuPop.h),
# linked against qhg4_b200/libqhg_b200.so: the reference's PopLooper::doStep drives the CUDA path (`make adapter`)
ADAPTER_PATH = os.path.join(_HERE, "_ref", "libqhgadapter.so")
PLUGIN_DIR = os.path.join(_HERE, "_ref", "plugins")  # *Wrapper.so files for the reference's DynPopFactory (`make adapter`)

_lib = None
_libs = {}


def available() -> bool:
    return os.path.exists(LIB_PATH)


def adapter_available() -> bool:
    return os.path.exists(ADAPTER_PATH)


def has_class(name: str) -> bool:
    """is this population class part of the reference build in oracle/_ref?"""
    if not available():
        return False
    L = lib()
    if not hasattr(L, "qref_has_class"):
        return name.startswith("tut_")
    L.qref_has_class.argtypes = [C.c_char_p]
    return bool(L.qref_has_class(name.encode()))


def lib(adapter: bool = False):
    global _lib
    if adapter not in _libs:
        # the adapter library is opened RTLD_GLOBAL: a plugin that DynPopFactory dlopens resolves the application's symbols
        # from the loading process (the reference's executables are linked with --export-dynamic, app/Makefile:132,147)
        L = C.CDLL(ADAPTER_PATH, mode=C.RTLD_GLOBAL) if adapter else C.CDLL(LIB_PATH)
        L.qref_create.restype = C.c_void_p
        L.qref_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.qref_add_agents.argtypes = [C.c_void_p, C.c_long] + [C.c_void_p] * 7
        L.qref_set_navigation.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.qref_start.argtypes = [C.c_void_p]
        L.qref_step.argtypes = [C.c_void_p, C.c_float]
        L.qref_run.restype = C.c_double
        L.qref_run.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.qref_num_agents.restype = C.c_long
        L.qref_num_agents.argtypes = [C.c_void_p]
        L.qref_get_agents.restype = C.c_long
        L.qref_get_agents.argtypes = [C.c_void_p, C.c_long] + [C.c_void_p] * 9
        L.qref_get_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.qref_get_weights.argtypes = [C.c_void_p, C.c_void_p]
        L.qref_get_bd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.qref_atan_prob.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.qref_geo_event.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        L.qref_timers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.qref_get_capacities.argtypes = [C.c_void_p, C.c_void_p]
        if hasattr(L, "qref_get_move_stats"):
            L.qref_get_move_stats.argtypes = [C.c_void_p] * 4
        if hasattr(L, "qref_qdf_write_agents"):
            L.qref_qdf_write_agents.restype = C.c_longlong
            L.qref_qdf_write_agents.argtypes = [C.c_void_p, C.c_float]
            L.qref_qdf_dataset.restype = C.c_long
            L.qref_qdf_dataset.argtypes = [C.c_longlong, C.c_void_p, C.c_void_p]
            L.qref_qdf_agent_member.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
            L.qref_qdf_read_agents.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.qref_set_env.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.qref_event.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
        L.qref_set_genomes.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_void_p]
        L.qref_get_genomes.restype = C.c_long
        L.qref_get_genomes.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.qref_genetics_well.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.qref_genetics_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        if hasattr(L, "qref_get_num_babies"):
            L.qref_get_num_babies.restype = C.c_long
            L.qref_get_num_babies.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
            L.qref_modify_attribute.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        for f in ("qref_gene2_crossover", "qref_gene2_freereco", "qref_gene2_mutate"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p] if f == "qref_gene2_crossover" else \
                ([C.c_void_p, C.c_void_p, C.c_int, C.c_void_p] if f == "qref_gene2_freereco" else [C.c_void_p, C.c_void_p, C.c_int, C.c_int])
        L.qref_destroy.argtypes = [C.c_void_p]
        L.qref_well_sequence.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.qref_polyline_eval.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _libs[adapter] = L
    return _libs[adapter]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def well_sequence(state16, n):
    st = np.ascontiguousarray(state16, dtype=np.uint32)
    out = np.zeros(n, dtype=np.uint32)
    lib().qref_well_sequence(_p(st), n, _p(out))
    return out


def polyline_eval(defn: str, x, float_cast=True):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    rc = lib().qref_polyline_eval(defn.encode(), len(x), _p(x), _p(out), int(float_cast))
    if rc != 0:
        raise ValueError(defn)
    return out


class RefSim:
    """One `tut_EnvironAltPop` on a given grid, driven through PopLooper::doStep."""

    def __init__(self, params, nbr, altitude, ice=None, threads=1, state16=None, quiet=True, layer_size=65536, env=None,
                 adapter=False):
        from qhg4_b200.params import DEFAULT_STATE
        self.ncells = len(nbr)
        self.adapter = bool(adapter)
        self.L = lib(self.adapter)
        self._nbr = np.ascontiguousarray(nbr, dtype=np.int32)
        alt = np.ascontiguousarray(altitude, dtype=np.float64)
        icea = None if ice is None else np.ascontiguousarray(ice, dtype=np.uint8)
        st = np.ascontiguousarray(DEFAULT_STATE if state16 is None else state16, dtype=np.uint32)
        gen = params.modules.get("Genetics")
        if gen is not None:  # the reference cannot take Genetics attributes from XML (see geneticsInit in oracle/ref_driver.cpp)
            params = params.copy()
            del params.modules["Genetics"]
        with tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False) as f:
            f.write(params.to_xml())
            path = f.name
        try:
            cls = params.class_name
            if adapter == "plugin":   # the reference's DynPopFactory loads integration/tut_EnvironAltGpuPopWrapper.cpp's .so
                os.environ["QHG_REF_SO_DIR"] = PLUGIN_DIR
                cls = "dyn:tut_EnvironAltGpuPop"
            elif self.adapter:        # integration/*GpuPop.h compiled into the driver library
                cls = {"tut_EnvironAltPop": "tut_EnvironAltGpuPop", "tut_EnvironCapAltPop": "tut_EnvironCapAltGpuPop",
                       "OoANavGenPop": "OoANavGenGpuPop"}[params.class_name]
            self.h = self.L.qref_create(path.encode(), cls.encode(), self.ncells, _p(self._nbr),
                                       _p(alt), _p(icea), int(threads), _p(st), int(layer_size), int(quiet))
        finally:
            os.unlink(path)
        if not self.h:
            raise RuntimeError("qref_create failed")
        self.threads = threads
        if gen is not None:
            rc = self.L.qref_genetics_init(self.h, int(gen["Genetics_genome_size"]), int(gen["Genetics_num_crossover"]),
                                           float(gen["Genetics_mutation_rate"]))
            if rc != 0:
                raise RuntimeError("Genetics::init failed")
        for k, v in (env or {}).items():
            self.set_env(k, v)

    def set_env(self, name, v):
        v = np.ascontiguousarray(v, np.float64)
        assert len(v) == self.ncells
        assert self.L.qref_set_env(self.h, name.encode(), _p(v)) == 0, name

    def set_navigation(self, port_cell, port_ptr, dest_cell, dist, bridges=()):
        pc, pp = np.ascontiguousarray(port_cell, np.int32), np.ascontiguousarray(port_ptr, np.int32)
        dc, dd = np.ascontiguousarray(dest_cell, np.int32), np.ascontiguousarray(dist, np.float64)
        br = np.ascontiguousarray(np.asarray(bridges, np.int32).reshape(-1, 2))
        assert self.L.qref_set_navigation(self.h, len(pc), _p(pc), _p(pp), _p(dc), _p(dd), len(br), _p(br) if len(br) else None) == 0

    def event(self, event_id, t=0.0, flush=True):
        return self.L.qref_event(self.h, int(event_id), float(t), int(flush))

    def capacities(self):
        out = np.zeros(self.ncells)
        assert self.L.qref_get_capacities(self.h, _p(out)) == 0
        return out

    # ---- the agent dataset of a QDF file through the reference's own writer / reader (HDF5 backed by memory, oracle/stubs) ----
    _QDF_TYPES = {101: "i1", 102: "u1", 103: "<i2", 104: "<u2", 105: "<i4", 106: "<u4", 107: "<i8", 108: "<u8", 109: "<i8", 110: "<u8",
                  111: "<f4", 112: "<f8", 113: "<i4", 114: "<u4", 115: "<i8", 116: "<u8", 118: "u1", 119: "i1", 120: "u1", 122: "<i2", 123: "<u2"}

    def qdf_agent_dtype(self, itemsize):
        """numpy dtype of the compound type the population registers for its agents (names and offsets as given to H5Tinsert)"""
        names, formats, offsets = [], [], []
        i = 0
        while True:
            name = C.create_string_buffer(128)
            off, code = C.c_size_t(0), C.c_int(0)
            if self.L.qref_qdf_agent_member(self.h, i, name, 128, C.byref(off), C.byref(code)) != 0:
                break
            names.append(name.value.decode()); offsets.append(off.value); formats.append(self._QDF_TYPES[code.value])
            i += 1
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": itemsize})

    def qdf_write_agents(self, t=0.0):
        """preWrite + the agent part of PopWriter::write; returns the records HDF5 was handed, as a structured array"""
        ds = self.L.qref_qdf_write_agents(self.h, float(t))
        if ds < 0:
            raise RuntimeError(f"writeAgentDataQDF failed ({ds})")
        elem = C.c_size_t(0)
        n = self.L.qref_qdf_dataset(ds, C.byref(elem), None)
        raw = np.zeros(n * elem.value, np.uint8)
        self.L.qref_qdf_dataset(ds, C.byref(elem), _p(raw))
        return raw.view(self.qdf_agent_dtype(elem.value))

    def qdf_read_agents(self, records):
        """the agent part of PopReader::read on a dataset holding `records`"""
        raw = np.ascontiguousarray(records).view(np.uint8)
        rc = self.L.qref_qdf_read_agents(self.h, len(records), _p(raw))
        if rc != 0:
            raise RuntimeError(f"readAgentDataQDF failed ({rc})")

    def move_stats(self):
        """MoveStats' per-cell arrays (m_aiHops, m_adDist, m_adTime)"""
        h, d, t = np.zeros(self.ncells, np.int32), np.zeros(self.ncells), np.zeros(self.ncells)
        assert self.L.qref_get_move_stats(self.h, _p(h), _p(d), _p(t)) == 0
        return h, d, t

    def add_agents(self, pop: dict):
        n = len(pop["cell"])
        arrs = [np.ascontiguousarray(pop["cell"], np.int32), np.ascontiguousarray(pop["id"], np.int64),
                np.ascontiguousarray(pop["birth"], np.float32), np.ascontiguousarray(pop["gender"], np.uint8),
                np.ascontiguousarray(pop["age"], np.float32), np.ascontiguousarray(pop["last_birth"], np.float32),
                np.ascontiguousarray(pop["life"], np.uint32)]
        rc = self.L.qref_add_agents(self.h, n, *[_p(a) for a in arrs])
        assert rc == 0

    def set_genomes(self, genomes, first_slot=0):
        """genome rows of the agents in slots first_slot.. (Genetics probe populations; agents are added from slot 0 on)"""
        g = np.ascontiguousarray(genomes, np.uint64)
        w = self.L.qref_set_genomes(self.h, int(first_slot), len(g), _p(g))
        assert w == g.shape[1], (w, g.shape)

    def genomes(self, row_words):
        n = self.num_agents()
        g = np.zeros((n, row_words), np.uint64)
        assert self.L.qref_get_genomes(self.h, n, _p(g)) == row_words
        return g

    def num_babies(self):
        """m_iNumBabies of every live agent in the order of agents() (populations/OoANavGenPop.cpp:243)"""
        n = self.num_agents()
        out = np.zeros(n, np.int32)
        assert self.L.qref_get_num_babies(self.h, n, _p(out)) == n
        return out

    def modify_attribute(self, name: str, value: float):
        """PopBase::modifyAttributes(name, value)"""
        rc = self.L.qref_modify_attribute(self.h, name.encode(), float(value))
        if rc != 0:
            raise RuntimeError(f"modifyAttributes({name}) -> {rc}")

    def genetics_well(self):
        st, idx = np.zeros(16, np.uint32), np.zeros(1, np.uint32)
        assert self.L.qref_genetics_well(self.h, _p(st), _p(idx)) == 0
        return st, int(idx[0])

    def start(self):
        rc = self.L.qref_start(self.h)
        if rc != 0:
            raise RuntimeError(f"qref_start -> {rc}")

    def step(self, t: float):
        return self.L.qref_step(self.h, float(t))

    def run(self, t0: float, nsteps: int):
        n = C.c_int64(0)
        sec = self.L.qref_run(self.h, float(t0), int(nsteps), C.byref(n))
        return sec, n.value

    def num_agents(self) -> int:
        return int(self.L.qref_num_agents(self.h))

    def agents(self) -> dict:
        n = self.num_agents()
        out = dict(cell=np.zeros(n, np.int32), id=np.zeros(n, np.int64), birth=np.zeros(n, np.float32),
                   gender=np.zeros(n, np.uint8), age=np.zeros(n, np.float32), last_birth=np.zeros(n, np.float32),
                   life=np.zeros(n, np.uint32), mate=np.zeros(n, np.int32), slot=np.zeros(n, np.int32))
        k = self.L.qref_get_agents(self.h, n, *[_p(out[f]) for f in
                                               ("cell", "id", "birth", "gender", "age", "last_birth", "life", "mate", "slot")])
        assert k == n, (k, n)
        return out

    def counts(self):
        out = np.zeros(self.ncells, np.uint64)
        self.L.qref_get_counts(self.h, _p(out))
        return out

    def weights(self):
        out = np.zeros((self.ncells, 7), np.float64)
        self.L.qref_get_weights(self.h, _p(out))
        return out

    def bd(self):
        b = np.zeros(self.ncells)
        d = np.zeros(self.ncells)
        rc = self.L.qref_get_bd(self.h, _p(b), _p(d))
        assert rc == 0
        return b, d

    def atan_prob(self, age):
        age = np.ascontiguousarray(age, np.float32)
        p = np.zeros(len(age))
        self.L.qref_atan_prob(self.h, len(age), _p(age), _p(p))
        return p

    def geo_event(self, altitude=None, ice=None, t=0.0):
        a = None if altitude is None else np.ascontiguousarray(altitude, np.float64)
        i = None if ice is None else np.ascontiguousarray(ice, np.uint8)
        return self.L.qref_geo_event(self.h, _p(a), _p(i), float(t))

    def timers(self):
        a, f = C.c_double(0), C.c_double(0)
        self.L.qref_timers(self.h, C.byref(a), C.byref(f))
        return a.value, f.value

    def close(self):
        if self.h:
            self.L.qref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
