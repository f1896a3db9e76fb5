"""`GpuPopulation`: the host-side mirror of the reference's population interface over the C ABI.

Method names and call order are the reference's `PopBase` / `PopLooper` ones
(core/PopBase.h:16-121, core/PopLooper.cpp:166-202): `read_species_data` (XML parameters and
priorities), `add_agents`, `pre_loop`, then per step `initialize_step(t)`, `do_actions(prio, t)` for
every priority level, `finalize_step()`; `update_event` / `flush_events` on environment events.
Every method goes through `include/qhg_b200.h`; errors raise `QhgError` carrying the text of
`qhgb_last_error()` (the C calls themselves return the reference's 0 / -1).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import QhgError, StepStats, check
from .params import DEFAULT_STATE, PopParams

EVENT_ID_GEO, EVENT_ID_CLIMATE, EVENT_ID_VEG, EVENT_ID_NAV, EVENT_ID_FLUSH = 2, 3, 4, 5, 20


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GpuPopulation:
    def __init__(self, pop_class: str, nbr, device: int = 0, capacity_hint: int = 0, global_id=None):
        self.L = capi.load()
        self._host_blocks = []
        nbr = np.ascontiguousarray(nbr, np.int32)
        self.ncells, self.max_neigh = nbr.shape
        h = C.c_void_p()
        check(self.L.qhgb_create(pop_class.encode(), device, self.ncells, self.max_neigh, int(capacity_hint), C.byref(h)),
              "qhgb_create")
        self.h = h
        gid = None if global_id is None else np.ascontiguousarray(global_id, np.int32)
        check(self.L.qhgb_set_cells(self.h, _p(nbr), _p(gid)), "qhgb_set_cells")
        self.prios = {}
        self.set_seed(DEFAULT_STATE)

    # ---- construction helpers ---------------------------------------------------------------
    @classmethod
    def from_params(cls, params: PopParams, nbr, altitude, ice=None, state16=None, device=0, capacity_hint=0, env=None):
        pop = cls(params.class_name, nbr, device=device, capacity_hint=capacity_hint)
        pop.set_env("Altitude", altitude)
        if ice is not None:
            pop.set_env("Ice", ice)
        for k, v in (env or {}).items():
            pop.set_env(k, v)
        pop.read_species_data(params)
        if state16 is not None:
            pop.set_seed(state16)
        return pop

    def set_env(self, name: str, values):
        v = np.ascontiguousarray(values, np.float64)
        check(self.L.qhgb_set_env_array(self.h, name.encode(), _p(v), len(v)), f"qhgb_set_env_array({name})")

    def set_env_delta(self, name: str, delta):
        """AutoInterpolator's per-step difference array of one target (core/AutoInterpolator.cpp:461-483); None removes it."""
        d = None if delta is None else np.ascontiguousarray(delta, np.float64)
        check(self.L.qhgb_set_env_delta(self.h, name.encode(), _p(d), 0 if d is None else len(d)), f"qhgb_set_env_delta({name})")

    def interpolate_env(self, steps: int = 1):
        """AutoInterpolator::interpolate(iSteps) on the device: every target array += steps * its difference array."""
        check(self.L.qhgb_interpolate_env(self.h, int(steps)), "qhgb_interpolate_env")

    def env_array(self, name: str):
        out = np.zeros(self.ncells, np.float64)
        check(self.L.qhgb_get_env_array(self.h, name.encode(), _p(out)), f"qhgb_get_env_array({name})")
        return out

    def read_species_data(self, params: PopParams):
        """SPopulation::readSpeciesData (core/SPopulation.cpp:1108-1142): priorities, then action attributes."""
        for name, pr in params.prios.items():
            check(self.L.qhgb_set_prio(self.h, name.encode(), int(pr)), f"qhgb_set_prio({name})")
            self.prios[name] = int(pr)
        for mod, pars in sorted(params.modules.items(), key=lambda kv: kv[0] != "Genetics"):  # genome size first
            for k, v in pars.items():
                check(self.L.qhgb_set_attribute_str(self.h, k.encode(), str(v).encode()), f"qhgb_set_attribute_str({k})")

    def modify_attributes(self, name: str, value: float):
        check(self.L.qhgb_set_attribute(self.h, name.encode(), float(value)), f"qhgb_set_attribute({name})")

    def set_prio(self, action: str, prio: int):
        check(self.L.qhgb_set_prio(self.h, action.encode(), int(prio)), f"qhgb_set_prio({action})")
        self.prios[action] = int(prio)

    def enable_action(self, action: str, on: bool = True):
        check(self.L.qhgb_enable_action(self.h, action.encode(), int(on)), f"qhgb_enable_action({action})")

    def disable_action(self, action: str):
        self.enable_action(action, False)

    def set_seed(self, state16):
        st = np.ascontiguousarray(state16, np.uint32)
        assert st.shape == (16,)
        check(self.L.qhgb_set_seed(self.h, _p(st)), "qhgb_set_seed")

    def add_agents(self, pop: dict):
        n = len(pop["cell"])
        arrs = [np.ascontiguousarray(pop["cell"], np.int32), np.ascontiguousarray(pop["id"], np.int64),
                np.ascontiguousarray(pop["birth"], np.float32), np.ascontiguousarray(pop["gender"], np.uint8),
                np.ascontiguousarray(pop["age"], np.float32), np.ascontiguousarray(pop["last_birth"], np.float32),
                np.ascontiguousarray(pop["life"], np.uint32)]
        check(self.L.qhgb_add_agents(self.h, n, *[_p(a) for a in arrs]), "qhgb_add_agents")

    def set_navigation(self, port_cell, port_ptr, dest_cell, dist, bridges=()):
        pc, pp = np.ascontiguousarray(port_cell, np.int32), np.ascontiguousarray(port_ptr, np.int32)
        dc, dd = np.ascontiguousarray(dest_cell, np.int32), np.ascontiguousarray(dist, np.float64)
        br = np.ascontiguousarray(np.asarray(bridges, np.int32).reshape(-1, 2))
        check(self.L.qhgb_set_navigation(self.h, len(pc), _p(pc), _p(pp), _p(dc), _p(dd), len(br), _p(br) if len(br) else None),
              "qhgb_set_navigation")

    def set_genomes(self, genomes):
        g = np.ascontiguousarray(genomes, np.uint64)
        check(self.L.qhgb_set_genomes(self.h, g.shape[0], _p(g)), "qhgb_set_genomes")

    def genomes(self, row_words: int):
        n = self.num_agents()
        g = np.zeros((n, row_words), np.uint64)
        nb = np.zeros(n, np.int32)
        k = self.L.qhgb_get_genomes(self.h, n, _p(g), _p(nb))
        if k != n:
            raise QhgError(f"qhgb_get_genomes -> {k}: {self.L.qhgb_last_error().decode()}")
        return g, nb

    # ---- the loop ---------------------------------------------------------------------------
    def pre_loop(self):
        check(self.L.qhgb_pre_loop(self.h), "qhgb_pre_loop")

    start = pre_loop

    def initialize_step(self, t: float):
        check(self.L.qhgb_initialize_step(self.h, float(t)), "qhgb_initialize_step")

    def do_actions(self, prio: int, t: float):
        check(self.L.qhgb_do_actions(self.h, int(prio), float(t)), "qhgb_do_actions")

    def finalize_step(self):
        check(self.L.qhgb_finalize_step(self.h), "qhgb_finalize_step")

    def step(self, t: float):
        check(self.L.qhgb_step(self.h, float(t)), "qhgb_step")

    def run(self, t0: float, nsteps: int):
        """nsteps steps queued on the device without a host round trip per step (same results as nsteps step() calls)"""
        check(self.L.qhgb_run(self.h, float(t0), int(nsteps)), "qhgb_run")

    def run_totals(self):
        """(agent-steps, agents sent, agents received) summed over all completed steps since pre_loop"""
        a, s, r = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(self.L.qhgb_get_run_totals(self.h, C.byref(a), C.byref(s), C.byref(r)), "qhgb_get_run_totals")
        return a.value, s.value, r.value

    def synchronize(self):
        check(self.L.qhgb_synchronize(self.h), "qhgb_synchronize")

    def update_event(self, event_id: int, t: float = 0.0):
        check(self.L.qhgb_update_event(self.h, int(event_id), float(t)), "qhgb_update_event")

    def flush_events(self, t: float = 0.0):
        check(self.L.qhgb_flush_events(self.h, float(t)), "qhgb_flush_events")

    # ---- read back --------------------------------------------------------------------------
    def num_agents(self) -> int:
        return int(self.L.qhgb_get_num_agents_effective(self.h))

    def counts(self, out=None):
        out = np.zeros(self.ncells, np.uint64) if out is None else out
        check(self.L.qhgb_get_num_agents_array(self.h, _p(out)), "qhgb_get_num_agents_array")
        return out

    def counts_range(self, c0: int, c1: int, out=None):
        """per-cell counts of the cells [c0, c1) only (a shard reads back its own range)"""
        out = np.zeros(c1 - c0, np.uint64) if out is None else out
        check(self.L.qhgb_get_num_agents_range(self.h, int(c0), int(c1), _p(out)), "qhgb_get_num_agents_range")
        return out

    def mirror_counts(self, host=None, c0: int = 0, c1: int = None):
        """keep `host` (ulong per cell of [c0, c1)) current after every step, the way m_aiNumAgentsPerCell is in the reference;
        None ends it"""
        if host is None:
            check(self.L.qhgb_mirror_num_agents_array(self.h, None, 0, 0), "qhgb_mirror_num_agents_array")
            self._mirror = None
            return None
        c1 = self.ncells if c1 is None else c1
        assert host.dtype == np.uint64 and len(host) >= c1 - c0
        check(self.L.qhgb_mirror_num_agents_array(self.h, _p(host), int(c0), int(c1)), "qhgb_mirror_num_agents_array")
        self._mirror = host  # keeps the array alive
        return host

    def path_counts(self):
        """(fast-path, generic-path, recovery-kernel) pipeline runs so far"""
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(self.L.qhgb_get_path_counts(self.h, C.byref(a), C.byref(b), C.byref(c)), "qhgb_get_path_counts")
        return a.value, b.value, c.value

    def move_stats(self):
        """MoveStats' per-cell arrays (hops, dist, time); -1 = never reached"""
        h, d, t = np.zeros(self.ncells, np.int32), np.zeros(self.ncells), np.zeros(self.ncells)
        check(self.L.qhgb_get_move_stats(self.h, _p(h), _p(d), _p(t)), "qhgb_get_move_stats")
        return h, d, t

    def occupied(self, cells):
        """OccTracker::calcBitMap for this population: one byte per listed cell, 1 = somebody is there"""
        c = np.ascontiguousarray(cells, np.int32)
        out = np.zeros(len(c), np.uint8)
        check(self.L.qhgb_get_occupied(self.h, len(c), _p(c), _p(out)), "qhgb_get_occupied")
        return out

    def step_stats(self) -> StepStats:
        s = StepStats()
        check(self.L.qhgb_get_step_stats(self.h, C.byref(s)), "qhgb_get_step_stats")
        return s

    def agents(self) -> dict:
        n = self.num_agents()
        out = dict(cell=np.zeros(n, np.int32), cell_id=np.zeros(n, np.int32), id=np.zeros(n, np.int64),
                   birth=np.zeros(n, np.float32), gender=np.zeros(n, np.uint8), age=np.zeros(n, np.float32),
                   last_birth=np.zeros(n, np.float32), life=np.zeros(n, np.uint32), mate_id=np.zeros(n, np.int64))
        k = self.L.qhgb_get_agents(self.h, n, *[_p(out[f]) for f in
                                                ("cell", "cell_id", "id", "birth", "gender", "age", "last_birth", "life", "mate_id")])
        if k != n:
            raise QhgError(f"qhgb_get_agents -> {k}: {self.L.qhgb_last_error().decode()}")
        return out

    def weights(self):
        out = np.zeros((self.ncells, self.max_neigh + 1))
        check(self.L.qhgb_get_env_weights(self.h, _p(out)), "qhgb_get_env_weights")
        return out

    def bd(self):
        b, d = np.zeros(self.ncells), np.zeros(self.ncells)
        check(self.L.qhgb_get_birth_death_probs(self.h, _p(b), _p(d)), "qhgb_get_birth_death_probs")
        return b, d

    def capacities(self):
        out = np.zeros(self.ncells)
        check(self.L.qhgb_get_capacities(self.h, _p(out)), "qhgb_get_capacities")
        return out

    def atan_prob(self, age):
        age = np.ascontiguousarray(age, np.float32)
        p = np.zeros(len(age))
        check(self.L.qhgb_atan_death_prob(self.h, len(age), _p(age), _p(p)), "qhgb_atan_death_prob")
        return p

    # ---- several GPUs ------------------------------------------------------------------------
    def comm_init(self, rank: int, nranks: int, unique_id: bytes, cell_begin):
        cb = np.ascontiguousarray(cell_begin, np.int32)
        assert len(cb) == nranks + 1 and len(unique_id) == 128
        uid = C.create_string_buffer(unique_id, 128)
        check(self.L.qhgb_comm_init(self.h, int(rank), int(nranks), uid, _p(cb)), "qhgb_comm_init")

    def dump_state(self, path: str):
        check(self.L.qhgb_dump_state(self.h, str(path).encode()), "qhgb_dump_state")

    def restore_state(self, path: str):
        """instead of add_agents + pre_loop, on a population configured like the dumped one"""
        check(self.L.qhgb_restore_state(self.h, str(path).encode()), "qhgb_restore_state")

    def comm_p2p_handle(self) -> bytes:
        buf = C.create_string_buffer(128)
        check(self.L.qhgb_comm_p2p_handle(self.h, buf, 128), "qhgb_comm_p2p_handle")
        return buf.raw

    def comm_p2p_connect(self, all_handles):
        """all_handles: the 128 bytes of every rank in rank order; None switches back to the NCCL exchange"""
        buf = None if all_handles is None else C.create_string_buffer(all_handles, len(all_handles))
        check(self.L.qhgb_comm_p2p_connect(self.h, buf), "qhgb_comm_p2p_connect")

    def comm_traffic(self):
        s, r = C.c_int64(0), C.c_int64(0)
        check(self.L.qhgb_comm_get_traffic(self.h, C.byref(s), C.byref(r)), "qhgb_comm_get_traffic")
        return s.value, r.value

    # ---- measurement ------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.L.qhgb_get_launch_count(self.h))

    def host_array(self, n: int, dtype=np.uint64):
        """numpy array over page-locked host memory (qhgb_host_alloc): the per-step result arrays copy into it by one DMA"""
        dt = np.dtype(dtype)
        ptr = self.L.qhgb_host_alloc(int(n) * dt.itemsize)
        if not ptr:
            raise QhgError(self.L.qhgb_last_error().decode())
        buf = (C.c_char * (int(n) * dt.itemsize)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dt, count=int(n))
        self._host_blocks.append(ptr)
        return arr

    def reset_kernel_times(self, enable=True):
        check(self.L.qhgb_reset_kernel_times(self.h, int(enable)), "qhgb_reset_kernel_times")

    def kernel_times(self) -> dict:
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        calls = (C.c_int64 * cap)()
        n = self.L.qhgb_get_kernel_times(self.h, cap, names, ms, calls)
        return {names[i].decode(): (ms[i], calls[i]) for i in range(min(n, cap))}

    def event_record(self, slot: int):
        check(self.L.qhgb_event_record(self.h, int(slot)), "qhgb_event_record")

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = float(self.L.qhgb_event_elapsed_ms(self.h, int(a), int(b)))
        if ms < 0:
            raise QhgError(self.L.qhgb_last_error().decode())
        return ms

    def close(self):
        if getattr(self, "h", None):
            self.L.qhgb_destroy(self.h)
            self.h = None
            for ptr in self._host_blocks:  # arrays from host_array() must not be used after close()
                self.L.qhgb_host_free(ptr)
            self._host_blocks = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
