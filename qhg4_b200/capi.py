"""ctypes binding of the C ABI in include/qhg_b200.h (qhg4_b200/libqhg_b200.so).

This is the same surface a cgo/JNI/C++ host would bind; Python is only the
test and benchmark driver.  There is no fallback: if the library is missing or no CUDA device
is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqhg_b200.so")

# every symbol include/qhg_b200.h declares: name -> (restype, argtypes)
vp, cp, i32, i64, f32, f64, u32 = C.c_void_p, C.c_char_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint


class StepStats(C.Structure):
    _fields_ = [("num_agents", i64), ("births", i64), ("deaths", i64), ("moves", i64), ("next_id", i64), ("steps_done", i64)]


SYMBOLS = {
    "qhgb_create": (i32, [cp, i32, i32, i32, i64, C.POINTER(vp)]),
    "qhgb_destroy": (i32, [vp]),
    "qhgb_last_error": (cp, []),
    "qhgb_version": (cp, []),
    "qhgb_set_cells": (i32, [vp, vp, vp]),
    "qhgb_set_env_array": (i32, [vp, cp, vp, i64]),
    "qhgb_set_env_delta": (i32, [vp, cp, vp, i64]),
    "qhgb_interpolate_env": (i32, [vp, i32]),
    "qhgb_get_env_array": (i32, [vp, cp, vp]),
    "qhgb_set_navigation": (i32, [vp, i32, vp, vp, vp, vp, i32, vp]),
    "qhgb_set_attribute": (i32, [vp, cp, f64]),
    "qhgb_set_attribute_str": (i32, [vp, cp, cp]),
    "qhgb_set_prio": (i32, [vp, cp, i32]),
    "qhgb_enable_action": (i32, [vp, cp, i32]),
    "qhgb_set_seed": (i32, [vp, vp]),
    "qhgb_add_agents": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, vp]),
    "qhgb_get_agents": (i64, [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "qhgb_set_genomes": (i32, [vp, i64, vp]),
    "qhgb_get_genomes": (i64, [vp, i64, vp, vp]),
    "qhgb_pre_loop": (i32, [vp]),
    "qhgb_initialize_step": (i32, [vp, f32]),
    "qhgb_do_actions": (i32, [vp, u32, f32]),
    "qhgb_finalize_step": (i32, [vp]),
    "qhgb_step": (i32, [vp, f32]),
    "qhgb_run": (i32, [vp, f32, i32]),
    "qhgb_get_run_totals": (i32, [vp, vp, vp, vp]),
    "qhgb_synchronize": (i32, [vp]),
    "qhgb_update_event": (i32, [vp, i32, f32]),
    "qhgb_flush_events": (i32, [vp, f32]),
    "qhgb_get_num_agents_effective": (i64, [vp]),
    "qhgb_get_num_agents_array": (i32, [vp, vp]),
    "qhgb_get_num_agents_range": (i32, [vp, i32, i32, vp]),
    "qhgb_mirror_num_agents_array": (i32, [vp, vp, i32, i32]),
    "qhgb_get_move_stats": (i32, [vp, vp, vp, vp]),
    "qhgb_get_path_counts": (i32, [vp, vp, vp, vp]),
    "qhgb_get_occupied": (i32, [vp, i32, vp, vp]),
    "qhgb_get_step_stats": (i32, [vp, C.POINTER(StepStats)]),
    "qhgb_get_env_weights": (i32, [vp, vp]),
    "qhgb_get_birth_death_probs": (i32, [vp, vp, vp]),
    "qhgb_atan_death_prob": (i32, [vp, i32, vp, vp]),
    "qhgb_get_capacities": (i32, [vp, vp]),
    "qhgb_dump_state": (i32, [vp, cp]),
    "qhgb_restore_state": (i32, [vp, cp]),
    "qhgb_comm_get_unique_id": (i32, [vp, i32]),
    "qhgb_comm_init": (i32, [vp, i32, i32, vp, vp]),
    "qhgb_comm_get_traffic": (i32, [vp, vp, vp]),
    "qhgb_comm_p2p_handle": (i32, [vp, vp, i32]),
    "qhgb_comm_p2p_connect": (i32, [vp, vp]),
    "qhgb_host_alloc": (vp, [C.c_size_t]),
    "qhgb_host_free": (i32, [vp]),
    "qhgb_get_launch_count": (i64, [vp]),
    "qhgb_get_stream": (vp, [vp]),
    "qhgb_get_kernel_times": (i32, [vp, i32, vp, vp, vp]),
    "qhgb_reset_kernel_times": (i32, [vp, i32]),
    "qhgb_event_record": (i32, [vp, i32]),
    "qhgb_event_elapsed_ms": (f64, [vp, i32, i32]),
}

_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m qhg4_b200.build` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            if os.environ.get("QHG_AB_OLD_LIB") and not hasattr(L, name):
                continue  # profiles/try_libs.sh only: A/B runs of the bench over libraries built from older commits
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class QhgError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = load().qhgb_last_error().decode(errors="replace")
        raise QhgError(f"{what} -> {rc}: {msg}")
    return rc
