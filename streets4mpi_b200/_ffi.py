"""ctypes binding of libs4mpi.so (include/s4mpi.h).

The library is the product path: if it is missing, or there is no CUDA device,
everything here raises -- there is deliberately no CPU fallback (the CPU
restatement under ``oracle/`` is test infrastructure and is never imported from
this package).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S4MPI_LIB") or os.path.join(_HERE, "libs4mpi.so")    # override: A/B kernel builds

ROOT_ORIGIN, ROOT_GOAL, ROOT_AUTO = 0, 1, 2


class S4Error(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libs4mpi error %d: %s" % (code, message))
        self.code = code


class Physics(C.Structure):
    """settings.py:36-44 as the library sees them."""
    _fields_ = [("car_length", C.c_double), ("min_breaking_distance", C.c_double),
                ("braking_deceleration", C.c_double), ("trip_volume", C.c_uint32), ("reserved", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "days", "sssp_instances", "arcs_algorithmic", "nodes_algorithmic", "arcs_relaxed", "nodes_popped",
        "stale_pops", "queue_pushes", "bucket_advances", "relax_rounds", "fallback_sweeps", "trips_routed",
        "trips_unreachable", "trips_stuck", "hops", "kernel_launches", "sssp_launches")] + [
        (n, C.c_double) for n in ("ms_weights", "ms_sssp", "ms_walk", "ms_adapt", "ms_step_total")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/s4mpi.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_i32p, _u32p, _f64p, _u8p = (C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_double),
                             C.POINTER(C.c_uint8))
SIGNATURES = {
    "s4_abi_version": (C.c_int, []),
    "s4_last_error": (C.c_char_p, []),
    "s4_ctx_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "s4_ctx_destroy": (C.c_int, [_vp]),
    "s4_ctx_sync": (C.c_int, [_vp]),
    "s4_ctx_set_option": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "s4_ctx_get_option": (C.c_int, [_vp, C.c_char_p, _f64p]),
    "s4_ctx_device_info": (C.c_int, [_vp, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "s4_comm_unique_id": (C.c_int, [_vp]),
    "s4_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "s4_comm_free": (C.c_int, [_vp]),
    "s4_comm_allreduce_host": (C.c_int, [_vp, _u32p, C.c_int64]),
    "s4_comm_size": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "s4_graph_upload": (C.c_int, [_vp, C.c_int32, C.c_int64, _i32p, _i32p, _f64p, _i32p, _i32p, C.POINTER(_vp)]),
    "s4_graph_free": (C.c_int, [_vp]),
    "s4_graph_dims": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "s4_graph_group_log": (C.c_int, [_vp, _u32p, C.c_int64, C.POINTER(C.c_int64)]),
    "s4_graph_sssp": (C.c_int, [_vp, _f64p, C.c_int32, _f64p, _i32p]),
    "s4_sim_create": (C.c_int, [_vp, C.c_int64, _i32p, _i32p, C.c_double, C.POINTER(Physics), C.POINTER(_vp)]),
    "s4_sim_free": (C.c_int, [_vp]),
    "s4_sim_create_sampled": (C.c_int, [_vp, C.c_int64, C.c_int64, _i32p, C.c_int64, _i32p, C.c_uint64, C.c_double,
                                        C.POINTER(Physics), C.POINTER(_vp)]),
    "s4_sim_get_trips": (C.c_int, [_vp, _i32p, _i32p]),
    "s4_sim_roots": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "s4_sim_step": (C.c_int, [_vp]),
    "s4_sim_weights": (C.c_int, [_vp, C.c_int, _f64p]),
    "s4_sim_get_load": (C.c_int, [_vp, _u32p]),
    "s4_sim_set_load": (C.c_int, [_vp, _u32p]),
    "s4_sim_load_device_ptr": (_vp, [_vp]),
    "s4_sim_load_add": (C.c_int, [_vp, _vp]),
    "s4_sim_load_copy": (C.c_int, [_vp, _vp]),
    "s4_sim_commit_total": (C.c_int, [_vp]),
    "s4_sim_exchange": (C.c_int, [_vp]),
    "s4_sim_load_hash": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "s4_sim_get_cumulative": (C.c_int, [_vp, _u32p, C.POINTER(C.c_int)]),
    "s4_sim_set_cumulative": (C.c_int, [_vp, _u32p]),
    "s4_sim_road_construction": (C.c_int, [_vp, _u8p]),
    "s4_sim_get_max_speed": (C.c_int, [_vp, _i32p, _i32p]),
    "s4_sim_set_max_speed": (C.c_int, [_vp, _i32p, _i32p]),
    "s4_sim_get_counters": (C.c_int, [_vp, C.POINTER(Counters)]),
    "s4_sim_reset_counters": (C.c_int, [_vp]),
    "s4_sim_trips_stuck": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "s4_sim_get_root_slots": (C.c_int, [_vp, C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "s4_driving_speed": (C.c_int, [_vp, C.c_int64, _f64p, _i32p, _u32p, C.POINTER(Physics), _f64p]),
}

_lib = None


def load():
    """Load libs4mpi.so (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libs4mpi.so is not built (%s); run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` -- there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the header and the library drifted
            fn.restype = res
            fn.argtypes = args
        if lib.s4_abi_version() != 1:
            raise ImportError("libs4mpi ABI mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise S4Error(rc, (load().s4_last_error() or b"").decode("utf-8", "replace"))


def ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def as_array(x, dtype):
    return np.ascontiguousarray(x, dtype=dtype)
