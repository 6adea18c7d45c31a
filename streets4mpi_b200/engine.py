"""Object layer over the C-ABI: device context, device-resident street graph and
per-rank simulation state.  Pure plumbing -- every number is produced by the
CUDA kernels in ``csrc/``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi
from ._ffi import Counters, Physics, as_array, check, ptr

_default_device = None


class Device(object):
    """One GPU + one stream (s4_ctx).  ``stream`` is a raw cudaStream_t (int) or
    None for a library-owned stream."""

    def __init__(self, index=None, stream=None):
        if index is None:
            index = int(os.environ.get("LOCAL_RANK", "0"))
        self.lib = _ffi.load()
        self.index = int(index)
        h = C.c_void_p()
        check(self.lib.s4_ctx_create(self.index, C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.stream = int(stream) if stream else None      # None: library-owned stream
        # tuning hook for A/B runs of the kernels: S4MPI_OPTIONS="key=value,key=value" (see s4_ctx_set_option)
        self.env_options = set()                           # keys set from the environment outrank device_settings
        for item in filter(None, os.environ.get("S4MPI_OPTIONS", "").split(",")):
            key, _, value = item.partition("=")
            self.set_option(key.strip(), float(value))
            self.env_options.add(key.strip())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.s4_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.s4_ctx_sync(self._h))

    def set_option(self, key, value):
        check(self.lib.s4_ctx_set_option(self._h, key.encode(), float(value)))

    def get_option(self, key):
        v = C.c_double()
        check(self.lib.s4_ctx_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    # ---- rank-level exchange (NCCL inside the library, streets4mpi.py:44-46)
    def comm_unique_id(self):
        """128-byte NCCL id; rank 0 creates it and ships it to the other ranks."""
        buf = C.create_string_buffer(128)
        check(self.lib.s4_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks, rank, unique_id=None):
        buf = C.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
        check(self.lib.s4_comm_init(self._h, int(nranks), int(rank), buf))

    def comm_free(self):
        check(self.lib.s4_comm_free(self._h))

    def comm_size(self):
        n, r = C.c_int(), C.c_int()
        check(self.lib.s4_comm_size(self._h, C.byref(n), C.byref(r)))
        return n.value, r.value

    def allreduce_host(self, arr_u32):
        """In-place SUM all-reduce of a host uint32 array over the ranks of comm_init."""
        a = as_array(arr_u32, np.uint32)
        check(self.lib.s4_comm_allreduce_host(self._h, ptr(a, C.c_uint32), len(a)))
        return a

    def info(self):
        name = C.create_string_buffer(256)
        sm = C.c_int()
        mem = C.c_uint64()
        check(self.lib.s4_ctx_device_info(self._h, name, 256, C.byref(sm), C.byref(mem)))
        return {"name": name.value.decode(), "sm_count": sm.value, "hbm_bytes": mem.value}

    def driving_speed(self, length, max_speed, trips, physics):
        length = as_array(length, np.float64)
        max_speed = as_array(max_speed, np.int32)
        trips = as_array(trips, np.uint32)
        out = np.empty(len(length), dtype=np.float64)
        check(self.lib.s4_driving_speed(self._h, len(length), ptr(length, C.c_double), ptr(max_speed, C.c_int32),
                                        ptr(trips, C.c_uint32), C.byref(physics), ptr(out, C.c_double)))
        return out


def default_device():
    """Process-wide device (cuda:LOCAL_RANK), created on first use."""
    global _default_device
    if _default_device is None:
        _default_device = Device()
    return _default_device


def set_default_device(dev):
    global _default_device
    _default_device = dev


def make_physics(settings):
    return Physics(float(settings["car_length"]), float(settings["min_breaking_distance"]),
                   float(settings["braking_deceleration"]), int(settings["trip_volume"]), 0)


class DeviceGraph(object):
    """Symmetric CSR of the street network in HBM (s4_graph)."""

    def __init__(self, device, V, u, v, length, max_speed, node_rank=None):
        """node_rank: optional permutation, position of every node in the device's internal numbering."""
        self.device = device
        self.lib = device.lib
        u = as_array(u, np.int32)
        v = as_array(v, np.int32)
        length = as_array(length, np.float64)
        max_speed = as_array(max_speed, np.int32)
        if not (len(u) == len(v) == len(length) == len(max_speed)):
            raise ValueError("street arrays differ in length")
        h = C.c_void_p()
        rank = None if node_rank is None else as_array(node_rank, np.int32)
        if rank is not None and len(rank) != int(V):
            raise ValueError("node_rank must have one entry per node")
        check(self.lib.s4_graph_upload(device._h, int(V), len(u), ptr(u, C.c_int32), ptr(v, C.c_int32),
                                       ptr(length, C.c_double), ptr(max_speed, C.c_int32),
                                       None if rank is None else ptr(rank, C.c_int32), C.byref(h)))
        self._h = h
        self.V, self.S = int(V), len(u)
        a = C.c_int64()
        check(self.lib.s4_graph_dims(h, None, None, C.byref(a)))
        self.A = a.value

    def close(self):
        if getattr(self, "_h", None):
            self.lib.s4_graph_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def group_log(self):
        """Diagnostics of the last day (option group_log=1): uint32[n_groups, 4] = SM cycles >> 10, rounds,
        entries popped, bucket advances per group of trees."""
        n = C.c_int64()
        check(self.lib.s4_graph_group_log(self._h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 0), dtype=np.uint32)
        if n.value:
            check(self.lib.s4_graph_group_log(self._h, ptr(out, C.c_uint32), n.value, C.byref(n)))
        return out.reshape(-1, 4)

    def sssp(self, w_street, origin, want_pred=True):
        """(dist float64[V], pred int32[V]) of one tree."""
        w = as_array(w_street, np.float64)
        if len(w) != self.S:
            raise ValueError("need one weight per street")
        dist = np.empty(self.V, dtype=np.float64)
        pred = np.empty(self.V, dtype=np.int32) if want_pred else None
        check(self.lib.s4_graph_sssp(self._h, ptr(w, C.c_double), int(origin), ptr(dist, C.c_double),
                                     ptr(pred, C.c_int32) if want_pred else None))
        return dist, pred


class DeviceSim(object):
    """One (virtual) rank: grouped trips, jam tolerance, load counters (s4_sim)."""

    def __init__(self, graph, origin, goal, jam_tolerance, physics):
        self.graph = graph
        self.lib = graph.lib
        origin = as_array(origin, np.int32)
        goal = as_array(goal, np.int32)
        if len(origin) != len(goal):
            raise ValueError("origin/goal arrays differ in length")
        h = C.c_void_p()
        check(self.lib.s4_sim_create(graph._h, len(origin), ptr(origin, C.c_int32), ptr(goal, C.c_int32),
                                     float(jam_tolerance), C.byref(physics), C.byref(h)))
        self._h = h
        self.T = len(origin)
        n = C.c_int64()
        m = C.c_int()
        check(self.lib.s4_sim_roots(h, C.byref(n), C.byref(m)))
        self.n_roots, self.root_mode = n.value, m.value

    @classmethod
    def sampled(cls, graph, n_trips, origins, goals, seed, jam_tolerance, physics):
        """A simulation whose trips are drawn on the device (s4_sim_create_sampled): one uniform origin and
        one uniform goal per trip from the two populations (dense node ids)."""
        self = cls.__new__(cls)
        self.graph = graph
        self.lib = graph.lib
        origins = as_array(origins, np.int32)
        goals = as_array(goals, np.int32)
        h = C.c_void_p()
        check(self.lib.s4_sim_create_sampled(graph._h, int(n_trips), len(origins), ptr(origins, C.c_int32), len(goals),
                                             ptr(goals, C.c_int32), C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF),
                                             float(jam_tolerance), C.byref(physics), C.byref(h)))
        self._h = h
        self.T = int(n_trips)
        n = C.c_int64()
        m = C.c_int()
        check(self.lib.s4_sim_roots(h, C.byref(n), C.byref(m)))
        self.n_roots, self.root_mode = n.value, m.value
        return self

    def get_trips(self):
        """(origin int32[T], goal int32[T]) in dense node ids, grouped by tree root."""
        o = np.empty(self.T, dtype=np.int32)
        g = np.empty(self.T, dtype=np.int32)
        check(self.lib.s4_sim_get_trips(self._h, ptr(o, C.c_int32), ptr(g, C.c_int32)))
        return o, g

    def close(self):
        if getattr(self, "_h", None):
            self.lib.s4_sim_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def S(self):
        return self.graph.S

    def step(self):
        check(self.lib.s4_sim_step(self._h))

    def weights(self, refresh=False):
        """Per-street driving times: the ones the last step routed on, or (refresh)
        recomputed from the resident load."""
        w = np.empty(self.S, dtype=np.float64)
        check(self.lib.s4_sim_weights(self._h, 1 if refresh else 0, ptr(w, C.c_double)))
        return w

    def get_load(self, out=None):
        if out is None:
            out = np.empty(self.S, dtype=np.uint32)
        check(self.lib.s4_sim_get_load(self._h, ptr(out, C.c_uint32)))
        return out

    def set_load(self, load):
        load = as_array(load, np.uint32)
        if len(load) != self.S:
            raise ValueError("traffic_load must have one entry per street")
        check(self.lib.s4_sim_set_load(self._h, ptr(load, C.c_uint32)))

    def load_device_ptr(self):
        return self.lib.s4_sim_load_device_ptr(self._h)

    def load_add(self, other):
        check(self.lib.s4_sim_load_add(self._h, other._h))

    def load_copy(self, other):
        check(self.lib.s4_sim_load_copy(self._h, other._h))

    def commit_total(self):
        check(self.lib.s4_sim_commit_total(self._h))

    def exchange(self):
        """streets4mpi.py:97-100 on the device: NCCL all-reduce over the ranks of Device.comm_init, then the
        cumulative merge."""
        check(self.lib.s4_sim_exchange(self._h))

    def load_hash(self):
        """(FNV-1a of traffic_load, FNV-1a of cumulative_traffic_load or 0)"""
        a, b = C.c_uint64(), C.c_uint64()
        check(self.lib.s4_sim_load_hash(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_cumulative(self):
        out = np.empty(self.S, dtype=np.uint32)
        none = C.c_int()
        check(self.lib.s4_sim_get_cumulative(self._h, ptr(out, C.c_uint32), C.byref(none)))
        return None if none.value else out

    def set_cumulative(self, cum):
        if cum is None:
            check(self.lib.s4_sim_set_cumulative(self._h, None))
            return
        cum = as_array(cum, np.uint32)
        if len(cum) != self.S:
            raise ValueError("cumulative_traffic_load must have one entry per street")
        check(self.lib.s4_sim_set_cumulative(self._h, ptr(cum, C.c_uint32)))

    def road_construction(self, adaptable=None):
        if adaptable is None:
            check(self.lib.s4_sim_road_construction(self._h, None))
        else:
            m = as_array(adaptable, np.uint8)
            if len(m) != self.S:
                raise ValueError("adaptable mask must have one entry per street")
            check(self.lib.s4_sim_road_construction(self._h, ptr(m, C.c_uint8)))

    def get_max_speed(self):
        vis = np.empty(self.S, dtype=np.int32)
        ins = np.empty(self.S, dtype=np.int32)
        check(self.lib.s4_sim_get_max_speed(self._h, ptr(vis, C.c_int32), ptr(ins, C.c_int32)))
        return vis, ins

    def set_max_speed(self, visible, insertion=None):
        vis = as_array(visible, np.int32)
        ins = None if insertion is None else as_array(insertion, np.int32)
        check(self.lib.s4_sim_set_max_speed(self._h, ptr(vis, C.c_int32),
                                            None if ins is None else ptr(ins, C.c_int32)))

    def root_slots(self):
        """(slots int32[groups * K], K): the day's roots as the SSSP kernel groups them, -1 = empty slot (diagnostics)."""
        n, k = C.c_int64(), C.c_int()
        check(self.lib.s4_sim_get_root_slots(self._h, None, 0, C.byref(n), C.byref(k)))
        out = np.empty(max(1, n.value), dtype=np.int32)
        check(self.lib.s4_sim_get_root_slots(self._h, ptr(out, C.c_int32), n.value, C.byref(n), C.byref(k)))
        return out[: n.value], int(k.value)

    def trips_stuck(self):
        n = C.c_uint64()
        check(self.lib.s4_sim_trips_stuck(self._h, C.byref(n)))
        return int(n.value)

    def counters(self):
        c = Counters()
        check(self.lib.s4_sim_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def reset_counters(self):
        check(self.lib.s4_sim_reset_counters(self._h))
