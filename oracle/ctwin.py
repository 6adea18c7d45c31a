"""ctypes binding of the C twin (oracle/c/s4_oracle.c).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libs4oracle.so")
_lib = None

PHYS = (4.0, 0.001, 7.5)   # settings.py:37-41 car_length, min_breaking_distance, braking_deceleration


def build(force=False):
    src = os.path.join(_HERE, "c", "s4_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp, ip, lp, up, bp = (C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                              C.POINTER(C.c_uint32), C.POINTER(C.c_uint8))
        L.o_driving_speed.restype = C.c_double
        L.o_driving_speed.argtypes = [C.c_double] * 6
        L.o_weights.restype = None
        L.o_weights.argtypes = [C.c_int64, dp, ip, up, C.c_double, C.c_double, C.c_double, C.c_double, dp]
        L.o_sssp.restype = C.c_int
        L.o_sssp.argtypes = [C.c_int32, lp, ip, ip, dp, C.c_int32, dp, ip, bp]
        L.o_day.restype = C.c_int
        L.o_day.argtypes = [C.c_int32, C.c_int64, lp, ip, ip, dp, C.c_int64, ip, lp, ip,
                            C.c_uint32, up, lp, C.c_int]
        L.o_merge.restype = None
        L.o_merge.argtypes = [C.c_int64, up, up]
        L.o_road_construction.restype = C.c_int
        L.o_road_construction.argtypes = [C.c_int64, up, ip, bp, lp]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Csr(object):
    """Symmetric CSR of an undirected street list (dense node ids).

    Arc order inside a row = street_index order (== python-graph's neighbour
    insertion order for a network built by add_street in index order)."""

    def __init__(self, V, u, v):
        u = np.ascontiguousarray(u, dtype=np.int32)
        v = np.ascontiguousarray(v, dtype=np.int32)
        S = len(u)
        self.V, self.S = int(V), int(S)
        loop = u == v
        tail = np.concatenate([u, v[~loop]])
        head = np.concatenate([v, u[~loop]])
        sid = np.concatenate([np.arange(S, dtype=np.int32), np.arange(S, dtype=np.int32)[~loop]])
        order = np.lexsort((sid, tail))
        self.col = np.ascontiguousarray(head[order], dtype=np.int32)
        self.arc_street = np.ascontiguousarray(sid[order], dtype=np.int32)
        counts = np.bincount(tail, minlength=V)
        self.row_ptr = np.zeros(V + 1, dtype=np.int64)
        np.cumsum(counts, out=self.row_ptr[1:])


def weights(length, max_speed, load, jam, phys=PHYS):
    length = np.ascontiguousarray(length, dtype=np.float64)
    max_speed = np.ascontiguousarray(max_speed, dtype=np.int32)
    load = np.ascontiguousarray(load, dtype=np.uint32)
    out = np.empty(len(length), dtype=np.float64)
    lib().o_weights(len(length), _p(length, C.c_double), _p(max_speed, C.c_int32), _p(load, C.c_uint32),
                    float(jam), float(phys[0]), float(phys[1]), float(phys[2]), _p(out, C.c_double))
    return out


def driving_speed(length, max_speed, trips, phys=PHYS):
    return lib().o_driving_speed(float(length), float(max_speed), float(trips), *map(float, phys))


def sssp(csr, w, src):
    w = np.ascontiguousarray(w, dtype=np.float64)
    dist = np.empty(csr.V, dtype=np.float64)
    pred = np.empty(csr.V, dtype=np.int32)
    tied = np.empty(csr.V, dtype=np.uint8)
    rc = lib().o_sssp(csr.V, _p(csr.row_ptr, C.c_int64), _p(csr.col, C.c_int32), _p(csr.arc_street, C.c_int32),
                      _p(w, C.c_double), int(src), _p(dist, C.c_double), _p(pred, C.c_int32), _p(tied, C.c_uint8))
    if rc:
        raise MemoryError("o_sssp")
    return dist, pred, tied


def group_trips(roots, targets):
    """(distinct roots ascending, offsets, targets sorted by root; stable)."""
    roots = np.asarray(roots, dtype=np.int32)
    targets = np.asarray(targets, dtype=np.int32)
    order = np.argsort(roots, kind="stable")
    r_sorted = roots[order]
    uniq, start = np.unique(r_sorted, return_index=True)
    off = np.concatenate([start, [len(roots)]]).astype(np.int64)
    return (np.ascontiguousarray(uniq, dtype=np.int32), np.ascontiguousarray(off),
            np.ascontiguousarray(targets[order], dtype=np.int32))


def day(csr, w, origins, goals, trip_volume=1, nthreads=1):
    """One simulation.py:68-88 day.  Returns (load uint32[S], counters dict)."""
    w = np.ascontiguousarray(w, dtype=np.float64)
    roots, off, tg = group_trips(origins, goals)
    load = np.zeros(csr.S, dtype=np.uint32)
    cnt = np.zeros(4, dtype=np.int64)
    rc = lib().o_day(csr.V, csr.S, _p(csr.row_ptr, C.c_int64), _p(csr.col, C.c_int32),
                     _p(csr.arc_street, C.c_int32), _p(w, C.c_double), len(roots), _p(roots, C.c_int32),
                     _p(off, C.c_int64), _p(tg, C.c_int32), int(trip_volume), _p(load, C.c_uint32),
                     _p(cnt, C.c_int64), int(nthreads))
    if rc:
        raise MemoryError("o_day")
    return load, dict(hops=int(cnt[0]), unreachable=int(cnt[1]), tied_trips=int(cnt[2]), stuck=int(cnt[3]),
                      n_roots=len(roots))


def road_construction(cumulative, max_speed, adaptable=None):
    cumulative = np.ascontiguousarray(cumulative, dtype=np.uint32)
    ms = np.array(max_speed, dtype=np.int32, copy=True)
    cnt = np.zeros(2, dtype=np.int64)
    mask = None if adaptable is None else np.ascontiguousarray(adaptable, dtype=np.uint8)
    rc = lib().o_road_construction(len(ms), _p(cumulative, C.c_uint32), _p(ms, C.c_int32),
                                   None if mask is None else _p(mask, C.c_uint8), _p(cnt, C.c_int64))
    if rc:
        raise MemoryError("o_road_construction")
    return ms, (int(cnt[0]), int(cnt[1]))
