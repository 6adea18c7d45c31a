"""StreetNetwork: the graph data model of the hot path (reference streetnetwork.py:28-125).

Same public surface as the reference class.  Storage differs: the reference kept
everything in python-graph dictionaries and ran its Dijkstra there; here the
network can be built either call by call (``add_node`` / ``add_street``, kept in
a small undirected-graph container with python-graph's observable behaviour) or
in one shot from arrays (``from_arrays``, for 10^6..10^8-node synthetic graphs),
and ``flat()`` exports the dense arrays the device CSR is built from.
``calculate_shortest_paths`` runs on the GPU (K2 + the predecessor kernel).
"""
from __future__ import annotations

import numpy as np


class _UndirectedGraph(object):
    """Minimal container with the python-graph ``graph`` behaviour the reference
    relies on: both orientations of an edge are keys; the attribute list is a
    separate object per orientation; weights are set on both orientations."""

    def __init__(self):
        self.node_neighbors = {}
        self.node_attr = {}
        self.edge_wt = {}
        self.edge_attr = {}

    def add_node(self, node, attrs=None):
        if node in self.node_neighbors:
            raise ValueError("Node %s already in graph" % (node,))
        self.node_neighbors[node] = []
        self.node_attr[node] = [] if attrs is None else attrs

    def has_node(self, node):
        return node in self.node_neighbors

    def nodes(self):
        return list(self.node_neighbors.keys())

    def node_attributes(self, node):
        return self.node_attr[node]

    def neighbors(self, node):
        return self.node_neighbors[node]

    def __getitem__(self, node):
        return iter(self.node_neighbors[node])

    def has_edge(self, edge):
        return edge in self.edge_wt and (edge[1], edge[0]) in self.edge_wt

    def edges(self):
        return list(self.edge_wt.keys())

    def add_edge(self, edge, wt=1, label="", attrs=None):
        u, v = edge
        if v in self.node_neighbors[u] or u in self.node_neighbors[v]:
            raise ValueError("Edge (%s, %s) already in graph" % (u, v))
        self.node_neighbors[u].append(v)
        self.edge_attr[(u, v)] = list(attrs or [])
        self.edge_wt[(u, v)] = wt
        if u != v:
            self.node_neighbors[v].append(u)
            self.edge_attr[(v, u)] = list(attrs or [])
            self.edge_wt[(v, u)] = wt

    def edge_attributes(self, edge):
        return self.edge_attr.get(edge, [])

    def set_edge_weight(self, edge, wt):
        self.edge_wt[edge] = wt
        self.edge_wt[(edge[1], edge[0])] = wt

    def edge_weight(self, edge):
        return self.edge_wt[edge]


class FlatNetwork(object):
    """Dense-array view of a StreetNetwork (what crosses the C-ABI)."""
    __slots__ = ("node_ids", "u", "v", "length", "max_speed", "max_speed_insertion", "adaptable", "driving_time",
                 "lon", "lat")

    @property
    def V(self):
        return len(self.node_ids)

    @property
    def S(self):
        return len(self.u)

    def locality_rank(self, curve="hilbert"):
        """Position of every node along a space-filling curve over the rank-quantised coordinates (Hilbert by
        default, "morton" for the Z-order curve), or None when the coordinates carry no information.  Pure
        memory-layout hint for the device: neighbouring nodes get neighbouring numbers, and the roots of a day,
        sorted by that number, fall into spatially compact groups (the source-batched SSSP kernel runs one group
        of trees per CTA; the Hilbert curve has no jumps, so no group straddles a quadrant boundary).  No result
        depends on it."""
        if self.lon is None or self.lat is None or len(self.node_ids) < 8:
            return None
        if np.ptp(self.lon) == 0 and np.ptp(self.lat) == 0:
            return None
        n = len(self.node_ids)

        def quantise(x):
            r = np.empty(n, dtype=np.uint64)
            r[np.argsort(x, kind="stable")] = np.arange(n, dtype=np.uint64)
            return (r * np.uint64(65536)) // np.uint64(n)

        qx, qy = quantise(self.lon), quantise(self.lat)
        if curve == "morton":
            def spread(x):
                x = (x | (x << np.uint64(8))) & np.uint64(0x00FF00FF)
                x = (x | (x << np.uint64(4))) & np.uint64(0x0F0F0F0F)
                x = (x | (x << np.uint64(2))) & np.uint64(0x33333333)
                x = (x | (x << np.uint64(1))) & np.uint64(0x55555555)
                return x
            key = spread(qx) | (spread(qy) << np.uint64(1))
        else:
            # Hilbert index of (qx, qy) on the 65536 x 65536 lattice, one bit plane per step (vectorised xy2d)
            x, y = qx.astype(np.int64), qy.astype(np.int64)
            key = np.zeros(n, dtype=np.int64)
            s = 32768
            while s > 0:
                rx = ((x & s) > 0).astype(np.int64)
                ry = ((y & s) > 0).astype(np.int64)
                key += s * s * ((3 * rx) ^ ry)
                flip = (ry == 0) & (rx == 1)
                x = np.where(flip, s - 1 - x, x)
                y = np.where(flip, s - 1 - y, y)
                swap = ry == 0
                x, y = np.where(swap, y, x), np.where(swap, x, y)
                s //= 2
        rank = np.empty(n, dtype=np.int32)
        rank[np.argsort(key, kind="stable")] = np.arange(n, dtype=np.int32)
        return rank

    def dense(self, nodes):
        """original node ids -> dense ids (rank in ascending id order)"""
        nodes = np.asarray(nodes)
        idx = np.searchsorted(self.node_ids, nodes)
        idx = np.clip(idx, 0, max(len(self.node_ids) - 1, 0))
        if len(self.node_ids) == 0 or not np.array_equal(self.node_ids[idx], nodes):
            raise KeyError("unknown node id in trip list")
        return idx.astype(np.int32)


class StreetNetwork(object):
    STREET_ATTRIBUTE_INDEX_INDEX = 0
    STREET_ATTRIBUTE_INDEX_LENGTH = 1
    STREET_ATTRIBUTE_INDEX_MAX_SPEED = 2
    NODE_ATTRIBUTE_INDEX_LONGITUDE = 0
    NODE_ATTRIBUTE_INDEX_LATITUDE = 1

    def __init__(self):
        self._g = _UndirectedGraph()
        self._arrays = None              # set by from_arrays until the dict form is needed
        self.bounds = None
        self.street_index = 0
        self._streets_by_index = dict()
        self._topology_version = 0       # nodes / streets
        self._speed_version = 0          # max_speed
        self._weight_version = 0         # driving times
        self._flat_cache = None
        self._device_graph = None        # (topology_version, device, DeviceGraph)

    # ------------------------------------------------------------------ bulk path
    @classmethod
    def from_arrays(cls, node_ids, u, v, length, max_speed, lon=None, lat=None):
        """Build from arrays: street i joins original node ids u[i], v[i]."""
        self = cls()
        node_ids = np.asarray(node_ids, dtype=np.int64)
        a = {"node_ids": node_ids, "u": np.asarray(u, dtype=np.int64), "v": np.asarray(v, dtype=np.int64),
             "length": np.asarray(length, dtype=np.float64).copy(),
             "max_speed": np.asarray(max_speed, dtype=np.int32).copy(),
             "lon": None if lon is None else np.asarray(lon, dtype=np.float64),
             "lat": None if lat is None else np.asarray(lat, dtype=np.float64)}
        a["max_speed_insertion"] = a["max_speed"].copy()
        a["driving_time"] = a["length"] / a["max_speed"]          # streetnetwork.py:52
        if not (len(a["u"]) == len(a["v"]) == len(a["length"]) == len(a["max_speed"])):
            raise ValueError("street arrays differ in length")
        self._arrays = a
        self.street_index = len(a["u"])
        return self

    def _materialize(self):
        """Switch an array-built network to the call-by-call container."""
        if self._arrays is None:
            return
        a, self._arrays = self._arrays, None
        lon = a["lon"] if a["lon"] is not None else np.zeros(len(a["node_ids"]))
        lat = a["lat"] if a["lat"] is not None else np.zeros(len(a["node_ids"]))
        for n, x, y in zip(a["node_ids"].tolist(), lon.tolist(), lat.tolist()):
            self._g.add_node(n, [x, y])
        self.street_index = 0
        us, vs = a["u"].tolist(), a["v"].tolist()
        for i in range(len(us)):
            key = (us[i], vs[i])
            self._g.add_edge(key, wt=float(a["driving_time"][i]), attrs=[i, float(a["length"][i]), int(a["max_speed"][i])])
            if us[i] != vs[i] and us[i] > vs[i]:
                self._g.edge_attr[key][2] = int(a["max_speed_insertion"][i])
            self._streets_by_index[i] = key
        self.street_index = len(us)

    @property
    def _graph(self):
        self._materialize()
        return self._g

    @property
    def streets_by_index(self):
        self._materialize()
        return self._streets_by_index

    # ------------------------------------------------------------------ reference API
    def has_street(self, street):
        return self._graph.has_edge(street)

    def add_street(self, street, length, max_speed):
        street_attributes = [self.street_index, length, max_speed]
        self._graph.add_edge(street, wt=length / max_speed, attrs=street_attributes)
        self._streets_by_index[self.street_index] = street
        self.street_index += 1
        self._topology_version += 1

    def set_driving_time(self, street, driving_time):
        self._graph.set_edge_weight(street, driving_time)
        self._weight_version += 1

    def get_driving_time(self, street):
        return self._graph.edge_weight(street)

    def get_street_index(self, street):
        return self._graph.edge_attributes(street)[StreetNetwork.STREET_ATTRIBUTE_INDEX_INDEX]

    def get_street_by_index(self, street_index):
        return self.streets_by_index.get(street_index)

    def change_maxspeed(self, street, max_speed_delta):
        attrs = self._graph.edge_attributes(street)
        cur = attrs[StreetNetwork.STREET_ATTRIBUTE_INDEX_MAX_SPEED]
        attrs[StreetNetwork.STREET_ATTRIBUTE_INDEX_MAX_SPEED] = max(1, min(140, cur + max_speed_delta))
        self._speed_version += 1
        # like the reference: no return value (simulation.py:101 depends on that)

    def set_bounds(self, min_latitude, max_latitude, min_longitude, max_longitude):
        self.bounds = ((min_latitude, max_latitude), (min_longitude, max_longitude))

    def add_node(self, node, longitude, latitude):
        self._graph.add_node(node, [longitude, latitude])
        self._topology_version += 1

    def get_nodes(self):
        if self._arrays is not None:
            return self._arrays["node_ids"].tolist()
        return self._g.nodes()

    def node_coordinates(self, node):
        attrs = self._graph.node_attributes(node)
        return (attrs[StreetNetwork.NODE_ATTRIBUTE_INDEX_LONGITUDE], attrs[StreetNetwork.NODE_ATTRIBUTE_INDEX_LATITUDE])

    def has_node(self, node):
        if self._arrays is not None:
            ids = self.flat().node_ids
            i = np.searchsorted(ids, node)
            return bool(i < len(ids) and ids[i] == node)
        return self._g.has_node(node)

    def __iter__(self):
        """(street, street_index, length, max_speed) once per street, street[0] <= street[1]."""
        if self._arrays is not None:
            a = self._arrays
            lo, hi = np.minimum(a["u"], a["v"]).tolist(), np.maximum(a["u"], a["v"]).tolist()
            L, ms = a["length"].tolist(), a["max_speed"].tolist()
            for i in range(len(lo)):
                yield (lo[i], hi[i]), i, L[i], ms[i]
            return
        g = self._g
        for street in g.edges():
            if street[0] > street[1]:
                continue
            at = g.edge_attributes(street)
            yield (street, at[0], at[1], at[2])

    def calculate_shortest_paths(self, origin_node):
        """{node: predecessor} of the shortest-path tree from origin_node on the current
        driving times; origin -> None, unreachable nodes absent.  Runs on the GPU."""
        flat = self.flat()
        dg = self.device_graph()
        origin = int(flat.dense([origin_node])[0])
        dist, pred = dg.sssp(flat.driving_time, origin)
        ids = flat.node_ids
        reach = np.nonzero(np.isfinite(dist))[0]
        paths = dict(zip(ids[reach].tolist(), ids[np.maximum(pred[reach], 0)].tolist()))
        paths[origin_node] = None
        return paths

    # ------------------------------------------------------------------ device plumbing
    def flat(self):
        """Dense arrays (cached until the network is modified)."""
        key = (self._topology_version, self._speed_version, self._weight_version, self._arrays is None)
        if self._flat_cache is not None and self._flat_cache[0] == key:
            return self._flat_cache[1]
        f = FlatNetwork()
        if self._arrays is not None:
            a = self._arrays
            ids = np.unique(a["node_ids"])
            if len(ids) != len(a["node_ids"]):
                raise ValueError("duplicate node ids")
            f.node_ids = ids
            f.u = np.searchsorted(ids, a["u"]).astype(np.int32)
            f.v = np.searchsorted(ids, a["v"]).astype(np.int32)
            if len(a["u"]) and (not np.array_equal(ids[np.clip(f.u, 0, len(ids) - 1)], a["u"]) or
                                not np.array_equal(ids[np.clip(f.v, 0, len(ids) - 1)], a["v"])):
                raise KeyError("street endpoint is not a node")
            f.length, f.max_speed = a["length"], a["max_speed"]
            f.max_speed_insertion, f.driving_time = a["max_speed_insertion"], a["driving_time"]
            f.adaptable = (a["u"] <= a["v"]).astype(np.uint8)
            order = np.argsort(a["node_ids"], kind="stable")
            f.lon = None if a["lon"] is None else a["lon"][order]
            f.lat = None if a["lat"] is None else a["lat"][order]
        else:
            g = self._g
            ids = np.array(sorted(g.node_neighbors.keys()), dtype=np.int64)
            S = self.street_index
            u = np.empty(S, dtype=np.int64)
            v = np.empty(S, dtype=np.int64)
            f.length = np.empty(S, dtype=np.float64)
            f.max_speed = np.empty(S, dtype=np.int32)
            f.max_speed_insertion = np.empty(S, dtype=np.int32)
            f.driving_time = np.empty(S, dtype=np.float64)
            for i in range(S):
                a, b = self._streets_by_index[i]
                u[i], v[i] = a, b
                seen = g.edge_attr[(a, b) if a <= b else (b, a)]       # the list __iter__ reads
                f.length[i] = seen[1]
                f.max_speed[i] = seen[2]
                f.max_speed_insertion[i] = g.edge_attr[(a, b)][2]       # the list change_maxspeed writes
                f.driving_time[i] = g.edge_wt[(a, b)]
            f.node_ids = ids
            f.u = np.searchsorted(ids, u).astype(np.int32)
            f.v = np.searchsorted(ids, v).astype(np.int32)
            f.adaptable = (u <= v).astype(np.uint8)
            coords = [g.node_attr[n] for n in ids.tolist()]
            ok = all(isinstance(c, (list, tuple)) and len(c) >= 2 for c in coords)
            f.lon = np.array([c[0] for c in coords], dtype=np.float64) if ok else None
            f.lat = np.array([c[1] for c in coords], dtype=np.float64) if ok else None
        self._flat_cache = (key, f)
        return f

    def device_graph(self, device=None):
        """The CSR of this network in HBM (uploaded once per topology)."""
        from . import engine
        device = device or engine.default_device()
        cached = self._device_graph
        if cached is not None and cached[0] == self._topology_version and cached[1] is device:
            return cached[2]
        f = self.flat()
        dg = engine.DeviceGraph(device, f.V, f.u, f.v, f.length, f.max_speed, node_rank=f.locality_rank())
        self._device_graph = (self._topology_version, device, dg)
        return dg

    def _apply_max_speed(self, visible, insertion):
        """Write the device's speed limits back (after road_construction on the GPU)."""
        if self._arrays is not None:
            self._arrays["max_speed"][:] = visible
            self._arrays["max_speed_insertion"][:] = insertion
        else:
            g = self._g
            for i in range(self.street_index):
                a, b = self._streets_by_index[i]
                g.edge_attr[(a, b)][2] = int(insertion[i])
                if a > b:
                    g.edge_attr[(b, a)][2] = int(visible[i])
        self._speed_version += 1

    def _apply_driving_time(self, w):
        """Write the device's weights of the day back (simulation.py:65 did set_driving_time per street)."""
        if self._arrays is not None:
            self._arrays["driving_time"][:] = w
        else:
            g = self._g
            for i in range(self.street_index):
                a, b = self._streets_by_index[i]
                g.edge_wt[(a, b)] = float(w[i])
                g.edge_wt[(b, a)] = float(w[i])
        self._weight_version += 1

    def __getstate__(self):
        # an array-built network is pickled in its array form (and materialises lazily after loading): writing
        # street_network_*.s4mpi must not turn a 10^6..10^8-node network into per-edge dicts
        d = dict(self.__dict__)
        d["_flat_cache"] = None
        d["_device_graph"] = None
        return d
