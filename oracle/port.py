"""Python-3 restatement of the Streets4MPI hot path (CPU oracle, "port").

TEST INFRASTRUCTURE.  Every routine names the reference lines it follows
(paths are relative to /root/reference).  The arithmetic is IEEE-754 double in
exactly the reference's operation order; integer loads behave like
``array('I')``.  The graph container sits on ``oracle.pygraph_standin`` (a
recalled restatement of python-graph: parity unpinned at the Dijkstra, see
``oracle/__init__.py``).

Layout differs from the reference on purpose (one module, free functions over a
small state object) -- it is a restatement, not a copy.
"""
from __future__ import annotations

import heapq
import random as _random
from array import array
from math import sqrt

from . import pygraph_standin as pg

# settings.py:24-44 -- the physics constants and knobs the path reads
DEFAULTS = {
    "random_seed": 3756917,
    "max_simulation_steps": 10,
    "number_of_residents": 100,
    "use_residential_origins": False,
    "traffic_period_duration": 8,
    "car_length": 4,
    "min_breaking_distance": 0.001,
    "braking_deceleration": 7.5,
    "steps_between_street_construction": 10,
    "trip_volume": 1,
}


# --------------------------------------------------------------------------
# A2  simulation.py:129-138
# --------------------------------------------------------------------------
def driving_speed(street_length, max_speed, number_of_trips, cfg=DEFAULTS):
    """km/h a car can drive on a street carrying ``number_of_trips`` cars.

    simulation.py:131  space per car  = length / max(trips, 1)
    simulation.py:132  braking room   = max(space - car_length, min_breaking_distance)
    simulation.py:134  potential m/s  = sqrt(decel * room * 2)   [(decel*room)*2]
    simulation.py:136  result km/h    = min(max_speed, potential * 3.6)
    """
    room_per_car = street_length / max(number_of_trips, 1)
    braking_room = max(room_per_car - cfg["car_length"], cfg["min_breaking_distance"])
    potential = sqrt(cfg["braking_deceleration"] * braking_room * 2)
    return min(max_speed, potential * 3.6)


def driving_time(length, max_speed, load, jam_tolerance, cfg=DEFAULTS):
    """simulation.py:57-63 for one street: the SSSP weight of the day."""
    ideal = driving_speed(length, max_speed, 0, cfg)            # :57
    actual = driving_speed(length, max_speed, load, cfg)        # :59
    perceived = actual + (ideal - actual) * jam_tolerance       # :61
    return length / perceived                                   # :63


# --------------------------------------------------------------------------
# A6  utils.py:26-34
# --------------------------------------------------------------------------
def merge_arrays(arrays):
    out = array("I", [0]) * len(arrays[0])
    for arr in arrays:
        if arr is None:                       # utils.py:30 (``!= None``)
            continue
        for i in range(len(arr)):
            out[i] += arr[i]                  # OverflowError past 2**32-1, like array('I')
    return out


# --------------------------------------------------------------------------
# A1  streetnetwork.py:28-125
# --------------------------------------------------------------------------
class Network(object):
    """streetnetwork.py:35-41 state: graph, bounds, running street index, index->street."""

    IDX, LENGTH, MAXSPEED = 0, 1, 2          # streetnetwork.py:29-31

    def __init__(self, sssp="canonical"):
        self._graph = pg.graph()
        self.bounds = None
        self.street_index = 0
        self.streets_by_index = {}
        self.sssp = sssp                      # "canonical" | "pygraph" | "pygraph_scan"

    def add_node(self, node, lon, lat):       # streetnetwork.py:89-91
        self._graph.add_node(node, [lon, lat])

    def has_node(self, node):                 # streetnetwork.py:104-105
        return self._graph.has_node(node)

    def get_nodes(self):                      # streetnetwork.py:94-95
        return self._graph.nodes()

    def has_street(self, street):             # streetnetwork.py:44-45
        return self._graph.has_edge(street)

    def add_street(self, street, length, max_speed):   # streetnetwork.py:48-57
        attrs = [self.street_index, length, max_speed]
        self._graph.add_edge(street, wt=length / max_speed, attrs=attrs)
        self.streets_by_index[self.street_index] = street
        self.street_index += 1

    def set_driving_time(self, street, t):    # streetnetwork.py:60-61
        self._graph.set_edge_weight(street, t)

    def get_driving_time(self, street):       # streetnetwork.py:64-65
        return self._graph.edge_weight(street)

    def get_street_index(self, street):       # streetnetwork.py:68-69
        return self._graph.edge_attributes(street)[self.IDX]

    def get_street_by_index(self, i):         # streetnetwork.py:72-76
        return self.streets_by_index.get(i)

    def change_maxspeed(self, street, delta):  # streetnetwork.py:79-82 (returns None!)
        attrs = self._graph.edge_attributes(street)
        attrs[self.MAXSPEED] = max(1, min(140, attrs[self.MAXSPEED] + delta))

    def streets(self):                        # streetnetwork.py:113-125 (__iter__)
        for key in self._graph.edges():
            if key[0] > key[1]:
                continue
            a = self._graph.edge_attributes(key)
            yield key, a[self.IDX], a[self.LENGTH], a[self.MAXSPEED]

    __iter__ = streets

    def calculate_shortest_paths(self, origin):   # streetnetwork.py:108-109
        if self.sssp == "pygraph":
            return pg.shortest_path(self._graph, origin)[0]
        if self.sssp == "pygraph_scan":
            return pg.shortest_path_scan(self._graph, origin)[0]
        return canonical_tree(self._graph, origin)[0]


# --------------------------------------------------------------------------
# A4  canonical Dijkstra (tie rule of DESIGN.md / SURVEY 8(c))
# --------------------------------------------------------------------------
PLATEAU_MAX = 32     # nodes a plateau search may touch before the trip counts as stuck (same bound on the GPU)


def canonical_tree(g, origin):
    """Dijkstra with strict ``<`` (distances equal python-graph's), then the
    schedule-independent predecessor rule.  Candidates of v are the neighbours u != v with
    ``fl(dist[u] + w(u,v)) == dist[v]``:

    * if some candidate is strictly closer to the origin (``dist[u] < dist[v]``), the one with the smallest id;
    * otherwise v sits on a *plateau* (zero or absorbed weights: its distance is only reproduced by neighbours at
      the same distance).  A breadth-first search over the plateau -- candidates in ascending id order, at most
      ``PLATEAU_MAX`` nodes -- finds the nearest node that has a strictly closer candidate; pred[v] is the first
      hop towards it.  Every step reduces the plateau level, so predecessor chains never cycle (a "smallest
      equal-cost neighbour" rule can oscillate between two nodes of a zero-weight chain).

    Returns ``(previous, dist, tied)``; ``tied`` = nodes with more than one candidate.
    """
    dist = {origin: 0}
    heap = [(0, origin)]
    done = set()
    while heap:
        du, u = heapq.heappop(heap)
        if u in done:
            continue
        done.add(u)
        for v in g.neighbors(u):
            if v in done:
                continue
            alt = du + g.edge_weight((u, v))
            if v not in dist or alt < dist[v]:
                dist[v] = alt
                heapq.heappush(heap, (alt, v))

    def candidates(v):
        dv = dist[v]
        return sorted(set(u for u in g.neighbors(v)
                          if u in dist and u != v and dist[u] + g.edge_weight((u, v)) == dv))

    def closer(v):
        return [u for u in candidates(v) if dist[u] < dist[v]]

    previous = {origin: None}
    tied = set()
    for v in dist:
        if v == origin:
            continue
        cands = candidates(v)
        near = [u for u in cands if dist[u] < dist[v]]
        if len(cands) > 1:
            tied.add(v)
        if near:
            previous[v] = near[0]
            continue
        # plateau search
        first = {v: None}
        queue = [v]
        found = None
        while queue and found is None:
            x = queue.pop(0)
            for u in candidates(x):                     # all at dist[v]: x has no closer candidate
                if u in first:
                    continue
                if len(first) >= PLATEAU_MAX:
                    queue = []
                    break
                first[u] = u if x == v else first[x]
                if u == origin or closer(u):
                    found = first[u]
                    break
                queue.append(u)
        previous[v] = found                              # None: stuck (plateau larger than the bound)
    return previous, dist, tied


# --------------------------------------------------------------------------
# A3 + A5 + A7  simulation.py:37-109
# --------------------------------------------------------------------------
class Sim(object):
    def __init__(self, network, trips, jam_tolerance, log=None, cfg=DEFAULTS):   # :37-45
        self.street_network = network
        self.trips = trips
        self.jam_tolerance = jam_tolerance
        self.log = log or (lambda *a: None)
        self.cfg = cfg
        self.step_counter = 0
        self.traffic_load = array("I", [0]) * network.street_index
        self.cumulative_traffic_load = None

    def step(self):                                                   # :48-88
        net = self.street_network
        self.step_counter += 1
        for street, idx, length, max_speed in net:                    # :53-65
            t = driving_time(length, max_speed, self.traffic_load[idx],
                             self.jam_tolerance, self.cfg)
            net.set_driving_time(street, t)
        self.traffic_load = array("I", [0]) * net.street_index         # :68
        volume = self.cfg["trip_volume"]
        for origin in self.trips.keys():                              # :71
            paths = net.calculate_shortest_paths(origin)              # :75
            for goal in self.trips[origin]:                           # :78
                if goal not in paths:                                 # :80
                    continue
                here = goal
                while here != origin:                                 # :83
                    nxt = paths[here]
                    key = (min(here, nxt), max(here, nxt))            # :84
                    here = nxt                                        # :85
                    self.traffic_load[net.get_street_index(key)] += volume   # :86-88

    def road_construction(self):                                      # :91-109
        net = self.street_network
        by_street = {}
        for i in range(len(self.cumulative_traffic_load)):            # :93-95
            by_street[net.get_street_by_index(i)] = self.cumulative_traffic_load[i]
        # :96 -- sorted() is stable; Py3 dicts iterate in insertion order, i.e. by
        # street_index (Py2's hash order is unpinned)
        ranked = sorted(by_street.items(), key=lambda kv: kv[1])
        n = len(ranked)
        dec_limit = 0.15 * n                                          # :97
        inc_limit = 0.95 * n                                          # :98
        for i in range(n):                                            # :99
            if i <= dec_limit:                                        # :100
                if not net.change_maxspeed(ranked[i][0], -20):        # :101 always true
                    dec_limit += 1                                    # :102
            j = n - i - 1                                             # :103
            if j >= inc_limit:                                        # :104
                if not net.change_maxspeed(ranked[j][0], 20):         # :105 always true
                    inc_limit -= 1                                    # :106
            if dec_limit >= inc_limit:                                # :107
                break
        self.cumulative_traffic_load = None                           # :109


def construction_plan(n):
    """Run simulation.py:97-108's control flow for ``n`` ranked streets and
    return the rank lists (decreased, increased) in application order."""
    dec, inc = [], []
    dec_limit = 0.15 * n
    inc_limit = 0.95 * n
    for i in range(n):
        if i <= dec_limit:
            dec.append(i)
            dec_limit += 1
        j = n - i - 1
        if j >= inc_limit:
            inc.append(j)
            inc_limit -= 1
        if dec_limit >= inc_limit:
            break
    return dec, inc


# --------------------------------------------------------------------------
# tripgenerator.py:30-44
# --------------------------------------------------------------------------
def generate_trips(n, origins, goals, rng=_random):
    """{origin: [goal, ...]} with one ``sample(pop, 1)[0]`` draw per endpoint
    (tripgenerator.py:34-35).  Python >=3.11 refuses sets in ``sample``; a set
    population is frozen with ``sorted`` first (Py2 used hash order: trip lists
    are inputs to parity tests)."""
    if not isinstance(origins, (list, tuple)):
        origins = sorted(origins)
    if not isinstance(goals, (list, tuple)):
        goals = sorted(goals)
    trips = {}
    for _ in range(n):
        o = rng.sample(origins, 1)[0]
        g = rng.sample(goals, 1)[0]
        trips.setdefault(o, []).append(g)      # :37-42
    return trips


# --------------------------------------------------------------------------
# A8 + A6  streets4mpi.py:44-106 with P virtual ranks in one process
# --------------------------------------------------------------------------
def run_days(make_network, trips_per_rank, jams, days, construction_every, cfg=DEFAULTS):
    """Serial stand-in for ``mpiexec -n P``: ``make_network()`` builds one graph
    copy per rank (streets4mpi.py:54-57), rank r routes ``trips_per_rank[r]``
    with ``jams[r]`` (:75-82); the Allreduce (:97-99) is a plain sum.

    Returns the list of per-day totals (``array('I')``) and the rank-0 network.
    """
    sims = [Sim(make_network(), trips_per_rank[r], jams[r], cfg=cfg)
            for r in range(len(jams))]
    totals = []
    for day in range(days):                                           # :84
        if day > 0 and day % construction_every == 0:                 # :86
            for s in sims:
                s.road_construction()                                 # :88
        for s in sims:
            s.step()                                                  # :93
        total = merge_arrays([s.traffic_load for s in sims])          # :97-98
        for s in sims:
            s.traffic_load = total                                    # :99
            s.cumulative_traffic_load = merge_arrays((total, s.cumulative_traffic_load))  # :100
        totals.append(total)
    return totals, sims[0].street_network
