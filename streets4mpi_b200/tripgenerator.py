"""Workload generation (reference tripgenerator.py:30-44).

``TripGenerator.generate_trips`` keeps the reference contract: it returns
``{origin: [goal, ...]}`` drawn with one ``random.sample(pop, 1)[0]`` per
endpoint from the module-level ``random`` stream.  (The reference's
``origin in trips.keys()`` made it O(n^2); the result is the same dict.)

``TripTable`` is the flat form the device consumes -- two int arrays -- with a
read-only dict view, for the 10^6..10^8-trip configurations where a Python dict
of lists is not an option (SURVEY section 7 "trip containers at scale").
"""
from random import sample

import numpy as np


class TripTable(object):
    """Trips as parallel arrays of (original) node ids, quacking like the dict."""

    def __init__(self, origins, goals):
        self.origins = np.ascontiguousarray(origins)
        self.goals = np.ascontiguousarray(goals)
        if self.origins.shape != self.goals.shape or self.origins.ndim != 1:
            raise ValueError("origins and goals must be 1-d arrays of equal length")
        self._index = None

    def __len__(self):
        return len(self.keys())

    @property
    def number_of_trips(self):
        return len(self.origins)

    def _build(self):
        if self._index is None:
            order = np.argsort(self.origins, kind="stable")
            so = self.origins[order]
            uniq, start = np.unique(so, return_index=True)
            end = np.append(start[1:], len(so))
            self._index = (uniq, start, end, self.goals[order])
        return self._index

    def keys(self):
        return self._build()[0].tolist()

    def __iter__(self):
        return iter(self.keys())

    def __contains__(self, origin):
        uniq = self._build()[0]
        i = np.searchsorted(uniq, origin)
        return bool(i < len(uniq) and uniq[i] == origin)

    def __getitem__(self, origin):
        uniq, start, end, goals = self._build()
        i = np.searchsorted(uniq, origin)
        if i >= len(uniq) or uniq[i] != origin:
            raise KeyError(origin)
        return goals[start[i]:end[i]].tolist()

    def items(self):
        uniq, start, end, goals = self._build()
        for i in range(len(uniq)):
            yield uniq[i].item(), goals[start[i]:end[i]].tolist()

    @staticmethod
    def from_dict(trips):
        o, g = [], []
        for origin, goals in trips.items():
            o.extend([origin] * len(goals))
            g.extend(goals)
        return TripTable(np.array(o, dtype=np.int64), np.array(g, dtype=np.int64))


def philox_trip_indices(n_trips, n_origins, n_goals, seed):
    """The draws of the device's trip sampler (csrc/s4_kernels.cuh sample_trips_kernel) in numpy: trip t takes the
    four words of Philox4x32-10 at counter (t, 0, 0, 0) under key (seed & 0xffffffff, seed >> 32); words 0,1 pick
    the origin, words 2,3 the goal, index = floor(r64 * n / 2**64).  Returns (origin_index, goal_index)."""
    t = np.arange(n_trips, dtype=np.uint64)
    c = [t & np.uint64(0xFFFFFFFF), t >> np.uint64(32), np.zeros(n_trips, np.uint64), np.zeros(n_trips, np.uint64)]
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & m32]
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF

    def scale(hi, lo, n):                       # floor(((hi << 32) | lo) * n / 2**64) without 128-bit integers
        n = np.uint64(n)
        low = (lo * n) >> np.uint64(32)         # lo, hi < 2**32 and n < 2**31: every product fits in 64 bits
        return ((hi * n + low) >> np.uint64(32)).astype(np.int64)
    return scale(c[0], c[1], n_origins), scale(c[2], c[3], n_goals)


class SampledTrips(object):
    """Trips that are drawn on the device when the Simulation is built (TripGenerator.sample_trips_on_device):
    only the two populations travel to the GPU.  ``table()`` gives the dict-like view (fetched back once)."""

    def __init__(self, number_of_residents, potential_origins, potential_goals, seed):
        self.number_of_trips = int(number_of_residents)
        self.origins = np.asarray(sorted(potential_origins) if not isinstance(potential_origins, np.ndarray) else potential_origins)
        self.goals = np.asarray(sorted(potential_goals) if not isinstance(potential_goals, np.ndarray) else potential_goals)
        if self.number_of_trips > 0 and (len(self.origins) == 0 or len(self.goals) == 0):
            raise ValueError("Sample larger than population or is negative")     # what random.sample raised
        self.seed = int(seed)
        self._table = None
        self._fetch = None                       # set by Simulation: () -> (origins, goals) as original node ids

    def host_table(self):
        """The same trips computed on the host (numpy Philox): what the device draws."""
        io, ig = philox_trip_indices(self.number_of_trips, len(self.origins), len(self.goals), self.seed)
        return TripTable(self.origins[io], self.goals[ig])

    def table(self):
        if self._table is None:
            self._table = TripTable(*self._fetch()) if self._fetch else self.host_table()
        return self._table

    def keys(self):
        return self.table().keys()

    def items(self):
        return self.table().items()

    def __getitem__(self, origin):
        return self.table()[origin]

    def __contains__(self, origin):
        return origin in self.table()

    def __iter__(self):
        return iter(self.table())

    def __len__(self):
        return len(self.table())


class TripGenerator(object):

    def sample_trips_on_device(self, number_of_residents, potential_origins, potential_goals, seed):
        """Like generate_trips (one uniform origin and one uniform goal per resident), but drawn by the GPU when
        the Simulation is created: nothing of size number_of_residents is built on the host."""
        return SampledTrips(number_of_residents, potential_origins, potential_goals, seed)


    def generate_trips(self, number_of_residents, potential_origins, potential_goals):
        # Python >= 3.11 no longer samples from sets; freeze them in sorted order
        # (Python 2 used hash order -- trip lists are run inputs, not outputs)
        if not isinstance(potential_origins, (list, tuple)):
            potential_origins = sorted(potential_origins)
        if not isinstance(potential_goals, (list, tuple)):
            potential_goals = sorted(potential_goals)
        trips = dict()
        for _ in range(0, number_of_residents):
            origin = sample(potential_origins, 1)[0]
            goal = sample(potential_goals, 1)[0]
            trips.setdefault(origin, []).append(goal)
        return trips

    def generate_trip_table(self, number_of_residents, potential_origins, potential_goals, seed):
        """Array form for large runs: uniform draws with a seeded numpy generator."""
        rng = np.random.default_rng(seed)
        po = np.asarray(sorted(potential_origins) if not isinstance(potential_origins, np.ndarray) else potential_origins)
        pg = np.asarray(sorted(potential_goals) if not isinstance(potential_goals, np.ndarray) else potential_goals)
        if len(po) == 0 or len(pg) == 0:
            raise ValueError("Sample larger than population or is negative")     # what random.sample raised
        return TripTable(po[rng.integers(0, len(po), size=number_of_residents)],
                         pg[rng.integers(0, len(pg), size=number_of_residents)])
