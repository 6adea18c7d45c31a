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


class TripGenerator(object):

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
