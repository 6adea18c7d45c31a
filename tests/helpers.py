"""Shared helpers for the test-suite (fixtures -> arrays, oracle drivers)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def fh(x):
    return float.fromhex(x)


def case_arrays(case):
    """Golden day-case -> dense arrays.  Dense node id = rank of the original id
    (ascending), the mapping the host StreetNetwork uses."""
    nodes = sorted(case["nodes"])
    dense = {n: i for i, n in enumerate(nodes)}
    u = np.array([dense[s[0]] for s in case["streets"]], dtype=np.int32)
    v = np.array([dense[s[1]] for s in case["streets"]], dtype=np.int32)
    length = np.array([fh(s[2]) for s in case["streets"]], dtype=np.float64)
    ms = np.array([s[3] for s in case["streets"]], dtype=np.int32)
    trips = []
    for t in case["trips"]:
        o = np.array([dense[p[0]] for p in t], dtype=np.int32)
        g = np.array([dense[p[1]] for p in t], dtype=np.int32)
        trips.append((o, g))
    jams = [fh(j) for j in case["jams"]]
    # orientation aliasing: streets inserted with u > v never show
    # a changed speed limit to the weight loop
    adaptable = np.array([1 if s[0] <= s[1] else 0 for s in case["streets"]], dtype=np.uint8)
    return dict(V=len(nodes), u=u, v=v, length=length, max_speed=ms, trips=trips, jams=jams,
                adaptable=adaptable, nodes=nodes, dense=dense)
