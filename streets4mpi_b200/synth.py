"""Seeded synthetic street networks and workloads of the BASELINE shapes.

The reference's own input (osm/test.osm) is a missing blob and there is no
network, so every configuration runs on generated graphs: 4-neighbour grids
(cfg3: 1024x1024) and thinned-Delaunay planar road graphs (cfg4/cfg5).  Lengths
are continuous (ties have probability ~0), speed limits follow a road-class
mix with a few arterial rows / columns, goals are a small "commercial +
industrial" node sample and origins are all nodes or a "residential" sample.
"""
from __future__ import annotations

import numpy as np

from .streetnetwork import StreetNetwork

SPEED_CLASSES = np.array([30, 50, 70, 100, 140], dtype=np.int32)
SPEED_WEIGHTS = np.array([0.45, 0.30, 0.12, 0.08, 0.05])


def grid_arrays(rows, cols, seed=1, arterial_every=64):
    """Street arrays of a rows x cols grid; node id = r * cols + c."""
    rng = np.random.default_rng(seed)
    idx = np.arange(rows * cols, dtype=np.int64).reshape(rows, cols)
    hu, hv = idx[:, :-1].ravel(), idx[:, 1:].ravel()        # horizontal streets
    vu, vv = idx[:-1, :].ravel(), idx[1:, :].ravel()        # vertical streets
    u = np.concatenate([hu, vu])
    v = np.concatenate([hv, vv])
    S = len(u)
    length = rng.uniform(20.0, 400.0, size=S)
    ms = rng.choice(SPEED_CLASSES, size=S, p=SPEED_WEIGHTS).astype(np.int32)
    if arterial_every:
        h_row = (hu // cols) % arterial_every == arterial_every // 2
        v_col = (vu % cols) % arterial_every == arterial_every // 2
        art = np.concatenate([h_row, v_col])
        ms[art] = rng.choice(np.array([100, 120, 140], dtype=np.int32), size=int(art.sum()))
    rr, cc = np.divmod(idx.ravel(), cols)
    return {"node_ids": idx.ravel(), "u": u, "v": v, "length": length, "max_speed": ms,
            "lon": cc.astype(np.float64), "lat": rr.astype(np.float64)}


def planar_arrays(n_nodes, seed=1, mean_degree=2.4, extent=None):
    """Random planar road graph: uniform points, Delaunay triangulation thinned to a
    spanning tree plus the shortest remaining edges up to ``mean_degree`` (real road
    graphs sit near 2.4).  Euclidean lengths in metres."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import minimum_spanning_tree
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    extent = extent or 150.0 * np.sqrt(n_nodes)
    pts = rng.uniform(0.0, extent, size=(n_nodes, 2))
    tri = Delaunay(pts)
    e = np.concatenate([tri.simplices[:, [0, 1]], tri.simplices[:, [1, 2]], tri.simplices[:, [0, 2]]])
    e.sort(axis=1)
    e = np.unique(e, axis=0)
    d = np.hypot(*(pts[e[:, 0]] - pts[e[:, 1]]).T)
    mst = minimum_spanning_tree(coo_matrix((d, (e[:, 0], e[:, 1])), shape=(n_nodes, n_nodes))).tocoo()
    key = e[:, 0].astype(np.int64) * n_nodes + e[:, 1]
    lo, hi = np.minimum(mst.row, mst.col).astype(np.int64), np.maximum(mst.row, mst.col).astype(np.int64)
    tree_mask = np.isin(key, lo * n_nodes + hi)
    want = int(mean_degree * n_nodes / 2)
    extra = max(0, want - int(tree_mask.sum()))
    rest = np.nonzero(~tree_mask)[0]
    pick = rest[np.argsort(d[rest] * rng.uniform(0.7, 1.3, size=len(rest)))[:extra]]
    keep = np.concatenate([np.nonzero(tree_mask)[0], pick])
    keep.sort()
    u, v, length = e[keep, 0].astype(np.int64), e[keep, 1].astype(np.int64), d[keep]
    length = np.maximum(length, 1.0)
    ms = rng.choice(SPEED_CLASSES, size=len(u), p=SPEED_WEIGHTS).astype(np.int32)
    return {"node_ids": np.arange(n_nodes, dtype=np.int64), "u": u, "v": v, "length": length, "max_speed": ms,
            "lon": pts[:, 0], "lat": pts[:, 1]}


def ladder_arrays(rows, cols, seed=1, rung_fraction=0.2):
    """Country-scale road-like graph in O(V) time and memory (BASELINE cfg5 shape: planar_arrays' Delaunay does not
    scale to 5e7 points): every row of a rows x cols lattice is a street chain, and a seeded `rung_fraction` of the
    vertical links joins neighbouring rows (at least one per pair of rows, so the graph is connected).  Mean degree
    2 + 2 * rung_fraction (2.4 like real road graphs), jittered lengths, the usual speed mix; node id = r * cols + c."""
    rng = np.random.default_rng(seed)
    n = rows * cols
    ids = np.arange(n, dtype=np.int64)
    hu = ids[(ids % cols) != cols - 1]
    hv = hu + 1
    vu_all = ids[: n - cols]
    keep = rng.random(len(vu_all)) < rung_fraction
    keep[np.arange(rows - 1, dtype=np.int64) * cols + rng.integers(0, cols, size=rows - 1)] = True
    vu = vu_all[keep]
    vv = vu + cols
    u = np.concatenate([hu, vu])
    v = np.concatenate([hv, vv])
    S = len(u)
    length = rng.uniform(20.0, 400.0, size=S)
    ms = rng.choice(SPEED_CLASSES, size=S, p=SPEED_WEIGHTS).astype(np.int32)
    rr, cc = np.divmod(ids, cols)
    return {"node_ids": ids, "u": u, "v": v, "length": length, "max_speed": ms,
            "lon": cc.astype(np.float64), "lat": rr.astype(np.float64)}


def road_arrays(n_nodes, seed=1, arterial_every=16, motorway_every=128, keep_local=0.62):
    """City-scale road-like graph with a speed hierarchy, O(V) time and memory (BASELINE cfg4 shape).

    A jittered rows x cols lattice of junctions.  Local streets (30 / 50 km/h) survive with probability
    ``keep_local`` (dead ends and irregular blocks, mean degree ~2.6 like real road graphs); every
    ``arterial_every``-th row and column is a complete arterial (70 / 100 km/h), every ``motorway_every``-th a
    motorway (120 / 140 km/h), so the graph is connected along the arterial mesh and a trip uses O(sqrt(V)) hops
    like on a real network -- unlike ``planar_arrays`` (spanning tree + shortest leftover Delaunay edges: no
    through roads, tens of thousands of hops per trip), which stays available as the tortuous stress case.
    Junctions cut off inside a block by the thinning keep their node ids; trips to them are skipped like any
    unreachable goal (simulation.py:80).  Lengths are Euclidean on the jittered positions (metres, ~120 m spacing)."""
    rng = np.random.default_rng(seed)
    cols = int(np.ceil(np.sqrt(n_nodes)))
    rows = int(np.ceil(n_nodes / cols))
    n = rows * cols
    ids = np.arange(n, dtype=np.int64)
    rr, cc = np.divmod(ids, cols)
    x = (cc + rng.uniform(-0.3, 0.3, n)) * 120.0
    y = (rr + rng.uniform(-0.3, 0.3, n)) * 120.0

    def level(line):            # 2 motorway, 1 arterial, 0 local
        return np.where(line % motorway_every == motorway_every // 2, 2, np.where(line % arterial_every == arterial_every // 2, 1, 0))

    hu = ids[cc != cols - 1]
    h_lvl = level(hu // cols)                                   # a horizontal street belongs to its row
    vu = ids[: n - cols]
    v_lvl = level(vu % cols)                                    # a vertical street to its column
    keep_h = (h_lvl > 0) | (rng.random(len(hu)) < keep_local)
    keep_v = (v_lvl > 0) | (rng.random(len(vu)) < keep_local)
    u = np.concatenate([hu[keep_h], vu[keep_v]])
    v = np.concatenate([hu[keep_h] + 1, vu[keep_v] + cols])
    lvl = np.concatenate([h_lvl[keep_h], v_lvl[keep_v]])
    length = np.maximum(np.hypot(x[u] - x[v], y[u] - y[v]) * rng.uniform(1.0, 1.3, len(u)), 5.0)
    ms = np.where(lvl == 2, rng.choice(np.array([120, 140], dtype=np.int32), len(u)),
                  np.where(lvl == 1, rng.choice(np.array([70, 100], dtype=np.int32), len(u)),
                           rng.choice(np.array([30, 50], dtype=np.int32), len(u), p=[0.6, 0.4]))).astype(np.int32)
    return {"node_ids": ids, "u": u, "v": v, "length": length, "max_speed": ms, "lon": x, "lat": y}


def network_from(arrays):
    return StreetNetwork.from_arrays(arrays["node_ids"], arrays["u"], arrays["v"], arrays["length"],
                                     arrays["max_speed"], arrays.get("lon"), arrays.get("lat"))


def node_categories(node_ids, seed=1, goal_fraction=0.02, residential_fraction=0.01):
    """Seeded stand-ins for osmdata's connected_{commercial,industrial,residential}_nodes."""
    rng = np.random.default_rng(seed + 7919)
    node_ids = np.asarray(node_ids)
    n = len(node_ids)
    n_goal = max(1, int(round(goal_fraction * n)))
    goals = rng.choice(node_ids, size=n_goal, replace=False)
    half = n_goal // 2
    residential = rng.choice(node_ids, size=max(1, int(round(residential_fraction * n))), replace=False)
    return {"commercial": np.sort(goals[:half]) if half else np.sort(goals), "industrial": np.sort(goals[half:]),
            "goals": np.sort(goals), "residential": np.sort(residential)}


def trip_arrays(n_trips, origins, goals, seed):
    """Uniform (origin, goal) draws, the array form of tripgenerator.py:33-35."""
    rng = np.random.default_rng(seed)
    origins, goals = np.asarray(origins), np.asarray(goals)
    return origins[rng.integers(0, len(origins), size=n_trips)], goals[rng.integers(0, len(goals), size=n_trips)]


def parse_spec(spec):
    """'synthetic:grid:ROWSxCOLS[:seed]', 'synthetic:planar:N[:seed]', 'synthetic:road:N[:seed]' or
    'synthetic:ladder:ROWSxCOLS[:seed]' -> arrays"""
    parts = spec.split(":")
    if len(parts) < 3 or parts[0] != "synthetic":
        raise ValueError("not a synthetic network spec: %r" % (spec,))
    seed = int(parts[3]) if len(parts) > 3 else 1
    if parts[1] == "grid":
        r, c = parts[2].lower().split("x")
        return grid_arrays(int(r), int(c), seed)
    if parts[1] == "planar":
        return planar_arrays(int(parts[2]), seed)
    if parts[1] == "road":
        return road_arrays(int(parts[2]), seed)
    if parts[1] == "ladder":
        r, c = parts[2].lower().split("x")
        return ladder_arrays(int(r), int(c), seed)
    raise ValueError("unknown synthetic network kind %r" % (parts[1],))


def write_osm_xml(path, rows=24, cols=30, seed=7, lat0=53.585, lon0=9.915, spacing_m=120.0):
    """A small 'Hamburg-Stellingen-like' OSM XML file (the reference's osm/test.osm is a missing blob):
    a jittered street grid with a road-class mix, a few explicit maxspeed tags (digits, mph, none), footways,
    and landuse polygons / one landuse relation that reuse street nodes, so that the residential, commercial
    and industrial node sets are connected to the network."""
    rng = np.random.default_rng(seed)
    dlat = spacing_m / 111320.0
    dlon = spacing_m / (111320.0 * np.cos(np.radians(lat0)))
    nid = lambda r, c: 100000 + r * cols + c
    out = ['<?xml version="1.0" encoding="UTF-8"?>', '<osm version="0.6" generator="streets4mpi_b200.synth">']
    for r in range(rows):
        for c in range(cols):
            la = lat0 + (r + rng.uniform(-0.25, 0.25)) * dlat
            lo = lon0 + (c + rng.uniform(-0.25, 0.25)) * dlon
            out.append('<node id="%d" lat="%.7f" lon="%.7f"/>' % (nid(r, c), la, lo))
    wid = [500000]

    def way(refs, tags):
        wid[0] += 1
        out.append('<way id="%d">' % wid[0])
        out.extend('<nd ref="%d"/>' % x for x in refs)
        out.extend('<tag k="%s" v="%s"/>' % kv for kv in tags.items())
        out.append("</way>")
        return wid[0]

    classes = ["residential"] * 6 + ["tertiary", "tertiary", "secondary", "unclassified", "service", "footway"]
    for r in range(rows):
        tags = {"highway": "primary" if r == rows // 2 else str(rng.choice(classes))}
        if r % 7 == 3:
            tags["maxspeed"] = str(rng.choice(["30", "50", "30 mph", "none"]))
        way([nid(r, c) for c in range(cols)], tags)
    for c in range(cols):
        tags = {"highway": "motorway" if c == cols // 3 else str(rng.choice(classes))}
        if c % 9 == 4:
            tags["maxspeed"] = str(rng.choice(["20", "60", "70"]))
        way([nid(r, c) for r in range(rows)], tags)

    def block(r0, c0, h, w_):
        ring = [nid(r0, c) for c in range(c0, c0 + w_)] + [nid(r, c0 + w_ - 1) for r in range(r0 + 1, r0 + h)]
        ring += [nid(r0 + h - 1, c) for c in range(c0 + w_ - 2, c0 - 1, -1)] + [nid(r, c0) for r in range(r0 + h - 2, r0, -1)]
        return ring + [ring[0]]
    way(block(1, 1, rows // 3, cols // 3), {"landuse": "residential"})
    way(block(rows // 2 + 1, 2, rows // 3, cols // 4), {"landuse": "residential"})
    w_ind = way(block(2, cols // 2 + 2, rows // 4, cols // 4), {"landuse": "industrial"})
    w_com = way(block(rows // 2, cols // 2, rows // 4, cols // 3), {"landuse": "commercial"})
    w_com2 = way(block(rows - 6, cols - 8, 4, 5), {})
    out.append('<relation id="900001"><member type="way" ref="%d" role="outer"/><tag k="landuse" v="commercial"/>'
               '<tag k="type" v="multipolygon"/></relation>' % w_com2)
    out.append("</osm>")
    with open(path, "w") as f:
        f.write("\n".join(out))
    return {"rows": rows, "cols": cols, "industrial_way": w_ind, "commercial_way": w_com}
