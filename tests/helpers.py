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


def visualization_lines(case_file, tmp_path, modes):
    """Run streets4mpi_b200.visualization on the inputs of tests/golden/visualization.json and return, per (mode, colour
    mode), the street lines it hands to ImageDraw.line for every step -- next to what the reference drew."""
    import os
    from array import array
    from PIL import ImageDraw
    from streets4mpi_b200 import StreetNetwork, persistence
    from streets4mpi_b200.visualization import Visualization
    gold = load_golden(case_file)
    net = StreetNetwork()
    for (n, lon, lat) in gold["nodes"]:
        net.add_node(n, fh(lon), fh(lat))
    lons, lats = [fh(x[1]) for x in gold["nodes"]], [fh(x[2]) for x in gold["nodes"]]
    net.set_bounds(min(lats), max(lats), min(lons), max(lons))
    for (u, v, L, ms) in gold["streets"]:
        net.add_street((u, v), fh(L), ms)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    out = []
    real_line = ImageDraw.ImageDraw.line
    try:
        persistence.persist_write("street_network_1.s4mpi", net)
        for day, load in enumerate(gold["loads"]):
            persistence.persist_write("traffic_load_%d.s4mpi" % (day + 1), array("I", load), is_array=True)
        for case in gold["cases"]:
            if case["mode"] not in modes:
                continue
            vis = Visualization("^street_network_[0-9]+.s4mpi$", "^traffic_load_[0-9]+.s4mpi$", mode=case["mode"],
                                color_mode=case["color_mode"], verbose=False)
            steps, cur, in_legend = [], [], [False]
            fin = vis.image_finalize

            def finalize(image, max_load, fin=fin, in_legend=in_legend, steps=steps, cur=cur):
                in_legend[0] = True
                try:
                    return fin(image, max_load)
                finally:
                    in_legend[0] = False
                    steps.append(list(cur))
                    del cur[:]
            vis.image_finalize = finalize

            def line(self, xy, fill=None, width=0, joint=None, in_legend=in_legend, cur=cur):
                if not in_legend[0]:
                    cur.append([[float(xy[0][0]).hex(), float(xy[0][1]).hex(), float(xy[1][0]).hex(), float(xy[1][1]).hex()],
                                fill if isinstance(fill, str) else list(fill), width])
                return real_line(self, xy, fill=fill, width=width)
            ImageDraw.ImageDraw.line = line
            try:
                vis.visualize()
            finally:
                ImageDraw.ImageDraw.line = real_line
            out.append((case, float(vis.zoom).hex(), steps))
    finally:
        os.chdir(cwd)
    return out
