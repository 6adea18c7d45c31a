#!/usr/bin/env python3
"""Generate golden fixtures by EXECUTING THE REFERENCE'S OWN SOURCE under Python 3.

Run in the authoring container only (needs /root/reference; the GPU box has no
reference checkout -- tests read the committed JSON, never this script):

    python tests/golden/make_golden.py

What is executed verbatim from /root/reference:
  settings.py       imported as is
  utils.py          imported as is                       (merge_arrays)
  streetnetwork.py  imported as is, on top of oracle.pygraph_standin registered as
                    ``pygraph`` (python-graph is third party and absent: the Dijkstra
                    and the dict layout behind it are a recalled stand-in => the
                    tie behaviour stays "parity unpinned"; every fixture below is
                    tie-free, which the generator asserts)
  simulation.py     source text up to the stale ``__main__`` block (it is Python-2
                    ``print`` syntax), with the single token ``.iteritems()`` ->
                    ``.items()``; ``osmdata`` (imposm) is stubbed -- simulation.py:30
                    imports it and never uses it
  osmdata.py        source text up to its stale ``__main__`` block, on top of a stub
                    ``imposm.parser.OSMParser`` (imposm is a C extension, absent) that reads
                    OSM XML with xml.etree and calls the four callbacks the way imposm does:
                    coords (id, lon, lat) for every node, nodes (id, tags, (lon, lat)) for
                    tagged nodes only, ways (id, tags, refs), relations (id, tags,
                    [(ref, type, role)]).  The construction rules, the speed table, the
                    haversine and the landuse closure are the reference's code (osm_ingest.json).
  visualization.py  source text up to its ``__main__`` block with the Python-2 syntax made runnable (``print``
                    statements -> calls, ``xrange`` -> ``range``, the three integer divisions of the legend
                    geometry -> ``//``, ``filter`` returning lists), ``persistence.persist_read`` serving the
                    fixture's objects, ``draw.textsize`` shimmed (gone from Pillow 10).  What is recorded is
                    WHAT IT DRAWS, not the pixels: every street line (end points as fp64 hex, fill colour) and
                    the zoom, for four modes x two colour modes (visualization.json).  COMPONENTS is left out:
                    the reference's own ``dict(edge_attributes(street))`` raises on the attribute list its
                    StreetNetwork stores (visualization.py:132).
Nothing else is shimmed: weights, the walk, the load counters, the adaptation
sweep and the speed clamp are the reference's code paths.
"""
import importlib.util
import json
import os
import random
import sys
import types
from array import array

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import pygraph_standin as pg  # noqa: E402


def _install_standins():
    for name in ("pygraph", "pygraph.classes", "pygraph.classes.graph",
                 "pygraph.algorithms", "pygraph.algorithms.minmax"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pygraph.classes.graph"].graph = pg.graph
    sys.modules["pygraph.algorithms.minmax"].shortest_path = pg.shortest_path
    osm = types.ModuleType("osmdata")
    osm.GraphBuilder = object
    sys.modules["osmdata"] = osm


def _load_file(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    _install_standins()
    settings = _load_file("settings")
    utils = _load_file("utils")
    streetnetwork = _load_file("streetnetwork")
    text = open(os.path.join(REF, "simulation.py")).read()
    text = text[:text.index('if __name__ == "__main__":')]
    assert text.count(".iteritems()") == 1
    text = text.replace(".iteritems()", ".items()")
    sim = types.ModuleType("simulation")
    sim.__file__ = os.path.join(REF, "simulation.py")
    sys.modules["simulation"] = sim
    exec(compile(text, sim.__file__, "exec"), sim.__dict__)
    return settings, utils, streetnetwork, sim


def hexf(x):
    return float(x).hex()


# ---------------------------------------------------------------------------
class StubOSMParser(object):
    """imposm.parser.OSMParser's callback protocol over xml.etree (concurrency ignored)."""

    def __init__(self, concurrency=None, coords_callback=None, nodes_callback=None, ways_callback=None,
                 relations_callback=None):
        self.cb = (coords_callback, nodes_callback, ways_callback, relations_callback)

    def parse(self, filename):
        import xml.etree.ElementTree as ET
        coords_cb, nodes_cb, ways_cb, rel_cb = self.cb
        coords, nodes, ways, rels = [], [], [], []
        for el in ET.parse(filename).getroot():
            tags = dict((t.get("k"), t.get("v")) for t in el.findall("tag"))
            if el.tag == "node":
                osmid, lon, lat = int(el.get("id")), float(el.get("lon")), float(el.get("lat"))
                coords.append((osmid, lon, lat))
                if tags:
                    nodes.append((osmid, tags, (lon, lat)))
            elif el.tag == "way":
                ways.append((int(el.get("id")), tags, [int(n.get("ref")) for n in el.findall("nd")]))
            elif el.tag == "relation":
                rels.append((int(el.get("id")), tags,
                             [(int(m.get("ref")), m.get("type"), m.get("role")) for m in el.findall("member")]))
        coords_cb(coords)
        nodes_cb(nodes)
        ways_cb(ways)
        rel_cb(rels)


def load_reference_osmdata():
    """The reference's osmdata module (call after load_reference(): it imports the reference's streetnetwork)."""
    for name in ("imposm", "imposm.parser"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["imposm.parser"].OSMParser = StubOSMParser
    text = open(os.path.join(REF, "osmdata.py")).read()
    text = text[:text.index('if __name__ == "__main__":')]
    mod = types.ModuleType("ref_osmdata")
    mod.__file__ = os.path.join(REF, "osmdata.py")
    exec(compile(text, mod.__file__, "exec"), mod.__dict__)
    return mod


HAND_OSM = """<?xml version="1.0" encoding="UTF-8"?>
<osm version="0.6">
 <node id="1" lat="53.5800" lon="9.9100"/>
 <node id="2" lat="53.5810" lon="9.9110"/>
 <node id="3" lat="53.5820" lon="9.9125"><tag k="landuse" v="commercial"/></node>
 <node id="4" lat="53.5830" lon="9.9140"/>
 <node id="5" lat="53.5830" lon="9.9140"/>
 <node id="6" lat="53.5845" lon="9.9150"/>
 <node id="7" lat="53.5790" lon="9.9080"/>
 <node id="8" lat="53.5795" lon="9.9070"><tag k="name" v="tagged but no landuse"/></node>
 <node id="9" lat="53.5900" lon="9.9300"/>
 <node id="10" lat="-33.9000" lon="151.2000"/>
 <node id="11" lat="-33.9010" lon="151.2020"/>
 <way id="100"><nd ref="1"/><nd ref="2"/><nd ref="3"/><tag k="highway" v="primary"/></way>
 <way id="101"><nd ref="3"/><nd ref="2"/><tag k="highway" v="motorway"/></way>
 <way id="102"><nd ref="3"/><nd ref="4"/><nd ref="5"/><nd ref="6"/><tag k="highway" v="residential"/><tag k="maxspeed" v="30 mph"/></way>
 <way id="103"><nd ref="6"/><nd ref="7"/><tag k="highway" v="trunk"/><tag k="maxspeed" v="none"/></way>
 <way id="104"><nd ref="7"/><nd ref="1"/><tag k="highway" v="bogus_class"/><tag k="maxspeed" v="walk"/></way>
 <way id="105"><nd ref="7"/><nd ref="8"/><tag k="highway" v="footway"/></way>
 <way id="106"><nd ref="8"/><nd ref="9"/><tag k="highway" v="secondary"/><tag k="maxspeed" v="60"/></way>
 <way id="107"><nd ref="10"/><nd ref="11"/><tag k="highway" v="service"/></way>
 <way id="108"><nd ref="4"/><nd ref="4"/><tag k="highway" v="path"/></way>
 <way id="200"><nd ref="1"/><nd ref="2"/><nd ref="7"/><nd ref="1"/><tag k="landuse" v="residential"/></way>
 <way id="201"><nd ref="6"/><nd ref="9"/><nd ref="10"/><tag k="landuse" v="industrial"/></way>
 <way id="202"><nd ref="5"/><nd ref="8"/><tag k="building" v="yes"/></way>
 <way id="203"><nd ref="9"/><nd ref="11"/><tag k="landuse" v="forest"/></way>
 <relation id="300"><member type="way" ref="202" role="outer"/><member type="relation" ref="301" role=""/><tag k="landuse" v="commercial"/></relation>
 <relation id="301"><member type="way" ref="107" role=""/><member type="node" ref="8" role=""/><member type="node" ref="2" role=""/><member type="node" ref="999" role=""/></relation>
 <relation id="302"><member type="way" ref="100" role=""/><tag k="type" v="route"/></relation>
</osm>
"""


def osm_record(name, xml_text, osmdata_mod):
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".osm", delete=False) as f:
        f.write(xml_text)
        path = f.name
    try:
        b = osmdata_mod.GraphBuilder(path)
        net = b.build_street_network()
        b.find_node_categories()
    finally:
        os.unlink(path)
    streets = []
    for i in range(net.street_index):
        st = net.get_street_by_index(i)
        at = net._graph.edge_attributes(st)
        assert at[0] == i
        streets.append([st[0], st[1], hexf(at[1]), at[2], hexf(net.get_driving_time(st))])
    nodes = sorted(net.get_nodes())
    return {"name": name, "xml": xml_text, "streets": streets,
            "nodes": [[n, hexf(net.node_coordinates(n)[0]), hexf(net.node_coordinates(n)[1])] for n in nodes],
            "bounds": None if net.bounds is None else [[hexf(x) for x in pair] for pair in net.bounds],
            "residential": sorted(b.residential_nodes), "industrial": sorted(b.industrial_nodes),
            "commercial": sorted(b.commercial_nodes),
            "connected_residential": sorted(b.connected_residential_nodes),
            "connected_industrial": sorted(b.connected_industrial_nodes),
            "connected_commercial": sorted(b.connected_commercial_nodes)}


def gen_visualization(out_dir, streetnetwork, sim_mod):
    """visualization.py:67-141 executed on a small network and two days of loads: the street lines it draws."""
    import builtins
    import re
    import tempfile
    from PIL import ImageDraw
    text = open(os.path.join(REF, "visualization.py")).read()
    text = text[:text.index('if __name__ == "__main__":')]
    text = re.sub(r"^(\s*)print (.*)$", r"\1print(\2)", text, flags=re.M)
    assert "print " not in re.sub(r"#.*", "", text)
    text = text.replace("xrange(", "range(")
    for old, new in (("self.max_resolution[0] / 40", "self.max_resolution[0] // 40"),
                     ("self.max_resolution[0] / 50", "self.max_resolution[0] // 50"),
                     ("int(bar_outer_width - bar_inner_width) / 2", "int(bar_outer_width - bar_inner_width) // 2")):
        assert text.count(old) == 1
        text = text.replace(old, new)
    rng = random.Random(77)
    nodes = [(100 + i, 9.90 + 0.002 * (i % 6) + rng.uniform(-4e-4, 4e-4), 53.55 + 0.0015 * (i // 6) + rng.uniform(-3e-4, 3e-4))
             for i in range(30)]
    streets = []
    for i in range(30):
        if i % 6 != 5:
            streets.append((100 + i, 100 + i + 1))
        if i + 6 < 30:
            streets.append((100 + i + 6, 100 + i) if i % 4 == 0 else (100 + i, 100 + i + 6))
    speeds = [rng.choice([1, 20, 30, 50, 70, 100, 140]) for _ in streets]
    lengths = [rng.uniform(40.0, 300.0) for _ in streets]
    loads = [[rng.choice([0, 0, 1, 3, 10, 40, 150]) for _ in streets] for _day in range(2)]
    net = streetnetwork.StreetNetwork()
    for (n, lon, lat) in nodes:
        net.add_node(n, lon, lat)
    net.set_bounds(min(x[2] for x in nodes), max(x[2] for x in nodes), min(x[1] for x in nodes), max(x[1] for x in nodes))
    for (st, L, ms) in zip(streets, lengths, speeds):
        net.add_street(st, L, ms)
    files = {"street_network_1.s4mpi": net, "traffic_load_1.s4mpi": array("I", loads[0]), "traffic_load_2.s4mpi": array("I", loads[1])}
    pers = types.ModuleType("persistence")
    pers.persist_read = lambda filename, is_array=False: files[filename]
    sys.modules["persistence"] = pers
    acc = types.ModuleType("pygraph.algorithms.accessibility")
    acc.connected_components = lambda g: {}
    sys.modules["pygraph.algorithms.accessibility"] = acc
    if not hasattr(ImageDraw.ImageDraw, "textsize"):
        def textsize(self, text, font=None):
            box = self.textbbox((0, 0), text, font=font)
            return box[2] - box[0], box[3] - box[1]
        ImageDraw.ImageDraw.textsize = textsize
    mod = types.ModuleType("ref_visualization")
    mod.__file__ = os.path.join(REF, "visualization.py")
    mod.filter = lambda f, xs: list(builtins.filter(f, xs))
    mod.print = lambda *a: None
    exec(compile(text, mod.__file__, "exec"), mod.__dict__)
    recs = []
    real_line = ImageDraw.ImageDraw.line
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            for name in files:
                open(name, "w").close()
            for mode in ("TRAFFIC_LOAD", "MAX_SPEED", "IDEAL_SPEED", "ACTUAL_SPEED"):
                for color_mode in ("HEATMAP", "MONOCHROME"):
                    vis = mod.Visualization("^street_network_[0-9]+.s4mpi$", "^traffic_load_[0-9]+.s4mpi$", mode=mode, color_mode=color_mode)
                    drawn, in_legend = [], [False]
                    fin = vis.image_finalize

                    def finalize(image, max_load, fin=fin, in_legend=in_legend):
                        in_legend[0] = True
                        try:
                            return fin(image, max_load)
                        finally:
                            in_legend[0] = False
                            drawn.append("end of step")
                    vis.image_finalize = finalize

                    def line(self, xy, fill=None, width=0, joint=None, in_legend=in_legend, drawn=drawn):
                        if not in_legend[0]:
                            drawn.append([[hexf(xy[0][0]), hexf(xy[0][1]), hexf(xy[1][0]), hexf(xy[1][1])],
                                          fill if isinstance(fill, str) else list(fill), width])
                        return real_line(self, xy, fill=fill, width=width)
                    ImageDraw.ImageDraw.line = line
                    try:
                        vis.visualize()
                    finally:
                        ImageDraw.ImageDraw.line = real_line
                    steps, cur = [], []
                    for item in drawn:
                        if item == "end of step":
                            steps.append(cur)
                            cur = []
                        else:
                            cur.append(item)
                    order = [[st[0], st[1]] for (st, _i, _l, _m) in net]
                    recs.append({"mode": mode, "color_mode": color_mode, "zoom": hexf(vis.zoom), "iteration_order": order,
                                 "lines_per_step": steps})
        finally:
            os.chdir(cwd)
    json.dump({"source": "visualization.py Visualization.visualize executed under Python 3 (print / xrange / three integer "
                         "divisions adapted); recorded: the street lines handed to ImageDraw.line, in iteration order",
               "nodes": [[n, hexf(lon), hexf(lat)] for (n, lon, lat) in nodes],
               "streets": [[st[0], st[1], hexf(L), ms] for (st, L, ms) in zip(streets, lengths, speeds)],
               "loads": loads, "cases": recs}, open(os.path.join(out_dir, "visualization.json"), "w"))


def gen_osm_ingest(out_dir):
    """osmdata.py:37-231 executed on two OSM XML files: a hand-written one with the rule corner cases (first way
    wins also against the reversed orientation, mph / none / unparsable maxspeed, unknown highway class, footway,
    two nodes on one coordinate (zero-length street), a self loop, the southern hemisphere, landuse on a node, on
    ways and on a relation with a nested relation, a dangling member) and a generated town (synth.write_osm_xml)."""
    import tempfile
    from streets4mpi_b200 import synth
    osmdata_mod = load_reference_osmdata()
    recs = [osm_record("hand_written_rules", HAND_OSM, osmdata_mod)]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "town.osm")
        synth.write_osm_xml(path, rows=7, cols=9, seed=5)
        recs.append(osm_record("generated_town_7x9", open(path).read(), osmdata_mod))
    json.dump({"source": "osmdata.py GraphBuilder.__init__/build_street_network/find_node_categories executed under "
                         "Python 3 behind a stub imposm.parser.OSMParser (xml.etree feeding the four callbacks)",
               "cases": recs}, open(os.path.join(out_dir, "osm_ingest.json"), "w"))


# ---------------------------------------------------------------------------
def gen_speed_vectors(sim):
    rng = random.Random(20261019)
    rows = []
    fixed = [(100.0, 50, 0, 0.5), (100.0, 50, 10, 0.0), (100.0, 50, 10, 0.5), (100.0, 50, 10, 1.0),
             (10.0, 50, 0, 0.3), (3.0, 30, 0, 0.3), (250.0, 140, 7, 0.25), (1000.0, 120, 500, 0.75),
             (16.86, 50, 0, 0.1), (0.0, 50, 0, 0.5), (4.001, 1, 0, 0.9), (104.8, 140, 1, 0.2),
             (1e-9, 140, 4000000000, 0.999), (25000.0, 1, 3, 0.0)]
    for _ in range(400):
        length = rng.choice([rng.uniform(0.5, 30.0), rng.uniform(20.0, 400.0), rng.uniform(400.0, 5000.0)])
        ms = rng.choice([1, 10, 20, 30, 50, 70, 80, 100, 120, 140, rng.randint(1, 140)])
        load = rng.choice([0, 1, 2, rng.randint(0, 50), rng.randint(0, 5000), rng.randint(0, 2 ** 32 - 1)])
        fixed.append((length, ms, load, rng.random()))
    for length, ms, load, jam in fixed:
        ideal = sim.calculate_driving_speed(length, ms, 0)
        actual = sim.calculate_driving_speed(length, ms, load)
        perceived = actual + (ideal - actual) * jam             # simulation.py:61
        rows.append({"length": hexf(length), "max_speed": ms, "load": load, "jam": hexf(jam),
                     "ideal": hexf(ideal), "actual": hexf(actual), "time": hexf(length / perceived)})
    return rows


# ---------------------------------------------------------------------------
def build_network(streetnetwork, nodes, streets):
    net = streetnetwork.StreetNetwork()
    for n in nodes:
        net.add_node(n, 0.0, 0.0)
    for (u, v, length, ms) in streets:
        net.add_street((u, v), length, ms)
    return net


def grid_case(rng, rows, cols, flip=0.0, id_of=lambda r, c, cols: r * cols + c + 1):
    nodes = [id_of(r, c, cols) for r in range(rows) for c in range(cols)]
    streets = []
    for r in range(rows):
        for c in range(cols):
            a = id_of(r, c, cols)
            for (rr, cc) in ((r, c + 1), (r + 1, c)):
                if rr < rows and cc < cols:
                    b = id_of(rr, cc, cols)
                    u, v = (b, a) if rng.random() < flip else (a, b)
                    streets.append((u, v, rng.uniform(20.0, 400.0), rng.choice([30, 50, 70, 100, 140])))
    return nodes, streets


def scatter_case(rng, n, extra):
    """Random connected graph on sparse, shuffled ids plus two isolated nodes and
    one detached pair (unreachable goals, simulation.py:80)."""
    ids = rng.sample(range(1000, 100000), n + 4)
    core, iso, pair = ids[:n], ids[n:n + 2], ids[n + 2:]
    streets, seen = [], set()
    for i in range(1, n):
        j = rng.randrange(0, i)
        u, v = core[i], core[j]
        seen.add((min(u, v), max(u, v)))
        streets.append((u, v, rng.uniform(5.0, 900.0), rng.choice([10, 30, 50, 80, 120])))
    while len(streets) < n - 1 + extra:
        u, v = rng.sample(core, 2)
        k = (min(u, v), max(u, v))
        if k in seen:
            continue
        seen.add(k)
        streets.append((u, v, rng.uniform(5.0, 900.0), rng.choice([10, 30, 50, 80, 120])))
    streets.append((pair[0], pair[1], 77.7, 50))
    return ids, streets, core, iso, pair


def assert_tie_free(net, trips):
    """Every tree used by a fixture must be unique: the predecessor returned by
    the stand-in must be the only neighbour reproducing the distance."""
    g = net._graph
    for origin in trips:
        prev, dist = pg.shortest_path(g, origin)
        prev2, dist2 = pg.shortest_path_scan(g, origin)
        assert dist == dist2 and prev == prev2, "variants disagree"
        for v, dv in dist.items():
            if v == origin:
                continue
            cands = [u for u in g.neighbors(v) if u in dist and dist[u] + g.edge_weight((u, v)) == dv]
            assert cands == [prev[v]], ("tie at", v, cands)


def run_days(utils, streetnetwork, sim_mod, settings, nodes, streets, trips_per_rank, jams, days, every,
             trip_volume=1, check_ties=True):
    """streets4mpi.py:84-106 with the Allreduce replaced by utils.merge_arrays."""
    settings.settings["trip_volume"] = trip_volume
    sims = [sim_mod.Simulation(build_network(streetnetwork, nodes, streets), trips_per_rank[r], jams[r],
                               lambda *a: None) for r in range(len(jams))]
    out = {"days": []}
    for day in range(days):
        rec = {}
        if day > 0 and day % every == 0:
            for s in sims:
                s.road_construction()
            net0 = sims[0].street_network
            rec["max_speed_after_construction"] = [
                dict((idx, ms) for (_st, idx, _l, ms) in net0)[i] for i in range(net0.street_index)]
        for s in sims:
            s.step()
            if check_ties:
                assert_tie_free(s.street_network, s.trips)
        rec["weights"] = []
        for s in sims:
            net = s.street_network
            w = dict((idx, net.get_driving_time(st)) for (st, idx, _l, _ms) in net)
            rec["weights"].append([hexf(w[i]) for i in range(net.street_index)])
        rec["local_loads"] = [list(s.traffic_load) for s in sims]
        total = utils.merge_arrays([s.traffic_load for s in sims])           # Allreduce(SUM)
        for s in sims:
            s.traffic_load = total
            s.cumulative_traffic_load = utils.merge_arrays((total, s.cumulative_traffic_load))
        rec["total"] = list(total)
        rec["cumulative"] = list(sims[0].cumulative_traffic_load)
        out["days"].append(rec)
    settings.settings["trip_volume"] = 1
    return out


def make_trips(rng, n, origins, goals):
    trips = {}
    for _ in range(n):
        o, g = rng.choice(origins), rng.choice(goals)
        trips.setdefault(o, []).append(g)
    return trips


def case_record(name, nodes, streets, trips_per_rank, jams, days, every, result, trip_volume=1):
    return {"name": name, "nodes": nodes,
            "streets": [[u, v, hexf(l), ms] for (u, v, l, ms) in streets],
            "trips": [[[o, g] for o, gs in t.items() for g in gs] for t in trips_per_rank],
            "jams": [hexf(j) for j in jams], "n_days": days, "construction_every": every,
            "trip_volume": trip_volume, **result}


def main():
    settings, utils, streetnetwork, sim = load_reference()
    out_dir = HERE

    json.dump({"source": "simulation.py:129-138 and :57-63 executed under Python 3",
               "settings": {k: settings.settings[k] for k in ("car_length", "min_breaking_distance",
                                                              "braking_deceleration", "trip_volume")},
               "rows": gen_speed_vectors(sim)},
              open(os.path.join(out_dir, "k1_weights.json"), "w"), indent=0)

    cases = []
    # (1) the only fixture the reference ships: simulation.py:147-155 (floats for lengths)
    nodes, streets = [1, 2, 3], [(1, 2, 10.0, 50), (2, 3, 100.0, 140)]
    trips = [{1: [3]}]
    cases.append(case_record("ref_three_node", nodes, streets, trips, [0.3], 4, 10,
                             run_days(utils, streetnetwork, sim, settings, nodes, streets, trips, [0.3], 4, 10)))

    rng = random.Random(3756917)
    # (2) 7x9 grid, one rank, congestion feedback over 6 days, adaptation every 2 days,
    #     30 % of the streets inserted as (larger id, smaller id): orientation aliasing
    nodes, streets = grid_case(rng, 7, 9, flip=0.3)
    trips = [make_trips(rng, 160, nodes, rng.sample(nodes, 6))]
    jams = [rng.random()]
    cases.append(case_record("grid7x9_1rank", nodes, streets, trips, jams, 6, 2,
                             run_days(utils, streetnetwork, sim, settings, nodes, streets, trips, jams, 6, 2)))
    # (3) 12x12 grid, three virtual ranks with their own jam tolerance (streets4mpi.py:78)
    nodes, streets = grid_case(rng, 12, 12, flip=0.0)
    goals = rng.sample(nodes, 9)
    trips = [make_trips(rng, 120, nodes, goals) for _ in range(3)]
    jams = [rng.random() for _ in range(3)]
    cases.append(case_record("grid12x12_3ranks", nodes, streets, trips, jams, 5, 3,
                             run_days(utils, streetnetwork, sim, settings, nodes, streets, trips, jams, 5, 3)))
    # (4) scattered ids, unreachable goals, origin == goal, duplicate trips, trip_volume 3
    ids, streets, core, iso, pair = scatter_case(rng, 70, 45)
    trips = [make_trips(rng, 90, core, rng.sample(core, 8)) for _ in range(2)]
    trips[0].setdefault(core[0], []).extend([core[0], iso[0], pair[0], core[5], core[5]])
    trips[1].setdefault(pair[0], []).extend([pair[1], core[3]])
    trips[1].setdefault(iso[1], []).append(core[1])
    jams = [0.0, 0.9999]
    cases.append(case_record("scatter70_2ranks", ids, streets, trips, jams, 4, 2,
                             run_days(utils, streetnetwork, sim, settings, ids, streets, trips, jams, 4, 2,
                                      trip_volume=3), trip_volume=3))
    json.dump({"source": "simulation.py Simulation.step/road_construction + utils.merge_arrays executed under "
                         "Python 3 on oracle.pygraph_standin; all trees asserted tie-free",
               "cases": cases}, open(os.path.join(out_dir, "days.json"), "w"))

    # road adaptation alone: sizes around the S<20 / S>=20 regimes
    recs = []
    for S in (1, 2, 10, 19, 20, 21, 100, 101, 257):
        nodes = list(range(1, S + 2))
        streets = []
        for i in range(S):
            u, v = i + 1, i + 2
            if rng.random() < 0.4:
                u, v = v, u
            streets.append((u, v, 50.0, rng.choice([1, 10, 15, 30, 50, 120, 130, 140])))
        net = build_network(streetnetwork, nodes, streets)
        s = sim.Simulation(net, {}, 0.5, lambda *a: None)
        cum = [rng.choice([0, 0, 1, 2, 5, rng.randint(0, 1000)]) for _ in range(S)]
        s.cumulative_traffic_load = array("I", cum)
        s.road_construction()
        seen_by_iter = dict((idx, ms) for (_st, idx, _l, ms) in net)
        by_index = [net._graph.edge_attributes(net.get_street_by_index(i))[2] for i in range(S)]
        recs.append({"S": S, "streets": [[u, v, ms] for (u, v, _l, ms) in streets], "cumulative": cum,
                     "max_speed_seen_by_weight_loop": [seen_by_iter[i] for i in range(S)],
                     "max_speed_insertion_orientation": by_index,
                     "cumulative_after": s.cumulative_traffic_load})
    json.dump({"source": "simulation.py:91-109 + streetnetwork.py:79-82,113-125 executed under Python 3",
               "cases": recs}, open(os.path.join(out_dir, "road_construction.json"), "w"))

    # merge_arrays (utils.py:26-34)
    a = array("I", [rng.randint(0, 2 ** 31) for _ in range(50)])
    b = array("I", [rng.randint(0, 2 ** 31 - 1) for _ in range(50)])
    json.dump({"a": list(a), "b": list(b), "a_b": list(utils.merge_arrays((a, b))),
               "a_none": list(utils.merge_arrays((a, None)))},
              open(os.path.join(out_dir, "merge_arrays.json"), "w"))
    gen_osm_ingest(out_dir)
    gen_visualization(out_dir, streetnetwork, sim)
    print("golden fixtures written to", out_dir)


if __name__ == "__main__":
    main()
