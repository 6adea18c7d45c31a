"""CPU-side checks: the C-ABI library loads and exports every declared symbol, and the
host mirror of the reference API (no compute calls -- those need a GPU)."""
import os
import pickle
import random
import re
from array import array

import numpy as np
import pytest

from tests.helpers import ROOT, fh, load_golden


def test_build_and_symbols():
    import __graft_entry__ as ge
    ge.build()
    from streets4mpi_b200 import _ffi
    lib = _ffi.load()
    header = open(os.path.join(ROOT, "include", "s4mpi.h")).read()
    declared = set(re.findall(r"\b(s4_[a-z0-9_]+)\s*\(", header))
    declared -= {"s4_ctx", "s4_graph", "s4_sim", "s4_physics", "s4_counters"}
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    for name in declared:
        assert getattr(lib, name)
    assert lib.s4_abi_version() == 1


def test_counters_struct_matches_header():
    from streets4mpi_b200 import _ffi
    header = open(os.path.join(ROOT, "include", "s4mpi.h")).read()
    body = header[header.index("typedef struct s4_counters {"):header.index("} s4_counters;")]
    fields = re.findall(r"(?:uint64_t|double)\s+([a-z_]+);", body)
    assert fields == [n for n, _ in _ffi.Counters._fields_]
    import ctypes
    assert ctypes.sizeof(_ffi.Physics) == 32


def test_no_device_fails_loudly():
    """No GPU in this container: creating a context must raise, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from streets4mpi_b200 import _ffi, engine
    with pytest.raises(_ffi.S4Error) as e:
        engine.Device(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "streets4mpi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text
                assert "libs4oracle" not in text and "oracle/_build" not in text and "ctwin" not in text


def test_street_network_api_matches_reference_fixture():
    from streets4mpi_b200 import StreetNetwork
    case = load_golden("days.json")["cases"][1]
    net = StreetNetwork()
    for n in case["nodes"]:
        net.add_node(n, float(n), -float(n))
    for (u, v, L, ms) in case["streets"]:
        assert not net.has_street((u, v))
        net.add_street((u, v), fh(L), ms)
    assert net.street_index == len(case["streets"])
    u, v, L, ms = case["streets"][5]
    assert net.has_street((u, v)) and net.has_street((v, u))
    assert net.get_street_index((u, v)) == 5 == net.get_street_index((v, u))
    assert net.get_street_by_index(5) == (u, v) and net.get_street_by_index(10 ** 6) is None
    assert net.get_driving_time((u, v)) == fh(L) / ms
    net.set_driving_time((u, v), 3.5)
    assert net.get_driving_time((v, u)) == 3.5
    assert net.node_coordinates(case["nodes"][0]) == (float(case["nodes"][0]), -float(case["nodes"][0]))
    assert sorted(net.get_nodes()) == sorted(case["nodes"]) and net.has_node(case["nodes"][3])
    rows = list(net)
    assert len(rows) == len(case["streets"]) and all(st[0] <= st[1] for st, _, _, _ in rows)
    assert net.change_maxspeed((u, v), 500) is None           # no return value, like the reference
    with pytest.raises(Exception):
        net.add_street((u, v), 1.0, 50)
    net.set_bounds(1, 2, 3, 4)
    assert net.bounds == ((1, 2), (3, 4))
    # orientation aliasing: the change is visible to __iter__ only for streets inserted as u <= v
    seen = dict((idx, m) for (_s, idx, _l, m) in net)[5]
    assert seen == (140 if u <= v else ms)
    flat = net.flat()
    assert flat.V == len(case["nodes"]) and flat.S == len(case["streets"])
    assert flat.max_speed_insertion[5] == 140 and flat.max_speed[5] == seen
    assert flat.adaptable.tolist() == [1 if s[0] <= s[1] else 0 for s in case["streets"]]
    clone = pickle.loads(pickle.dumps(net, protocol=2))
    assert list(clone) == list(net) and clone.street_index == net.street_index


def test_street_network_from_arrays_equals_call_by_call():
    from streets4mpi_b200 import StreetNetwork, synth
    arr = synth.grid_arrays(6, 7, seed=2)
    a = synth.network_from(arr)
    b = StreetNetwork()
    for n in arr["node_ids"].tolist():
        b.add_node(n, 0.0, 0.0)
    for u, v, L, ms in zip(arr["u"].tolist(), arr["v"].tolist(), arr["length"].tolist(), arr["max_speed"].tolist()):
        b.add_street((u, v), L, ms)
    fa, fb = a.flat(), b.flat()
    for name in ("node_ids", "u", "v", "length", "max_speed", "max_speed_insertion", "adaptable", "driving_time"):
        assert np.array_equal(getattr(fa, name), getattr(fb, name)), name
    assert sorted(a) == sorted(b)
    assert a.has_node(5) and not a.has_node(10 ** 7)
    # pickling (street_network_*.s4mpi) keeps the array form of the live object and of the clone
    c = synth.network_from(arr)
    clone = pickle.loads(pickle.dumps(c, protocol=2))
    assert c._arrays is not None and clone._arrays is not None
    assert np.array_equal(clone.flat().length, fa.length) and sorted(clone) == sorted(b)
    assert a.get_street_index((1, 0)) == 0            # materialises the dict form
    assert a.flat().S == fb.S


def test_trip_generator_and_table():
    from streets4mpi_b200 import TripGenerator, TripTable
    random.seed(3756917)
    trips = TripGenerator().generate_trips(500, list(range(100)), {200, 201, 202})
    assert sum(len(g) for g in trips.values()) == 500
    assert set(trips).issubset(range(100)) and all(set(g) <= {200, 201, 202} for g in trips.values())
    # same stream, same draws as the reference loop (tripgenerator.py:33-35)
    random.seed(3756917)
    want = {}
    for _ in range(500):
        o = random.sample(list(range(100)), 1)[0]
        g = random.sample([200, 201, 202], 1)[0]
        want.setdefault(o, []).append(g)
    assert trips == want
    with pytest.raises(ValueError):
        TripGenerator().generate_trips(1, [], [1])
    t = TripTable.from_dict(trips)
    assert t.number_of_trips == 500 and sorted(t.keys()) == sorted(trips)
    for o in trips:
        assert sorted(t[o]) == sorted(trips[o]) and o in t
    assert 12345 not in t
    big = TripGenerator().generate_trip_table(10000, np.arange(50), np.arange(50, 60), seed=1)
    assert big.number_of_trips == 10000 and len(big) <= 50


def test_merge_arrays_and_persistence(tmp_path):
    from streets4mpi_b200 import merge_arrays, persistence
    g = load_golden("merge_arrays.json")
    a, b = array("I", g["a"]), array("I", g["b"])
    m = merge_arrays((a, b))
    assert isinstance(m, array) and m.typecode == "I" and list(m) == g["a_b"]
    assert list(merge_arrays((a, None))) == g["a_none"]
    with pytest.raises(OverflowError):
        merge_arrays((array("I", [2 ** 32 - 1]), array("I", [1])))
    f = str(tmp_path / "traffic_load_1.s4mpi")
    persistence.persist_write(f, m, is_array=True)
    import zlib
    assert zlib.decompress(open(f, "rb").read()) == np.array(g["a_b"], dtype="<u4").tobytes()
    assert list(persistence.persist_read(f, is_array=True)) == g["a_b"]
    f2 = str(tmp_path / "obj.s4mpi")
    persistence.persist_write(f2, {"x": [1, 2]})
    assert persistence.persist_read(f2) == {"x": [1, 2]}
    assert persistence.persist_deserialize(persistence.persist_serialize([3], compress=False), compressed=False) == [3]


def test_settings_keys_match_reference():
    from streets4mpi_b200.settings import settings
    assert set(settings) == {"osm_file", "logging", "persist_traffic_load", "random_seed", "max_simulation_steps",
                             "number_of_residents", "use_residential_origins", "traffic_period_duration",
                             "car_length", "min_breaking_distance", "braking_deceleration",
                             "steps_between_street_construction", "trip_volume"}
    assert settings["random_seed"] == 3756917 and settings["braking_deceleration"] == 7.5


def test_synthetic_generators():
    from streets4mpi_b200 import synth
    g = synth.grid_arrays(16, 16, seed=1)
    assert len(g["u"]) == 2 * 16 * 15 and g["length"].min() >= 20 and g["max_speed"].min() >= 30
    assert np.array_equal(synth.grid_arrays(16, 16, seed=1)["length"], g["length"])
    p = synth.planar_arrays(1500, seed=2)
    assert abs(2 * len(p["u"]) / 1500 - 2.4) < 0.05
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    m = sp.coo_matrix((np.ones(len(p["u"])), (p["u"], p["v"])), shape=(1500, 1500))
    assert connected_components(m, directed=False)[0] == 1
    cats = synth.node_categories(g["node_ids"], seed=1)
    assert len(cats["goals"]) == 5 and set(cats["commercial"]) | set(cats["industrial"]) == set(cats["goals"])
    assert synth.parse_spec("synthetic:grid:4x5:9")["node_ids"].shape == (20,)
    lad = synth.ladder_arrays(40, 50, seed=3)                      # cfg5 shape: street chains joined by sparse rungs
    assert abs(2 * len(lad["u"]) / 2000 - 2.4) < 0.1 and lad["length"].min() >= 20
    m = sp.coo_matrix((np.ones(len(lad["u"])), (lad["u"], lad["v"])), shape=(2000, 2000))
    assert connected_components(m, directed=False)[0] == 1
    assert len(synth.parse_spec("synthetic:ladder:6x7")["node_ids"]) == 42


def test_async_writer_matches_persist_write(tmp_path):
    """persistence.AsyncWriter: same bytes as persist_write, snapshot at submission, errors surface at close()."""
    from array import array
    from streets4mpi_b200 import persistence
    load = array("I", range(5000))
    obj = {"a": [1, 2, 3], "b": "street"}
    persistence.persist_write(str(tmp_path / "ref_load.s4mpi"), load, is_array=True)
    persistence.persist_write(str(tmp_path / "ref_obj.s4mpi"), obj)
    with persistence.AsyncWriter(max_pending=2) as w:
        w.write(str(tmp_path / "load.s4mpi"), load, is_array=True)
        load[0] = 77                                                   # later changes must not reach the file
        w.write(str(tmp_path / "obj.s4mpi"), obj)
        obj["a"].append(4)
        w.write(str(tmp_path / "load_np.s4mpi"), np.arange(10, dtype=np.uint32), is_array=True)
        w.flush()
    assert (tmp_path / "load.s4mpi").read_bytes() == (tmp_path / "ref_load.s4mpi").read_bytes()
    assert (tmp_path / "obj.s4mpi").read_bytes() == (tmp_path / "ref_obj.s4mpi").read_bytes()
    assert persistence.persist_read(str(tmp_path / "load_np.s4mpi"), is_array=True).tolist() == list(range(10))
    w = persistence.AsyncWriter()
    w.write(str(tmp_path / "no_such_dir" / "x.s4mpi"), load, is_array=True)
    with pytest.raises(OSError):
        w.close()


def test_philox_trip_indices_known_answers():
    """The host mirror of the device's trip sampler: Philox4x32-10 known answers (Random123 kat_vectors: counter 0 / key 0
    -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8; all ones -> 408f276d 41c83b0e a20bc7c6 6d5451fd) and uniform indices."""
    from streets4mpi_b200 import tripgenerator as tg
    # counter (t, 0, 0, 0) with t = 0 and key (0, 0): index = floor(r64 * n / 2^64) with n = 2^31 - 1 exposes the high bits
    io, ig = tg.philox_trip_indices(1, 2 ** 31 - 1, 2 ** 31 - 1, 0)
    r_o, r_g = (0x6627e8d5 << 32) | 0xe169c58d, (0xbc57ac4c << 32) | 0x9b00dbd8
    assert int(io[0]) == (r_o * (2 ** 31 - 1)) >> 64 and int(ig[0]) == (r_g * (2 ** 31 - 1)) >> 64
    io, ig = tg.philox_trip_indices(200000, 37, 11, 3756917)
    assert io.min() == 0 and io.max() == 36 and ig.min() == 0 and ig.max() == 10
    counts = np.bincount(io, minlength=37)
    assert abs(counts - 200000 / 37).max() < 6 * np.sqrt(200000 / 37)         # uniform within 6 sigma
    st = tg.TripGenerator().sample_trips_on_device(1000, [5, 9, 2], {7, 8}, seed=42)
    t = st.host_table()
    assert t.number_of_trips == 1000 and set(t.origins.tolist()) <= {2, 5, 9} and set(t.goals.tolist()) <= {7, 8}
    with pytest.raises(ValueError):
        tg.TripGenerator().sample_trips_on_device(10, [], [1], seed=1)


def test_rank_seeding_known_answers():
    """streets4mpi.py:50-51,78: CPython's Mersenne Twister is the same in Py2 and Py3."""
    from streets4mpi_b200 import distributed
    random.seed(distributed.rank_seed(3756917, 0))
    assert random.random() == 0.461979755332244
    random.seed(distributed.rank_seed(3756917, 1))
    assert random.random() == 0.010228509196787305
    random.seed(distributed.rank_seed(3756917, 7))
    assert random.random() == 0.0975399640380532
    assert distributed.residents_of_rank(100, 8) == 12 and distributed.residents_of_rank(100000, 3) == 33333


def test_osm_ingest_rules(tmp_path):
    """osmdata.GraphBuilder on a generated OSM XML file: the reference's construction rules."""
    from streets4mpi_b200 import synth
    from streets4mpi_b200.osmdata import GraphBuilder, haversine_m, speed_limit
    path = str(tmp_path / "test.osm")
    meta = synth.write_osm_xml(path, rows=10, cols=12, seed=3)
    b = GraphBuilder(path)
    net = b.build_street_network()
    assert len(net.get_nodes()) == 120 and net.street_index == 10 * 11 + 12 * 9
    assert net.bounds is not None and net.bounds[0][0] < net.bounds[0][1]
    # first way wins, every consecutive ref pair is a street, lengths are haversine metres
    (st, idx, length, ms) = next(iter(net))
    assert 60.0 < length < 200.0 and ms >= 1
    la1, lo1 = b.coords[st[0]]
    la2, lo2 = b.coords[st[1]]
    assert length == haversine_m(la1, lo1, la2, lo2)
    assert abs(haversine_m(53.5, 9.9, 53.6, 9.9) - 6367000 * np.radians(0.1)) < 1e-6
    # speed-limit rules (osmdata.py:119-129)
    assert speed_limit({"highway": "motorway"}) == 140 and speed_limit({"highway": "footway"}) == 1
    assert speed_limit({"highway": "bogus"}) == 50 and speed_limit({"highway": "residential", "maxspeed": "70"}) == 70
    assert speed_limit({"highway": "primary", "maxspeed": "30 mph"}) == 30          # no unit conversion
    assert speed_limit({"highway": "primary", "maxspeed": "none"}) == 140
    assert speed_limit({"highway": "primary", "maxspeed": "walk"}) == 100
    b.find_node_categories()
    assert b.connected_residential_nodes and b.connected_industrial_nodes and b.connected_commercial_nodes
    assert b.connected_industrial_nodes == set(b.all_osm_ways[meta["industrial_way"]][2])
    assert b.connected_commercial_nodes > set(b.all_osm_ways[meta["commercial_way"]][2])    # + the relation's way
    assert b.get_all_child_nodes(123456789) == set()
    flat = net.flat()
    assert flat.locality_rank() is not None and sorted(flat.locality_rank().tolist()) == list(range(120))


def test_osm_ingest_matches_reference_fixture(tmp_path):
    """f3 pinned: tests/golden/osm_ingest.json holds what the reference's own osmdata.py (executed behind a stub
    imposm parser, tests/golden/make_golden.py) builds from two OSM XML files; this package's reader must build the
    same street list (order, fp64 lengths and driving times bit for bit, speed limits), node coordinates, bounds
    and landuse node sets."""
    from streets4mpi_b200.osmdata import GraphBuilder
    for case in load_golden("osm_ingest.json")["cases"]:
        path = str(tmp_path / (case["name"] + ".osm"))
        with open(path, "w") as f:
            f.write(case["xml"])
        b = GraphBuilder(path)
        net = b.build_street_network()
        b.find_node_categories()
        assert net.street_index == len(case["streets"])
        for i, (u, v, length, ms, w) in enumerate(case["streets"]):
            st = net.get_street_by_index(i)
            assert tuple(st) == (u, v), (case["name"], i)
            at = net._graph.edge_attributes(st)
            assert at[0] == i and float(at[1]).hex() == length and at[2] == ms
            assert float(net.get_driving_time(st)).hex() == w
        assert sorted(net.get_nodes()) == [n for n, _, _ in case["nodes"]]
        for n, lon, lat in case["nodes"]:
            assert [float(x).hex() for x in net.node_coordinates(n)] == [lon, lat]
        assert [[float(x).hex() for x in pair] for pair in net.bounds] == case["bounds"]
        for key in ("residential", "industrial", "commercial"):
            assert sorted(getattr(b, key + "_nodes")) == case[key], (case["name"], key)
            assert sorted(getattr(b, "connected_" + key + "_nodes")) == case["connected_" + key], (case["name"], key)
        # the array view the device consumes carries the same numbers
        flat = net.flat()
        assert [float(x).hex() for x in flat.length] == [s[2] for s in case["streets"]]
        assert flat.max_speed.tolist() == [s[3] for s in case["streets"]]


def test_visualization_draws_what_the_reference_draws(tmp_path):
    """f4 pinned where it is meaningful: tests/golden/visualization.json records every street line the reference's
    own visualization.py hands to ImageDraw (end points as fp64 hex, colour) for a small network and two days of
    loads; this package's Visualization must draw the same lines in the same colours (the device-free modes here,
    the two speed modes in tests/test_gpu_api.py)."""
    from tests.helpers import visualization_lines
    runs = visualization_lines("visualization.json", tmp_path, ("TRAFFIC_LOAD", "MAX_SPEED"))
    assert len(runs) == 4
    for case, zoom, steps in runs:
        assert zoom == case["zoom"]
        assert len(steps) == len(case["lines_per_step"]) == 2
        for got, want in zip(steps, case["lines_per_step"]):
            key = lambda rec: tuple(rec[0])
            assert sorted(got, key=key) == sorted(want, key=key), (case["mode"], case["color_mode"])


def test_visualization_modes_without_gpu(tmp_path):
    """The offline heat maps read the .s4mpi files and write one PNG per day (no device needed for the
    load / speed-limit / component modes)."""
    from array import array as _array
    from PIL import Image
    from streets4mpi_b200 import persistence, synth
    from streets4mpi_b200.osmdata import GraphBuilder
    from streets4mpi_b200.visualization import Visualization
    osm = str(tmp_path / "t.osm")
    synth.write_osm_xml(osm, rows=8, cols=9, seed=1)
    net = GraphBuilder(osm).build_street_network()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        persistence.persist_write("street_network_1.s4mpi", net)
        rng = np.random.default_rng(0)
        for day in (1, 2):
            persistence.persist_write("traffic_load_%d.s4mpi" % day,
                                      _array("I", rng.integers(0, 50, net.street_index).tolist()), is_array=True)
        for mode, cm in (("TRAFFIC_LOAD", "HEATMAP"), ("MAX_SPEED", "MONOCHROME"), ("COMPONENTS", "HEATMAP")):
            out = Visualization("^street_network_[0-9]+.s4mpi$", "^traffic_load_[0-9]+.s4mpi$", mode=mode,
                                color_mode=cm, verbose=False).visualize()
            assert out == ["traffic_load_1.png", "traffic_load_2.png"]
            img = Image.open(out[0])
            assert img.size[0] > 500 and img.size[1] > 500 and len(img.getcolors(1 << 20)) > 3
        v = Visualization("x", "y", verbose=False)
        assert v.value_to_color(0.0) == "hsl(260,100%,5%)" and v.value_to_color(1.0) == "hsl(0,100%,50%)"
        assert Visualization("x", "y", color_mode="MONOCHROME", verbose=False).value_to_color(0.5) == (135, 135, 135, 0)
        comps = Visualization.calculate_components(net)
        assert set(comps.values()) == {1}
        with pytest.raises(ValueError):
            Visualization("x", "y", mode="NOPE")
    finally:
        os.chdir(cwd)
