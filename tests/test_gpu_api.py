"""The reference-facing Python API on the GPU: StreetNetwork / Simulation / TripGenerator /
persistence used the way streets4mpi.py uses them, against the golden fixtures."""
import os
from array import array

import numpy as np
import pytest

from tests.helpers import fh, load_golden

pytestmark = pytest.mark.gpu


def _network(case):
    from streets4mpi_b200 import StreetNetwork
    net = StreetNetwork()
    for n in case["nodes"]:
        net.add_node(n, 0.0, 0.0)
    for (u, v, L, ms) in case["streets"]:
        net.add_street((u, v), fh(L), ms)
    return net


def _trips(pairs):
    d = {}
    for o, g in pairs:
        d.setdefault(o, []).append(g)
    return d


def test_reference_driver_protocol_golden():
    """Exactly the host-array protocol of streets4mpi.py:84-106: step(), read traffic_load,
    'Allreduce' on the host, assign traffic_load / cumulative_traffic_load back, road_construction."""
    from streets4mpi_b200 import Simulation, merge_arrays
    from streets4mpi_b200.settings import settings
    for case in load_golden("days.json")["cases"]:
        settings["trip_volume"] = case["trip_volume"]
        try:
            sims = [Simulation(_network(case), _trips(t), fh(j), lambda *a: None)
                    for t, j in zip(case["trips"], case["jams"])]
            for day, rec in enumerate(case["days"]):
                if day > 0 and day % case["construction_every"] == 0:
                    for s in sims:
                        s.road_construction()
                        assert s.cumulative_traffic_load is None
                    seen = dict((idx, ms) for (_st, idx, _l, ms) in sims[0].street_network)
                    assert [seen[i] for i in range(len(seen))] == rec["max_speed_after_construction"]
                for s in sims:
                    s.step()
                for r, s in enumerate(sims):
                    assert isinstance(s.traffic_load, array) and s.traffic_load.typecode == "I"
                    assert list(s.traffic_load) == rec["local_loads"][r], (case["name"], day, r)
                total = merge_arrays([s.traffic_load for s in sims])
                for s in sims:
                    s.traffic_load = total
                    s.cumulative_traffic_load = merge_arrays((total, s.cumulative_traffic_load))
                assert list(total) == rec["total"]
                assert list(sims[0].cumulative_traffic_load) == rec["cumulative"]
            sims[0].sync_street_network()
            net = sims[0].street_network
            w = dict((idx, net.get_driving_time(st)) for (st, idx, _l, _ms) in net)
            assert [float(w[i]).hex() for i in range(len(w))] == case["days"][-1]["weights"][0]
        finally:
            settings["trip_volume"] = 1


def test_device_resident_exchange_matches_golden():
    """Same days through Simulation.exchange(): loads never leave the GPU between days."""
    from streets4mpi_b200 import Simulation
    case = [c for c in load_golden("days.json")["cases"] if c["name"] == "grid12x12_3ranks"][0]
    sims = [Simulation(_network(case), _trips(t), fh(j), lambda *a: None) for t, j in zip(case["trips"], case["jams"])]
    for day, rec in enumerate(case["days"]):
        if day > 0 and day % case["construction_every"] == 0:
            for s in sims:
                s.road_construction()
        for s in sims:
            s.step()
        sims[0].exchange(virtual_ranks=sims[1:])
        for s in sims:
            assert list(s.traffic_load) == rec["total"]
            assert list(s.cumulative_traffic_load) == rec["cumulative"]


def test_three_node_fixture_and_shortest_paths():
    """simulation.py:147-155, the only fixture the reference ships (with float lengths)."""
    from streets4mpi_b200 import Simulation, StreetNetwork
    net = StreetNetwork()
    for n in (1, 2, 3):
        net.add_node(n, 0, 0)
    net.add_street((1, 2), 10.0, 50)
    net.add_street((2, 3), 100.0, 140)
    logged = []
    sim = Simulation(net, {1: [3]}, 0.3, lambda *a: logged.append(a))
    for _ in range(10):
        sim.step()
        assert list(sim.traffic_load) == [1, 1]
        sim.exchange()
    assert sim.step_counter == 10 and logged
    sim.sync_street_network()
    assert float(net.get_driving_time((1, 2))).hex() == "0x1.2bd4ad642ee50p-2"
    assert float(net.get_driving_time((3, 2))).hex() == "0x1.76c9d8bd3a9e4p-1"
    paths = net.calculate_shortest_paths(1)
    assert paths == {1: None, 2: 1, 3: 2}
    net.add_node(9, 0, 0)                                  # unreachable node is absent from the result
    assert 9 not in net.calculate_shortest_paths(1)


def test_calculate_driving_speed_function():
    from streets4mpi_b200 import calculate_driving_speed
    assert calculate_driving_speed(100.0, 50, 10) == 34.1525987298185
    assert calculate_driving_speed(3.0, 30, 0) == 0.440908153700972
    assert calculate_driving_speed(250.0, 140, 0) == 140


def test_streets4mpi_driver_end_to_end(tmp_path):
    """The zero-argument driver on a synthetic network, persisting .s4mpi files."""
    from streets4mpi_b200 import persistence
    from streets4mpi_b200.settings import settings
    from streets4mpi_b200.streets4mpi import Streets4MPI
    old = dict(settings)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        settings.update({"osm_file": "synthetic:grid:24x24:3", "logging": None, "max_simulation_steps": 5,
                         "number_of_residents": 300, "steps_between_street_construction": 2})
        app = Streets4MPI()
        files = sorted(os.listdir(tmp_path))
        assert [f for f in files if f.startswith("traffic_load_")] == ["traffic_load_%d.s4mpi" % i for i in range(1, 6)]
        assert "street_network_1.s4mpi" in files and "street_network_3.s4mpi" in files and "street_network_5.s4mpi" in files
        day5 = persistence.persist_read("traffic_load_5.s4mpi", is_array=True)
        assert list(day5) == list(app.simulation.traffic_load) and sum(day5) > 0
        net = persistence.persist_read("street_network_5.s4mpi")
        assert net.street_index == 2 * 24 * 23
    finally:
        os.chdir(cwd)
        settings.clear()
        settings.update(old)


def test_streets4mpi_driver_with_device_trip_sampling(tmp_path):
    """device_settings["sample_trips_on_device"]: the driver's trips come from the GPU's Philox sampler (seed =
    random_seed + 37 * rank); the day loop, persistence and conservation are unchanged."""
    from streets4mpi_b200 import persistence, tripgenerator
    from streets4mpi_b200.settings import device_settings, settings
    from streets4mpi_b200.streets4mpi import Streets4MPI
    old, old_dev = dict(settings), dict(device_settings)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        settings.update({"osm_file": "synthetic:grid:30x20:4", "logging": None, "max_simulation_steps": 3,
                         "number_of_residents": 2000, "steps_between_street_construction": 10})
        device_settings["sample_trips_on_device"] = True
        app = Streets4MPI()
        trips = app.simulation.trips
        assert isinstance(trips, tripgenerator.SampledTrips) and trips.number_of_trips == 2000 and trips.seed == 3756917
        table, host = trips.table(), trips.host_table()
        assert sorted(zip(table.origins.tolist(), table.goals.tolist())) == sorted(zip(host.origins.tolist(), host.goals.tolist()))
        day3 = persistence.persist_read("traffic_load_3.s4mpi", is_array=True)
        assert list(day3) == list(app.simulation.traffic_load) and sum(day3) > 0
    finally:
        os.chdir(cwd)
        settings.clear()
        settings.update(old)
        device_settings.clear()
        device_settings.update(old_dev)


@pytest.mark.parametrize("residents,every", [(100, 10), (3000, 4)])
def test_streets4mpi_on_osm_file_matches_port(tmp_path, residents, every):
    """BASELINE configs[0]/[1] stand-in: the zero-argument driver on an OSM XML file (the reference's
    osm/test.osm is a missing blob; synth.write_osm_xml generates a Stellingen-like one) with settings.py
    defaults, against the oracle's line-by-line port fed with the same trips and jam tolerance."""
    import random
    from oracle import port
    from streets4mpi_b200 import persistence, synth
    from streets4mpi_b200.osmdata import GraphBuilder
    from streets4mpi_b200.settings import device_settings, settings
    from streets4mpi_b200.streets4mpi import Streets4MPI
    osm = str(tmp_path / "test.osm")
    synth.write_osm_xml(osm, rows=14, cols=17, seed=5)
    old, old_dev = dict(settings), dict(device_settings)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        settings.update({"osm_file": osm, "logging": None, "number_of_residents": residents,
                         "steps_between_street_construction": every})
        device_settings["root_mode"] = 0
        app = Streets4MPI()
        days = settings["max_simulation_steps"]
        got = [list(persistence.persist_read("traffic_load_%d.s4mpi" % (d + 1), is_array=True)) for d in range(days)]
        # the same run through the oracle port
        b = GraphBuilder(osm)
        src = b.build_street_network()
        b.find_node_categories()
        def make():
            net = port.Network(sssp="canonical")
            for n in src.get_nodes():
                net.add_node(n, *src.node_coordinates(n))
            for i in range(src.street_index):
                u, v = src.get_street_by_index(i)
                at = src._graph.edge_attributes((u, v))
                net.add_street((u, v), at[1], at[2])
            return net
        random.seed(settings["random_seed"])
        goals = b.connected_commercial_nodes | b.connected_industrial_nodes
        trips = port.generate_trips(residents, src.get_nodes(), goals)
        jam = random.random()
        assert jam == app.simulation.jam_tolerance
        totals, net0 = port.run_days(make, [trips], [jam], days, every)
        for d in range(days):
            assert got[d] == list(totals[d]), "day %d" % (d + 1)
        assert sum(got[-1]) > 0
        seen_gpu = dict((idx, ms) for (_s, idx, _l, ms) in app.simulation.street_network)
        seen_port = dict((idx, ms) for (_s, idx, _l, ms) in net0)
        assert seen_gpu == seen_port
    finally:
        os.chdir(cwd)
        settings.clear(); settings.update(old)
        device_settings.clear(); device_settings.update(old_dev)


def test_visualization_speed_modes(tmp_path):
    """IDEAL_SPEED / ACTUAL_SPEED evaluate calculate_driving_speed for every street on the device."""
    from array import array as _array
    from streets4mpi_b200 import persistence, synth
    from streets4mpi_b200.osmdata import GraphBuilder
    from streets4mpi_b200.visualization import Visualization
    osm = str(tmp_path / "t.osm")
    synth.write_osm_xml(osm, rows=6, cols=7, seed=2)
    net = GraphBuilder(osm).build_street_network()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        persistence.persist_write("street_network_1.s4mpi", net)
        persistence.persist_write("traffic_load_1.s4mpi", _array("I", [3] * net.street_index), is_array=True)
        for mode in ("IDEAL_SPEED", "ACTUAL_SPEED"):
            v = Visualization("^street_network_[0-9]+.s4mpi$", "^traffic_load_[0-9]+.s4mpi$", mode=mode, verbose=False)
            assert v.visualize() == ["traffic_load_1.png"]
            vals = v._street_values(_array("I", [3] * net.street_index), 3)
            assert len(vals) == net.street_index and 0.0 < max(vals) <= 1.0
    finally:
        os.chdir(cwd)


def test_trips_sampled_on_the_device():
    from streets4mpi_b200 import engine as _engine
    from streets4mpi_b200.settings import settings as _settings
    dev, phys = _engine.default_device(), _engine.make_physics(_settings)
    _check_trips_sampled_on_the_device(dev, phys)


def _check_trips_sampled_on_the_device(dev, phys):
    """s4_sim_create_sampled: the device's Philox draws equal the numpy mirror trip for trip, both root modes; a day
    on those trips is bit-equal to the oracle's; the Simulation / SampledTrips wrapper gives the dict view."""
    from oracle import ctwin
    from streets4mpi_b200 import Simulation, TripGenerator, engine, synth, tripgenerator
    arr = synth.grid_arrays(40, 45, seed=8)
    net = synth.network_from(arr)
    flat = net.flat()
    cats = synth.node_categories(arr["node_ids"], seed=8, goal_fraction=0.03)
    origins, goals = np.sort(arr["node_ids"][::3]), cats["goals"]
    g = engine.DeviceGraph(dev, flat.V, flat.u, flat.v, flat.length, flat.max_speed, node_rank=flat.locality_rank())
    T, seed = 5000, 3756917 + 37
    io, ig = tripgenerator.philox_trip_indices(T, len(origins), len(goals), seed)
    want = sorted(zip(flat.dense(origins)[io].tolist(), flat.dense(goals)[ig].tolist()))
    old = dev.get_option("root_mode")
    try:
        for mode in (0, 1, 2):
            dev.set_option("root_mode", mode)
            sim = engine.DeviceSim.sampled(g, T, flat.dense(origins), flat.dense(goals), seed, 0.3, phys)
            o, gl = sim.get_trips()
            assert sorted(zip(o.tolist(), gl.tolist())) == want, mode
            if mode == 2:
                assert sim.root_mode == 1                                    # fewer goals than origins
            sim.step()
            csr = ctwin.Csr(flat.V, flat.u, flat.v)
            w = ctwin.weights(flat.length, flat.max_speed, np.zeros(flat.S, dtype=np.uint32), 0.3)
            load, cnt = ctwin.day(csr, w, o, gl, nthreads=4)
            assert cnt["tied_trips"] == 0 and np.array_equal(sim.get_load(), load), mode
            sim.close()
    finally:
        dev.set_option("root_mode", old)
    g.close()
    st = TripGenerator().sample_trips_on_device(T, origins, goals, seed)
    s2 = Simulation(net, st, 0.3, lambda *a: None, device=dev)
    table = st.table()
    host = st.host_table()
    assert sorted(zip(table.origins.tolist(), table.goals.tolist())) == sorted(zip(host.origins.tolist(), host.goals.tolist()))
    some = table.keys()[0]
    assert some in st and len(st[some]) >= 1 and len(st) == len(table.keys())
    s2.step()
    assert int(np.frombuffer(s2.traffic_load, dtype=np.uint32).astype(np.int64).sum()) == s2.counters()["hops"]



def test_stuck_trips_are_surfaced():
    """a walk that cannot find its way off a zero-weight plateau (more than 32 nodes here) is counted,
    the oracle counts the same trip, and the Simulation warns as soon as the host looks at the load."""
    from oracle import ctwin
    from streets4mpi_b200 import Simulation, StreetNetwork
    n = 41                                                # node 0 = root side, 1..40 a chain of zero-length streets
    net = StreetNetwork()
    for i in range(n + 1):
        net.add_node(i, float(i), 0.0)
    net.add_street((0, 1), 100.0, 50)
    for i in range(1, n):
        net.add_street((i, i + 1), 0.0, 50)
    sim = Simulation(net, {0: [n, 5]}, 0.5, lambda *a: None)
    sim.step()
    with pytest.warns(RuntimeWarning, match="could not be walked back"):
        load = np.frombuffer(sim.traffic_load, dtype=np.uint32)
    flat = net.flat()
    csr = ctwin.Csr(flat.V, flat.u, flat.v)
    w = ctwin.weights(flat.length, flat.max_speed, np.zeros(flat.S, dtype=np.uint32), 0.5)
    want, cnt = ctwin.day(csr, w, flat.dense([0, 0]), flat.dense([n, 5]))
    assert cnt["stuck"] == 1 and sim.counters()["trips_stuck"] == 1
    assert np.array_equal(load, want) and load[0] == 1     # the trip to node 5 is 4 plateau nodes from its exit


def test_visualization_speed_modes_draw_what_the_reference_draws(tmp_path):
    """The two speed modes of the heat maps evaluate calculate_driving_speed (simulation.py:129-138) per street -- on the
    device here: same lines, same colours as the reference's visualization.py drew (tests/golden/visualization.json)."""
    from tests.helpers import visualization_lines
    runs = visualization_lines("visualization.json", tmp_path, ("IDEAL_SPEED", "ACTUAL_SPEED"))
    assert len(runs) == 4
    for case, zoom, steps in runs:
        assert zoom == case["zoom"] and len(steps) == 2
        for got, want in zip(steps, case["lines_per_step"]):
            key = lambda rec: tuple(rec[0])
            assert sorted(got, key=key) == sorted(want, key=key), (case["mode"], case["color_mode"])
