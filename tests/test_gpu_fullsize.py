"""BASELINE-sized checks (cfg3: 1024x1024 grid, V=1,048,576, A=4,190,208).  The oracle cannot route a whole
day at this size in test time, so the day is checked through size-independent properties and a sample of
trees / trips is compared with the oracle's C twin exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg3():
    from streets4mpi_b200 import engine, synth
    from streets4mpi_b200.settings import settings
    dev = engine.default_device()
    arr = synth.grid_arrays(1024, 1024, seed=1)
    net = synth.network_from(arr)
    flat = net.flat()
    g = engine.DeviceGraph(dev, flat.V, flat.u, flat.v, flat.length, flat.max_speed, node_rank=flat.locality_rank())
    cats = synth.node_categories(arr["node_ids"], seed=1, goal_fraction=0.02)
    return dict(dev=dev, arr=arr, flat=flat, g=g, cats=cats, phys=engine.make_physics(settings))


def test_fullsize_trees_bit_exact(cfg3):
    """Three full trees on the 1M-node graph: distances bit-equal, predecessors equal."""
    from oracle import ctwin
    arr, g = cfg3["arr"], cfg3["g"]
    rng = np.random.default_rng(11)
    w = rng.uniform(0.1, 13.0, size=len(arr["u"]))
    csr = ctwin.Csr(len(arr["node_ids"]), arr["u"], arr["v"])
    for src in (0, 524800, 1048575):
        dist, pred = g.sssp(w, src)
        d0, p0, tied = ctwin.sssp(csr, w, src)
        assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64))
        assert np.array_equal(pred, p0)
        # symmetry of the undirected graph: d(src, t) == d(t, src) up to the rounding of the reversed sum
        t = int(rng.integers(0, len(dist)))
        back, _ = g.sssp(w, t, want_pred=False)
        assert abs(back[src] - dist[t]) <= 1e-6 * dist[t]          # the north_star's cost tolerance


def test_fullsize_day_properties(cfg3):
    """100k trips on cfg3 (5 % of the day's goals -> ~1000 trees): conservation (sum of loads == hops x
    volume), no stuck / unreachable trips, repeatability, root-mode invariance, and a 60-trip sample routed
    by the oracle equals the GPU's load for exactly those trips."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    dev, arr, g, cats, phys = cfg3["dev"], cfg3["arr"], cfg3["g"], cfg3["cats"], cfg3["phys"]
    goals = cats["goals"][:1000]
    o, gl = synth.trip_arrays(100000, arr["node_ids"], goals, seed=3756917)
    o, gl = o.astype(np.int32), gl.astype(np.int32)
    loads = {}
    for mode in (1, 1, 0):                                   # goal-rooted twice, then a sample origin-rooted
        dev.set_option("root_mode", mode)
        if mode == 0:
            oo, gg = o[:1500], gl[:1500]
        else:
            oo, gg = o, gl
        sim = engine.DeviceSim(g, oo, gg, 0.4, phys)
        sim.step()
        load = sim.get_load()
        c = sim.counters()
        assert c["trips_routed"] == len(oo) and c["trips_unreachable"] == 0 and c["trips_stuck"] == 0
        assert c["fallback_sweeps"] == 0
        assert int(load.astype(np.int64).sum()) == c["hops"] * phys.trip_volume
        loads.setdefault(mode, []).append(load)
        if mode == 0:
            dev.set_option("root_mode", 1)
            sim2 = engine.DeviceSim(g, oo, gg, 0.4, phys)
            sim2.step()
            assert np.array_equal(sim2.get_load(), load)      # origin- vs goal-rooted, unique paths
            sim2.close()
        sim.close()
    dev.set_option("root_mode", 0)
    assert np.array_equal(loads[1][0], loads[1][1])           # repeatable to the bit
    # sample against the oracle
    sample = slice(0, 60)
    csr = ctwin.Csr(len(arr["node_ids"]), arr["u"], arr["v"])
    w = ctwin.weights(arr["length"], arr["max_speed"], np.zeros(len(arr["u"]), dtype=np.uint32), 0.4)
    want, cnt = ctwin.day(csr, w, o[sample], gl[sample], nthreads=8)
    sim = engine.DeviceSim(g, o[sample], gl[sample], 0.4, phys)
    sim.step()
    assert cnt["tied_trips"] == 0 and np.array_equal(sim.get_load(), want)
    sim.close()


def test_road_like_graph_at_scale():
    """cfg4 / cfg5 shape at 1.44 M nodes (street chains joined by sparse rungs, mean degree 2.4: thin waves, narrow CTAs,
    adaptive bucket width, overlapped buckets): two whole trees bit-equal to the oracle's, and a day of 5 k trips from
    144 residential origins bit-equal to the oracle's load."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    from streets4mpi_b200.settings import settings
    dev = engine.default_device()
    phys = engine.make_physics(settings)
    arr = synth.ladder_arrays(1200, 1200, seed=2)
    flat = synth.network_from(arr).flat()
    g = engine.DeviceGraph(dev, flat.V, flat.u, flat.v, flat.length, flat.max_speed, node_rank=flat.locality_rank())
    csr = ctwin.Csr(flat.V, flat.u, flat.v)
    rng = np.random.default_rng(5)
    w = rng.uniform(0.1, 13.0, size=flat.S)
    for src in (0, flat.V // 2 + 77):
        dist, pred = g.sssp(w, src)
        d0, p0, tied = ctwin.sssp(csr, w, src)
        assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64))
        assert np.array_equal(pred, p0)
    cats = synth.node_categories(arr["node_ids"], seed=2, goal_fraction=0.02, residential_fraction=0.0001)
    o, gl = synth.trip_arrays(5000, cats["residential"], cats["goals"], seed=9)
    old = dev.get_option("root_mode")
    try:
        dev.set_option("root_mode", 0)                        # 144 residential origins -> 144 trees
        sim = engine.DeviceSim(g, flat.dense(o), flat.dense(gl), 0.4, phys)
        sim.step()
        c = sim.counters()
        w0 = ctwin.weights(flat.length, flat.max_speed, np.zeros(flat.S, dtype=np.uint32), 0.4)
        want, cnt = ctwin.day(csr, w0, flat.dense(o), flat.dense(gl), nthreads=16)
        assert cnt["tied_trips"] == 0 and c["fallback_sweeps"] == 0 and c["trips_unreachable"] == 0
        assert np.array_equal(sim.get_load(), want)
        sim.close()
    finally:
        dev.set_option("root_mode", old)
    g.close()



def test_cfg4_road_graph_at_5m_nodes():
    """BASELINE.json configs[3] size: the 5 M-node road-like graph with a speed hierarchy (synth.road_arrays, mean degree
    2.6, dead ends and cut-off junctions).  Two whole trees bit-equal to the oracle's (distances and canonical
    predecessors), then a day of 5 k trips from 32 origins bit-equal to the oracle's load, followed by the road
    adaptation and a second day on the changed limits."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    from streets4mpi_b200.settings import settings
    dev = engine.default_device()
    phys = engine.make_physics(settings)
    arr = synth.road_arrays(5_000_000, seed=1)
    flat = synth.network_from(arr).flat()
    assert flat.V >= 5_000_000
    g = engine.DeviceGraph(dev, flat.V, flat.u, flat.v, flat.length, flat.max_speed, node_rank=flat.locality_rank())
    csr = ctwin.Csr(flat.V, flat.u, flat.v)
    rng = np.random.default_rng(5)
    w = ctwin.weights(flat.length, flat.max_speed, rng.integers(0, 400, flat.S).astype(np.uint32), 0.4)
    for src in (flat.V // 2 + 1234, 17):
        dist, pred = g.sssp(w, src)
        d0, p0, tied = ctwin.sssp(csr, w, src)
        assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64))
        assert np.array_equal(pred, p0)
    cats = synth.node_categories(arr["node_ids"], seed=1, goal_fraction=0.02, residential_fraction=32.5 / flat.V)
    o, gl = synth.trip_arrays(5000, cats["residential"], cats["goals"], seed=9)
    od, gd = flat.dense(o), flat.dense(gl)
    old = dev.get_option("root_mode")
    try:
        dev.set_option("root_mode", 0)
        sim = engine.DeviceSim(g, od, gd, 0.4, phys)
        assert sim.n_roots <= 33
        prev = np.zeros(flat.S, dtype=np.uint32)
        ms_now = flat.max_speed.copy()
        for day in range(2):
            if day == 1:
                sim.road_construction(flat.adaptable)
                ms_now, _ = ctwin.road_construction(prev, ms_now, flat.adaptable)
            sim.step()
            want, cnt = ctwin.day(csr, ctwin.weights(flat.length, ms_now, prev, 0.4), od, gd, nthreads=16)
            c = sim.counters()
            assert c["fallback_sweeps"] == 0 and c["trips_stuck"] == 0 and cnt["stuck"] == 0
            assert np.array_equal(sim.get_load(), want), day
            sim.commit_total()
            prev = want
        assert c["trips_unreachable"] == 2 * cnt["unreachable"]
        sim.close()
    finally:
        dev.set_option("root_mode", old)
    g.close()


def test_cluster_shared_groups_at_scale():
    """The large-graph configuration (BASELINE configs[4]: a slot per group is gigabytes, fewer slots than SMs, so the
    CTAs of a thread-block cluster share one group) forced onto a 4 M-node road-like graph: two whole trees bit-equal
    to the oracle's, and a day of 4 k trips from 60 residential origins bit-equal to the oracle's load."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    from streets4mpi_b200.settings import settings
    dev = engine.default_device()
    phys = engine.make_physics(settings)
    arr = synth.ladder_arrays(2000, 2000, seed=3)
    flat = synth.network_from(arr).flat()
    keys = ("cluster", "block_threads", "group_words", "root_mode")
    old = {k: dev.get_option(k) for k in keys}
    try:
        dev.set_option("group_words", 2)
        dev.set_option("block_threads", 1024)
        dev.set_option("cluster", 4)
        dev.set_option("root_mode", 0)
        g = engine.DeviceGraph(dev, flat.V, flat.u, flat.v, flat.length, flat.max_speed, node_rank=flat.locality_rank())
        csr = ctwin.Csr(flat.V, flat.u, flat.v)
        rng = np.random.default_rng(7)
        w = rng.uniform(0.1, 13.0, size=flat.S)
        for src in (3, flat.V // 2 + 999):
            dist, pred = g.sssp(w, src)
            d0, p0, _ = ctwin.sssp(csr, w, src)
            assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64))
            assert np.array_equal(pred, p0)
        cats = synth.node_categories(arr["node_ids"], seed=3, goal_fraction=0.02, residential_fraction=60.5 / flat.V)
        o, gl = synth.trip_arrays(4000, cats["residential"], cats["goals"], seed=11)
        sim = engine.DeviceSim(g, flat.dense(o), flat.dense(gl), 0.4, phys)
        sim.step()
        c = sim.counters()
        w0 = ctwin.weights(flat.length, flat.max_speed, np.zeros(flat.S, dtype=np.uint32), 0.4)
        want, cnt = ctwin.day(csr, w0, flat.dense(o), flat.dense(gl), nthreads=16)
        assert cnt["tied_trips"] == 0 and c["trips_unreachable"] == cnt["unreachable"] and c["hops"] == cnt["hops"]
        assert np.array_equal(sim.get_load(), want)
        sim.close()
        g.close()
    finally:
        for k, val in old.items():
            dev.set_option(k, val)
