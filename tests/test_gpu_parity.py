"""Parity of the CUDA path (through the C-ABI) with the oracle and with the golden
fixtures generated from the reference's own source.  Integer / index results must be
bit-exact; fp64 weights and distances must be bit-exact too (same operation order,
no FMA), which is stronger than the 1e-6 relative tolerance BASELINE.json allows."""
import numpy as np
import pytest

from tests.helpers import case_arrays, fh, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from streets4mpi_b200 import engine
    return engine.default_device()


@pytest.fixture(scope="module")
def phys():
    from streets4mpi_b200 import engine
    from streets4mpi_b200.settings import settings
    return engine.make_physics(settings)


def test_library_is_the_cuda_path(dev):
    info = dev.info()
    assert info["sm_count"] > 0 and "NVIDIA" in info["name"]


def test_k1_golden_bit_exact(dev, phys):
    rows = load_golden("k1_weights.json")["rows"]
    L = np.array([fh(r["length"]) for r in rows])
    ms = np.array([r["max_speed"] for r in rows], dtype=np.int32)
    ld = np.array([r["load"] for r in rows], dtype=np.uint32)
    got = dev.driving_speed(L, ms, ld, phys)
    assert [float(x).hex() for x in got] == [r["actual"] for r in rows]
    got0 = dev.driving_speed(L, ms, np.zeros_like(ld), phys)
    assert [float(x).hex() for x in got0] == [r["ideal"] for r in rows]
    # driving times: one street per row, each row its own jam tolerance -> one sim per jam is too slow;
    # group rows by jam instead (fixed rows share a few values, random rows are unique): use a 2-node graph per row
    from streets4mpi_b200 import engine
    n = len(rows)
    u = np.arange(n, dtype=np.int32) * 2
    v = u + 1
    ok = L >= 0
    g = engine.DeviceGraph(dev, 2 * n, u, v, L, ms)
    for jam in sorted(set(r["jam"] for r in rows))[:40]:
        sim = engine.DeviceSim(g, np.zeros(0, np.int32), np.zeros(0, np.int32), fh(jam), phys)
        sim.set_load(ld)
        w = sim.weights(refresh=True)
        for i, r in enumerate(rows):
            if r["jam"] == jam:
                assert float(w[i]).hex() == r["time"], (i, r)
        sim.close()
    assert ok.all()


def test_k1_random_vs_oracle(dev, phys):
    from oracle import ctwin
    from streets4mpi_b200 import engine
    rng = np.random.default_rng(7)
    n = 200000
    L = np.concatenate([rng.uniform(0.0, 30.0, n // 2), rng.uniform(20.0, 5000.0, n - n // 2)])
    ms = rng.integers(1, 141, n).astype(np.int32)
    ld = rng.choice([0, 1, 2, 10, 1000, 2 ** 32 - 1], size=n).astype(np.uint32)
    u = np.arange(n, dtype=np.int32)
    v = (u + 1) % n
    g = engine.DeviceGraph(dev, n, u, v, L, ms)
    for jam in (0.0, 0.123456789, 0.999999):
        sim = engine.DeviceSim(g, np.zeros(0, np.int32), np.zeros(0, np.int32), jam, phys)
        sim.set_load(ld)
        got = sim.weights(refresh=True)
        want = ctwin.weights(L, ms, ld, jam)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
        sim.close()


def _run_case_on_gpu(case, dev, phys, root_mode=0):
    from streets4mpi_b200 import engine
    a = case_arrays(case)
    dev.set_option("root_mode", root_mode)
    g = engine.DeviceGraph(dev, a["V"], a["u"], a["v"], a["length"], a["max_speed"])
    sims = [engine.DeviceSim(g, o, gl, jam, engine.Physics(phys.car_length, phys.min_breaking_distance,
                                                           phys.braking_deceleration, case["trip_volume"], 0))
            for (o, gl), jam in zip(a["trips"], a["jams"])]
    dev.set_option("root_mode", 0)
    for day, rec in enumerate(case["days"]):
        if day > 0 and day % case["construction_every"] == 0:
            for s in sims:
                s.road_construction(a["adaptable"])
            vis, ins = sims[0].get_max_speed()
            assert vis.tolist() == rec["max_speed_after_construction"], (case["name"], day)
        for s in sims:
            s.step()
        for r, s in enumerate(sims):
            assert s.get_load().tolist() == rec["local_loads"][r], (case["name"], day, r, root_mode)
            w = s.weights()
            assert [float(x).hex() for x in w] == rec["weights"][r], (case["name"], day, r)
        for s in sims[1:]:
            sims[0].load_add(s)
        for s in sims[1:]:
            s.load_copy(sims[0])
        for s in sims:
            s.commit_total()
        assert sims[0].get_load().tolist() == rec["total"]
        assert sims[-1].get_cumulative().tolist() == rec["cumulative"]
    c = sims[0].counters()
    assert c["fallback_sweeps"] == 0 and c["trips_stuck"] == 0
    for s in sims:
        s.close()
    g.close()


@pytest.mark.parametrize("root_mode", [0, 1, 2])
def test_days_golden(dev, phys, root_mode):
    """Reference-generated fixtures: weights, local loads, totals, cumulative loads and the
    adapted speed limits, day by day, several virtual ranks.  All trees are tie-free, so
    goal-rooted trees (root_mode 1/2) must give the same loads."""
    for case in load_golden("days.json")["cases"]:
        _run_case_on_gpu(case, dev, phys, root_mode)


def test_road_construction_golden(dev, phys):
    from streets4mpi_b200 import engine
    for c in load_golden("road_construction.json")["cases"]:
        S = c["S"]
        u = np.array([s[0] for s in c["streets"]], dtype=np.int32) - 1
        v = np.array([s[1] for s in c["streets"]], dtype=np.int32) - 1
        ms = np.array([s[2] for s in c["streets"]], dtype=np.int32)
        g = engine.DeviceGraph(dev, S + 1, u, v, np.full(S, 50.0), ms)
        sim = engine.DeviceSim(g, np.zeros(0, np.int32), np.zeros(0, np.int32), 0.5, phys)
        with pytest.raises(Exception):
            sim.road_construction(None)                       # cumulative is None
        sim.set_cumulative(np.array(c["cumulative"], dtype=np.uint32))
        sim.road_construction((u <= v).astype(np.uint8))
        vis, ins = sim.get_max_speed()
        assert vis.tolist() == c["max_speed_seen_by_weight_loop"], S
        assert ins.tolist() == c["max_speed_insertion_orientation"], S
        assert sim.get_cumulative() is None
        sim.close()
        g.close()


@pytest.mark.parametrize("shape", ["grid", "planar"])
@pytest.mark.parametrize("opts", [dict(), dict(block_threads=256, ctas_per_sm=4, delta_factor=0.5),
                                  dict(block_threads=1024, delta_factor=8.0, near_smem_entries=64),
                                  dict(delta_factor=0.0), dict(preread=1), dict(preread=1, near_smem_entries=64),
                                  dict(block_threads=32, near_smem_entries=32, delta_factor=1.0),
                                  # overlapped buckets: off, and so eager that nearly every advance finds entries pending
                                  dict(early_advance=0), dict(early_advance=64, delta_factor=0.5),
                                  dict(early_advance=64, block_threads=32, near_smem_entries=32),
                                  # thread-block clusters sharing one group (the large-graph configuration)
                                  dict(cluster=2, block_threads=512), dict(cluster=8, block_threads=1024),
                                  dict(cluster=4, block_threads=1024, delta_factor=0.5, early_advance=64),
                                  # other group widths: 7 and 31 trees per group (rows of 64 / 256 bytes), 3 per group on 4 lanes
                                  dict(group_words=1), dict(group_words=2), dict(group_words=4), dict(group_words=4, block_threads=512),
                                  dict(group_lanes=4, group_words=1)])
@pytest.mark.parametrize("ranked", [False, True])
def test_sssp_dist_and_pred_vs_oracle(dev, shape, opts, ranked):
    """Single trees: distances bit-equal to Dijkstra, predecessors equal under the canonical rule;
    kernel variants and bucket widths (incl. the tiny shared-memory queue that forces spills)."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    arr = synth.grid_arrays(37, 53, seed=3) if shape == "grid" else synth.planar_arrays(3000, seed=4)
    V = len(arr["node_ids"])
    rng = np.random.default_rng(1)
    w = rng.uniform(0.05, 13.0, size=len(arr["u"]))
    rank = synth.network_from(arr).flat().locality_rank() if ranked else None      # device numbering along a Z curve
    assert (rank is not None) == ranked
    g = engine.DeviceGraph(dev, V, arr["u"], arr["v"], arr["length"], arr["max_speed"], node_rank=rank)
    csr = ctwin.Csr(V, arr["u"], arr["v"])
    defaults = {k: dev.get_option(k) for k in opts}
    try:
        for k, val in opts.items():
            dev.set_option(k, val)
        for src in (0, V // 3, V - 1):
            dist, pred = g.sssp(w, src)
            d0, p0, tied = ctwin.sssp(csr, w, src)
            assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64))
            assert np.array_equal(pred, p0)
    finally:
        for k, val in defaults.items():
            dev.set_option(k, val)
    g.close()


def test_ties_zero_weights_and_components(dev):
    """Integer weights (massive ties), zero-weight streets, self loops, isolated nodes and a
    detached component: distances and canonical predecessors still equal the oracle's."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    arr = synth.grid_arrays(20, 20, seed=9)
    V = 400 + 5
    u = np.concatenate([arr["u"], [400, 401, 7]]).astype(np.int32)       # 400-401 detached pair, 401 self loop, 7 self loop
    v = np.concatenate([arr["v"], [401, 401, 7]]).astype(np.int32)
    S = len(u)
    rng = np.random.default_rng(2)
    for k, w in enumerate((np.ones(S), rng.integers(0, 3, S).astype(np.float64), rng.integers(1, 4, S).astype(np.float64),
                           np.ones(S), rng.integers(0, 3, S).astype(np.float64))):
        # the last two run on a shuffled device numbering: ties must still break on the caller's ids
        rank = rng.permutation(V).astype(np.int32) if k >= 3 else None
        g = engine.DeviceGraph(dev, V, u, v, np.full(S, 10.0), np.full(S, 50, dtype=np.int32), node_rank=rank)
        csr = ctwin.Csr(V, u, v)
        for src in (0, 210, 400, 404):
            dist, pred = g.sssp(w, src)
            d0, p0, _ = ctwin.sssp(csr, w, src)
            assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64))
            assert np.array_equal(pred, p0), (src,)
        g.close()


@pytest.mark.parametrize("ranked", [False, True])
def test_zero_weight_plateaus_match_oracle(dev, phys, ranked):
    """Zero-length streets (two OSM nodes on one coordinate give haversine 0): the plateau rule keeps predecessor
    chains acyclic (R-C w=1, C-A w=0, A-B w=0, id(B) < id(A) < id(C) must load [1, 1, 1]), and a day on a
    grid with 25 % zero-length streets equals the oracle's, stuck counter included."""
    from oracle import ctwin
    from streets4mpi_b200 import engine
    B, A, Cn, R = 0, 1, 2, 3
    u = np.array([R, Cn, A], dtype=np.int32)
    v = np.array([Cn, A, B], dtype=np.int32)
    g = engine.DeviceGraph(dev, 4, u, v, np.array([100.0, 0.0, 0.0]), np.full(3, 50, dtype=np.int32),
                           node_rank=np.array([2, 0, 3, 1], dtype=np.int32) if ranked else None)
    dist, pred = g.sssp(np.array([1.0, 0.0, 0.0]), R)
    assert dist.tolist() == [1.0, 1.0, 1.0, 0.0] and pred.tolist() == [A, Cn, R, -1]
    for mode in (0, 1):
        old = dev.get_option("root_mode")
        try:
            dev.set_option("root_mode", mode)
            sim = engine.DeviceSim(g, np.array([R, B], dtype=np.int32), np.array([B, R], dtype=np.int32), 0.5, phys)
            sim.step()
            assert sim.get_load().tolist() == [2, 2, 2]
            assert sim.counters()["trips_stuck"] == 0
            sim.close()
        finally:
            dev.set_option("root_mode", old)
    g.close()
    rng = np.random.default_rng(21)
    rows, cols = 30, 34
    V = rows * cols
    idx = np.arange(V).reshape(rows, cols)
    u = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()]).astype(np.int32)
    v = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()]).astype(np.int32)
    S = len(u)
    L = np.where(rng.random(S) < 0.25, 0.0, rng.integers(1, 4, S) * 50.0)
    ms = np.full(S, 50, dtype=np.int32)
    o = rng.integers(0, V, 4000).astype(np.int32)
    gl = rng.integers(0, V, 4000).astype(np.int32)
    rank = rng.permutation(V).astype(np.int32) if ranked else None
    g = engine.DeviceGraph(dev, V, u, v, L, ms, node_rank=rank)
    sim = engine.DeviceSim(g, o, gl, 0.5, phys)
    csr = ctwin.Csr(V, u, v)
    prev = np.zeros(S, dtype=np.uint32)
    stuck = 0
    for _ in range(2):
        sim.step()
        w = ctwin.weights(L, ms, prev, 0.5)
        want, cnt = ctwin.day(csr, w, o, gl, nthreads=4)
        assert np.array_equal(sim.get_load(), want)
        stuck += cnt["stuck"]
        sim.commit_total()
        prev = want
    assert sim.counters()["trips_stuck"] == stuck
    for src in (0, V // 3):
        dist, pred = g.sssp(w, src)
        d0, p0, _ = ctwin.sssp(csr, w, src)
        assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64)) and np.array_equal(pred, p0)


@pytest.mark.parametrize("ranked", [False, True])
def test_day_with_ties_matches_oracle(dev, phys, ranked):
    """Equal lengths and speeds on a grid: every trip has many shortest paths; the CUDA walk
    and the oracle apply the same lowest-predecessor rule, so loads still agree bit for bit."""
    from oracle import ctwin
    from streets4mpi_b200 import engine
    rows = cols = 24
    idx = np.arange(rows * cols).reshape(rows, cols)
    u = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()]).astype(np.int32)
    v = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()]).astype(np.int32)
    S = len(u)
    L = np.full(S, 100.0)
    ms = np.full(S, 50, dtype=np.int32)
    rng = np.random.default_rng(3)
    o = rng.integers(0, rows * cols, 3000).astype(np.int32)
    gl = rng.integers(0, rows * cols, 3000).astype(np.int32)
    rank = rng.permutation(rows * cols).astype(np.int32) if ranked else None
    g = engine.DeviceGraph(dev, rows * cols, u, v, L, ms, node_rank=rank)
    sim = engine.DeviceSim(g, o, gl, 0.5, phys)
    csr = ctwin.Csr(rows * cols, u, v)
    prev = np.zeros(S, dtype=np.uint32)
    for _ in range(3):
        sim.step()
        w = ctwin.weights(L, ms, prev, 0.5)
        want, cnt = ctwin.day(csr, w, o, gl, nthreads=4)
        assert np.array_equal(sim.get_load(), want)
        sim.commit_total()
        prev = want
    assert cnt["tied_trips"] > 0
    sim.close()
    g.close()


@pytest.mark.parametrize("shape,n_trips", [("grid", 30000), ("planar", 30000)])
def test_medium_day_vs_oracle(dev, phys, shape, n_trips):
    """~10^5-node graphs, 3 days with congestion feedback and one adaptation, against the C twin."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    arr = synth.grid_arrays(300, 333, seed=21) if shape == "grid" else synth.planar_arrays(100000, seed=22)
    V = len(arr["node_ids"])
    cats = synth.node_categories(arr["node_ids"], seed=5)
    o, gl = synth.trip_arrays(n_trips, arr["node_ids"], cats["goals"], seed=6)
    o, gl = o.astype(np.int32), gl.astype(np.int32)
    dev.set_option("root_mode", 2)                      # goal-rooted: ~2000 trees instead of ~26000
    g = engine.DeviceGraph(dev, V, arr["u"], arr["v"], arr["length"], arr["max_speed"])
    sim = engine.DeviceSim(g, o, gl, 0.25, phys)
    dev.set_option("root_mode", 0)
    assert sim.root_mode == 1
    csr = ctwin.Csr(V, arr["u"], arr["v"])
    ms = arr["max_speed"].copy()
    prev = np.zeros(len(ms), dtype=np.uint32)
    cum = None
    hops = 0
    for day in range(3):
        if day == 2:
            sim.road_construction(None)
            ms, _ = ctwin.road_construction(cum, ms, None)
            assert np.array_equal(sim.get_max_speed()[0], ms)
            cum = None
        sim.step()
        w = ctwin.weights(arr["length"], ms, prev, 0.25)
        want, cnt = ctwin.day(csr, w, gl, o, nthreads=8)      # oracle rooted the same way
        got = sim.get_load()
        assert cnt["stuck"] == 0
        if cnt["tied_trips"] == 0:
            assert np.array_equal(got, want), (shape, day)
        else:                                                   # ties: same rule on both sides anyway
            assert np.array_equal(got, want), (shape, day, "tied")
        sim.commit_total()
        prev = want
        hops += int(want.astype(np.int64).sum())
        cum = want.copy() if cum is None else cum + want
    c = sim.counters()
    assert c["trips_routed"] == 3 * n_trips and c["fallback_sweeps"] == 0
    assert c["hops"] == hops
    sim.close()
    g.close()


def test_small_pool_ping_pong_batches(dev, phys):
    """A pool of 64 trees forces ~50 launches per day that alternate between the two pool halves while
    the walks run on the second stream; the day must still equal the oracle's."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    arr = synth.grid_arrays(60, 70, seed=41)
    V = len(arr["node_ids"])
    rng = np.random.default_rng(9)
    o = rng.integers(0, V, 2500).astype(np.int32)
    gl = rng.integers(0, V, 2500).astype(np.int32)
    old = dev.get_option("batch_roots")
    try:
        dev.set_option("batch_roots", 64)
        dev.set_option("root_mode", 0)
        g = engine.DeviceGraph(dev, V, arr["u"], arr["v"], arr["length"], arr["max_speed"])
        sim = engine.DeviceSim(g, o, gl, 0.3, phys)
        csr = ctwin.Csr(V, arr["u"], arr["v"])
        prev = np.zeros(len(arr["u"]), dtype=np.uint32)
        for _ in range(2):
            sim.step()
            want, cnt = ctwin.day(csr, ctwin.weights(arr["length"], arr["max_speed"], prev, 0.3), o, gl, nthreads=4)
            assert np.array_equal(sim.get_load(), want)
            sim.commit_total()
            prev = want
        c = sim.counters()
        assert c["sssp_launches"] >= 2 * (sim.n_roots // 32) and c["hops"] > 0
        sim.close()
        g.close()
    finally:
        dev.set_option("batch_roots", old)


def test_root_mode_invariance_unique_paths(dev, phys):
    """Origin-rooted (reference grouping) and goal-rooted trees give identical loads on continuous weights."""
    from streets4mpi_b200 import engine, synth
    arr = synth.grid_arrays(90, 110, seed=31)
    cats = synth.node_categories(arr["node_ids"], seed=2, goal_fraction=0.01)
    o, gl = synth.trip_arrays(5000, arr["node_ids"], cats["goals"], seed=8)
    g = engine.DeviceGraph(dev, len(arr["node_ids"]), arr["u"], arr["v"], arr["length"], arr["max_speed"])
    loads = []
    for mode in (0, 1):
        dev.set_option("root_mode", mode)
        sim = engine.DeviceSim(g, o.astype(np.int32), gl.astype(np.int32), 0.6, phys)
        sim.step()
        loads.append(sim.get_load())
        sim.close()
    dev.set_option("root_mode", 0)
    assert np.array_equal(loads[0], loads[1])
    g.close()


@pytest.mark.parametrize("preread", [1, 2])
def test_queue_overflow_falls_back_to_sweeps(dev, phys, preread):
    """Starve the frontier queues: entries are dropped, the Bellman-Ford safety net must
    still deliver exact distances."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    arr = synth.grid_arrays(64, 64, seed=5)
    V = 64 * 64
    rng = np.random.default_rng(4)
    w = rng.uniform(0.1, 10.0, len(arr["u"]))
    old = {k: dev.get_option(k) for k in ("near_smem_entries", "queue_entries", "delta_factor", "preread")}
    try:
        dev.set_option("preread", preread)
        dev.set_option("near_smem_entries", 64)
        dev.set_option("queue_entries", 16)
        dev.set_option("delta_factor", 50.0)
        g = engine.DeviceGraph(dev, V, arr["u"], arr["v"], arr["length"], arr["max_speed"])
        dist, pred = g.sssp(w, 5)
        d0, p0, _ = ctwin.sssp(ctwin.Csr(V, arr["u"], arr["v"]), w, 5)
        assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64))
        assert np.array_equal(pred, p0)
        g.close()
    finally:
        for k, val in old.items():
            dev.set_option(k, val)


@pytest.mark.parametrize("words", [1, 2, 4])
@pytest.mark.parametrize("case", ["unique", "ties", "plateaus", "wide_nodes"])
def test_walk_with_predecessor_codes_matches_oracle(dev, phys, words, case):
    """walk_codes=1: a streaming pass writes the canonical predecessor of every (node, tree) as a 2-bit arc-slot code, one
    lane per trip follows the codes, and ties / zero-weight plateaus / nodes with more than four arcs fall back to the
    full rule from that node on.  Loads and counters must equal the oracle's on unique shortest paths, on a grid of equal
    lengths (nearly every node tied), with 25 % zero-length streets, and on a planar graph with wide nodes."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    rng = np.random.default_rng(31)
    if case == "wide_nodes":
        arr = synth.planar_arrays(4000, seed=6)
        V, u, v, L, ms = len(arr["node_ids"]), arr["u"], arr["v"], arr["length"], arr["max_speed"]
    else:
        rows, cols = 40, 50
        V = rows * cols
        idx = np.arange(V).reshape(rows, cols)
        u = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()]).astype(np.int32)
        v = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()]).astype(np.int32)
        S = len(u)
        ms = np.full(S, 50, dtype=np.int32)
        if case == "unique":
            L = rng.uniform(20.0, 400.0, S)
        elif case == "ties":
            L = np.full(S, 100.0)
        else:
            L = np.where(rng.random(S) < 0.25, 0.0, rng.integers(1, 4, S) * 50.0)
    o = rng.integers(0, V, 5000).astype(np.int32)
    gl = rng.choice(V, 300, replace=False)[rng.integers(0, 300, 5000)].astype(np.int32)
    keys = ("group_words", "walk_codes", "root_mode")
    old = {k: dev.get_option(k) for k in keys}
    try:
        dev.set_option("group_words", words)
        dev.set_option("walk_codes", 1)
        dev.set_option("root_mode", 1)
        g = engine.DeviceGraph(dev, V, u, v, L, ms, node_rank=rng.permutation(V).astype(np.int32))
        sim = engine.DeviceSim(g, o, gl, 0.4, phys)
        csr = ctwin.Csr(V, u, v)
        prev = np.zeros(len(u), dtype=np.uint32)
        hops = stuck = unreach = 0
        for _ in range(2):
            sim.step()
            want, cnt = ctwin.day(csr, ctwin.weights(L, ms, prev, 0.4), gl, o, nthreads=4)
            assert np.array_equal(sim.get_load(), want)
            hops, stuck, unreach = hops + cnt["hops"], stuck + cnt["stuck"], unreach + cnt["unreachable"]
            sim.commit_total()
            prev = want
        c = sim.counters()
        assert (c["hops"], c["trips_stuck"], c["trips_unreachable"], c["trips_routed"]) == (hops, stuck, unreach, 10000)
        sim.close()
        g.close()
    finally:
        for k, val in old.items():
            dev.set_option(k, val)


@pytest.mark.parametrize("words,block", [(1, 0), (2, 0), (4, 0), (4, 512), (4, 128)])
def test_group_widths_day(dev, phys, words, block):
    """7 and 31 trees per group (rows of 64 and 256 bytes; 31 trees take the neighbour rows two at a time): a day of
    6 k trips to 1500 goals equals the oracle's, on two days with the congestion feedback between them."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    arr = synth.grid_arrays(80, 100, seed=23)
    V = len(arr["node_ids"])
    rng = np.random.default_rng(18)
    o = rng.integers(0, V, 6000).astype(np.int32)
    gl = rng.choice(V, 1500, replace=False)[rng.integers(0, 1500, 6000)].astype(np.int32)
    keys = ("group_words", "block_threads", "root_mode")
    old = {k: dev.get_option(k) for k in keys}
    try:
        dev.set_option("group_words", words)
        dev.set_option("block_threads", block)
        dev.set_option("root_mode", 1)
        g = engine.DeviceGraph(dev, V, arr["u"], arr["v"], arr["length"], arr["max_speed"])
        sim = engine.DeviceSim(g, o, gl, 0.3, phys)
        csr = ctwin.Csr(V, arr["u"], arr["v"])
        prev = np.zeros(len(arr["u"]), dtype=np.uint32)
        for _ in range(2):
            sim.step()
            want, cnt = ctwin.day(csr, ctwin.weights(arr["length"], arr["max_speed"], prev, 0.3), gl, o, nthreads=4)
            assert np.array_equal(sim.get_load(), want)
            sim.commit_total()
            prev = want
        c = sim.counters()
        assert c["trips_routed"] == 12000 and c["trips_stuck"] == 0
        sim.close()
        g.close()
    finally:
        for k, val in old.items():
            dev.set_option(k, val)


@pytest.mark.parametrize("cluster,block", [(2, 512), (4, 1024), (8, 1024)])
def test_cluster_shared_groups_day_and_overflow(dev, phys, cluster, block):
    """The CTAs of a thread-block cluster share one group of trees (frontier queues, far piles and counters in global
    memory, a cluster barrier per round): a day of 4 k trips over ~60 groups equals the oracle's, the counters add up,
    and with starved queues the cluster-wide Bellman-Ford safety net still delivers exact distances."""
    from oracle import ctwin
    from streets4mpi_b200 import engine, synth
    arr = synth.grid_arrays(70, 90, seed=17)
    V = len(arr["node_ids"])
    rng = np.random.default_rng(8)
    o = rng.integers(0, V, 4000).astype(np.int32)
    gl = rng.choice(V, 900, replace=False)[rng.integers(0, 900, 4000)].astype(np.int32)
    keys = ("cluster", "block_threads", "root_mode", "queue_entries", "delta_factor")
    old = {k: dev.get_option(k) for k in keys}
    try:
        dev.set_option("cluster", cluster)
        dev.set_option("block_threads", block)
        dev.set_option("root_mode", 1)
        # room for the duplicates of a tiny graph under 4096 .. 8192 threads (the far pile is not de-duplicated; an
        # overflow would be survivable, but the second half of this test is where it is provoked on purpose)
        dev.set_option("queue_entries", 65536)
        g = engine.DeviceGraph(dev, V, arr["u"], arr["v"], arr["length"], arr["max_speed"])
        sim = engine.DeviceSim(g, o, gl, 0.3, phys)
        csr = ctwin.Csr(V, arr["u"], arr["v"])
        prev = np.zeros(len(arr["u"]), dtype=np.uint32)
        hops = 0
        for _ in range(2):
            sim.step()
            want, cnt = ctwin.day(csr, ctwin.weights(arr["length"], arr["max_speed"], prev, 0.3), gl, o, nthreads=4)
            assert np.array_equal(sim.get_load(), want)
            hops += cnt["hops"]
            sim.commit_total()
            prev = want
        c = sim.counters()
        assert c["hops"] == hops and c["trips_routed"] == 8000
        if cluster == 2:          # 4096 .. 8192 threads on a 6300-node graph re-expand enough to fill any pile; survivable
            assert c["fallback_sweeps"] == 0
        sim.close()
        dev.set_option("queue_entries", 16)
        dev.set_option("delta_factor", 50.0)
        w = rng.uniform(0.1, 10.0, len(arr["u"]))
        dist, pred = g.sssp(w, 77)
        d0, p0, _ = ctwin.sssp(csr, w, 77)
        assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64)) and np.array_equal(pred, p0)
        g.close()
    finally:
        for k, val in old.items():
            dev.set_option(k, val)


def test_options_from_environment(monkeypatch):
    """S4MPI_OPTIONS="key=value,..." (the A/B hook of tools/ab_variants.sh) reaches s4_ctx_set_option."""
    from streets4mpi_b200 import _ffi, engine
    monkeypatch.setenv("S4MPI_OPTIONS", "early_advance=5, delta_factor=2.5")
    d = engine.Device()
    assert d.get_option("early_advance") == 5 and d.get_option("delta_factor") == 2.5
    d.close()
    monkeypatch.setenv("S4MPI_OPTIONS", "no_such_option=1")
    with pytest.raises(_ffi.S4Error):
        engine.Device()


def test_c_abi_error_behaviour(dev, phys):
    from streets4mpi_b200 import _ffi, engine
    with pytest.raises(_ffi.S4Error):
        engine.DeviceGraph(dev, 4, [0, 9], [1, 2], [1.0, 1.0], [50, 50])          # node id out of range
    with pytest.raises(_ffi.S4Error):
        engine.DeviceGraph(dev, 4, [0], [1], [1.0], [0])                           # max_speed 0
    with pytest.raises(_ffi.S4Error):
        engine.DeviceGraph(dev, 4, [0], [1], [1.0], [50], node_rank=[0, 1, 1, 3])  # not a permutation
    g = engine.DeviceGraph(dev, 4, [0, 1], [1, 2], [10.0, 20.0], [50, 50])
    with pytest.raises(_ffi.S4Error):
        engine.DeviceSim(g, [0], [7], 0.1, phys)                                   # goal out of range
    with pytest.raises(_ffi.S4Error):
        dev.set_option("block_threads", 300)
    with pytest.raises(_ffi.S4Error):
        dev.set_option("no_such_option", 1)
    with pytest.raises(Exception):
        dev.set_option("early_advance", 65)
    sim = engine.DeviceSim(g, np.zeros(0, np.int32), np.zeros(0, np.int32), 0.1, phys)   # empty trip list
    sim.step()
    assert sim.get_load().tolist() == [0, 0]
    sim.close()
    g.close()


def test_degenerate_graphs(dev, phys):
    """Single node without streets, a lone self loop, an origin == goal trip, a trip into another component."""
    from streets4mpi_b200 import engine
    g = engine.DeviceGraph(dev, 1, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), np.zeros(0, np.int32))
    sim = engine.DeviceSim(g, [0, 0], [0, 0], 0.5, phys)
    sim.step()
    assert sim.get_load().tolist() == [] and sim.counters()["hops"] == 0
    dist, pred = g.sssp(np.zeros(0), 0)
    assert dist.tolist() == [0.0] and pred.tolist() == [-1]
    sim.close()
    g.close()
    g = engine.DeviceGraph(dev, 4, [1, 2], [1, 3], [5.0, 7.0], [30, 50])           # self loop on 1, street 2-3, node 0 isolated
    sim = engine.DeviceSim(g, [2, 0, 1, 3], [3, 2, 1, 3], 0.1, phys)
    for _ in range(2):
        sim.step()
        assert sim.get_load().tolist() == [0, 1]
        sim.commit_total()
    c = sim.counters()
    assert c["trips_unreachable"] == 2 and c["trips_routed"] == 8 and c["hops"] == 2
    sim.close()
    g.close()


def test_random_small_graphs_property(dev, phys):
    """Hypothesis-driven: arbitrary small multigraphs (parallel streets, self loops, several components,
    integer / zero / continuous weights, shuffled device numbering): distances bit-equal, predecessors equal,
    and a day's loads equal to the oracle's."""
    from hypothesis import HealthCheck, given, settings as hsettings, strategies as st
    from oracle import ctwin
    from streets4mpi_b200 import engine

    @st.composite
    def cases(draw):
        V = draw(st.integers(2, 28))
        S = draw(st.integers(1, 70))
        u = draw(st.lists(st.integers(0, V - 1), min_size=S, max_size=S))
        v = draw(st.lists(st.integers(0, V - 1), min_size=S, max_size=S))
        kind = draw(st.sampled_from(["int", "zero", "float"]))
        if kind == "int":
            w = [float(x) for x in draw(st.lists(st.integers(1, 3), min_size=S, max_size=S))]
        elif kind == "zero":
            w = [float(x) for x in draw(st.lists(st.integers(0, 2), min_size=S, max_size=S))]
        else:
            w = draw(st.lists(st.floats(0.01, 50.0, allow_nan=False), min_size=S, max_size=S))
        perm = draw(st.permutations(list(range(V))))
        ranked = draw(st.booleans())
        src = draw(st.integers(0, V - 1))
        trips = draw(st.lists(st.tuples(st.integers(0, V - 1), st.integers(0, V - 1)), min_size=1, max_size=30))
        return V, u, v, w, (perm if ranked else None), src, trips

    @hsettings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
    @given(cases())
    def run(case):
        V, u, v, w, perm, src, trips = case
        u, v, w = np.array(u, np.int32), np.array(v, np.int32), np.array(w, np.float64)
        S = len(u)
        rank = None if perm is None else np.array(perm, np.int32)
        g = engine.DeviceGraph(dev, V, u, v, np.full(S, 10.0), np.full(S, 50, np.int32), node_rank=rank)
        csr = ctwin.Csr(V, u, v)
        dist, pred = g.sssp(w, src)
        d0, p0, _ = ctwin.sssp(csr, w, src)
        assert np.array_equal(dist.view(np.uint64), d0.view(np.uint64))
        assert np.array_equal(pred, p0)
        g.close()
        # a day on K1 weights (lengths double as the varying quantity so that K1 produces the ties / zeros)
        length = np.where(w == 0.0, 0.0, w * 40.0)
        g = engine.DeviceGraph(dev, V, u, v, length, np.full(S, 50, np.int32), node_rank=rank)
        o = np.array([t[0] for t in trips], np.int32)
        gl = np.array([t[1] for t in trips], np.int32)
        sim = engine.DeviceSim(g, o, gl, 0.25, phys)
        sim.step()
        ww = ctwin.weights(length, np.full(S, 50, np.int32), np.zeros(S, np.uint32), 0.25)
        want, cnt = ctwin.day(csr, ww, o, gl, nthreads=1)
        c = sim.counters()
        assert np.array_equal(sim.get_load(), want)
        assert c["trips_unreachable"] == cnt["unreachable"] and c["trips_stuck"] == cnt["stuck"] and c["hops"] == cnt["hops"]
        sim.close()
        g.close()

    old = dev.get_option("root_mode")
    dev.set_option("root_mode", 0)
    try:
        run()
    finally:
        dev.set_option("root_mode", old)
