"""The oracle (oracle/port.py restatement and its C twin) against the fixtures that
tests/golden/make_golden.py produced by executing the reference's own source."""
from array import array

import numpy as np
import pytest

from oracle import ctwin, port
from tests.helpers import case_arrays, fh, load_golden


def test_k1_known_answers_port_and_ctwin():
    g = load_golden("k1_weights.json")
    rows = g["rows"]
    assert len(rows) > 400
    length = np.array([fh(r["length"]) for r in rows])
    ms = np.array([r["max_speed"] for r in rows], dtype=np.int32)
    load = np.array([r["load"] for r in rows], dtype=np.uint32)
    for r in rows:
        L, jam = fh(r["length"]), fh(r["jam"])
        assert float(port.driving_speed(L, r["max_speed"], 0)).hex() == r["ideal"]
        assert float(port.driving_speed(L, r["max_speed"], r["load"])).hex() == r["actual"]
        assert port.driving_time(L, r["max_speed"], r["load"], jam).hex() == r["time"]
        assert float(ctwin.driving_speed(L, r["max_speed"], r["load"])).hex() == r["actual"]
        w = ctwin.weights([L], [r["max_speed"]], [r["load"]], jam)[0]
        assert float(w).hex() == r["time"]
    # SURVEY 8(c) table, spot values
    assert port.driving_time(100.0, 50, 10, 0.5).hex() == "0x1.3035923ac0021p+1"
    assert port.driving_time(1000.0, 120, 500, 0.75).hex() == "0x1.631ee12cb97b3p+3"
    assert port.driving_time(0.0, 50, 0, 0.5) == 0.0
    del length, ms, load


def test_merge_arrays_golden():
    g = load_golden("merge_arrays.json")
    a, b = array("I", g["a"]), array("I", g["b"])
    assert list(port.merge_arrays((a, b))) == g["a_b"]
    assert list(port.merge_arrays((a, None))) == g["a_none"]
    acc = np.array(g["a"], dtype=np.uint32)
    ctwin.lib().o_merge(len(acc), ctwin._p(acc, __import__("ctypes").c_uint32),
                        ctwin._p(np.array(g["b"], dtype=np.uint32), __import__("ctypes").c_uint32))
    assert acc.tolist() == g["a_b"]


def _port_network(case, sssp):
    net = port.Network(sssp=sssp)
    for n in case["nodes"]:
        net.add_node(n, 0.0, 0.0)
    for (u, v, L, ms) in case["streets"]:
        net.add_street((u, v), fh(L), ms)
    return net


@pytest.mark.parametrize("sssp", ["canonical", "pygraph", "pygraph_scan"])
def test_port_days_match_reference(sssp):
    for case in load_golden("days.json")["cases"]:
        if sssp == "pygraph_scan" and len(case["nodes"]) > 100:
            continue
        cfg = dict(port.DEFAULTS, trip_volume=case["trip_volume"])
        trips = []
        for t in case["trips"]:
            d = {}
            for o, g in t:
                d.setdefault(o, []).append(g)
            trips.append(d)
        jams = [fh(j) for j in case["jams"]]
        totals, net0 = port.run_days(lambda: _port_network(case, sssp), trips, jams,
                                     case["n_days"], case["construction_every"], cfg)
        for day, rec in enumerate(case["days"]):
            assert list(totals[day]) == rec["total"], (case["name"], day)


def test_ctwin_days_match_reference():
    """C twin: weights bit-equal, loads equal, adaptation equal, day by day."""
    for case in load_golden("days.json")["cases"]:
        a = case_arrays(case)
        csr = ctwin.Csr(a["V"], a["u"], a["v"])
        S = len(a["u"])
        nr = len(a["jams"])
        ms = a["max_speed"].copy()
        prev = np.zeros(S, dtype=np.uint32)
        cum = None
        for day, rec in enumerate(case["days"]):
            if day > 0 and day % case["construction_every"] == 0:
                ms, _ = ctwin.road_construction(cum, ms, a["adaptable"])
                assert ms.tolist() == rec["max_speed_after_construction"], (case["name"], day)
                cum = None
            total = np.zeros(S, dtype=np.uint32)
            for r in range(nr):
                w = ctwin.weights(a["length"], ms, prev, a["jams"][r])
                assert [float(x).hex() for x in w] == rec["weights"][r], (case["name"], day, r)
                o, g = a["trips"][r]
                load, cnt = ctwin.day(csr, w, o, g, trip_volume=case["trip_volume"], nthreads=2)
                assert load.tolist() == rec["local_loads"][r], (case["name"], day, r)
                assert cnt["tied_trips"] == 0 and cnt["stuck"] == 0
                total += load
            assert total.tolist() == rec["total"]
            prev = total
            cum = total.copy() if cum is None else cum + total
            assert cum.tolist() == rec["cumulative"]


def test_road_construction_golden():
    for c in load_golden("road_construction.json")["cases"]:
        S = c["S"]
        ms0 = np.array([s[2] for s in c["streets"]], dtype=np.int32)
        mask = np.array([1 if s[0] <= s[1] else 0 for s in c["streets"]], dtype=np.uint8)
        cum = np.array(c["cumulative"], dtype=np.uint32)
        seen, counts = ctwin.road_construction(cum, ms0, mask)
        assert seen.tolist() == c["max_speed_seen_by_weight_loop"], S
        hidden, _ = ctwin.road_construction(cum, ms0, None)
        assert hidden.tolist() == c["max_speed_insertion_orientation"], S
        dec, inc = port.construction_plan(S)
        assert counts == (len(dec), len(inc))
        assert c["cumulative_after"] is None


def test_construction_counts_survey_table():
    want = {10: (8, 0), 19: (16, 0), 20: (8, 8), 21: (9, 9), 100: (40, 40), 101: (41, 41), 1000: (400, 400)}
    for S, (d, i) in want.items():
        dec, inc = port.construction_plan(S)
        assert (len(dec), len(inc)) == (d, i)


def test_sssp_against_scipy():
    """Independent cross-check of the distances (SURVEY section 4 item v)."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import dijkstra
    rng = np.random.default_rng(5)
    rows, cols = 23, 31
    V = rows * cols
    idx = np.arange(V).reshape(rows, cols)
    u = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()]).astype(np.int32)
    v = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()]).astype(np.int32)
    w = rng.uniform(0.1, 13.0, size=len(u))
    csr = ctwin.Csr(V, u, v)
    m = sp.coo_matrix((np.concatenate([w, w]), (np.concatenate([u, v]), np.concatenate([v, u]))), shape=(V, V)).tocsr()
    for src in (0, 17, V - 1):
        dist, pred, tied = ctwin.sssp(csr, w, src)
        ref = dijkstra(m, directed=True, indices=src)
        np.testing.assert_allclose(dist, ref, rtol=1e-12)
        assert pred[src] == -1 and not tied.any()
        # the tree reproduces the distances exactly, left to right
        wmap = {(int(a), int(b)): float(x) for a, b, x in zip(u, v, w)}
        for t in (5, V // 2, V - 2):
            path = [t]
            while path[-1] != src:
                path.append(int(pred[path[-1]]))
            acc = 0.0
            for a, b in zip(path[::-1][:-1], path[::-1][1:]):
                acc = acc + wmap.get((a, b), wmap.get((b, a)))
            assert acc == dist[t]


def _zero_weight_chain():
    """R-C w=1, C-A w=0, A-B w=0 with id(B) < id(A) < id(C): a "smallest equal-cost neighbour" rule
    oscillates B -> A -> B; the plateau rule climbs B -> A -> C -> R."""
    B, A, Cn, R = 0, 1, 2, 3
    u = np.array([R, Cn, A], dtype=np.int32)
    v = np.array([Cn, A, B], dtype=np.int32)
    w = np.array([1.0, 0.0, 0.0])
    return (B, A, Cn, R), u, v, w


def test_plateau_rule_has_no_cycles_on_zero_weights():
    (B, A, Cn, R), u, v, w = _zero_weight_chain()
    csr = ctwin.Csr(4, u, v)
    dist, pred, _ = ctwin.sssp(csr, w, R)
    assert dist.tolist() == [1.0, 1.0, 1.0, 0.0]
    assert pred.tolist() == [A, Cn, R, -1]
    load, cnt = ctwin.day(csr, w, np.array([R], dtype=np.int32), np.array([B], dtype=np.int32))
    assert load.tolist() == [1, 1, 1] and cnt["stuck"] == 0 and cnt["hops"] == 3
    # the Python port applies the same rule
    import types
    nbrs = {R: [Cn], Cn: [R, A], A: [Cn, B], B: [A]}
    wmap = {(R, Cn): 1.0, (Cn, R): 1.0, (Cn, A): 0.0, (A, Cn): 0.0, (A, B): 0.0, (B, A): 0.0}
    g = types.SimpleNamespace(neighbors=lambda n: nbrs[n], edge_weight=lambda e: wmap[e])
    previous, d, _ = port.canonical_tree(g, R)
    assert previous == {R: None, Cn: R, A: Cn, B: A}


def test_plateau_rule_random_zero_weight_graphs():
    """Grids with 25 % zero-weight streets: every predecessor chain of the C twin reaches the root without
    repeating a node, reproduces the distance exactly, and the port's tree is the same tree."""
    import types
    rng = np.random.default_rng(12)
    rows, cols = 9, 11
    V = rows * cols
    idx = np.arange(V).reshape(rows, cols)
    u = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()]).astype(np.int32)
    v = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()]).astype(np.int32)
    for trial in range(4):
        w = np.where(rng.random(len(u)) < 0.25, 0.0, rng.integers(1, 4, len(u))).astype(np.float64)
        csr = ctwin.Csr(V, u, v)
        nbrs = {n: [] for n in range(V)}
        wmap = {}
        for a, b, x in zip(u.tolist(), v.tolist(), w.tolist()):
            nbrs[a].append(b); nbrs[b].append(a)
            wmap[(a, b)] = x; wmap[(b, a)] = x
        g = types.SimpleNamespace(neighbors=lambda n: nbrs[n], edge_weight=lambda e: wmap[e])
        for src in (0, V // 2 + trial):
            dist, pred, _ = ctwin.sssp(csr, w, src)
            previous, d, _ = port.canonical_tree(g, src)
            for n in range(V):
                assert (previous[n] if previous[n] is not None else -1) == pred[n]
                seen, here, acc = set(), n, []
                while here != src:
                    assert here not in seen
                    seen.add(here)
                    assert pred[here] >= 0          # plateaus stay below the bound at this density
                    acc.append(wmap[(int(pred[here]), here)])
                    here = int(pred[here])
                total = 0.0
                for x in reversed(acc):
                    total = total + x
                assert total == dist[n]
