"""Tuning probe: time a sample of trees on a BASELINE-shaped graph for several kernel options."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from streets4mpi_b200 import engine, synth  # noqa: E402
from streets4mpi_b200.settings import settings  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1024)
ap.add_argument("--cols", type=int, default=1024)
ap.add_argument("--planar", type=int, default=0)
ap.add_argument("--roots", type=int, default=1184)
ap.add_argument("--configs", type=str, default="")
ap.add_argument("--days", type=int, default=2)
ap.add_argument("--morton", type=int, default=0)
ap.add_argument("--rank", type=int, default=1)
ap.add_argument("--goals", type=int, default=0, help="1: the trees are rooted at the cfg3 goal sample (2 %% of the nodes), like a bench day")
ap.add_argument("--share", type=int, default=0, help="N: keep the first of N contiguous runs of the goals along the locality curve (one GPU's share of an N-GPU day)")
args = ap.parse_args()

dev = engine.default_device()
arr = synth.planar_arrays(args.planar, seed=1) if args.planar else synth.grid_arrays(args.rows, args.cols, seed=1)
V = len(arr["node_ids"])
if args.morton and not args.planar:
    # experiment: renumber the grid nodes along a Z-order curve (2x2 nodes per 32-byte sector of dist[])
    def spread(x):
        x = x.astype(np.uint64)
        x = (x | (x << 16)) & 0x0000FFFF0000FFFF
        x = (x | (x << 8)) & 0x00FF00FF00FF00FF
        x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0F
        x = (x | (x << 2)) & 0x3333333333333333
        x = (x | (x << 1)) & 0x5555555555555555
        return x
    ids = arr["node_ids"]
    r, c = ids // args.cols, ids % args.cols
    key = spread(c) | (spread(r) << 1)
    rank = np.empty(V, dtype=np.int64)
    rank[np.argsort(key, kind="stable")] = np.arange(V)
    arr["u"], arr["v"] = rank[arr["u"]], rank[arr["v"]]
cats = synth.node_categories(arr["node_ids"], seed=1)
rng = np.random.default_rng(0)
origins = rng.choice(arr["node_ids"], size=args.roots, replace=False)
goals = cats["goals"][rng.integers(0, len(cats["goals"]), size=args.roots)]
if args.goals:
    goals = cats["goals"] if args.roots <= 0 else cats["goals"][: args.roots]
    if args.share > 1:
        curve = synth.network_from(arr).flat().locality_rank()
        along = cats["goals"][np.argsort(curve[cats["goals"]], kind="stable")]
        n = len(along) // args.share
        goals = along[3 * n: 4 * n] if args.share >= 4 else along[:n]
    origins = rng.choice(arr["node_ids"], size=len(goals))
    dev.set_option("root_mode", 1)
phys = engine.make_physics(settings)
rank = synth.network_from(arr).flat().locality_rank() if args.rank else None
g = engine.DeviceGraph(dev, V, arr["u"], arr["v"], arr["length"], arr["max_speed"], node_rank=rank)
print("graph V=%d S=%d A=%d" % (V, g.S, g.A), flush=True)
configs = [dict()]
if args.configs:
    configs = [dict((k, float(v)) for k, v in (kv.split("=") for kv in c.split(","))) if c else dict()
               for c in args.configs.split(";")]
for cfg in configs:
    base = {k: dev.get_option(k) for k in cfg}
    for k, v in cfg.items():
        dev.set_option(k, v)
    sim = engine.DeviceSim(g, origins.astype(np.int32), goals.astype(np.int32), 0.4, phys)
    for d in range(args.days):
        sim.reset_counters()
        t0 = time.time()
        sim.step()
        dev.sync()
        wall = time.time() - t0
        c = sim.counters()
        sim.commit_total()
    alg_bytes = 20.0 * c["arcs_algorithmic"] + 20.0 * c["nodes_algorithmic"]
    out = dict(cfg=cfg, roots=sim.n_roots, ms_sssp=round(c["ms_sssp"], 2), ms_walk=round(c["ms_walk"], 2),
               ms_w=round(c["ms_weights"], 3), wall_ms=round(wall * 1e3, 1),
               us_per_tree=round(1e3 * c["ms_sssp"] / sim.n_roots, 1),
               gteps=round(c["arcs_algorithmic"] / c["ms_sssp"] / 1e6, 2),
               hbm_frac=round(alg_bytes / (c["ms_sssp"] * 1e-3) / 6540.5e9, 4),
               relax_x=round(c["arcs_relaxed"] / max(1, c["arcs_algorithmic"]), 2),
               pops_x=round(c["nodes_popped"] / max(1, c["nodes_algorithmic"]), 2),
               stale_x=round(c["stale_pops"] / max(1, c["nodes_algorithmic"]), 2),
               push_x=round(c["queue_pushes"] / max(1, c["nodes_algorithmic"]), 2),
               rounds=c["relax_rounds"] // sim.n_roots, adv=c["bucket_advances"] // sim.n_roots,
               fb=c["fallback_sweeps"], hops=c["hops"])
    if cfg.get("group_log"):
        gl = g.group_log().astype(np.float64)
        if len(gl):
            cyc = gl[:, 0] * 1024 / 1.965e6      # ms at 1965 MHz
            q = lambda a: [round(float(x), 1) for x in np.percentile(a, [0, 10, 50, 90, 99, 100])]
            out["group_ms_pct"] = q(cyc)
            out["group_rounds_pct"] = q(gl[:, 1])
            out["group_exp_per_node_pct"] = [round(x / V, 2) for x in np.percentile(gl[:, 2], [0, 10, 50, 90, 99, 100])]
            out["group_ms_sum_over_sms"] = round(float(cyc.sum()) / dev.info()["sm_count"], 1)
    print(json.dumps(out), flush=True)
    sim.close()
    for k, v in base.items():
        dev.set_option(k, v)
