"""Diagnostic: days with road adaptation on a road-like graph, per-day device time and SSSP counters."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from streets4mpi_b200 import engine, synth  # noqa: E402
from streets4mpi_b200.settings import settings  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=1_000_000)
ap.add_argument("--trips", type=int, default=1_000_000)
ap.add_argument("--days", type=int, default=9)
ap.add_argument("--every", type=int, default=2)
ap.add_argument("--options", default="")
args = ap.parse_args()
dev = engine.default_device()
for kv in filter(None, args.options.split(",")):
    k, v = kv.split("=")
    dev.set_option(k, float(v))
arr = synth.road_arrays(args.nodes, seed=1)
net = synth.network_from(arr)
flat = net.flat()
cats = synth.node_categories(arr["node_ids"], seed=1, goal_fraction=0.02)
o, g = synth.trip_arrays(args.trips, arr["node_ids"], cats["goals"], seed=3756917)
dev.set_option("root_mode", 2)
dev.set_option("group_log", 1)
graph = engine.DeviceGraph(dev, flat.V, flat.u, flat.v, flat.length, flat.max_speed, node_rank=flat.locality_rank())
sim = engine.DeviceSim(graph, flat.dense(o), flat.dense(g), 0.451, engine.make_physics(settings))
for day in range(args.days):
    if day > 0 and day % args.every == 0:
        sim.road_construction(flat.adaptable)
    sim.reset_counters()
    t0 = time.time()
    sim.step()
    dev.sync()
    wall = time.time() - t0
    c = sim.counters()
    sim.commit_total()
    w = sim.weights()
    vis, _ = sim.get_max_speed()
    gl = graph.group_log().astype(np.float64)
    out = dict(day=day, wall_ms=round(wall * 1e3, 1), ms_sssp=round(c["ms_sssp"], 1), ms_walk=round(c["ms_walk"], 1),
               rounds=c["relax_rounds"], adv=c["bucket_advances"], fb=c["fallback_sweeps"], pops=c["nodes_popped"],
               pushes=c["queue_pushes"], dups=c["stale_pops"], stuck=c["trips_stuck"], unreach=c["trips_unreachable"],
               w_mean=float(w.mean()), w_median=float(np.median(w)), w_max=float(w.max()),
               speed_hist=np.bincount(np.minimum(vis, 140) // 20, minlength=8).tolist())
    if len(gl):
        cyc = gl[:, 0] * 1024 / 1.965e6
        out["group_ms_pct"] = [round(float(x), 1) for x in np.percentile(cyc, [0, 50, 90, 99, 100])]
        out["group_rounds_pct"] = [round(float(x), 1) for x in np.percentile(gl[:, 1], [0, 50, 90, 99, 100])]
        out["group_exp_per_node_pct"] = [round(float(x) / flat.V, 2) for x in np.percentile(gl[:, 2], [0, 50, 90, 99, 100])]
    print(json.dumps(out), flush=True)
