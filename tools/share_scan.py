"""Diagnostic: every GPU's share of an N-GPU cfg3 day (bench.py's own partition), one after the other on one GPU."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from streets4mpi_b200 import distributed, engine, synth  # noqa: E402
from streets4mpi_b200.settings import settings  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = engine.default_device()
for kv in filter(None, (sys.argv[2] if len(sys.argv) > 2 else "").split(",")):
    k, v = kv.split("=")
    dev.set_option(k, float(v))
arrays, cats, per_rank = bench.make_workload("cfg3_grid1024", N, "strong")
net = synth.network_from(arrays)
flat = net.flat()
curve = flat.locality_rank()
o_all, g_all, jam = bench.rank_trips(arrays, cats, per_rank, 0)
phys = engine.make_physics(settings)
g = engine.DeviceGraph(dev, flat.V, flat.u, flat.v, flat.length, flat.max_speed, node_rank=curve)
for r in range(N):
    od, gd, mode = distributed.partition_by_root(flat.dense(o_all), flat.dense(g_all), r, N, 2, order=np.asarray(curve))
    dev.set_option("root_mode", mode)
    sim = engine.DeviceSim(g, od, gd, jam, phys)
    for day in range(3):
        sim.reset_counters()
        dev.sync()
        t0 = time.perf_counter()
        sim.step()
        dev.sync()
        wall = (time.perf_counter() - t0) * 1e3
        c = sim.counters()
        sim.commit_total()
    slots, K = sim.root_slots()
    extra = {}
    gl = g.group_log().astype(np.float64)
    if len(gl):
        msg = gl[:, 0] * 1024 / 1.965e6
        extra = {"group_ms_pct": [round(float(x), 1) for x in np.percentile(msg, [0, 50, 90, 99, 100])],
                 "group_ms_sum_over_sms": round(float(msg.sum()) / 148, 1), "first_groups_ms": [round(float(x), 1) for x in msg[:6]],
                 "last_groups_ms": [round(float(x), 1) for x in msg[-6:]]}
    print(json.dumps({"rank": r, **extra, "trips": len(od), "trees": sim.n_roots, "groups": len(slots) // K, "K": K, "wall_ms": round(wall, 1),
                      "ms_sssp": round(c["ms_sssp"], 1), "ms_walk": round(c["ms_walk"], 1), "launches": c["sssp_launches"],
                      "fb": c["fallback_sweeps"]}), flush=True)
    sim.close()
