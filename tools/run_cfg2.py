"""BASELINE.json configs[0] / configs[1] stand-in, end to end on one GPU.

The reference's osm/test.osm is a missing blob; synth.write_osm_xml generates a Stellingen-sized OSM XML file
(12,000 nodes).  The run follows streets4mpi.py:42-106 for P virtual ranks (default: the host's cores, i.e.
what `mpiexec -n <cores>` would start) hosted on one GPU: GraphBuilder -> per-rank trips (seed + 37 r) and jam
tolerance -> 10 days of Simulation.step() + exchange.  Every day's total load is compared with the oracle's C
twin routing the same trips on the host cores (bit-exact), and both are timed.

    python tools/run_cfg2.py [--residents 100000] [--ranks P] [--days 10]
"""
import argparse
import json
import os
import random
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ctwin  # noqa: E402  (checker / CPU baseline only)
from streets4mpi_b200 import Simulation, TripGenerator, TripTable, engine, synth  # noqa: E402
from streets4mpi_b200.osmdata import GraphBuilder  # noqa: E402
from streets4mpi_b200.settings import device_settings, settings  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--residents", type=int, default=100000)
ap.add_argument("--ranks", type=int, default=os.cpu_count() or 1)
ap.add_argument("--days", type=int, default=10)
ap.add_argument("--rows", type=int, default=100)
ap.add_argument("--cols", type=int, default=120)
args = ap.parse_args()

osm = os.path.join(tempfile.mkdtemp(), "test.osm")
synth.write_osm_xml(osm, rows=args.rows, cols=args.cols, seed=7)
t0 = time.perf_counter()
builder = GraphBuilder(osm)
net = builder.build_street_network()
builder.find_node_categories()
t_ingest = time.perf_counter() - t0
goals = builder.connected_commercial_nodes | builder.connected_industrial_nodes
flat = net.flat()
per_rank = args.residents // args.ranks                                  # streets4mpi.py:69
trips, jams = [], []
for r in range(args.ranks):
    random.seed(settings["random_seed"] + 37 * r)                         # :50-51
    trips.append(TripGenerator().generate_trips(per_rank, net.get_nodes(), goals))   # :75
    jams.append(random.random())                                          # :78
tables = [TripTable.from_dict(t) for t in trips]

dev = engine.default_device()
sims = [Simulation(net, tables[r], jams[r], lambda *a: None, device=dev) for r in range(args.ranks)]
dev.sync()
t0 = time.perf_counter()
gpu_totals = []
for day in range(args.days):
    for s in sims:
        s.step()
    sims[0].exchange(virtual_ranks=sims[1:])
    gpu_totals.append(np.frombuffer(sims[0].traffic_load, dtype=np.uint32).copy())
dev.sync()
t_gpu = time.perf_counter() - t0

csr = ctwin.Csr(flat.V, flat.u, flat.v)
dense = [(flat.dense(tb.origins), flat.dense(tb.goals)) for tb in tables]
prev = np.zeros(flat.S, dtype=np.uint32)
t0 = time.perf_counter()
ok = True
for day in range(args.days):
    total = np.zeros(flat.S, dtype=np.uint32)
    for r in range(args.ranks):
        w = ctwin.weights(flat.length, flat.max_speed, prev, jams[r])
        load, cnt = ctwin.day(csr, w, dense[r][0], dense[r][1], nthreads=os.cpu_count() or 1)
        total += load
    ok = ok and bool(np.array_equal(total, gpu_totals[day]))
    prev = total
t_cpu = time.perf_counter() - t0
trips_total = per_rank * args.ranks * args.days
print(json.dumps({
    "workload": "cfg2 stand-in: generated OSM XML %dx%d (V=%d, S=%d), %d residents over %d days, %d virtual ranks on 1 GPU"
                % (args.rows, args.cols, flat.V, flat.S, args.residents, args.days, args.ranks),
    "root_mode": {0: "origin", 1: "goal", 2: "auto"}[device_settings["root_mode"]],
    "loads_bit_exact_every_day": ok, "ingest_seconds": t_ingest,
    "gpu_seconds": t_gpu, "gpu_trips_per_sec": trips_total / t_gpu,
    "cpu_ctwin_seconds": t_cpu, "cpu_ctwin_trips_per_sec": trips_total / t_cpu, "cpu_threads": os.cpu_count(),
}))
