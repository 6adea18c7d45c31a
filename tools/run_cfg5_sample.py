"""BASELINE.json configs[4] shape on one GPU, bounded sample.

cfg5 = a country-scale road graph (~5e7 nodes) replicated per GPU, 1e8 trips/day with residential-only origins
(a seeded 1 % node sample, 5e5 distinct origins = trees per day), weak scaling.  A full day is 5e5 trees per GPU; this
tool builds the full-size graph (synth.ladder_arrays: street chains joined by sparse rungs, mean degree 2.4), keeps
the day's trips of a SAMPLE of the residential origins (default 256 trees, every one with its share of the day's
trips), runs two days with the congestion feedback between them and checks

  * one whole tree (distances bit-equal, predecessors equal) against the oracle's C twin,
  * conservation (sum of the loads == hops x trip_volume), no unreachable / stuck trips, no fallback sweeps,

and reports trees/s, trips/s and the extrapolated day.

    python tools/run_cfg5_sample.py [--rows 7071 --cols 7071] [--trees 256] [--trips-per-day 100000000]

--contiguous (default): the sampled origins are a contiguous run of the residential nodes along the locality curve, as
dense as in the full day (1 % of the nodes), so that the groups of neighbouring trees the SSSP kernel forms look like
the full day's; --scattered draws them over the whole graph (r1's sample).
Under torchrun (weak scaling, BASELINE cfg5): every rank takes its own run of origins and its own trips, the daily
exchange (NCCL all-reduce of the uint32[S] load through the library communicator) is timed, rank 0 prints all ranks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ctwin  # noqa: E402  (checker only)
from streets4mpi_b200 import engine, synth  # noqa: E402
from streets4mpi_b200.settings import settings  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=7071)
ap.add_argument("--cols", type=int, default=7071)
ap.add_argument("--trees", type=int, default=256)
ap.add_argument("--trips-per-day", type=int, default=100_000_000)
ap.add_argument("--days", type=int, default=2)
ap.add_argument("--no-oracle", action="store_true")
ap.add_argument("--scattered", action="store_true")
ap.add_argument("--options", default="", help="library options, key=value,...")
args = ap.parse_args()
rank_id, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))

t0 = time.perf_counter()
arr = synth.ladder_arrays(args.rows, args.cols, seed=1)
V, S = len(arr["node_ids"]), len(arr["u"])
net = synth.network_from(arr)
flat = net.flat()
rank = flat.locality_rank()
t_gen = time.perf_counter() - t0
cats = synth.node_categories(arr["node_ids"], seed=1, goal_fraction=0.02, residential_fraction=0.01)
residential, goals = cats["residential"], cats["goals"]
# the day's trips of the sampled origins: every residential origin gets trips_per_day / len(residential) on average
rng = np.random.default_rng(settings["random_seed"] + 37 * rank_id)
if args.scattered:
    origins = rng.choice(residential, size=min(args.trees, len(residential)), replace=False)
else:
    along = residential[np.argsort(rank[flat.dense(residential)], kind="stable")]      # residential nodes along the curve
    n_take = min(args.trees, len(along))
    start = (len(along) // (world + 1)) * (rank_id + 1) - n_take // 2
    origins = along[max(0, start): max(0, start) + n_take]
per_origin = rng.poisson(args.trips_per_day / len(residential), size=len(origins)).clip(min=1)
o = np.repeat(origins, per_origin).astype(np.int32)
g = goals[rng.integers(0, len(goals), size=len(o))].astype(np.int32)

dev = engine.Device(int(os.environ.get("LOCAL_RANK", "0")))
engine.set_default_device(dev)
dev.set_option("root_mode", 0)                          # residential origins: the reference's own grouping
for kv in filter(None, args.options.split(",")):
    k, v = kv.split("=")
    dev.set_option(k, float(v))
if world > 1:
    from streets4mpi_b200 import distributed
    distributed.init(dev)
t0 = time.perf_counter()
graph = engine.DeviceGraph(dev, V, flat.u, flat.v, flat.length, flat.max_speed, node_rank=rank)
phys = engine.make_physics(settings)
sim = engine.DeviceSim(graph, o, g, 0.4, phys)
dev.sync()
t_upload = time.perf_counter() - t0

out = {"workload": "cfg5 shape: ladder road graph %dx%d (V=%d, S=%d, A=%d), %d sampled residential origins of %d, %d trips "
                   "(%.0f per origin: %d trips/day over all origins)"
                   % (args.rows, args.cols, V, S, graph.A, len(origins), len(residential), len(o), len(o) / len(origins),
                      args.trips_per_day),
       "generate_seconds": round(t_gen, 1), "upload_seconds": round(t_upload, 1), "days": []}
for day in range(args.days):
    sim.reset_counters()
    t0 = time.perf_counter()
    sim.step()
    dev.sync()
    wall = time.perf_counter() - t0
    c = sim.counters()
    load = sim.get_load()
    ok = (c["trips_routed"] == len(o) and c["trips_unreachable"] == 0 and c["trips_stuck"] == 0 and
          c["fallback_sweeps"] == 0 and int(load.astype(np.int64).sum()) == c["hops"] * phys.trip_volume)
    alg_bytes = 20.0 * c["arcs_algorithmic"] + 20.0 * c["nodes_algorithmic"]
    out["days"].append({"day": day + 1, "wall_seconds": round(wall, 3), "trees": sim.n_roots,
                        "trees_per_sec": round(sim.n_roots / wall, 1), "trips_per_sec": round(len(o) / wall, 1),
                        "ms_sssp": round(c["ms_sssp"], 1), "ms_walk": round(c["ms_walk"], 1),
                        "gteps": round(c["arcs_algorithmic"] / wall / 1e9, 2),
                        "sssp_hbm_frac": round(alg_bytes / (c["ms_sssp"] * 1e-3) / 6540.5e9, 4),
                        "relaxations_per_arc": round(c["arcs_relaxed"] / max(1, c["arcs_algorithmic"]), 2),
                        "rounds_per_tree": c["relax_rounds"] // max(1, sim.n_roots),
                        "hops_per_trip": round(c["hops"] / len(o), 1), "properties_ok": bool(ok)})
    t0 = time.perf_counter()
    sim.exchange()                                      # NCCL all-reduce of uint32[S] (several ranks) + cumulative merge
    dev.sync()
    out["days"][-1]["exchange_seconds"] = round(time.perf_counter() - t0, 4)
full_day = len(residential) / out["days"][-1]["trees_per_sec"]
out["extrapolated_full_day_seconds_1gpu"] = round(full_day, 1)
if not args.no_oracle:
    w = sim.weights()
    src = int(flat.dense(origins[:1])[0])
    t0 = time.perf_counter()
    dist, pred = graph.sssp(w, src)
    t_gpu_tree = time.perf_counter() - t0
    csr = ctwin.Csr(V, flat.u, flat.v)
    t0 = time.perf_counter()
    d0, p0, tied = ctwin.sssp(csr, w, src)
    t_cpu_tree = time.perf_counter() - t0
    out["oracle_tree"] = {"dist_bit_equal": bool(np.array_equal(dist.view(np.uint64), d0.view(np.uint64))),
                          "pred_equal": bool(np.array_equal(pred, p0)), "cpu_seconds_one_tree": round(t_cpu_tree, 1),
                          "gpu_seconds_one_tree_with_copies": round(t_gpu_tree, 2)}
out["rank"], out["world"] = rank_id, world
out["options"] = args.options or "defaults"
if world > 1:
    import torch.distributed as dist
    box = [None] * world
    dist.all_gather_object(box, out)
    if rank_id == 0:
        print(json.dumps({"ranks": box}))
    dist.barrier()
    dev.comm_free()
    dist.destroy_process_group()
else:
    print(json.dumps(out))
