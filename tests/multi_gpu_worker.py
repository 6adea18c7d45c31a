"""Worker of tests/test_gpu_multi.py (one process per GPU, launched with torch.distributed.run): the rank-level
exchange of streets4mpi.py:97-100 through the library's NCCL communicator, checked against the oracle.

mode "ranks":  the reference's own axis -- every rank routes its own trips with its own jam tolerance; every day's
               all-reduced total and the cumulative load must equal the oracle routing all ranks' trips.
mode "trees":  one virtual rank whose trees are dealt to the GPUs (distributed.partition_by_root): the totals must
               equal the single-rank oracle day, whatever the number of GPUs.
mode "driver": the zero-argument ``Streets4MPI()`` driver itself under the launcher (it must join the job
               and exchange every day): every rank ends with the same total, equal to the oracle routing every
               rank's trips with that rank's jam tolerance.
Rank 0 prints one JSON line with the verdict and the load hashes of every day.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ctwin  # noqa: E402  (checker only)
from streets4mpi_b200 import Simulation, TripTable, distributed, engine, synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "ranks"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = engine.Device(local)
engine.set_default_device(dev)
if mode == "driver":
    import torch.distributed as dist
    from streets4mpi_b200.settings import settings
    from streets4mpi_b200.streets4mpi import Streets4MPI
    settings.update({"osm_file": "synthetic:grid:40x50:3", "logging": None, "max_simulation_steps": 3,
                     "number_of_residents": 3001, "steps_between_street_construction": 10, "persist_traffic_load": False})
    app = Streets4MPI()                                     # joins the job itself (distributed.init inside)
    sim = app.simulation
    assert dev.comm_size() == (world, rank) and app.process_rank == rank
    flat = sim.street_network.flat()
    o = np.array([a for a, goals in sim.trips.items() for _ in goals], dtype=np.int64)
    g = np.array([b for _, goals in sim.trips.items() for b in goals], dtype=np.int64)
    assert len(o) == 3001 // world                          # streets4mpi.py:69
    mine = (flat.dense(o), flat.dense(g), float(sim.jam_tolerance), "%016x" % sim._sim.load_hash()[0])
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    ok = len(set(e[3] for e in everyone)) == 1              # the same total on every rank
    ok = ok and len(set(e[2] for e in everyone)) == world   # own jam tolerance per rank (:78)
    if rank == 0:
        csr = ctwin.Csr(flat.V, flat.u, flat.v)
        prev = np.zeros(flat.S, dtype=np.uint32)
        for day in range(3):
            want = np.zeros(flat.S, dtype=np.uint32)
            for (oo, gg, jj, _) in everyone:
                load, _ = ctwin.day(csr, ctwin.weights(flat.length, flat.max_speed, prev, jj), oo, gg, nthreads=os.cpu_count() or 1)
                want += load
            prev = want
        ok = ok and bool(np.array_equal(np.frombuffer(sim.traffic_load, dtype=np.uint32), prev))
        print(json.dumps({"mode": mode, "world": world, "ok": bool(ok), "load_hash_per_day": [everyone[0][3]]}), flush=True)
    box = [ok]
    dist.broadcast_object_list(box, src=0)
    dist.barrier()
    dev.comm_free()
    dist.destroy_process_group()
    sys.exit(0 if box[0] else 1)
distributed.init(dev)
assert dev.comm_size() == (world, rank)
arrays = synth.grid_arrays(200, 220, seed=5)
cats = synth.node_categories(arrays["node_ids"], seed=5, goal_fraction=0.02)
net = synth.network_from(arrays)
flat = net.flat()
n_virtual = world if mode == "ranks" else 1


def trips_of(r):
    o, g = synth.trip_arrays(20000, arrays["node_ids"], cats["goals"], seed=3756917 + 37 * r)
    return o, g, (0.1 + 0.8 * r / max(1, n_virtual - 1)) if n_virtual > 1 else 0.5


if mode == "ranks":
    o, g, jam = trips_of(rank)
else:
    o, g, jam = trips_of(0)
    order = flat.locality_rank()
    from streets4mpi_b200.settings import device_settings
    # order is indexed by dense node id; the grid's node ids are dense already
    o, g, forced = distributed.partition_by_root(o, g, rank, world, 2, order=order)
    device_settings["root_mode"] = forced
sim = Simulation(net, TripTable(o, g), jam, lambda *a: None, device=dev)
ok = True
csr = ctwin.Csr(flat.V, flat.u, flat.v) if rank == 0 else None
prev = np.zeros(flat.S, dtype=np.uint32)
cum = np.zeros(flat.S, dtype=np.uint64)
hashes = []
ms_now = flat.max_speed.copy()
for day in range(4):
    if day == 2:
        sim.road_construction()
    sim.step()
    sim.exchange()
    hashes.append("%016x" % sim._sim.load_hash()[0])
    total = np.frombuffer(sim.traffic_load, dtype=np.uint32).copy()
    # the host-array protocol of the reference driver gives the same total: all-reduce of the (already total) load
    # would multiply it by the world size, so run it on this rank's share of a fresh vector instead
    share = np.full(8, rank + 1, dtype=np.uint32)
    summed = np.frombuffer(distributed.allreduce_array_I(share, device=dev), dtype=np.uint32)
    ok = ok and bool((summed == world * (world + 1) // 2).all())
    if rank == 0:
        if day == 2:
            ms_now, _ = ctwin.road_construction(cum.astype(np.uint32), ms_now, flat.adaptable)
            cum[:] = 0
        want = np.zeros(flat.S, dtype=np.uint32)
        for r in range(n_virtual):
            oo, gg, jj = trips_of(r)
            w = ctwin.weights(flat.length, ms_now, prev, jj)
            load, _ = ctwin.day(csr, w, flat.dense(oo), flat.dense(gg), nthreads=os.cpu_count() or 1)
            want += load
        ok = ok and bool(np.array_equal(want, total))
        prev = want
        cum += want
        cumg = np.frombuffer(sim.cumulative_traffic_load, dtype=np.uint32)
        ok = ok and bool(np.array_equal(cumg, cum.astype(np.uint32)))
box = [ok]
if world > 1:
    import torch.distributed as dist  # noqa: E402
    dist.broadcast_object_list(box, src=0)
if rank == 0:
    print(json.dumps({"mode": mode, "world": world, "ok": bool(box[0]), "load_hash_per_day": hashes}), flush=True)
if world > 1:
    dist.barrier()
    dev.comm_free()
    dist.destroy_process_group()
sys.exit(0 if box[0] else 1)
