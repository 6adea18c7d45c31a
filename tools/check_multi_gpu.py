"""Multi-GPU parity check (run under torchrun): every rank routes its own trips with its own jam tolerance,
the loads are summed by the NCCL all-reduce inside Simulation.exchange(), and rank 0 compares every day's
total with the oracle's C twin routing all ranks' trips on the host.

    torchrun --nproc-per-node N tools/check_multi_gpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ctwin  # noqa: E402  (checker only)
from streets4mpi_b200 import Simulation, TripTable, engine, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
arrays = synth.grid_arrays(200, 220, seed=5)
cats = synth.node_categories(arrays["node_ids"], seed=5, goal_fraction=0.02)
net = synth.network_from(arrays)
flat = net.flat()


def trips_of(r):
    o, g = synth.trip_arrays(20000, arrays["node_ids"], cats["goals"], seed=3756917 + 37 * r)
    return o, g, 0.1 + 0.8 * r / max(1, world - 1) if world > 1 else 0.5


o, g, jam = trips_of(rank)
dev = engine.Device(local)
engine.set_default_device(dev)
sim = Simulation(net, TripTable(o, g), jam, lambda *a: None, device=dev)
ok = True
csr = ctwin.Csr(flat.V, flat.u, flat.v) if rank == 0 else None
prev = np.zeros(flat.S, dtype=np.uint32)
cum = np.zeros(flat.S, dtype=np.uint64)
for day in range(4):
    if day == 2:
        sim.road_construction()
    sim.step()
    sim.exchange()
    total = np.frombuffer(sim.traffic_load, dtype=np.uint32).copy()
    if rank == 0:
        if day == 2:
            ms, _ = ctwin.road_construction(cum.astype(np.uint32), ms_now, flat.adaptable)
            ms_now = ms
            cum[:] = 0
        elif day == 0:
            ms_now = flat.max_speed.copy()
        want = np.zeros(flat.S, dtype=np.uint32)
        for r in range(world):
            oo, gg, jj = trips_of(r)
            w = ctwin.weights(flat.length, ms_now, prev, jj)
            load, _ = ctwin.day(csr, w, flat.dense(oo), flat.dense(gg), nthreads=os.cpu_count() or 1)
            want += load
        ok = ok and bool(np.array_equal(want, total))
        prev = want
        cum += want
        cumg = np.frombuffer(sim.cumulative_traffic_load, dtype=np.uint32)
        ok = ok and bool(np.array_equal(cumg, cum.astype(np.uint32)))
        print("day %d: total %d trips-hops, matches oracle: %s" % (day + 1, int(total.astype(np.int64).sum()), ok), flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda:%d" % local)
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
