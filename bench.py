#!/usr/bin/env python3
"""Benchmark of the per-day traffic-assignment hot path (BASELINE.json metric: routed trips/s
and GTEPS; sim-day time against the CPU reference path).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU algorithm (oracle port)

A step is one simulated day: K1 weights -> K2 one shortest-path tree per root -> K3 walk of every
trip -> exchange (NCCL all-reduce of the uint32 load + cumulative merge).  Workload (config.workload):
BASELINE.json configs[2], the synthetic 1024x1024 grid (V=1,048,576 S=2,095,104 A=4,190,208), 1M
trips/day, goals = seeded 2 % node sample.  Default --scaling strong: ONE virtual rank (one trip list, one jam
tolerance: the reference run without mpiexec) whose distinct roots, sorted along the Hilbert curve, are cut into one
contiguous run per GPU (with the trips hanging off them); the daily NCCL all-reduce inside the library sums the loads
(bit-identical to the 1-GPU result: every line carries `load_hash`).
--scaling weak: one virtual rank with its own trips and jam tolerance per GPU (streets4mpi.py:50-51,69,78);
--scaling ranks: the 1M trips divided over per-GPU virtual ranks, every rank needing its own trees.

One JSON line on rank 0.  `value` = trips routed per second by the whole job with everything resident
in HBM (CUDA events, max over ranks); `e2e` = the same through the reference-facing Python API with
host arrays (array('I') traffic_load assigned / read every day, host<->device copies inside the timed
region); `roofline` = the SSSP kernel against the measured HBM copy bandwidth; `cpu_baseline` = the
oracle's C twin on this box's host cores on a bounded sample of the same day.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

RANDOM_SEED = 3756917          # settings.py:29
TRIPS_OVERRIDE = 0
WORKLOADS = {
    # name: (rows, cols, trips per day in total); rows == 0: random planar road graph with `cols` nodes
    "cfg3_grid1024": (1024, 1024, 1_000_000),
    "grid256": (256, 256, 100_000),           # quick functional run
    "cfg4_planar5m": (0, 5_000_000, 10_000_000),   # BASELINE.json configs[3] (adaptation every 5 days)
    "planar200k": (0, 200_000, 200_000),
    "planar1m": (0, 1_000_000, 1_000_000),          # tortuous stress case (no through roads)
    "cfg4_road5m": (-1, 5_000_000, 10_000_000),     # BASELINE.json configs[3] with a speed hierarchy (synth.road_arrays)
    "road1m": (-1, 1_000_000, 1_000_000),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# workload (identical for both arms)
# ----------------------------------------------------------------------------------------------
def make_workload(name, world, scaling="strong"):
    """scaling == "weak": every GPU (= one virtual rank with its own trip stream and jam tolerance) routes the
    workload's full day, the job routes world x that.  "strong": ONE virtual rank (one trip list, one jam
    tolerance -- the reference started without mpiexec) whose roots are cut into one contiguous run per GPU
    (distributed.partition_by_root); the all-reduce sums the loads.  "ranks": the day's trips divided over
    per-GPU virtual ranks like streets4mpi.py:69 divides number_of_residents (every rank needs its own trees)."""
    from streets4mpi_b200 import synth
    rows, cols, trips_total = WORKLOADS[name]
    if TRIPS_OVERRIDE:
        trips_total = TRIPS_OVERRIDE
    if scaling == "weak":
        trips_total *= world
    arrays = graph_arrays(name)
    cats = synth.node_categories(arrays["node_ids"], seed=1, goal_fraction=0.02)
    per_rank = trips_total if scaling == "strong" else trips_total // world      # streets4mpi.py:69
    return arrays, cats, per_rank


def graph_arrays(name):
    from streets4mpi_b200 import synth
    rows, cols, _ = WORKLOADS[name]
    if rows > 0:
        return synth.grid_arrays(rows, cols, seed=1)
    return synth.planar_arrays(cols, seed=1) if rows == 0 else synth.road_arrays(cols, seed=1)


def rank_trips(arrays, cats, per_rank, rank):
    """Trip arrays and jam tolerance of one virtual rank (streets4mpi.py:50-51,75,78)."""
    from streets4mpi_b200 import synth
    seed = RANDOM_SEED + 37 * rank
    o, g = synth.trip_arrays(per_rank, arrays["node_ids"], cats["goals"], seed=seed)
    jam = float(np.random.default_rng(seed ^ 0x5DEECE66D).random())
    return o, g, jam


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]                      # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU arms (oracle: test infrastructure / reported baseline only)
# ----------------------------------------------------------------------------------------------
def cpu_sample(arrays, o, g, jam, seconds_target, threads):
    """Oracle C twin (oracle/c/s4_oracle.c) on a bounded sample of the day's origins, the reference's
    grouping (one tree per distinct origin, simulation.py:71).  Returns (trips/s, description)."""
    from oracle import ctwin
    ctwin.build()
    V = len(arrays["node_ids"])
    csr = ctwin.Csr(V, arrays["u"], arrays["v"])
    w = ctwin.weights(arrays["length"], arrays["max_speed"], np.zeros(len(arrays["u"]), dtype=np.uint32), jam)
    roots, off, tg = ctwin.group_trips(o.astype(np.int32), g.astype(np.int32))
    # calibrate on `threads` trees, then size the sample for ~seconds_target
    def run(n_roots):
        sel = np.arange(n_roots)
        oo = np.repeat(roots[sel], np.diff(off)[sel])
        gg = np.concatenate([tg[off[i]:off[i + 1]] for i in sel]) if n_roots else np.zeros(0, np.int32)
        t0 = time.perf_counter()
        _load, cnt = ctwin.day(csr, w, oo, gg, nthreads=threads)
        return time.perf_counter() - t0, len(oo), cnt
    t_cal, _, _ = run(min(len(roots), threads))
    n = int(max(threads, min(len(roots), threads * seconds_target / max(t_cal, 1e-3))))
    t, trips, cnt = run(n)
    desc = ("%d of %d distinct origins of the day (%d trips), binary-heap Dijkstra + canonical walk, %d threads, %.1f s"
            % (n, len(roots), trips, threads, t))
    return trips / t, desc, dict(trees=n, trips=trips, seconds=t, trees_per_s=n / t)


_REF_GRAPH = None      # (row_ptr, col, arc_w) lists, inherited by the forked workers


def _py_tree_worker(jobs):
    """Pure-Python heapq Dijkstra + walk for a few origins: the reference algorithm at the reference's
    language level (simulation.py:71-88 over python-graph's strict-'<' Dijkstra)."""
    import heapq
    row_ptr, col, arc_w = _REF_GRAPH
    hops = 0
    t0 = time.perf_counter()
    for origin, goals in jobs:
        dist = {origin: 0.0}
        prev = {origin: None}
        done = set()
        heap = [(0.0, origin)]
        while heap:
            du, u = heapq.heappop(heap)
            if u in done:
                continue
            done.add(u)
            for e in range(row_ptr[u], row_ptr[u + 1]):
                v = col[e]
                if v in done:
                    continue
                alt = du + arc_w[e]
                if v not in dist or alt < dist[v]:
                    dist[v] = alt
                    prev[v] = u
                    heapq.heappush(heap, (alt, v))
        for goal in goals:
            if goal in prev:
                cur = goal
                while cur != origin:
                    cur = prev[cur]
                    hops += 1
    return time.perf_counter() - t0, hops


_REF_PG = None         # python-graph stand-in of the same network (faithful variant), inherited by the forked workers


def _py_faithful_worker(jobs):
    """The reference's Dijkstra as python-graph ran it (recalled 1.8.x: sorted list + insort, dict-of-dict graph;
    oracle/pygraph_standin.shortest_path), then the walk of simulation.py:78-88."""
    from oracle import pygraph_standin as pg
    hops = 0
    t0 = time.perf_counter()
    for origin, goals in jobs:
        prev, _dist = pg.shortest_path(_REF_PG, origin)
        for goal in goals:
            if goal in prev:
                cur = goal
                while cur != origin:
                    cur = prev[cur]
                    hops += 1
    return time.perf_counter() - t0, hops


def reference_arm(args):
    """--impl reference: the reference's CPU path.  The reference itself (Python 2 + python-graph +
    mpi4py + imposm) cannot run in this image (DESIGN.md 'Oracle'), so this times the oracle port: a
    pure-Python Dijkstra per distinct origin, one process per host core standing in for
    `mpiexec -n <cores>`, on a bounded sample of the same day."""
    import multiprocessing as mp
    from oracle import ctwin
    cores = os.cpu_count() or 1
    arrays, cats, per_rank = make_workload(args.workload, 1)
    o, g, jam = rank_trips(arrays, cats, per_rank, 0)
    V = len(arrays["node_ids"])
    csr = ctwin.Csr(V, arrays["u"], arrays["v"])
    w = ctwin.weights(arrays["length"], arrays["max_speed"], np.zeros(len(arrays["u"]), dtype=np.uint32), jam)
    roots, off, tg = ctwin.group_trips(o.astype(np.int32), g.astype(np.int32))
    global _REF_GRAPH
    _REF_GRAPH = (csr.row_ptr.tolist(), csr.col.tolist(), w[csr.arc_street].tolist())
    per_step = args.ref_trees_per_core * cores
    steps = args.warmup + args.steps
    times, trips_done = [], []
    with mp.get_context("fork").Pool(cores) as pool:
        for s in range(steps):
            sel = [(s * per_step + i) % len(roots) for i in range(per_step)]
            jobs = [(int(roots[i]), tg[off[i]:off[i + 1]].tolist()) for i in sel]
            chunks = [jobs[c::cores] for c in range(cores)]
            t0 = time.perf_counter()
            pool.map(_py_tree_worker, chunks)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
                trips_done.append(int(sum(off[i + 1] - off[i] for i in sel)))
    # the labelled *faithful* variant: python-graph's own data structures and queue discipline
    faithful = None
    if args.faithful_trees_per_core > 0:
        from oracle import pygraph_standin as pg
        global _REF_PG
        t0 = time.perf_counter()
        gobj = pg.graph()
        for n_ in range(V):
            gobj.add_node(n_)
        w_street = w.tolist()
        for s_, (a_, b_) in enumerate(zip(arrays["u"].tolist(), arrays["v"].tolist())):
            if a_ != b_ and not gobj.has_edge((a_, b_)):
                gobj.add_edge((a_, b_), wt=w_street[s_])
        build_s = time.perf_counter() - t0
        _REF_PG = gobj
        workers = min(cores, 8)
        n_f = args.faithful_trees_per_core * workers
        sel = list(range(n_f))
        jobs = [(int(roots[i]), tg[off[i]:off[i + 1]].tolist()) for i in sel]
        with mp.get_context("fork").Pool(workers) as pool:
            t0 = time.perf_counter()
            pool.map(_py_faithful_worker, [jobs[c::workers] for c in range(workers)])
            dt = time.perf_counter() - t0
        f_trips = int(sum(off[i + 1] - off[i] for i in sel))
        faithful = {"value": f_trips / dt, "unit": "trips/s", "cores": workers, "kind": "port",
                    "trees_per_sec": n_f / dt, "graph_build_seconds": build_s,
                    "sample": "%d distinct origins (%d per process, %d processes): sorted-list insort Dijkstra on a dict-of-dict "
                              "graph (oracle/pygraph_standin.shortest_path, the recalled python-graph 1.8.x algorithm), %.1f s"
                              % (n_f, args.faithful_trees_per_core, workers, dt)}
        _REF_PG = None
        del gobj
    total_t = sum(times)
    value = sum(trips_done) / total_t
    sample = ("%d distinct origins per step (%d per core) of %d; full day extrapolates linearly in the number of "
              "distinct origins" % (per_step, args.ref_trees_per_core, len(roots)))
    line = {
        "impl": "reference", "metric": "routed_trips_per_sec", "value": value, "unit": "trips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong" if args.scaling == "ranks" else args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + workload_text(args.workload), "root_mode": "origin (reference grouping)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "trips/s", "cores": cores, "kind": "port", "sample": sample,
                         "variant": "optimistic: heapq Dijkstra on flat lists (same trees as the faithful variant)"},
        "faithful_variant": faithful,
        "e2e": {"value": value, "unit": "trips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_text(name):
    rows, cols, trips = WORKLOADS[name]
    trips = TRIPS_OVERRIDE or trips
    if rows < 0:
        return ("synthetic road-like graph with a speed hierarchy (thinned jittered lattice, arterials every 16 and motorways "
                "every 128 junctions, V=%d, mean degree 2.6), %d trips/day, goals = 2%% node sample, BASELINE.json "
                "configs[3] shape" % (cols, trips))
    if not rows:
        return ("synthetic random planar road graph (thinned Delaunay, V=%d, S=%d, A=%d), %d trips/day, goals = 2%% node "
                "sample, BASELINE.json configs[3] shape" % (cols, int(1.2 * cols), int(2.4 * cols), trips))
    return ("synthetic %dx%d grid street network (V=%d, S=%d, A=%d), %d trips/day, goals = 2%% node sample, "
            "BASELINE.json configs[2]" % (rows, cols, rows * cols, 2 * rows * cols - rows - cols,
                                          2 * (2 * rows * cols - rows - cols), trips))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created; stdout must carry
        # exactly one JSON line, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize(local)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from streets4mpi_b200 import Simulation, TripTable, distributed, engine, synth
    from streets4mpi_b200.settings import device_settings, settings

    stream = torch.cuda.Stream(device=local)
    dev = engine.Device(local, stream=stream.cuda_stream)
    engine.set_default_device(dev)
    device_settings["root_mode"] = {"origin": 0, "goal": 1, "auto": 2}[args.root_mode]
    for kv in filter(None, args.options.split(",")):
        k, v = kv.split("=")
        device_settings[k] = float(v)
        dev.set_option(k, float(v))
    settings["logging"] = None

    if world > 1:
        distributed.init(dev)          # the library's own NCCL communicator (s4_comm_init): the daily exchange runs there
    arrays, cats, per_rank = make_workload(args.workload, world, args.scaling)
    net = synth.network_from(arrays)
    curve = net.flat().locality_rank()

    def strong_share(o_all, g_all, mode_in):
        # one virtual rank for the whole job: its trees, sorted along the Hilbert curve, are cut into one contiguous
        # run per GPU (neighbouring roots stay together: the SSSP kernel runs groups of neighbouring trees)
        flat0 = net.flat()
        order = None if curve is None else np.asarray(curve)
        od, gd = flat0.dense(o_all), flat0.dense(g_all)
        od, gd, forced_mode = distributed.partition_by_root(od, gd, rank, world, mode_in, order=order)
        return flat0.node_ids[od], flat0.node_ids[gd], forced_mode

    if args.scaling == "strong":
        o, g, jam = rank_trips(arrays, cats, per_rank, 0)
        o, g, forced = strong_share(o, g, device_settings["root_mode"])
        device_settings["root_mode"] = forced
        job_trips = per_rank
        per_rank = len(o)
    else:
        o, g, jam = rank_trips(arrays, cats, per_rank, rank)
        job_trips = per_rank * world
    t0 = time.perf_counter()
    sim = Simulation(net, TripTable(o, g), jam, lambda *a: None, device=dev)
    dev.sync()
    setup_s = time.perf_counter() - t0
    S = sim._sim.S
    mode = {0: "origin", 1: "goal", 2: "auto"}[sim._sim.root_mode]
    log("[rank %d] setup %.1f s, %d trips, %d trees/day (%s-rooted), jam %.3f" % (rank, setup_s, per_rank, sim._sim.n_roots, mode, jam))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(local)

    day_no = [0]

    def day_resident():
        with torch.cuda.stream(stream):
            if args.construction_every and day_no[0] > 0 and day_no[0] % args.construction_every == 0:
                sim.road_construction()                     # streets4mpi.py:86-88
            sim.step()
            sim.exchange()
            day_no[0] += 1

    # ---- resident-in-HBM timing: W warm-up days, then exactly K days between two events
    for _ in range(args.warmup):
        day_resident()
    barrier()
    sim._sim.reset_counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    day_marks = []
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            day_resident()
            if args.per_day:                                # one more event per day: where the days of a long run differ
                day_marks.append(torch.cuda.Event(enable_timing=True))
                day_marks[-1].record(stream)
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    per_day_ms = [a.elapsed_time(b) for a, b in zip([e0] + day_marks[:-1], day_marks)]
    clocks = sampler.stop() if rank == 0 else None
    c = sim.counters()
    adapt_events = sum(1 for d in range(args.warmup, args.warmup + args.steps)
                       if args.construction_every and d > 0 and d % args.construction_every == 0)
    # the state after warm-up + K days: uint32 sums are exact, so in strong mode these do not depend on N
    load_hash, cum_hash = sim._sim.load_hash()
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda:%d" % local)
    agg = torch.tensor([c["sssp_instances"], c["arcs_algorithmic"], c["nodes_algorithmic"], c["trips_routed"],
                        c["kernel_launches"], c["hops"], c["arcs_relaxed"]], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_max = float(t_ms.item())
    trees, arcs_alg, nodes_alg, trips, launches, hops, arcs_rel = [float(x) for x in agg.tolist()]

    # ---- end to end through the reference-facing API: host arrays every day (streets4mpi.py:93-100)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    from streets4mpi_b200 import merge_arrays
    sim.traffic_load = sim.traffic_load                     # the driver holds host copies from "yesterday":
    sim.cumulative_traffic_load = sim.cumulative_traffic_load   # every timed day starts with their upload
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        with torch.cuda.stream(stream):
            sim.step()                                      # uploads what the driver assigned yesterday (H2D 8*S)
            local_load = sim.traffic_load                   # D2H 4*S -> array('I')
            total = distributed.allreduce_array_I(local_load, device=dev) if world > 1 else local_load   # :97-98
            sim.traffic_load = total                        # :99
            sim.cumulative_traffic_load = merge_arrays((total, sim.cumulative_traffic_load))   # :100
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())

    # ---- the dominant kernel alone: one more day with strictly serial launches (lanes=1), so that the CUDA
    # events around every launch time that kernel and nothing else (in the timed region above two launch
    # lanes overlap, which is faster but makes per-launch times meaningless)
    barrier()
    lanes_default = dev.get_option("lanes")
    dev.set_option("lanes", 1)
    sim._sim.reset_counters()
    with torch.cuda.stream(stream):
        day_resident()
    barrier()
    cs = sim.counters()
    dev.set_option("lanes", lanes_default)
    # what every GPU did (rank 0 prints it): a day ends with an all-reduce, so the slowest GPU sets the pace
    slots_r, k_r = sim._sim.root_slots()
    mine = {"rank": rank, "trees": int(sim._sim.n_roots), "groups": int(len(slots_r) // max(1, k_r)), "trees_per_group_slots": int(k_r),
            "ms_timed_region_per_day": ms / args.steps, "ms_sssp_serial_day": cs["ms_sssp"], "ms_walk_serial_day": cs["ms_walk"],
            "sssp_launches_serial_day": int(cs["sssp_launches"]), "fallback_sweeps": int(c["fallback_sweeps"]),
            "trips_stuck": int(c["trips_stuck"])}
    rank_report = [mine]
    if world > 1:
        rank_report = [None] * world
        dist.all_gather_object(rank_report, mine)

    # ---- secondary measurements: every number the design notes quote gets a driver-run twin (a few seconds each)
    peak_gbs, _ = measured_peak()

    def measured_days(sim_x, days=2, warm=1):
        """{ms_per_day, trips/s ...} of `days` resident days (device events, max over ranks) + the SSSP kernel's
        roofline fraction from one more strictly serial day."""
        def one():
            with torch.cuda.stream(stream):
                sim_x.step()
                sim_x.exchange()
        for _ in range(warm):
            one()
        barrier()
        sim_x._sim.reset_counters()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(stream)
            for _ in range(days):
                one()
            b.record(stream)
        barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda:%d" % local)
        cx = sim_x.counters()
        tot = torch.tensor([cx["trips_routed"], cx["sssp_instances"], cx["hops"]], dtype=torch.float64, device="cuda:%d" % local)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        ms_day = float(t.item()) / days
        trips_d, trees_d, hops_d = [float(x) / days for x in tot.tolist()]
        dev.set_option("lanes", 1)
        sim_x._sim.reset_counters()
        one()
        barrier()
        cy = sim_x.counters()
        dev.set_option("lanes", lanes_default)
        algb = 20.0 * cy["arcs_algorithmic"] + 20.0 * cy["nodes_algorithmic"]
        frac = algb / (cy["ms_sssp"] * 1e-3) / 1e9 / peak_gbs if cy["ms_sssp"] > 0 else None
        return {"ms_per_day": ms_day, "trips_per_sec": trips_d / (ms_day * 1e-3), "trees_per_day": trees_d,
                "trees_per_sec": trees_d / (ms_day * 1e-3), "hops_per_trip": hops_d / max(1.0, trips_d),
                "sssp_roofline_frac_rank0": frac, "sssp_ms_serial_day_rank0": cy["ms_sssp"], "walk_ms_serial_day_rank0": cy["ms_walk"],
                "load_hash": "%016x" % sim_x._sim.load_hash()[0]}

    secondary = {}
    saved_mode = device_settings["root_mode"]
    if not args.no_secondary:
        if world == 1:
            # the reference's own grouping, measured: one tree per distinct origin (simulation.py:71)
            o2, g2, _ = rank_trips(arrays, cats, args.origin_rooted_trips, 0)
            device_settings["root_mode"] = 0
            sim_o = Simulation(net, TripTable(o2, g2), jam, lambda *a: None, device=dev)
            secondary["origin_rooted_measured"] = dict(measured_days(sim_o), trips_per_day=len(o2), root_mode="origin",
                                                       distinct_origins=int(sim_o._sim.n_roots))
            sim_o._sim.close()
        if world > 1:
            # the reference's mpiexec -n N semantics: every rank its own trips (1/N of the day) and jam tolerance
            o3, g3, jam3 = rank_trips(arrays, cats, job_trips // world if args.scaling == "strong" else per_rank, rank)
            device_settings["root_mode"] = {"origin": 0, "goal": 1, "auto": 2}[args.root_mode]
            sim_r = Simulation(net, TripTable(o3, g3), jam3, lambda *a: None, device=dev)
            secondary["ranks_mode"] = dict(measured_days(sim_r), trips_per_day_per_rank=len(o3),
                                           note="per-GPU virtual ranks with their own jam tolerance (streets4mpi.py:50-51,69,78): "
                                                "every rank needs its own trees")
            sim_r._sim.close()
        if world == 8 or args.north_star:
            # BASELINE.json north_star: a day with 10 M trips on the ~4 M-arc graph
            o4, g4, jam4 = rank_trips(arrays, cats, 10_000_000, 0)
            o4, g4, forced4 = strong_share(o4, g4, {"origin": 0, "goal": 1, "auto": 2}[args.root_mode])
            device_settings["root_mode"] = forced4
            sim_n = Simulation(net, TripTable(o4, g4), jam4, lambda *a: None, device=dev)
            secondary["north_star_10M"] = dict(measured_days(sim_n), trips_per_day=10_000_000, gpus=world)
            sim_n._sim.close()
        if world == 1:
            for wname in args.extra_graphs.split(","):
                if not wname:
                    continue
                arr_x = graph_arrays(wname)
                cats_x = synth.node_categories(arr_x["node_ids"], seed=1, goal_fraction=0.02)
                net_x = synth.network_from(arr_x)
                ox, gx, jx = rank_trips(arr_x, cats_x, WORKLOADS[wname][2], 0)
                device_settings["root_mode"] = {"origin": 0, "goal": 1, "auto": 2}[args.root_mode]
                sim_x = Simulation(net_x, TripTable(ox, gx), jx, lambda *a: None, device=dev)
                secondary[wname + "_day"] = dict(measured_days(sim_x), workload=workload_text(wname))
                sim_x._sim.close()
                net_x._device_graph = None
    device_settings["root_mode"] = saved_mode

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = 20.0 * cs["arcs_algorithmic"] + 20.0 * cs["nodes_algorithmic"]        # rank 0's launches that day
        n_launch = max(1, cs["sssp_launches"])
        ms_launch = cs["ms_sssp"] / n_launch
        achieved = alg_bytes / n_launch / (ms_launch * 1e-3) / 1e9 if ms_launch > 0 else 0.0
        serial_total = cs["ms_weights"] + cs["ms_sssp"] + cs["ms_walk"]
        traffic = None
        prof = os.path.join(ROOT, "profiles", "sssp_dram_bytes_per_tree.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof))["dram_bytes_per_tree"] * cs["sssp_instances"] / n_launch
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, desc, _ = cpu_sample(arrays, net.flat().dense(o), net.flat().dense(g), jam, args.cpu_seconds, threads)
            cpu = {"value": v, "unit": "trips/s", "cores": threads, "kind": "port", "sample": desc}
        total_trips = job_trips * args.steps
        line = {
            "metric": "routed_trips_per_sec", "value": total_trips / (ms_max * 1e-3), "unit": "trips/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.scaling == "ranks" else args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + workload_text(args.workload),
                       "trips_per_day": {"per_gpu_rank0": per_rank, "job": job_trips},
                       "decomposition": {"weak": "one virtual rank (own trips, own jam tolerance) per GPU",
                                         "strong": "one virtual rank; its trees, sorted along the Hilbert curve, cut into one contiguous run per GPU",
                                         "ranks": "the day's trips divided over per-GPU virtual ranks"}[args.scaling],
                       "root_mode": "%s (trees per day and GPU: %d)" % (mode, sim._sim.n_roots),
                       "virtual_ranks": 1 if args.scaling == "strong" else world, "construction_every": args.construction_every,
                       "l2": "inputs larger than L2 (per-tree distance arrays, 8 B x V x trees in flight >> 126 MB)",
                       "options": args.options or "defaults"},
            "load_hash": "%016x" % load_hash, "cumulative_hash": "%016x" % cum_hash,
            "hash_note": "FNV-1a of traffic_load / cumulative_traffic_load (uint32[S]) after warm-up + steps days; "
                         "identical for every N in strong mode",
            "gteps": arcs_alg / (ms_max * 1e-3) / 1e9,
            "sssp_trees_per_sec": trees / (ms_max * 1e-3),
            "relaxations_per_algorithmic_arc": arcs_rel / max(1.0, arcs_alg),
            # context: with the reference's own grouping (root_mode origin, one tree per distinct origin) the same
            # kernels route this many trips/s -- estimate from the measured tree rate, not a separate run
            "reference_grouping_estimate": {
                "distinct_origins_rank0": int(np.unique(o).size),
                "trips_per_sec": trees / (ms_max * 1e-3) * per_rank / max(1, int(np.unique(o).size)),
                "note": "root_mode=origin needs one tree per distinct origin instead of per distinct goal"},
            "secondary": secondary,
            "hops_per_trip": hops / max(1.0, trips),
            "kernel_ms_serial_day": {"weights": cs["ms_weights"], "sssp": cs["ms_sssp"], "walk": cs["ms_walk"],
                                     "sssp_share": cs["ms_sssp"] / serial_total if serial_total else None,
                                     "note": "one extra day with lanes=1 (separate kernels, no overlap), CUDA events per launch"},
            "roofline": {"bound": "hbm", "kernel": "sssp", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / n_launch, "ms_per_launch": ms_launch,
                         "launches": n_launch, "timing": "CUDA events around each launch, serial day (lanes=1)",
                         "frac_overlapped_region": (20.0 * c["arcs_algorithmic"] + 20.0 * c["nodes_algorithmic"])
                         / (ms * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu,
            "e2e": {"value": job_trips * e2e_steps / e2e_s, "unit": "trips/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": 8 * S, "d2h_bytes_per_step": 4 * S,
                    "api": "reference driver protocol (streets4mpi.py:93-100): Simulation.step(), array('I') "
                           "traffic_load read + assign, merge_arrays into cumulative_traffic_load; wall clock"},
            "gpu_launches": int(launches),
            "per_rank": rank_report,
            "per_day_ms_rank0": [round(x, 3) for x in per_day_ms] if args.per_day else None,
            "road_construction": {"events_in_timed_region": adapt_events,
                                  "ms_per_event_rank0": c["ms_adapt"] / adapt_events if adapt_events else None,
                                  "note": "K5: CUB radix sort of (cumulative load, street) + adapt_kernel, simulation.py:91-109"},
            "clocks": clocks,
            "setup_seconds": setup_s,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg3_grid1024")
    ap.add_argument("--root-mode", choices=["origin", "goal", "auto"], default="auto")
    ap.add_argument("--options", default="", help="library options, key=value,...")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-trees-per-core", type=int, default=1)
    ap.add_argument("--trips", type=int, default=0, help="override the workload's trips per day")
    ap.add_argument("--scaling", choices=["weak", "strong", "ranks"], default="strong",
                    help="weak: the workload's day per GPU (one virtual rank each); strong: one virtual rank, trees "
                         "dealt to the GPUs; ranks: the day's trips divided over per-GPU virtual ranks")
    ap.add_argument("--construction-every", type=int, default=0, help="road adaptation every N days (0 = never)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary measurements (origin-rooted day, "
                    "per-rank jam tolerances, north-star day, road-like graphs)")
    ap.add_argument("--origin-rooted-trips", type=int, default=30000)
    ap.add_argument("--north-star", action="store_true", help="measure the 10 M-trip day at any GPU count (default: at 8)")
    ap.add_argument("--extra-graphs", default="road1m,planar1m", help="workloads measured as secondary days at N=1")
    ap.add_argument("--faithful-trees-per-core", type=int, default=1)
    ap.add_argument("--per-day", action="store_true", help="record the device time of every timed day (rank 0)")
    args = ap.parse_args()
    global TRIPS_OVERRIDE
    TRIPS_OVERRIDE = args.trips
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
