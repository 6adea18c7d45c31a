"""The rank-level exchange of the traffic load (reference streets4mpi.py:44-46, :97-100).

One process per GPU.  The collective itself lives in the library (include/s4mpi.h: ``s4_comm_init`` +
``s4_sim_exchange`` = ``ncclAllReduce(ncclUint32, ncclSum)`` in place on the device-resident counters,
stream-ordered behind the day's kernels, then ``cumulative += total``); this module only bootstraps it: the
launcher's environment (torchrun) names the ranks, ``torch.distributed`` ships NCCL's 128-byte id from rank 0 to
the others once, and from then on no torch call sits on the day path.  uint32 addition is exact and
associative, so the result does not depend on the number of GPUs or on how trips were partitioned.

Without a GPU (the world_size-2 ``gloo`` tests of the host logic) the same protocol runs on host arrays through
``torch.distributed`` -- int32 view of the uint32 buffer, two's-complement addition produces the same bits.
"""
from __future__ import annotations

import os

import numpy as np


def world():
    """(rank, world_size) -- torch.distributed if initialised, else the launcher's
    environment, else a single process (the reference's no-mpiexec case, README.md:73-80)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def rank_seed(random_seed, rank):
    """streets4mpi.py:50: every rank owns the stream seeded with random_seed + 37 * rank."""
    return random_seed + 37 * rank


def residents_of_rank(number_of_residents, world_size):
    """streets4mpi.py:69 (Python-2 integer division: remainder trips are dropped)."""
    return number_of_residents // world_size


def init(device=None, backend=None):
    """Join the job the launcher started (streets4mpi.py:44-46): initialise ``torch.distributed`` if WORLD_SIZE > 1
    and it is not yet, and -- when a GPU ``device`` (engine.Device) is given -- create the library's NCCL
    communicator on it.  Returns (rank, world_size).  Idempotent."""
    rank, size = world()
    if size <= 1:
        return rank, size
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if (device is not None and torch.cuda.is_available()) else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", device.index)
        dist.init_process_group(backend, rank=rank, world_size=size, **kw)
    if device is not None and device.comm_size()[0] != size:
        box = [device.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        device.comm_init(size, rank, box[0])
    return rank, size


def partition_by_root(origins, goals, rank, world_size, root_mode=2, order=None):
    """Tree-parallel decomposition of ONE virtual rank (one trip list, one jam tolerance) over several GPUs.

    The reference's only parallel axis is "ranks own trips" and every rank owns its jam tolerance, hence its own
    weights and trees (streets4mpi.py:69,78).  A single rank's day is however a sum over independent
    shortest-path trees, so its trees -- with the trips hanging off them -- can be dealt to the GPUs; the uint32
    loads are summed by the same all-reduce and equal the single-process result bit for bit.
    ``order`` (optional, one number per node id, e.g. ``FlatNetwork.locality_rank()`` indexed by dense id): the
    distinct roots are sorted by it and cut into ``world_size`` contiguous runs with (nearly) equally many trips,
    so that every GPU keeps spatially neighbouring roots (the source-batched SSSP kernel groups neighbouring
    roots); without it the roots are dealt round-robin.
    Returns (origins, goals, forced_root_mode) of this GPU's share; root_mode as in include/s4mpi.h
    (2 = whichever endpoint set has fewer distinct nodes, decided on the whole list)."""
    origins, goals = np.asarray(origins), np.asarray(goals)
    if root_mode == 2:
        root_mode = 1 if len(np.unique(goals)) < len(np.unique(origins)) else 0
    roots = goals if root_mode == 1 else origins
    distinct, index, counts = np.unique(roots, return_inverse=True, return_counts=True)
    if order is None or world_size == 1:
        mine = (index % world_size) == rank
    else:
        by_order = np.argsort(np.asarray(order)[distinct], kind="stable")     # distinct roots along the curve
        cum = np.cumsum(counts[by_order])
        owner_sorted = np.minimum(world_size - 1, (cum - 1) * world_size // max(1, int(cum[-1])))
        owner = np.empty(len(distinct), dtype=np.int64)
        owner[by_order] = owner_sorted
        mine = owner[index] == rank
    return origins[mine], goals[mine], root_mode


def allreduce_u32_(tensor, group=None):
    """In-place SUM all-reduce of a uint32 buffer held as an int32 torch tensor
    (CPU tensor + gloo in the tests, CUDA tensor + NCCL on the GPUs)."""
    import torch
    import torch.distributed as dist
    if tensor.dtype != torch.int32:
        raise TypeError("pass the uint32 buffer as its int32 view")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def allreduce_array_I(arr, group=None, device=None):
    """Host-array form for drivers written against the reference API:
    ``total = allreduce_array_I(simulation.traffic_load)`` (streets4mpi.py:97-98).  With a GPU ``device`` that
    holds the library communicator the sum runs through its pinned staging buffer and NCCL; otherwise through
    torch.distributed on the host (gloo)."""
    from array import array
    a = np.frombuffer(arr, dtype=np.uint32).copy() if isinstance(arr, array) else np.array(arr, dtype=np.uint32)
    _, size = world()
    if device is not None and device.comm_size()[0] > 1:
        a = device.allreduce_host(a)
    elif size > 1:
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("WORLD_SIZE=%d but no communicator: call streets4mpi_b200.distributed.init() first" % size)
        t = torch.from_numpy(a.view(np.int32))
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t = t.cpu()
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        a = t.numpy().view(np.uint32)
    out = array("I")
    out.frombytes(a.tobytes())
    return out


class _DeviceSpan(object):
    """Expose a raw device pointer through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def load_tensor(device_sim, device_index):
    """int32 CUDA tensor aliasing the simulation's device-resident traffic_load."""
    import torch
    if device_sim.S == 0:
        return torch.empty(0, dtype=torch.int32, device="cuda:%d" % device_index)
    return torch.as_tensor(_DeviceSpan(device_sim.load_device_ptr(), device_sim.S), device="cuda:%d" % device_index)


def exchange(simulations, group=None):
    """streets4mpi.py:97-100 for the Simulations hosted on this GPU (virtual ranks):
    sum their local loads, all-reduce across GPUs, hand every one of them the total
    and fold it into its cumulative load.  Nothing leaves the device: the all-reduce is the library's
    (s4_sim_exchange).  Raises when the launcher announced several ranks and no communicator exists -- a rank
    that silently kept its partial load would route the next day on wrong weights."""
    sims = list(simulations)
    for s in sims:
        s._push_host_state()
    lead = sims[0]._sim
    for s in sims[1:]:
        lead.load_add(s._sim)
    _, size = world()
    dev = sims[0]._device
    if size > 1 and dev.comm_size()[0] != size:
        raise RuntimeError("WORLD_SIZE=%d but this device has no communicator of that size: call "
                           "streets4mpi_b200.distributed.init(device) before the first exchange" % size)
    lead.exchange()                                   # NCCL all-reduce (if several ranks) + cumulative merge
    for s in sims[1:]:
        s._sim.load_copy(lead)
        s._sim.commit_total()
    for s in sims:
        s._host_load = None
        s._host_load_dirty = False
        s._cum_on_host = False
        s._cum_dirty = False
