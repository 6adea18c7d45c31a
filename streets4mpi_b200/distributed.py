"""The rank-level exchange of the traffic load (reference streets4mpi.py:44-46, :97-100).

One process per GPU, ``torch.distributed`` (NCCL over NVLink / NVSwitch) in place
of ``mpi4py``: a single in-place sum all-reduce of ``uint32[S]`` per simulated
day on the device-resident counters, followed by the local
``cumulative += total``.  uint32 addition is exact and associative, so the
result does not depend on the number of GPUs or on how trips were partitioned.

torch has no unsigned 32-bit reductions; the buffer is reduced through an int32
view -- two's-complement addition produces the same bits.
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


def partition_by_root(origins, goals, rank, world_size, root_mode=2):
    """Tree-parallel decomposition of ONE virtual rank (one trip list, one jam tolerance) over several GPUs.

    The reference's only parallel axis is "ranks own trips" and every rank owns its jam tolerance, hence its own
    weights and trees (streets4mpi.py:69,78).  A single rank's day is however a sum over independent
    shortest-path trees, so its trees -- with the trips hanging off them -- can be dealt round-robin to the
    GPUs; the uint32 loads are summed by the same all-reduce and equal the single-process result bit for bit.
    Returns (origins, goals, forced_root_mode) of this GPU's share; root_mode as in include/s4mpi.h
    (2 = whichever endpoint set has fewer distinct nodes, decided on the whole list)."""
    origins, goals = np.asarray(origins), np.asarray(goals)
    if root_mode == 2:
        root_mode = 1 if len(np.unique(goals)) < len(np.unique(origins)) else 0
    roots = goals if root_mode == 1 else origins
    _, index = np.unique(roots, return_inverse=True)          # rank of every trip's root among the distinct roots
    mine = (index % world_size) == rank
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


def allreduce_array_I(arr, group=None):
    """Host-array form for drivers written against the reference API:
    ``total = allreduce_array_I(simulation.traffic_load)`` (streets4mpi.py:97-98)."""
    import torch
    from array import array
    a = np.frombuffer(arr, dtype=np.uint32).copy() if isinstance(arr, array) else np.array(arr, dtype=np.uint32)
    t = torch.from_numpy(a.view(np.int32))
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t = t.cpu()
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = array("I")
    out.frombytes(t.numpy().view(np.uint32).tobytes())
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
    and fold it into its cumulative load.  Nothing leaves the device."""
    sims = list(simulations)
    for s in sims:
        s._push_host_state()
    lead = sims[0]._sim
    for s in sims[1:]:
        lead.load_add(s._sim)
    _, size = world()
    if size > 1:
        import torch
        import torch.distributed as dist
        if dist.is_initialized():
            dev = sims[0]._device
            # NCCL orders itself after torch's current stream; if the library runs on that very stream the
            # whole day stays stream-ordered, otherwise fence the two streams through the host
            shared = dev.stream is not None and torch.cuda.current_stream(dev.index).cuda_stream == dev.stream
            if not shared:
                dev.sync()
            t = load_tensor(lead, dev.index)
            allreduce_u32_(t, group)
            if not shared:
                torch.cuda.current_stream(dev.index).synchronize()
    for s in sims[1:]:
        s._sim.load_copy(lead)
    for s in sims:
        s._sim.commit_total()
        s._host_load = None
        s._host_load_dirty = False
        s._cum_on_host = False
        s._cum_dirty = False
