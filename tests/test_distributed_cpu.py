"""World-size-2 gloo runs (CPU) of the host-side exchange: the uint32 sum all-reduce through
the int32 view, the array('I') driver protocol and the per-rank partition of the residents."""
import os
import socket
import sys
from array import array

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from streets4mpi_b200 import distributed
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert distributed.world() == (rank, world)
        rng = np.random.default_rng(100 + rank)
        local = rng.integers(0, 2 ** 32, size=1000, dtype=np.uint64).astype(np.uint32)   # wraps on purpose
        t = torch.from_numpy(local.copy().view(np.int32))
        distributed.allreduce_u32_(t)
        total_a = distributed.allreduce_array_I(array("I", local.tolist()))
        out.put((rank, local, t.numpy().view(np.uint32).copy(), np.frombuffer(total_a, dtype=np.uint32).copy(),
                 distributed.residents_of_rank(1001, world), distributed.rank_seed(3756917, rank)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = (res[0][1].astype(np.uint64) + res[1][1].astype(np.uint64)).astype(np.uint32)     # mod 2**32
    for rank, _local, total_t, total_a, residents, seed in res:
        assert np.array_equal(total_t, want) and np.array_equal(total_a, want)
        assert residents == 500 and seed == 3756917 + 37 * rank


def test_single_process_is_identity():
    import torch
    from streets4mpi_b200 import distributed
    a = np.arange(10, dtype=np.uint32)
    t = torch.from_numpy(a.copy().view(np.int32))
    assert np.array_equal(distributed.allreduce_u32_(t).numpy().view(np.uint32), a)
    assert list(distributed.allreduce_array_I(array("I", a.tolist()))) == a.tolist()
    with pytest.raises(TypeError):
        distributed.allreduce_u32_(torch.zeros(3))


def test_partition_by_root_is_an_exact_cover():
    """Tree-parallel decomposition of one virtual rank: every trip lands on exactly one GPU, no tree is split,
    the tree counts are balanced, and the forced root mode is the globally smaller endpoint set."""
    from streets4mpi_b200 import distributed
    rng = np.random.default_rng(3)
    o = rng.integers(0, 5000, 20000)
    g = rng.choice(np.arange(0, 5000, 50), 20000)               # 100 distinct goals
    for world in (1, 2, 3, 8):
        parts = [distributed.partition_by_root(o, g, r, world, 2) for r in range(world)]
        assert all(p[2] == 1 for p in parts)                     # goal-rooted: 100 goals < ~4900 origins
        assert sum(len(p[0]) for p in parts) == len(o)
        roots = [set(np.unique(p[1]).tolist()) for p in parts]
        assert sum(len(r) for r in roots) == 100 and len(set().union(*roots)) == 100
        assert max(len(r) for r in roots) - min(len(r) for r in roots) <= 1
        merged = np.sort(np.concatenate([p[0] * 10000 + p[1] for p in parts]))
        assert np.array_equal(merged, np.sort(o * 10000 + g))
    assert distributed.partition_by_root(g, o, 0, 2, 2)[2] == 0   # fewer distinct origins -> origin-rooted
    assert distributed.partition_by_root(o, g, 0, 2, 0)[2] == 0   # explicit mode is respected
