"""Multi-GPU parity (-m gpu, needs >= 2 devices): the daily exchange of streets4mpi.py:97-100 through the library's
NCCL communicator (s4_comm_init / s4_sim_exchange), one process per GPU, against the oracle -- for the reference's
own decomposition (every rank its own trips and jam tolerance) and for the tree-parallel one, whose totals must
not depend on the number of GPUs."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(mode, world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), mode]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert p.returncode == 0 and lines, (p.returncode, p.stdout[-2000:], p.stderr[-4000:])
    return json.loads(lines[-1])


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_rank_parallel_exchange_matches_oracle():
    out = _run("ranks", 2)
    assert out["ok"] and out["world"] == 2


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_tree_parallel_totals_do_not_depend_on_gpu_count():
    worlds = [w for w in (1, 2, 4, 8) if w <= _gpus()]
    outs = [_run("trees", w) for w in worlds]
    assert all(o["ok"] for o in outs)
    assert all(o["load_hash_per_day"] == outs[0]["load_hash_per_day"] for o in outs), outs


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_streets4mpi_driver_under_the_launcher():
    """``Streets4MPI()`` started by torchrun joins the job, exchanges every day, and every rank
    ends with the oracle's total over all ranks' trips."""
    out = _run("driver", 2)
    assert out["ok"] and out["world"] == 2
