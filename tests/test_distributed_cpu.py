"""world_size-2 gloo test of the multi-rank host logic: the reference's load/rest packet
split (iteration_mod.f90:477-493) + integer-tally all-reduce gives exactly the
single-rank result.  Runs the oracle per rank (CPU); the GPU equivalent is
tests/test_gpu_parity.py::test_rank_partition_invariance_on_device."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cases import make
    from mocassin_b200.api import partition
    from oracle.oracle import Oracle

    m, _ = make("multigrid_sym")
    o = Oracle(m, fp32_tallies=False)
    first, cnt = partition(n, rank, world)
    c, _ = o.transport(1, first, cnt)
    sums = []
    for out in o.out:
        for k in ("JsteQ", "escapedQ"):
            t = torch.from_numpy(np.ascontiguousarray(out[k].reshape(-1, order="F")))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)      # replaces MPI_ALLREDUCE iteration_mod.f90:627,653
            sums.append(t.numpy().copy())
    seg = torch.tensor([c["nSegments"]], dtype=torch.int64)
    dist.all_reduce(seg)
    if rank == 0:
        q.put((sums, int(seg.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_rank():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import make
    from oracle.oracle import Oracle

    n = 3001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    sums, seg = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, _ = make("multigrid_sym")
    o = Oracle(m, fp32_tallies=False)
    c, _ = o.transport(1, 0, n)
    want = []
    for out in o.out:
        for k in ("JsteQ", "escapedQ"):
            want.append(out[k].reshape(-1, order="F"))
    assert seg == c["nSegments"]
    for a, b in zip(sums, want):
        assert np.array_equal(a, b)
