"""Multi-GPU path: one process per GPU, packets sharded with the reference's load/rest rule,
integer tallies summed by an NCCL all-reduce (replacing MPI_ALLREDUCE,
iteration_mod.f90:627-659), then folded.  Result must equal the single-GPU run bit for bit
on every rank.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, n, q, overlap=False, sed_local=False, sparse=True, pipelined=False):
    try:
        _worker_body(rank, world, port, name, n, q, overlap, sed_local, sparse, pipelined)
    except BaseException as ex:              # a dead worker must not leave the parent waiting on the queue
        import traceback

        q.put((rank, "error: " + "".join(traceback.format_exception(type(ex), ex, ex.__traceback__))[-3000:]))
        raise


def _worker_body(rank, world, port, name, n, q, overlap=False, sed_local=False, sparse=True, pipelined=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from cases import make
    from mocassin_b200.api import PacketEngine

    m, _ = make(name)
    e = PacketEngine(m, device=rank, rank=rank, nranks=world, seed=12345)
    e.upload_iteration_inputs()
    if sed_local:
        e.set_sed_local(True)
    e.sparse_escaped = sparse
    e.pipelined_fold = pipelined
    if overlap:
        e.zero_estimators()
        e.energyPacketDriverOverlapped(1, n)
    else:
        e.lucy_transport([n])
    want = ["Jste", "escapedPackets"] + (["Jdif", "linePackets"] if m.lgDebug else [])
    out = [e.fetch(iG, want=want) for iG in range(1, m.nGrids + 1)]
    if sparse and not overlap and not sed_local:
        assert e.last_escaped_exchange is not None
    if sed_local:
        out.append(e.fetch_sed())
    q.put((rank, out))
    dist.barrier()
    e.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,overlap,sparse", [("multigrid_sym", False, True), ("multigrid_sym", True, True),
                                                 ("cube_clumpy_gasdust", False, True), ("cube_clumpy_gasdust", True, True),
                                                 ("viewing_angles", False, True), ("multigrid_sym", False, False),
                                                 ("multigrid_sym", "pipelined", True), ("hii_sym_gas_debug", False, True)])
def test_nccl_allreduce_matches_single_gpu(name, overlap, sparse):
    pipelined = overlap == "pipelined"
    overlap = overlap is True
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import make
    from mocassin_b200.api import PacketEngine

    n = 20001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, n, q, overlap, False, sparse, pipelined)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = dict(q.get(timeout=240) for _ in range(2))
        for v in got.values():
            assert not isinstance(v, str), v
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        for p in procs:                      # a rank stuck in a collective behind a dead peer
            if p.is_alive():
                p.kill()
    m, _ = make(name)
    e = PacketEngine(m, seed=12345)
    e.upload_iteration_inputs()
    e.lucy_transport([n])
    want = ["Jste", "escapedPackets"] + (["Jdif", "linePackets"] if m.lgDebug else [])
    for iG in range(1, m.nGrids + 1):
        ref = e.fetch(iG, want=want)
        for r in (0, 1):
            for k in want:
                assert np.array_equal(got[r][iG - 1][k], ref[k]), (iG, r, k)
    e.close()


@pytest.mark.parametrize("overlap", [False, True])
def test_nccl_sed_exchange_replaces_escaped_packets(overlap):
    """sed_local: the ranks all-reduce the (nu, angle) escape counts instead of the per-cell
    escapedPackets tallies.  Jste and the SED equal the single-GPU run bit for bit; the per-cell
    escapedPackets are rank-local and add up to the global array."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import make
    from mocassin_b200.api import PacketEngine

    name, n = "viewing_angles", 20001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, n, q, overlap, True)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = dict(q.get(timeout=240) for _ in range(2))
        for v in got.values():
            assert not isinstance(v, str), v
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        for p in procs:                      # a rank stuck in a collective behind a dead peer
            if p.is_alive():
                p.kill()
    m, _ = make(name)
    e = PacketEngine(m, seed=12345)
    e.upload_iteration_inputs()
    e.lucy_transport([n])
    ref = e.fetch(1)
    sed, cnt = e.fetch_sed()
    for r in (0, 1):
        assert np.array_equal(got[r][0]["Jste"], ref["Jste"])
        assert np.array_equal(got[r][-1][0].view(np.uint32), sed.view(np.uint32))
        assert np.array_equal(got[r][-1][1], cnt)
    both = got[0][0]["escapedPackets"].astype(np.float64) + got[1][0]["escapedPackets"].astype(np.float64)
    assert np.allclose(both, ref["escapedPackets"], rtol=1e-6)
    assert not np.array_equal(got[0][0]["escapedPackets"], ref["escapedPackets"])
    e.close()
