"""The library's own NCCL exchange (mcb200_comm_init / mcb200_exchange): what a Fortran/MPI
host calls in place of the MPI_ALLREDUCE block iteration_mod.f90:564,627,649,653,659.

One GPU: a one-rank communicator with option defer_fold walks the whole exchange path
(dlopen of NCCL, flag max-reduce, plane all-reduces, compact -> all-gather -> scatter of the
escape counts, dense variant, sed_local) and must leave every estimator bit-identical to the
run that folded immediately.  Two GPUs (skipped otherwise): two processes bootstrap the
communicator from the 128 id bytes alone (a file stands in for MPI_BCAST; no torch.distributed)
and must reproduce the single-GPU result bit for bit on both ranks."""
import os
import sys
import tempfile
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _want(m):
    return ["Jste", "escapedPackets"] + (["Jdif", "linePackets"] if m.lgDebug else [])


def _run(name, n, native, dense=False, sed_local=False):
    from cases import make
    from mocassin_b200.api import PacketEngine

    m, _ = make(name)
    e = PacketEngine(m, seed=12345)
    e.upload_iteration_inputs()
    if native:
        e.set_option("defer_fold", 1)
        e.set_option("exchange_dense", 1 if dense else 0)
        e.comm_init(e.comm_unique_id())
    if sed_local:
        e.set_sed_local(True)
    cnt = e.lucy_transport([n] * m.nStars)
    out = [e.fetch(iG, want=_want(m)) for iG in range(1, m.nGrids + 1)]
    sed = e.fetch_sed()
    plane = e.plane_distribution() if m.lgPlaneIonization else None
    info = getattr(e, "last_exchange", None)
    e.close()
    return out, sed, plane, info, cnt


@pytest.mark.parametrize("name,dense", [("multigrid_sym", False), ("multigrid_sym", True), ("cube_clumpy_gasdust", False),
                                        ("hii_sym_gas_debug", False), ("viewing_angles", False),
                                        ("plane_slab_gasdust", False)])
def test_one_rank_exchange_is_the_identity(cuda_lib, name, dense):
    from cases import make

    m, _ = make(name)
    n = 20001
    ref, sed0, plane0, _, c0 = _run(name, n, native=False)
    got, sed1, plane1, info, c1 = _run(name, n, native=True, dense=dense)
    assert info["nccl_version"] >= 20000, info
    assert info["bytes"] > 0
    assert info["sparse_grids"] == (0 if dense else m.nGrids)
    for iG in range(m.nGrids):
        for k in _want(m):
            assert np.array_equal(got[iG][k].view(np.uint32), ref[iG][k].view(np.uint32)), (iG, k)
    assert np.array_equal(sed1[0].view(np.uint32), sed0[0].view(np.uint32))
    assert np.array_equal(sed1[1], sed0[1])
    if plane0 is not None:
        assert np.array_equal(plane0, plane1)
    assert [c["nSegments"] for c in c0] == [c["nSegments"] for c in c1]


def test_one_rank_exchange_sed_local(cuda_lib):
    ref, sed0, _, _, _ = _run("viewing_angles", 20001, native=False)
    got, sed1, _, info, _ = _run("viewing_angles", 20001, native=True, sed_local=True)
    assert info["sparse_grids"] == 0            # the (nu, angle) counts travelled instead of the per-cell array
    assert np.array_equal(got[0]["Jste"], ref[0]["Jste"])
    assert np.array_equal(got[0]["escapedPackets"], ref[0]["escapedPackets"])    # one rank: local = global
    assert np.array_equal(sed1[0].view(np.uint32), sed0[0].view(np.uint32))
    assert np.array_equal(sed1[1], sed0[1])


def test_exchange_without_communicator_is_an_error(cuda_lib):
    from cases import make
    from mocassin_b200.api import MocassinError, PacketEngine

    m, _ = make("hii_sym_gas")
    e = PacketEngine(m, rank=0, nranks=2, seed=12345)     # rank 0 of 2: tallies stay pending
    e.upload_iteration_inputs()
    e.zero_estimators()
    e.energyPacketDriver(1, 2000)
    with pytest.raises(MocassinError) as ei:
        e.exchange()
    assert "mcb200_comm_init" in str(ei.value)
    e.close()


def _worker(rank, world, idfile, name, n, q, allreduce=False):
    try:
        _worker_body(rank, world, idfile, name, n, q, allreduce)
    except BaseException as ex:              # a dead worker must not leave the parent waiting on the queue
        import traceback

        q.put((rank, "error: " + "".join(traceback.format_exception(type(ex), ex, ex.__traceback__))[-3000:], None))
        raise


def _worker_body(rank, world, idfile, name, n, q, allreduce=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import make
    from mocassin_b200.api import PacketEngine

    m, _ = make(name)
    e = PacketEngine(m, device=rank, rank=rank, nranks=world, seed=12345)
    e.upload_iteration_inputs()
    if rank == 0:                                          # the host's MPI_BCAST of 128 bytes
        with open(idfile + ".tmp", "wb") as f:
            f.write(e.comm_unique_id())
        os.replace(idfile + ".tmp", idfile)
    t0 = time.time()
    while not os.path.exists(idfile):
        if time.time() - t0 > 120:
            raise TimeoutError("no unique id")
        time.sleep(0.05)
    with open(idfile, "rb") as f:
        e.comm_init(f.read())
    if allreduce == "p2p+slabs":
        # the CDF table goes up as one slab of nu-planes per rank and is all-gathered over NVLink
        e.set_option("pdf_slabs", 1)
        e.set_pdfs()
        allreduce = "p2p"
    if allreduce == "p2p-unpacked":
        e.set_option("exchange_pack", 0)           # 64-bit pushes (copy engines) instead of the packed push kernel
        allreduce = "p2p"
    if allreduce == "p2p-kernelpush":
        e.set_option("exchange_pack", 0)
        e.set_option("exchange_push", 2)
        allreduce = "p2p"
    if allreduce == "p2p-pull":
        e.set_option("exchange_push", 0)           # the owner pulls the partial sums with peer loads
        allreduce = "p2p"
    if allreduce == "allreduce" or allreduce is True:
        e.set_option("exchange_allreduce", 1)
    elif allreduce == "nccl":
        e.set_option("exchange_p2p", 0)
    elif allreduce == "p2p" and not m.lgDebug:
        e.set_option("exchange_p2p", 1)            # required: an error if peer memory cannot be mapped
    e.lucy_transport([n] * m.nStars)
    want_path = {"allreduce": "nccl all-reduce", True: "nccl all-reduce", "nccl": "nccl reduce-scatter + all-gather",
                 "p2p": "nccl reduce-scatter + all-gather" if m.lgDebug else "fused peer-memory kernel (NVLink)"}.get(allreduce)
    if want_path:
        assert e.last_exchange["path"] == want_path, e.last_exchange
    out = [e.fetch(iG, want=_want(m)) for iG in range(1, m.nGrids + 1)]
    sums = [e.checksum(iG, w) for iG in range(1, m.nGrids + 1) for w in (0, 1)]
    # the same packets by this rank alone (option solo) on the same context: what bench.py's
    # nrank_parity does at the full size
    e.set_option("solo", 1)
    e.lucy_transport([n] * m.nStars)
    solo = [e.checksum(iG, w) for iG in range(1, m.nGrids + 1) for w in (0, 1)]
    e.set_option("solo", 0)
    assert sums == solo, (rank, sums, solo)
    q.put((rank, out, e.last_exchange))
    e.comm_destroy()
    e.close()


@pytest.mark.parametrize("name,allreduce", [("multigrid_sym", "p2p"), ("cube_clumpy_gasdust", "p2p"), ("hii_sym_gas_debug", "p2p"),
                                            ("viewing_angles", "p2p"), ("multigrid_sym", "allreduce"), ("multigrid_sym", "nccl"), ("multigrid_sym", "p2p+slabs"),
                                            ("multigrid_sym", "p2p-pull"), ("cube_clumpy_gasdust", "p2p-pull"),
                                            ("multigrid_sym", "p2p-unpacked"), ("cube_clumpy_gasdust", "p2p-unpacked"),
                                            ("viewing_angles", "p2p-kernelpush"), ("dust_shell_hg", "p2p-unpacked"),
                                            ("cube_clumpy_gasdust", "p2p+slabs"), ("dust_shell_hg", "p2p+slabs"),
                                            ("cube_clumpy_gasdust", "nccl"), ("hii_sym_gas_debug", "nccl"),
                                            ("multigrid_nonsym", False), ("plane_slab_gasdust", False)])
def test_two_ranks_native_exchange_matches_single_gpu(cuda_lib, name, allreduce):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from cases import make

    m, _ = make(name)
    n = 20001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        idfile = os.path.join(d, "nccl_id")
        procs = [ctx.Process(target=_worker, args=(r, 2, idfile, name, n, q, allreduce)) for r in range(2)]
        for p in procs:
            p.start()
        got = {}
        try:
            for _ in range(2):
                r, out, info = q.get(timeout=240)
                assert not isinstance(out, str), out
                got[r] = out
                assert info["bytes"] > 0
            for p in procs:
                p.join(timeout=120)
                assert p.exitcode == 0
        finally:
            for p in procs:                  # a rank stuck in a collective behind a dead peer
                if p.is_alive():
                    p.kill()
    ref, _, _, _, _ = _run(name, n, native=False)
    for iG in range(m.nGrids):
        for r in (0, 1):
            for k in _want(m):
                assert np.array_equal(got[r][iG][k], ref[iG][k]), (iG, r, k)


def test_fetch_cells_is_the_round_robin_slice_of_jste(cuda_lib):
    """mcb200_fetch_estimators_cells: the rows of cells r+1, r+1+N, ... (iteration_mod.f90:832), compact."""
    from cases import make
    from mocassin_b200.api import PacketEngine

    m, _ = make("cube_clumpy_gasdust")
    e = PacketEngine(m, seed=12345)
    e.upload_iteration_inputs()
    e.lucy_transport([20001])
    J = e.fetch(1)["Jste"]
    nC = m.grids[0].nCells
    for first, stride in ((1, 1), (1, 2), (2, 2), (3, 8), (8, 8), (nC, 8), (nC + 1, 8)):
        got = e.fetch_cells(1, first, stride)
        want = J[first::stride] if first <= nC else J[:0]
        assert got.shape == want.shape and np.array_equal(got, want), (first, stride)
    e.close()


def test_solo_option_and_checksum(cuda_lib):
    """A context created as rank 1 of 2 transports every packet itself under option solo and folds at
    once; its estimators (and their checksums) equal a plain single-rank run."""
    from cases import make
    from mocassin_b200.api import PacketEngine

    m, _ = make("cube_clumpy_gasdust")
    n = 20001
    ref = PacketEngine(m, seed=12345)
    ref.upload_iteration_inputs()
    ref.lucy_transport([n])
    want = ref.fetch(1)
    sums = [ref.checksum(1, 0), ref.checksum(1, 1)]
    ref.close()
    e = PacketEngine(m, rank=1, nranks=2, seed=12345)
    e.upload_iteration_inputs()
    e.set_option("solo", 1)
    e.zero_estimators()
    c = e.energyPacketDriver(1, n)
    assert c["nPackets"] == n
    got = e.fetch(1)
    assert np.array_equal(got["Jste"], want["Jste"]) and np.array_equal(got["escapedPackets"], want["escapedPackets"])
    assert [e.checksum(1, 0), e.checksum(1, 1)] == sums
    # a changed element changes the sum; a zero array sums to zero
    e.zero_estimators()
    assert e.checksum(1, 0) == 0
    e.set_option("solo", 0)
    e.zero_estimators()
    c = e.energyPacketDriver(1, n)
    assert c["nPackets"] == n // 2                      # rank 1 of 2 again
    e.close()


def test_exchange_twice_or_transport_after_exchange_is_refused(cuda_lib):
    from cases import make
    from mocassin_b200.api import MocassinError, PacketEngine

    m, _ = make("hii_sym_gas")
    e = PacketEngine(m, seed=12345)
    e.upload_iteration_inputs()
    e.set_option("defer_fold", 1)
    e.comm_init(e.comm_unique_id())
    e.zero_estimators()
    e.energyPacketDriver(1, 3000)
    e.exchange()
    with pytest.raises(MocassinError):
        e.exchange()                         # would sum the other ranks' tallies twice
    with pytest.raises(MocassinError):
        e.energyPacketDriver(1, 3000)        # new tallies on top of already-global ones
    e._check(e.lib.mcb200_reduce(e.h))
    e.energyPacketDriver(1, 3000)            # fine again after the fold
    e.reduce()
    e.close()
