"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Bar: bit exact -- per-packet fates, every counter, every integer
tally and the folded float32 estimators are identical."""
import ctypes as C

import numpy as np
import pytest

from cases import CASES, make
from mocassin_b200 import _lib
from mocassin_b200 import workloads as W
from mocassin_b200.api import MocassinError, PacketEngine, partition
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

SEED = 12345


def _engine(m, **kw):
    e = PacketEngine(m, seed=SEED, **kw)
    e.upload_iteration_inputs()
    return e


@pytest.fixture(scope="module")
def ctx(cuda_lib):
    h = C.c_void_p()
    assert cuda_lib.mcb200_create(C.byref(h), 0, 0, 1, C.c_uint64(1)) == 0
    yield h
    cuda_lib.mcb200_destroy(h)


@pytest.mark.parametrize("which,lo,hi", [(0, 2.0 ** -24, 1.0), (1, -7.0, 7.0), (2, -7.0, 7.0), (3, -1.0, 1.0), (4, -1e4, 1e4),
                                         (5, -100.0, 88.0)])
def test_device_detmath_is_bit_identical_to_oracle(cuda_lib, oracle_lib, ctx, which, lo, hi):
    rng = np.random.default_rng(which + 10)
    x = rng.uniform(lo, hi, 1 << 20).astype(np.float32)
    x[:8] = np.float32([lo, hi, 1.0, 0.5, 0.70710677, 0.41421357, 0.99999994, 2.0 ** -24])[:8].clip(lo, hi)
    a, b = np.zeros_like(x), np.zeros_like(x)
    fp = _lib.c_float_p
    assert cuda_lib.mcb200_test_detmath(ctx, which, x.ctypes.data_as(fp), a.ctypes.data_as(fp), x.shape[0]) == 0
    oracle_lib.oracle_detmath(which, x.ctypes.data_as(fp), b.ctypes.data_as(fp), x.shape[0])
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_device_division_is_ieee(cuda_lib, ctx):
    """The transport divides wall distances by direction cosines with a reciprocal-multiply
    plus one FMA correction (div_rn, transport_core.cuh).  It must equal IEEE division -- what
    the oracle and the reference do -- for every operand pair."""
    rng = np.random.default_rng(99)
    n = 1 << 22
    fp = _lib.c_float_p
    for trial in range(3):
        x = np.empty(n, np.float32)
        # even slots: numerators over 2^-40..2^69 (both signs, some zeros); odd: cosines 1e-10..1
        x[0::2] = (np.exp2(rng.uniform(-40, 69, n // 2)) * rng.choice([-1.0, 1.0], n // 2)).astype(np.float32)
        x[0::2][rng.integers(0, n // 2, 1000)] = 0.0
        v = (np.exp2(rng.uniform(-33.2, 0, n // 2)) * rng.choice([-1.0, 1.0], n // 2)).astype(np.float32)
        edge = rng.integers(0, n // 2, 4000)
        v[edge] = (np.float32(1.0) - np.float32(2.0 ** -24) * rng.integers(0, 4, 4000)).astype(np.float32) * np.exp2(rng.integers(-30, 1, 4000)).astype(np.float32)
        x[1::2] = v
        a, b = np.zeros_like(x), np.zeros_like(x)
        assert cuda_lib.mcb200_test_detmath(ctx, 6, x.ctypes.data_as(fp), a.ctypes.data_as(fp), n) == 0
        assert cuda_lib.mcb200_test_detmath(ctx, 7, x.ctypes.data_as(fp), b.ctypes.data_as(fp), n) == 0
        ev = slice(0, n, 2)                      # numerator / cosine pairs
        assert np.array_equal(a[ev].view(np.uint32), b[ev].view(np.uint32))
        with np.errstate(over="ignore"):
            want = (x[0::2] / x[1::2]).astype(np.float32)
        assert np.array_equal(b[ev].view(np.uint32), want.view(np.uint32))


def test_set_grid_rejects_axes_outside_the_division_range(cuda_lib, ctx):
    m, _ = make("hii_sym_gas")
    e = _engine(m)
    g = m.grids[0]
    for bad in (1.0e28, 1.0e-12):
        ax = g.xAxis.copy()
        ax[-1 if bad > 1 else 1] = np.float32(bad)
        if bad < 1:
            ax[0] = 0.0
        fp, ip = _lib.c_float_p, _lib.c_int32_p
        act = np.asfortranarray(g.active, dtype=np.int32)
        rc = cuda_lib.mcb200_set_grid(e.h, 1, g.nx, g.ny, g.nz, g.nCells, g.motherP, np.sort(ax).ctypes.data_as(fp),
                                      g.yAxis.ctypes.data_as(fp), g.zAxis.ctypes.data_as(fp), act.ctypes.data_as(ip))
        assert rc == -6, rc            # MCB200_EUNSUPPORTED
    e.close()


def test_device_philox_stream_is_identical_to_oracle(cuda_lib, oracle_lib, ctx):
    fp = _lib.c_float_p
    for seed, pid, stream in [(12345, 0, 1), (2 ** 40 + 3, 2 ** 33 + 17, 0), (0, 999999937, 7)]:
        a, b = np.zeros(1001, np.float32), np.zeros(1001, np.float32)
        assert cuda_lib.mcb200_test_uniforms(ctx, seed, pid, stream, a.shape[0], a.ctypes.data_as(fp)) == 0
        oracle_lib.oracle_uniforms(seed, pid, stream, b.shape[0], b.ctypes.data_as(fp))
        assert np.array_equal(a, b)


def _compare(m, n, e, o, iStar=1):
    e.set_option("trace", 1)
    e.zero_estimators()
    cg = e.energyPacketDriver(iStar, n)
    co, fo = o.transport(iStar, 0, n, seed=SEED, want_fates=True)
    fg = e.fates(n)
    bad = np.flatnonzero((fg != fo).any(axis=1))
    assert bad.size == 0, f"{bad.size} packets differ, first {bad[:5]}: gpu {fg[bad[:5]]} oracle {fo[bad[:5]]}"
    for k in ("nAbs", "nSca", "trapped", "nLinePackets", "nDropped", "nSegments", "nFlights", "nEscaped", "nEarlyEscaped"):
        assert cg[k] == co[k], k
    assert cg["nPackets"] == n
    assert np.array_equal(e.qphot_counts(), o.qphotCounts)
    if m.lgPlaneIonization:
        assert np.array_equal(e.plane_distribution(), o.planeIonDistribution)
        assert o.planeIonDistribution.sum() == n
    dE = float(m.deltaE[iStar])
    want = ["Jste", "escapedPackets"] + (["Jdif", "linePackets"] if m.lgDebug else [])
    for iG in range(1, m.nGrids + 1):
        assert e.len_unit(iG) == np.ldexp(1.0, o.out[iG - 1]["lenExp"])
        fold = o.folded(iG, dE)
        got = e.fetch(iG, want=want)
        for k in want:
            a, b = got[k], fold[k]
            if k in ("Jste", "Jdif"):
                a, b = a[1:], b[1:]
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (iG, k)
        # the fold cleared the integer tallies
        t = e.fetch_tallies(iG)
        assert not t["JsteQ"].any() and not t["escapedQ"].any()
    return cg, co


SCHEDULES = {
    "persistent": dict(wavefront=0, order=0),          # one thread carries a packet through all phases
    "persistent_ordered": dict(wavefront=0, order=1, agg_steps=6),   # + packets sorted by first nu, aggregated tallies
    "wavefront": dict(wavefront=1, tail=64, wave0_order=2),   # per-wave event kernels + nu-sorted FLY kernel; wave 0 pre-ordered
    "wavefront_budget3": dict(wavefront=1, step_budget=3, tail=0, fly_batch=1, wave0_order=0),   # flights continue across many waves; no batching
    "wavefront_tail": dict(wavefront=1, step_budget=5, tail=10 ** 9, fly_batch=32),    # wave 0, then the persistent kernel resumes all
}


@pytest.mark.parametrize("schedule", list(SCHEDULES))
@pytest.mark.parametrize("name", list(CASES))
def test_transport_bit_exact_vs_oracle(name, schedule):
    """Every schedule must give the oracle's answer bit for bit (tallies are
    order-independent integers, every packet owns its Philox stream)."""
    m, n = make(name)
    e = _engine(m)
    for k, v in SCHEDULES[schedule].items():
        e.set_option(k, v)
    o = Oracle(m)
    cg, co = _compare(m, n, e, o)
    # Qphot (float64 from exact counts) agrees with the reference's float32 running sum
    if co["Qphot"] > 0:
        assert abs(cg["Qphot"] / co["Qphot"] - 1) < 1e-3
    e.close()


def test_integer_tallies_before_fold_match_oracle():
    """nranks=2 keeps the integer tallies pending: compare them raw (rows >= 1 of J; row 0
    is the inactive-cell sink the device skips), then reduce."""
    m, n = make("dust_shell_hg")
    e0 = _engine(m, rank=0, nranks=2)
    e1 = _engine(m, rank=1, nranks=2)
    o = Oracle(m, fp32_tallies=False)
    o.transport(1, 0, n, seed=SEED)
    e0.zero_estimators(); e1.zero_estimators()
    c0 = e0.energyPacketDriver(1, n)
    c1 = e1.energyPacketDriver(1, n)
    assert c0["nPackets"] == partition(n, 0, 2)[1] and c1["nPackets"] == partition(n, 1, 2)[1]
    t0, t1 = e0.fetch_tallies(1), e1.fetch_tallies(1)
    assert np.array_equal((t0["JsteQ"] + t1["JsteQ"])[1:], o.out[0]["JsteQ"][1:])
    assert np.array_equal(t0["escapedQ"] + t1["escapedQ"], o.out[0]["escapedQ"])
    with pytest.raises(MocassinError):
        e0.fetch(1)                      # pending tallies must be reduced first
    e0.close(); e1.close()


def test_determinism_and_seed_sensitivity():
    m, n = make("cube_clumpy_gasdust")
    res = []
    for seed, bps in ((SEED, 0), (SEED, 1), (SEED + 1, 0)):
        e = PacketEngine(m, seed=seed)
        e.upload_iteration_inputs()
        if bps:
            e.set_option("blocks_per_sm", bps)      # different schedule, same answer
        e.zero_estimators()
        e.energyPacketDriver(1, n)
        res.append(e.fetch(1))
        e.close()
    assert np.array_equal(res[0]["Jste"], res[1]["Jste"])
    assert np.array_equal(res[0]["escapedPackets"], res[1]["escapedPackets"])
    assert not np.array_equal(res[0]["Jste"], res[2]["Jste"])


def test_two_calls_accumulate_like_the_reference():
    """Two stars' worth of transport calls accumulate into the same estimators
    (iteration_mod.f90:474-496 loops over iStar without zeroing in between)."""
    m, n = make("hii_sym_gas")
    e = _engine(m)
    o = Oracle(m)
    e.zero_estimators()
    e.energyPacketDriver(1, n)
    a = e.fetch(1)
    e.energyPacketDriver(1, n, deltaE=float(m.deltaE[1]) * 0.5)
    b = e.fetch(1)
    o.transport(1, 0, n, seed=SEED)
    f = o.folded(1, float(m.deltaE[1]))
    h = o.folded(1, float(m.deltaE[1]) * 0.5)
    assert np.array_equal(a["Jste"][1:], f["Jste"][1:])
    assert np.array_equal(b["Jste"][1:], (f["Jste"] + h["Jste"]).astype(np.float32)[1:])
    assert np.array_equal(b["escapedPackets"], (f["escapedPackets"] + h["escapedPackets"]).astype(np.float32))
    e.zero_estimators()
    assert not e.fetch(1)["Jste"].any()
    e.close()


def test_diffuse_external_source():
    """energyPacketDriver(iStar=0, n, grid, gpLoc, cellLoc) (iteration_mod.f90:498-550)."""
    m, n = make("cube_clumpy_gasdust")
    m.inSpectrumProbDen[0, :] = W.blackbody_cdf(20000.0, m.nuArray, np.gradient(m.nuArray).astype(np.float32))
    m.deltaE[0] = 3.0e-6
    e = _engine(m)
    o = Oracle(m)
    e.set_option("trace", 1)
    e.set_option("wavefront", 1)
    e.zero_estimators()
    cell = [5, 7, 9]
    cg = e.energyPacketDriver(0, 4000, gpLoc=1, cellLoc=cell)
    co, fo = o.transport(0, 0, 4000, seed=SEED, gpLoc=1, cellLoc=cell, want_fates=True)
    assert np.array_equal(e.fates(4000), fo)
    assert cg["nSegments"] == co["nSegments"]
    got, want = e.fetch(1), o.folded(1, float(m.deltaE[0]))
    assert np.array_equal(got["Jste"][1:], want["Jste"][1:])
    assert np.array_equal(got["escapedPackets"], want["escapedPackets"])
    e.close()


def test_error_behaviour():
    m, n = make("hii_sym_gas")
    e = PacketEngine(m, seed=SEED)
    with pytest.raises(MocassinError) as ei:          # transport before opacities are set
        e.energyPacketDriver(1, 10)
    assert ei.value.code == -3
    e.set_opacity()
    bad = m.grids[0].recPDF.copy(order="F")
    bad[5, 100] = 2.0                                  # non-monotone CDF row
    good = m.grids[0].recPDF
    m.grids[0].recPDF = bad
    with pytest.raises(MocassinError) as ei:
        e.set_pdfs()
    assert ei.value.code == -7
    m.grids[0].recPDF = good
    e.set_pdfs()
    with pytest.raises(MocassinError):
        e.energyPacketDriver(2, 10, deltaE=1.0)        # iStar out of range
    with pytest.raises(MocassinError):
        e.energyPacketDriver(1, -1)
    assert e.energyPacketDriver(1, 0)["nPackets"] == 0  # empty input
    assert e.energyPacketDriver(1, 1)["nPackets"] == 1  # ragged: fewer packets than a warp
    e.close()
    m2 = W.hii_region()
    m2.lgPlaneIonization = True                       # symmetricXYZ + planeIonization: the reference stops
    with pytest.raises(MocassinError) as ei:
        PacketEngine(m2)
    assert ei.value.code == -2


def test_packet_reaching_a_reference_stop_is_reported():
    """nuP >= nbins is fatal in the reference (photon_mod.f90:826-832): a stellar CDF that
    reaches 1 only in its last bin can return nbins-1 at most, so force the condition with
    a CDF that is 0 everywhere (u >= cdf for all bins -> nuP = nbins)."""
    m, n = make("hii_sym_gas")
    m.inSpectrumProbDen[1, :] = 0.0
    e = _engine(m)
    with pytest.raises(MocassinError) as ei:
        e.energyPacketDriver(1, 100)
    assert ei.value.code == -5
    o = Oracle(m)
    with pytest.raises(RuntimeError):
        o.transport(1, 0, 100, seed=SEED)
    e.close()


def test_async_pdf_upload_gives_same_answer_and_reports_bad_tables():
    """option async_pdfs: set_pdfs only enqueues upload+transpose on a copy stream (overlapping
    the stellar wave); results are unchanged and a non-monotone table is reported by the
    next transport call."""
    m, n = make("cube_clumpy_gasdust")
    n = 200000                                   # wave-front path
    ref = _engine(m)
    ref.zero_estimators(); ref.energyPacketDriver(1, n)
    want = ref.fetch(1); ref.close()
    e = PacketEngine(m, seed=SEED)
    e.set_option("async_pdfs", 1)
    e.set_opacity(); e.set_dust_state(); e.set_pdfs()
    e.zero_estimators(); e.energyPacketDriver(1, n)
    got = e.fetch(1)
    assert np.array_equal(got["Jste"], want["Jste"]) and np.array_equal(got["escapedPackets"], want["escapedPackets"])
    good = m.grids[0].recPDF
    bad = good.copy(order="F"); bad[7, 50] = 3.0
    m.grids[0].recPDF = bad
    e.set_pdfs()                                 # accepted: verdict is deferred
    with pytest.raises(MocassinError) as ei:
        e.energyPacketDriver(1, n)
    assert ei.value.code == -7
    m.grids[0].recPDF = good
    e.set_pdfs(); e.zero_estimators(); e.energyPacketDriver(1, n)
    assert np.array_equal(e.fetch(1)["Jste"], want["Jste"])
    e.close()


@pytest.mark.parametrize("wavefront", [0, 1])
def test_resonance_line_packet_transfer(wavefront):
    """Second half of energyPacketDriver (photon_mod.f90:180-266): resLinePackets(cell) diffuse
    packets from every cell centre, cells dealt round-robin to ranks."""
    m, _ = make("multigrid_sym")
    rng = np.random.default_rng(4)
    for g in m.grids:
        r = rng.integers(0, 5, g.nCells + 1).astype(np.int32)
        r[0] = 0
        g.resLinePackets = r
    o = Oracle(m)
    co, nrun = o.transport_reslines(1, seed=SEED)
    e = _engine(m)
    e.set_option("wavefront", wavefront)
    e.zero_estimators()
    cg = e.resLinePacketsTransfer(1)
    assert cg["nPackets"] == nrun == sum(int(g.resLinePackets.sum()) for g in m.grids)
    for k in ("nAbs", "nSca", "nSegments", "nEscaped", "nLinePackets", "nFlights"):
        assert cg[k] == co[k], k
    for iG in (1, 2):
        got, want = e.fetch(iG), o.folded(iG, float(m.deltaE[1]))
        assert np.array_equal(got["Jste"][1:], want["Jste"][1:])
        assert np.array_equal(got["escapedPackets"], want["escapedPackets"])
    e.close()
    # two ranks, one after the other on this GPU: integer tallies add up to the oracle's
    parts = []
    for r in range(2):
        e = _engine(m, rank=r, nranks=2)
        e.set_option("wavefront", wavefront)
        e.zero_estimators()
        e.resLinePacketsTransfer(1)
        parts.append([e.fetch_tallies(iG) for iG in (1, 2)])
        e.close()
    for iG in (0, 1):
        assert np.array_equal((parts[0][iG]["JsteQ"] + parts[1][iG]["JsteQ"])[1:], o.out[iG]["JsteQ"][1:])
        assert np.array_equal(parts[0][iG]["escapedQ"] + parts[1][iG]["escapedQ"], o.out[iG]["escapedQ"])


@pytest.mark.parametrize("name", ["viewing_angles", "multigrid_sym", "hii_sym_gas", "dust_shell_hg"])
def test_sparse_escaped_fetch_equals_dense_fetch(name):
    """mcb200_fetch_escaped_sparse writes exactly the non-zero entries of escapedPackets: the
    array it fills equals the dense fetch, also when the same array is reused for a second,
    different call (clear_previous), and without clear_previous into a zeroed array."""
    m, n = make(name)
    e = _engine(m)
    e.zero_estimators()
    e.energyPacketDriver(1, n)
    bufs = {}
    for iG in range(1, m.nGrids + 1):
        dense = e.fetch(iG)["escapedPackets"]
        sparse, k = e.fetch_escaped_sparse(iG)
        bufs[iG] = sparse
        assert k == int(np.count_nonzero(dense)) or k == -1
        assert np.array_equal(sparse.view(np.uint32), dense.view(np.uint32)), iG
    assert any(np.count_nonzero(b) for b in bufs.values())
    # a second, different call into the same arrays
    e.zero_estimators()
    e.energyPacketDriver(1, max(n // 7, 50), deltaE=float(m.deltaE[1]) * 3.0)
    for iG in range(1, m.nGrids + 1):
        dense = e.fetch(iG)["escapedPackets"]
        again, k = e.fetch_escaped_sparse(iG, out=bufs[iG], clear_previous=True)
        assert again is bufs[iG] and np.array_equal(again.view(np.uint32), dense.view(np.uint32)), iG
        fresh, _ = e.fetch_escaped_sparse(iG, clear_previous=False)
        assert np.array_equal(fresh.view(np.uint32), dense.view(np.uint32))
        both, _ = e.fetch_sparse(iG)                      # Jste dense + escapedPackets sparse in one call
        assert np.array_equal(both["escapedPackets"].view(np.uint32), dense.view(np.uint32))
        assert np.array_equal(both["Jste"].view(np.uint32), e.fetch(iG)["Jste"].view(np.uint32))
    e.close()
