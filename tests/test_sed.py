"""SED reduction (head of writeSED, output_mod.f90:2561-2568) and the host post-processing
(:2626-2660).  CPU part: oracle counts vs the per-cell float tallies, conservation of the
energy budget, file layout.  GPU part: device reduction == oracle, bit for bit."""
import os

import numpy as np
import pytest

from cases import make
from mocassin_b200 import workloads as W
from mocassin_b200.output import sed_from_raw, write_sed
from oracle.oracle import Oracle

F32 = np.float32
SEED = 12345


def _run_oracle(name):
    m, n = make(name)
    o = Oracle(m)
    c, _ = o.transport(1, 0, n, seed=SEED)
    return m, n, o, c


@pytest.mark.parametrize("name", ["dust_shell_hg", "viewing_angles", "multigrid_sym"])
def test_oracle_sed_is_the_cell_sum_of_escaped_packets(name):
    m, n, o, c = _run_oracle(name)
    dE = float(m.deltaE[1])
    cnt, sed = o.sed(dE)
    assert cnt.shape == (m.nbins, m.nAngleBins + 1)
    tot = np.zeros((m.nbins, m.nAngleBins + 1))
    for iG in range(1, m.nGrids + 1):
        tot += o.folded(iG, dE)["escapedPackets"].astype(np.float64).sum(axis=0)[1:, :]
    assert np.allclose(sed, tot, rtol=2e-6, atol=0)
    assert cnt[:, 0].sum() == c["nEscaped"]
    if m.nAngleBins:
        assert np.all(cnt[:, 1:].sum(axis=1) <= cnt[:, 0])


def test_sed_energy_budget_dust_shell(tmp_path):
    """No gas: every packet escapes, so writeSED's totalE equals Lstar (the reference's own
    check, output_mod.f90:2693-2694)."""
    m, n, o, c = _run_oracle("dust_shell_hg")
    dE = float(m.deltaE[1])
    cnt, raw = o.sed(dE)
    wid = W.wid_flx(m.nuArray)
    sed, totalE = sed_from_raw(m, wid, raw)
    assert c["nEscaped"] == n
    assert abs(totalE - n * dE) <= 1e-5 * n * dE
    assert np.all(sed >= 0) and sed[:, 0].max() > 0
    # Jy conversion of column 0: F = 1e23 * E / (4 pi 3.08^2) / (3.2898e15 widFlx)
    f = int(np.argmax(raw[:, 0]))
    expect = 1e23 * float(raw[f, 0]) / (4 * np.pi * 3.08 ** 2) / (3.2898e15 * float(wid[f]))
    assert abs(sed[f, 0] - expect) <= 1e-5 * expect
    path = os.path.join(tmp_path, "SED.out")
    te = write_sed(path, m, wid, raw)
    lines = open(path).read().splitlines()
    assert len(lines) == 4 + m.nbins + 4 and "Total energy radiated" in lines[4 + m.nbins + 1]
    assert te == totalE
    first = lines[4].split()
    assert len(first) == 2 + m.nAngleBins + 1 and abs(float(first[0]) - float(m.nuArray[0])) < 1e-6 * float(m.nuArray[0])


def test_sed_viewing_angle_columns():
    m, n, o, c = _run_oracle("viewing_angles")
    cnt, raw = o.sed(float(m.deltaE[1]))
    sed, _ = sed_from_raw(m, W.wid_flx(m.nuArray), raw)
    assert sed.shape == (m.nbins, m.nAngleBins + 1)
    at = m.angle_tables()
    for imu in range(1, m.nAngleBins + 1):
        th1 = F32(int(F32(m.viewPointTheta[imu]) / F32(at["dTheta"]))) * F32(at["dTheta"])
        solid = float(at["dPhi"]) * 3.08 ** 2 * abs(np.cos(float(th1)) - np.cos(float(th1) + float(at["dTheta"])))
        f = int(np.argmax(raw[:, imu]))
        if raw[f, imu] > 0:
            mult = 4.0 / 8.0 if m.lgSymmetricXYZ else 1.0
            expect = 1e23 * float(raw[f, imu]) * mult / solid / (3.2898e15 * float(W.wid_flx(m.nuArray)[f]))
            assert abs(sed[f, imu] - expect) <= 2e-5 * expect


# ------------------------------------------------------------------------------- GPU
def _engine(m, **kw):
    from mocassin_b200.api import PacketEngine
    e = PacketEngine(m, seed=SEED, **kw)
    e.upload_iteration_inputs()
    return e


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dust_shell_hg", "viewing_angles", "multigrid_sym", "hii_sym_gas", "cube_clumpy_gasdust"])
def test_device_sed_matches_oracle(name):
    m, n, o, c = _run_oracle(name)
    e = _engine(m)
    e.zero_estimators()
    e.energyPacketDriver(1, n)
    e.reduce()
    sed, cnt = e.fetch_sed()
    wc, ws = o.sed(float(m.deltaE[1]))
    assert np.array_equal(cnt, wc)
    assert np.array_equal(sed.view(np.uint32), ws.view(np.uint32))
    assert cnt.sum() > 0
    # and it is the cell sum of what fetch returns
    tot = sum(e.fetch(iG)["escapedPackets"].astype(np.float64).sum(axis=0)[1:, :] for iG in range(1, m.nGrids + 1))
    assert np.allclose(sed, tot, rtol=2e-6, atol=0)
    e.close()


@pytest.mark.gpu
def test_device_sed_accumulates_over_calls_and_zeroes():
    m, n, o, c = _run_oracle("dust_shell_hg")
    e = _engine(m)
    e.zero_estimators()
    e.energyPacketDriver(1, n)
    e.energyPacketDriver(1, n, deltaE=float(m.deltaE[1]) * 0.5)
    sed, cnt = e.fetch_sed()
    wc, ws = o.sed(float(m.deltaE[1]))
    assert np.array_equal(cnt, 2 * wc)           # same seed, same packets twice
    half = (wc.astype(F32) * F32(float(m.deltaE[1]) * 0.5)).astype(F32)
    assert np.array_equal(sed, (ws + half).astype(F32))
    e.zero_estimators()
    sed, cnt = e.fetch_sed()
    assert not sed.any() and not cnt.any()
    e.close()


@pytest.mark.gpu
def test_sed_local_on_two_ranks_of_one_gpu():
    """sed_local: each rank tallies its own escapes per (nu, angle); summing the two ranks'
    buffer-6 counts by hand (what the all-reduce does) gives the single-rank SED exactly."""
    import ctypes as C
    m, n, o, c = _run_oracle("viewing_angles")
    wc, ws = o.sed(float(m.deltaE[1]))
    engines = [_engine(m, rank=r, nranks=2) for r in range(2)]
    import torch
    from mocassin_b200.api import _as_cuda_tensor
    bufs = []
    for e in engines:
        e.set_sed_local(True)
        e.zero_estimators()
        e.energyPacketDriver(1, n)
        ptr, cnt = e.tally_buffer(1, 6)
        bufs.append(_as_cuda_tensor(ptr, cnt, "<i8", 0))
    total = bufs[0] + bufs[1]
    for b in bufs:
        b.copy_(total)
    torch.cuda.synchronize()
    for e in engines:
        e._check(e.lib.mcb200_reduce(e.h))
        sed, cnt = e.fetch_sed()
        assert np.array_equal(cnt, wc)
        assert np.array_equal(sed.view(np.uint32), ws.view(np.uint32))
    # escapedPackets stays rank-local: the two ranks' arrays add up to the global one
    loc = [e.fetch(1)["escapedPackets"].astype(np.float64) for e in engines]
    glob = o.folded(1, float(m.deltaE[1]))["escapedPackets"].astype(np.float64)
    assert np.allclose(loc[0] + loc[1], glob, rtol=1e-6)
    for e in engines:
        e.close()


def test_plane_ion_distribution_file(tmp_path):
    """planeIonDistribution.out: one `i k count` record per (x, z) column, x outermost
    (iteration_mod.f90:570-577)."""
    from mocassin_b200 import output

    p = np.arange(12, dtype=np.int32).reshape(3, 4, order="F")
    output.write_plane_ion_distribution(str(tmp_path / "p.out"), p)
    rows = [[int(x) for x in ln.split()] for ln in open(tmp_path / "p.out")]
    assert len(rows) == 12 and rows[0] == [1, 1, 0] and rows[1] == [1, 2, 3] and rows[-1] == [3, 4, 11]
