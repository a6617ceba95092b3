"""Photo-rate pre-integration (update_mod.f90:170-262, :1160-1214): oracle
checks on CPU against float64 numpy, device == oracle bit for bit on GPU."""
import numpy as np
import pytest

from cases import make
from mocassin_b200.api import scale_estimators
from oracle import oracle as O
from oracle.oracle import Oracle

F32 = np.float32
HCRYD = 2.1799153e-11


def _bands(nbins, rng, nBands=9):
    low = rng.integers(1, nbins - 5, nBands).astype(np.int32)
    high = np.minimum(low + rng.integers(0, nbins, nBands), nbins + rng.integers(0, 3, nBands)).astype(np.int32)
    high[0] = nbins + 7                       # clipped to nbins like min(highNuP, nbins)
    low[1], high[1] = nbins, nbins            # single-bin band
    xs = [np.zeros(3, F32)]
    off = np.zeros(nBands, np.int32)
    pos = 4
    for b in range(nBands):
        n = min(int(high[b]), nbins) - int(low[b]) + 1
        x = (1e-18 * rng.lognormal(0.0, 2.0, n) * (np.arange(1, n + 1) ** -3.0) * 50).astype(F32)
        if b % 3 == 2 and n > 4:
            x[n // 2] = F32(1e-37)            # thermBalance leaves its loop here; updateCell treats it as 0
        off[b] = pos
        xs.append(x)
        pos += n
    return off, low, high, np.concatenate(xs).astype(F32)


def _ref64(nbins, off, low, high, xs, nu, J):
    nR = J.shape[0]
    nP = np.zeros((nR, len(off)))
    ht = np.zeros((nR, len(off)))
    for b in range(len(off)):
        hi = min(int(high[b]), nbins)
        j = np.arange(int(low[b]), hi + 1)
        x = xs[off[b] - 1 + (j - low[b])].astype(np.float64)
        stop = np.nonzero(x < 1e-35)[0]
        x0 = np.where(x < 1e-35, 0.0, x)
        Jb = np.maximum(J[:, j - 1].astype(np.float64), 0.0)
        nuj = nu[j - 1].astype(np.float64)
        nP[:, b] = 1e-20 + (Jb * x0 / (HCRYD * nuj)).sum(axis=1)
        k = stop[0] if len(stop) else len(j)
        ht[:, b] = (x0[:k] * Jb[:, :k] * (nuj[:k] - nuj[0]) / nuj[:k]).sum(axis=1)
    return nP, ht


@pytest.fixture(scope="module")
def run():
    m, n = make("hii_sym_gas")
    o = Oracle(m)
    o.transport(1, 0, n, seed=12345)
    J = o.folded(1, float(m.deltaE[1]))["Jste"]
    Js, _ = scale_estimators(m, J, np.zeros((1, 1, 1), F32))
    rng = np.random.default_rng(17)
    return m, n, Js, _bands(m.nbins, rng)


def test_oracle_photo_integrals_against_float64(run):
    m, n, Js, (off, low, high, xs) = run
    nP, ht = O.photo_integrals(m.nbins, off, low, high, xs, m.nuArray, Js)
    rP, rH = _ref64(m.nbins, off, low, high, xs, m.nuArray, Js)
    assert nP.shape == (m.grids[0].nCells + 1, len(off))
    assert np.allclose(nP, rP, rtol=2e-5, atol=0)
    assert np.allclose(ht, rH, rtol=2e-5, atol=1e-30)
    assert np.all(nP[0] == F32(1e-20)) and np.all(ht[0] == 0)     # cell 0 never holds Jste
    assert (nP[1:] > 1e-20).any() and (ht[1:] > 0).any()
    # the heating sum stops at the tiny cross-section, the rate sum does not
    xs_open = xs.copy()
    xs_open[3:][xs[3:] < 1e-35] = F32(1e-34)          # not below the threshold: no early exit
    full = _ref64(m.nbins, off, low, high, xs_open, m.nuArray, Js)[1]
    stopped = [b for b in range(len(off)) if b % 3 == 2]
    assert (full[:, stopped] >= rH[:, stopped] * (1 - 1e-12)).all()
    assert (full[:, stopped] > rH[:, stopped] * 1.001).any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hii_sym_gas", "hii_sym_gas_debug", "cube_uniform_gas"])
def test_device_photo_integrals_match_oracle(name):
    from mocassin_b200.api import PacketEngine
    m, n = make(name)
    rng = np.random.default_rng(5)
    off, low, high, xs = _bands(m.nbins, rng, nBands=9 if name != "cube_uniform_gas" else 130)
    e = PacketEngine(m, seed=12345)
    e.upload_iteration_inputs()
    e.set_xsec(xs)
    e.zero_estimators()
    e.energyPacketDriver(1, n)
    e.reduce()
    got = e.photo_integrals(1, off, low, high, dif=m.lgDebug)
    f = e.fetch(1, want=["Jste", "escapedPackets"] + (["Jdif"] if m.lgDebug else []))
    Js, _ = scale_estimators(m, f["Jste"], np.zeros((1, 1, 1), F32))
    wP, wH = O.photo_integrals(m.nbins, off, low, high, xs, m.nuArray, Js)
    assert np.array_equal(got["nPhotoSte"].view(np.uint32), wP.view(np.uint32))
    assert np.array_equal(got["heatSte"].view(np.uint32), wH.view(np.uint32))
    assert (wP[1:] > 1e-20).any()
    if m.lgDebug:
        Jd, _ = scale_estimators(m, f["Jdif"], np.zeros((1, 1, 1), F32))
        dP, dH = O.photo_integrals(m.nbins, off, low, high, xs, m.nuArray, Jd)
        assert np.array_equal(got["nPhotoDif"].view(np.uint32), dP.view(np.uint32))
        assert np.array_equal(got["heatDif"].view(np.uint32), dH.view(np.uint32))
    e.close()


@pytest.mark.gpu
def test_photo_integrals_error_behaviour():
    from mocassin_b200.api import MocassinError, PacketEngine
    m, n = make("hii_sym_gas")
    e = PacketEngine(m, seed=1)
    e.upload_iteration_inputs()
    one = np.ones(1, np.int32)
    with pytest.raises(MocassinError):           # set_xsec missing
        e.photo_integrals(1, one, one, one)
    e.set_xsec(np.ones(10, F32))
    with pytest.raises(MocassinError):           # band runs past the end of xSecArray
        e.photo_integrals(1, one * 5, one, one * 100)
    with pytest.raises(MocassinError):           # Jdif without debug mode
        e.photo_integrals(1, one, one, one, dif=True)
    e.close()
