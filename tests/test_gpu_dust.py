"""GPU parity of the dust-only closure (K5 getDustT/updateCell, K6 setDustPDF) against the
oracle, through the C ABI.  Bit-exact: float32 with the reference's operation order."""
import numpy as np
import pytest

from mocassin_b200 import workloads as W
from mocassin_b200.api import MocassinError, PacketEngine, scale_estimators
from oracle import oracle as O

pytestmark = pytest.mark.gpu
F32 = np.float32


def _engine(model, t, **kw):
    eng = PacketEngine(model, **kw)
    eng.set_xsec(t["xSecArray"])
    eng.set_dust_tables(t["widFlx"], t["grainWeight"], t["dustAbsXsecP"], t["dustEmIntegral"])
    eng.set_opacity()
    eng.set_dust_state()
    return eng


def _same(a, b):
    """Bit-identical, except that a NaN matches a NaN (0/0 has a different payload and sign
    on x86 and on the GPU; the reference produces NaN rows for cells whose grains all sublimed)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | both_nan))


@pytest.mark.parametrize("multi", [True, False])
def test_dust_pdf_matches_oracle(multi):
    model, t = W.dust_closure(n=9, nbins=120, multiChem=multi)
    g = model.grids[0]
    rng = np.random.default_rng(5)
    g.Tdust[1:, 1:, 1:] = rng.uniform(5.0, 1500.0, size=g.Tdust[1:, 1:, 1:].shape).astype(F32)
    g.Tdust[1, 1, 1:4] = F32(0.0)
    want = O.dust_pdf(model, g, t)
    eng = _engine(model, t)
    got = eng.setDustPDF(1, fetch=True)
    assert _same(got, want)
    assert (not multi) or np.isnan(want).any()        # rows of cells whose grains all sublimed are 0/0, as in the reference


def test_dust_pdf_feeds_transport():
    """Transport sampling the device-built table == oracle transport sampling the oracle table."""
    model, t = W.dust_closure(n=9, nbins=120, nPhotons=30000)
    g = model.grids[0]
    g.dustPDF = O.dust_pdf(model, g, t)
    orc = O.Oracle(model)
    orc.transport(1, 0, 30000, seed=77)
    eng = _engine(model, t, seed=77)
    eng.setDustPDF(1)
    eng.zero_estimators()
    eng.energyPacketDriver(1, 30000)
    eng.reduce()
    got, want = eng.fetch(1), orc.folded(1, float(model.deltaE[1]))
    assert _same(got["Jste"][1:], want["Jste"][1:])
    assert _same(got["escapedPackets"], want["escapedPackets"])
    assert got["Jste"].sum() > 0


@pytest.mark.parametrize("multi,sym", [(True, True), (False, True)])
def test_dust_update_matches_oracle(multi, sym):
    model, t = W.dust_closure(n=9, nbins=120, nPhotons=40000, multiChem=multi)
    g = model.grids[0]
    eng = _engine(model, t, seed=9)
    eng.setDustPDF(1)
    eng.zero_estimators()
    eng.energyPacketDriver(1, 40000)
    eng.reduce()
    J = eng.fetch(1)["Jste"]
    Js, _ = scale_estimators(model, J, np.zeros((1, 1, 1), F32))
    wantT, wantC = O.dust_update(model, g, t, Js, 0.05)
    T, conv, nconv = eng.getDustT(1, 0.05)
    assert _same(T, wantT)
    assert np.array_equal(conv, wantC)
    assert nconv == int(wantC.sum())
    assert 0 < (T[0, 0, 1:] > 1.0).sum()


def test_dust_update_table_ends_and_unlit_cells():
    model, t = W.dust_closure(n=7, nbins=90)
    g = model.grids[0]
    eng = _engine(model, t)
    T0 = g.Tdust.copy()
    eng.zero_estimators()                       # J = 0 everywhere: no cell was crossed by a packet,
    zero = np.zeros((g.nCells + 1, model.nbins), F32, order="F")       # updateCell leaves them all alone
    wantT, wantC = O.dust_update(model, g, t, zero, 0.05)
    T, conv, nconv = eng.getDustT(1, 0.05)
    assert _same(T, wantT) and np.array_equal(conv, wantC) and nconv == 0
    assert _same(T, T0)
    # a very faint field: crossed cells fall below the table (1 K), the others keep their T
    eng.setDustPDF(1)
    eng.zero_estimators()
    eng.energyPacketDriver(1, 40, deltaE=1.0e-25)
    eng.reduce()
    Js, _ = scale_estimators(model, eng.fetch(1)["Jste"], np.zeros((1, 1, 1), F32))
    lit = (Js > 0).any(axis=1)
    assert lit[1:].any() and not lit[1:].all()
    wantT, wantC = O.dust_update(model, g, t, Js, 0.05)
    T, conv, nconv = eng.getDustT(1, 0.05)
    assert _same(T, wantT) and np.array_equal(conv, wantC)
    assert np.all(T[1, 1:, lit] == 1.0) and _same(T[:, :, ~lit], T0[:, :, ~lit])


def test_device_lucy_loop_matches_oracle_loop():
    """Three dust-only Lucy iterations entirely on the device (no estimator or PDF leaves
    the GPU between passes) == the same loop on the oracle, bit for bit, including the
    sublimation flags the new temperatures imply for scattering."""
    n = 30000
    model, t = W.dust_closure(n=9, nbins=120, nPhotons=n, T0=100.0)
    g = model.grids[0]
    eng = _engine(model, t, seed=4242)
    ref_model, _ = W.dust_closure(n=9, nbins=120, nPhotons=n, T0=100.0)
    rg = ref_model.grids[0]
    for it in range(3):
        eng.setDustPDF(1)
        eng.zero_estimators()
        eng.energyPacketDriver(1, n)
        eng.reduce()
        T, conv, nconv = eng.getDustT(1, 0.05)

        rg.dustPDF = O.dust_pdf(ref_model, rg, t)
        orc = O.Oracle(ref_model)
        orc.transport(1, 0, n, seed=4242)
        Jr = orc.folded(1, float(ref_model.deltaE[1]))["Jste"]
        assert _same(eng.fetch(1)["Jste"][1:], Jr[1:]), f"iteration {it}"
        Js, _ = scale_estimators(ref_model, Jr, np.zeros((1, 1, 1), F32))
        # iterateMC zeroes grid%lgConverged before every iteration (iteration_mod.f90:87)
        wantT, wantC = O.dust_update(ref_model, rg, t, Js, 0.05)
        rg.Tdust = wantT
        assert _same(T, wantT), f"iteration {it}"
        assert np.array_equal(conv, wantC)


def test_dust_closure_error_behaviour():
    model, t = W.dust_closure(n=7, nbins=90)
    eng = PacketEngine(model)
    eng.set_opacity()
    eng.set_dust_state()
    with pytest.raises(MocassinError):          # tables not set
        eng.setDustPDF(1)
    eng.set_xsec(t["xSecArray"])
    bad = t["dustEmIntegral"].copy(order="F")
    bad[0, 0, 10] = bad[0, 0, 9] * F32(0.5)     # locate needs an ascending table
    with pytest.raises(MocassinError):
        eng.set_dust_tables(t["widFlx"], t["grainWeight"], t["dustAbsXsecP"], bad)
    eng.set_dust_tables(t["widFlx"], t["grainWeight"], t["dustAbsXsecP"], t["dustEmIntegral"])
    with pytest.raises(MocassinError):          # set_dust_state must follow set_dust_tables
        eng.getDustT(1, 0.05)
    gas = W.hii_region(n=5, nbins=60, nPhotons=100)
    e2 = PacketEngine(gas)
    with pytest.raises(MocassinError):
        e2.setDustPDF(1)
