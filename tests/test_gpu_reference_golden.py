"""GPU tests against the golden vectors of the reference's OWN code (tests/golden/ref_*.npz,
produced by running the f90py translation of photon_mod.f90, see tests/test_reference_pin.py).

The CUDA path keeps order-independent integer tallies instead of the reference's sequential
float32 sums, so the bar per quantity is:
* per packet: cell crossings and energyPacketRun calls          -- equal
* escapedPackets: packet count per (cell, nu, angle)             -- equal (count = sum/deltaE)
* linePackets (debug): packet count per (cell, line)             -- equal
* planeIonDistribution                                           -- equal
* Jste/Jdif: same non-zero pattern, and EVERY element within its own rigorous bound of the
  reference's sequential float32 sum: n_i * unit/2 * deltaE/dV (quantisation of the n_i path
  lengths added to it) + (n_i + 8) * 2^-24 * J (the float32 roundings of the running sum and
  of the fold) -- ref_cases.j_error_bound; the grid total within 1e-5
"""
import os

import numpy as np
import pytest

import ref_cases
from mocassin_b200.api import PacketEngine

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _counts(E, dE):
    c = np.rint(E.astype(np.float64) / dE)
    assert np.allclose(c * dE, E, rtol=1e-3, atol=0)
    return c.astype(np.int64)


@pytest.mark.parametrize("wavefront", [0, 1])
@pytest.mark.parametrize("name", list(ref_cases.REF_CASES))
def test_cuda_matches_reference_golden(name, wavefront):
    want = dict(np.load(os.path.join(GOLD, f"ref_{name}.npz")))
    m, n, mode = ref_cases.make(name)
    e = PacketEngine(m, seed=ref_cases.SEED)
    e.upload_iteration_inputs()
    e.set_option("trace", 1)
    e.set_option("wavefront", wavefront)
    e.zero_estimators()
    if mode == "stellar":
        iStar = 1
        cg = e.energyPacketDriver(1, n)
    elif mode[0] == "stars":                 # one call per source, different packet energies
        nseg, k0 = 0, 0
        for iStar in mode[1]:
            c = e.energyPacketDriver(iStar, n)
            nseg += c["nSegments"]
            assert np.array_equal(e.fates(n)[:, :2], want["fates"][k0:k0 + n]), f"star {iStar}"
            k0 += n
        assert nseg == int(want["nSegments"])
        got = e.fetch(1)
        for k in ("Jste", "escapedPackets"):
            g, w = got[k][1:].astype(np.float64), want[f"{k}_g1"][1:].astype(np.float64)
            assert np.array_equal(g > 0, w > 0), k
            sel = w > 0
            rel = np.abs(g[sel] - w[sel]) / w[sel]
            if k == "escapedPackets":        # count * deltaE per source, summed: float32 rounding only
                assert rel.max() < 2e-6, rel.max()
            else:
                lim = ref_cases.j_error_bound(m, mode, n, e.len_unit)(1, k, g, w)
                assert np.all(np.abs(g - w) <= lim), float((np.abs(g - w) / np.where(lim > 0, lim, 1)).max())
        e.close()
        return
    elif mode == "reslines":
        iStar = 1
        cg = e.resLinePacketsTransfer(1)
        n = 0
        assert cg["nPackets"] == want["fates"].shape[0]
    else:
        iStar = 0
        cg = e.energyPacketDriver(0, n, gpLoc=mode[1], cellLoc=list(mode[2]))
    assert cg["nSegments"] == int(want["nSegments"])
    if n:
        fg = e.fates(n)
        bad = np.flatnonzero((fg[:, :2] != want["fates"]).any(axis=1))
        assert bad.size == 0, f"{bad.size} packets differ from the reference, first {bad[:5]}"
    if m.lgPlaneIonization:
        assert np.array_equal(e.plane_distribution(), want["plane"])
    dE = float(m.deltaE[iStar])
    npk = want["fates"].shape[0]
    names = ["Jste", "escapedPackets"] + (["Jdif", "linePackets"] if m.lgDebug else [])
    bound = ref_cases.j_error_bound(m, mode, n, e.len_unit)
    for iG in range(1, m.nGrids + 1):
        got = e.fetch(iG, want=names)
        for k in names:
            g, w = got[k], want[f"{k}_g{iG}"]
            assert g.shape == w.shape, (k, g.shape, w.shape)
            if k in ("escapedPackets", "linePackets"):
                assert np.array_equal(_counts(g, dE), _counts(w, dE)), (iG, k)
                continue
            g, w = g[1:].astype(np.float64), w[1:].astype(np.float64)     # row 0: inactive-cell sink, never read
            assert np.array_equal(g > 0, w > 0), (iG, k)
            sel = w > 0
            if not sel.any():
                continue
            # every element within its own bound: quantisation of its n_i path lengths + float32 roundings
            lim = bound(iG, k, g, w)
            assert np.all(np.abs(g - w) <= lim), (iG, k, float((np.abs(g - w) / np.where(lim > 0, lim, 1)).max()))
            assert abs(g.sum() - w.sum()) / w.sum() < 1e-5, (iG, k)
    e.close()


def _same_nan(a, b):
    a, b = np.asarray(a), np.asarray(b)
    both = np.isnan(a) & np.isnan(b)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | both))


def test_device_dust_pdf_matches_reference_setdustpdf():
    """K6 (dust_pdf_kernel) against what the reference's own emissionDriver -> setDustPDF produced
    (tests/golden/ref_aux_dust_closure.npz)."""
    want = dict(np.load(os.path.join(GOLD, "ref_aux_dust_closure.npz")))
    model, g, t, J, Jd = ref_cases._dust_inputs(False)
    eng = PacketEngine(model)
    eng.set_xsec(t["xSecArray"])
    eng.set_dust_tables(t["widFlx"], t["grainWeight"], t["dustAbsXsecP"], t["dustEmIntegral"])
    eng.set_opacity()
    eng.set_dust_state()
    got = eng.setDustPDF(1, fetch=True)
    assert _same_nan(got[1:], want["dustPDF"][1:])
    eng.close()


@pytest.mark.parametrize("name", ["opacity_multichem", "opacity_singlechem"])
def test_device_opacity_matches_reference_opacity_block(name):
    """K1 (opacity_kernel + the host band flattening) against the output of the reference's own
    opacity block of iterateMC (ionizationDriver/addOpacity + dust loop; ref_aux_opacity_*.npz)."""
    from mocassin_b200 import workloads as W
    from mocassin_b200.model import Grid

    want = dict(np.load(os.path.join(GOLD, f"ref_aux_{name}.npz")))
    c = ref_cases._opacity_inputs(name == "opacity_multichem")
    t, nb, dm = c["t"], c["nbins"], c["dm"]
    m = W.dust_shell(n=8, nbins=nb)
    ax = np.arange(4, dtype=np.float32) * np.float32(1e15)
    m.grids[0] = Grid(xAxis=ax, yAxis=ax.copy(), zAxis=ax.copy(), active=c["active"], nCells=c["nCells"])
    m.lgGas = True
    m.lgMultiDustChemistry = bool(dm.lgMultiDustChemistry)
    m.nSpeciesMax, m.nSizes = dm.nSpeciesMax, dm.nSizes
    m.nSpeciesPart, m.dustComPoint = dm.nSpeciesPart, dm.dustComPoint
    m.grainAbun, m.TdustSublime = dm.grainAbun, dm.TdustSublime
    m.starIndeces[0, :3] = 1
    e = PacketEngine(m)
    e.set_xsec(t.xSecArray)
    den = t.species_densities(c["ionDen"], c["elemAbun"], c["abIndex"], c["Hden"])
    e.assemble_opacity(1, t.band_list(nb), den, want["ff1"], c["dust"])
    op, sca, ab = e.get_opacity(1, want_abs=True)
    for got, key in ((sca, "scaOpac"), (ab, "absOpac"), (op, "opacity")):
        assert np.array_equal(got[1:].view(np.uint32), want[key][1:].view(np.uint32)), key
    e.close()
