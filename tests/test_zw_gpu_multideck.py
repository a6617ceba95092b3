"""The shipped multi-grid examples on the device (fixtures tests/golden/deck_multigridgas.npz,
deck_multigridgasdust.npz): K1 per grid on the real band list (+ 10 grain sizes for the dust deck)
against the oracle's addOpacity restatement, then the multi-grid transport -- star inside the
sub-grid, packets leaving it into the mother grid -- against the oracle, bit for bit, persistent
and wave-front schedules."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden")


def _fixture(name):
    from mocassin_b200 import multideck

    return multideck.multideck_from_arrays(dict(np.load(os.path.join(GOLD, f"deck_{name}.npz"))))


@pytest.mark.parametrize("name", ["multigridgas", "multigridgasdust"])
def test_device_opacity_and_transport_on_the_shipped_multigrid_deck(cuda_lib, name):
    from mocassin_b200.api import PacketEngine
    from oracle import oracle as O
    from oracle.oracle import Oracle

    m, t, _ = _fixture(name)
    n = 30000
    e = PacketEngine(m, seed=12345)
    e.set_xsec(t["xsec"].xSecArray)
    for iG, (g, ent) in enumerate(zip(m.grids, t["grids"]), start=1):
        e.assemble_opacity(iG, t["bands"], ent["den"], None, ent["dust"])                   # K1 on the device ...
        g.opacity, g.scaOpac, _ = O.opacity(t["xsec"], m.nbins, ent["ionDen"], t["elemAbun"], ent["abIndex"], g.Hden,
                                            dust=ent["dust"], model=m if m.lgDust else None)   # ... and on the oracle
        op, sca, _ = e.get_opacity(iG)
        assert np.array_equal(op[1:].view(np.uint32), g.opacity[1:].view(np.uint32)), iG
        if m.lgDust:
            assert np.array_equal(sca[1:].view(np.uint32), g.scaOpac[1:].view(np.uint32)), iG
    e.set_pdfs()
    e.set_dust_state()
    o = Oracle(m)
    co, fo = o.transport(1, 0, n, seed=12345, want_fates=True)
    want = [o.folded(iG, float(m.deltaE[1])) for iG in (1, 2)]
    for wavefront in (0, 1):
        e.set_option("trace", 1)
        e.set_option("wavefront", wavefront)
        e.zero_estimators()
        cg = e.energyPacketDriver(1, n)
        assert np.array_equal(e.fates(n), fo), wavefront
        for k in ("nAbs", "nSca", "nSegments", "nEscaped", "nLinePackets", "nDropped"):
            assert cg[k] == co[k], (wavefront, k)
        for iG in (1, 2):
            got = e.fetch(iG)
            assert np.array_equal(got["Jste"][1:], want[iG - 1]["Jste"][1:]), (wavefront, iG)
            assert np.array_equal(got["escapedPackets"], want[iG - 1]["escapedPackets"]), (wavefront, iG)
    assert cg["nSegments"] > 5 * n
    e.close()
