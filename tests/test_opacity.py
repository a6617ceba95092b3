"""Opacity assembly (K1): host band flattening + device kernel vs the oracle's restatement
of addOpacity (ionization_mod.f90:349-484) and the dust loop (iteration_mod.f90:166-227)."""
import types

import numpy as np
import pytest

import opacity_case
from mocassin_b200 import workloads as W
from oracle import oracle as orc


def _dust_model(case):
    return types.SimpleNamespace(**case["dust_model"])


def test_band_list_reproduces_oracle_on_cpu():
    """numpy evaluation of the flattened band list == oracle_opacity (gas part): checks the
    host flattening (species columns, xSecP-nuLowP offsets, clipping rules) on the CPU."""
    c = opacity_case.make(nCells=400)
    t, nb = c["t"], c["nbins"]
    op, _, _ = orc.opacity(t, nb, c["ionDen"], c["elemAbun"], c["abIndex"], c["Hden"], ff1=c["ff1"])
    bands = t.band_list(nb)
    den = t.species_densities(c["ionDen"], c["elemAbun"], c["abIndex"], c["Hden"])
    ref = np.zeros_like(op)
    ref[1:, 0] = c["ff1"][1:]
    xs = t.xSecArray
    for sp, off, lo, hi in zip(bands["species"], bands["off"], bands["low"], bands["high"]):
        up = max(lo, min(hi, nb))
        d = den[:, sp - 1]
        for nu in range(lo, up + 1):
            add = (xs[nu + off - 1] * d).astype(np.float32)
            ref[:, nu - 1] = np.where(d > 0, (ref[:, nu - 1] + add).astype(np.float32), ref[:, nu - 1])
    assert np.array_equal(op.view(np.uint32), ref.view(np.uint32))
    assert op[1:].any() and not op[0].any()


@pytest.mark.gpu
@pytest.mark.parametrize("multi_chem", [True, False])
def test_device_opacity_assembly_bit_exact(multi_chem):
    from mocassin_b200.api import PacketEngine

    c = opacity_case.make(multi_chem=multi_chem)
    t, nb = c["t"], c["nbins"]
    dm = _dust_model(c)
    # a model whose only job is to carry the grid shape and dust species tables
    m = W.dust_shell(n=8, nbins=nb)
    g = m.grids[0]
    # replace the grid by a flat list of nCells cells: 1 x 1 x n would do, keep it simple
    n = c["nCells"]
    side = int(round(n ** (1 / 3))) + 1
    ax = np.arange(side, dtype=np.float32) * np.float32(1e15)
    from mocassin_b200.model import Grid, number_active
    mask = np.zeros((side, side, side), bool)
    mask.reshape(-1)[:n] = True
    active, nc = number_active(mask)
    assert nc == n
    m.grids[0] = Grid(xAxis=ax, yAxis=ax.copy(), zAxis=ax.copy(), active=active, nCells=n)
    m.lgGas = True
    m.lgMultiDustChemistry = multi_chem
    m.nSpeciesMax, m.nSizes = dm.nSpeciesMax, dm.nSizes
    m.nSpeciesPart, m.dustComPoint = dm.nSpeciesPart, dm.dustComPoint
    m.grainAbun, m.TdustSublime = dm.grainAbun, dm.TdustSublime
    m.starIndeces[0, :3] = 1
    e = PacketEngine(m)
    e.set_xsec(t.xSecArray)
    den = t.species_densities(c["ionDen"], c["elemAbun"], c["abIndex"], c["Hden"])
    e.assemble_opacity(1, t.band_list(nb), den, c["ff1"], c["dust"])
    op, sca, ab = e.get_opacity(1, want_abs=True)
    dm.lgMultiDustChemistry = multi_chem
    rop, rsca, rab = orc.opacity(t, nb, c["ionDen"], c["elemAbun"], c["abIndex"], c["Hden"], ff1=c["ff1"],
                                 dust=c["dust"], model=dm)
    assert np.array_equal(sca.view(np.uint32), rsca.view(np.uint32))
    assert np.array_equal(ab.view(np.uint32), rab.view(np.uint32))
    assert np.array_equal(op.view(np.uint32), rop.view(np.uint32))
    assert sca[1:].any() and ab[1:].any()
    e.close()


@pytest.mark.gpu
def test_bench_workload_tables_device_vs_numpy():
    """The bench builds its 128^3 opacity tables with K1; the CPU arm builds them with numpy.
    Same operation order -> identical bits (checked here at 24^3)."""
    import argparse

    import bench
    from mocassin_b200.api import PacketEngine

    args = argparse.Namespace(grid=24, nbins=150, workload="clumpy")
    m = bench.build_model(args, tables=False)
    xsec, bands, den, dust = bench.compact_inputs(m)
    e = PacketEngine(m)
    e.set_xsec(xsec)
    e.assemble_opacity(1, bands, den, None, dust)
    op, sca, _ = e.get_opacity(1)
    m2 = bench.build_model(args, tables=False)
    bench.host_tables(m2)
    assert np.array_equal(op.view(np.uint32), m2.grids[0].opacity.view(np.uint32))
    assert np.array_equal(sca.view(np.uint32), m2.grids[0].scaOpac.view(np.uint32))
    e.close()
