"""The shipped multi-grid examples (mocassin_b200/multideck.py): examples/multigridgas loads as it
is; examples/multigridgasdust stops in the reference itself (its sub-grid file lacks the Ndust
column a gas+dust run reads) and loads with that column read as 0.  The sub-grid reader is pinned
against the reference's own setSubGrids reading code, executed on the shipped files through the
translator's list-directed READ (tests/golden/ref_aux_subgrid_<deck>.npz; live where the reference is
mounted)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("MOCASSIN_REFERENCE", "/root/reference")
DECKS = ["multigridgas", "multigridgasdust"]
have_ref = os.path.isdir(os.path.join(REF, "examples", "multigridgas"))


def _fixture(name):
    from mocassin_b200 import multideck

    return multideck.multideck_from_arrays(dict(np.load(os.path.join(GOLD, f"deck_{name}.npz"))))


@pytest.mark.parametrize("name", DECKS)
def test_fixture_geometry(name):
    m, t, s = _fixture(name)
    mother, sub = m.grids
    assert (mother.nx, mother.ny, mother.nz) == (16, 16, 16) and (sub.nx, sub.ny, sub.nz) == (11, 11, 11)
    assert sub.motherP == 1 and m.nbins == 600 and m.lgSymmetricXYZ and m.lgGas
    assert m.lgDust == (name == "multigridgasdust")
    # the sub-grid spans [0, 2e16]^3: exactly one mother cell (the one at the origin) lies inside it
    assert int((mother.active < 0).sum()) == 1 and mother.active[0, 0, 0] == -2
    assert np.allclose(sub.xAxis, np.linspace(0, 2e16, 11), rtol=1e-6)
    # the star sits in the sub-grid (setStarPosition finds the origin there)
    assert list(m.starIndeces[0]) == [1, 1, 1, 2]
    assert t["bands"]["species"].shape[0] > 60
    if m.lgDust:
        assert t["dust"]["dustScaXsecP"].shape == (1, 10) and mother.Ndust.max() > 0 and sub.Ndust.max() == 0
        assert t["dust"]["dustScaXsecP"].min() > t["xsec"].HeIIXSecP1      # dust blocks sit behind the gas stack


def test_subgrid_reader_equals_reference_setsubgrids_golden():
    """multigridgas: axes (rescaled row by row from normalised coordinates), active cells and densities
    == the reference's own reading code on the shipped subgrid.in / subgrid0.dat."""
    m, _, _ = _fixture("multigridgas")
    ref = dict(np.load(os.path.join(GOLD, "ref_aux_subgrid_multigridgas.npz")))
    assert bytes(ref["outcome"]).decode() == "ok"
    sub = m.grids[1]
    for k, v in (("xAxis", sub.xAxis), ("yAxis", sub.yAxis), ("zAxis", sub.zAxis)):
        assert np.array_equal(ref[k].view(np.uint32), v.view(np.uint32)), k
    assert np.array_equal(ref["active"], sub.active) and int(ref["nCells"]) == sub.nCells and int(ref["motherP"]) == 1
    sel = sub.active > 0
    assert np.array_equal(ref["Hden3"][sel], sub.Hden[sub.active[sel]])                   # denfac = 1


def test_reference_itself_stops_on_multigridgasdust_as_shipped():
    ref = dict(np.load(os.path.join(GOLD, "ref_aux_subgrid_multigridgasdust.npz")))
    assert bytes(ref["outcome"]).decode().startswith("STOP")


@pytest.mark.parametrize("name", DECKS)
def test_oracle_transport_on_the_deck(name, oracle_lib):
    from oracle import oracle as O
    from oracle.oracle import Oracle

    m, t, _ = _fixture(name)
    for g, e in zip(m.grids, t["grids"]):
        g.opacity, g.scaOpac, _ = O.opacity(t["xsec"], m.nbins, e["ionDen"], t["elemAbun"], e["abIndex"], g.Hden,
                                            dust=e["dust"], model=m if m.lgDust else None)
    o = Oracle(m)
    n = 4000
    c, _ = o.transport(1, 0, n, seed=12345)
    assert c["nEscaped"] + c["nLinePackets"] + c["nDropped"] == n
    assert all(np.count_nonzero(o.folded(iG, float(m.deltaE[1]))["Jste"][1:]) > 0 for iG in (1, 2))   # both grids crossed


@pytest.mark.skipif(not have_ref, reason="needs the reference tree")
def test_live_loader_and_reference(oracle_lib):
    from mocassin_b200 import multideck as M
    from oracle.f90ref import rt
    from oracle.f90ref.harness_aux import AuxReference

    for name in DECKS:
        run = os.path.join(REF, "examples", name)
        dust = name.endswith("dust")
        if dust:
            with pytest.raises(M.DeckError, match="insanity occurred in setting yAxis"):
                M.load_multigrid_deck(run, REF)
        m, t, d = M.load_multigrid_deck(run, REF, pad_missing_ndust=dust)
        kept = dict(np.load(os.path.join(GOLD, f"deck_{name}.npz")))
        for k, v in M.multideck_to_arrays(m, t, d).items():
            assert np.array_equal(np.asarray(v), kept[k]), (name, k)
        A = AuxReference(oracle_lib, math="libm")
        xA, yA, zA, _ = M.read_density_file(os.path.join(run, "bipolar_lobes.dat"), 16, 16, 16)
        args = (open(os.path.join(run, "subgrid.in")).read(), open(os.path.join(run, "subgrid0.dat")).read(), (xA, yA, zA),
                (11, 11, 11), True, dust, 1.0e15, 1.0e18)
        if dust:
            with pytest.raises(rt.FortranStop):
                A.sub_grid_read(*args)
        else:
            r = A.sub_grid_read(*args)
            sub = m.grids[1]
            assert np.array_equal(r["xAxis"], sub.xAxis) and np.array_equal(r["active"], sub.active)
