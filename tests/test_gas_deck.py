"""The shipped gas benchmarks benchmarks/gas/HII40 and PN150 in their first-iteration state
(mocassin_b200/gasdata.py, gasdeck.py): frequency mesh, cross-section stack, pointer tables, band
list, initial ion state -- pinned bit for bit against the reference's own initCartesianGrid /
setPointers / setShells / initXSecArray / phFitEl / makeOpacity / setMotherGrid / ionizationDriver,
executed through the Fortran translator (live where /root/reference is mounted, and through
tests/golden/ref_aux_gas_<deck>.npz everywhere)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("MOCASSIN_REFERENCE", "/root/reference")
DECKS = ["HII40", "PN150"]
have_ref = os.path.isdir(os.path.join(REF, "benchmarks", "gas"))


def _fixture(name):
    from mocassin_b200 import gasdeck

    return gasdeck.gas_deck_from_arrays(dict(np.load(os.path.join(GOLD, f"deck_{name}.npz"))))


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", DECKS)
def test_fixture_shapes_as_the_survey_lists_them(name):
    m, t, s = _fixture(name)
    g = m.grids[0]
    assert (g.nx, g.ny, g.nz) == (13, 13, 13) and m.lgSymmetricXYZ and m.lgGas and not m.lgDust
    assert m.nbins == (600 if name == "HII40" else 700)
    assert np.all(np.diff(m.nuArray) > 0) and m.nuArray[-1] < 1e29          # no sortUp duplicates left over
    nb = t["bands"]["species"].shape[0]
    assert nb == (98 if name == "HII40" else 148)                            # the real band list, not 3 toy bands
    assert t["den"].shape == (g.nCells + 1, len(t["xsec"].species()))
    assert s["recPDF_kind"].startswith("stand-in")


@pytest.mark.parametrize("name", DECKS)
def test_restatement_equals_the_reference_golden(name):
    """gasdata's mesh, stack and pointers == what the reference's own routines produced."""
    m, t, _ = _fixture(name)
    ref = dict(np.load(os.path.join(GOLD, f"ref_aux_gas_{name}.npz")))
    xt = t["xsec"]
    assert np.array_equal(_bits(m.nuArray), _bits(ref["nuArray"]))
    assert np.array_equal(_bits(t["widFlx"]), _bits(ref["widFlx"]))
    assert np.array_equal(_bits(xt.xSecArray), _bits(ref["xSecArray"])) and xt.xSecArray.shape[0] == int(ref["xSecTop"])
    assert np.array_equal(xt.elementP, ref["elementP"]) and np.array_equal(xt.nShells, ref["nShells"])
    assert (xt.HlevNuP1, xt.HeIlevNuP1, xt.HeIIlevNuP1) == (ref["HlevNuP"][0], ref["HeIlevNuP"][0], ref["HeIIlevNuP"][0])
    assert (xt.HlevXSecP1, xt.HeISingXSecP1, xt.HeIIXSecP1) == (ref["HlevXSecP"][0], ref["HeISingXSecP"][0], ref["HeIIXSecP"][0])
    assert float(m.ionEdge1) == float(ref["ionEdge"][0])


@pytest.mark.parametrize("name", DECKS)
def test_oracle_opacity_on_the_real_tables_equals_ionization_driver(name, oracle_lib):
    """K1's CPU restatement on the deck's ~100-150 bands == the reference's ionizationDriver /
    addOpacity run cell by cell on the same tables (golden), bit for bit; so does the numpy twin."""
    from mocassin_b200 import gasdeck
    from oracle import oracle as O

    m, t, _ = _fixture(name)
    ref = dict(np.load(os.path.join(GOLD, f"ref_aux_gas_{name}.npz")))
    g = m.grids[0]
    op, _, _ = O.opacity(t["xsec"], m.nbins, t["ionDen"], t["elemAbun"], t["abIndex"], g.Hden, ff1=ref["ff1"])
    assert np.array_equal(_bits(op[1:]), _bits(ref["opacity"][1:]))
    twin = gasdeck.host_opacity(m, t)
    twin[:, 0] = (twin[:, 0] + 0).astype(np.float32)
    want = ref["opacity"].copy()
    assert np.array_equal(_bits(twin[1:, 1:]), _bits(want[1:, 1:]))          # bin 1 also holds the free-free term
    assert op[1:].max() > 0 and np.count_nonzero(op[1]) > m.nbins // 4


@pytest.mark.parametrize("name", DECKS)
def test_oracle_transport_on_the_deck_conserves_packets(name, oracle_lib):
    from mocassin_b200 import gasdeck
    from oracle.oracle import Oracle

    m, t, _ = _fixture(name)
    m.grids[0].opacity = gasdeck.host_opacity(m, t)
    o = Oracle(m)
    n = 4000
    c, _ = o.transport(1, 0, n, seed=12345)
    assert c["nEscaped"] + c["nLinePackets"] + c["nDropped"] == n
    assert 0 < c["nAbs"] < n                       # X(H0) = 1e-5: the first iteration sees a thin nebula
    J = o.folded(1, float(m.deltaE[1]))["Jste"]
    assert np.count_nonzero(J[1:]) > 0


@pytest.mark.skipif(not have_ref, reason="needs the reference tree")
@pytest.mark.parametrize("name", DECKS)
def test_fresh_load_equals_fixture_and_live_reference(name, oracle_lib):
    """Where the reference is mounted: the loader on the shipped files reproduces the committed
    fixture, and the translated reference routines reproduce the loader -- live."""
    from mocassin_b200 import gasdeck
    from oracle.f90ref.harness_aux import AuxReference

    m, t, d = gasdeck.load_gas_deck(os.path.join(REF, "benchmarks", "gas", name), REF)
    fresh = gasdeck.gas_deck_to_arrays(m, t, d)
    kept = dict(np.load(os.path.join(GOLD, f"deck_{name}.npz")))
    for k, v in fresh.items():
        assert np.array_equal(np.asarray(v), kept[k]), k
    A = AuxReference(oracle_lib, math="libm")
    nu, wid, edges = A.gas_nu_mesh(t["ph1"], t["ph2"], t["lgElementOn"], t["nstages"], d.nbins, d.nuMin, d.nuMax)
    assert np.array_equal(_bits(nu), _bits(m.nuArray)) and np.array_equal(_bits(wid), _bits(t["widFlx"]))
    assert np.array_equal(_bits(edges[:len(t["ionEdge"])]), _bits(t["ionEdge"]))
    r = A.gas_xsec(nu, t["ph1"], t["ph2"], t["lgElementOn"], t["nstages"])
    assert np.array_equal(_bits(r["xSecArray"][:r["xSecTop"]]), _bits(t["xsec"].xSecArray))
    assert np.array_equal(r["elementP"], t["ptr"]["elementP"]) and np.array_equal(r["nShells"], t["ptr"]["nShells"])
    for k in ("HlevNuP", "HeIlevNuP", "HeIIlevNuP"):
        assert np.array_equal(r[k], t["ptr"][k]), k
    for k in ("HlevXSecP", "HeISingXSecP", "HeIIXSecP"):
        assert np.array_equal(r[k], t["xp"][k]), k
    ion, _ = A.initial_ions(m.grids[0].active, t["lgElementOn"], t["elementXref"], t["nstages"])
    assert np.array_equal(_bits(ion), _bits(t["ionDen"]))
