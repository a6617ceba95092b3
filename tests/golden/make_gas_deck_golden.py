"""Fixtures of the shipped gas benchmarks (benchmarks/gas/HII40, PN150) in their first-iteration
state: what mocassin_b200.gasdeck.load_gas_deck builds from the reference's own input.in,
abundance file, data/ph1.dat and data/ph2.dat (tests/golden/deck_<name>.npz), and beside it what the
REFERENCE'S OWN code -- run through oracle/f90ref -- makes of the same inputs
(tests/golden/ref_aux_gas_<name>.npz): the frequency mesh and thresholds of initCartesianGrid, the
cross-section stack and pointer tables of setPointers / initXSecArray, and the opacity of every cell
from ionizationDriver / addOpacity on those real tables (98 / 148 bands).  Run here, where
/root/reference is mounted:

    python tests/golden/make_gas_deck_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mocassin_b200 import gasdeck  # noqa: E402

REF = os.environ.get("MOCASSIN_REFERENCE", "/root/reference")
CONTBOLTZ1, GAUNTFF1 = 0.731, 1.17          # what the harness's BoltGaunt stand-in returns for bin 1


def reference_side(name):
    from oracle import oracle as O
    from oracle.f90ref.harness_aux import AuxReference

    O.build()
    m, t, d = gasdeck.load_gas_deck(os.path.join(REF, "benchmarks", "gas", name), REF)
    A = AuxReference(O.load(), math="libm")
    nu, wid, edges = A.gas_nu_mesh(t["ph1"], t["ph2"], t["lgElementOn"], t["nstages"], d.nbins, d.nuMin, d.nuMax)
    r = A.gas_xsec(nu, t["ph1"], t["ph2"], t["lgElementOn"], t["nstages"])
    g = m.grids[0]
    B = AuxReference(O.load())
    op, ff1 = B.gas_opacity(t["xsec"], m.nbins, t["ionDen"], t["elemAbun"], t["abIndex"], g.Hden, g.active, CONTBOLTZ1,
                            GAUNTFF1, g.Ne, g.Te, int(r["bremsXSecP"]))
    out = dict(nuArray=nu, widFlx=wid, ionEdge=edges[:len(t["ionEdge"])], opacity=op, ff1=ff1)
    out.update({k: np.asarray(v) for k, v in r.items()})
    out["xSecArray"] = out["xSecArray"][:int(r["xSecTop"])]
    return m, t, d, out


if __name__ == "__main__":
    for name in ("HII40", "PN150"):
        m, t, d, ref = reference_side(name)
        out = os.path.join(HERE, f"deck_{name}.npz")
        np.savez_compressed(out, **gasdeck.gas_deck_to_arrays(m, t, d))
        out2 = os.path.join(HERE, f"ref_aux_gas_{name}.npz")
        np.savez_compressed(out2, **ref)
        print(out, os.path.getsize(out), out2, os.path.getsize(out2), "nCells", m.grids[0].nCells, "nbins", m.nbins,
              "bands", t["bands"]["species"].shape[0])
