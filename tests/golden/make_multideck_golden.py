"""Fixtures of the shipped multi-grid examples (examples/multigridgas as it is; examples/multigridgasdust
with the sub-grid's missing Ndust column read as 0 -- as shipped the reference itself stops on it, see
mocassin_b200/multideck.py), and beside them what the REFERENCE'S OWN setSubGrids reading code, run
through oracle/f90ref on the shipped list and sub-grid files, produced (ref_aux_subgrid_<deck>.npz).

    python tests/golden/make_multideck_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mocassin_b200 import multideck as M  # noqa: E402

REF = os.environ.get("MOCASSIN_REFERENCE", "/root/reference")


def reference_subgrid(deck, lgDust):
    from oracle import oracle as O
    from oracle.f90ref import rt
    from oracle.f90ref.harness_aux import AuxReference

    O.build()
    run = os.path.join(REF, "examples", deck)
    A = AuxReference(O.load(), math="libm")
    xA, yA, zA, _ = M.read_density_file(os.path.join(run, "bipolar_lobes.dat"), 16, 16, 16)
    try:
        r = A.sub_grid_read(open(os.path.join(run, "subgrid.in")).read(), open(os.path.join(run, "subgrid0.dat")).read(),
                            (xA, yA, zA), (11, 11, 11), True, lgDust, 1.0e15, 1.0e18)
        r = {k: np.asarray(v) for k, v in r.items()}
        r["outcome"] = np.frombuffer(b"ok", dtype=np.uint8)
        return r
    except rt.FortranStop as ex:
        return dict(outcome=np.frombuffer(("STOP: " + str(ex)).encode(), dtype=np.uint8))
    except rt.FortranEOF as ex:
        return dict(outcome=np.frombuffer(("EOF: " + str(ex)).encode(), dtype=np.uint8))


if __name__ == "__main__":
    for deck, dust in (("multigridgas", False), ("multigridgasdust", True)):
        m, t, d = M.load_multigrid_deck(os.path.join(REF, "examples", deck), REF, pad_missing_ndust=dust)
        out = os.path.join(HERE, f"deck_{deck}.npz")
        np.savez_compressed(out, **M.multideck_to_arrays(m, t, d))
        ref = reference_subgrid(deck, dust)
        out2 = os.path.join(HERE, f"ref_aux_subgrid_{deck}.npz")
        np.savez_compressed(out2, **ref)
        print(out, os.path.getsize(out), "grids", [(g.nx, g.ny, g.nz, g.nCells) for g in m.grids], "reference:",
              bytes(ref["outcome"]).decode()[:80])
