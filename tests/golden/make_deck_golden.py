"""Fixtures of the shipped dust decks (benchmarks/dust/1D/p0tau{1,10,100}, 2D/tau1.000): what
mocassin_b200.deck.load_dust_deck builds from the reference's own input.in, density, grain and
optical-constant files, stored as plain arrays (tests/golden/deck_<name>.npz) because those
files do not exist on the GPU box.  Run here, where /root/reference is mounted:

    python tests/golden/make_deck_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mocassin_b200 import deck  # noqa: E402

REF = os.environ.get("MOCASSIN_REFERENCE", "/root/reference")

if __name__ == "__main__":
    for name in ("p0tau1", "p0tau10", "p0tau100", "tau1.000"):
        sub = "2D" if name.startswith("tau") else "1D"
        m, t, d = deck.load_dust_deck(os.path.join(REF, "benchmarks", "dust", sub, name), REF)
        out = os.path.join(HERE, f"deck_{name}.npz")
        np.savez_compressed(out, **deck.deck_to_arrays(m, t, d))
        print(out, os.path.getsize(out), "bytes", "nCells", m.grids[0].nCells, "nbins", m.nbins)
