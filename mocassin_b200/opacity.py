"""Host side of the opacity assembly (K1): turns the reference's atomic-data pointer
tables into the flat band list the device kernel consumes.

Mirrors ``ionizationDriver`` / ``addOpacity`` of the reference
(``source/ionization_mod.f90:26-129,349-484``): the per-cell species densities
``density(elem,ion) = ionDen*elemAbun*Hden`` (:65-80) become the columns of ``den`` and every
``inOpacity`` call -- H0, He0, He+ (:396-407) and each (element, ion, shell) of
``putOpacity`` (:418-443) -- becomes one band ``{species column, xSecP-nuLowP, nuLowP,
nuHighP}`` in the reference's call order, which is the order the kernel sums in.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .model import F32, I32

NELEMENTS = 30      # constants_mod.f90:47


@dataclass
class XSecTables:
    """The xSec_mod / elements_mod globals read by addOpacity (ph_mod.f90:204-402,
    hydro_mod.f90:208-475)."""

    xSecArray: np.ndarray        # (nXsec,) float32
    nstages: int
    lgElementOn: np.ndarray      # (30,) int32 0/1
    elementXref: np.ndarray      # (30,) int32: column of ionDen for element n (1-based)
    HlevXSecP1: int
    HlevNuP1: int
    HeISingXSecP1: int
    HeIlevNuP1: int
    HeIIXSecP1: int
    HeIIlevNuP1: int
    elementP: np.ndarray         # (30,30,7,3) int32 F-order: [nuLowP, nuHighP, xSecP]
    nShells: np.ndarray          # (30,30) int32 F-order

    def species(self):
        """(elem, ion) pairs that own a column of `den`, in addOpacity's order."""
        sp = [(1, 1), (2, 1), (2, 2)]
        for el in range(3, NELEMENTS + 1):
            if self.lgElementOn[el - 1]:
                for ion in range(1, min(el, self.nstages) + 1):
                    sp.append((el, ion))
        return sp

    def band_list(self, nbins: int):
        """Flatten addOpacity's inOpacity calls into bands (1-based Fortran values)."""
        sp = self.species()
        col = {s: i + 1 for i, s in enumerate(sp)}
        spec, off, lo, hi = [], [], [], []

        def add(s, xSecP, nuLowP, nuHighP):
            spec.append(col[s]); off.append(xSecP - nuLowP); lo.append(nuLowP); hi.append(nuHighP)

        add((1, 1), self.HlevXSecP1, self.HlevNuP1, nbins)
        add((2, 1), self.HeISingXSecP1, self.HeIlevNuP1, nbins)
        add((2, 2), self.HeIIXSecP1, self.HeIIlevNuP1, nbins)
        for el in range(3, NELEMENTS + 1):
            if not self.lgElementOn[el - 1]:
                continue
            for ion in range(1, min(el, self.nstages) + 1):
                for sh in range(1, int(self.nShells[el - 1, ion - 1]) + 1):
                    nuLowP, nuHighP, xSecP = (int(v) for v in self.elementP[el - 1, ion - 1, sh - 1, :])
                    add((el, ion), xSecP, nuLowP, nuHighP)
        return dict(species=np.array(spec, I32), off=np.array(off, I32), low=np.array(lo, I32), high=np.array(hi, I32))

    def species_densities(self, ionDen: np.ndarray, elemAbun: np.ndarray, abIndex: np.ndarray, Hden: np.ndarray):
        """density(n,i) = ionDen(cell,xref(n),i)*elemAbun(abIndex(cell),n)*Hden(cell)
        (ionization_mod.f90:76-78) for every cell, float32, left to right."""
        nR = Hden.shape[0]
        sp = self.species()
        den = np.zeros((nR, len(sp)), dtype=F32, order="F")
        ab_rows = np.maximum(abIndex, 1) - 1
        for c, (el, ion) in enumerate(sp):
            if not self.lgElementOn[el - 1]:
                continue
            xr = int(self.elementXref[el - 1]) - 1
            d = (ionDen[:, xr, ion - 1].astype(F32) * elemAbun[ab_rows, el - 1].astype(F32)).astype(F32)
            den[:, c] = (d * Hden.astype(F32)).astype(F32)
        den[0, :] = 0.0
        return den
