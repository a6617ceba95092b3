"""A seeded synthetic atomic-data set in the reference's own table shapes (xSecArray stack,
elementP/nShells pointer tables, multi-component dust) for the opacity-assembly tests."""
import numpy as np

from mocassin_b200.model import F32, I32
from mocassin_b200.opacity import XSecTables


def make(nCells=3000, nbins=200, seed=3, multi_chem=True):
    rng = np.random.default_rng(seed)
    nstages = 5
    on = np.zeros(30, I32)
    for el in (1, 2, 6, 7, 8, 10, 16, 26):
        on[el - 1] = 1
    xref = np.zeros(30, I32)
    xref[on > 0] = np.arange(1, int(on.sum()) + 1)
    nUsed = int(on.sum())
    xs = []

    def push(n):
        start = len(xs) + 1
        xs.extend((rng.random(n) * 1e-18).astype(F32).tolist())
        return start

    tabs = dict(HlevNuP1=40, HeIlevNuP1=70, HeIIlevNuP1=120)
    tabs["HlevXSecP1"] = push(nbins - 40 + 1)
    tabs["HeISingXSecP1"] = push(nbins - 70 + 1)
    tabs["HeIIXSecP1"] = push(nbins - 120 + 1)
    elementP = np.zeros((30, 30, 7, 3), I32, order="F")
    nShells = np.zeros((30, 30), I32, order="F")
    for el in range(3, 31):
        if not on[el - 1]:
            continue
        for ion in range(1, min(el, nstages) + 1):
            ns = int(rng.integers(1, 4))
            nShells[el - 1, ion - 1] = ns
            for sh in range(1, ns + 1):
                lo = int(rng.integers(2, nbins - 5))
                kind = rng.integers(0, 4)
                if kind == 0:
                    hi = nbins + int(rng.integers(0, 50))      # extends past the mesh: clipped to nbins
                elif kind == 1:
                    hi = lo - 1                                  # "(2,1)"-style flag: one bin only
                else:
                    hi = int(rng.integers(lo, nbins + 1))
                n = max(hi, lo) - lo + 1
                elementP[el - 1, ion - 1, sh - 1, :] = (lo, hi, push(min(n, nbins - lo + 1) + 2))
    # dust cross-sections: 3 species in total, 2 sizes, two chemistry components
    nSpeciesTot, nSizes = 3, 2
    scaP = np.zeros((nSpeciesTot, nSizes), I32, order="F")
    absP = np.zeros((nSpeciesTot, nSizes), I32, order="F")
    for s in range(nSpeciesTot):
        for a in range(nSizes):
            scaP[s, a] = push(nbins)
            absP[s, a] = push(nbins)
    t = XSecTables(xSecArray=np.array(xs, F32), nstages=nstages, lgElementOn=on, elementXref=xref,
                   elementP=elementP, nShells=nShells, **tabs)
    nR = nCells + 1
    ionDen = np.asfortranarray(rng.random((nR, nUsed, nstages)).astype(F32))
    ionDen[rng.random(ionDen.shape) < 0.15] = 0.0                # some stages empty (density>0 test)
    elemAbun = np.asfortranarray((rng.random((2, 30)) * 1e-3).astype(F32))
    elemAbun[:, 0] = 1.0
    abIndex = rng.integers(1, 3, nR).astype(I32)
    Hden = (rng.random(nR) * 1e3).astype(F32)
    ff1 = (rng.random(nR) * 1e-20).astype(F32)
    # dust state
    dust_model = dict(nSpeciesMax=2, nSizes=nSizes, nSpeciesPart=np.array([2, 1], I32),
                      dustComPoint=np.array([1, 3], I32),
                      grainAbun=np.asfortranarray(np.array([[0.7, 0.3], [1.0, 0.0]], F32)),
                      TdustSublime=np.array([1400.0, 1200.0, 900.0], F32), lgMultiDustChemistry=multi_chem)
    Tdust = np.asfortranarray((rng.random((3, nSizes + 1, nR)) * 2000.0).astype(F32))
    dust = dict(Ndust=(rng.random(nR) * 1e-9).astype(F32), Tdust=Tdust,
                dustAbunIndex=rng.integers(1, 3, nR).astype(I32) if multi_chem else None,
                grainWeight=np.array([0.4, 0.6], F32), dustScaXsecP=scaP, dustAbsXsecP=absP)
    return dict(t=t, nbins=nbins, ionDen=ionDen, elemAbun=elemAbun, abIndex=abIndex, Hden=Hden, ff1=ff1,
                dust=dust, dust_model=dust_model, nCells=nCells)
