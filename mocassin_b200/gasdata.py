"""Gas-side input data of the shipped gas benchmarks (benchmarks/gas/HII40, PN150): everything the
opacity assembly (K1) and the transport read that the reference derives from its atomic data files
before the first Lucy iteration.  Host-side Python -- in a drop-in build the Fortran host keeps its
own -- restated in the reference's float32 operation order and pinned against the reference's own
code, translated and executed (tests/test_gas_deck.py).

=========================================  =======================================================
reference                                  here
=========================================  =======================================================
``setComposition`` composition_mod.f90:10   :func:`read_abundances`
``phInit`` ph_mod.f90:408-466               :func:`read_ph_tables` (data/ph1.dat, data/ph2.dat)
``getOuterShell`` hydro_mod.f90:89-182      :func:`outer_shell`
ν-mesh ``initCartesianGrid``                :func:`nu_mesh_gas` (+ ``sortUp`` interpolation_mod.f90:7-40)
grid_mod.f90:132-177,213-258,333-338
``makeHydro``/``setPointers``/``setShells`` :func:`set_pointers`
/``limitShell`` hydro_mod.f90:56-85,208-475
``phFitEl`` ph_mod.f90:477-552              :func:`ph_fit_el`
``phFitHIon`` ph_mod.f90:560-606            :func:`ph_fit_h_ion`
``initXSecArray`` (gas part), ``powLawXSec``, :func:`init_xsec_array`
``makeOpacity`` ph_mod.f90:204-402,712-803
initial ion state grid_mod.f90:1562-1607    :func:`initial_ion_state`
=========================================  =======================================================
"""
from __future__ import annotations

import numpy as np

from .model import F32, I32, locate
from .opacity import NELEMENTS, XSecTables

RYD_TO_EV = F32(13.6056981)          # constants_mod.f90:41
RADIO_4P9_GHZ = F32(1.489434e-6)     # :19
H_ION_POT = F32(0.99946)             # :15
K_SHELL_LIMIT = F32(7.35e4)          # :29
N_H_LEVEL, N_HEI_LEVEL, N_HEII_LEVEL = 10, 9, 9
N_SERIES = 17
SERIES_EDGE = np.array([0.0069, 0.0083, 0.01, 0.0123, 0.0156, 0.0204, 0.0278, 0.04, 0.0625, 0.11117,
                        0.11610, 0.12248, 0.13732, 0.24763, 0.24994, 0.26630, 0.29189], dtype=F32)   # grid_mod.f90:136-138

# phInit, ph_mod.f90:415-419
LEVEL = (0, 0, 1, 0, 1, 2, 0)
N_INN = (0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 3, 3, 3, 3, 3, 3, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5)
N_TOT = (1, 1, 2, 2, 3, 3, 3, 3, 3, 3, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 7, 7)


def _f(x):
    return F32(x)


def read_abundances(path: str):
    """setComposition for one abundance file: 30 reals (list-directed, one per line; anything after
    the number is a comment).  Returns (elemAbun(30), lgElementOn(30), elementXref(30), nElementsUsed):
    elements with abundance <= 1e-12 are switched off (composition_mod.f90:44-70)."""
    ab = np.zeros(NELEMENTS, dtype=F32)
    with open(path) as fh:
        vals = []
        for line in fh:
            t = line.replace(",", " ").split()
            if not t:
                continue
            vals.append(F32(float(t[0].lower().replace("d", "e"))))
            if len(vals) == NELEMENTS:
                break
    if len(vals) < NELEMENTS:
        raise ValueError(f"{path}: {len(vals)} abundances, 30 expected")
    ab[:] = vals
    on = (ab > F32(1.0e-12)).astype(I32)
    xref = np.zeros(NELEMENTS, dtype=I32)
    xref[on > 0] = np.arange(1, int(on.sum()) + 1, dtype=I32)
    return ab, on, xref, int(on.sum())


def _numbers(path: str):
    with open(path) as fh:
        for line in fh:
            for t in line.replace(",", " ").split():
                yield float(t.lower().replace("d", "e"))


def read_ph_tables(ph1_path: str, ph2_path: str):
    """phInit (ph_mod.f90:408-466): ph1(6, 30, 30, 7) and ph2(7, 30, 30), float32, in the
    reference's index order [i, element, electrons, shell] / [i, element, electrons]."""
    ph1 = np.zeros((6, 30, 30, 7), dtype=F32)
    ph2 = np.zeros((7, 30, 30), dtype=F32)
    it = _numbers(ph1_path)
    for j in range(1, 31):
        for k in range(1, j + 1):
            nt = N_TOT[k - 1]
            if j == k and k > 18:
                nt = 7
            if j == k + 1 and j in (20, 21, 22, 25, 26):
                nt = 7
            for l in range(1, nt + 1):
                for i in range(6):
                    ph1[i, j - 1, k - 1, l - 1] = F32(next(it))
    it = _numbers(ph2_path)
    for k in range(1, 31):
        for j in range(1, k + 1):
            if k not in (15, 17, 19) and (k <= 20 or k == 26):
                for i in range(7):
                    ph2[i, k - 1, j - 1] = F32(next(it))
    return ph1, ph2


def outer_shell(z: int, nElec: int) -> int:
    """getOuterShell (hydro_mod.f90:89-182), the shell number only (the statistical weights are not
    used on this path)."""
    out = N_TOT[nElec - 1]
    if z == nElec and z > 18:
        out = 7
    if z == nElec + 1 and z in (20, 21, 22, 25, 26):
        out = 7
    return out


def sort_up(a: np.ndarray) -> np.ndarray:
    """sortUp (interpolation_mod.f90:7-40): ascending selection of DISTINCT values; every duplicate
    leaves a 1e30 at the tail."""
    u = np.unique(np.asarray(a, dtype=F32))
    out = np.full(a.shape[0], F32(1.0e30), dtype=F32)
    out[:u.shape[0]] = u
    return out


def ion_edges(ph1: np.ndarray, lgElementOn: np.ndarray, nstages: int, nuMax: float):
    """The ionisation thresholds inside the frequency range (grid_mod.f90:140-177), sorted.  A
    threshold above nuMax leaves its slot to the next one (the counter only advances on a hit)."""
    edges = []
    last = None
    for elem in range(1, NELEMENTS + 1):
        for ion in range(1, min(elem, nstages - 1) + 1):
            if not lgElementOn[elem - 1]:
                break
            if elem > 2:
                nElec = elem - ion + 1
                e = F32(ph1[0, elem - 1, nElec - 1, outer_shell(elem, nElec) - 1] / RYD_TO_EV)
            elif elem == 1:
                e = F32(0.999434)
            elif ion == 1:
                e = F32(1.80804)
            else:
                e = F32(4.0)
            last = e
            if e <= F32(nuMax):
                edges.append(e)
    nEdges = len(edges)
    # the reference sorts ionEdge(1:nEdges): the slot nEdges+1 holds the last rejected value and is not read
    return sort_up(np.asarray(edges, dtype=F32)), nEdges


def nu_mesh_gas(nbins: int, nuMin: float, nuMax: float, ionEdge: np.ndarray, nEdges: int) -> np.ndarray:
    """The gas-only frequency mesh of initCartesianGrid (grid_mod.f90:213-258): the 4.9 GHz point,
    the 17 H-series edges and every ionisation threshold each with a point 0.0003 Ryd either side,
    a logarithmic mesh over the rest, then sortUp."""
    nu = np.zeros(nbins, dtype=F32)
    k = 0                                        # 0-based nuCount-1
    if F32(nuMin) < RADIO_4P9_GHZ:
        nu[0] = RADIO_4P9_GHZ
        k = 1
    nuMinA, nuMaxA = F32(nuMin), F32(nuMax)
    d3, d6 = F32(0.0003), F32(0.0006)
    for i in range(N_SERIES):
        nu[k], nu[k + 1], nu[k + 2] = SERIES_EDGE[i], F32(SERIES_EDGE[i] - d3), F32(SERIES_EDGE[i] + d3)
        if nu[k] < nuMinA:
            nuMinA = F32(nu[k] - d6)
        if nu[k] > nuMaxA:
            nuMaxA = nu[k]
        k += 3
    for i in range(nEdges):
        if ionEdge[i] < nuMaxA:
            nu[k], nu[k + 1], nu[k + 2] = ionEdge[i], F32(ionEdge[i] - d3), F32(ionEdge[i] + d3)
            if nu[k] < nuMinA:
                nuMinA = F32(nu[k] - d6)
            k += 3
    iCount = nbins - (k + 1) + 1
    step = F32(F32(np.log10(nuMaxA) - np.log10(nuMinA)) / F32(iCount - 1))
    nu[k] = nuMinA
    for i in range(k + 1, nbins):
        nu[i] = F32(np.power(F32(10.0), F32(np.log10(nu[i - 1]) + step)))
    return sort_up(nu)


def _loc(nu, x):
    return int(locate(nu, F32(x)))


def set_pointers(nu: np.ndarray, ph1: np.ndarray, lgElementOn: np.ndarray):
    """makeHydro + setPointers + setShells + limitShell (hydro_mod.f90:56-85,208-475): level
    energies, their pointers into nuArray, nShells(30,30) and elementP(30,30,7,1:2)."""
    nb = nu.shape[0]
    HlevEn = np.zeros(N_H_LEVEL + 1, dtype=F32)
    for i in range(1, N_H_LEVEL + 1):
        HlevEn[i - 1] = F32(H_ION_POT / F32(float(i * i)))
    HlevEn[N_H_LEVEL] = HlevEn[1]
    HeIlevEn = np.zeros(N_HEI_LEVEL + 1, dtype=F32)
    for i in range(1, N_HEI_LEVEL + 1):
        HeIlevEn[i - 1] = F32(F32(1.0) / F32(float(i * i)))
    HeIlevEn[0], HeIlevEn[1] = F32(1.80802), F32(0.2478)
    HeIlevEn[N_HEI_LEVEL] = HeIlevEn[1]
    HeIIlevEn = np.zeros(N_HEII_LEVEL + 1, dtype=F32)
    for i in range(1, N_HEII_LEVEL + 1):
        HeIIlevEn[i - 1] = F32(F32(4.0) / F32(float(i * i)))
    HeIIlevEn[N_HEII_LEVEL] = HeIIlevEn[1]

    P = dict(KshellLimitP=_loc(nu, K_SHELL_LIMIT), secIonP=_loc(nu, 7.353), cRecoilHP=_loc(nu, 194.0),
             cRecoilHeP=_loc(nu, 260.0), xrayP=_loc(nu, 20.6), BjumpP=_loc(nu, 0.25))
    HlevNuP = np.array([_loc(nu, e) for e in HlevEn[:N_H_LEVEL]] + [0], dtype=I32)
    HlevNuP[N_H_LEVEL] = HlevNuP[1]
    HeIlevNuP = np.array([_loc(nu, e) for e in HeIlevEn[:N_HEI_LEVEL]] + [0], dtype=I32)
    HeIlevNuP[N_HEI_LEVEL] = HeIlevNuP[1]
    HeIIlevNuP = np.array([_loc(nu, e) for e in HeIIlevEn[:N_HEII_LEVEL]] + [0], dtype=I32)
    HeIIlevNuP[N_HEII_LEVEL] = HeIIlevNuP[1]

    nShells = np.zeros((30, 30), dtype=I32, order="F")
    elementP = np.zeros((30, 30, 7, 3), dtype=I32, order="F")
    nShells[0, 0] = nShells[1, 0] = nShells[1, 1] = 1
    kP = P["KshellLimitP"]

    def limit_shell(el, ion, shell):
        if shell <= 3:
            return kP
        if shell <= 6:
            return int(elementP[el - 1, ion - 1, 0, 0]) - 1
        if elementP[el - 1, ion - 1, 5, 0] < 3:
            return int(elementP[el - 1, ion - 1, 4, 0]) - 1
        return int(elementP[el - 1, ion - 1, 5, 0]) - 1

    def set_shells(el):
        for ion in range(1, el + 1):
            nElec = el - ion + 1
            out = outer_shell(el, nElec)
            nShells[el - 1, ion - 1] = out
            thres = F32(0.0)
            for shell in range(1, out + 1):
                thres = F32(ph1[0, el - 1, nElec - 1, shell - 1] / RYD_TO_EV)
                if thres <= F32(0.1):
                    elementP[el - 1, ion - 1, shell - 1, 0] = 2
                    elementP[el - 1, ion - 1, shell - 1, 1] = 1
                else:
                    elementP[el - 1, ion - 1, shell - 1, 0] = _loc(nu, thres)
                    elementP[el - 1, ion - 1, shell - 1, 1] = limit_shell(el, ion, shell)
            elementP[el - 1, ion - 1, out - 1, 0] = _loc(nu, thres)      # the valence pointer (:268)

    # hydro_mod.f90:389-470: "more or less in decreasing abundance"
    for el in (6, 8, 7, 10, 11, 12, 13, 14, 15, 16, 17, 26, 18, 19, 20, 21, 22, 23, 24, 25, 9, 3, 4, 5, 27, 28, 29, 30):
        if lgElementOn[el - 1]:
            set_shells(el)
    return dict(HlevEn=HlevEn, HeIlevEn=HeIlevEn, HeIIlevEn=HeIIlevEn, HlevNuP=HlevNuP, HeIlevNuP=HeIlevNuP,
                HeIIlevNuP=HeIIlevNuP, nShells=nShells, elementP=elementP, nbins=nb, **P)


def _pow(x, y):
    with np.errstate(over="ignore", under="ignore", invalid="ignore", divide="ignore"):
        return F32(np.power(F32(x), F32(y)))


def ph_fit_el(ph1, ph2, nz: int, ne: int, shell: int, photEn) -> np.float32:
    """phFitEl (ph_mod.f90:477-552): Verner & Yakovlev 1995 / Verner et al. 1996 fits, in Mb."""
    photEn = F32(photEn)
    zero = F32(0.0)
    if nz < 1 or nz > 30 or ne < 1 or ne > nz:
        return zero
    nOut = N_TOT[ne - 1]
    if nz == ne and nz > 18:
        nOut = 7
    if nz == ne + 1 and nz in (20, 21, 22, 25, 26):
        nOut = 7
    if shell > nOut:
        return zero
    p = ph1[:, nz - 1, ne - 1, shell - 1]
    if photEn < p[0]:
        return zero
    nInt = N_INN[ne - 1]
    if nz in (15, 17, 19) or (nz > 20 and nz != 26):
        eInn = zero
    elif ne < 3:
        eInn = F32(1.0e30)
    else:
        eInn = ph1[0, nz - 1, ne - 1, nInt - 1]
    if shell < nOut and shell > nInt and photEn < eInn:
        return zero
    if shell <= nInt or photEn >= eInn:
        p1 = F32(-p[4])
        y = F32(photEn / p[1])
        q = F32(F32(F32(-0.5) * p1) - F32(LEVEL[shell - 1])) - F32(5.5)
        q = F32(q)
        ym1 = F32(y - F32(1.0))
        a = F32(p[2] * F32(F32(ym1 * ym1) + F32(p[5] * p[5])))      # **2 is a multiplication
        b = F32(np.sqrt(F32(y / p[3])) + F32(1.0))
        return F32(F32(a * _pow(y, q)) * _pow(b, p1))
    r = ph2[:, nz - 1, ne - 1]
    p1 = F32(-r[3])
    q = F32(F32(F32(-0.5) * p1) - F32(5.5))
    x = F32(F32(photEn / r[0]) - r[5])
    z = F32(np.sqrt(F32(F32(x * x) + F32(r[6] * r[6]))))
    xm1 = F32(x - F32(1.0))
    a = F32(r[1] * F32(F32(xm1 * xm1) + F32(r[4] * r[4])))
    b = F32(np.sqrt(F32(z / r[2])) + F32(1.0))
    return F32(F32(a * _pow(z, q)) * _pow(b, p1))


_HA = np.array([-17.2004, -16.8584, -16.6670, -16.5339, -16.4319, -16.3491, -16.2795, -16.2194, -16.1666, -16.1195], dtype=F32)
_HB = np.array([-2.6671, -2.8068, -2.8549, -2.8812, -2.8983, -2.9103, -2.9193, -2.9264, -2.9321, -2.9369], dtype=F32)
_HC = np.array([-0.3228, -0.1323, -0.0931, -0.0735, -0.0612, -0.0527, -0.0465, -0.0418, -0.0380, -0.0350], dtype=F32)
_HD = np.array([0.0608, 0.0224, 0.0224, 0.0204, 0.0181, 0.0163, 0.0147, 0.0135, 0.0125, 0.0117], dtype=F32)
_HE = np.array([-16.9991, -16.7709, -16.6188, -16.5012, -16.4069, -16.3289, -16.2624, -16.2046, -16.1536, -16.1078], dtype=F32)
_HF = np.array([-3.1304, -3.0041, -2.9738, -2.9671, -2.9662, -2.9669, -2.9681, -2.9694, -2.9707, -2.9719], dtype=F32)


def ph_fit_h_ion(x, n: int, z) -> np.float32:
    """phFitHIon (ph_mod.f90:560-606): hydrogenic photo-ionisation cross-section [cm^2], x = log10(W/W0)."""
    x, z = F32(x), F32(z)
    k = n - 1
    if x <= F32(1.0):
        v = F32(_HA[k] + F32(x * F32(_HB[k] + F32(x * F32(_HC[k] + F32(x * _HD[k]))))))
    else:
        v = F32(_HE[k] + F32(x * _HF[k]))
    v = _pow(10.0, v)
    return F32(v / F32(z * z))


def init_xsec_array(nu: np.ndarray, ph1, ph2, ptr: dict, lgElementOn: np.ndarray, nstages: int):
    """The gas part of initXSecArray (ph_mod.f90:204-402) with powLawXSec (:712-738) and makeOpacity
    (:742-803): the cross-section stack and the pointers into it.  Returns (xSecArray, pointers) with
    elementP(:,:,:,3) filled in `ptr["elementP"]`."""
    nb = nu.shape[0]
    xs = np.zeros(1_000_000, dtype=F32)
    top = 0                                        # xSecTop
    HlevNuP, HeIlevNuP, HeIIlevNuP = ptr["HlevNuP"], ptr["HeIlevNuP"], ptr["HeIIlevNuP"]
    elementP, nShells = ptr["elementP"], ptr["nShells"]
    e18 = F32(1.0e-18)
    log10 = lambda v: F32(np.log10(F32(v)))

    HlevXSecP = np.zeros(N_H_LEVEL, dtype=I32)
    HlevXSecP[0] = top + 1
    for i in range(int(HlevNuP[0]), nb + 1):
        thres = max(F32(nu[i - 1] * RYD_TO_EV), ph1[0, 0, 0, 0])
        xs[i - HlevNuP[0] + HlevXSecP[0] - 1] = F32(ph_fit_el(ph1, ph2, 1, 1, 1, thres) * e18)
    top += nb - int(HlevNuP[0]) + 1
    for n in range(2, N_H_LEVEL + 1):
        HlevXSecP[n - 1] = top + 1
        for i in range(int(HlevNuP[n - 1]), int(HlevNuP[0]) + 1):
            x = F32(log10(nu[i - 1]) - log10(nu[HlevNuP[n - 1] - 1]))
            xs[i - HlevNuP[n - 1] + HlevXSecP[n - 1] - 1] = ph_fit_h_ion(x, n, 1.0)
        top += int(HlevNuP[0]) - int(HlevNuP[n - 1]) + 1
    bremsXSecP = top + 1
    for i in range(1, nb + 1):
        xs[i - 1 + bremsXSecP - 1] = F32(F32(1.03680e-18) / F32(F32(nu[i - 1] * nu[i - 1]) * nu[i - 1]))   # **3: x*x*x
    top += nb

    def pow_law(low, high, cross, s):
        nonlocal top
        xSecP = top + 1
        thres = nu[low - 1]
        for i in range(low, high + 1):
            xs[i - low + xSecP - 1] = F32(F32(cross) * _pow(F32(nu[i - 1] / thres), F32(-F32(s))))
        top += high - low + 1
        return xSecP

    HeISingXSecP = np.zeros(N_HEI_LEVEL, dtype=I32)
    HeISingXSecP[0] = top + 1
    for i in range(int(HeIlevNuP[0]), nb + 1):
        xs[i - HeIlevNuP[0] + HeISingXSecP[0] - 1] = F32(ph_fit_el(ph1, ph2, 2, 2, 1, F32(nu[i - 1] * RYD_TO_EV)) * e18)
    top += nb - int(HeIlevNuP[0]) + 1
    HeISingXSecP[1] = pow_law(int(HeIlevNuP[1]), int(HeIlevNuP[0]), F32(F32(0.4) * F32(8.7e-18)), 1.5)
    for n in range(3, N_HEI_LEVEL + 1):
        cross = F32(F32(F32(7.906e-18) * F32(float(n))) / F32(1.0))
        HeISingXSecP[n - 1] = pow_law(int(HeIlevNuP[n - 1]), int(HeIlevNuP[0]), cross, 3.0)

    HeIIXSecP = np.zeros(N_HEII_LEVEL, dtype=I32)
    HeIIXSecP[0] = top + 1
    for i in range(int(HeIIlevNuP[0]), nb + 1):
        thres = max(F32(nu[i - 1] * RYD_TO_EV), ph1[0, 1, 0, 0])
        xs[i - HeIIlevNuP[0] + HeIIXSecP[0] - 1] = F32(ph_fit_el(ph1, ph2, 2, 1, 1, thres) * e18)
    top += nb - int(HeIIlevNuP[0]) + 1
    for n in range(2, N_HEII_LEVEL + 1):
        HeIIXSecP[n - 1] = top + 1
        for i in range(int(HeIIlevNuP[n - 1]), int(HeIIlevNuP[0]) + 1):
            x = F32(log10(nu[i - 1]) - log10(nu[HeIIlevNuP[n - 1] - 1]))
            xs[i - HeIIlevNuP[n - 1] + HeIIXSecP[n - 1] - 1] = ph_fit_h_ion(x, n, 2.0)
        top += int(HeIIlevNuP[0]) - int(HeIIlevNuP[n - 1]) + 1

    for el in range(3, NELEMENTS + 1):                 # makeOpacity
        if not lgElementOn[el - 1]:
            continue
        for ion in range(1, min(el, nstages) + 1):
            nElec = el - ion + 1
            for shell in range(1, int(nShells[el - 1, ion - 1]) + 1):
                lo, hi = int(elementP[el - 1, ion - 1, shell - 1, 0]), int(elementP[el - 1, ion - 1, shell - 1, 1])
                elementP[el - 1, ion - 1, shell - 1, 2] = top + 1
                if lo > hi and not (lo == 2 and hi == 1):
                    raise ValueError(f"makeOpacity: upper energy limit below the threshold [elem {el}, ion {ion}, shell {shell}]")
                for i in range(lo, hi + 1):
                    energy = max(F32(nu[i - 1] * RYD_TO_EV), ph1[0, el - 1, nElec - 1, shell - 1])
                    xs[i - lo + top] = F32(ph_fit_el(ph1, ph2, el, nElec, shell, energy) * e18)
                if hi - lo + 1 >= 1:
                    top += hi - lo + 1
    return xs[:top].copy(), dict(HlevXSecP=HlevXSecP, HeISingXSecP=HeISingXSecP, HeIIXSecP=HeIIXSecP,
                                 bremsXSecP=bremsXSecP, xSecTop=top)


def build_xsec_tables(nu, ph1, ph2, lgElementOn, elementXref, nstages: int = 7):
    """XSecTables (mocassin_b200/opacity.py) for a gas model on the mesh `nu`, plus the raw pointer
    dictionaries for the pin tests."""
    ptr = set_pointers(nu, ph1, lgElementOn)
    xs, xp = init_xsec_array(nu, ph1, ph2, ptr, lgElementOn, nstages)
    t = XSecTables(xSecArray=xs, nstages=nstages, lgElementOn=np.asarray(lgElementOn, dtype=I32),
                   elementXref=np.asarray(elementXref, dtype=I32),
                   HlevXSecP1=int(xp["HlevXSecP"][0]), HlevNuP1=int(ptr["HlevNuP"][0]),
                   HeISingXSecP1=int(xp["HeISingXSecP"][0]), HeIlevNuP1=int(ptr["HeIlevNuP"][0]),
                   HeIIXSecP1=int(xp["HeIIXSecP"][0]), HeIIlevNuP1=int(ptr["HeIIlevNuP"][0]),
                   elementP=ptr["elementP"], nShells=ptr["nShells"])
    return t, ptr, xp


def initial_ion_state(nCells: int, lgElementOn, elementXref, nElementsUsed: int, nstages: int = 7):
    """ionDen(0:nCells, nElementsUsed, nstages) as setMotherGrid leaves it (grid_mod.f90:1562-1607):
    X(H0) = 1e-5, He0 = H0, every heavy element 1e-5 neutral and nothing else."""
    ion = np.zeros((nCells + 1, nElementsUsed, nstages), dtype=F32, order="F")
    H0 = F32(1.0e-5)
    if lgElementOn[0]:
        c = elementXref[0] - 1
        ion[1:, c, 0] = H0
        ion[1:, c, 1] = F32(F32(1.0) - H0)
    if lgElementOn[1]:
        c = elementXref[1] - 1
        ion[1:, c, 0] = ion[1:, elementXref[0] - 1, 0]
        ion[1:, c, 1] = (F32(1.0) - ion[1:, c, 0]).astype(F32)
        ion[1:, c, 2] = 0.0
    for el in range(3, NELEMENTS + 1):
        if lgElementOn[el - 1]:
            c = elementXref[el - 1] - 1
            ion[1:, c, 0] = ion[1:, 0, 0]                    # sic: ionDen(cell, 1, 1)
            ion[1:, c, 1:min(el + 1, nstages)] = 0.0
    return ion
