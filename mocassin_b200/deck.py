"""Input side of a dust-only run: the reference's own deck (input.in), density, grain and
optical-constant files turned into the arrays the transport path reads.

This is the caller side of the hot path for the `benchmarks/dust` configurations (SURVEY.md
section 8d: "decks as shipped"), restated from

* ``readInput``            set_input_mod.f90:30-140 (defaults), :143-570 (keywords)
* ``initCartesianGrid``    grid_mod.f90:180-211 (frequency mesh of a noGas run = the points of
                           dustData/nuDustRyd.dat up to nuMax), :333-338 (widFlx)
* ``setMotherGrid``        grid_mod.f90:1187-1218 (`Ndust file`: axes and number densities),
                           :1226-1294 (active cells)
* ``makeDustXsec``         ph_mod.f90:808-1544 (grain sizes and weights, n,k files, mapping on the
                           mesh, cross-sections, xSecArray layout, gSca)
* ``getQs`` / ``BHmie``    ph_mod.f90:1548-1586, :1600-1757 (Mie efficiencies; Bohren & Huffman
                           with Draine's <cos>), same mixed single/double precision
* ``dustDriver``           dust_mod.f90:26-181 (Tdust = 50 K, dustEmIntegral)
* ``setProbDen``           continuum_mod.f90:418-474 (stellar CDF), deltaE set_input_mod.f90:900
* tail of ``iterateMC``    iteration_mod.f90:950-1170 (convergence, autoPackets) in `iterate_dust`

The Fortran host keeps doing all of this itself in a drop-in build; this module exists so that
the shipped decks can be run through the C ABI (and through the CPU oracle) without a Fortran
compiler.  BHmie, getQs, linearMap and the assembly tail of makeDustXsec are bit-equal to the
reference's own code run through the Fortran translator (tests/test_reference_pin.py); BHmie is
also checked against an independent Mie series (tests/test_deck.py).
"""
from __future__ import annotations

import os
import shlex
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from .model import F32, I32, Grid, Model, locate, number_active, set_star_position
from .workloads import wid_flx

C_LIGHT = F32(2.9979250e10)       # constants_mod.f90
FR1RYD = F32(3.28984e15)
HPLANCK = F32(6.6262e-27)
N_TEMPS = 3000                    # constants_mod.f90:55


# ---------------------------------------------------------------------------------------
# readInput
# ---------------------------------------------------------------------------------------
@dataclass
class Deck:
    """The keywords of input.in a dust-only run uses, with readInput's defaults."""

    lgGas: bool = True
    lgDust: bool = False
    lgSymmetricXYZ: bool = False
    lgIsotropic: bool = False
    lgDustScattering: bool = True
    lgAutoPackets: bool = False
    lgOutput: bool = False
    convIncPercent: float = 0.0
    nPhotIncrease: float = 0.0
    maxPhotons: int = 0
    contShape: str = "none"
    abundanceFile: str = "none"
    maxIterateMC: int = 30
    minConvergence: float = 95.0
    nPhotons: int = 0
    nx: int = 30
    ny: int = 30
    nz: int = 30
    nbins: int = 600
    LStar: float = 0.0
    LPhot: float = 0.0
    TStellar: float = 0.0
    nuMax: float = 15.0
    nuMin: float = 1.001e-5
    R_in: float = -1.0
    R_out: float = 0.0
    edges: Optional[tuple] = None
    NdustFile: str = "none"
    NdustValue: float = 0.0
    dustFile: Optional[tuple] = None         # (species file, sizes file)
    convWriteGrid: float = 0.0
    XHILimit: float = 0.05
    nAngleBins: int = 0
    viewPointTheta: List[float] = field(default_factory=lambda: [0.0])
    viewPointPhi: List[float] = field(default_factory=lambda: [0.0])
    starPosition: tuple = (0.0, 0.0, 0.0)
    other: dict = field(default_factory=dict)      # keywords read but not used by a dust-only run


def _tokens(line: str) -> list:
    lex = shlex.shlex(line, posix=True)
    lex.whitespace += ","
    lex.whitespace_split = True
    lex.commenters = ""
    return list(lex)


def _real(tok: str) -> float:
    return float(tok.lower().replace("d", "e"))


def read_input(path: str) -> Deck:
    """List-directed keyword file (set_input_mod.f90:143-570): first token = keyword, the rest
    its values; unknown keywords are a fatal error there and here."""
    d = Deck()
    with open(path) as fh:
        for raw in fh:
            t = _tokens(raw)
            if not t:
                continue
            k, v = t[0], t[1:]
            if k == "autoPackets":
                d.convIncPercent, d.nPhotIncrease, d.maxPhotons = _real(v[0]), _real(v[1]), int(_real(v[2]))
                d.lgAutoPackets = True
                if d.maxPhotons == 0:
                    raise ValueError("readInput: autoPackets input invalid - maximum number of photons is zero")
            elif k == "output":
                d.lgOutput = True
            elif k == "symmetricXYZ":
                d.lgSymmetricXYZ = True
            elif k == "isotropicScattering":
                d.lgIsotropic = True
            elif k == "noScattering":
                d.lgDustScattering = False
            elif k == "contShape":
                d.contShape = v[0]
            elif k == "nebComposition":
                d.abundanceFile = v[0]
                if v[0] == "noGas":
                    d.lgGas = False
            elif k == "maxIterateMC":
                d.maxIterateMC, d.minConvergence = int(_real(v[0])), _real(v[1])
            elif k == "nPhotons":
                d.nPhotons = int(_real(v[0]))
            elif k in ("nx", "ny", "nz", "nbins"):
                setattr(d, k, int(_real(v[0])))
            elif k == "LStar":
                d.LStar = _real(v[0])
            elif k == "LPhot":
                d.LPhot = _real(v[0])
            elif k == "TStellar":
                d.TStellar = _real(v[0])
            elif k in ("nuMax", "nuMin"):
                setattr(d, k, _real(v[0]))
            elif k == "Rin":
                d.R_in = _real(v[0])
            elif k == "Rout":
                d.R_out = _real(v[0])
            elif k == "edges":
                d.edges = tuple(_real(x) for x in v[:3])
            elif k == "Ndust":
                d.lgDust = True
                if v[0] == "constant":
                    d.NdustValue = _real(v[1])
                elif v[0] == "file":
                    d.NdustFile = v[1]
                else:
                    raise ValueError(f"readInput: invalid keyword in Ndust field {v[0]}")
            elif k == "dustFile":
                d.dustFile = (v[0], v[1])
            elif k == "writeGrid":
                d.convWriteGrid = _real(v[0])
            elif k == "convLimit":
                d.XHILimit = _real(v[0])
            elif k == "starPosition":
                d.starPosition = tuple(_real(x) for x in v[:3])
            elif k == "inclination":
                d.nAngleBins = int(_real(v[0]))
                if d.nAngleBins > 2:
                    raise ValueError("readInput: only two inclination anges are allowed per simulation")
                d.viewPointTheta = [0.0] + [_real(v[1 + 2 * j]) for j in range(d.nAngleBins)]
                d.viewPointPhi = [0.0] + [_real(v[2 + 2 * j]) for j in range(d.nAngleBins)]
            elif k in ("Hdensity", "TeStart", "NeStart", "H0Start", "nstages", "densityFile", "densityLaw", "talk",
                       "debug", "fillingFactor", "multiChemistry", "recombinationLines", "resLinesTransfer",
                       "multiGrids", "getEquivalentTau", "noPhotoelectric", "traceHeating", "TDust", "echo",
                       "NoSourceSED", "2D", "diffuseSource", "quantumHeatGrain", "quantumHeatGrainParameters",
                       "multiPhotoSources", "continuumCube", "slit", "inputNe", "multiDustChemistry",
                       "planeIonization", "oneD", "MdMg", "MdMh", "dustMass", "gasMass"):
                d.other[k] = v
            else:
                raise ValueError(f"readInput: unknown keyword {k!r} in {path}")
    return d


# ---------------------------------------------------------------------------------------
# frequency mesh, density file
# ---------------------------------------------------------------------------------------
def nu_mesh_dust(nu_file: str, nbins: int, nuMax: float) -> np.ndarray:
    """grid_mod.f90:180-211: the points of dustData/nuDustRyd.dat up to (excluding the first
    one above) nuMax; nbins is reset to their number."""
    vals = []
    with open(nu_file) as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if len(vals) + 1 > nbins + 1:
                raise ValueError("initCartesianGrid: nbins is smaller that the number of frequency points in "
                                 "dustData/nuDustRyd.dat file - enlarge nbins")
            x = F32(_real(t[0]))
            if x > F32(nuMax):
                break
            vals.append(x)
    return np.asarray(vals, dtype=F32)


def read_ndust(path: str, nx: int, ny: int, nz: int):
    """`Ndust file` (grid_mod.f90:1187-1218): rows `x y z Ndust`, x outermost and z innermost;
    the axes are whatever the rows say.  An optional first line `# nx ny nz` gives the extents
    (set_input_mod.f90:516-521)."""
    with open(path) as fh:
        first = fh.readline().split()
        if first and first[0] == "#":
            nx, ny, nz = int(first[1]), int(first[2]), int(first[3])
            rows = np.loadtxt(fh, dtype=np.float64)
        else:
            fh.seek(0)
            rows = np.loadtxt(fh, dtype=np.float64)
    if rows.shape[0] < nx * ny * nz:
        raise ValueError(f"{path}: {rows.shape[0]} rows for a {nx}x{ny}x{nz} grid")
    rows = rows[:nx * ny * nz].reshape(nx, ny, nz, -1)
    # the reference overwrites xAxis(i) on every row: the last row that carries index i wins
    xAxis = rows[:, -1, -1, 0].astype(F32)
    yAxis = rows[-1, :, -1, 1].astype(F32)
    zAxis = rows[-1, -1, :, 2].astype(F32)
    return xAxis, yAxis, zAxis, rows[..., 3].astype(F32), (nx, ny, nz)


# ---------------------------------------------------------------------------------------
# grains: sizes, species, optical constants, Mie
# ---------------------------------------------------------------------------------------
def normalise_grain_weights(rad: np.ndarray, w: np.ndarray) -> np.ndarray:
    """ph_mod.f90:986-1012: weights times the trapezoid widths da of the size grid, normalised to
    sum 1 (float32, running sum); a single size gets weight 1."""
    n = rad.shape[0]
    rad, w = np.asarray(rad, F32), np.asarray(w, F32)
    if n == 1:
        return np.ones(1, dtype=F32)
    da = np.zeros(n, dtype=F32)
    da[0] = rad[1] - rad[0]
    da[1:-1] = (rad[2:] - rad[:-2]) / F32(2.0)
    da[-1] = rad[-1] - rad[-2]
    norm = F32(0.0)
    for i in range(n):
        norm = F32(norm + F32(w[i] * da[i]))
    out = ((w * da).astype(F32) / norm).astype(F32)
    if not np.all(out >= 0):
        raise ValueError("makeDustXSec : Invalid grain weight")
    return out


def read_grain_sizes(path: str):
    """ph_mod.f90:958-1012: radii [um] and weights of the sizes file, weights normalised."""
    with open(path) as fh:
        n = int(fh.readline().split()[0])
        if n < 1:
            raise ValueError("makeDustXSec : Invalid nSizes")
        rad = np.zeros(n, dtype=F32)
        w = np.zeros(n, dtype=F32)
        for i in range(n):
            t = fh.readline().split()
            rad[i], w[i] = F32(_real(t[1])), F32(_real(t[2]))
    return rad, normalise_grain_weights(rad, w)


def read_grain_species(path: str):
    """Species file: count, then `'<file under share/mocassin>' abundance` rows."""
    with open(path) as fh:
        n = int(fh.readline().split()[0])
        out = []
        for _ in range(n):
            t = _tokens(fh.readline())
            out.append((t[0], F32(_real(t[1]))))
    return out


def read_nk(path: str):
    """An `nk` optical-constant file (ph_mod.f90:1063-1170): type line, `label Tsublime rho Vn
    MsurfAtom`, then `wavelength[um] n k` rows."""
    with open(path) as fh:
        kind = fh.readline().split()[0]
        if kind != "nk":
            raise NotImplementedError(f"{path}: dust file type {kind!r} (only 'nk' files are restated)")
        t = fh.readline().split()
        label, Tsub = t[0], F32(_real(t[1]))
        rows = np.array([[_real(x) for x in line.split()[:3]] for line in fh if line.split()], dtype=np.float64)
    return label, Tsub, rows[:, 0].astype(F32), rows[:, 1].astype(F32), rows[:, 2].astype(F32)


def linear_map(y: np.ndarray, x: np.ndarray, x_new: np.ndarray) -> np.ndarray:
    """interpolation_mod.f90:86-106, float32."""
    out = np.zeros(x_new.shape[0], dtype=F32)
    n = x.shape[0]
    for i, xn in enumerate(x_new):
        ii = locate(x, xn)
        if ii == 0:
            out[i] = y[0]
        elif ii == n:
            out[i] = y[-1]
        else:
            out[i] = F32(y[ii - 1] + F32(F32(F32(y[ii] - y[ii - 1]) * F32(xn - x[ii - 1])) / F32(x[ii] - x[ii - 1])))
    return out


def bhmie(x, refrel):
    """BHmie (ph_mod.f90:1600-1757): Qext, Qsca, <cos> of a sphere of size parameter x (REAL)
    and relative refractive index refrel (COMPLEX), with the reference's mix of single and
    double precision: logarithmic derivative by downward recurrence in double complex,
    Riccati-Bessel functions upward, sums kept in single precision."""
    x = F32(x)
    refrel = np.complex64(refrel)
    dx = np.float64(x)
    y = np.complex128(np.complex64(x * refrel))
    xstop = np.float64(F32(F32(x + F32(F32(4.0) * F32(np.power(x, F32(0.3333))))) + F32(2.0)))
    nstop = int(xstop)
    ymod = abs(y)
    nmx = int(max(xstop, ymod)) + 15
    if nmx > 3000:
        raise ValueError("BHmie: nmx exceeds the reference's d(3000)")
    d = np.zeros(nmx + 2, dtype=np.complex128)
    for n in range(1, nmx):
        rn = nmx - n + 1
        d[nmx - n] = (rn / y) - (1.0 / (d[nmx - n + 1] + rn / y))
    m = np.complex128(refrel)
    psi0, psi1 = np.cos(dx), np.sin(dx)
    chi0, chi1 = np.float64(F32(-np.sin(x))), np.float64(F32(np.cos(x)))
    apsi1 = psi1
    xi1 = complex(apsi1, -chi1)
    qsca = F32(0.0)
    gg = F32(0.0)
    pi0, pi1 = 0.0, 1.0                   # angle j = 1 (theta = 0) is all qext needs
    s1 = 0j
    an = bn = an1 = bn1 = 0j
    n = 1
    x64 = np.float64(x)
    while True:
        rn = n
        dn = np.float64(n)
        fn = np.float64(F32(F32(F32(2.0) * F32(rn) + F32(1.0)) / F32(F32(rn) * F32(F32(rn) + F32(1.0)))))
        psi = (2.0 * dn - 1.0) * psi1 / dx - psi0
        apsi = psi
        chi = np.float64(F32(F32(2.0) * F32(rn) - F32(1.0))) * chi1 / x64 - chi0
        xi = complex(apsi, -chi)
        if n > 1:
            an1, bn1 = an, bn
        rnx = np.float64(F32(F32(rn) / x))
        an = (d[n] / m + rnx) * apsi - apsi1
        an = an / ((d[n] / m + rnx) * xi - xi1)
        bn = (m * d[n] + rnx) * apsi - apsi1
        bn = bn / ((m * d[n] + rnx) * xi - xi1)
        f2 = np.float64(F32(F32(2.0) * F32(rn) + F32(1.0)))
        qsca = F32(qsca + F32(f2 * (abs(an) * abs(an) + abs(bn) * abs(bn))))
        inner = F32(an.real * bn.real + an.imag * bn.imag)
        gg = F32(gg + F32(F32(F32(F32(2.0) * F32(rn) + F32(1.0)) / F32(F32(rn) * F32(F32(rn) + F32(1.0)))) * inner))
        if n > 1:
            fac = np.float64(F32(F32(F32(F32(rn) - F32(1.0)) * F32(F32(rn) + F32(1.0))) / F32(rn)))
            gg = F32(gg + F32(fac * (an1.real * an.real + an1.imag * an.imag + bn1.real * bn.real + bn1.imag * bn.imag)))
        pii = pi1
        tau = rn * 1.0 * pii - np.float64(F32(F32(rn) + F32(1.0))) * pi0
        s1 = s1 + fn * (an * pii + bn * tau)
        psi0, psi1 = psi1, psi
        apsi1 = psi1
        chi0, chi1 = chi1, chi
        xi1 = complex(apsi1, -chi1)
        n += 1
        rn = n
        pi1 = np.float64(F32(F32(F32(2.0) * F32(rn) - F32(1.0)) / F32(F32(rn) - F32(1.0)))) * 1.0 * pii
        pi1 = pi1 - rn * pi0 / np.float64(F32(F32(rn) - F32(1.0)))
        pi0 = pii
        if n - 1 - nstop >= 0:
            break
    gg = F32(F32(F32(2.0) * gg) / qsca)
    qsca = F32(F32(F32(2.0) / F32(x * x)) * qsca)
    qext = F32(F32(F32(4.0) / F32(x * x)) * F32(s1.real))
    return qext, qsca, gg


def get_qs(Ere: np.ndarray, Eim: np.ndarray, radius: np.ndarray, nu: np.ndarray, scattering: bool = True):
    """getQs (ph_mod.f90:1548-1586): Qabs, Qsca, <cos> as (nSizes, nbins)."""
    nS, nb = radius.shape[0], nu.shape[0]
    Qa = np.zeros((nS, nb), dtype=F32)
    Qs = np.zeros((nS, nb), dtype=F32)
    G = np.zeros((nS, nb), dtype=F32)
    for i in range(nb):
        ref = np.complex64(complex(Ere[i], Eim[i]))
        for ai in range(nS):
            lam = F32(F32(2.9979250e14) / F32(nu[i] * FR1RYD))
            sp = F32(F32(F32(F32(2.0) * F32(3.14159265)) * radius[ai]) / lam)
            if sp > F32(100.0):
                sp = F32(100.0)
            qe, qs, g = bhmie(sp, ref)
            Qa[ai, i] = F32(qe - qs)
            Qs[ai, i] = qs if scattering else F32(0.0)
            G[ai, i] = g
    return Qa, Qs, G


def dust_efficiencies(species, radius, nu, share_dir: str, scattering: bool = True):
    """n,k files -> Qsca, Qabs, <cos> (nSpecies, nSizes, nbins) on the frequency mesh, plus the
    abundances, sublimation temperatures and labels (ph_mod.f90:1063-1201)."""
    nb, nSp, nSz = nu.shape[0], len(species), radius.shape[0]
    Qsca = np.zeros((nSp, nSz, nb), dtype=F32)
    Qabs = np.zeros((nSp, nSz, nb), dtype=F32)
    gCos = np.zeros((nSp, nSz, nb), dtype=F32)
    abun = np.zeros(nSp, dtype=F32)
    Tsub = np.zeros(nSp, dtype=F32)
    labels = []
    for s, (fname, ab) in enumerate(species):
        label, ts, wav, n_re, k_im = read_nk(os.path.join(share_dir, fname))
        labels.append(label)
        abun[s], Tsub[s] = ab, ts
        # wavelength [um] -> energy [Ryd], reversed so the table ascends (:1172-1180)
        e = (C_LIGHT / (wav * FR1RYD * F32(1.0e-4))).astype(F32)[::-1].copy()
        Ere = linear_map(n_re[::-1].copy(), e, nu)
        Eim = linear_map(k_im[::-1].copy(), e, nu)
        Qabs[s], Qsca[s], gCos[s] = get_qs(Ere, Eim, radius, nu, scattering)
    return Qsca, Qabs, gCos, abun, Tsub, labels


def make_dust_xsec(species, radius, weight, nu, share_dir: str, scattering: bool = True):
    """makeDustXsec for ONE dust component of `nk` species (ph_mod.f90:808-1544).  Returns the
    dust part of xSecArray in the reference's layout -- [CTsca, CTabs, then Csca, Cabs per
    (species, size)], cross-sections in cm^2 -- with the 1-based pointer tables
    dustScaXsecP/dustAbsXsecP(0:nSpecies, nSizes), gSca, grainAbun, TdustSublime, labels."""
    Qsca, Qabs, gCos, abun, Tsub, labels = dust_efficiencies(species, radius, nu, share_dir, scattering)
    out = assemble_dust_xsec(Qsca, Qabs, gCos, radius, weight, abun)
    out.update(grainAbun=abun, TdustSublime=Tsub, grainLabel=labels)
    return out


def assemble_dust_xsec(Qsca, Qabs, gCos, radius, weight, abun):
    """The tail of makeDustXsec's component loop (ph_mod.f90:1456-1538), same operation order."""
    nSp, nSz, nb = Qsca.shape
    PI = F32(3.141592654)
    Csca = np.zeros((nSp, nSz, nb), dtype=F32)
    Cabs = np.zeros((nSp, nSz, nb), dtype=F32)
    for s in range(nSp):
        for ai in range(nSz):
            # ((Q*Pi)*a)*a*1e-8, left to right (:1460-1461)
            Csca[s, ai] = (((Qsca[s, ai] * PI).astype(F32) * radius[ai]).astype(F32) * radius[ai]).astype(F32) * F32(1.0e-8)
            Cabs[s, ai] = (((Qabs[s, ai] * PI).astype(F32) * radius[ai]).astype(F32) * radius[ai]).astype(F32) * F32(1.0e-8)
    CTsca = np.zeros(nb, dtype=F32)
    CTabs = np.zeros(nb, dtype=F32)
    for s in range(nSp):
        for ai in range(nSz):
            CTsca = (CTsca + ((abun[s] * Csca[s, ai]).astype(F32) * weight[ai]).astype(F32)).astype(F32)
            CTabs = (CTabs + ((abun[s] * Cabs[s, ai]).astype(F32) * weight[ai]).astype(F32)).astype(F32)
    blocks = [CTsca, CTabs]
    scaP = np.full((nSp + 1, nSz), -1, dtype=I32, order="F")
    absP = np.full((nSp + 1, nSz), -1, dtype=I32, order="F")
    scaP[0, :] = 1
    absP[0, :] = 1 + nb
    nn = 2
    for s in range(nSp):
        for ai in range(nSz):
            scaP[s + 1, ai] = 1 + nn * nb; blocks.append(Csca[s, ai]); nn += 1
            absP[s + 1, ai] = 1 + nn * nb; blocks.append(Cabs[s, ai]); nn += 1
    # xSecTop advances by 2*nbins*(nSpecies+1)*nSizes (:1507), more than the 2+2*nSpecies*nSizes
    # blocks written when nSizes > 1: the tail stays zero, the next component would start after it
    xSecTop = 2 * nb * (nSp + 1) * nSz
    xSec = np.zeros(xSecTop, dtype=F32)
    flat = np.concatenate(blocks).astype(F32)
    xSec[:flat.shape[0]] = flat
    gS = np.zeros(nb, dtype=F32)
    norm = np.zeros(nb, dtype=F32)
    for s in range(nSp):
        for ai in range(nSz):
            a2 = F32(F32(PI * F32(radius[ai] * radius[ai])))
            gS = (gS + (((gCos[s, ai] * PI).astype(F32) * F32(radius[ai] * radius[ai])).astype(F32) * weight[ai]).astype(F32)
                  * abun[s]).astype(F32)
            norm = (norm + F32(F32(a2 * weight[ai]) * abun[s])).astype(F32)
    gS = (gS / norm).astype(F32)
    return dict(xSecArray=xSec, dustScaXsecP=scaP, dustAbsXsecP=absP, gSca=gS)


def dust_em_integral(xSec, absP, nu, widFlx, nTemps: int = N_TEMPS) -> np.ndarray:
    """dustEmissionInt (dust_mod.f90:145-181): (nSpecies, nSizes, nTemps), T = 1..nTemps K, in the
    reference's float32: per temperature the running sum over the bins of
    ((xSec*getFlux)*fr1Ryd)*widFlx, then *hPlanck*4."""
    nSp, nSz = absP.shape[0] - 1, absP.shape[1]
    nb = nu.shape[0]
    T = np.arange(1, nTemps + 1, dtype=F32)
    w = np.asarray(widFlx, dtype=F32)
    bb = [get_flux_blackbody(nu[i], T) for i in range(nb)]     # (nbins)(nTemps)
    em = np.zeros((nSp, nSz, nTemps), dtype=F32, order="F")
    for s in range(nSp):
        for ai in range(nSz):
            o = int(absP[s + 1, ai]) - 1
            acc = np.zeros(nTemps, dtype=F32)
            for i in range(nb):
                acc = (acc + (((xSec[o + i] * bb[i]).astype(F32) * FR1RYD).astype(F32) * w[i]).astype(F32)).astype(F32)
            em[s, ai, :] = ((acc * HPLANCK).astype(F32) * F32(4.0)).astype(F32)
    return em


def get_flux_blackbody(nu: np.ndarray, T) -> np.ndarray:
    """getFlux(nu, T, 'blackbody') (continuum_mod.f90:359-401) in the reference's float32: Planck
    function / h, Wien form above h nu / k T = 86 (with its double-precision exp), Rayleigh-Jeans
    where exp(x) - 1 rounds to zero."""
    e = np.asarray(nu, dtype=F32)
    T = np.asarray(T, dtype=F32)         # nu and T broadcast against each other
    const = F32(F32(0.5250229) / HPLANCK)
    x = (F32(157893.94) * e).astype(F32) / T
    e3 = (((const * e).astype(F32) * e).astype(F32) * e).astype(F32)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        wien = (e3.astype(np.float64) * np.exp((-x).astype(F32).astype(np.float64))).astype(F32)
        den = (np.exp(x).astype(F32) - F32(1.0)).astype(F32)
        rj = ((((F32(3.32154e-6) * e).astype(F32) * e).astype(F32) * T).astype(F32) / HPLANCK).astype(F32)
        planck = (e3 / den).astype(F32)
    return np.where(x > F32(86.0), wien, np.where(den <= 0, rj, planck)).astype(F32)


def stellar_cdf(T, nu: np.ndarray, widFlx: np.ndarray) -> np.ndarray:
    """inSpectrumProbDen of a blackbody source: setContinuum's inSpectrumErg = getFlux (:112) and
    setProbDen (continuum_mod.f90:418-474) with its mix of REAL and DOUBLE PRECISION."""
    f = get_flux_blackbody(nu, T).astype(np.float64)           # inSpSumErg = real(inSpectrumErg)
    w = np.asarray(widFlx, dtype=F32)
    norm = F32(0.0)
    for i in range(f.shape[0]):
        norm = F32(np.float64(norm) + f[i] * np.float64(w[i]))
    cdf = np.zeros(f.shape[0], dtype=F32)
    cdf[0] = F32(f[0] * np.float64(w[0]) / np.float64(norm))
    for i in range(1, f.shape[0]):
        cdf[i] = F32(np.float64(cdf[i - 1]) + f[i] * np.float64(w[i]) / np.float64(norm))
    cdf[cdf >= cdf.max()] = F32(1.0)
    return cdf


def dust_opacity(g: Grid, tables: dict) -> None:
    """Dust opacities of a single-component dust-only grid (iteration_mod.f90:166-227):
    scaOpac/absOpac(cell, nu) = sum over (species, size) with Tdust < TdustSublime of
    grainAbun*grainWeight*Ndust * xSec; opacity = their sum.  Recomputed every iteration, as
    the reference does, because grains above their sublimation temperature drop out."""
    xSec, weight = tables["xSecArray"], tables["grainWeight"]
    abun, Tsub = tables["grainAbun1"], tables["TdustSublime"]
    nSp, nSz = tables["dustAbsXsecP"].shape
    nCells, nbins = g.nCells, tables["widFlx"].shape[0]
    g.scaOpac = np.zeros((nCells + 1, nbins), dtype=F32, order="F")
    g.absOpac = np.zeros((nCells + 1, nbins), dtype=F32, order="F")
    for s in range(nSp):
        for ai in range(nSz):
            on = g.Tdust[s + 1, ai + 1, :] < Tsub[s]
            coef = np.where(on, (F32(abun[s] * weight[ai]) * g.Ndust).astype(F32), F32(0.0)).astype(F32)
            o = int(tables["dustScaXsecP"][s, ai]) - 1
            g.scaOpac += (coef[:, None] * xSec[None, o:o + nbins]).astype(F32)
            o = int(tables["dustAbsXsecP"][s, ai]) - 1
            g.absOpac += (coef[:, None] * xSec[None, o:o + nbins]).astype(F32)
    g.scaOpac[0, :] = 0
    g.absOpac[0, :] = 0
    g.opacity = (g.scaOpac + g.absOpac).astype(F32, order="F")


# ---------------------------------------------------------------------------------------
# the whole deck
# ---------------------------------------------------------------------------------------
def load_dust_deck(run_dir: str, share_dir: str, input_file: str = "input.in"):
    """Build (Model, tables, Deck) for a dust-only deck whose files sit in `run_dir` (the deck's
    own `input/...` paths are resolved against it, with or without the `input/` prefix) and
    whose data files (`dustData/...`) sit under `share_dir` (= PREFIX/share/mocassin)."""
    d = read_input(os.path.join(run_dir, input_file))
    if d.lgGas or not d.lgDust:
        raise NotImplementedError("only dust-only decks (nebComposition noGas + Ndust) are restated")
    if d.contShape != "blackbody":
        raise NotImplementedError(f"contShape {d.contShape}")
    if d.dustFile is None:
        raise ValueError("readInput: dust present but no dustFile given")
    if d.NdustFile == "none" and not d.NdustValue > 0:
        raise NotImplementedError("dust-only deck without an Ndust keyword (MdMg / MdMh are gas-relative)")

    def resolve(p):
        for cand in (os.path.join(run_dir, p), os.path.join(run_dir, os.path.basename(p))):
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError(p)

    nu = nu_mesh_dust(os.path.join(share_dir, "dustData", "nuDustRyd.dat"), d.nbins, d.nuMax)
    nbins = nu.shape[0]
    widFlx = wid_flx(nu)
    if d.NdustFile != "none":
        xA, yA, zA, Nd3, (nx, ny, nz) = read_ndust(resolve(d.NdustFile), d.nx, d.ny, d.nz)
    else:
        # `Ndust constant`: automatic axes from the `edges` keyword (fillGrid, grid_mod.f90:530-601;
        # without it readInput stops, set_input_mod.f90:720-723), the same density in every cell
        from .model import auto_axis

        if d.edges is None or min(d.edges) < 0:
            raise ValueError("readInput: Grid edges unspecified or non-valid grid edges")
        if not d.lgSymmetricXYZ and (d.nx % 2 == 0 or d.ny % 2 == 0 or d.nz % 2 == 0):
            raise ValueError("fillGrid: the automatic grid option requires odd integer nx, ny, nz if not symmetric")
        nx, ny, nz = d.nx, d.ny, d.nz
        xA, yA, zA = (auto_axis(n, e, d.lgSymmetricXYZ) for n, e in zip((nx, ny, nz), d.edges))
        Nd3 = np.full((nx, ny, nz), F32(d.NdustValue), dtype=F32)
    # active cells (grid_mod.f90:1226-1294): inside [R_in, R_out] and Ndust > 0
    r = F32(1.0e10) * np.sqrt(((xA / F32(1.0e10)) ** 2)[:, None, None] + ((yA / F32(1.0e10)) ** 2)[None, :, None]
                              + ((zA / F32(1.0e10)) ** 2)[None, None, :]).astype(F32)
    mask = (Nd3 > 0) & ~(r < F32(d.R_in))
    if d.R_out > 0:
        mask &= ~(r > F32(d.R_out))
    active, nCells = number_active(mask)
    g = Grid(xAxis=xA, yAxis=yA, zAxis=zA, active=active, nCells=nCells)
    nd = np.zeros(nCells + 1, dtype=F32)
    nd[active[mask]] = Nd3[mask]
    g.Ndust = nd
    g.dustAbunIndex = np.ones(nCells + 1, dtype=I32)

    radius, weight = read_grain_sizes(resolve(d.dustFile[1]))
    species = read_grain_species(resolve(d.dustFile[0]))
    xs = make_dust_xsec(species, radius, weight, nu, share_dir, d.lgDustScattering)
    nSp, nSz = len(species), radius.shape[0]
    xSec = xs["xSecArray"]
    em = dust_em_integral(xSec, xs["dustAbsXsecP"], nu, widFlx)
    grainAbun = np.zeros((1, nSp), dtype=F32, order="F")
    grainAbun[0, :] = xs["grainAbun"]

    g.Tdust = np.zeros((nSp + 1, nSz + 1, nCells + 1), dtype=F32, order="F")
    g.Tdust[...] = F32(50.0)                                   # dust_mod.f90:65
    tables = dict(xSecArray=xSec, dustAbsXsecP=np.asfortranarray(xs["dustAbsXsecP"][1:, :]),
                  dustScaXsecP=np.asfortranarray(xs["dustScaXsecP"][1:, :]), grainWeight=weight, widFlx=widFlx,
                  dustEmIntegral=em, grainRadius=radius, grainLabel=xs["grainLabel"],
                  grainAbun1=xs["grainAbun"], TdustSublime=xs["TdustSublime"])
    dust_opacity(g, tables)

    wid = widFlx                                               # setProbDen uses widFlx
    cdf = stellar_cdf(d.TStellar, nu, wid)
    pos, sidx = set_star_position([g], [list(d.starPosition)])   # the keyword is in units of the axis ends
    nPhot = int(d.nPhotons)
    view = {}
    if d.nAngleBins > 0:
        if d.lgSymmetricXYZ and any(F32(th) > F32(F32(3.141592654) / F32(2.0)) for th in d.viewPointTheta[1:]):
            raise ValueError("initCartesianGrid: the inclination theta required is not available for symmetricXYZ "
                             "models (theta > Pi/2)")          # grid_mod.f90:462-465
        view = dict(nAngleBins=d.nAngleBins, viewPointTheta=np.asarray(d.viewPointTheta, dtype=F32),
                    viewPointPhi=np.asarray(d.viewPointPhi, dtype=F32))
    model = Model(grids=[g], nbins=nbins, nuArray=nu,
                  inSpectrumProbDen=np.stack([np.zeros(nbins, F32), cdf]).astype(F32),
                  deltaE=np.asarray([0.0, F32(d.LStar) / F32(nPhot)], dtype=F32),
                  starPosition=np.asarray(pos, dtype=F32), starIndeces=np.asarray(sidx, dtype=I32),
                  lgDust=True, lgGas=False, lgSymmetricXYZ=d.lgSymmetricXYZ, lgIsotropic=d.lgIsotropic,
                  R_out=float(d.R_out), gSca=xs["gSca"], nSpeciesMax=nSp, nSizes=nSz,
                  nSpeciesPart=np.asarray([nSp], dtype=I32), grainAbun=grainAbun,
                  dustComPoint=np.asarray([1], dtype=I32), TdustSublime=xs["TdustSublime"], **view)
    return model, tables, d


_GOLDEN_ARRAYS = ("nuArray", "xAxis", "yAxis", "zAxis", "active", "Ndust", "cdf", "gSca", "xSecArray", "dustAbsXsecP",
                  "dustScaXsecP", "grainWeight", "grainRadius", "grainAbun1", "TdustSublime", "dustEmIntegral")


def deck_to_arrays(model: Model, tables: dict, d: Deck) -> dict:
    """Everything load_dust_deck produced, as plain arrays (the fixture format of
    tests/golden/deck_*.npz: the decks' data files do not travel to the GPU box)."""
    import json

    g = model.grids[0]
    scal = {k: getattr(d, k) for k in ("lgSymmetricXYZ", "lgIsotropic", "lgAutoPackets", "convIncPercent", "nPhotIncrease",
                                       "maxPhotons", "maxIterateMC", "minConvergence", "nPhotons", "LStar", "TStellar",
                                       "R_in", "R_out", "XHILimit", "nAngleBins", "viewPointTheta", "viewPointPhi",
                                       "starPosition")}
    out = dict(nuArray=model.nuArray, xAxis=g.xAxis, yAxis=g.yAxis, zAxis=g.zAxis, active=g.active, Ndust=g.Ndust,
               cdf=model.inSpectrumProbDen[1], gSca=model.gSca, deck_json=np.frombuffer(json.dumps(scal).encode(), dtype=np.uint8))
    for k in _GOLDEN_ARRAYS[8:]:
        out[k] = tables[k]
    return out


def deck_from_arrays(a: dict):
    """Inverse of deck_to_arrays: (Model, tables, Deck) without touching the deck's files."""
    import json

    scal = json.loads(bytes(a["deck_json"]).decode())
    d = Deck(lgGas=False, lgDust=True, contShape="blackbody")
    for k, v in scal.items():
        setattr(d, k, tuple(v) if k == "starPosition" else v)
    nu = np.asarray(a["nuArray"], dtype=F32)
    nbins = nu.shape[0]
    active = np.asfortranarray(a["active"], dtype=I32)
    g = Grid(xAxis=np.asarray(a["xAxis"], F32), yAxis=np.asarray(a["yAxis"], F32), zAxis=np.asarray(a["zAxis"], F32),
             active=active, nCells=int(active.max()))
    g.Ndust = np.asarray(a["Ndust"], F32)
    g.dustAbunIndex = np.ones(g.nCells + 1, dtype=I32)
    tables = {k: np.asfortranarray(a[k]) for k in _GOLDEN_ARRAYS[8:]}
    tables["widFlx"] = wid_flx(nu)
    nSp, nSz = tables["dustAbsXsecP"].shape
    g.Tdust = np.full((nSp + 1, nSz + 1, g.nCells + 1), F32(50.0), dtype=F32, order="F")
    dust_opacity(g, tables)
    grainAbun = np.zeros((1, nSp), dtype=F32, order="F")
    grainAbun[0, :] = tables["grainAbun1"]
    view = {}
    if d.nAngleBins > 0:
        view = dict(nAngleBins=d.nAngleBins, viewPointTheta=np.asarray(d.viewPointTheta, dtype=F32),
                    viewPointPhi=np.asarray(d.viewPointPhi, dtype=F32))
    model = Model(grids=[g], nbins=nbins, nuArray=nu,
                  inSpectrumProbDen=np.stack([np.zeros(nbins, F32), np.asarray(a["cdf"], F32)]).astype(F32),
                  deltaE=np.asarray([0.0, F32(d.LStar) / F32(int(d.nPhotons))], dtype=F32),
                  starPosition=set_star_position([g], [list(d.starPosition)])[0],
                  starIndeces=set_star_position([g], [list(d.starPosition)])[1],
                  lgDust=True, lgGas=False, lgSymmetricXYZ=d.lgSymmetricXYZ, lgIsotropic=d.lgIsotropic,
                  R_out=float(d.R_out), gSca=np.asarray(a["gSca"], F32), nSpeciesMax=nSp, nSizes=nSz,
                  nSpeciesPart=np.asarray([nSp], dtype=I32), grainAbun=grainAbun,
                  dustComPoint=np.asarray([1], dtype=I32), TdustSublime=np.asarray(tables["TdustSublime"], F32), **view)
    return model, tables, d


# ---------------------------------------------------------------------------------------
# the Lucy iteration of a dust-only run
# ---------------------------------------------------------------------------------------
def iterate_dust(deck: Deck, model: Model, step: Callable[[int, float], tuple], log: Optional[Callable] = None):
    """Outer loop of iterateMC for a dust-only run (iteration_mod.f90:31, :950-1170):
    `step(nPhotons, deltaE) -> (nConverged, nCells)` does one iteration (setDustPDF ->
    energyPacketDriver -> getDustT).  Mirrors the convergence test against minConvergence,
    maxIterateMC and the autoPackets rule (packets doubled when the converged fraction grew
    by <= convIncPercent, with the reference's double count of star 1 in nPhotonsTot)."""
    nPhotons = int(deck.nPhotons)
    deltaE = F32(model.deltaE[1])
    totOld = F32(0.0)
    hist = []
    for it in range(1, deck.maxIterateMC + 1):
        nconv, ncells = step(nPhotons, float(deltaE))
        # float32 as in the reference (:982-1064): per grid 100*conv/totCells, weighted by nCells/100,
        # then 100*sum/totCells
        conv = F32(F32(100.0) * F32(nconv)) / F32(ncells) if ncells else F32(0.0)
        tot = F32(F32(100.0) * F32(F32(conv * F32(ncells)) / F32(100.0))) / F32(ncells) if ncells else F32(0.0)
        hist.append(dict(iteration=it, converged_pct=float(tot), nPhotons=nPhotons))
        if log:
            log(hist[-1])
        nTot = 2 * nPhotons               # nPhotonsTot = nPhotons(1) + sum over stars (:1106-1109)
        if it > 1 and tot < F32(95.0) and deck.lgAutoPackets and nTot < deck.maxPhotons and totOld > 0:
            if F32(F32(tot - totOld) / totOld) <= F32(deck.convIncPercent):
                # nint(): half away from zero (iteration_mod.f90:1113), not numpy's half-to-even
                nPhotons = int(np.floor(np.float64(F32(nPhotons) * F32(deck.nPhotIncrease)) + 0.5))
                deltaE = F32(deltaE / F32(deck.nPhotIncrease))
        totOld = tot
        if tot >= F32(deck.minConvergence):
            break
    return hist
