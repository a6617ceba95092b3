"""Checkpoint wire formats of the reference (writeGrid, grid_mod.f90:2646-2870; reader resetGrid,
:2967-3567): ``grid0.out`` (geometry, active map, convergence flags), ``grid1.out`` (Te, Ne, Hden
[, abFileIndex]), ``grid2.out`` (ionDen), ``grid3.out`` (run parameters), ``dustGrid.out`` (Ndust,
dustAbunIndex, Tdust) and ``photoSource.out``.  Lets the harness hand a state over to (or
warm-start from) a real mocassin run (``mocassinWarm``).

The files are Fortran list-directed text: one record per line, blank separated, read back with
``read(unit,*)`` -- so any whitespace layout round-trips; numbers are written with 9 significant
digits (float32 round-trip exact), logicals as T / F.  The record sequences are checked against
the records the reference's own writeGrid writes (tests/test_reference_pin.py)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .model import F32, I32, Grid, Model


def _r(x) -> str:
    return f"{float(x):.9G}"


def write_grid0(path: str, model: Model, lgConverged: Optional[Sequence[np.ndarray]] = None,
                lgBlack: Optional[Sequence[np.ndarray]] = None, lg2D: bool = False) -> None:
    """grid0.out, grid_mod.f90:2692-2727: per grid ``nGrids``; ``nx ny nz nCells motherP R_out``;
    the three axes one value per line; then for x (slowest), y, z (fastest) one line
    ``active lgConverged lgBlack`` (flags of cell 0 for inactive / sub-grid cells)."""
    with open(path, "w") as fh:
        for iG, g in enumerate(model.grids):
            conv = np.zeros(g.nCells + 1, I32) if lgConverged is None else np.asarray(lgConverged[iG], I32)
            black = np.zeros(g.nCells + 1, I32) if lgBlack is None else np.asarray(lgBlack[iG], I32)
            fh.write(f" {model.nGrids}\n")
            fh.write(f" {g.nx} {g.ny} {g.nz} {g.nCells} {g.motherP} {_r(F32(model.R_out))}\n")
            for ax in (g.xAxis, g.yAxis, g.zAxis):
                for v in ax:
                    fh.write(f" {_r(v)}\n")
            yTop = 1 if (iG == 0 and lg2D) else g.ny
            act = np.asarray(g.active)
            lines = []
            for i in range(g.nx):
                for j in range(yTop):
                    a = act[i, j, :]
                    c = np.where(a < 0, 0, a)
                    lines.extend(f" {int(a[k])} {int(conv[c[k]])} {int(black[c[k]])}\n" for k in range(g.nz))
            fh.writelines(lines)


def _ytop(iG: int, g, lg2D: bool) -> int:
    """Planes of y a checkpoint file holds for grid iG (0-based): writeGrid writes only j = 1 of the
    mother grid of a 2D run (grid_mod.f90:2709-2713).  (resetGrid, :3412-3416, applies yTop = 1 to
    every grid it reads; the files it reads are writeGrid's, so the readers here follow the writer.)"""
    return 1 if (iG == 0 and lg2D) else g.ny


def fill_2d_planes(g: Grid) -> np.ndarray:
    """resetGrid's completion of a 2D mother grid (grid_mod.f90:3476-3500): the cells of the planes
    j >= 2 point at the cell of plane 1 whose x is nearest to their cylindrical radius
    sqrt(x^2 + y^2); returns TwoDscaleJ(1:nCells) (index 0 unused) = cells sharing each id."""
    from .model import locate

    act = np.asarray(g.active)
    scale = np.ones(g.nCells + 1, F32)
    x = np.asarray(g.xAxis, F32)
    for i in range(g.nx):
        for j in range(1, g.ny):
            a, b = F32(x[i] / F32(1.e10)), F32(g.yAxis[j] / F32(1.e10))
            radius = F32(F32(1.e10) * np.sqrt(F32(F32(a * a) + F32(b * b))))
            xP = locate(x, radius)
            if xP < g.nx:
                if radius >= F32(F32(x[xP - 1] + x[xP]) / F32(2.0)):
                    xP += 1
            for k in range(g.nz):
                act[i, j, k] = act[xP - 1, 0, k]
                if act[xP - 1, 0, k] > 0:
                    scale[act[xP - 1, 0, k]] += F32(1.0)
    return scale


def read_grid0(path: str, lg2D: bool = False):
    """Inverse of write_grid0 -> (grids [axes, active, nCells, motherP], R_out, lgConverged, lgBlack).
    lg2D: only plane j = 1 of the mother grid is in the file; the other planes are completed as
    resetGrid does (:func:`fill_2d_planes`)."""
    tok = open(path).read().split()
    p = 0
    grids: List[Grid] = []
    convs, blacks = [], []
    R_out = 0.0
    nGrids = None
    while p < len(tok):
        nG = int(tok[p]); p += 1
        nGrids = nG if nGrids is None else nGrids
        nx, ny, nz, nCells, motherP = (int(t) for t in tok[p:p + 5])
        R_out = float(tok[p + 5]); p += 6
        ax = []
        for n in (nx, ny, nz):
            ax.append(np.array([float(t) for t in tok[p:p + n]], dtype=F32)); p += n
        yTop = 1 if (lg2D and not grids) else ny
        vals = np.array([int(t) for t in tok[p:p + 3 * nx * yTop * nz]], dtype=np.int64).reshape(nx, yTop, nz, 3); p += 3 * nx * yTop * nz
        active = np.zeros((nx, ny, nz), I32, order="F")
        active[:, :yTop, :] = vals[..., 0]
        conv = np.zeros(nCells + 1, I32); black = np.zeros(nCells + 1, I32)
        m = vals[..., 0] > 0
        conv[vals[..., 0][m]] = vals[..., 1][m]; black[vals[..., 0][m]] = vals[..., 2][m]
        grids.append(Grid(xAxis=ax[0], yAxis=ax[1], zAxis=ax[2], active=active, nCells=nCells, motherP=motherP))
        if yTop != ny:
            fill_2d_planes(grids[-1])
        convs.append(conv); blacks.append(black)
        if len(grids) == nGrids:
            break
    return grids, R_out, convs, blacks


def write_dust_grid(path: str, model: Model, lgMultiChemistry: bool = False, totalDustMass: float = 0.0,
                    lg2D: bool = False) -> None:
    """dustGrid.out, grid_mod.f90:2746-2757,2771-2775: per cell of the x,y,z loop ``Ndust``
    (``Ndust dustAbunIndex`` with lgMultiChemistry -- sic, the gas flag) and nSizes+1 lines of
    ``Tdust(0:nSpeciesMax, ai, cell)``; cell 0 stands in for inactive cells."""
    with open(path, "w") as fh:
        for iG, g in enumerate(model.grids):
            act = np.asarray(g.active)
            T = np.asarray(g.Tdust, dtype=F32)
            lines = []
            for i in range(g.nx):
                for j in range(_ytop(iG, g, lg2D)):
                    for k in range(g.nz):
                        c = int(act[i, j, k])
                        c = 0 if c < 0 else c
                        if lgMultiChemistry:
                            lines.append(f" {_r(g.Ndust[c])} {int(g.dustAbunIndex[c])}\n")
                        else:
                            lines.append(f" {_r(g.Ndust[c])}\n")
                        for ai in range(model.nSizes + 1):
                            lines.append(" " + "   ".join(_r(T[e, ai, c]) for e in range(model.nSpeciesMax + 1)) + "\n")
            fh.writelines(lines)
        fh.write("  \n")
        fh.write(f" Total dust mass [1.e45 g]:  {_r(F32(totalDustMass))}\n")
        fh.write(f" Total dust mass [Msol]:  {_r(F32(totalDustMass) * F32(5.028e11))}\n")


def read_dust_grid(path: str, model: Model, lgMultiChemistry: bool = False, lg2D: bool = False) -> None:
    """Fill Ndust, dustAbunIndex and Tdust of model.grids from dustGrid.out (resetGrid's dust part)."""
    tok = open(path).read().split()
    p = 0
    nS, nZ = model.nSpeciesMax + 1, model.nSizes + 1
    per = (2 if lgMultiChemistry else 1) + nS * nZ
    for iG, g in enumerate(model.grids):
        yTop = _ytop(iG, g, lg2D)
        n = g.nx * yTop * g.nz
        block = tok[p:p + per * n]; p += per * n
        a = np.array([float(t) for t in block], dtype=np.float64).reshape(g.nx, yTop, g.nz, per)
        act = np.asarray(g.active)[:, :yTop, :]
        m = act > 0
        g.Ndust = np.zeros(g.nCells + 1, F32)
        g.Ndust[act[m]] = a[..., 0][m]
        off = 1
        if lgMultiChemistry:
            g.dustAbunIndex = np.ones(g.nCells + 1, I32)
            g.dustAbunIndex[act[m]] = a[..., 1][m].astype(I32)
            off = 2
        T = np.zeros((nS, nZ, g.nCells + 1), dtype=F32, order="F")
        body = a[..., off:].reshape(g.nx, yTop, g.nz, nZ, nS)       # lines: ai, values: species
        T[:, :, act[m]] = np.transpose(body[m], (2, 1, 0))
        g.Tdust = T


def write_photo_source(path: str, model: Model, contShape: Sequence[str], TStellar: Sequence[float],
                       LStar: Sequence[float], nPhotons: Sequence[int], spID: Optional[Sequence[str]] = None,
                       tStep: Optional[Sequence[float]] = None) -> None:
    """photoSource.out, grid_mod.f90:2787-2798: positions in units of the mother grid's last axis point."""
    g = model.grids[0]
    with open(path, "w") as fh:
        fh.write(f" {model.nStars}  number of photon sources\n")
        for i in range(model.nStars):
            x, y, z = (float(v) for v in model.starPosition[i])
            fh.write(f" '{contShape[i]}' {_r(TStellar[i])} {_r(LStar[i])} {int(nPhotons[i])} "
                     f"{_r(F32(x) / g.xAxis[-1])} {_r(F32(y) / g.yAxis[-1])} {_r(F32(z) / g.zAxis[-1])} "
                     f"{(spID[i] if spID else 'mocassin')} {_r(tStep[i] if tStep else 0.0)}\n")
        fh.write(" (contShape, T_eff[K], L_* [E36 erg/s], nPackets, (x,y,z) position, spID, tstep)\n")


def read_photo_source(path: str):
    """-> list of dict(contShape, TStellar, LStar, nPhotons, position (relative), spID, tStep)."""
    lines = open(path).read().splitlines()
    n = int(lines[0].split()[0])
    out = []
    for ln in lines[1:1 + n]:
        t = ln.replace("'", " ").split()
        out.append(dict(contShape=t[0], TStellar=float(t[1]), LStar=float(t[2]), nPhotons=int(t[3]),
                        position=(float(t[4]), float(t[5]), float(t[6])), spID=t[7], tStep=float(t[8])))
    return out


def _l(x) -> str:
    return "T" if x else "F"


def write_grid1(path: str, model: Model, Te: Sequence[np.ndarray], Ne: Sequence[np.ndarray],
                abFileIndex: Optional[Sequence[np.ndarray]] = None, lg2D: bool = False) -> None:
    """grid1.out, grid_mod.f90:2730-2737: per cell of the x, y, z loop ``Te Ne Hden`` (cell 0 for
    inactive cells), plus ``abFileIndex(i,j,k)`` when abFileIndex is given (lgMultiChemistry)."""
    with open(path, "w") as fh:
        for iG, g in enumerate(model.grids):
            act = np.asarray(g.active)
            yTop = 1 if (iG == 0 and lg2D) else g.ny
            lines = []
            for i in range(g.nx):
                for j in range(yTop):
                    for k in range(g.nz):
                        c = max(int(act[i, j, k]), 0)
                        rec = f" {_r(Te[iG][c])} {_r(Ne[iG][c])} {_r(g.Hden[c])}"
                        if abFileIndex is not None:
                            rec += f" {int(abFileIndex[iG][i, j, k])}"
                        lines.append(rec + "\n")
            fh.writelines(lines)


def read_grid1(path: str, grids: Sequence[Grid], multi_chemistry: bool = False, lg2D: bool = False):
    """-> (Te, Ne, Hden, abFileIndex) lists per grid, arrays (0:nCells) / (nx,ny,nz) (abFileIndex of a
    2D mother grid: plane j = 1 only, the other planes stay 1 as resetGrid leaves them)."""
    tok = open(path).read().split()
    p = 0
    per = 4 if multi_chemistry else 3
    out = ([], [], [], [])
    for iG, g in enumerate(grids):
        yTop = _ytop(iG, g, lg2D)
        n = g.nx * yTop * g.nz
        a = np.array([float(t) for t in tok[p:p + per * n]]).reshape(g.nx, yTop, g.nz, per); p += per * n
        act = np.asarray(g.active)[:, :yTop, :]
        m = act > 0
        for q in range(3):
            v = np.zeros(g.nCells + 1, F32)
            v[act[m]] = a[..., q][m]
            out[q].append(v)
        if multi_chemistry:
            ab = np.ones((g.nx, g.ny, g.nz), I32, order="F")
            ab[:, :yTop, :] = a[..., 3].astype(I32)
            out[3].append(ab)
        else:
            out[3].append(None)
    return out


def write_grid2(path: str, model: Model, ionDen: Sequence[np.ndarray], lgElementOn, elementXref, nstages: int,
                lg2D: bool = False) -> None:
    """grid2.out, grid_mod.f90:2739-2745: per cell, one record per switched-on element with
    ``ionDen(cell, elementXref(elem), 1:min(elem+1, nstages))``."""
    on = [e for e in range(1, 31) if lgElementOn[e - 1]]
    with open(path, "w") as fh:
        for iG, g in enumerate(model.grids):
            act = np.asarray(g.active)
            yTop = 1 if (iG == 0 and lg2D) else g.ny
            lines = []
            for i in range(g.nx):
                for j in range(yTop):
                    for k in range(g.nz):
                        c = max(int(act[i, j, k]), 0)
                        for e in on:
                            row = ionDen[iG][c, int(elementXref[e - 1]) - 1, :min(e + 1, nstages)]
                            lines.append(" " + " ".join(_r(v) for v in row) + "\n")
            fh.writelines(lines)


def read_grid2(path: str, grids: Sequence[Grid], lgElementOn, elementXref, nstages: int, lg2D: bool = False):
    """-> ionDen per grid, (0:nCells, nElementsUsed, nstages) F-order (resetGrid, :3440-3452)."""
    tok = open(path).read().split()
    p = 0
    on = [e for e in range(1, 31) if lgElementOn[e - 1]]
    out = []
    for iG, g in enumerate(grids):
        d = np.zeros((g.nCells + 1, len(on), nstages), dtype=F32, order="F")
        act = np.asarray(g.active)
        for i in range(g.nx):
            for j in range(_ytop(iG, g, lg2D)):
                for k in range(g.nz):
                    c = int(act[i, j, k])
                    for e in on:
                        n = min(e + 1, nstages)
                        if c > 0:
                            d[c, int(elementXref[e - 1]) - 1, :n] = [float(t) for t in tok[p:p + n]]
                        p += n
        out.append(d)
    return out


@dataclass
class RunParams:
    """The run parameters writeGrid stores in grid3.out (grid_mod.f90:2812-2866) that are not part
    of :class:`Model`; defaults are those of set_input_mod.f90."""

    convWriteGrid: float = 0.0
    lgAutoPackets: bool = False
    convIncPercent: float = 0.0
    nPhotIncrease: float = 0.0
    maxPhotons: int = 0
    lgTalk: bool = False
    lg1D: bool = False
    nuStepSize: float = 0.075
    nuMax: float = 15.0
    nuMin: float = 1.001e-5
    R_in: float = -1.0
    XHIlimit: float = 0.05
    maxIterateMC: int = 30
    minConvergence: float = 95.0
    nAbComponents: int = 1
    abundanceFile: Sequence[str] = ("none",)
    lgOutput: bool = False
    dxSlit: float = 0.0
    dySlit: float = 0.0
    lgDustConstant: bool = False
    nDustComponents: int = 1
    dustSpeciesFile: Sequence[str] = ("none",)
    dustFile2: str = "none"
    lgRecombination: bool = False
    nSpecies: int = 1
    resLinesTransfer: float = 101.0
    lgDustScattering: bool = True
    contCube: Sequence[float] = (-1.0, -1.0)
    lgPhotoelectric: bool = True
    lgTraceHeating: bool = False
    Ldiffuse: float = 0.0
    Tdiffuse: float = 0.0
    shapeDiffuse: str = "none"
    nPhotonsDiffuse: int = 0
    emittingGrid: int = 0
    nstages: int = 7
    lg2D: bool = False
    lgEcho: bool = False
    echot1: float = 0.0
    echot2: float = 0.0
    echoTemp: float = 0.0
    lgNosource: bool = False


def grid3_records(model: Model, rp: RunParams) -> list:
    """The records of grid3.out in writeGrid's order (:2812-2866), as tuples of values."""
    rec = [(model.nGrids,), (F32(rp.convWriteGrid), " convWriteGrid"),
           (rp.lgAutoPackets, F32(rp.convIncPercent), F32(rp.nPhotIncrease), int(rp.maxPhotons), " lgAutoPackets"),
           (bool(model.lgSymmetricXYZ), " lgSymmetricXYZ"), (rp.lgTalk, " lgTalk"), (rp.lg1D, " lg1D"),
           (int(model.nbins), " nbins"), (F32(rp.nuStepSize), " nuStepSize"), (F32(rp.nuMax), " nuMax"),
           (F32(rp.nuMin), " nuMin"), (F32(rp.R_in), " R_in"), (F32(rp.XHIlimit), " XHIlimit"),
           (int(rp.maxIterateMC), F32(rp.minConvergence), " maxIterateMC"), (bool(model.lgDebug), " lgDebug"),
           (bool(model.lgPlaneIonization), " lgPlaneIonization"), (int(rp.nAbComponents), " nAbComponents")]
    rec += [('"', f, '"') for f in list(rp.abundanceFile)[:rp.nAbComponents]]
    rec += [(rp.lgOutput, " lgOutput"), (F32(rp.dxSlit), F32(rp.dySlit), " dxSlit,dySlit"),
            (bool(model.lgDust), rp.lgDustConstant, " lgDust, lgDustConstant"),
            (bool(model.lgMultiDustChemistry), int(rp.nDustComponents), " lgMultiDustChemistry, nDustComponents")]
    if model.lgDust:
        rec += [('"', f.rstrip(), '"', " dustFile") for f in list(rp.dustSpeciesFile)[:rp.nDustComponents]]
    else:
        rec += [("none", " dustFile")]
    rec += [('"', rp.dustFile2.rstrip(), '"', " dustFile"), (bool(model.lgGas), " lgGas"),
            (rp.lgRecombination, " lgRecombination")]
    if model.lgDust:
        rec += [(int(model.nSpeciesMax), int(rp.nSpecies), int(model.nSizes), " nSpeciesMax, nSpecies, nSizes"),
                tuple(int(v) for v in np.asarray(model.nSpeciesPart)[:rp.nDustComponents]) + (" Partial nspecies",)]
    else:
        rec += [(1, 1, 1, " nSpeciesMax, nSpecies, nSizes"), (1, " Partial nspecies")]
    rec += [(F32(rp.resLinesTransfer), "resLinesTransfer"), (rp.lgDustScattering, "lgDustScattering"),
            (int(model.nAngleBins), " nAngleBins")]
    if model.nAngleBins > 0:        # the whole arrays (0:nAngleBins) are written (:2851-2852), 1: is read back
        rec += [tuple(F32(v) for v in np.asarray(model.viewPointTheta, F32)) + (" inclination theta",),
                tuple(F32(v) for v in np.asarray(model.viewPointPhi, F32)) + (" inclination theta",)]
    rec += [(F32(rp.contCube[0]), F32(rp.contCube[1]), " continuumCube"), (rp.lgPhotoelectric, " lgPhotoelectric"),
            (rp.lgTraceHeating, " lgTraceHeating"), (F32(rp.Ldiffuse), " Ldiffuse"), (F32(rp.Tdiffuse), " Tdiffuse"),
            (rp.shapeDiffuse, " shapeDiffuse"), (int(rp.nPhotonsDiffuse), "nPhotonsDiffuse"),
            (int(rp.emittingGrid), " emittingGrid"), (int(rp.nstages), " emittingGrid"),
            (bool(model.lgMultistars), " lgMultiStars"), (rp.lg2D, " 2D geometry?"),
            (rp.lgEcho, F32(rp.echot1), F32(rp.echot2), F32(rp.echoTemp), " Echo on/off"), (rp.lgNosource, " NoSourceSED")]
    return rec


def _fmt(v) -> str:
    if isinstance(v, (bool, np.bool_)):
        return _l(v)
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, str):
        return v
    return _r(v)


def _join(rec) -> str:
    """list-directed output: values are blank separated, except that adjacent character items
    follow each other directly (so  '"', name, '"'  reads back as one quoted string)"""
    out = ""
    prev_str = True
    for v in rec:
        is_str = isinstance(v, str)
        out += ("" if (is_str and prev_str) or not out else " ") + _fmt(v)
        prev_str = is_str
    return out


def write_grid3(path: str, model: Model, rp: Optional[RunParams] = None) -> None:
    """grid3.out, grid_mod.f90:2805-2868."""
    with open(path, "w") as fh:
        for rec in grid3_records(model, rp or RunParams()):
            fh.write(" " + _join(rec) + "\n")


def read_grid3(path: str) -> dict:
    """The values resetGrid reads back (:3038-3111), keyed by name."""
    raw = iter(open(path).read().splitlines())
    last = [""]

    def nxt():
        last[0] = next(raw)
        return last[0].split()

    def L(t):
        return t.upper().strip(".") in ("T", "TRUE")

    def q(t):                                   # a record holding one quoted string: "name"
        ln = last[0]
        a = ln.index('"')
        return ln[a + 1:ln.index('"', a + 1)].strip()
    d = {}
    d["nGrids"] = int(nxt()[0]); d["convWriteGrid"] = float(nxt()[0])
    t = nxt(); d["lgAutoPackets"], d["convIncPercent"], d["nPhotIncrease"], d["maxPhotons"] = L(t[0]), float(t[1]), float(t[2]), int(t[3])
    d["lgSymmetricXYZ"] = L(nxt()[0]); d["lgTalk"] = L(nxt()[0]); d["lg1D"] = L(nxt()[0])
    d["nbins"] = int(nxt()[0]); d["nuStepSize"] = float(nxt()[0]); d["nuMax"] = float(nxt()[0]); d["nuMin"] = float(nxt()[0])
    d["R_in"] = float(nxt()[0]); d["XHIlimit"] = float(nxt()[0])
    t = nxt(); d["maxIterateMC"], d["minConvergence"] = int(t[0]), float(t[1])
    d["lgDebug"] = L(nxt()[0]); d["lgPlaneIonization"] = L(nxt()[0]); d["nAbComponents"] = int(nxt()[0])
    d["abundanceFile"] = [q(nxt()) for _ in range(d["nAbComponents"])]
    d["lgOutput"] = L(nxt()[0])
    t = nxt(); d["dxSlit"], d["dySlit"] = float(t[0]), float(t[1])
    t = nxt(); d["lgDust"], d["lgDustConstant"] = L(t[0]), L(t[1])
    t = nxt(); d["lgMultiDustChemistry"], d["nDustComponents"] = L(t[0]), int(t[1])
    d["dustSpeciesFile"] = [q(nxt()) for _ in range(d["nDustComponents"])] if d["lgDust"] else [nxt()[0]]
    d["dustFile2"] = q(nxt()); d["lgGas"] = L(nxt()[0]); d["lgRecombination"] = L(nxt()[0])
    t = nxt(); d["nSpeciesMax"], d["nSpecies"], d["nSizes"] = int(t[0]), int(t[1]), int(t[2])
    d["nSpeciesPart"] = [int(v) for v in nxt()[:d["nDustComponents"]]]
    d["resLinesTransfer"] = float(nxt()[0]); d["lgDustScattering"] = L(nxt()[0]); d["nAngleBins"] = int(nxt()[0])
    if d["nAngleBins"] > 0:
        n = d["nAngleBins"] + 1
        d["viewPointTheta"] = [float(v) for v in nxt()[:n]]; d["viewPointPhi"] = [float(v) for v in nxt()[:n]]
    t = nxt(); d["contCube"] = (float(t[0]), float(t[1]))
    d["lgPhotoelectric"] = L(nxt()[0]); d["lgTraceHeating"] = L(nxt()[0]); d["Ldiffuse"] = float(nxt()[0])
    d["Tdiffuse"] = float(nxt()[0]); d["shapeDiffuse"] = nxt()[0]; d["nPhotonsDiffuse"] = int(nxt()[0])
    d["emittingGrid"] = int(nxt()[0]); d["nstages"] = int(nxt()[0]); d["lgMultistars"] = L(nxt()[0]); d["lg2D"] = L(nxt()[0])
    t = nxt(); d["lgEcho"], d["echot1"], d["echot2"], d["echoTemp"] = L(t[0]), float(t[1]), float(t[2]), float(t[3])
    d["lgNosource"] = L(nxt()[0])
    return d


def write_checkpoint(outdir: str, model: Model, lgConverged=None, **kw) -> None:
    os.makedirs(outdir, exist_ok=True)
    write_grid0(os.path.join(outdir, "grid0.out"), model, lgConverged=lgConverged)
    if model.lgDust:
        write_dust_grid(os.path.join(outdir, "dustGrid.out"), model, **kw)
