"""Checkpoint wire formats of the reference (writeGrid, grid_mod.f90:2646-2870; reader resetGrid,
:2967-3567) for the arrays this repository owns: ``grid0.out`` (geometry, active map, convergence
flags), ``dustGrid.out`` (Ndust, dustAbunIndex, Tdust) and ``photoSource.out``.  Lets the harness
hand a dust-only state over to (or warm-start from) a real mocassin run (``mocassinWarm``).

The files are Fortran list-directed text: one record per line, blank separated, read back with
``read(unit,*)`` -- so any whitespace layout round-trips; numbers are written with 9 significant
digits (float32 round-trip exact).  ``grid1.out`` / ``grid2.out`` (Te, Ne, ionDen) belong to the
ionisation solver, which stays with the host code, and are not written here."""
from __future__ import annotations

import os
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
            fh.write(f" {g.nx} {g.ny} {g.nz} {g.nCells} {g.motherP} {_r(model.R_out)}\n")
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


def read_grid0(path: str):
    """Inverse of write_grid0 -> (grids [axes, active, nCells, motherP], R_out, lgConverged, lgBlack)."""
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
        vals = np.array([int(t) for t in tok[p:p + 3 * nx * ny * nz]], dtype=np.int64).reshape(nx, ny, nz, 3); p += 3 * nx * ny * nz
        active = np.asfortranarray(vals[..., 0].astype(I32))
        conv = np.zeros(nCells + 1, I32); black = np.zeros(nCells + 1, I32)
        m = active > 0
        conv[active[m]] = vals[..., 1][m]; black[active[m]] = vals[..., 2][m]
        grids.append(Grid(xAxis=ax[0], yAxis=ax[1], zAxis=ax[2], active=active, nCells=nCells, motherP=motherP))
        convs.append(conv); blacks.append(black)
        if len(grids) == nGrids:
            break
    return grids, R_out, convs, blacks


def write_dust_grid(path: str, model: Model, lgMultiChemistry: bool = False, totalDustMass: float = 0.0) -> None:
    """dustGrid.out, grid_mod.f90:2746-2757,2771-2775: per cell of the x,y,z loop ``Ndust``
    (``Ndust dustAbunIndex`` with lgMultiChemistry -- sic, the gas flag) and nSizes+1 lines of
    ``Tdust(0:nSpeciesMax, ai, cell)``; cell 0 stands in for inactive cells."""
    with open(path, "w") as fh:
        for g in model.grids:
            act = np.asarray(g.active)
            T = np.asarray(g.Tdust, dtype=F32)
            lines = []
            for i in range(g.nx):
                for j in range(g.ny):
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
        fh.write(f" Total dust mass [1.e45 g]:  {_r(totalDustMass)}\n")
        fh.write(f" Total dust mass [Msol]:  {_r(totalDustMass * 5.028e11)}\n")


def read_dust_grid(path: str, model: Model, lgMultiChemistry: bool = False) -> None:
    """Fill Ndust, dustAbunIndex and Tdust of model.grids from dustGrid.out (resetGrid's dust part)."""
    tok = open(path).read().split()
    p = 0
    nS, nZ = model.nSpeciesMax + 1, model.nSizes + 1
    per = (2 if lgMultiChemistry else 1) + nS * nZ
    for g in model.grids:
        n = g.nx * g.ny * g.nz
        block = tok[p:p + per * n]; p += per * n
        a = np.array([float(t) for t in block], dtype=np.float64).reshape(g.nx, g.ny, g.nz, per)
        act = np.asarray(g.active)
        m = act > 0
        g.Ndust = np.zeros(g.nCells + 1, F32)
        g.Ndust[act[m]] = a[..., 0][m]
        off = 1
        if lgMultiChemistry:
            g.dustAbunIndex = np.ones(g.nCells + 1, I32)
            g.dustAbunIndex[act[m]] = a[..., 1][m].astype(I32)
            off = 2
        T = np.zeros((nS, nZ, g.nCells + 1), dtype=F32, order="F")
        body = a[..., off:].reshape(g.nx, g.ny, g.nz, nZ, nS)       # lines: ai, values: species
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


def write_checkpoint(outdir: str, model: Model, lgConverged=None, **kw) -> None:
    os.makedirs(outdir, exist_ok=True)
    write_grid0(os.path.join(outdir, "grid0.out"), model, lgConverged=lgConverged)
    if model.lgDust:
        write_dust_grid(os.path.join(outdir, "dustGrid.out"), model, **kw)
