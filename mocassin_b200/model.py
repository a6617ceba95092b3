"""Host-side description of the data the transport hot path reads and writes.

These classes mirror, member for member, the parts of the reference's shared
state that `photon_mod` touches:

* :class:`Grid`  -- the hot members of ``type grid_type``
  (reference ``source/common_mod.f90:241-302``);
* :class:`Model` -- the module globals of ``common_mod`` used by
  ``energyPacketDriver`` (``source/photon_mod.f90:26-2974``): flags, frequency
  mesh, stellar CDFs, viewing-angle tables, dust species tables.

All arrays keep the reference's Fortran layout (column major, cell index fastest in
``T(0:nCells, 1:nbins)``) so that a Fortran host can hand its own arrays to the C ABI
without any repacking; numpy arrays are therefore created with ``order='F'``.
Indices stored *inside* arrays (``active``, ``starIndeces``) are 1-based as in the
reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

F32 = np.float32
I32 = np.int32


def farray(shape, dtype=F32):
    return np.zeros(shape, dtype=dtype, order="F")


@dataclass
class Grid:
    """Hot members of ``grid_type`` (common_mod.f90:241-302)."""

    xAxis: np.ndarray
    yAxis: np.ndarray
    zAxis: np.ndarray
    active: np.ndarray                 # int32 (nx,ny,nz) F-order; >0 cell id, 0 inactive, <0 -subgrid
    nCells: int
    motherP: int = 0                   # 0 for the mother grid, 1 for sub-grids
    opacity: Optional[np.ndarray] = None      # (nCells+1, nbins) F
    scaOpac: Optional[np.ndarray] = None
    absOpac: Optional[np.ndarray] = None
    recPDF: Optional[np.ndarray] = None
    dustPDF: Optional[np.ndarray] = None
    linePDF: Optional[np.ndarray] = None      # (nCells+1, nLines) F   (debug only)
    totalLines: Optional[np.ndarray] = None   # (nCells+1,)
    Tdust: Optional[np.ndarray] = None        # (nSpeciesMax+1, nSizes+1, nCells+1) F
    dustAbunIndex: Optional[np.ndarray] = None  # (nCells+1,) int32
    Ndust: Optional[np.ndarray] = None        # (nCells+1,)
    Hden: Optional[np.ndarray] = None         # (nCells+1,)
    resLinePackets: Optional[np.ndarray] = None   # (nCells+1,) int32, resonance-line transfer

    @property
    def nx(self):
        return int(self.xAxis.shape[0])

    @property
    def ny(self):
        return int(self.yAxis.shape[0])

    @property
    def nz(self):
        return int(self.zAxis.shape[0])

    @property
    def geoCorr(self):
        """grid_mod.f90:809-812."""
        gx = F32((self.xAxis[-1] - self.xAxis[-2]) / F32(2.0))
        gy = F32((self.yAxis[-1] - self.yAxis[-2]) / F32(2.0))
        gz = F32((self.zAxis[-1] - self.zAxis[-2]) / F32(2.0))
        return gx, gy, gz

    def min_cell_width(self) -> float:
        w = min(float(np.min(np.diff(a.astype(np.float64)))) for a in (self.xAxis, self.yAxis, self.zAxis))
        return w / 2.0  # the symmetric first cell is half a spacing wide

    def max_cell_width(self) -> float:
        w = max(float(np.max(np.diff(a.astype(np.float64)))) for a in (self.xAxis, self.yAxis, self.zAxis))
        return w / 2.0

    def cell_volumes(self, symmetric: bool) -> np.ndarray:
        """dV(0:nCells) in 1e45 cm^3, float32 arithmetic of photon_mod.f90:1469-1512
        (== getVolume, grid_mod.f90:2876-2965). Entry 0 is unused."""

        def widths(a):
            n = a.shape[0]
            w = np.empty(n, dtype=F32)
            w[1:-1] = np.abs(a[2:] - a[:-2]) / F32(2.0)
            w[0] = np.abs(a[1] - a[0]) / F32(2.0) if symmetric else np.abs(a[1] - a[0])
            w[-1] = np.abs(a[-1] - a[-2])
            return (w / F32(1.0e15)).astype(F32)

        dx, dy, dz = widths(self.xAxis), widths(self.yAxis), widths(self.zAxis)
        vol = ((dx[:, None, None] * dy[None, :, None]).astype(F32) * dz[None, None, :]).astype(F32)
        out = np.zeros(self.nCells + 1, dtype=F32)
        m = self.active > 0
        out[self.active[m]] = vol[m]
        return out


def number_active(mask: np.ndarray) -> tuple[np.ndarray, int]:
    """Active-cell numbering of setMotherGrid (grid_mod.f90:1226-1294): cells are
    numbered in loop order i (x) outermost, k (z) innermost."""
    nx, ny, nz = mask.shape
    active = np.zeros((nx, ny, nz), dtype=I32, order="F")
    ids = np.arange(1, int(mask.sum()) + 1, dtype=I32)
    # C-order ravel of (x,y,z) has z fastest == reference loop order
    flat = np.zeros(nx * ny * nz, dtype=I32)
    flat[np.flatnonzero(mask.ravel(order="C"))] = ids
    active[...] = flat.reshape((nx, ny, nz), order="C")
    return active, int(ids.shape[0])


def auto_axis(n: int, R: float, symmetric: bool) -> np.ndarray:
    """fillGrid automatic axes (grid_mod.f90:536-599), float32 arithmetic."""
    i = np.arange(n, dtype=F32)
    if symmetric:
        a = (i / F32(n - 1)).astype(F32)
    else:
        a = (F32(2.0) * i / F32(n - 1) - F32(1.0)).astype(F32)
    return (a * F32(R)).astype(F32)


@dataclass
class Model:
    """Module globals of common_mod read by the hot path."""

    grids: List[Grid]
    nbins: int
    nuArray: np.ndarray                     # (nbins,) Ryd
    inSpectrumProbDen: np.ndarray           # (nStars+1, nbins) C-order rows; row 0 = diffuse source
    deltaE: np.ndarray                      # (nStars+1,)  [1e36 erg/s]
    starPosition: np.ndarray                # (nStars,3) cm
    starIndeces: np.ndarray                 # (nStars,4) int32: xP,yP,zP,grid (1-based)
    lgDust: bool = False
    lgGas: bool = True
    lgSymmetricXYZ: bool = False
    lgIsotropic: bool = False
    lgPlaneIonization: bool = False
    lgDebug: bool = False
    lgMultistars: bool = False
    lgMultiDustChemistry: bool = False
    R_out: float = 0.0
    ionEdge1: float = 0.99946               # ionEdge(1) = H I threshold [Ryd]
    gSca: Optional[np.ndarray] = None       # (nbins,)
    nAngleBins: int = 0
    totAngleBinsTheta: int = 10             # common_mod.f90:399
    totAngleBinsPhi: int = 20               # common_mod.f90:400
    viewPointTheta: Optional[np.ndarray] = None   # (nAngleBins+1,) radians, [0] = 0
    viewPointPhi: Optional[np.ndarray] = None
    nLines: int = 0
    # dust species tables
    nSpeciesMax: int = 0
    nSizes: int = 0
    nSpeciesPart: np.ndarray = field(default_factory=lambda: np.ones(1, dtype=I32))
    grainAbun: np.ndarray = field(default_factory=lambda: np.ones((1, 1), dtype=F32, order="F"))
    dustComPoint: np.ndarray = field(default_factory=lambda: np.ones(1, dtype=I32))
    TdustSublime: np.ndarray = field(default_factory=lambda: np.full(1, 1.0e30, dtype=F32))

    @property
    def nStars(self) -> int:
        return int(self.starPosition.shape[0])

    @property
    def nGrids(self) -> int:
        return len(self.grids)

    # -- viewing-angle tables, initCartesianGrid (grid_mod.f90:416-468) ------------
    def angle_tables(self):
        PI = F32(3.141592654)
        totT = int(self.totAngleBinsTheta)
        totP = int(self.totAngleBinsPhi)
        vtheta = np.zeros(self.nAngleBins + 1, dtype=F32)
        vphi = np.zeros(self.nAngleBins + 1, dtype=F32)
        if self.nAngleBins > 0:
            vtheta[1:] = np.asarray(self.viewPointTheta, dtype=F32)[1:]
            vphi[1:] = np.asarray(self.viewPointPhi, dtype=F32)[1:]
            if np.any(vphi[1:] < 0):
                totP = 1
                vphi[:] = F32(-1.0)
        dTheta = F32(PI / F32(totT))
        dPhi = F32(F32(2.0) * PI / F32(totP))
        pT = np.zeros(totT + 1, dtype=I32)
        pP = np.zeros(totP + 1, dtype=I32)
        for i in range(1, self.nAngleBins + 1):
            pT[int(vtheta[i] / dTheta) + 1] = i
            pP[int(vphi[i] / dPhi) + 1] = i
        return dict(dTheta=dTheta, dPhi=dPhi, totAngleBinsTheta=totT, totAngleBinsPhi=totP,
                    viewPointPtheta=pT, viewPointPphi=pP, viewPointTheta=vtheta, viewPointPhi=vphi)

    def len_unit_exponent(self, g: Grid) -> int:
        """Exponent e of the power-of-two path-length unit 2^e [cm] of grid g's
        fixed-point J tally: smallest cell width / 2^24, rounded down to a power of 2 -- but not
        below the largest cell width / 2^33 (strongly graded axes: head room of the 64-bit sums).
        The same rule as mcb200_set_grid (capi.cu), on the same float32 axis values."""
        e_fine = int(np.floor(np.log2(g.min_cell_width()))) - 24
        e_cap = int(np.floor(np.log2(g.max_cell_width()))) - 33
        return max(e_fine, e_cap)


def locate(xa: np.ndarray, x: float) -> int:
    """interpolation_mod.f90:48-81 for ascending xa; returns the 1-based ns."""
    n = xa.shape[0]
    x = F32(x)
    if x > xa[-1]:
        return n
    if x < xa[0]:
        return 0
    pos = np.flatnonzero(xa > x)
    if pos.size == 0:
        return 1
    return max(int(pos[0]), 1)  # (first 1-based index with xa>x) - 1


def star_indices(grid: Grid, pos) -> list[int]:
    """setStarPosition for a star in the mother grid (grid_mod.f90:3569-3610)."""
    out = []
    for a, p in zip((grid.xAxis, grid.yAxis, grid.zAxis), pos):
        ns = locate(a, p)
        if ns < a.shape[0] and ns >= 1:
            if F32(p) > (a[ns - 1] + a[ns]) / F32(2.0):
                ns += 1
        out.append(max(ns, 1))
    return out


def set_star_position(grids, relative):
    """setStarPosition (grid_mod.f90:3569-3648): the `starPosition` keyword gives coordinates in
    units of the mother grid's last axis points; returns (positions [cm] float32 (nStars, 3),
    starIndeces int32 (nStars, 4)).  A star whose mother cell points into a sub-grid
    (active < 0) is located in that sub-grid -- with the reference's yAxis in the z mid-point test,
    and with its other quirk: nxA/nyA/nzA are overwritten by the sub-grid's extents and not
    restored, so every LATER star is scaled by, and searched up to, mother-axis point number
    nx(sub-grid) instead of the last one (:3581-3583, :3607-3609)."""
    g1 = grids[0]
    pos = np.zeros((len(relative), 3), dtype=F32)
    idx = np.zeros((len(relative), 4), dtype=I32)
    nA = [g1.nx, g1.ny, g1.nz]

    def find(a, v, n, mid=None):
        mid = a if mid is None else mid
        p = locate(a, v)
        if p < n:
            # p = 0 (below the axis) indexes xA(0) in the reference; only reachable off-grid
            if p >= 1 and F32(v) > F32(F32(mid[p - 1] + mid[p]) / F32(2.0)):
                p += 1
        return p
    for i, rel in enumerate(relative):
        axes = (g1.xAxis, g1.yAxis, g1.zAxis)
        pos[i] = [F32(F32(rel[k]) * axes[k][nA[k] - 1]) for k in range(3)]
        x, y, z = (find(axes[k], pos[i, k], nA[k]) for k in range(3))
        a = int(g1.active[x - 1, y - 1, z - 1])
        if a >= 0:
            idx[i] = [x, y, z, 1]
        else:
            s = grids[-a - 1]
            nA = [s.nx, s.ny, s.nz]
            idx[i] = [find(s.xAxis, pos[i, 0], nA[0]), find(s.yAxis, pos[i, 1], nA[1]),
                      find(s.zAxis, pos[i, 2], nA[2], mid=s.yAxis), -a]
    return pos, idx


def mask_subgrids(grids, symmetric: bool) -> None:
    """fillGrid's last block (grid_mod.f90:835-889), in place: every cell of grid iG whose centre
    lies strictly inside the box of another grid jG >= 2 gets active = -jG (for symmetricXYZ a
    box starting at 0 includes the centres at 0).  Runs after the active cells were numbered, so
    the masked mother cells keep their (now unused) numbers, as in the reference."""
    if len(grids) < 2:
        return
    for iG, g in enumerate(grids, start=1):
        for jG in range(2, len(grids) + 1):
            if jG == iG:
                continue
            s = grids[jG - 1]
            sel = []
            for a, b in ((g.xAxis, s.xAxis), (g.yAxis, s.yAxis), (g.zAxis, s.zAxis)):
                lo = (a > b[0]) | ((a >= b[0]) & (b[0] == 0)) if symmetric else (a > b[0])
                sel.append(lo & (a < b[-1]))
            box = sel[0][:, None, None] & sel[1][None, :, None] & sel[2][None, None, :]
            g.active[box] = -jG
