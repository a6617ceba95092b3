"""The shipped multi-grid examples (examples/multigridgas, examples/multigridgasdust) as transport
workloads in their first-iteration state: `input.in` with `densityFile`, `multiGrids`, and for the
second deck `Ndust file` + `dustFile`.  Host-side Python restating, in the reference's float32
operation order,

=======================================================  ======================================
reference                                                here
=======================================================  ======================================
density file of the mother grid: axes from its rows,      :func:`read_density_file`
`setMotherGrid` grid_mod.f90:956-1031,1226-1294
gas+dust frequency mesh `initCartesianGrid` :262-331     :func:`nu_mesh_gasdust`
sub-grid list (`readGridList` set_input_mod.f90,           :func:`read_grid_list`
`setSubGrids` grid_mod.f90:1848-1857)
sub-grid density file, normalised coordinates rescaled     :func:`read_subgrid_file`
to the box, active cells, `denfac` (:2072-2236,:2398-2413)
cross-section stack: gas part + `makeDustXsec` appended    :func:`load_multigrid_deck`
at xSecTop (ph_mod.f90:387,1492-1507)
=======================================================  ======================================

List-directed input is read the way Fortran reads it (:class:`ListReader`): a READ starts on a new
record, runs on over the following records until it has all its items, and drops what is left of
the last record it touched.  This matters: examples/multigridgasdust ships a 4-column sub-grid file
(`x y z Hden`) although a gas+dust run reads FIVE items per cell (`x, y, z, Hden, Ndust`,
grid_mod.f90:2103) -- the reference takes the fifth from the next record and loses the rest of that
record, so every second row is skipped and the coordinates stop lining up with the loop indices;
its own sanity check then stops the run ("setSubGrids: insanity occurred in setting yAxis",
:2128-2133; executed and confirmed by running the reference's reading code on the shipped files,
tests/golden/ref_aux_subgrid_multigridgasdust.npz).  The loader reproduces that
(:class:`DeckError`) unless `pad_missing_ndust` asks for the evident intent, Ndust = 0 inside the
sub-grid.  examples/multigridgas (gas only, four items per cell) loads exactly as shipped.
"""
from __future__ import annotations

import os
import shlex

import numpy as np

from . import gasdata as G
from .deck import (make_dust_xsec, read_grain_sizes, read_grain_species, read_input, stellar_cdf, _real)
from .model import F32, I32, Grid, Model, mask_subgrids, number_active, set_star_position
from .workloads import recombination_cdf, wid_flx


class DeckError(RuntimeError):
    """A condition on which the reference stops (`print*; stop`, or a Fortran run-time error)."""


class ListReader:
    """List-directed READs on one file, record by record."""

    def __init__(self, path):
        self.path = path
        self.rows = []
        with open(path) as fh:
            for raw in fh:
                lex = shlex.shlex(raw, posix=True)
                lex.whitespace += ","
                lex.whitespace_split = True
                lex.commenters = ""
                self.rows.append(list(lex))
        while self.rows and not self.rows[-1]:
            self.rows.pop()
        self.pos = 0

    def read(self, n: int):
        out = []
        col = 0
        while len(out) < n:
            if self.pos >= len(self.rows):
                raise DeckError(f"{self.path}: end of file in a list-directed READ of {n} items")
            row = self.rows[self.pos]
            if col < len(row):
                out.append(row[col])
                col += 1
            else:
                self.pos += 1
                col = 0
        self.pos += 1
        return out

    def backspace(self):
        if self.pos > 0:
            self.pos -= 1


def _r(tok):
    return F32(float(tok.lower().replace("d", "e")))


def _radius(x, y, z):
    t = F32(1.0e10)
    return F32(t * np.sqrt(F32(F32(F32(x / t) * F32(x / t)) + F32(F32(y / t) * F32(y / t))) + F32(F32(z / t) * F32(z / t))))


def read_density_file(path: str, nx: int, ny: int, nz: int, ncol: int = 4):
    """Rows `x y z value` of a mother-grid density / Ndust file (grid_mod.f90:956-963,1014,
    :1187-1218): an optional `#` header record, then nx*ny*nz READs of `ncol` items, x outermost and
    z innermost; every row overwrites xAxis(i), yAxis(j), zAxis(k)."""
    rd = ListReader(path)
    if rd.read(1)[0] != "#":
        rd.backspace()
    xA, yA, zA = np.zeros(nx, F32), np.zeros(ny, F32), np.zeros(nz, F32)
    val = np.zeros((nx, ny, nz), dtype=F32)
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                t = rd.read(ncol)
                xA[i], yA[j], zA[k], val[i, j, k] = _r(t[0]), _r(t[1]), _r(t[2]), _r(t[3])
    return xA, yA, zA, val


def nu_mesh_gasdust(nbins: int, nuMin: float, nuMax: float, ionEdge: np.ndarray, nEdges: int, nu_dust_file: str) -> np.ndarray:
    """The gas+dust frequency mesh (grid_mod.f90:262-331): series edges and thresholds as in the gas
    mesh, then points of dustData/nuDustRyd.dat -- read until `nuArray(i) >= seriesEdge(1)` where i
    counts the points READ, not the slot written (:306), so with nuMin above the 4.9 GHz point the
    loop stops after ONE dust point -- then the logarithmic fill and sortUp."""
    nu = np.zeros(nbins, dtype=F32)
    k = 0
    if F32(nuMin) < G.RADIO_4P9_GHZ:
        nu[0] = G.RADIO_4P9_GHZ
        k = 1
    nuMinA, nuMaxA = F32(nuMin), F32(nuMax)
    d3, d6 = F32(0.0003), F32(0.0006)
    for i in range(G.N_SERIES):
        nu[k], nu[k + 1], nu[k + 2] = G.SERIES_EDGE[i], F32(G.SERIES_EDGE[i] - d3), F32(G.SERIES_EDGE[i] + d3)
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
    rd = ListReader(nu_dust_file)
    for i in range(1, 10_000_001):
        if i > nbins:
            raise DeckError("initCartesianGrid: nbins is smaller than the number of frequency points in nuDustRyd.dat")
        try:
            nu[k] = _r(rd.read(1)[0])
            eof = False
        except DeckError:
            eof = True
        k += 1
        if eof:
            break
        if nu[i - 1] >= G.SERIES_EDGE[0]:
            break
    iCount = nbins - (k + 1) + 1
    step = F32(F32(np.log10(nuMaxA) - np.log10(nuMinA)) / F32(iCount - 1))
    nu[k] = nuMinA
    for i in range(k + 1, nbins):
        nu[i] = F32(np.power(F32(10.0), F32(np.log10(nu[i - 1]) + step)))
    return G.sort_up(nu)


def read_grid_list(path: str, nGrids: int):
    """The sub-grid list (grid_mod.f90:1848-1857): per sub-grid `motherP nx ny nz 'file' denfac`
    and `xmin xmax ymin ymax zmin zmax`, each a list-directed READ."""
    rd = ListReader(path)
    out = []
    for _ in range(2, nGrids + 1):
        t = rd.read(6)
        b = rd.read(6)
        out.append(dict(motherP=int(_r(t[0])), nx=int(_r(t[1])), ny=int(_r(t[2])), nz=int(_r(t[3])), file=t[4],
                        denfac=_r(t[5]), box=[_r(v) for v in b]))
    return out


def read_subgrid_file(path: str, spec: dict, lgDust: bool, R_in: float, R_out: float, pad_missing_ndust: bool = False):
    """One sub-grid (grid_mod.f90:2072-2236, gas or gas+dust, single chemistry): READ
    `x, y, z, Hden[, Ndust]` per cell with x, y, z normalised to the box, axes rescaled row by row
    (a row overwrites xAxis(ix) -- also the end points the rescaling itself uses), the active-cell
    rule of the mother grid.  Returns (Grid, HdenTemp, NdustTemp) before `denfac`."""
    nx, ny, nz = spec["nx"], spec["ny"], spec["nz"]
    xA, yA, zA = np.zeros(nx, F32), np.zeros(ny, F32), np.zeros(nz, F32)
    xA[0], xA[-1], yA[0], yA[-1], zA[0], zA[-1] = spec["box"]
    rd = ListReader(path)
    nItems = 5 if lgDust else 4
    if lgDust and pad_missing_ndust and all(len(r) == 4 for r in rd.rows):
        nItems = 4                                       # the shipped file: no Ndust column -> no dust in the sub-grid
    H = np.zeros((nx, ny, nz), dtype=F32)
    Nd = np.zeros((nx, ny, nz), dtype=F32)
    act = np.ones((nx, ny, nz), dtype=bool)
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                t = rd.read(nItems)
                x, y, z, H[ix, iy, iz] = _r(t[0]), _r(t[1]), _r(t[2]), _r(t[3])
                if nItems == 5:
                    Nd[ix, iy, iz] = _r(t[4])
                x = F32(xA[0] + F32(x * F32(xA[-1] - xA[0])))
                y = F32(yA[0] + F32(y * F32(yA[-1] - yA[0])))
                z = F32(zA[0] + F32(z * F32(zA[-1] - zA[0])))
                for a, i, n, v, name in ((xA, ix, nx, x, "x"), (yA, iy, ny, y, "y"), (zA, iz, nz, z, "z")):
                    if i == n - 1 and abs(F32(v - a[i])) >= abs(F32(a[i] - a[i - 1])):
                        raise DeckError(f"setSubGrids: insanity occurred in setting {name}Axis for a subGrid")
                xA[ix], yA[iy], zA[iz] = x, y, z
                rad = _radius(xA[ix], yA[iy], zA[iz])
                if rad < F32(R_in) or (R_out > 0 and rad > F32(R_out)):
                    act[ix, iy, iz] = False
                if not act[ix, iy, iz]:
                    H[ix, iy, iz] = 0
                    Nd[ix, iy, iz] = 0
                if not (H[ix, iy, iz] > 0 or (lgDust and Nd[ix, iy, iz] > 0)):
                    act[ix, iy, iz] = False
                    H[ix, iy, iz] = 0
                    Nd[ix, iy, iz] = 0
    active, nCells = number_active(act)
    g = Grid(xAxis=xA, yAxis=yA, zAxis=zA, active=active, nCells=nCells)
    g.motherP = spec["motherP"]
    return g, H, Nd


def _per_cell(active, field, nCells):
    out = np.zeros(nCells + 1, dtype=F32)
    m = active > 0
    out[active[m]] = field[m]
    return out


def load_multigrid_deck(run_dir: str, share_dir: str, input_file: str = "input.in", pad_missing_ndust: bool = False):
    """(Model, tables, Deck) of a `multiGrids` gas or gas+dust deck with a mother-grid density file.
    tables["grids"][iG-1] carries the per-grid K1 inputs (den, ionDen, and the dust dictionary)."""
    d = read_input(os.path.join(run_dir, input_file))
    o = d.other
    if not d.lgGas or "densityFile" not in o or "multiGrids" not in o:
        raise NotImplementedError("gas(+dust) decks with densityFile and multiGrids")
    if d.contShape != "blackbody":
        raise NotImplementedError(f"contShape {d.contShape}")
    nGrids, gridList = int(_real(o["multiGrids"][0])), o["multiGrids"][1]
    TeStart = F32(_real(o["TeStart"][0])) if "TeStart" in o else F32(10000.0)
    nstages = int(_real(o["nstages"][0])) if "nstages" in o else 7
    lgDust = bool(d.lgDust)

    def resolve(p):
        for cand in (os.path.join(run_dir, p), os.path.join(run_dir, os.path.basename(p))):
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError(p)

    ab, on, xref, nUsed = G.read_abundances(resolve(d.abundanceFile))
    ph1, ph2 = G.read_ph_tables(os.path.join(share_dir, "data", "ph1.dat"), os.path.join(share_dir, "data", "ph2.dat"))
    ionEdge, nEdges = G.ion_edges(ph1, on, nstages, d.nuMax)
    if lgDust:
        nu = nu_mesh_gasdust(d.nbins, d.nuMin, d.nuMax, ionEdge, nEdges, os.path.join(share_dir, "dustData", "nuDustRyd.dat"))
    else:
        nu = G.nu_mesh_gas(d.nbins, d.nuMin, d.nuMax, ionEdge, nEdges)
    nbins = d.nbins
    widFlx = wid_flx(nu)
    xt, ptr, xp = G.build_xsec_tables(nu, ph1, ph2, on, xref, nstages)
    dustT = None
    if lgDust:
        if d.dustFile is None:
            raise DeckError("readInput: dust present but no dustFile given")
        radius, weight = read_grain_sizes(resolve(d.dustFile[1]))
        species = read_grain_species(resolve(d.dustFile[0]))
        xs = make_dust_xsec(species, radius, weight, nu, share_dir, d.lgDustScattering)
        top = int(xp["xSecTop"])                               # makeDustXsec appends at xSecTop (ph_mod.f90:387)
        xt.xSecArray = np.concatenate([xt.xSecArray, xs["xSecArray"]]).astype(F32)
        nSp, nSz = len(species), radius.shape[0]
        dustT = dict(grainWeight=weight, grainRadius=radius, grainAbun1=xs["grainAbun"], TdustSublime=xs["TdustSublime"],
                     dustScaXsecP=np.asfortranarray(xs["dustScaXsecP"][1:, :] + top),
                     dustAbsXsecP=np.asfortranarray(xs["dustAbsXsecP"][1:, :] + top), gSca=xs["gSca"], nSp=nSp, nSz=nSz)

    # ---- mother grid: axes and densities from the files
    nx, ny, nz = d.nx, d.ny, d.nz
    xA, yA, zA, H3 = read_density_file(resolve(o["densityFile"][0]), nx, ny, nz)
    Nd3 = np.zeros_like(H3)
    if lgDust:
        if d.NdustFile == "none":
            raise NotImplementedError("gas+dust multigrid deck without an Ndust file")
        xA, yA, zA, Nd3 = read_density_file(resolve(d.NdustFile), nx, ny, nz)    # setMotherGrid reads the axes again (:1206)
    act = np.ones((nx, ny, nz), dtype=bool)
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                rad = _radius(xA[i], yA[j], zA[k])
                if rad < F32(d.R_in) or (d.R_out > 0 and rad > F32(d.R_out)):
                    act[i, j, k] = False
    H3 = np.where(act, H3, F32(0)).astype(F32)
    Nd3 = np.where(act, Nd3, F32(0)).astype(F32)
    act &= (H3 > 0) | (Nd3 > 0)
    H3 = np.where(act, H3, F32(0)).astype(F32)
    Nd3 = np.where(act, Nd3, F32(0)).astype(F32)
    active, nCells = number_active(act)
    mother = Grid(xAxis=xA, yAxis=yA, zAxis=zA, active=active, nCells=nCells)
    mother.motherP = 0
    grids, fields = [mother], [(H3, Nd3, F32(1.0))]
    for spec in read_grid_list(resolve(gridList), nGrids):
        g, H, Nd = read_subgrid_file(resolve(spec["file"]), spec, lgDust, d.R_in, d.R_out, pad_missing_ndust)
        grids.append(g)
        fields.append((H, Nd, spec["denfac"]))
    per_grid = []
    for g, (H, Nd, fac) in zip(grids, fields):
        g.Hden = (_per_cell(g.active, H, g.nCells) * F32(fac)).astype(F32)
        g.Te = np.zeros(g.nCells + 1, dtype=F32)
        g.Te[1:] = TeStart
        g.Ne = g.Hden.copy()
        ionDen = G.initial_ion_state(g.nCells, on, xref, nUsed, nstages)
        abIndex = np.ones(g.nCells + 1, dtype=I32)
        elemAbun = np.asfortranarray(ab.reshape(1, 30))
        entry = dict(ionDen=ionDen, den=xt.species_densities(ionDen, elemAbun, abIndex, g.Hden), abIndex=abIndex, dust=None)
        if lgDust:
            g.Ndust = (_per_cell(g.active, Nd, g.nCells) * F32(fac)).astype(F32)
            g.dustAbunIndex = np.ones(g.nCells + 1, dtype=I32)
            g.Tdust = np.full((dustT["nSp"] + 1, dustT["nSz"] + 1, g.nCells + 1), F32(50.0), dtype=F32, order="F")
            entry["dust"] = dict(Ndust=g.Ndust, Tdust=g.Tdust, dustAbunIndex=None, grainWeight=dustT["grainWeight"],
                                 dustScaXsecP=dustT["dustScaXsecP"], dustAbsXsecP=dustT["dustAbsXsecP"])
        # the re-emission tables are the host solver's (emissionDriver): smooth stand-in, see gasdeck.py
        row = recombination_cdf(nu, widFlx, float(TeStart)).astype(F32)
        g.recPDF = np.zeros((g.nCells + 1, nbins), dtype=F32, order="F")
        g.recPDF[1:, :] = row[None, :]
        g.totalLines = np.zeros(g.nCells + 1, dtype=F32)
        g.totalLines[1:] = F32(0.6)
        per_grid.append(entry)
    mask_subgrids(grids, d.lgSymmetricXYZ)
    pos, sidx = set_star_position(grids, [list(d.starPosition)])
    cdf = stellar_cdf(d.TStellar, nu, widFlx)
    LStar = F32(d.LStar)
    kw = {}
    if lgDust:
        grainAbun = np.zeros((1, dustT["nSp"]), dtype=F32, order="F")
        grainAbun[0, :] = dustT["grainAbun1"]
        kw = dict(gSca=dustT["gSca"], nSpeciesMax=dustT["nSp"], nSizes=dustT["nSz"], nSpeciesPart=np.asarray([dustT["nSp"]], I32),
                  grainAbun=grainAbun, dustComPoint=np.asarray([1], I32), TdustSublime=np.asarray(dustT["TdustSublime"], F32))
    model = Model(grids=grids, nbins=nbins, nuArray=nu,
                  inSpectrumProbDen=np.stack([np.zeros(nbins, F32), cdf]).astype(F32),
                  deltaE=np.asarray([0.0, F32(LStar / F32(d.nPhotons))], dtype=F32),
                  starPosition=np.asarray(pos, dtype=F32), starIndeces=np.asarray(sidx, dtype=I32),
                  lgDust=lgDust, lgGas=True, lgSymmetricXYZ=d.lgSymmetricXYZ, lgIsotropic=d.lgIsotropic,
                  R_out=float(d.R_out), ionEdge1=float(ionEdge[0]), **kw)
    tables = dict(xsec=xt, bands=xt.band_list(nbins), grids=per_grid, elemAbun=np.asfortranarray(ab.reshape(1, 30)),
                  widFlx=widFlx, nstages=nstages, lgElementOn=on, elementXref=xref, ptr=ptr, xp=xp, dust=dustT,
                  ph1=ph1, ph2=ph2, ionEdge=ionEdge[:nEdges], gridList=read_grid_list(resolve(gridList), nGrids),
                  recPDF_kind="stand-in (workloads.recombination_cdf at TeStart; emissionDriver is the host solver's)")
    return model, tables, d


# ---------------------------------------------------------------------------------------
# fixture format (tests/golden/deck_multigridgas.npz, deck_multigridgasdust.npz): the decks' files and
# the atomic / optical data do not travel to the GPU box
# ---------------------------------------------------------------------------------------
def multideck_to_arrays(model: Model, tables: dict, d) -> dict:
    import json

    from .gasdeck import _XT_SCALARS

    xt = tables["xsec"]
    scal = dict(lgSymmetricXYZ=bool(d.lgSymmetricXYZ), lgIsotropic=bool(d.lgIsotropic), lgDust=bool(model.lgDust),
                nPhotons=int(d.nPhotons), R_out=float(d.R_out), ionEdge1=float(model.ionEdge1), TeStart=float(model.grids[0].Te[1]),
                starPosition=list(d.starPosition), nGrids=len(model.grids), recPDF_kind=tables["recPDF_kind"],
                motherP=[int(getattr(g, "motherP", 0)) for g in model.grids],
                **{k: int(getattr(xt, k)) for k in _XT_SCALARS})
    out = dict(nuArray=model.nuArray, cdf=model.inSpectrumProbDen[1], deltaE=model.deltaE, xSecArray=xt.xSecArray,
               lgElementOn=xt.lgElementOn, elementXref=xt.elementXref, elementP=xt.elementP, nShells=xt.nShells,
               elemAbun=tables["elemAbun"], recRow=model.grids[0].recPDF[1], totalLines1=model.grids[0].totalLines[1:2],
               starPositionAbs=model.starPosition, starIndeces=model.starIndeces,
               deck_json=np.frombuffer(json.dumps(scal).encode(), dtype=np.uint8))
    for i, g in enumerate(model.grids, start=1):
        out.update({f"g{i}_xAxis": g.xAxis, f"g{i}_yAxis": g.yAxis, f"g{i}_zAxis": g.zAxis, f"g{i}_active": g.active,
                    f"g{i}_Hden": g.Hden, f"g{i}_nCells": np.asarray([g.nCells], I32)})
        if model.lgDust:
            out[f"g{i}_Ndust"] = g.Ndust
    if model.lgDust:
        t = tables["dust"]
        out.update(grainWeight=t["grainWeight"], grainAbun1=t["grainAbun1"], TdustSublime=t["TdustSublime"],
                   dustScaXsecP=t["dustScaXsecP"], dustAbsXsecP=t["dustAbsXsecP"], gSca=t["gSca"])
    return out


def multideck_from_arrays(a: dict):
    """Inverse of multideck_to_arrays: (Model, tables, scalars) without the deck's files."""
    import json

    from .gasdeck import _XT_SCALARS
    from .opacity import XSecTables

    s = json.loads(bytes(a["deck_json"]).decode())
    nu = np.asarray(a["nuArray"], dtype=F32)
    nbins = nu.shape[0]
    xt = XSecTables(xSecArray=np.asarray(a["xSecArray"], F32), lgElementOn=np.asarray(a["lgElementOn"], I32),
                    elementXref=np.asarray(a["elementXref"], I32), elementP=np.asfortranarray(a["elementP"], dtype=I32),
                    nShells=np.asfortranarray(a["nShells"], dtype=I32), **{k: int(s[k]) for k in _XT_SCALARS})
    on, xref = xt.lgElementOn, xt.elementXref
    elemAbun = np.asfortranarray(a["elemAbun"], dtype=F32)
    lgDust = bool(s["lgDust"])
    dustT = None
    if lgDust:
        dustT = dict(grainWeight=np.asarray(a["grainWeight"], F32), grainAbun1=np.asarray(a["grainAbun1"], F32),
                     TdustSublime=np.asarray(a["TdustSublime"], F32), dustScaXsecP=np.asfortranarray(a["dustScaXsecP"], dtype=I32),
                     dustAbsXsecP=np.asfortranarray(a["dustAbsXsecP"], dtype=I32), gSca=np.asarray(a["gSca"], F32))
        dustT["nSp"], dustT["nSz"] = dustT["dustScaXsecP"].shape
    grids, per_grid = [], []
    for i in range(1, int(s["nGrids"]) + 1):
        act = np.asfortranarray(a[f"g{i}_active"], dtype=I32)
        g = Grid(xAxis=np.asarray(a[f"g{i}_xAxis"], F32), yAxis=np.asarray(a[f"g{i}_yAxis"], F32),
                 zAxis=np.asarray(a[f"g{i}_zAxis"], F32), active=act, nCells=int(a[f"g{i}_nCells"][0]))
        g.motherP = int(s["motherP"][i - 1])
        g.Hden = np.asarray(a[f"g{i}_Hden"], F32)
        g.Te = np.zeros(g.nCells + 1, dtype=F32)
        g.Te[1:] = F32(s["TeStart"])
        g.Ne = g.Hden.copy()
        ionDen = G.initial_ion_state(g.nCells, on, xref, int(on.sum()), xt.nstages)
        abIndex = np.ones(g.nCells + 1, dtype=I32)
        entry = dict(ionDen=ionDen, den=xt.species_densities(ionDen, elemAbun, abIndex, g.Hden), abIndex=abIndex, dust=None)
        if lgDust:
            g.Ndust = np.asarray(a[f"g{i}_Ndust"], F32)
            g.dustAbunIndex = np.ones(g.nCells + 1, dtype=I32)
            g.Tdust = np.full((dustT["nSp"] + 1, dustT["nSz"] + 1, g.nCells + 1), F32(50.0), dtype=F32, order="F")
            entry["dust"] = dict(Ndust=g.Ndust, Tdust=g.Tdust, dustAbunIndex=None, grainWeight=dustT["grainWeight"],
                                 dustScaXsecP=dustT["dustScaXsecP"], dustAbsXsecP=dustT["dustAbsXsecP"])
        g.recPDF = np.zeros((g.nCells + 1, nbins), dtype=F32, order="F")
        g.recPDF[1:, :] = np.asarray(a["recRow"], F32)[None, :]
        g.totalLines = np.zeros(g.nCells + 1, dtype=F32)
        g.totalLines[1:] = F32(a["totalLines1"][0])
        grids.append(g)
        per_grid.append(entry)
    kw = {}
    if lgDust:
        grainAbun = np.zeros((1, dustT["nSp"]), dtype=F32, order="F")
        grainAbun[0, :] = dustT["grainAbun1"]
        kw = dict(gSca=dustT["gSca"], nSpeciesMax=dustT["nSp"], nSizes=dustT["nSz"], nSpeciesPart=np.asarray([dustT["nSp"]], I32),
                  grainAbun=grainAbun, dustComPoint=np.asarray([1], I32), TdustSublime=dustT["TdustSublime"])
    model = Model(grids=grids, nbins=nbins, nuArray=nu,
                  inSpectrumProbDen=np.stack([np.zeros(nbins, F32), np.asarray(a["cdf"], F32)]).astype(F32),
                  deltaE=np.asarray(a["deltaE"], dtype=F32), starPosition=np.asarray(a["starPositionAbs"], dtype=F32),
                  starIndeces=np.asarray(a["starIndeces"], dtype=I32), lgDust=lgDust, lgGas=True,
                  lgSymmetricXYZ=bool(s["lgSymmetricXYZ"]), lgIsotropic=bool(s["lgIsotropic"]), R_out=float(s["R_out"]),
                  ionEdge1=float(s["ionEdge1"]), **kw)
    tables = dict(xsec=xt, bands=xt.band_list(nbins), grids=per_grid, elemAbun=elemAbun, widFlx=wid_flx(nu),
                  nstages=xt.nstages, lgElementOn=on, elementXref=xref, dust=dustT, recPDF_kind=s["recPDF_kind"])
    return model, tables, s
