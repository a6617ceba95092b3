"""Synthetic transport workloads (numpy only): the inputs the hot path needs at the
start of a Lucy iteration -- grids, opacities, re-emission CDFs, stellar CDF -- built with
the reference's recipes where they are simple (axes, active numbering, star placement,
blackbody CDF) and with smooth analytic stand-ins for the atomic physics that is out of
scope (cross-sections, recombination spectrum, grain optics).  Everything is seeded.

Shapes follow BASELINE.md section 4:
  hii_region      13^3 octant, symmetricXYZ, gas only (HII40-like), nbins 600
  dust_shell      16^3 octant, symmetricXYZ, dust only (benchmarks/dust/1D-like), nbins 215
  multigrid       16^3 mother + 11^3 sub-grid in the inner corner, gas+dust
  synthetic_cube  n^3 full cube, star at the centre, uniform or clumpy, gas(+dust)
  viewing_angles  small non-symmetric dust cube with `inclination` viewing angles
"""
from __future__ import annotations

import numpy as np

from .model import F32, I32, Grid, Model, auto_axis, farray, number_active, star_indices

TE1RYD = 1.578866e5
HC_RYD_K = 1.578866e5      # h c Ryd / k  [K]  (the reference's getFlux uses 157893.94 in float32: these builders are
                           # synthetic stand-ins; mocassin_b200/deck.py has the faithful, pinned versions)


# ---------------------------------------------------------------------------------------
def nu_mesh(nbins: int, nuMin: float = 1.001e-5, nuMax: float = 15.0, edges=(0.99946, 1.8071406, 3.9996377)):
    """Log mesh with a pair of points +-0.0003 around each ionisation edge (the gist of
    grid_mod.f90:213-258), ascending, float32; widFlx = centred differences (:333-338)."""
    pts = []
    for e in edges:
        if nuMin < e < nuMax:
            pts += [e - 0.0003, e + 0.0003]
    nlog = nbins - len(pts)
    mesh = np.exp(np.linspace(np.log(nuMin), np.log(nuMax), nlog))
    nu = np.sort(np.concatenate([mesh, np.array(pts)])).astype(F32)
    assert nu.shape[0] == nbins and np.all(np.diff(nu) > 0)
    wid = np.empty(nbins, dtype=F32)
    wid[1:-1] = (nu[2:] - nu[:-2]) / F32(2.0)
    wid[0] = nu[1] - nu[0]
    wid[-1] = nu[-1] - nu[-2]
    return nu, wid


def blackbody_cdf(T: float, nu: np.ndarray, wid: np.ndarray) -> np.ndarray:
    """setProbDen (continuum_mod.f90:418-474) for contShape blackbody: running sum of
    B_nu*widFlx, normalised, entries >= max forced to 1."""
    x = HC_RYD_K * nu.astype(np.float64) / T
    with np.errstate(over="ignore"):
        b = np.where(x > 86.0, nu.astype(np.float64) ** 3 * np.exp(-np.minimum(x, 700.0)),
                     nu.astype(np.float64) ** 3 / np.expm1(np.minimum(x, 700.0)))
    w = b * wid
    cdf = np.cumsum((w / w.sum()).astype(F32), dtype=F32)
    cdf[cdf >= cdf.max()] = F32(1.0)
    return cdf.astype(F32)


def gas_cross_sections(nu: np.ndarray):
    """Smooth stand-ins for the H0, He0, He+ photoionisation cross-sections [cm^2]."""
    nu = nu.astype(np.float64)
    sH = np.where(nu >= 0.99946, 6.30e-18 * (nu / 0.99946) ** -3.0, 0.0)
    x = nu / 1.8071406
    sHe = np.where(x >= 1.0, 7.4e-18 * (1.66 * x ** -2.05 - 0.66 * x ** -3.05), 0.0)
    sHe2 = np.where(nu >= 3.9996377, 1.58e-18 * (nu / 3.9996377) ** -3.0, 0.0)
    return sH.astype(F32), sHe.astype(F32), sHe2.astype(F32)


def recombination_cdf(nu: np.ndarray, wid: np.ndarray, Te: float) -> np.ndarray:
    """A smooth recombination-continuum emissivity (free-bound edges of H n=1..4, He0,
    He+, plus free-free) turned into a normalised cumulative table like setDiffusePDF
    (emission_mod.f90:1186-1205): entries > 0.999998 forced to 1."""
    nu64 = nu.astype(np.float64)
    j = 0.05 * np.exp(-nu64 * TE1RYD / Te)
    for edge, wgt in ((0.99946, 1.0), (0.99946 / 4, 0.35), (0.99946 / 9, 0.15), (0.99946 / 16, 0.08),
                      (1.8071406, 0.12), (3.9996377, 0.02)):
        j += np.where(nu64 >= edge, wgt * np.exp(-(nu64 - edge) * TE1RYD / Te), 0.0)
    w = j * wid
    cdf = np.cumsum((w / w.sum()).astype(F32), dtype=F32)
    cdf = (cdf / cdf[-1]).astype(F32)
    cdf[cdf > F32(0.999998)] = F32(1.0)
    return cdf


def dust_optics(nu: np.ndarray, a_um: float = 0.16):
    """Smooth grain optics for one size: Cabs, Csca [cm^2] and asymmetry g(nu)."""
    lam_um = 0.0911267 / nu.astype(np.float64)          # 1 Ryd = 911.267 A
    x = 2.0 * np.pi * a_um / lam_um
    qabs = np.minimum(1.0, x) * (1.0 + 0.5 * np.exp(-((np.log(lam_um / 9.7)) ** 2) / 0.02))
    qsca = np.minimum(1.5, x ** 4 / (1.0 + x ** 3 / 1.5))
    g = 0.65 * x ** 2 / (1.0 + x ** 2)
    area = np.pi * (a_um * 1.0e-4) ** 2
    return (qabs * area).astype(F32), (qsca * area).astype(F32), g.astype(F32)


def planck_cdf(T: float, nu: np.ndarray, wid: np.ndarray, cabs: np.ndarray) -> np.ndarray:
    """setDustPDF-like row (emission_mod.f90:1313-1387): cumulative B_nu(T)*Cabs, last = 1."""
    x = HC_RYD_K * nu.astype(np.float64) / T
    b = nu.astype(np.float64) ** 3 / np.expm1(np.minimum(x, 700.0))
    w = b * cabs * wid
    cdf = np.cumsum((w / w.sum()).astype(F32), dtype=F32)
    cdf = (cdf / cdf[-1]).astype(F32)
    cdf[-1] = F32(1.0)
    return np.maximum.accumulate(cdf).astype(F32)


def _radius(g_x, g_y, g_z):
    x, y, z = np.meshgrid(g_x.astype(np.float64), g_y.astype(np.float64), g_z.astype(np.float64), indexing="ij")
    return np.sqrt(x * x + y * y + z * z)


def _per_cell(active: np.ndarray, field3d: np.ndarray, nCells: int, dtype=F32) -> np.ndarray:
    out = np.zeros(nCells + 1, dtype=dtype)
    m = active > 0
    out[active[m]] = field3d[m]
    return out


def _expand(percell: np.ndarray, spectrum: np.ndarray) -> np.ndarray:
    """(nCells+1,) x (nbins,) -> (nCells+1, nbins) F-order float32 outer product."""
    out = np.empty((percell.shape[0], spectrum.shape[0]), dtype=F32, order="F")
    np.multiply(percell[:, None].astype(F32), spectrum[None, :].astype(F32), out=out)
    return out


def _rows_from_table(idx: np.ndarray, table: np.ndarray) -> np.ndarray:
    """pdf(cell, :) = table[idx[cell], :] as an (nCells+1, nbins) F-order array."""
    out = np.empty((idx.shape[0], table.shape[1]), dtype=F32, order="F")
    np.take(table.T, idx, axis=1, out=out.T)
    return out


# ---------------------------------------------------------------------------------------
def _gas_tables(model: Model, g: Grid, Hden3d, xH0_3d, nu, wid, rng, nTe=16):
    """opacity, recPDF, totalLines of a gas grid from a neutral-fraction field."""
    sH, sHe, sHe2 = gas_cross_sections(nu)
    nH0 = _per_cell(g.active, Hden3d * xH0_3d, g.nCells)
    nHe0 = _per_cell(g.active, 0.1 * Hden3d * np.minimum(1.0, 3.0 * xH0_3d), g.nCells)
    nHe1 = _per_cell(g.active, 0.1 * Hden3d * (1.0 - np.minimum(1.0, 3.0 * xH0_3d)) * 0.9, g.nCells)
    op = _expand(nH0, sH)
    op += _expand(nHe0, sHe)
    op += _expand(nHe1, sHe2)
    g.opacity = op
    g.Hden = _per_cell(g.active, Hden3d, g.nCells)
    Tes = np.linspace(6000.0, 12000.0, nTe)
    table = np.stack([recombination_cdf(nu, wid, T) for T in Tes]).astype(F32)
    idx = rng.integers(0, nTe, size=g.nCells + 1)
    g.recPDF = _rows_from_table(idx, table)
    g.recPDF[0, :] = 0.0
    tl = np.zeros(g.nCells + 1, dtype=F32)
    tl[1:] = (0.55 + 0.2 * rng.random(g.nCells)).astype(F32)
    g.totalLines = tl


def _dust_tables(model: Model, g: Grid, Nd3d, nu, wid, rng, Tdust3d=None, add_to_opacity=True, nT=24):
    cabs, csca, gs = dust_optics(nu)
    model.gSca = gs
    nd = _per_cell(g.active, Nd3d, g.nCells)
    g.Ndust = nd
    g.scaOpac = _expand(nd, csca)
    g.absOpac = _expand(nd, cabs)
    if g.opacity is None or not add_to_opacity:
        g.opacity = (g.scaOpac + g.absOpac).astype(F32, order="F")
    else:
        g.opacity = (g.opacity + (g.scaOpac + g.absOpac)).astype(F32, order="F")
    g.Tdust = np.zeros((model.nSpeciesMax + 1, model.nSizes + 1, g.nCells + 1), dtype=F32, order="F")
    if Tdust3d is None:
        Td = np.full(g.nCells + 1, 50.0, dtype=F32)
    else:
        Td = _per_cell(g.active, Tdust3d, g.nCells)
    g.Tdust[:, :, :] = Td[None, None, :]
    g.dustAbunIndex = np.ones(g.nCells + 1, dtype=I32)
    if not model.lgGas:
        Ts = np.exp(np.linspace(np.log(30.0), np.log(1500.0), nT))
        table = np.stack([planck_cdf(T, nu, wid, cabs) for T in Ts]).astype(F32)
        idx = np.clip(np.searchsorted(Ts, Td), 0, nT - 1)
        g.dustPDF = _rows_from_table(idx, table)
        g.dustPDF[0, :] = 0.0


def _finish_model(grids, nu, cdf_rows, deltaE, star_pos, star_idx, **kw) -> Model:
    return Model(grids=grids, nbins=int(nu.shape[0]), nuArray=nu,
                 inSpectrumProbDen=np.asarray(cdf_rows, dtype=F32), deltaE=np.asarray(deltaE, dtype=F32),
                 starPosition=np.asarray(star_pos, dtype=F32).reshape(-1, 3),
                 starIndeces=np.asarray(star_idx, dtype=I32).reshape(-1, 4), **kw)


def _dust_model_kw():
    return dict(nSpeciesMax=1, nSizes=1, nSpeciesPart=np.ones(1, dtype=I32),
                grainAbun=np.ones((1, 1), dtype=F32, order="F"), dustComPoint=np.ones(1, dtype=I32),
                TdustSublime=np.full(1, 1400.0, dtype=F32))


# ---------------------------------------------------------------------------------------
def hii_region(n: int = 13, nbins: int = 600, Tstar: float = 40000.0, Rin: float = 3.0e18, Rout: float = 1.46e19,
               edge: float = 1.4e19, Hden: float = 100.0, nPhotons: int = 1_000_000, Lstar: float = 308.2,
               nuMax: float = 15.0, seed: int = 7, debug: bool = False) -> Model:
    """HII40-like (benchmarks/gas/HII40/input.in): 13^3 octant, symmetricXYZ, gas only."""
    rng = np.random.default_rng(seed)
    nu, wid = nu_mesh(nbins, nuMax=nuMax)
    ax = auto_axis(n, edge, True)
    r = _radius(ax, ax, ax)
    mask = (r >= Rin) & (r <= Rout)
    active, nCells = number_active(mask)
    g = Grid(xAxis=ax, yAxis=ax.copy(), zAxis=ax.copy(), active=active, nCells=nCells)
    xH0 = np.clip(1.0e-4 * np.exp((r - Rin) / (0.12 * (Rout - Rin))), 1.0e-4, 1.0)
    model = _finish_model([g], nu, np.stack([np.zeros(nbins, F32), blackbody_cdf(Tstar, nu, wid)]),
                          [0.0, Lstar / nPhotons], [[0.0, 0.0, 0.0]], [[1, 1, 1, 1]],
                          lgDust=False, lgGas=True, lgSymmetricXYZ=True, R_out=Rout, lgDebug=debug)
    _gas_tables(model, g, np.full(r.shape, Hden), xH0, nu, wid, rng)
    if debug:
        model.nLines = 20
        lp = np.cumsum(rng.random((g.nCells + 1, model.nLines)), axis=1)
        g.linePDF = np.asfortranarray((lp / lp[:, -1:]).astype(F32))
    return model


def dust_shell(n: int = 16, nbins: int = 215, Tstar: float = 2500.0, Rout: float = 2.18e17, Rin: float = 2.18e16,
               tauV: float = 1.0, nPhotons: int = 100_000, Lstar: float = 38.26, isotropic: bool = False,
               seed: int = 11) -> Model:
    """benchmarks/dust/1D-like: octant, symmetricXYZ, dust only, shell with n ~ r^-2."""
    rng = np.random.default_rng(seed)
    nu, wid = nu_mesh(nbins, nuMin=1.0e-4, nuMax=15.0, edges=())
    ax = auto_axis(n, Rout, True)
    r = _radius(ax, ax, ax)
    mask = (r >= Rin) & (r <= Rout)
    active, nCells = number_active(mask)
    g = Grid(xAxis=ax, yAxis=ax.copy(), zAxis=ax.copy(), active=active, nCells=nCells)
    cabs, csca, _ = dust_optics(nu)
    iV = int(np.argmin(np.abs(nu - 0.1657)))          # 0.55 um
    n0 = tauV / ((cabs[iV] + csca[iV]) * max(Rin, 1.0) * (1.0 - Rin / Rout))
    rr = np.maximum(r, max(Rin, 1.0e-3 * Rout))
    if Rin > 0.0:
        Nd = np.where(mask, n0 * (Rin / rr) ** 2, 0.0)
        Td = np.where(mask, 900.0 * (Rin / rr) ** 0.45, 0.0)
    else:                                   # filled sphere: uniform density
        Nd = np.where(mask, tauV / ((cabs[iV] + csca[iV]) * Rout), 0.0)
        Td = np.where(mask, 300.0, 0.0)
    model = _finish_model([g], nu, np.stack([np.zeros(nbins, F32), blackbody_cdf(Tstar, nu, wid)]),
                          [0.0, Lstar / nPhotons], [[0.0, 0.0, 0.0]], [[1, 1, 1, 1]],
                          lgDust=True, lgGas=False, lgSymmetricXYZ=True, R_out=Rout, lgIsotropic=isotropic,
                          **_dust_model_kw())
    _dust_tables(model, g, Nd, nu, wid, rng, Tdust3d=Td)
    return model


def wid_flx(nu: np.ndarray) -> np.ndarray:
    """widFlx of initCartesianGrid (grid_mod.f90:333-338)."""
    w = np.empty_like(nu, dtype=F32)
    w[0] = nu[1] - nu[0]
    w[1:-1] = (nu[2:] - nu[:-2]) / F32(2.0)
    w[-1] = nu[-1] - nu[-2]
    return w


def dust_closure(n: int = 9, nbins: int = 120, nPhotons: int = 200_000, Lstar: float = 38.26, Tstar: float = 2500.0,
                 Rout: float = 2.18e17, Rin: float = 2.18e16, tauV: float = 1.0, multiChem: bool = True,
                 T0: float = 100.0, nTemps: int = 3000, seed: int = 23):
    """Dust-only shell with a full dust description, for the device closure of the iteration
    (getDustT + setDustPDF): 3 grain species in 2 chemistry components ([sp1,sp2] and [sp3]),
    3 grain sizes.  Returns (model, tables) where tables holds xSecArray, dustAbsXsecP,
    dustScaXsecP, grainWeight, widFlx and dustEmIntegral (dust_mod.f90:155-181)."""
    nu, wid = nu_mesh(nbins, nuMin=1.0e-4, nuMax=15.0, edges=())
    widFlx = wid_flx(nu)
    ax = auto_axis(n, Rout, True)
    r = _radius(ax, ax, ax)
    mask = (r >= Rin) & (r <= Rout)
    active, nCells = number_active(mask)
    g = Grid(xAxis=ax, yAxis=ax.copy(), zAxis=ax.copy(), active=active, nCells=nCells)
    sizes = np.array([0.05, 0.16, 0.5])
    nSizes, nSpecies = 3, 3
    w = sizes ** -2.5
    grainWeight = (w / w.sum()).astype(F32)
    qscale = np.array([1.0, 0.55, 1.6])               # per-species absorption efficiency factor
    xs = [np.zeros(1, F32)]                           # entry 1 unused: offsets are 1-based
    absP = np.zeros((nSpecies, nSizes), dtype=I32, order="F")
    scaP = np.zeros((nSpecies, nSizes), dtype=I32, order="F")
    pos = 2
    gs = None
    for nS in range(nSpecies):
        for ai in range(nSizes):
            cabs, csca, gsc = dust_optics(nu, float(sizes[ai]))
            gs = gsc if ai == 1 else gs
            absP[nS, ai] = pos; xs.append((cabs * qscale[nS]).astype(F32)); pos += nbins
            scaP[nS, ai] = pos; xs.append(csca.astype(F32)); pos += nbins
    xSecArray = np.concatenate(xs).astype(F32)
    nSpeciesPart = np.array([2, 1], dtype=I32)
    dustComPoint = np.array([1, 3], dtype=I32)
    grainAbun = np.array([[0.6, 0.4], [1.0, 0.0]], dtype=F32, order="F")
    TdustSublime = np.array([1400.0, 1200.0, 900.0], dtype=F32)
    # dustEmIntegral(nS,ai,T) = 4 h fr1Ryd sum_i Cabs(i) (B_nu(T)/h) widFlx(i), T = 1..nTemps K
    T = np.arange(1, nTemps + 1, dtype=np.float64)
    x = HC_RYD_K * nu.astype(np.float64)[None, :] / T[:, None]
    bb = (0.5250229 / 6.6262e-27) * nu.astype(np.float64)[None, :] ** 3 / np.expm1(np.minimum(x, 700.0))
    em = np.zeros((nSpecies, nSizes, nTemps), dtype=F32, order="F")
    for nS in range(nSpecies):
        for ai in range(nSizes):
            cabs = xSecArray[absP[nS, ai] - 1: absP[nS, ai] - 1 + nbins].astype(np.float64)
            em[nS, ai, :] = ((bb * (cabs * 3.28984e15 * widFlx.astype(np.float64))[None, :]).sum(axis=1)
                             * 6.6262e-27 * 4.0).astype(F32)
    comp3d = np.where(r < 0.55 * Rout, 1, 2) if multiChem else np.ones_like(r, dtype=np.int64)
    # mean cross-sections per component for the opacity
    def mean_xs(P, comp):
        out = np.zeros(nbins, dtype=np.float64)
        dcp = dustComPoint[comp - 1]
        for k in range(nSpeciesPart[comp - 1]):
            for ai in range(nSizes):
                o = P[dcp - 1 + k, ai] - 1
                out += grainAbun[comp - 1, k] * grainWeight[ai] * xSecArray[o:o + nbins].astype(np.float64)
        return out
    iV = int(np.argmin(np.abs(nu - 0.1657)))
    ext1 = mean_xs(absP, 1) + mean_xs(scaP, 1)
    n0 = tauV / (ext1[iV] * Rin * (1.0 - Rin / Rout))
    Nd3 = np.where(mask, n0 * (Rin / np.maximum(r, Rin)) ** 2, 0.0)
    nd = _per_cell(active, Nd3, nCells)
    comp = _per_cell(active, comp3d, nCells, dtype=I32)
    comp[0] = 1
    g.Ndust = nd
    g.dustAbunIndex = comp
    g.absOpac = np.zeros((nCells + 1, nbins), dtype=F32, order="F")
    g.scaOpac = np.zeros((nCells + 1, nbins), dtype=F32, order="F")
    for c in (1, 2):
        sel = comp == c
        g.absOpac[sel, :] = (nd[sel, None] * mean_xs(absP, c)[None, :]).astype(F32)
        g.scaOpac[sel, :] = (nd[sel, None] * mean_xs(scaP, c)[None, :]).astype(F32)
    g.opacity = (g.absOpac + g.scaOpac).astype(F32, order="F")
    g.Tdust = np.zeros((2 + 1, nSizes + 1, nCells + 1), dtype=F32, order="F")
    g.Tdust[:, :, 1:] = F32(T0)
    model = _finish_model([g], nu, np.stack([np.zeros(nbins, F32), blackbody_cdf(Tstar, nu, wid)]),
                          [0.0, Lstar / nPhotons], [[0.0, 0.0, 0.0]], [[1, 1, 1, 1]],
                          lgDust=True, lgGas=False, lgSymmetricXYZ=True, R_out=Rout, gSca=gs,
                          lgMultiDustChemistry=multiChem, nSpeciesMax=2, nSizes=nSizes, nSpeciesPart=nSpeciesPart,
                          grainAbun=grainAbun, dustComPoint=dustComPoint, TdustSublime=TdustSublime)
    tables = dict(xSecArray=xSecArray, dustAbsXsecP=absP, dustScaXsecP=scaP, grainWeight=grainWeight,
                  widFlx=widFlx, dustEmIntegral=em)
    return model, tables


def multigrid(n: int = 16, nsub: int = 11, nbins: int = 600, Tstar: float = 80000.0, Rin: float = 1.0e15,
              Rout: float = 1.0e18, sub_hi: float = 2.0e17, nPhotons: int = 1_000_000, Lstar: float = 1.0,
              symmetric: bool = True, seed: int = 13) -> Model:
    """examples/multigridgasdust-like: mother grid + one denser sub-grid covering the
    mother cells inside [0, sub_hi]^3 (symmetric) or [-sub_hi/2, sub_hi/2]^3."""
    rng = np.random.default_rng(seed)
    nu, wid = nu_mesh(nbins)
    ax = auto_axis(n, Rout, symmetric)
    r = _radius(ax, ax, ax)
    mask = (r >= Rin) & (r <= Rout)
    lo = 0.0 if symmetric else -0.5 * sub_hi
    hi = sub_hi if symmetric else 0.5 * sub_hi
    # mother cells whose centre lies inside the sub-grid box point to grid 2
    # (setSubGrids, grid_mod.f90:835-889)
    inside = (ax > lo) & (ax < hi) if not symmetric else (ax >= lo) & (ax < hi)
    box = inside[:, None, None] & inside[None, :, None] & inside[None, None, :]
    active, nCells = number_active(mask & ~box)
    active[box] = -2
    gm = Grid(xAxis=ax, yAxis=ax.copy(), zAxis=ax.copy(), active=active, nCells=nCells, motherP=0)
    sx = (lo + (hi - lo) * np.arange(nsub, dtype=np.float64) / (nsub - 1)).astype(F32)
    rs = _radius(sx, sx, sx)
    smask = rs >= Rin
    sactive, snCells = number_active(smask)
    gs = Grid(xAxis=sx, yAxis=sx.copy(), zAxis=sx.copy(), active=sactive, nCells=snCells, motherP=1)
    star = [0.0, 0.0, 0.0]
    # star sits in a mother cell that maps to the sub-grid: starIndeces(:,4) = 2
    sidx = star_indices(gs, star) + [2]
    model = _finish_model([gm, gs], nu, np.stack([np.zeros(nbins, F32), blackbody_cdf(Tstar, nu, wid)]),
                          [0.0, Lstar / nPhotons], [star], [sidx],
                          lgDust=True, lgGas=True, lgSymmetricXYZ=symmetric, R_out=Rout, **_dust_model_kw())
    for g, rr, dens in ((gm, r, 3.0), (gs, rs, 30.0)):
        xH0 = np.clip(3.0e-4 * np.exp(rr / (0.15 * Rout)), 1.0e-4, 1.0)
        _gas_tables(model, g, np.full(rr.shape, dens), xH0, nu, wid, rng)
        _dust_tables(model, g, np.where(g.active > 0, 3.0e-10 * dens, 0.0), nu, wid, rng)
    return model


def clumpy_field(n: int, rng, sigma: float = 1.5, corr: float = 4.0, ff: float = 0.1, contrast: float = 30.0):
    """Log-normal density field (Gaussian-filtered white noise, correlation length `corr`
    cells) with a clump mask of filling factor `ff` and density contrast `contrast`."""
    white = rng.standard_normal((n, n, n)).astype(np.float32)
    k = np.fft.fftfreq(n).astype(np.float32)
    k2 = k[:, None, None] ** 2 + k[None, :, None] ** 2 + k[None, None, : n // 2 + 1] ** 2
    filt = np.exp(-0.5 * k2 * (2.0 * np.pi * corr) ** 2)
    f = np.fft.irfftn(np.fft.rfftn(white, axes=(0, 1, 2)) * filt, s=(n, n, n), axes=(0, 1, 2)).astype(np.float64)
    f = (f - f.mean()) / f.std()
    rho = np.exp(sigma * f - 0.5 * sigma ** 2)
    thr = np.quantile(f, 1.0 - ff)
    rho = np.where(f >= thr, rho * contrast, rho)
    return rho / rho.mean()


def synthetic_cube(n: int = 128, nbins: int = 600, clumpy: bool = True, dust: bool = True, edge: float = 3.0e18,
                   Tstar: float = 40000.0, Hden: float = 100.0, nPhotons: int = 1_000_000_000,
                   Lstar: float = 308.2, seed: int = 2024, build_tables: bool = True) -> Model:
    """S-uniform / S-clumpy of BASELINE.md: n^3 full cube (non-symmetric), star at the
    centre, gas opacity from a prescribed neutral fraction, optional dust."""
    rng = np.random.default_rng(seed)
    nu, wid = nu_mesh(nbins)
    ax = auto_axis(n, edge, False)
    mask = np.ones((n, n, n), dtype=bool)
    active, nCells = number_active(mask)
    g = Grid(xAxis=ax, yAxis=ax.copy(), zAxis=ax.copy(), active=active, nCells=nCells)
    star = [0.0, 0.0, 0.0]
    sidx = star_indices(g, star) + [1]
    kw = _dust_model_kw() if dust else {}
    model = _finish_model([g], nu, np.stack([np.zeros(nbins, F32), blackbody_cdf(Tstar, nu, wid)]),
                          [0.0, Lstar / nPhotons], [star], [sidx],
                          lgDust=dust, lgGas=True, lgSymmetricXYZ=False, R_out=0.0, **kw)
    r = _radius(ax, ax, ax)
    rho = clumpy_field(n, rng) if clumpy else np.ones((n, n, n))
    Hd = Hden * rho
    # neutral fraction: ionised inside a density-dependent Stroemgren-like radius
    rs = 0.55 * edge * (rho.clip(0.05, 50.0)) ** (-1.0 / 3.0)
    xH0 = np.clip(1.0e-4 * np.exp(np.clip((r - rs) / (0.05 * edge), -20.0, 9.3)), 1.0e-4, 1.0)
    model._fields3d = dict(Hden=Hd, xH0=xH0, Ndust=(1.0e-12 * Hd if dust else None))
    if build_tables:
        _gas_tables(model, g, Hd, xH0, nu, wid, rng)
        if dust:
            _dust_tables(model, g, 1.0e-12 * Hd, nu, wid, rng)
    return model


def viewing_angles(n: int = 9, nbins: int = 64, nPhotons: int = 100_000, seed: int = 5, phi_free: bool = False) -> Model:
    """Small non-symmetric dust-only cube with two `inclination` viewing angles
    (nAngleBins=2) to exercise the escape binning (photon_mod.f90:414-462)."""
    rng = np.random.default_rng(seed)
    nu, wid = nu_mesh(nbins, nuMin=1.0e-3, nuMax=10.0, edges=())
    edge = 1.0e17
    ax = auto_axis(n, edge, False)
    r = _radius(ax, ax, ax)
    mask = (r <= 1.05 * edge) & (r >= 0.15 * edge)
    active, nCells = number_active(mask)
    g = Grid(xAxis=ax, yAxis=ax.copy(), zAxis=ax.copy(), active=active, nCells=nCells)
    star = [0.0, 0.0, 0.0]
    sidx = star_indices(g, star) + [1]
    vt = np.array([0.0, np.deg2rad(30.0), np.deg2rad(110.0)], dtype=F32)
    vp = np.array([0.0, -1.0, -1.0], dtype=F32) if phi_free else np.array([0.0, np.deg2rad(40.0), np.deg2rad(200.0)], dtype=F32)
    model = _finish_model([g], nu, np.stack([np.zeros(nbins, F32), blackbody_cdf(6000.0, nu, wid)]),
                          [0.0, 1.0 / nPhotons], [star], [sidx],
                          lgDust=True, lgGas=False, lgSymmetricXYZ=False, R_out=1.05 * edge,
                          nAngleBins=2, viewPointTheta=vt, viewPointPhi=vp, **_dust_model_kw())
    cabs, csca, _ = dust_optics(nu)
    iV = int(np.argmin(np.abs(nu - 0.1657)))
    Nd = np.where(mask, 2.0 / ((cabs[iV] + csca[iV]) * edge), 0.0)
    _dust_tables(model, g, Nd, nu, wid, rng, Tdust3d=np.where(mask, 300.0, 0.0))
    return model


def plane_slab(nx: int = 9, ny: int = 17, nz: int = 9, nbins: int = 120, dust: bool = True, Tstar: float = 35000.0,
               Hden: float = 300.0, nPhotons: int = 100_000, seed: int = 21) -> Model:
    """Plane-parallel ionisation (`planeIonization` keyword): packets enter through the y=0
    face along +y, are mirrored at the x and z faces and leave through the y faces
    (photon_mod.f90:561-646, 2199-2414)."""
    rng = np.random.default_rng(seed)
    nu, wid = nu_mesh(nbins)
    ax = auto_axis(nx, 1.0e17, False)
    ay = (np.arange(ny, dtype=F32) / F32(ny - 1) * F32(4.0e17)).astype(F32)
    az = auto_axis(nz, 1.0e17, False)
    mask = np.ones((nx, ny, nz), dtype=bool)
    active, nCells = number_active(mask)
    g = Grid(xAxis=ax, yAxis=ay, zAxis=az, active=active, nCells=nCells)
    kw = _dust_model_kw() if dust else {}
    model = _finish_model([g], nu, np.stack([np.zeros(nbins, F32), blackbody_cdf(Tstar, nu, wid)]),
                          [0.0, 1.0 / nPhotons], [[0.0, 0.0, 0.0]], [star_indices(g, [0.0, 0.0, 0.0]) + [1]],
                          lgDust=dust, lgGas=True, lgSymmetricXYZ=False, lgPlaneIonization=True, R_out=0.0, **kw)
    y = np.broadcast_to(ay.astype(np.float64)[None, :, None], (nx, ny, nz))
    xH0 = np.clip(1.0e-3 * np.exp(y / 6.0e16), 1.0e-3, 1.0)
    _gas_tables(model, g, np.full((nx, ny, nz), Hden), xH0, nu, wid, rng)
    if dust:
        _dust_tables(model, g, np.full((nx, ny, nz), 3.0e-10 * Hden), nu, wid, rng)
    return model
