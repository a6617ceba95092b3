"""Run the reference's own routines either side of the transport (oracle/_ref/mocassin_ref_aux.py,
the f90py translation) -- TEST INFRASTRUCTURE:

* `gas_opacity`  : `ionizationDriver(grid, ix, iy, iz)` for every cell (ionization_mod.f90:26-129:
                   density loop, eDenSum, addOpacity/putOpacity/inOpacity).  BoltGaunt, whose only
                   trace in the result is the free-free term of bin 1, is supplied by the harness.
* `photo`        : lines 168-269 of updateCell (nPhotoSte/nPhotoDif per element and ion) and lines
                   1123-1234 of thermBalance (heatSte/heatDif), with getOuterShell (hydro_mod.f90).
* `write_sed`    : `writeSED(grid)` (output_mod.f90:2508-2719); the records it writes to unit 16
* `bhmie`       : `BHmie` (ph_mod.f90:1600-1757), COMPLEX arithmetic and statement functions
* `write_tau_nu`  : `writeTauNu(grid)` over `integratePathTauNu`; unit 73 records
* `write_cont_cube`: `writeContCube(grid, freq1, freq2)` (output_mod.f90:2722-2806); unit 19 records
                   are captured instead of going to output/SED.out.
* `write_grid`   : `writeGrid(grid)` (grid_mod.f90:2646-2870): the records of grid0-3.out, dustGrid.out
                   and photoSource.out, captured per unit.
* `dust_pdf`     : `emissionDriver(grids, ix, iy, iz, iG)` on a dust-only model = setDustPDF
                   (emission_mod.f90:1313-1387) with getFlux (continuum_mod.f90:359-416).
* `dust_update`  : `updateCell(grid, xP, yP, zP)` on a dust-only model = the no-gas branch
                   (update_mod.f90:308-334) with getDustT (:1836-1945).
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

from . import build_ref, rt


def load_aux():
    path = build_ref.build(target='aux')
    d = os.path.dirname(path)
    if d not in sys.path:
        sys.path.insert(0, d)
    name = os.path.splitext(os.path.basename(path))[0]
    if name in sys.modules:
        return sys.modules[name]
    return importlib.import_module(name)


def _F(a, dt):
    return np.array(a, dtype=dt, order='F', copy=True)


def _cells(active):
    """(x, y, z) 1-based of every active cell in the reference's loop order"""
    nx, ny, nz = active.shape
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            for k in range(1, nz + 1):
                if active[i - 1, j - 1, k - 1] > 0:
                    yield i, j, k


class AuxReference:
    def __init__(self, oracle_lib, math: str = 'detmath'):
        self.ref = load_aux()
        self.lib = oracle_lib
        if math == 'detmath':
            rt.use_detmath(oracle_lib)
        else:
            rt.use_libm()
        rt.UNINIT_INT = 0
        self.G = self.ref.init_globals()
        G = self.G
        G.taskid, G.numtasks, G.lgtalk, G.lgecho = 0, 1, False, False
        G.lgforcetdust, G.lgtraceheating, G.lgqheat, G.lgphotoelectric = False, False, False, False

    # ------------------------------------------------------------------------------------------
    def _gas_setup(self, t, nbins, ionDen, elemAbun, abIndex, Hden, active, contBoltz1, gauntFF1, Ne, Te,
                   bremsXSecP):
        G, ref = self.G, self.ref
        nR = Hden.shape[0]
        G.lggas = True
        G.nbins, G.nstages = int(nbins), int(t.nstages)
        G.lgelementon = rt.wrap(np.asarray(t.lgElementOn) != 0)
        G.elementxref = rt.wrap(_F(t.elementXref, np.int64))
        G.xsecarray = rt.wrap(_F(t.xSecArray, np.float32))
        G.bremsxsecp = int(bremsXSecP)
        for name, v in (('hlevxsecp', t.HlevXSecP1), ('hlevnup', t.HlevNuP1), ('heisingxsecp', t.HeISingXSecP1),
                        ('heilevnup', t.HeIlevNuP1), ('heiixsecp', t.HeIIXSecP1), ('heiilevnup', t.HeIIlevNuP1)):
            arr = getattr(G, name)
            if arr is None:
                arr = rt.alloc('i', [(1, 10)])
                setattr(G, name, arr)
            arr[1] = int(v)
        G.elementp = rt.wrap(_F(t.elementP, np.int64))
        G.nshells = rt.wrap(_F(t.nShells, np.int64))
        G.nuarray = rt.wrap(np.linspace(0.1, 10.0, nbins).astype(np.float32))
        G.iondenused = rt.alloc('r', [(1, ionDen.shape[1]), (1, t.nstages)])
        g = ref.T_grid_type()
        g.nx, g.ny, g.nz = active.shape
        g.ncells = nR - 1
        g.active = rt.wrap(_F(active, np.int64))
        ab3 = np.zeros(active.shape, np.int64, order='F')
        ab3[active > 0] = abIndex[active[active > 0]]
        g.abfileindex = rt.wrap(ab3)
        g.ionden = rt.wrap(_F(ionDen, np.float32), (0, 1, 1))
        g.elemabun = rt.wrap(_F(elemAbun, np.float32))
        g.hden = rt.wrap(_F(Hden, np.float32), (0,))
        g.ne = rt.wrap(_F(Ne, np.float32), (0,))
        g.te = rt.wrap(_F(Te, np.float32), (0,))
        op = np.zeros((nR, nbins), np.float32, order='F')
        g.opacity = rt.wrap(op, (0, 1))

        def boltgaunt():                      # stand-in for BoltGaunt: bin 1 only matters
            G.contboltz[1] = np.float32(contBoltz1)
            G.gauntff[1] = np.float32(gauntFF1)
        rt.externs['boltgaunt'] = boltgaunt
        return g, op

    def gas_opacity(self, t, nbins, ionDen, elemAbun, abIndex, Hden, active, contBoltz1, gauntFF1, Ne, Te,
                    bremsXSecP):
        """opacity(0:nCells, nbins) from ionizationDriver called cell by cell; also returns
        FFOpacity(1) per cell (the value the oracle takes as its ff1 input).  abIndex is per
        cell (0:nCells); the reference indexes abFileIndex(ix,iy,iz), built here from `active`."""
        g, op = self._gas_setup(t, nbins, ionDen, elemAbun, abIndex, Hden, active, contBoltz1, gauntFF1, Ne, Te,
                                bremsXSecP)
        ff1 = np.zeros(Hden.shape[0], np.float32)
        with np.errstate(all='ignore'):
            for (i, j, k) in _cells(active):
                self.ref.p_ionizationdriver(g, i, j, k)
                ff1[active[i - 1, j - 1, k - 1]] = self.G.ffopacity[1]
        return op, ff1

    def opacity_block(self, t, nbins, ionDen, elemAbun, abIndex, Hden, active, contBoltz1, gauntFF1, Ne, Te,
                      bremsXSecP, dust=None, dust_model=None):
        """the opacity block of iterateMC (iteration_mod.f90:106-230) on one grid: ionizationDriver
        over all cells, the single-rank all-reduce, and the dust contribution.  Returns
        (opacity, scaOpac, absOpac), each (0:nCells, nbins); the last two None without dust."""
        G = self.G
        g, op = self._gas_setup(t, nbins, ionDen, elemAbun, abIndex, Hden, active, contBoltz1, gauntFF1, Ne, Te,
                                bremsXSecP)
        G.ngrids, G.lg2d, G.lgequivalenttau, G.niteratemc = 1, False, False, 1
        G.lgdust = dust is not None
        sca = ab = None
        if dust is not None:
            nR = Hden.shape[0]
            G.lgmultidustchemistry = bool(dust_model.lgMultiDustChemistry)
            G.nsizes = int(dust_model.nSizes)
            G.nspeciespart = rt.wrap(_F(dust_model.nSpeciesPart, np.int64))
            G.dustcompoint = rt.wrap(_F(dust_model.dustComPoint, np.int64))
            G.grainabun = rt.wrap(_F(dust_model.grainAbun, np.float32))
            G.tdustsublime = rt.wrap(_F(dust_model.TdustSublime, np.float32))
            G.grainweight = rt.wrap(_F(dust['grainWeight'], np.float32))
            G.dustscaxsecp = rt.wrap(_F(dust['dustScaXsecP'], np.int64))
            G.dustabsxsecp = rt.wrap(_F(dust['dustAbsXsecP'], np.int64))
            g.ndust = rt.wrap(_F(dust['Ndust'], np.float32), (0,))
            g.tdust = rt.wrap(_F(dust['Tdust'], np.float32), (0, 0, 0))
            if dust.get('dustAbunIndex') is not None:
                g.dustabunindex = rt.wrap(_F(dust['dustAbunIndex'], np.int64), (0,))
            sca = np.zeros((nR, nbins), np.float32, order='F')
            ab = np.zeros((nR, nbins), np.float32, order='F')
            g.scaopac, g.absopac = rt.wrap(sca, (0, 1)), rt.wrap(ab, (0, 1))
        grids = np.empty(1, dtype=object)
        grids[0] = g
        with np.errstate(all='ignore'):
            self.ref.p_opacity_block(rt.wrap(grids))
        return op, sca, ab

    # ------------------------------------------------------------------------------------------
    def photo(self, t, nbins, nuArray, active, J, Jdif, ionDen, elemAbun, abIndex):
        """per cell: nPhotoSte/nPhotoDif(30, nstages) of updateCell and heatSte/heatDif of
        thermBalance.  J/Jdif are the host-scaled estimators (0:nCells, nbins); Jdif None = not
        debug mode.  Also returns outShell(30, nstages) from getOuterShell for the band list."""
        G, ref = self.G, self.ref
        nR = J.shape[0]
        G.lggas, G.lgdebug = True, Jdif is not None
        G.nbins, G.nstages = int(nbins), int(t.nstages)
        G.lgelementon = rt.wrap(np.asarray(t.lgElementOn) != 0)
        G.elementxref = rt.wrap(_F(t.elementXref, np.int64))
        G.xsecarray = rt.wrap(_F(t.xSecArray, np.float32))
        G.nuarray = rt.wrap(_F(nuArray, np.float32))
        for name, v in (('hlevxsecp', t.HlevXSecP1), ('hlevnup', t.HlevNuP1), ('heisingxsecp', t.HeISingXSecP1),
                        ('heilevnup', t.HeIlevNuP1), ('heiixsecp', t.HeIIXSecP1), ('heiilevnup', t.HeIIlevNuP1)):
            arr = getattr(G, name)
            if arr is None:
                arr = rt.alloc('i', [(1, 10)])
                setattr(G, name, arr)
            arr[1] = int(v)
        G.elementp = rt.wrap(_F(t.elementP, np.int64))
        G.iondenused = rt.alloc('r', [(1, ionDen.shape[1]), (1, t.nstages)])
        g = ref.T_grid_type()
        g.nx, g.ny, g.nz = active.shape
        g.ncells = nR - 1
        g.active = rt.wrap(_F(active, np.int64))
        ab3 = np.zeros(active.shape, np.int64, order='F')
        ab3[active > 0] = abIndex[active[active > 0]]
        g.abfileindex = rt.wrap(ab3)
        g.elemabun = rt.wrap(_F(elemAbun, np.float32))
        g.jste = rt.wrap(_F(J, np.float32), (0, 1))
        g.jdif = rt.wrap(_F(Jdif, np.float32), (0, 1)) if Jdif is not None else None
        ns = int(t.nstages)
        res = dict(nPhotoSte=np.zeros((nR, 30, ns), np.float32), nPhotoDif=np.zeros((nR, 30, ns), np.float32),
                   heatSte=np.zeros(nR, np.float32), heatDif=np.zeros(nR, np.float32))
        oste, odif = rt.alloc('r', [(1, 30), (1, ns)]), rt.alloc('r', [(1, 30), (1, ns)])
        with np.errstate(all='ignore'):
            for (i, j, k) in _cells(active):
                c = int(active[i - 1, j - 1, k - 1])
                G.iondenused.setall(rt.wrap(_F(ionDen[c], np.float32)))
                ref.p_photo_rates(g, i, j, k, oste, odif)
                res['nPhotoSte'][c], res['nPhotoDif'][c] = oste.a, odif.a
                hs, hd = ref.p_photo_heat(g, i, j, k, rt.ZERO32, rt.ZERO32)
                res['heatSte'][c], res['heatDif'][c] = hs, hd
            shell = np.zeros((30, ns), np.int64)
            for el in range(3, 31):
                if t.lgElementOn[el - 1]:
                    for ion in range(1, min(el, ns - 1) + 1):
                        shell[el - 1, ion - 1] = ref.p_getoutershell(el, el - ion + 1, 0, 0, 0)[0]
        res['outShell'] = shell
        return res

    # ------------------------------------------------------------------------------------------
    def write_sed(self, model, widFlx, escaped):
        """writeSED on host-scaled escapedPackets arrays (one (0:nCells,0:nbins,0:nAngleBins)
        float32 array per grid).  Returns (SED (nbins, nAngleBins+1) as written, totalE)."""
        G, ref = self.G, self.ref
        at = model.angle_tables()
        G.nbins, G.ngrids, G.nanglebins = int(model.nbins), int(model.nGrids), int(model.nAngleBins)
        G.lgecho, G.lgnosource, G.lgequivalenttau, G.niteratemc = False, False, False, 1
        G.lgsymmetricxyz = bool(model.lgSymmetricXYZ)
        G.dtheta, G.dphi = np.float32(at['dTheta']), np.float32(at['dPhi'])
        G.viewpointtheta = rt.wrap(_F(at['viewPointTheta'], np.float32), (0,))
        G.viewpointphi = rt.wrap(_F(at['viewPointPhi'], np.float32), (0,))
        G.nuarray = rt.wrap(_F(model.nuArray, np.float32))
        G.widflx = rt.wrap(_F(widFlx, np.float32))
        G.radio4p9ghzp = 1
        G.lstar = rt.wrap(np.zeros(1, np.float32))
        G.nphotons = rt.wrap(np.zeros(1, np.int64))
        grids = np.empty(model.nGrids, dtype=object)
        for i, (g, e) in enumerate(zip(model.grids, escaped)):
            t = ref.T_grid_type()
            t.nx, t.ny, t.nz, t.ncells = g.nx, g.ny, g.nz, int(g.nCells)
            t.escapedpackets = rt.wrap(_F(e, np.float32), (0, 0, 0))
            grids[i] = t
        rt.io_log.clear()
        with np.errstate(all='ignore'):
            ref.p_writesed(rt.wrap(grids))
        recs = rt.io_log.get(16, [])
        rows = [r for r in recs if len(r) == model.nAngleBins + 3 and not isinstance(r[0], str)]
        assert len(rows) == model.nbins, (len(rows), model.nbins)
        sed = np.array([[r[2 + a] for a in range(model.nAngleBins + 1)] for r in rows], np.float32)
        tot = [r for r in recs if isinstance(r[0], str) and r[0].startswith('Total energy')][0][1]
        return sed, np.float32(tot), rows

    # ------------------------------------------------------------------------------------------
    def bhmie(self, x, refrel):
        """BHmie(x, refrel, qext, qsca, ggsca) (ph_mod.f90:1600-1757): Mie efficiencies of a sphere,
        REAL size parameter and COMPLEX refractive index in, three REALs out."""
        with np.errstate(all='ignore'):
            qext, qsca, gg = self.ref.p_bhmie(np.float32(x), np.complex64(refrel), np.float32(0), np.float32(0),
                                              np.float32(0))
        return np.float32(qext), np.float32(qsca), np.float32(gg)

    def get_qs(self, Ere, Eim, radius, nu, scattering=True):
        """getQs (ph_mod.f90:1548-1586): Qabs, Qsca, <cos> as (nSizes, nbins) for optical constants
        already mapped on the frequency mesh."""
        G, ref = self.G, self.ref
        nS, nb = len(radius), len(nu)
        G.nsizes, G.nbins = int(nS), int(nb)
        G.grainradius = rt.wrap(_F(radius, np.float32))
        G.nuarray = rt.wrap(_F(nu, np.float32))
        G.lgdustscattering = bool(scattering)
        out = [rt.wrap(np.zeros((nS, nb), np.float32, order='F')) for _ in range(3)]
        with np.errstate(all='ignore'):
            ref.p_getqs(rt.wrap(_F(Ere, np.float32)), rt.wrap(_F(Eim, np.float32)), *out)
        return tuple(o.a.copy() for o in out)

    def dust_xsec_assembly(self, Qsca, Qabs, gCos, radius, weight, abun, nbins):
        """The tail of makeDustXsec's component loop (ph_mod.f90:1456-1538) for ONE dust component:
        efficiencies (nSpecies, nSizes, nbins) -> the dust part of xSecArray, dustScaXsecP /
        dustAbsXsecP (0:nSpecies, nSizes), gSca, absOpacSpecies."""
        G, ref = self.G, self.ref
        nSp, nSz = Qsca.shape[0], Qsca.shape[1]
        G.nbins, G.nsizes, G.nspecies, G.ndustcomponents = int(nbins), int(nSz), int(nSp), 1
        G.nspeciespart = rt.wrap(np.array([nSp], np.int64))
        G.dustcompoint = rt.wrap(np.array([1], np.int64))
        G.grainradius = rt.wrap(_F(radius, np.float32))
        G.grainweight = rt.wrap(_F(weight, np.float32))
        G.grainabun = rt.wrap(_F(np.asarray(abun, np.float32).reshape(1, nSp), np.float32))
        G.xsecarraytemp = rt.wrap(np.zeros(2 * nbins * (nSp + 1) * nSz + 8, np.float32))
        G.xsectop = 0
        G.dustscaxsecp = rt.wrap(np.full((nSp + 1, nSz), -1, np.int64, order='F'), (0, 1))
        G.dustabsxsecp = rt.wrap(np.full((nSp + 1, nSz), -1, np.int64, order='F'), (0, 1))
        G.absopacspecies = rt.wrap(np.zeros((nSp, nbins), np.float32, order='F'))
        G.gsca = rt.wrap(np.zeros(nbins, np.float32))

        def pad(q):                      # the host arrays carry an unused size index 0
            a = np.zeros((nSp, nSz + 1, nbins), np.float32, order='F')
            a[:, 1:, :] = q
            return rt.wrap(a, (1, 0, 1))
        with np.errstate(all='ignore'):
            ref.p_dust_xsec_assembly(1, pad(Qsca), pad(Qabs), pad(gCos))
        return dict(xSecArray=G.xsecarraytemp.a[:int(G.xsectop)].copy(), xSecTop=int(G.xsectop),
                    dustScaXsecP=G.dustscaxsecp.a.astype(np.int32), dustAbsXsecP=G.dustabsxsecp.a.astype(np.int32),
                    gSca=G.gsca.a.copy(), absOpacSpecies=G.absopacspecies.a.copy())

    def stellar_cdf(self, T, nu, widFlx):
        """inSpectrumErg(i) = getFlux(nuArray(i), T, 'blackbody') (setContinuum, continuum_mod.f90:112)
        followed by setProbDen(1) (:418-474).  Returns (getFlux values, inSpectrumProbDen(1, :))."""
        G, ref = self.G, self.ref
        nb = len(nu)
        G.nbins = int(nb)
        G.nuarray = rt.wrap(_F(nu, np.float32))
        G.widflx = rt.wrap(_F(widFlx, np.float32))
        G.lymanp = 1
        shape = 'blackbody'.ljust(50)
        flux = np.array([ref.p_getflux(np.float32(v), np.float32(T), shape) for v in nu], np.float32)
        G.inspectrumerg = rt.wrap(flux.astype(np.float64))
        G.inspectrumphot = rt.wrap(np.zeros(nb, np.float64))
        G.inspectrumprobden = rt.wrap(np.zeros((2, nb), np.float32, order='F'), (0, 1))
        G.contshape = rt.wrap(np.array([shape, shape], dtype=object), (0,))
        with np.errstate(all='ignore'):
            ref.p_setprobden(1)
        return flux, G.inspectrumprobden.a[1, :].copy()

    def dust_emission_int(self, xSec, absP, nu, widFlx, nTemps):
        """dustEmissionInt (dust_mod.f90:145-181, statement-range slice): dustEmIntegral
        (nSpecies, nSizes, nTemps) from xSecArray and the 1-based pointers absP(nSpecies, nSizes).
        nTemps is a parameter of the reference (3000); the harness may shrink it (Python loops)."""
        G, ref = self.G, self.ref
        nSp, nSz = absP.shape
        G.nspecies, G.nsizes, G.nbins, G.ntemps = int(nSp), int(nSz), int(len(nu)), int(nTemps)
        G.nuarray = rt.wrap(_F(nu, np.float32))
        G.widflx = rt.wrap(_F(widFlx, np.float32))
        G.xsecarray = rt.wrap(_F(xSec, np.float32))
        G.dustabsxsecp = rt.wrap(_F(np.vstack([np.zeros((1, nSz), np.int64), np.asarray(absP, np.int64)]), np.int64), (0, 1))
        with np.errstate(all='ignore'):
            ref.p_dust_emission_int()
        return G.dustemintegral.a.copy()

    def grain_weights(self, radius, weight):
        """makeDustXsec's normalisation of the size distribution (ph_mod.f90:986-1012, slice)."""
        G = self.G
        G.nsizes = int(len(radius))
        G.grainradius = rt.wrap(_F(radius, np.float32))
        G.grainweight = rt.wrap(_F(weight, np.float32).copy())
        self.ref.p_grain_weights()
        return G.grainweight.a.copy()

    def set_star_position(self, grids, relative):
        """setStarPosition(xA, yA, zA, grid) (grid_mod.f90:3569-3648) for stars at `relative`
        positions (units of the mother grid's last axis points).  Returns (positions, starIndeces)."""
        G, ref = self.G, self.ref
        G.nstars = len(relative)
        sp = np.empty(len(relative), dtype=object)
        for i, r in enumerate(relative):
            v = ref.T_vector()
            v.x, v.y, v.z = (np.float32(c) for c in r)
            sp[i] = v
        G.starposition = rt.wrap(sp)
        G.starindeces = None
        gs = np.empty(len(grids), dtype=object)
        for i, g in enumerate(grids):
            t = ref.T_grid_type()
            t.nx, t.ny, t.nz, t.ncells = g.nx, g.ny, g.nz, int(g.nCells)
            t.xaxis, t.yaxis, t.zaxis = (rt.wrap(_F(a, np.float32)) for a in (g.xAxis, g.yAxis, g.zAxis))
            t.active = rt.wrap(_F(g.active, np.int32))
            gs[i] = t
        g1 = grids[0]
        ref.p_setstarposition(rt.wrap(_F(g1.xAxis, np.float32)), rt.wrap(_F(g1.yAxis, np.float32)),
                              rt.wrap(_F(g1.zAxis, np.float32)), rt.wrap(gs))
        pos = np.array([[G.starposition.a[i].x, G.starposition.a[i].y, G.starposition.a[i].z] for i in range(len(relative))], np.float32)
        return pos, np.asarray(G.starindeces.a, np.int32)

    def get_volume(self, g, symmetric, cells):
        """getVolume(grid, xP, yP, zP) (grid_mod.f90:2876-2965) for a list of (xP, yP, zP)."""
        G, ref = self.G, self.ref
        G.lg1d, G.lgsymmetricxyz, G.ngrids = False, bool(symmetric), 1
        t = ref.T_grid_type()
        t.nx, t.ny, t.nz = g.nx, g.ny, g.nz
        t.xaxis, t.yaxis, t.zaxis = (rt.wrap(_F(a, np.float32)) for a in (g.xAxis, g.yAxis, g.zAxis))
        return np.array([ref.p_getvolume(t, int(x), int(y), int(z)) for x, y, z in cells], np.float32)

    def angle_tables(self, viewPointTheta, viewPointPhi, symmetric, totT=10, totP=20):
        """The angular-bin block of initCartesianGrid (grid_mod.f90:416-468, slice): dTheta, dPhi,
        totAngleBinsPhi, viewPointPtheta/Pphi and the (possibly reset) viewPointPhi."""
        G, ref = self.G, self.ref
        n = len(viewPointTheta) - 1
        G.nanglebins, G.totanglebinstheta, G.totanglebinsphi = int(n), int(totT), int(totP)
        G.lgsymmetricxyz = bool(symmetric)
        G.viewpointtheta = rt.wrap(_F(viewPointTheta, np.float32).copy(), (0,))
        G.viewpointphi = rt.wrap(_F(viewPointPhi, np.float32).copy(), (0,))
        G.viewpointptheta, G.viewpointpphi = None, None
        ref.p_angle_tables()
        return dict(dTheta=np.float32(G.dtheta), dPhi=np.float32(G.dphi), totAngleBinsPhi=int(G.totanglebinsphi),
                    viewPointPtheta=np.asarray(G.viewpointptheta.a, np.int32), viewPointPphi=np.asarray(G.viewpointpphi.a, np.int32),
                    viewPointPhi=G.viewpointphi.a.copy())

    def active_cells(self, xAxis, yAxis, zAxis, Hden3, Ndust3, lgGas, lgDust, R_in, R_out, plane=False):
        """The active-cell block of setMotherGrid (grid_mod.f90:1226-1294, slice; grid%active = 1
        beforehand as at :977).  Returns (active (nx,ny,nz) int32, nCells)."""
        G, ref = self.G, self.ref
        G.lg1d, G.lgplaneionization, G.lgmultidustchemistry = False, bool(plane), False
        G.lggas, G.lgdust = bool(lgGas), bool(lgDust)
        G.r_in, G.r_out = np.float32(R_in), np.float32(R_out)
        t = ref.T_grid_type()
        t.nx, t.ny, t.nz = len(xAxis), len(yAxis), len(zAxis)
        t.xaxis, t.yaxis, t.zaxis = (rt.wrap(_F(a, np.float32)) for a in (xAxis, yAxis, zAxis))
        t.active = rt.wrap(np.zeros((t.nx, t.ny, t.nz), np.int64, order='F'))
        ref.p_active_cells(t, rt.wrap(_F(Hden3, np.float32)), rt.wrap(_F(Ndust3, np.float32)), int(t.ny))
        return np.asarray(t.active.a, np.int32), int(t.ncells)

    def _grids_for_fill(self, grids):
        gs = np.empty(len(grids), dtype=object)
        for i, g in enumerate(grids):
            t = self.ref.T_grid_type()
            t.nx, t.ny, t.nz = g.nx, g.ny, g.nz
            t.xaxis, t.yaxis, t.zaxis = (rt.wrap(_F(a, np.float32).copy()) for a in (g.xAxis, g.yAxis, g.zAxis))
            t.active = rt.wrap(_F(g.active, np.int64).copy())
            gs[i] = t
        return gs

    def fill_axes(self, n, R, symmetric):
        """The automatic axes of fillGrid (grid_mod.f90:530-601, slice) for an n^3 mother grid with
        edges Rnx = Rny = Rnz = R."""
        G, ref = self.G, self.ref
        G.lgdfile, G.lg1d, G.lgsymmetricxyz = False, False, bool(symmetric)
        G.rnx = G.rny = G.rnz = np.float32(R)
        t = ref.T_grid_type()
        t.nx = t.ny = t.nz = int(n)
        t.xaxis, t.yaxis, t.zaxis = (rt.wrap(np.zeros(n, np.float32)) for _ in range(3))
        gs = np.empty(1, dtype=object)
        gs[0] = t
        ref.p_fill_axes(rt.wrap(gs))
        return t.xaxis.a.copy(), t.yaxis.a.copy(), t.zaxis.a.copy()

    def fill_mask(self, grids, symmetric):
        """geoCorr and the masking of cells inside other grids (grid_mod.f90:807-816, 835-889, slice).
        Returns ([active per grid], [(geoCorrX, Y, Z) per grid])."""
        G, ref = self.G, self.ref
        G.ngrids, G.lg1d, G.lgsymmetricxyz = len(grids), False, bool(symmetric)
        gs = self._grids_for_fill(grids)
        ref.p_fill_mask(rt.wrap(gs))
        return ([np.asarray(t.active.a, np.int32) for t in gs],
                [(np.float32(t.geocorrx), np.float32(t.geocorry), np.float32(t.geocorrz)) for t in gs])

    def linear_map(self, y, x, x_new):
        """linearMap (interpolation_mod.f90:86-106)."""
        out = rt.wrap(np.zeros(len(x_new), np.float32))
        self.ref.p_linearmap(rt.wrap(_F(y, np.float32)), rt.wrap(_F(x, np.float32)), int(len(x)), out,
                             rt.wrap(_F(x_new, np.float32)), int(len(x_new)))
        return out.a.copy()

    # ------------------------------------------------------------------------------------------
    def write_tau_nu(self, model):
        """writeTauNu(grid) (output_mod.f90:2384-2505; integratePathTauNu, pathIntegration_mod.f90:
        241-470) on a single-grid model with grid%opacity set.  Returns the three outTau(1:nbins)
        arrays (directions 1,0,0 / 0,0,1 / 0,1,0) and the lambda column, as written to unit 73."""
        G, ref = self.G, self.ref
        g = model.grids[0]
        G.nbins, G.ngrids = int(model.nbins), 1
        G.lgsymmetricxyz = bool(model.lgSymmetricXYZ)
        G.r_out = np.float32(model.R_out)
        G.nuarray = rt.wrap(_F(model.nuArray, np.float32))
        t = ref.T_grid_type()
        t.nx, t.ny, t.nz, t.ncells = g.nx, g.ny, g.nz, int(g.nCells)
        t.xaxis, t.yaxis, t.zaxis = (rt.wrap(_F(a, np.float32)) for a in (g.xAxis, g.yAxis, g.zAxis))
        t.active = rt.wrap(_F(g.active, np.int32))
        t.opacity = rt.wrap(_F(g.opacity, np.float32), (0, 1))
        grids = np.empty(1, dtype=object)
        grids[0] = t
        rt.io_log.clear()
        with np.errstate(all='ignore'):
            ref.p_writetaunu(rt.wrap(grids))
        rows = [r for r in rt.io_log.get(73, []) if len(r) == 2 and not isinstance(r[0], str)]
        assert len(rows) == 3 * model.nbins, (len(rows), model.nbins)
        a = np.array(rows, np.float32).reshape(3, model.nbins, 2)
        return [a[k, :, 1].copy() for k in range(3)], a[0, :, 0].copy()

    # ------------------------------------------------------------------------------------------
    def write_cont_cube(self, model, escaped, freq1, freq2, origin=(1, 1, 1)):
        """writeContCube(grid, freq1, freq2) (output_mod.f90:2722-2806) on host-scaled
        escapedPackets arrays.  Returns the unit-19 records [(iG, ix, iy, iz, contI(0:nAngleBins))]."""
        G, ref = self.G, self.ref
        at = model.angle_tables()
        G.nbins, G.ngrids, G.nanglebins = int(model.nbins), int(model.nGrids), int(model.nAngleBins)
        G.dtheta, G.dphi = np.float32(at['dTheta']), np.float32(at['dPhi'])
        G.viewpointtheta = rt.wrap(_F(at['viewPointTheta'], np.float32), (0,))
        G.viewpointphi = rt.wrap(_F(at['viewPointPhi'], np.float32), (0,))
        G.nuarray = rt.wrap(_F(model.nuArray, np.float32))
        G.iorigin, G.jorigin, G.korigin = (int(v) for v in origin)
        grids = np.empty(model.nGrids, dtype=object)
        for i, (g, e) in enumerate(zip(model.grids, escaped)):
            t = ref.T_grid_type()
            t.nx, t.ny, t.nz, t.ncells = g.nx, g.ny, g.nz, int(g.nCells)
            t.active = rt.wrap(_F(g.active, np.int32))
            t.escapedpackets = rt.wrap(_F(e, np.float32), (0, 0, 0))
            grids[i] = t
        rt.io_log.clear()
        with np.errstate(all='ignore'):
            ref.p_writecontcube(rt.wrap(grids), np.float32(freq1), np.float32(freq2))
        recs = rt.io_log.get(19, [])
        return [r for r in recs if len(r) == model.nAngleBins + 5 and not isinstance(r[0], str)]

    # ------------------------------------------------------------------------------------------
    def write_grid(self, model, rp, state):
        """writeGrid on a Model + checkpoint.RunParams + per-grid state arrays (dict with lists
        lgConverged, lgBlack, and for gas Te, Ne, ionDen, abFileIndex plus lgElementOn/elementXref;
        star: contShape, TStellar, LStar, nPhotons, spID, tStep; lgMultiChemistry, totalDustMass).
        Returns {unit: [record tuples]} (21 grid0, 20 grid1, 30 grid2, 50 dustGrid, 42 photoSource,
        40 grid3)."""
        G, ref = self.G, self.ref
        m = model
        G.ngrids, G.nbins, G.nstars = int(m.nGrids), int(m.nbins), int(m.nStars)
        G.lggas, G.lgdust, G.lg2d = bool(m.lgGas), bool(m.lgDust), bool(rp.lg2D)
        G.lgmultichemistry = bool(state.get('lgMultiChemistry', False))
        G.r_out = np.float32(m.R_out)
        G.nsizes, G.nspeciesmax, G.nspecies = int(m.nSizes), int(m.nSpeciesMax), int(rp.nSpecies)
        G.totaldustmass = np.float32(state.get('totalDustMass', 0.0))
        G.nstages = int(rp.nstages)
        if m.lgGas:
            G.lgelementon = rt.wrap(np.asarray(state['lgElementOn']) != 0)
            G.elementxref = rt.wrap(_F(state['elementXref'], np.int64))
        # photoSource.out
        n = m.nStars
        strs = lambda v, w: rt.wrap(np.array([str(x).ljust(w) for x in v], dtype=object))
        G.contshapein = strs(state['contShape'], 50)
        G.spid = strs(state['spID'], 50)
        G.tstellar = rt.wrap(_F(state['TStellar'], np.float32))
        G.lstar = rt.wrap(_F(state['LStar'], np.float32))
        G.nphotons = rt.wrap(_F(state['nPhotons'], np.int64))
        G.tstep = rt.wrap(_F(state['tStep'], np.float32))
        sp = np.empty(n, dtype=object)
        for i in range(n):
            sp[i] = ref.T_vector(*[np.float32(v) for v in m.starPosition[i]])
        G.starposition = rt.wrap(sp)
        G.pwlindex = np.float32(0.0)
        # grid3.out
        G.convwritegrid = np.float32(rp.convWriteGrid)
        G.lgautopackets, G.convincpercent = bool(rp.lgAutoPackets), np.float32(rp.convIncPercent)
        G.nphotincrease, G.maxphotons = np.float32(rp.nPhotIncrease), int(rp.maxPhotons)
        G.lgsymmetricxyz, G.lgtalk, G.lg1d = bool(m.lgSymmetricXYZ), bool(rp.lgTalk), bool(rp.lg1D)
        G.nustepsize, G.numax, G.numin = np.float32(rp.nuStepSize), np.float32(rp.nuMax), np.float32(rp.nuMin)
        G.r_in, G.xhilimit = np.float32(rp.R_in), np.float32(rp.XHIlimit)
        G.maxiteratemc, G.minconvergence = int(rp.maxIterateMC), np.float32(rp.minConvergence)
        G.lgdebug, G.lgplaneionization = bool(m.lgDebug), bool(m.lgPlaneIonization)
        G.nabcomponents = int(rp.nAbComponents)
        G.abundancefile = strs(rp.abundanceFile, 50)
        G.lgoutput, G.dxslit, G.dyslit = bool(rp.lgOutput), np.float32(rp.dxSlit), np.float32(rp.dySlit)
        G.lgdustconstant = bool(rp.lgDustConstant)
        G.lgmultidustchemistry, G.ndustcomponents = bool(m.lgMultiDustChemistry), int(rp.nDustComponents)
        G.dustspeciesfile = strs(rp.dustSpeciesFile, 50)
        G.dustfile = strs(['', rp.dustFile2], 50)
        G.lgrecombination = bool(rp.lgRecombination)
        G.nspeciespart = rt.wrap(_F(m.nSpeciesPart, np.int64))
        G.reslinestransfer, G.lgdustscattering = np.float32(rp.resLinesTransfer), bool(rp.lgDustScattering)
        G.nanglebins = int(m.nAngleBins)
        if m.nAngleBins > 0:
            G.viewpointtheta = rt.wrap(_F(m.viewPointTheta, np.float32), (0,))
            G.viewpointphi = rt.wrap(_F(m.viewPointPhi, np.float32), (0,))
        G.contcube = rt.wrap(_F(rp.contCube, np.float32))
        G.lgphotoelectric, G.lgtraceheating = bool(rp.lgPhotoelectric), bool(rp.lgTraceHeating)
        G.ldiffuse, G.tdiffuse = np.float32(rp.Ldiffuse), np.float32(rp.Tdiffuse)
        G.shapediffuse = str(rp.shapeDiffuse).ljust(50)
        G.nphotonsdiffuse, G.emittinggrid = int(rp.nPhotonsDiffuse), int(rp.emittingGrid)
        G.lgmultistars, G.lgecho = bool(m.lgMultistars), bool(rp.lgEcho)
        G.echot1, G.echot2, G.echotemp = np.float32(rp.echot1), np.float32(rp.echot2), np.float32(rp.echoTemp)
        G.lgnosource = bool(rp.lgNosource)
        grids = np.empty(m.nGrids, dtype=object)
        for i, g in enumerate(m.grids):
            t = ref.T_grid_type()
            t.nx, t.ny, t.nz, t.ncells, t.motherp = g.nx, g.ny, g.nz, int(g.nCells), int(g.motherP)
            t.xaxis, t.yaxis, t.zaxis = (rt.wrap(_F(a, np.float32)) for a in (g.xAxis, g.yAxis, g.zAxis))
            t.active = rt.wrap(_F(g.active, np.int64))
            t.lgconverged = rt.wrap(_F(state['lgConverged'][i], np.int64), (0,))
            t.lgblack = rt.wrap(_F(state['lgBlack'][i], np.int64), (0,))
            if m.lgGas:
                t.te = rt.wrap(_F(state['Te'][i], np.float32), (0,))
                t.ne = rt.wrap(_F(state['Ne'][i], np.float32), (0,))
                t.hden = rt.wrap(_F(g.Hden, np.float32), (0,))
                t.ionden = rt.wrap(_F(state['ionDen'][i], np.float32), (0, 1, 1))
                t.abfileindex = rt.wrap(_F(state['abFileIndex'][i], np.int64))
            if m.lgDust:
                t.ndust = rt.wrap(_F(g.Ndust, np.float32), (0,))
                t.tdust = rt.wrap(_F(g.Tdust, np.float32), (0, 0, 0))
                if g.dustAbunIndex is not None:
                    t.dustabunindex = rt.wrap(_F(g.dustAbunIndex, np.int64), (0,))
            grids[i] = t
        rt.io_log.clear()
        with np.errstate(all='ignore'):
            ref.p_writegrid(rt.wrap(grids))
        return {u: list(v) for u, v in rt.io_log.items()}

    # ------------------------------------------------------------------------------------------
    def _dust_globals(self, model, tables, lgDebug=False):
        G = self.G
        G.lgdust, G.lggas, G.lgdebug = True, False, bool(lgDebug)
        G.lgmultidustchemistry = bool(model.lgMultiDustChemistry)
        G.nbins, G.nsizes, G.nspeciesmax = int(model.nbins), int(model.nSizes), int(model.nSpeciesMax)
        G.nuarray = rt.wrap(_F(model.nuArray, np.float32))
        G.widflx = rt.wrap(_F(tables['widFlx'], np.float32))
        G.xsecarray = rt.wrap(_F(tables['xSecArray'], np.float32))
        G.dustabsxsecp = rt.wrap(_F(tables['dustAbsXsecP'], np.int64))
        G.grainweight = rt.wrap(_F(tables['grainWeight'], np.float32))
        G.grainabun = rt.wrap(_F(model.grainAbun, np.float32))
        G.nspeciespart = rt.wrap(_F(model.nSpeciesPart, np.int64))
        G.dustcompoint = rt.wrap(_F(model.dustComPoint, np.int64))
        G.tdustsublime = rt.wrap(_F(model.TdustSublime, np.float32))
        G.dustemintegral = rt.wrap(_F(tables['dustEmIntegral'], np.float32))
        G.convpercent, G.niteratemc = np.float32(0.0), 1

    def _grid(self, model, g):
        t = self.ref.T_grid_type()
        t.nx, t.ny, t.nz, t.ncells = g.nx, g.ny, g.nz, int(g.nCells)
        t.active = rt.wrap(_F(g.active, np.int64))
        t.dustabunindex = rt.wrap(_F(g.dustAbunIndex, np.int64), (0,)) if g.dustAbunIndex is not None else None
        t.tdust = rt.wrap(_F(g.Tdust, np.float32), (0, 0, 0))
        return t

    def dust_pdf(self, model, g, tables):
        """dustPDF(0:nCells, nbins): emissionDriver on every cell of a dust-only model"""
        self._dust_globals(model, tables)
        t = self._grid(model, g)
        pdf = np.zeros((g.nCells + 1, model.nbins), np.float32, order='F')
        t.dustpdf = rt.wrap(pdf, (0, 1))
        grids = np.empty(1, dtype=object)
        grids[0] = t
        grids = rt.wrap(grids)
        with np.errstate(all='ignore'):
            for (i, j, k) in _cells(g.active):
                self.ref.p_emissiondriver(grids, i, j, k, 1)
        return pdf

    def dust_update(self, model, g, tables, Jste, XHILimit, Jdif=None):
        """(Tdust, lgConverged) after updateCell on every cell of a dust-only model; Jste is the
        host-scaled estimator the reference holds at that point"""
        self._dust_globals(model, tables, lgDebug=Jdif is not None)
        G = self.G
        G.xhilimit = np.float32(XHILimit)
        t = self._grid(model, g)
        t.jste = rt.wrap(_F(Jste, np.float32), (0, 1))
        t.jdif = rt.wrap(_F(Jdif, np.float32), (0, 1)) if Jdif is not None else None
        t.lgblack = rt.alloc('i', [(0, g.nCells)])
        t.lgconverged = rt.alloc('i', [(0, g.nCells)])
        G.tdusttemp = rt.alloc('r', [(0, model.nSpeciesMax), (0, model.nSizes), (0, g.nCells)])
        with np.errstate(all='ignore'):
            for (i, j, k) in _cells(g.active):
                self.ref.p_updatecell(t, i, j, k)
        return t.tdust.a, t.lgconverged.a.astype(np.int32)


# ---- gas-side input data (tests/test_reference_pin_gas.py) ------------------------------------
def _gas_nu_mesh(self, ph1, ph2, lgElementOn, nstages, nbins, nuMin, nuMax):
    """The ionisation thresholds + the gas-only frequency mesh + widFlx of initCartesianGrid
    (grid_mod.f90:132-175, 215-258, 333-338; statement-range slice) with getOuterShell and sortUp.
    ph1 / ph2 are what phInit reads from data/ph1.dat, data/ph2.dat (the translator has no READ:
    the harness parses the files).  Returns (nuArray, widFlx, ionEdge(1:nEdges) sorted)."""
    G, ref = self.G, self.ref
    G.lggas, G.lgdust = True, False
    G.nbins, G.nstages = int(nbins), int(nstages)
    G.numin, G.numax = np.float32(nuMin), np.float32(nuMax)
    G.ph1 = rt.wrap(_F(ph1, np.float32))
    G.ph2 = rt.wrap(_F(ph2, np.float32))
    G.lgelementon = rt.wrap(np.asarray(lgElementOn) != 0)
    G.ionedge = rt.wrap(np.zeros(450, np.float32))
    with np.errstate(all='ignore'):
        ref.p_gas_nu_mesh()
    edges = G.ionedge.a.copy()
    return G.nuarray.a.copy(), G.widflx.a.copy(), edges


def _gas_xsec(self, nu, ph1, ph2, lgElementOn, nstages):
    """setPointers (makeHydro, setShells, limitShell) + initXSecArray (phFitEl, phFitHIon,
    powLawXSec, makeOpacity) of the reference on the mesh `nu` (hydro_mod.f90:56-85,208-475;
    ph_mod.f90:204-402,477-606,712-803).  The file readers on the way (makeCollIonData,
    makeAugerData, readHeIRecLines) are no-ops.  Returns the cross-section stack and every pointer."""
    G, ref = self.G, self.ref
    G.lggas, G.lgdust, G.lgcompton = True, False, False
    G.nbins, G.nstages = int(len(nu)), int(nstages)
    G.nuarray = rt.wrap(_F(nu, np.float32))
    G.ph1 = rt.wrap(_F(ph1, np.float32))
    G.ph2 = rt.wrap(_F(ph2, np.float32))
    G.level = rt.wrap(np.array([0, 0, 1, 0, 1, 2, 0], np.int64))
    G.ninn = rt.wrap(np.array([0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 3, 3, 3, 3, 3, 3, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5], np.int64))
    G.ntot = rt.wrap(np.array([1, 1, 2, 2, 3, 3, 3, 3, 3, 3, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 7, 7], np.int64))
    G.lgelementon = rt.wrap(np.asarray(lgElementOn) != 0)
    G.elementp = rt.wrap(np.zeros((30, 30, 7, 3), np.int64, order='F'))
    G.nshells = rt.wrap(np.zeros((30, 30), np.int64, order='F'))
    for name in ('makecolliondata', 'makeaugerdata', 'readheireclines', 'setcompton', 'makedustxsec', 'phinit'):
        rt.externs[name] = lambda: None
    with np.errstate(all='ignore'):
        ref.p_initxsecarray()
    return dict(xSecArray=G.xsecarray.a.copy(), xSecTop=int(G.xsectop), elementP=G.elementp.a.copy(), nShells=G.nshells.a.copy(),
                HlevNuP=G.hlevnup.a.copy(), HeIlevNuP=G.heilevnup.a.copy(), HeIIlevNuP=G.heiilevnup.a.copy(),
                HlevXSecP=G.hlevxsecp.a.copy(), HeISingXSecP=G.heisingxsecp.a.copy(), HeIIXSecP=G.heiixsecp.a.copy(),
                bremsXSecP=int(G.bremsxsecp), KshellLimitP=int(G.kshelllimitp), HlevEn=G.hleven.a.copy(),
                HeIlevEn=G.heileven.a.copy(), HeIIlevEn=G.heiileven.a.copy())


AuxReference.gas_nu_mesh = _gas_nu_mesh
AuxReference.gas_xsec = _gas_xsec


def _initial_ions(self, active, lgElementOn, elementXref, nstages):
    """The initial ionisation state of setMotherGrid (grid_mod.f90:1564-1607; statement-range
    slice): ionDen(0:nCells, nElementsUsed, nstages) and Ne for every active cell."""
    G, ref = self.G, self.ref
    G.lggas, G.nstages = True, int(nstages)
    G.lgelementon = rt.wrap(np.asarray(lgElementOn) != 0)
    G.elementxref = rt.wrap(_F(elementXref, np.int64))
    nUsed = int((np.asarray(lgElementOn) != 0).sum())
    nCells = int(active.max())
    g = ref.T_grid_type()
    g.nx, g.ny, g.nz = active.shape
    g.ncells = nCells
    g.active = rt.wrap(_F(active, np.int64))
    ion = np.zeros((nCells + 1, nUsed, nstages), np.float32, order='F')
    g.ionden = rt.wrap(ion, (0, 1, 1))
    hden = np.full(nCells + 1, 100.0, np.float32)
    ne = np.zeros(nCells + 1, np.float32)
    g.hden, g.ne = rt.wrap(hden, (0,)), rt.wrap(ne, (0,))
    ref.p_initial_ions(g, int(active.shape[1]))
    return ion, ne


AuxReference.initial_ions = _initial_ions


def _sub_grid_read(self, list_text, file_text, mother_axes, nsub, lgGas, lgDust, R_in, R_out, symmetric=True):
    """setSubGrids' own reading code (grid_mod.f90:1847-1857, 2028-2232; statement-range slice) on the text
    of a sub-grid list (unit 71) and of ONE sub-grid density file (unit 72): list-directed READs, axes
    rescaled from normalised coordinates, the routine's own `insanity` stops, the active-cell rule.
    Returns dict(xAxis, yAxis, zAxis, active, nCells, motherP, Hden3, Ndust3) -- or raises what the
    reference raises (rt.FortranStop for `print*; stop`, rt.FortranEOF for a READ past the end)."""
    G, ref = self.G, self.ref
    G.ngrids, G.lggas, G.lgdust = 2, bool(lgGas), bool(lgDust)
    G.lgmultichemistry, G.lgmultidustchemistry, G.lg1d, G.lgecho, G.lgplaneionization = False, False, False, False, False
    G.r_in, G.r_out = np.float32(R_in), np.float32(R_out)
    G.taskid = 0
    rt.bind_unit(71, list_text)
    rt.bind_unit(72, file_text)
    grids = np.empty(2, dtype=object)
    for i, (nx, ny, nz) in enumerate([tuple(len(a) for a in mother_axes), nsub]):
        g = ref.T_grid_type()
        g.nx, g.ny, g.nz = int(nx), int(ny), int(nz)
        g.xaxis = rt.wrap(np.zeros(nx, np.float32))
        g.yaxis = rt.wrap(np.zeros(ny, np.float32))
        g.zaxis = rt.wrap(np.zeros(nz, np.float32))
        g.active = rt.wrap(np.zeros((nx, ny, nz), np.int64, order='F'))
        g.elemabun = rt.wrap(np.zeros((1, 30), np.float32, order='F'))
        g.ncells = 0
        grids[i] = g
    for a, src in zip((grids[0].xaxis, grids[0].yaxis, grids[0].zaxis), mother_axes):
        a.a[:] = np.asarray(src, np.float32)
    nx, ny, nz = nsub
    H = rt.wrap(np.zeros((nx, ny, nz), np.float32, order='F'))
    Nd = rt.wrap(np.zeros((nx, ny, nz), np.float32, order='F'))
    with np.errstate(all='ignore'):
        ref.p_sub_grid_read(rt.wrap(grids), H, Nd)
    s = grids[1]
    return dict(xAxis=s.xaxis.a.copy(), yAxis=s.yaxis.a.copy(), zAxis=s.zaxis.a.copy(), active=s.active.a.copy(),
                nCells=int(s.ncells), motherP=int(s.motherp), Hden3=H.a.copy(), Ndust3=Nd.a.copy())


AuxReference.sub_grid_read = _sub_grid_read
