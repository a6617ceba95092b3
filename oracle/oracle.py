"""ctypes binding of the CPU oracle (oracle/libmc_oracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package (mocassin_b200)
never does.  Pinned against the reference's own code run through oracle/f90ref (see
oracle/mc_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmc_oracle.so")

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)


class OrGrid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("nCells", C.c_int32),
                ("motherP", C.c_int32), ("geoCorrX", C.c_float), ("geoCorrY", C.c_float),
                ("geoCorrZ", C.c_float), ("invLenUnit", C.c_float),
                ("xAxis", fp), ("yAxis", fp), ("zAxis", fp), ("active", ip),
                ("opacity", fp), ("scaOpac", fp), ("recPDF", fp), ("dustPDF", fp), ("linePDF", fp),
                ("totalLines", fp), ("Tdust", fp), ("dustAbunIndex", ip),
                ("Jste", fp), ("Jdif", fp), ("escapedPackets", fp), ("linePackets", fp),
                ("JsteQ", lp), ("JdifQ", lp), ("escapedQ", lp), ("linePacketsQ", lp), ("resLinePackets", ip),
                ("JsteN", ip), ("JdifN", ip)]


class OrParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nGrids", "nbins", "nAngleBins", "totAngleBinsTheta", "totAngleBinsPhi", "nLines", "nStars",
        "lgDust", "lgGas", "lgSymmetricXYZ", "lgIsotropic", "lgPlaneIonization", "lgDebug",
        "lgMultistars", "lgMultiDustChemistry", "nSpeciesMax", "nSizes", "nDustComp")] + [
        (n, C.c_float) for n in ("dTheta", "dPhi", "R_out", "ionEdge1")] + [
        ("nuArray", fp), ("gSca", fp), ("viewPointPtheta", ip), ("viewPointPphi", ip),
        ("viewPointTheta", fp), ("viewPointPhi", fp), ("starPosition", fp), ("starIndeces", ip),
        ("deltaE", fp), ("inSpectrumProbDen", fp), ("nSpeciesPart", ip), ("grainAbun", fp),
        ("dustComPoint", ip), ("TdustSublime", fp), ("planeIonDistribution", ip)]


class OrCounters(C.Structure):
    _fields_ = [("Qphot", C.c_float), ("absInt", C.c_float), ("scaInt", C.c_float)] + [
        (n, C.c_int64) for n in ("nAbs", "nSca", "trapped", "nLinePackets", "nDropped", "nSegments",
                                 "nFlights", "nEscaped", "nEarlyEscaped")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class OrOpacityIn(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nCells", "nbins", "nstages", "nElementsUsed", "nAbComp")] + [
        ("lgElementOn", ip), ("elementXref", ip), ("ionDen", fp), ("elemAbun", fp), ("Hden", fp), ("ff1", fp),
        ("abIndex", ip), ("xSecArray", fp)] + [
        (n, C.c_int32) for n in ("HlevXSecP1", "HlevNuP1", "HeISingXSecP1", "HeIlevNuP1", "HeIIXSecP1", "HeIIlevNuP1")] + [
        ("elementP", ip), ("nShells", ip)] + [
        (n, C.c_int32) for n in ("lgDust", "lgMultiDustChemistry", "nSpeciesMax", "nSizes", "nDustComp", "nSpeciesTot")] + [
        ("nSpeciesPart", ip), ("dustComPoint", ip), ("dustAbunIndex", ip), ("dustScaXsecP", ip), ("dustAbsXsecP", ip),
        ("grainAbun", fp), ("grainWeight", fp), ("TdustSublime", fp), ("Tdust", fp), ("Ndust", fp)]


def opacity(t, nbins, ionDen, elemAbun, abIndex, Hden, ff1=None, dust=None, model=None):
    """oracle_opacity on reference-form inputs; returns (opacity, scaOpac, absOpac)."""
    lib = load()
    keep = []

    def A(a, dt):
        a = np.asfortranarray(a, dtype=dt)
        keep.append(a)
        return a

    I = OrOpacityIn()
    nR = Hden.shape[0]
    I.nCells, I.nbins, I.nstages = nR - 1, nbins, t.nstages
    I.nElementsUsed, I.nAbComp = ionDen.shape[1], elemAbun.shape[0]
    I.lgElementOn = _p(A(t.lgElementOn, np.int32), ip); I.elementXref = _p(A(t.elementXref, np.int32), ip)
    I.ionDen = _p(A(ionDen, np.float32), fp); I.elemAbun = _p(A(elemAbun, np.float32), fp)
    I.Hden = _p(A(Hden, np.float32), fp); I.ff1 = _p(A(ff1, np.float32), fp) if ff1 is not None else fp()
    I.abIndex = _p(A(abIndex, np.int32), ip); I.xSecArray = _p(A(t.xSecArray, np.float32), fp)
    I.HlevXSecP1, I.HlevNuP1 = t.HlevXSecP1, t.HlevNuP1
    I.HeISingXSecP1, I.HeIlevNuP1, I.HeIIXSecP1, I.HeIIlevNuP1 = t.HeISingXSecP1, t.HeIlevNuP1, t.HeIIXSecP1, t.HeIIlevNuP1
    I.elementP = _p(A(t.elementP, np.int32), ip); I.nShells = _p(A(t.nShells, np.int32), ip)
    op = np.zeros((nR, nbins), np.float32, order="F")
    sca = ab = None
    if dust is not None:
        I.lgDust, I.lgMultiDustChemistry = 1, int(model.lgMultiDustChemistry)
        I.nSpeciesMax, I.nSizes, I.nDustComp = model.nSpeciesMax, model.nSizes, int(model.nSpeciesPart.shape[0])
        I.nSpeciesTot = int(np.asarray(dust["dustScaXsecP"]).shape[0])
        I.nSpeciesPart = _p(A(model.nSpeciesPart, np.int32), ip); I.dustComPoint = _p(A(model.dustComPoint, np.int32), ip)
        I.dustAbunIndex = _p(A(dust["dustAbunIndex"], np.int32), ip) if dust.get("dustAbunIndex") is not None else ip()
        I.dustScaXsecP = _p(A(dust["dustScaXsecP"], np.int32), ip); I.dustAbsXsecP = _p(A(dust["dustAbsXsecP"], np.int32), ip)
        I.grainAbun = _p(A(model.grainAbun, np.float32), fp); I.grainWeight = _p(A(dust["grainWeight"], np.float32), fp)
        I.TdustSublime = _p(A(model.TdustSublime, np.float32), fp); I.Tdust = _p(A(dust["Tdust"], np.float32), fp)
        I.Ndust = _p(A(dust["Ndust"], np.float32), fp)
        sca = np.zeros((nR, nbins), np.float32, order="F")
        ab = np.zeros((nR, nbins), np.float32, order="F")
    lib.oracle_opacity.argtypes = [C.POINTER(OrOpacityIn), fp, fp, fp]
    lib.oracle_opacity.restype = None
    lib.oracle_opacity(C.byref(I), _p(op, fp), _p(sca, fp), _p(ab, fp))
    return op, sca, ab


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("mc_oracle.c", "mc_oracle.h", "detmath.h", "Makefile")]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src)
    if stale:
        res = subprocess.run(["make", "-C", HERE, "-B", "libmc_oracle.so"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.oracle_transport.restype = C.c_int
        lib.oracle_transport.argtypes = [C.POINTER(OrParams), C.POINTER(OrGrid), C.c_int32, C.c_int64,
                                         C.c_int64, C.c_uint64, C.c_int32, ip, C.POINTER(OrCounters), lp, ip]
        lib.oracle_transport_mt.restype = C.c_int
        lib.oracle_transport_mt.argtypes = [C.POINTER(OrParams), C.POINTER(OrGrid), C.c_int32, C.c_int64,
                                            C.c_int64, C.c_uint64, C.c_int32, C.POINTER(OrCounters), lp]
        lib.oracle_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        lib.oracle_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int32, fp]
        lib.oracle_locate.restype = C.c_int32
        lib.oracle_locate.argtypes = [fp, C.c_int32, C.c_float]
        lib.oracle_getnu2.restype = C.c_int32
        lib.oracle_getnu2.argtypes = [fp, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint32]
        lib.oracle_random_unit_vector.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, fp]
        lib.oracle_hg.restype = C.c_int32
        lib.oracle_hg.argtypes = [C.c_float, fp, C.c_uint64, C.c_uint64, C.c_uint32, fp]
        lib.oracle_detmath.argtypes = [C.c_int32, fp, fp, C.c_int64]
        lib.oracle_escape_bins.restype = C.c_int32
        lib.oracle_escape_bins.argtypes = [C.POINTER(OrParams), fp, ip, ip]
        lib.oracle_cell_volume.restype = C.c_float
        lib.oracle_cell_volume.argtypes = [C.POINTER(OrParams), C.POINTER(OrGrid), C.c_int32, C.c_int32, C.c_int32]
        _lib = lib
    return _lib


def _p(a, typ):
    if a is None:
        return typ()
    return a.ctypes.data_as(typ)


def _F(a, dtype):
    return None if a is None else np.asfortranarray(a, dtype=dtype)


class Oracle:
    """Runs the CPU restatement on a mocassin_b200.model.Model-shaped object (duck typed:
    only attribute access, so the oracle does not import the product package)."""

    def __init__(self, model, len_unit=None, fp32_tallies: bool = True, int_tallies: bool = True,
                 count_segments: bool = False):
        self.lib = load()
        self.m = m = model
        self._keep = []
        at = m.angle_tables()
        self.at = at
        P = OrParams()
        P.nGrids, P.nbins, P.nAngleBins = m.nGrids, m.nbins, m.nAngleBins
        P.totAngleBinsTheta, P.totAngleBinsPhi = at["totAngleBinsTheta"], at["totAngleBinsPhi"]
        P.nLines, P.nStars = m.nLines, m.nStars
        P.lgDust, P.lgGas, P.lgSymmetricXYZ = int(m.lgDust), int(m.lgGas), int(m.lgSymmetricXYZ)
        P.lgIsotropic, P.lgPlaneIonization, P.lgDebug = int(m.lgIsotropic), int(m.lgPlaneIonization), int(m.lgDebug)
        P.lgMultistars, P.lgMultiDustChemistry = int(m.lgMultistars), int(m.lgMultiDustChemistry)
        P.nSpeciesMax, P.nSizes, P.nDustComp = m.nSpeciesMax, m.nSizes, int(m.nSpeciesPart.shape[0])
        P.dTheta, P.dPhi, P.R_out, P.ionEdge1 = float(at["dTheta"]), float(at["dPhi"]), float(m.R_out), float(m.ionEdge1)

        def keep(a, dtype):
            a = np.ascontiguousarray(a, dtype=dtype)
            self._keep.append(a)
            return a

        P.nuArray = _p(keep(m.nuArray, np.float32), fp)
        P.gSca = _p(keep(m.gSca, np.float32), fp) if m.gSca is not None else fp()
        P.viewPointPtheta = _p(keep(at["viewPointPtheta"], np.int32), ip)
        P.viewPointPphi = _p(keep(at["viewPointPphi"], np.int32), ip)
        P.viewPointTheta = _p(keep(at["viewPointTheta"], np.float32), fp)
        P.viewPointPhi = _p(keep(at["viewPointPhi"], np.float32), fp)
        P.starPosition = _p(keep(np.asarray(m.starPosition).reshape(-1), np.float32), fp)
        P.starIndeces = _p(keep(np.asarray(m.starIndeces).reshape(-1), np.int32), ip)   # row-major [i][k]
        self.deltaE = keep(m.deltaE, np.float32)
        P.deltaE = _p(self.deltaE, fp)
        P.inSpectrumProbDen = _p(keep(np.asarray(m.inSpectrumProbDen).reshape(-1), np.float32), fp)
        P.nSpeciesPart = _p(keep(m.nSpeciesPart, np.int32), ip)
        self._ga = _F(m.grainAbun, np.float32); P.grainAbun = _p(self._ga, fp)
        P.dustComPoint = _p(keep(m.dustComPoint, np.int32), ip)
        P.TdustSublime = _p(keep(m.TdustSublime, np.float32), fp)
        self.planeIonDistribution = np.zeros((m.grids[0].nx, m.grids[0].nz), np.int32, order="F")
        P.planeIonDistribution = _p(self.planeIonDistribution, ip)
        self.P = P

        G = (OrGrid * m.nGrids)()
        self.out = []
        for i, g in enumerate(m.grids):
            og = G[i]
            og.nx, og.ny, og.nz, og.nCells, og.motherP = g.nx, g.ny, g.nz, g.nCells, g.motherP
            gx, gy, gz = g.geoCorr
            og.geoCorrX, og.geoCorrY, og.geoCorrZ = float(gx), float(gy), float(gz)
            e = m.len_unit_exponent(g) if len_unit is None else int(round(np.log2(len_unit[i])))
            og.invLenUnit = float(np.ldexp(1.0, -e))
            arrs = dict(xAxis=keep(g.xAxis, np.float32), yAxis=keep(g.yAxis, np.float32), zAxis=keep(g.zAxis, np.float32))
            og.xAxis, og.yAxis, og.zAxis = (_p(arrs[k], fp) for k in ("xAxis", "yAxis", "zAxis"))
            act = _F(g.active, np.int32); self._keep.append(act); og.active = _p(act, ip)
            for name in ("opacity", "scaOpac", "recPDF", "dustPDF", "linePDF", "totalLines", "Tdust"):
                a = _F(getattr(g, name), np.float32)
                self._keep.append(a)
                setattr(og, name, _p(a, fp))
            dai = _F(g.dustAbunIndex, np.int32); self._keep.append(dai); og.dustAbunIndex = _p(dai, ip)
            rlp = _F(getattr(g, 'resLinePackets', None), np.int32); self._keep.append(rlp); og.resLinePackets = _p(rlp, ip)
            tshape = (g.nCells + 1, m.nbins)
            eshape = (g.nCells + 1, m.nbins + 1, m.nAngleBins + 1)
            lshape = (g.nCells + 1, max(m.nLines, 1))
            o = dict(lenExp=e)
            if fp32_tallies:
                o["Jste"] = np.zeros(tshape, np.float32, order="F")
                o["escapedPackets"] = np.zeros(eshape, np.float32, order="F")
                og.Jste, og.escapedPackets = _p(o["Jste"], fp), _p(o["escapedPackets"], fp)
                if m.lgDebug:
                    o["Jdif"] = np.zeros(tshape, np.float32, order="F")
                    o["linePackets"] = np.zeros(lshape, np.float32, order="F")
                    og.Jdif, og.linePackets = _p(o["Jdif"], fp), _p(o["linePackets"], fp)
            if int_tallies:
                o["JsteQ"] = np.zeros(tshape, np.int64, order="F")
                o["escapedQ"] = np.zeros(eshape, np.int64, order="F")
                og.JsteQ, og.escapedQ = _p(o["JsteQ"], lp), _p(o["escapedQ"], lp)
                if m.lgDebug:
                    o["JdifQ"] = np.zeros(tshape, np.int64, order="F")
                    o["linePacketsQ"] = np.zeros(lshape, np.int64, order="F")
                    og.JdifQ, og.linePacketsQ = _p(o["JdifQ"], lp), _p(o["linePacketsQ"], lp)
            if count_segments:                   # segments added per (cell, nu): out[iG-1]["JsteN"]
                o["JsteN"] = np.zeros(tshape, np.int32, order="F")
                og.JsteN = _p(o["JsteN"], ip)
                if m.lgDebug:
                    o["JdifN"] = np.zeros(tshape, np.int32, order="F")
                    og.JdifN = _p(o["JdifN"], ip)
            self.out.append(o)
        self.G = G
        self.qphotCounts = np.zeros(m.nbins, np.int64)

    def transport(self, iStar: int, first: int, n: int, seed: int = 12345, gpLoc: int = 0, cellLoc=None,
                  want_fates: bool = False, deltaE=None):
        if deltaE is not None:
            self.deltaE[iStar] = np.float32(deltaE)
        cnt = OrCounters()
        fates = np.zeros((n, 4), np.int32) if want_fates else None
        cl = np.asarray(cellLoc, np.int32) if cellLoc is not None else None
        rc = self.lib.oracle_transport(C.byref(self.P), self.G, iStar, first, n, C.c_uint64(seed), gpLoc,
                                       _p(cl, ip), C.byref(cnt), _p(self.qphotCounts, lp), _p(fates, ip))
        if rc != 0:
            raise RuntimeError(f"oracle stop condition {rc}")
        return cnt.as_dict(), fates

    def transport_reslines(self, iStar: int, seed: int = 12345, rank: int = 0, nranks: int = 1):
        cnt = OrCounters()
        n = C.c_int64()
        self.lib.oracle_transport_reslines.restype = C.c_int
        self.lib.oracle_transport_reslines.argtypes = [C.POINTER(OrParams), C.POINTER(OrGrid), C.c_int32, C.c_uint64,
                                                       C.c_int32, C.c_int32, C.POINTER(OrCounters), lp]
        rc = self.lib.oracle_transport_reslines(C.byref(self.P), self.G, iStar, C.c_uint64(seed), rank, nranks,
                                                C.byref(cnt), C.byref(n))
        if rc != 0:
            raise RuntimeError(f"oracle stop condition {rc}")
        return cnt.as_dict(), int(n.value)

    def transport_mt(self, iStar: int, first: int, n: int, seed: int = 12345, threads: int = 1):
        cnt = OrCounters()
        rc = self.lib.oracle_transport_mt(C.byref(self.P), self.G, iStar, first, n, C.c_uint64(seed), threads,
                                          C.byref(cnt), _p(self.qphotCounts, lp))
        if rc != 0:
            raise RuntimeError(f"oracle stop condition {rc}")
        return cnt.as_dict()

    # the K4 fold in numpy: integer tallies -> float32 estimators, same arithmetic as
    # mocassin_b200/csrc/tables.cu fold_j_kernel / fold_count_kernel
    def folded(self, iG: int, deltaE: float, symmetric=None):
        m = self.m
        g = m.grids[iG - 1]
        o = self.out[iG - 1]
        sym = m.lgSymmetricXYZ if symmetric is None else symmetric
        dV = g.cell_volumes(sym)
        dV[0] = 1.0
        unit = np.ldexp(1.0, o["lenExp"])
        res = {}
        for name in ("JsteQ", "JdifQ"):
            if name in o:
                length = (o[name].astype(np.float64) * unit).astype(np.float32)
                J = (length * np.float32(deltaE)).astype(np.float32) / dV[:, None]
                J = J.astype(np.float32)
                J[0, :] = 0
                res[name[:-1]] = np.asfortranarray(J)
        res["escapedPackets"] = np.asfortranarray((o["escapedQ"].astype(np.float32) * np.float32(deltaE)).astype(np.float32))
        if "linePacketsQ" in o:
            res["linePackets"] = np.asfortranarray((o["linePacketsQ"].astype(np.float32) * np.float32(deltaE)).astype(np.float32))
        return res


def _sed(self, deltaE: float):
    """Head of writeSED (output_mod.f90:2561-2568): escaped packets per (nu, viewing angle)
    summed over cells and grids -> (counts int64, raw SED float32 = float(count) * deltaE),
    both (nbins, nAngleBins+1)."""
    cnt = None
    for o in self.out:
        c = o["escapedQ"].sum(axis=0)[1:, :]
        cnt = c if cnt is None else cnt + c
    cnt = np.asfortranarray(cnt.astype(np.int64))
    return cnt, np.asfortranarray((cnt.astype(np.float32) * np.float32(deltaE)).astype(np.float32))


Oracle.sed = _sed


class OrDustIn(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nCells", "nbins", "nSpeciesMax", "nSizes", "nDustComp", "nSpeciesTot", "nTemps",
        "lgMultiDustChemistry", "lgDebug")] + [
        ("nuArray", fp), ("widFlx", fp), ("xSecArray", fp), ("dustAbsXsecP", ip),
        ("nSpeciesPart", ip), ("dustComPoint", ip), ("dustAbunIndex", ip),
        ("grainAbun", fp), ("grainWeight", fp), ("TdustSublime", fp), ("dustEmIntegral", fp)]


def _dust_in(model, nCells, tables, dustAbunIndex, keep):
    def A(a, dt):
        a = np.asfortranarray(a, dtype=dt)
        keep.append(a)
        return a

    I = OrDustIn()
    ap = A(tables["dustAbsXsecP"], np.int32)
    em = A(tables["dustEmIntegral"], np.float32)
    I.nCells, I.nbins, I.nSpeciesMax, I.nSizes = nCells, model.nbins, model.nSpeciesMax, model.nSizes
    I.nDustComp, I.nSpeciesTot, I.nTemps = int(model.nSpeciesPart.shape[0]), int(ap.shape[0]), int(em.shape[2])
    I.lgMultiDustChemistry, I.lgDebug = int(model.lgMultiDustChemistry), int(model.lgDebug)
    I.nuArray = _p(A(model.nuArray, np.float32), fp); I.widFlx = _p(A(tables["widFlx"], np.float32), fp)
    I.xSecArray = _p(A(tables["xSecArray"], np.float32), fp); I.dustAbsXsecP = _p(ap, ip)
    I.nSpeciesPart = _p(A(model.nSpeciesPart, np.int32), ip); I.dustComPoint = _p(A(model.dustComPoint, np.int32), ip)
    I.dustAbunIndex = _p(A(dustAbunIndex, np.int32), ip) if dustAbunIndex is not None else ip()
    I.grainAbun = _p(A(model.grainAbun, np.float32), fp); I.grainWeight = _p(A(tables["grainWeight"], np.float32), fp)
    I.TdustSublime = _p(A(model.TdustSublime, np.float32), fp); I.dustEmIntegral = _p(em, fp)
    return I


def dust_pdf(model, grid, tables):
    """oracle_dust_pdf (setDustPDF over all cells) -> dustPDF (0:nCells, nbins)."""
    lib = load()
    keep = []
    I = _dust_in(model, grid.nCells, tables, grid.dustAbunIndex, keep)
    T = np.asfortranarray(grid.Tdust, dtype=np.float32)
    out = np.zeros((grid.nCells + 1, model.nbins), np.float32, order="F")
    lib.oracle_dust_pdf.argtypes = [C.POINTER(OrDustIn), fp, fp]
    lib.oracle_dust_pdf.restype = None
    lib.oracle_dust_pdf(C.byref(I), _p(T, fp), _p(out, fp))
    return out


def dust_update(model, grid, tables, Jste, XHILimit, Jdif=None, lgConverged=None):
    """oracle_dust_update (dust-only updateCell + getDustT over all cells); Jste is the
    host-scaled estimator.  Returns (Tdust, lgConverged); the grid is not modified.
    lgConverged: the flags before the call (kept by cells no packet crossed), default 0."""
    lib = load()
    keep = []
    I = _dust_in(model, grid.nCells, tables, grid.dustAbunIndex, keep)
    T = np.array(grid.Tdust, dtype=np.float32, order="F", copy=True)
    conv = np.zeros(grid.nCells + 1, np.int32) if lgConverged is None else np.array(lgConverged, np.int32)
    J = np.asfortranarray(Jste, dtype=np.float32)
    Jd = np.asfortranarray(Jdif, dtype=np.float32) if Jdif is not None else None
    lib.oracle_dust_update.argtypes = [C.POINTER(OrDustIn), fp, fp, C.c_float, fp, ip]
    lib.oracle_dust_update.restype = None
    lib.oracle_dust_update(C.byref(I), _p(J, fp), _p(Jd, fp) if Jd is not None else fp(), float(XHILimit), _p(T, fp), _p(conv, ip))
    return T, conv


def get_flux(energy, temperature):
    lib = load()
    lib.oracle_get_flux.argtypes = [C.c_float, C.c_float]
    lib.oracle_get_flux.restype = C.c_float
    return float(lib.oracle_get_flux(float(energy), float(temperature)))


def photo_integrals(nbins, off, low, high, xSecArray, nuArray, J):
    """oracle_photo_integrals -> (nPhoto, heat), each (nCells+1, nBands) F-order."""
    lib = load()
    J = np.asfortranarray(J, dtype=np.float32)
    nR = J.shape[0]
    off, low, high = (np.ascontiguousarray(a, dtype=np.int32) for a in (off, low, high))
    xs = np.ascontiguousarray(xSecArray, dtype=np.float32)
    nu = np.ascontiguousarray(nuArray, dtype=np.float32)
    nb = int(off.shape[0])
    nPhoto = np.zeros((nR, nb), np.float32, order="F")
    heat = np.zeros((nR, nb), np.float32, order="F")
    lib.oracle_photo_integrals.argtypes = [C.c_int32, C.c_int32, C.c_int32, ip, ip, ip, fp, fp, fp, fp, fp]
    lib.oracle_photo_integrals.restype = None
    lib.oracle_photo_integrals(nR - 1, nbins, nb, _p(off, ip), _p(low, ip), _p(high, ip), _p(xs, fp), _p(nu, fp),
                               _p(J, fp), _p(nPhoto, fp), _p(heat, fp))
    return nPhoto, heat
