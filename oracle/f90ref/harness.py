"""Run the f90py-translated reference (oracle/_ref/mocassin_ref.py) on a Model
(TEST INFRASTRUCTURE).

`Reference(model)` fills the translated module's variables (the module variables of
common_mod / continuum_mod the hot path reads, and `grid(:)`) from a
mocassin_b200.model.Model-shaped object (duck typed), installs the oracle's Philox stream
as RANDOM_NUMBER, and calls the reference's own `energyPacketDriver`
(photon_mod.f90:26).  Outputs are the reference's own arrays: `grid(iG)%Jste`,
`%escapedPackets`, `%Jdif`, `%linePackets` (float32, accumulated sequentially in packet
order exactly as the Fortran does), `Qphot`, `absInt`, `scaInt`, `planeIonDistribution`,
plus per-packet instrumentation from two hooks the translator plants (trips of the
cell-crossing loop `do j` of pathSegment, calls of energyPacketRun).
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import sys

import numpy as np

from . import build_ref, rt

HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref():
    """import oracle/_ref/mocassin_ref.py, translating it first if needed"""
    path = build_ref.build()
    d = os.path.dirname(path)
    if d not in sys.path:
        sys.path.insert(0, d)
    name = os.path.splitext(os.path.basename(path))[0]
    if name in sys.modules:
        return sys.modules[name]
    return importlib.import_module(name)


class PhiloxStream:
    """RANDOM_NUMBER bound to the oracle's per-packet stream (oracle_uniforms): uniform k of
    packet `pid` of source `stream` under `seed`."""

    def __init__(self, lib, seed):
        self.lib = lib
        self.seed = int(seed)
        self.buf = np.zeros(0, np.float32)
        self.k = 0
        self.pid = None
        self.stream = 0

    def start(self, pid, stream):
        self.pid, self.stream, self.k = int(pid), int(stream), 0
        self._fill(64)

    def _fill(self, n):
        self.buf = np.zeros(n, np.float32)
        self.lib.oracle_uniforms(C.c_uint64(self.seed), C.c_uint64(self.pid), C.c_uint32(self.stream), n,
                                 self.buf.ctypes.data_as(C.POINTER(C.c_float)))

    def next(self):
        if self.k >= self.buf.shape[0]:
            self._fill(self.buf.shape[0] * 4)
        v = self.buf[self.k]
        self.k += 1
        return v


def _F(a, dt):
    return np.array(a, dtype=dt, order='F', copy=True)


class Reference:
    def __init__(self, model, oracle_lib, math: str = 'detmath', uninit_int: int = 0):
        self.ref = ref = load_ref()
        self.m = m = model
        self.lib = oracle_lib
        if math == 'detmath':
            rt.use_detmath(oracle_lib)
        else:
            rt.use_libm()
        rt.UNINIT_INT = int(uninit_int)
        G = self.G = ref.init_globals()
        at = m.angle_tables()
        # flags and scalars (common_mod)
        G.lgdust, G.lggas = bool(m.lgDust), bool(m.lgGas)
        G.lgsymmetricxyz, G.lgisotropic = bool(m.lgSymmetricXYZ), bool(m.lgIsotropic)
        G.lgplaneionization, G.lgdebug = bool(m.lgPlaneIonization), bool(m.lgDebug)
        G.lgmultistars, G.lgmultidustchemistry = bool(m.lgMultistars), bool(m.lgMultiDustChemistry)
        G.lg1d, G.lgtalk = False, False
        G.taskid, G.numtasks = 0, 1
        G.nbins, G.ngrids, G.nstars = int(m.nbins), int(m.nGrids), int(m.nStars)
        G.nanglebins = int(m.nAngleBins)
        G.totanglebinstheta, G.totanglebinsphi = int(at['totAngleBinsTheta']), int(at['totAngleBinsPhi'])
        G.dtheta, G.dphi = np.float32(at['dTheta']), np.float32(at['dPhi'])
        G.r_out = np.float32(m.R_out)
        G.ionedge[1] = np.float32(m.ionEdge1)
        G.nlines = int(m.nLines)
        G.nspeciesmax, G.nsizes = int(m.nSpeciesMax), int(m.nSizes)
        # resonance-line transfer off unless asked for (photon_mod.f90:175-178)
        G.lgreslinesfirst = True
        G.niteratemc = 1
        G.convpercent = np.float32(0.0)
        G.reslinestransfer = np.float32(101.0)
        # frequency-indexed tables
        G.nuarray = rt.wrap(_F(m.nuArray, np.float32))
        G.gsca = rt.wrap(_F(m.gSca, np.float32)) if m.gSca is not None else None
        G.viewpointptheta = rt.wrap(_F(at['viewPointPtheta'], np.int64), (0,))
        G.viewpointpphi = rt.wrap(_F(at['viewPointPphi'], np.int64), (0,))
        G.viewpointtheta = rt.wrap(_F(at['viewPointTheta'], np.float32), (0,))
        G.viewpointphi = rt.wrap(_F(at['viewPointPhi'], np.float32), (0,))
        sp = np.asarray(m.starPosition, np.float32).reshape(-1, 3)
        pos = np.empty(sp.shape[0], dtype=object)
        for i in range(sp.shape[0]):
            pos[i] = ref.T_vector(sp[i, 0], sp[i, 1], sp[i, 2])
        G.starposition = rt.wrap(pos)
        G.starindeces = rt.wrap(_F(np.asarray(m.starIndeces).reshape(-1, 4), np.int64))
        G.deltae = rt.wrap(_F(m.deltaE, np.float32), (0,))
        G.lstar = rt.wrap(np.zeros(max(m.nStars, 1), np.float32))
        G.inspectrumprobden = rt.wrap(_F(np.asarray(m.inSpectrumProbDen).reshape(m.nStars + 1, m.nbins), np.float32), (0, 1))
        G.nspeciespart = rt.wrap(_F(m.nSpeciesPart, np.int64))
        G.grainabun = rt.wrap(_F(m.grainAbun, np.float32))
        G.dustcompoint = rt.wrap(_F(m.dustComPoint, np.int64))
        G.tdustsublime = rt.wrap(_F(m.TdustSublime, np.float32))
        G.nphotonsdiffuseloc = 1
        g0 = m.grids[0]
        G.planeiondistribution = rt.wrap(np.zeros((g0.nx, g0.nz), np.int64, order='F'))
        # grid(:)
        grids = np.empty(m.nGrids, dtype=object)
        self.out = []
        for i, g in enumerate(m.grids):
            t = ref.T_grid_type()
            t.nx, t.ny, t.nz, t.ncells, t.motherp = g.nx, g.ny, g.nz, int(g.nCells), int(g.motherP)
            gx, gy, gz = g.geoCorr
            t.geocorrx, t.geocorry, t.geocorrz = np.float32(gx), np.float32(gy), np.float32(gz)
            t.xaxis, t.yaxis, t.zaxis = (rt.wrap(_F(a, np.float32)) for a in (g.xAxis, g.yAxis, g.zAxis))
            t.active = rt.wrap(_F(g.active, np.int64))
            for name in ('opacity', 'scaOpac', 'absOpac', 'recPDF', 'dustPDF', 'linePDF'):
                a = getattr(g, name, None)
                setattr(t, name.lower(), rt.wrap(_F(a, np.float32), (0, 1)) if a is not None else None)
            t.totallines = rt.wrap(_F(g.totalLines, np.float32), (0,)) if g.totalLines is not None else None
            t.tdust = rt.wrap(_F(g.Tdust, np.float32), (0, 0, 0)) if g.Tdust is not None else None
            t.dustabunindex = rt.wrap(_F(g.dustAbunIndex, np.int64), (0,)) if g.dustAbunIndex is not None else None
            rlp = getattr(g, 'resLinePackets', None)
            t.reslinepackets = rt.wrap(_F(rlp, np.int64), (0,)) if rlp is not None else None
            t.ldiffuseloc = rt.wrap(np.zeros(g.nCells + 1, np.float32), (0,))
            o = {}
            o['Jste'] = np.zeros((g.nCells + 1, m.nbins), np.float32, order='F')
            o['escapedPackets'] = np.zeros((g.nCells + 1, m.nbins + 1, m.nAngleBins + 1), np.float32, order='F')
            t.jste = rt.wrap(o['Jste'], (0, 1))
            t.escapedpackets = rt.wrap(o['escapedPackets'], (0, 0, 0))
            if m.lgDebug:
                o['Jdif'] = np.zeros((g.nCells + 1, m.nbins), np.float32, order='F')
                o['linePackets'] = np.zeros((g.nCells + 1, max(m.nLines, 1)), np.float32, order='F')
                t.jdif = rt.wrap(o['Jdif'], (0, 1))
                t.linepackets = rt.wrap(o['linePackets'], (0, 1))
            self.out.append(o)
            grids[i] = t
        self.grid = rt.wrap(grids)
        self.draws = None

    # ------------------------------------------------------------------------------------------
    def _run(self, iStar, n, pid_of, stream, gpLoc=None, cellLoc=None):
        """call energyPacketDriver with hooks; pid_of(k) = Philox packet id of the k-th
        packet (k = 0, 1, ...) in the order the driver starts them"""
        G = self.G
        rng = PhiloxStream(self.lib, self.seed)
        rec = []          # per packet: [segments, energyPacketRun calls, draws]
        state = {'k': -1}

        def finish():
            if state['k'] >= 0:
                rec[-1][2] = rng.k

        def loop_hook(which, val):
            if which == 'energypacketdriver.iphot':
                finish()
                state['k'] += 1
                rng.start(pid_of(state['k']), stream)
                rec.append([0, 0, 0])
            else:                                   # pathsegment.j
                rec[-1][0] += 1

        def proc_hook(which):
            if which == 'energypacketrun':
                rec[-1][1] += 1

        rt.loop_hook, rt.proc_hook, rt.rng = loop_hook, proc_hook, rng
        G.absint, G.scaint = np.float32(0.0), np.float32(0.0)
        cl = None
        if cellLoc is not None:
            cl = rt.wrap(np.asarray(cellLoc, np.int64).copy())
        try:
            with np.errstate(all='ignore'):
                self.ref.p_energypacketdriver(int(iStar), int(n), self.grid, gpLoc, cl)
        finally:
            finish()
            rt.loop_hook = rt.proc_hook = rt._no_hook
        fates = np.asarray(rec, np.int64).reshape(-1, 3)
        return dict(Qphot=np.float32(G.qphot), absInt=np.float32(G.absint), scaInt=np.float32(G.scaint),
                    nSegments=int(fates[:, 0].sum()) if fates.size else 0), fates

    def transport(self, iStar: int, first: int, n: int, seed: int = 12345, gpLoc=None, cellLoc=None):
        """packets with global ids [first, first+n) of source iStar (0 = extra diffuse source
        in cell cellLoc of grid gpLoc), same keying as oracle_transport"""
        self.seed = seed
        if iStar == 0:          # the extra diffuse source: streams of its own per emitting cell (mc_oracle.c run_packet)
            g = self.m.grids[gpLoc - 1]
            lin = (cellLoc[0] - 1) + g.nx * ((cellLoc[1] - 1) + g.ny * (cellLoc[2] - 1))
            return self._run(0, n, lambda k: (int(gpLoc) << 48) + first + k, 0x80000000 + int(lin), gpLoc, cellLoc)
        return self._run(iStar, n, lambda k: first + k, iStar, gpLoc, cellLoc)

    def transport_reslines(self, iStar: int, seed: int = 12345):
        """the resonance-line packet loop of energyPacketDriver (photon_mod.f90:180-266) alone:
        no stellar packets (n = 0), transfer conditions forced on; packet j of the loop is
        keyed 2^40 + j as in oracle_transport_reslines"""
        G = self.G
        self.seed = seed
        G.lgreslinesfirst, G.niteratemc = False, 2
        G.convpercent, G.reslinestransfer = np.float32(100.0), np.float32(0.0)
        try:
            return self._run(iStar, 0, lambda k: (1 << 40) + k, iStar)
        finally:
            G.lgreslinesfirst, G.niteratemc = True, 1
