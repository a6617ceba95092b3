"""Host-side mirror of the reference's interface to the transport hot path.

The reference's seam is a handful of Fortran calls inside ``iterateMC``
(``source/iteration_mod.f90``); :class:`PacketEngine` exposes the same operations with
the same names, argument meaning and error behaviour, on top of the C ABI
(``include/mcb200.h``):

=============================================  ==========================================
reference                                      here
=============================================  ==========================================
``call ionizationDriver`` loop + dust add      :meth:`PacketEngine.assemble_opacity` /
(iteration_mod.f90:117-227)                    :meth:`PacketEngine.set_opacity`
recPDF/dustPDF/totalLines after emissionDriver :meth:`PacketEngine.set_pdfs`
(:279-424)
zero Jste/escapedPackets (:458-472)            :meth:`PacketEngine.zero_estimators`
``call energyPacketDriver(iStar, load, grid)`` :meth:`PacketEngine.energyPacketDriver`
(:474-550, photon_mod.f90:26)
``MPI_ALLREDUCE`` block (:583-703)             :meth:`PacketEngine.reduce`
copy back + scaling (:664-724)                 :meth:`PacketEngine.fetch` (raw sums) and
                                               :func:`scale_estimators` (host scaling)
=============================================  ==========================================

Errors: the reference does ``print*; stop``; here every failing call raises
:class:`MocassinError` carrying the library's message (the Fortran shim turns the
same status codes into ``stop``).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .model import F32, I32, Grid, Model


class MocassinError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mcb200 error {code}: {msg}")
        self.code = code


def _fp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float32
    return a.ctypes.data_as(_lib.c_float_p)


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32
    return a.ctypes.data_as(_lib.c_int32_p)


def _lp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int64
    return a.ctypes.data_as(_lib.c_int64_p)


def _f(a, dtype=F32):
    """Fortran-contiguous array of the given dtype (no copy when already so)."""
    if a is None:
        return None
    return np.asfortranarray(a, dtype=dtype)


def partition(n_global: int, rank: int, nranks: int) -> tuple[int, int]:
    """The reference's packet split over MPI ranks (iteration_mod.f90:477-493):
    ``load=int(n/numtasks)``, ranks below ``mod(n,numtasks)`` take one more.
    Returns (first global packet id, count) of `rank`; ids are contiguous per rank so the
    Philox streams, and therefore the results, do not depend on `nranks`."""
    load, rest = divmod(int(n_global), int(nranks))
    mine = load + (1 if rank < rest else 0)
    first = rank * load + min(rank, rest)
    return first, mine


class PacketEngine:
    """One rank's (one GPU's) transport engine for a :class:`Model`."""

    def __init__(self, model: Model, device: int = 0, rank: int = 0, nranks: int = 1, seed: int = 12345):
        self.lib = _lib.load()
        self.model = model
        self.rank, self.nranks = rank, nranks
        h = C.c_void_p()
        rc = self.lib.mcb200_create(C.byref(h), device, rank, nranks, C.c_uint64(seed))
        if rc != 0:
            raise MocassinError(rc, "mcb200_create failed (no CUDA device? this library has no CPU fallback)")
        self.h = h
        self._keep = []
        self.sed_local = False
        self.sparse_escaped = True       # N>1: exchange the escape counts as (index, count) lists
        self.pipelined_fold = False      # N>1: fold plane chunks while later chunks are exchanged (measured: no gain)
        self.exchange_chunk_planes = 64
        self.last_escaped_exchange = None
        self.native_comm = False         # N>1: the library's own NCCL communicator does the exchange
        self.solo = False                # option solo: this rank runs every packet itself, no exchange
        self._upload_static()

    # -- plumbing ---------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise MocassinError(rc, self.lib.mcb200_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mcb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        self._check(self.lib.mcb200_set_option(self.h, name.encode(), int(value)))
        if name == "solo":
            self.solo = bool(value)

    # -- static inputs ------------------------------------------------------------------
    def _upload_static(self):
        m = self.model
        at = m.angle_tables()
        cfg = _lib.Config(
            nGrids=m.nGrids, nbins=m.nbins, nStars=m.nStars, nAngleBins=m.nAngleBins,
            totAngleBinsTheta=at["totAngleBinsTheta"], totAngleBinsPhi=at["totAngleBinsPhi"],
            nLines=m.nLines, lgDust=int(m.lgDust), lgGas=int(m.lgGas),
            lgSymmetricXYZ=int(m.lgSymmetricXYZ), lgIsotropic=int(m.lgIsotropic),
            lgPlaneIonization=int(m.lgPlaneIonization), lgDebug=int(m.lgDebug),
            lgMultistars=int(m.lgMultistars), lgMultiDustChemistry=int(m.lgMultiDustChemistry),
            nSpeciesMax=m.nSpeciesMax, nSizes=m.nSizes, nDustComp=int(m.nSpeciesPart.shape[0]),
            dTheta=float(at["dTheta"]), dPhi=float(at["dPhi"]), R_out=float(m.R_out),
            ionEdge1=float(m.ionEdge1))
        self._check(self.lib.mcb200_set_config(self.h, C.byref(cfg)))
        for i, g in enumerate(m.grids, start=1):
            act = _f(g.active, I32)
            self._check(self.lib.mcb200_set_grid(
                self.h, i, g.nx, g.ny, g.nz, g.nCells, g.motherP,
                _fp(_f(g.xAxis)), _fp(_f(g.yAxis)), _fp(_f(g.zAxis)), _ip(act)))
        cdf = _f(np.asarray(m.inSpectrumProbDen, dtype=F32))      # (0:nStars, nbins) star fastest
        self._check(self.lib.mcb200_set_spectra(self.h, _fp(_f(m.nuArray)),
                                                _fp(_f(m.gSca)) if m.gSca is not None else None, _fp(cdf)))
        if m.nStars > 0:
            pos = np.ascontiguousarray(m.starPosition, dtype=F32)     # x,y,z triplets per star
            idx = _f(np.asarray(m.starIndeces, dtype=I32), I32)       # (nStars,4) star fastest
            self._check(self.lib.mcb200_set_stars(self.h, _fp(pos), _ip(idx)))
        if m.nAngleBins > 0:
            self._check(self.lib.mcb200_set_viewpoints(
                self.h, _ip(at["viewPointPtheta"]), _ip(at["viewPointPphi"]),
                _fp(at["viewPointTheta"]), _fp(at["viewPointPhi"])))
        if m.lgDust:
            self._check(self.lib.mcb200_set_dust_species(
                self.h, _ip(_f(m.nSpeciesPart, I32)), _fp(_f(m.grainAbun)), _ip(_f(m.dustComPoint, I32)),
                _fp(_f(m.TdustSublime)), int(m.TdustSublime.shape[0])))

    # -- per-iteration inputs -----------------------------------------------------------
    def set_opacity(self, iG: int = 0):
        """Upload host-assembled ``opacity``/``scaOpac`` (all grids when iG==0)."""
        for i, g in self._grids(iG):
            self._check(self.lib.mcb200_set_opacity(self.h, i, _fp(_f(g.opacity)),
                                                    _fp(_f(g.scaOpac)) if self.model.lgDust else None))

    def set_pdfs(self, iG: int = 0):
        m = self.model
        for i, g in self._grids(iG):
            self._check(self.lib.mcb200_set_pdfs(
                self.h, i,
                _fp(_f(g.recPDF)) if m.lgGas else None,
                _fp(_f(g.dustPDF)) if not m.lgGas else None,
                _fp(_f(g.totalLines)) if m.lgGas else None,
                _fp(_f(g.linePDF)) if (m.lgDebug and m.lgGas) else None))

    def set_dust_state(self, iG: int = 0):
        m = self.model
        if not m.lgDust:
            return
        for i, g in self._grids(iG):
            self._check(self.lib.mcb200_set_dust_state(
                self.h, i, _fp(_f(g.Tdust)),
                _ip(_f(g.dustAbunIndex, I32)) if g.dustAbunIndex is not None else None))

    # -- dust-only closure (update_mod.f90 getDustT, emission_mod.f90 setDustPDF) --------
    def set_dust_tables(self, widFlx, grainWeight, dustAbsXsecP, dustEmIntegral):
        """Tables of the dust closure; call after set_xsec and before set_dust_state.
        ``dustAbsXsecP`` (nSpecies, nSizes) 1-based offsets into xSecArray,
        ``dustEmIntegral`` (nSpecies, nSizes, nTemps), both in Fortran order."""
        ap = _f(dustAbsXsecP, I32)
        em = _f(dustEmIntegral)
        self._check(self.lib.mcb200_set_dust_tables(
            self.h, _fp(_f(widFlx)), _fp(_f(grainWeight)), _ip(ap), int(ap.shape[0]), _fp(em), int(em.shape[2])))

    def getDustT(self, iG: int, XHILimit: float):
        """Dust-only updateCell over every cell of grid iG (update_mod.f90:308-334,
        :1836-1945) from the device-resident Jste.  Returns (Tdust, lgConverged, nConverged);
        also stored on the host grid as the reference does."""
        g = self.model.grids[iG - 1]
        T = np.zeros_like(_f(g.Tdust), order="F")
        conv = np.zeros(g.nCells + 1, dtype=I32)
        n = C.c_int64(0)
        self._check(self.lib.mcb200_dust_update(self.h, iG, float(XHILimit), _fp(T), _ip(conv), C.byref(n)))
        g.Tdust = T
        g.lgConverged = conv
        return T, conv, int(n.value)

    def setDustPDF(self, iG: int, fetch: bool = False):
        """emissionDriver's dust-only work for every cell of grid iG (emission_mod.f90:1313-1387),
        from the device Tdust into the device re-emission tables.  ``fetch`` returns dustPDF
        (0:nCells, nbins)."""
        g = self.model.grids[iG - 1]
        out = np.zeros((g.nCells + 1, self.model.nbins), dtype=F32, order="F") if fetch else None
        self._check(self.lib.mcb200_dust_pdf(self.h, iG, _fp(out) if fetch else None))
        return out

    def photo_integrals(self, iG: int, off, low, high, dif: bool = False) -> dict:
        """Photo-ionisation and heating integrals of updateCell / thermBalance
        (update_mod.f90:170-262, :1160-1214) per (cell, band) from the device-resident Jste.
        Bands are 1-based (xSecArray offset, first bin, last bin).  Returns arrays
        (nCells+1, nBands): nPhotoSte, heatSte (and nPhotoDif, heatDif with ``dif``)."""
        g = self.model.grids[iG - 1]
        off, low, high = (_f(a, I32) for a in (off, low, high))
        nb = int(off.shape[0])
        mk = lambda: np.zeros((g.nCells + 1, nb), dtype=F32, order="F")
        out = dict(nPhotoSte=mk(), heatSte=mk())
        if dif:
            out.update(nPhotoDif=mk(), heatDif=mk())
        self._check(self.lib.mcb200_photo_integrals(
            self.h, iG, nb, _ip(off), _ip(low), _ip(high), _fp(out["nPhotoSte"]), _fp(out["heatSte"]),
            _fp(out.get("nPhotoDif")), _fp(out.get("heatDif"))))
        return out

    def upload_iteration_inputs(self):
        self.set_opacity()
        self.set_pdfs()
        self.set_dust_state()

    def set_xsec(self, xSecArray: np.ndarray):
        x = _f(xSecArray)
        self._check(self.lib.mcb200_set_xsec(self.h, _fp(x), int(x.shape[0])))

    def assemble_opacity(self, iG: int, bands, den: np.ndarray, ff1: Optional[np.ndarray] = None, dust=None):
        """K1 on device.  `bands` = dict(species, off, low, high) int32 arrays (1-based
        Fortran values as in addOpacity); `den` (nCells+1, nSpeciesDen) F-order; `dust` =
        dict(Ndust, Tdust, dustAbunIndex, grainWeight, dustScaXsecP, dustAbsXsecP) or None;
        Tdust = None takes the sublimation mask from the device-resident dust state (after
        set_dust_state / getDustT), so a dust-only Lucy iteration never brings Tdust to the host."""
        sp, off, lo, hi = (_f(bands[k], I32) for k in ("species", "off", "low", "high"))
        den = _f(den)
        nsd = int(den.shape[1]) if den.ndim == 2 else 0
        d = dust or {}
        nTot = int(np.asarray(d["dustScaXsecP"]).shape[0]) if dust else 0
        self._check(self.lib.mcb200_assemble_opacity(
            self.h, iG, int(sp.shape[0]), _ip(sp), _ip(off), _ip(lo), _ip(hi), nsd, _fp(den),
            _fp(_f(ff1)) if ff1 is not None else None,
            _fp(_f(d["Ndust"])) if dust else None, _fp(_f(d["Tdust"])) if dust and d.get("Tdust") is not None else None,
            _ip(_f(d["dustAbunIndex"], I32)) if dust and d.get("dustAbunIndex") is not None else None,
            _fp(_f(d["grainWeight"])) if dust else None,
            _ip(_f(d["dustScaXsecP"], I32)) if dust else None,
            _ip(_f(d["dustAbsXsecP"], I32)) if dust else None, nTot))

    def get_opacity_rows(self, iG: int, cells) -> np.ndarray:
        """(len(cells), nbins) rows opacity(cell, :) gathered on the device (for writeTauNu:
        mocassin_b200.output.tau_nu(model, lambda c: eng.get_opacity_rows(1, c)))."""
        cells = np.ascontiguousarray(cells, dtype=I32)
        out = np.zeros((cells.shape[0], self.model.nbins), dtype=F32, order="F")
        self._check(self.lib.mcb200_get_opacity_rows(self.h, iG, int(cells.shape[0]), _ip(cells), _fp(out)))
        return out

    def get_opacity(self, iG: int, want_abs: bool = False):
        g = self.model.grids[iG - 1]
        shape = (g.nCells + 1, self.model.nbins)
        op = np.zeros(shape, dtype=F32, order="F")
        sca = np.zeros(shape, dtype=F32, order="F") if self.model.lgDust else None
        ab = np.zeros(shape, dtype=F32, order="F") if want_abs else None
        self._check(self.lib.mcb200_get_opacity(self.h, iG, _fp(op), _fp(sca) if sca is not None else None,
                                                _fp(ab) if ab is not None else None))
        return op, sca, ab

    # -- the hot path -------------------------------------------------------------------
    def zero_estimators(self):
        self._check(self.lib.mcb200_zero_estimators(self.h))

    def energyPacketDriver(self, iStar: int, n: int, gpLoc: Optional[int] = None, cellLoc=None,
                           deltaE: Optional[float] = None) -> dict:
        """``call energyPacketDriver(iStar, n, grid[, gpLoc, cellLoc])`` (photon_mod.f90:26)
        for *all* ranks at once: `n` is the global packet count ``nPhotons(iStar)``, this
        rank transports its :func:`partition` share."""
        cnt = _lib.Counters()
        if deltaE is None:
            deltaE = float(self.model.deltaE[iStar])
        if iStar >= 1:
            rc = self.lib.mcb200_transport(self.h, iStar, int(n), C.c_float(deltaE), C.byref(cnt))
        else:
            cl = np.asarray(cellLoc, dtype=I32)
            rc = self.lib.mcb200_transport_diffuse(self.h, int(gpLoc), _ip(cl), int(n), C.c_float(deltaE), C.byref(cnt))
        self._check(rc)
        return cnt.as_dict()

    def resLinePacketsTransfer(self, iStar: int, deltaE: Optional[float] = None) -> dict:
        """The resonance-line packet loop of energyPacketDriver (photon_mod.f90:180-266):
        `grid.resLinePackets(cell)` diffuse packets from the centre of every cell of this rank."""
        for i, g in enumerate(self.model.grids, start=1):
            r = getattr(g, "resLinePackets", None)
            self._check(self.lib.mcb200_set_res_line_packets(self.h, i, _ip(_f(r, I32)) if r is not None else None))
        cnt = _lib.Counters()
        if deltaE is None:
            deltaE = float(self.model.deltaE[iStar])
        self._check(self.lib.mcb200_transport_reslines(self.h, iStar, C.c_float(deltaE), C.byref(cnt)))
        return cnt.as_dict()

    def tally_buffer(self, iG: int, which: int) -> tuple[int, int]:
        p = C.c_void_p()
        n = C.c_int64()
        self._check(self.lib.mcb200_tally_buffer(self.h, iG, which, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def _exchange(self, tset: int = 0, group=None, async_op: bool = False, sed: bool = True) -> list:
        """All-reduce the pending integer tallies of tally set `tset` over ranks (NCCL over
        NVLink, replacing MPI_ALLREDUCE at iteration_mod.f90:627-659): first the touched-bin
        flags (tiny max-reduce), then only the flagged nu-planes -- int64 path lengths, uint32
        packet counts (summed as int32 bit patterns).  Returns the NCCL work handles when
        async_op (the transfers then overlap whatever the library stream does next)."""
        import torch
        import torch.distributed as dist

        m = self.model
        dev = self._device_index()
        base = 16 * tset
        works = []
        for iG in range(1, m.nGrids + 1):
            nR = m.grids[iG - 1].nCells + 1
            fptr, fn = self.tally_buffer(iG, base + 4)
            flags = _as_cuda_tensor(fptr, fn, "<i4", dev)
            dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
            ranges = _touched_ranges(flags.cpu().numpy())
            self.last_exchange_planes = (sum(b - a + 1 for a, b in ranges), m.nbins + 1, len(ranges))
            whichs = [0] + ([] if self.sed_local else [1]) + ([2] if (m.lgDebug and tset == 0) else [])
            for w in whichs:
                if w == 1 and self.sparse_escaped and not async_op and self._exchange_escaped_sparse(iG, tset, ranges, group):
                    continue
                ptr, n = self.tally_buffer(iG, base + w)
                if n == 0:
                    continue
                t = _as_cuda_tensor(ptr, n, "<i8" if w in (0, 2) else "<i4", dev)
                for p0, p1 in ranges:
                    if w == 1:
                        for ang in range(m.nAngleBins + 1):
                            off = nR * (p0 + (m.nbins + 1) * ang)
                            works.append(dist.all_reduce(t[off:off + (p1 - p0 + 1) * nR], op=dist.ReduceOp.SUM,
                                                         group=group, async_op=async_op))
                    elif p1 >= max(p0, 1):
                        q0 = max(p0, 1)
                        works.append(dist.all_reduce(t[(q0 - 1) * nR:p1 * nR], op=dist.ReduceOp.SUM, group=group,
                                                     async_op=async_op))
            if m.lgDebug and tset == 0:
                ptr, n = self.tally_buffer(iG, 3)
                if n:
                    works.append(dist.all_reduce(_as_cuda_tensor(ptr, n, "<i4", dev), op=dist.ReduceOp.SUM,
                                                 group=group, async_op=async_op))
            if m.lgPlaneIonization and tset == 0 and iG == 1:
                ptr, n = self.tally_buffer(1, 5)
                if n:
                    works.append(dist.all_reduce(_as_cuda_tensor(ptr, n, "<i4", dev), op=dist.ReduceOp.SUM,
                                                 group=group, async_op=async_op))
        if self.sed_local and sed:
            ptr, n = self.tally_buffer(1, 6)
            works.append(dist.all_reduce(_as_cuda_tensor(ptr, n, "<i8", dev), op=dist.ReduceOp.SUM, group=group,
                                         async_op=async_op))
        return [w for w in works if w is not None] if async_op else []

    def _exchange_escaped_sparse(self, iG: int, tset: int, ranges, group=None) -> bool:
        """All-gather the non-zero (index, count) pairs of the escape counts instead of
        all-reducing the dense array (mcb200_escaped_compact / _scatter).  Returns False, with
        the array restored, when the lists would move more bytes than the dense exchange."""
        import torch
        import torch.distributed as dist

        m = self.model
        dev = self._device_index()
        world = self.nranks
        p = C.c_void_p()
        n = C.c_int64()
        self._check(self.lib.mcb200_escaped_compact(self.h, iG, tset, C.byref(p), C.byref(n)))
        n = int(n.value)
        sizes = torch.tensor([n], dtype=torch.int64, device=f"cuda:{dev}")
        allsz = torch.empty(world, dtype=torch.int64, device=f"cuda:{dev}")
        dist.all_gather_into_tensor(allsz, sizes, group=group)
        maxn = int(allsz.max().item())
        nR = m.grids[iG - 1].nCells + 1
        dense = 2 * 4 * nR * (m.nAngleBins + 1) * sum(b - a + 1 for a, b in ranges)
        mine = _as_cuda_tensor(int(p.value or 0), 2 * n, "<i8", dev) if n else None
        self.last_escaped_exchange = dict(entries=n, max_entries=maxn, sparse_bytes=16 * maxn * world, dense_bytes=dense)
        if 16 * maxn * world >= dense:
            if n:
                self._check(self.lib.mcb200_escaped_scatter(self.h, iG, tset, C.c_void_p(mine.data_ptr()), n))
            return False
        if maxn == 0:
            return True
        pad = torch.zeros(2 * maxn, dtype=torch.int64, device=f"cuda:{dev}")
        if n:
            pad[:2 * n] = mine
        gathered = torch.empty(2 * maxn * world, dtype=torch.int64, device=f"cuda:{dev}")
        dist.all_gather_into_tensor(gathered, pad, group=group)
        torch.cuda.synchronize()
        self._check(self.lib.mcb200_escaped_scatter(self.h, iG, tset, C.c_void_p(gathered.data_ptr()), maxn * world))
        return True

    def set_sed_local(self, on: bool = True):
        """Exchange the per-(nu, angle) escape counts (a few KB) instead of the per-cell
        escapedPackets tallies (5 GB at 128^3 x 600): the SED stays exact and global,
        escapedPackets becomes rank-local."""
        self.set_option("sed_local", 1 if on else 0)
        self.sed_local = bool(on)

    def fetch_sed(self):
        """(SED, counts), each (nbins, nAngleBins+1): raw sums over cells and grids of
        escapedPackets (head of writeSED, output_mod.f90:2561-2568) and the integer packet counts."""
        m = self.model
        sed = np.zeros((m.nbins, m.nAngleBins + 1), dtype=F32, order="F")
        cnt = np.zeros((m.nbins, m.nAngleBins + 1), dtype=np.int64, order="F")
        self._check(self.lib.mcb200_fetch_sed(self.h, _fp(sed), cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        return sed, cnt

    def fetch_contcube(self, iG: int = 1) -> np.ndarray:
        """(nCells+1, nAngleBins+1): sum over the frequency bins of escapedPackets per cell and
        viewing angle (the reduction of writeContCube, output_mod.f90:2762-2772), raw sums."""
        g = self.model.grids[iG - 1]
        out = np.zeros((g.nCells + 1, self.model.nAngleBins + 1), dtype=F32, order="F")
        self._check(self.lib.mcb200_fetch_contcube(self.h, iG, _fp(out)))
        return out

    def _exchange_pipelined(self, group=None) -> bool:
        """Exchange with the fold hidden behind it: the escape counts go first (sparse), then
        the touched JsteQ planes are all-reduced in chunks on NCCL's stream and every chunk is
        folded (mcb200_reduce_range, library stream) as soon as it has arrived, while the next
        chunk is in flight.  Returns False if the plain path has to be used."""
        import torch
        import torch.distributed as dist

        m = self.model
        if m.lgDebug or self.sed_local or not self.sparse_escaped:
            return False
        dev = self._device_index()
        plan = []
        for iG in range(1, m.nGrids + 1):
            nR = m.grids[iG - 1].nCells + 1
            fptr, fn = self.tally_buffer(iG, 4)
            flags = _as_cuda_tensor(fptr, fn, "<i4", dev)
            dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
            ranges = _touched_ranges(flags.cpu().numpy())
            self.last_exchange_planes = (sum(b - a + 1 for a, b in ranges), m.nbins + 1, len(ranges))
            if not self._exchange_escaped_sparse(iG, 0, ranges, group):
                # dense fallback for the escape counts, then the plain fold
                ptr, n = self.tally_buffer(iG, 1)
                t = _as_cuda_tensor(ptr, n, "<i4", dev)
                for p0, p1 in ranges:
                    for ang in range(m.nAngleBins + 1):
                        off = nR * (p0 + (m.nbins + 1) * ang)
                        dist.all_reduce(t[off:off + (p1 - p0 + 1) * nR], op=dist.ReduceOp.SUM, group=group)
            ptr, n = self.tally_buffer(iG, 0)
            t = _as_cuda_tensor(ptr, n, "<i8", dev)
            step = max(1, self.exchange_chunk_planes)
            for p0, p1 in ranges:
                a = p0
                while a <= p1:
                    b = min(a + step - 1, p1)
                    q0 = max(a, 1)
                    w = dist.all_reduce(t[(q0 - 1) * nR:b * nR], op=dist.ReduceOp.SUM, group=group, async_op=True) if b >= q0 else None
                    plan.append((iG, a, b, w))
                    a = b + 1
            if m.lgPlaneIonization and iG == 1:
                ptr, n = self.tally_buffer(1, 5)
                if n:
                    dist.all_reduce(_as_cuda_tensor(ptr, n, "<i4", dev), op=dist.ReduceOp.SUM, group=group)
        for iG, a, b, w in plan:
            if w is not None:
                w.wait()
            torch.cuda.current_stream().synchronize()
            self._check(self.lib.mcb200_reduce_range(self.h, iG, a, b))
        return True

    # -- the library's own communicator (what a Fortran/MPI host uses) -------------------
    def comm_unique_id(self) -> bytes:
        """128-byte ncclUniqueId (rank 0 makes it, the host broadcasts it: MPI_BCAST in the
        reference, any byte channel here)."""
        buf = C.create_string_buffer(128)
        self._check(self.lib.mcb200_comm_unique_id(self.h, buf))
        return buf.raw

    def comm_init(self, unique_id: bytes):
        """mcb200_comm_init: join the NCCL communicator of `unique_id` as (rank, nranks) of
        the constructor; reduce() then runs mcb200_exchange instead of torch.distributed."""
        if len(unique_id) != 128:
            raise ValueError("ncclUniqueId is 128 bytes")
        self._check(self.lib.mcb200_comm_init(self.h, C.create_string_buffer(unique_id, 128)))
        self.native_comm = True

    def comm_init_from_group(self, group=None):
        """Bootstrap the native communicator over an existing torch.distributed group (any
        backend: only the 128 id bytes travel)."""
        import torch.distributed as dist

        box = [self.comm_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        self.comm_init(box[0])

    def comm_destroy(self):
        self._check(self.lib.mcb200_comm_destroy(self.h))
        self.native_comm = False

    def exchange(self) -> dict:
        """mcb200_exchange: sum the pending integer tallies over the ranks of the native
        communicator (no fold).  Returns mcb200_exchange_info."""
        self._check(self.lib.mcb200_exchange(self.h))
        return self.exchange_info()

    def exchange_info(self) -> dict:
        b, sp, v = C.c_int64(), C.c_int32(), C.c_int32()
        self._check(self.lib.mcb200_exchange_info(self.h, C.byref(b), C.byref(sp), C.byref(v)))
        path = C.c_int32()
        why = C.create_string_buffer(256)
        ph = (C.c_double * 4)()
        self._check(self.lib.mcb200_exchange_path(self.h, C.byref(path), why, 256, ph))
        return dict(bytes=b.value, sparse_grids=sp.value, nccl_version=v.value,
                    exchange_ms=ph[0], reduce_ms=ph[1], j_merge_device_ms=ph[2], j_push_device_ms=ph[3],
                    path={0: "none", 1: "nccl all-reduce", 2: "nccl reduce-scatter + all-gather",
                          3: "fused peer-memory kernel (NVLink)"}.get(path.value, str(path.value)),
                    p2p_unavailable=why.value.decode() or None)

    def reduce(self, group=None):
        """Sum the pending integer tallies over ranks and fold them into the float32
        estimators.  Exact integer sums -> identical bits on every rank and for every rank
        count."""
        if self.solo:                        # acting as a single rank: the transport call folded already
            pass
        elif self.native_comm:
            self.last_exchange = self.exchange()
        elif self.nranks > 1:
            import torch

            if not (self.pipelined_fold and self._exchange_pipelined(group)):
                self._exchange(0, group)
            torch.cuda.synchronize()
        self._check(self.lib.mcb200_reduce(self.h))
        if self.native_comm and not self.solo:   # + the float32 all-gather of the fold
            self.last_exchange = self.exchange_info()

    def energyPacketDriverOverlapped(self, iStar: int, n: int, deltaE: Optional[float] = None, group=None) -> dict:
        """energyPacketDriver + exchange + fold for nranks > 1 with the exchange hidden behind
        the transport: this rank's packets are run in two halves into two tally sets; while
        the second half is transported the first half's tallies are all-reduced on NCCL's
        stream.  The sets are merged as integers before the single fold, so the result is
        bit-identical to the plain path."""
        import torch

        if self.nranks == 1 or self.model.lgDebug:
            c = self.energyPacketDriver(iStar, n, deltaE=deltaE)
            self.reduce(group)
            return c
        self.set_option("parts", 2)
        try:
            self.set_option("part", 0); self.set_option("tally_set", 0)
            c0 = self.energyPacketDriver(iStar, n, deltaE=deltaE)
            works = self._exchange(0, group, async_op=True, sed=False)
            self.set_option("part", 1); self.set_option("tally_set", 1)
            c1 = self.energyPacketDriver(iStar, n, deltaE=deltaE)
            for w in works:
                w.wait()
            self._exchange(1, group)
            torch.cuda.synchronize()
        finally:
            self.set_option("parts", 1); self.set_option("tally_set", 0)
        self._check(self.lib.mcb200_reduce(self.h))
        out = dict(c0)
        for k, v in c1.items():
            out[k] = out[k] + v
        return out

    def _device_index(self):
        import torch

        return torch.cuda.current_device()

    def fetch(self, iG: int = 1, want=("Jste", "escapedPackets"), out: Optional[dict] = None) -> dict:
        """Raw estimator sums in the reference's layouts (before the host scaling of
        iteration_mod.f90:705-724).  `out` may hold preallocated (e.g. pinned) F-order
        arrays for "Jste" / "escapedPackets"."""
        m = self.model
        g = m.grids[iG - 1]
        pre = out or {}
        out = {}
        J = pre.get("Jste") if "Jste" in want else None
        E = pre.get("escapedPackets") if "escapedPackets" in want else None
        if J is None and "Jste" in want:
            J = np.zeros((g.nCells + 1, m.nbins), dtype=F32, order="F")
        if E is None and "escapedPackets" in want:
            E = np.zeros((g.nCells + 1, m.nbins + 1, m.nAngleBins + 1), dtype=F32, order="F")
        D = np.zeros((g.nCells + 1, m.nbins), dtype=F32, order="F") if "Jdif" in want else None
        Lp = np.zeros((g.nCells + 1, max(m.nLines, 1)), dtype=F32, order="F") if "linePackets" in want else None
        self._check(self.lib.mcb200_fetch_estimators(self.h, iG, _fp(J), _fp(E), _fp(D), _fp(Lp)))
        for k, v in (("Jste", J), ("escapedPackets", E), ("Jdif", D), ("linePackets", Lp)):
            if v is not None:
                out[k] = v
        return out

    def fetch_sparse(self, iG: int = 1, out: Optional[dict] = None, clear_previous: bool = True):
        """Jste (dense) and escapedPackets (sparse, see fetch_escaped_sparse) in one call
        (mcb200_fetch_estimators_sparse): the Jste copy overlaps the host-side scatter.  `out` may
        hold preallocated "Jste" / "escapedPackets" arrays.  Returns (dict, entries written)."""
        m = self.model
        g = m.grids[iG - 1]
        pre = out or {}
        J = pre.get("Jste")
        E = pre.get("escapedPackets")
        if J is None:
            J = np.zeros((g.nCells + 1, m.nbins), dtype=F32, order="F")
        if E is None:
            E = np.zeros((g.nCells + 1, m.nbins + 1, m.nAngleBins + 1), dtype=F32, order="F")
        n = C.c_int64()
        self._check(self.lib.mcb200_fetch_estimators_sparse(self.h, iG, _fp(J), _fp(E), int(bool(clear_previous)), C.byref(n)))
        return {"Jste": J, "escapedPackets": E}, int(n.value)

    def fetch_escaped_sparse(self, iG: int = 1, out: Optional[np.ndarray] = None, clear_previous: bool = True):
        """escapedPackets of grid iG through the sparse path (mcb200_fetch_escaped_sparse): only
        the non-zero entries cross PCIe and are written into `out`, which must be zero elsewhere
        (as iterateMC leaves it, iteration_mod.f90:466-470); with `clear_previous` the entries the
        previous call wrote into the same array are zeroed first.  Returns (array, entries written;
        -1 = dense fallback)."""
        m = self.model
        g = m.grids[iG - 1]
        if out is None:
            out = np.zeros((g.nCells + 1, m.nbins + 1, m.nAngleBins + 1), dtype=F32, order="F")
        n = C.c_int64()
        self._check(self.lib.mcb200_fetch_escaped_sparse(self.h, iG, _fp(out), int(bool(clear_previous)), C.byref(n)))
        return out, int(n.value)

    def fetch_tallies(self, iG: int = 1) -> dict:
        m = self.model
        g = m.grids[iG - 1]
        JQ = np.zeros((g.nCells + 1, m.nbins), dtype=np.int64, order="F")
        EQ = np.zeros((g.nCells + 1, m.nbins + 1, m.nAngleBins + 1), dtype=np.int64, order="F")
        DQ = np.zeros((g.nCells + 1, m.nbins), dtype=np.int64, order="F") if m.lgDebug else None
        LQ = np.zeros((g.nCells + 1, max(m.nLines, 1)), dtype=np.int64, order="F") if m.lgDebug else None
        self._check(self.lib.mcb200_fetch_tallies(self.h, iG, _lp(JQ), _lp(EQ), _lp(DQ), _lp(LQ)))
        return dict(JsteQ=JQ, escapedQ=EQ, JdifQ=DQ, linePacketsQ=LQ)

    def fetch_cells(self, iG: int = 1, first: Optional[int] = None, stride: Optional[int] = None,
                    out: Optional[np.ndarray] = None) -> np.ndarray:
        """Jste rows of the cells this rank owns under the reference's round-robin rule
        (iteration_mod.f90:832: cells rank+1, rank+1+nranks, ...), compact (nMine, nbins), F-order:
        mcb200_fetch_estimators_cells.  `out` may be a preallocated (pinned) array."""
        g = self.model.grids[iG - 1]
        first = self.rank + 1 if first is None else int(first)
        stride = self.nranks if stride is None else int(stride)
        nMine = 0 if first > g.nCells else (g.nCells - first) // stride + 1
        if out is None:
            out = np.zeros((nMine, self.model.nbins), dtype=F32, order="F")
        assert out.shape == (nMine, self.model.nbins) and out.flags.f_contiguous
        n = C.c_int64()
        self._check(self.lib.mcb200_fetch_estimators_cells(self.h, iG, first, stride, _fp(out), None, C.byref(n)))
        assert n.value == nMine
        return out

    def checksum(self, iG: int = 1, which: int = 0) -> int:
        """mcb200_checksum: position-sensitive 64-bit sum of the device-resident estimator
        (0 Jste, 1 escapedPackets, 2 Jdif, 3 linePackets)."""
        v = C.c_uint64()
        self._check(self.lib.mcb200_checksum(self.h, iG, which, C.byref(v)))
        return int(v.value)

    def len_unit(self, iG: int = 1) -> float:
        v = C.c_double()
        self._check(self.lib.mcb200_len_unit(self.h, iG, C.byref(v)))
        return float(v.value)

    def plane_distribution(self) -> np.ndarray:
        g = self.model.grids[0]
        d = np.zeros((g.nx, g.nz), dtype=I32, order="F")
        self._check(self.lib.mcb200_fetch_plane_distribution(self.h, _ip(d)))
        return d

    def qphot_counts(self) -> np.ndarray:
        q = np.zeros(self.model.nbins, dtype=np.int64)
        self._check(self.lib.mcb200_fetch_qphot_counts(self.h, _lp(q)))
        return q

    def fates(self, n: int) -> np.ndarray:
        f = np.zeros((n, 4), dtype=I32)
        self._check(self.lib.mcb200_fetch_fates(self.h, _ip(f), int(n)))
        return f

    def _grids(self, iG: int):
        if iG:
            return [(iG, self.model.grids[iG - 1])]
        return list(enumerate(self.model.grids, start=1))

    # -- the packet loop of iterateMC (iteration_mod.f90:458-726) -----------------------
    def lucy_transport(self, nPhotons, group=None, iteration: Optional[int] = None) -> list[dict]:
        """zero estimators; for every star energyPacketDriver; reduce.  `nPhotons[i]` is
        the global packet count of star i+1.  `iteration` (the host's nIterateMC) becomes the
        Philox epoch, so that every Lucy iteration draws fresh histories as the reference's
        wall-clock-seeded generator does (option "epoch", include/mcb200.h)."""
        if iteration is not None:
            self.set_option("epoch", int(iteration))
        self.zero_estimators()
        out = []
        for iStar in range(1, self.model.nStars + 1):
            out.append(self.energyPacketDriver(iStar, int(nPhotons[iStar - 1])))
            if (self.nranks > 1 or self.native_comm) and not self.solo:
                self.reduce(group)
        if (self.nranks == 1 and not self.native_comm) or self.solo:
            self.reduce()
        return out


def call_seed(seed: int, epoch: int = 0) -> int:
    """The Philox key of a transport call under option "epoch" (capi.cu run_transport): what a
    checker passes as `seed` to reproduce iteration `epoch` of a context created with `seed`."""
    return (int(seed) + int(epoch) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF


def _touched_ranges(flag) -> list:
    """Contiguous runs [first,last] of touched frequency bins; gaps of <= 2 bins are bridged
    (same rule as the fold in capi.cu)."""
    out = []
    for i, f in enumerate(flag):
        if not f:
            continue
        if out and i - out[-1][1] <= 3:
            out[-1][1] = i
        else:
            out.append([i, i])
    return [(a, b) for a, b in out]


def _as_cuda_tensor(ptr: int, n: int, typestr: str, device_index: int):
    """Wrap a raw device pointer as a torch tensor (no copy) via __cuda_array_interface__."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {
        "shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None,
    }
    return torch.as_tensor(h, device=f"cuda:{device_index}")


def scale_estimators(model: Model, Jste: np.ndarray, escapedPackets: np.ndarray):
    """The host's own post-scaling, iteration_mod.f90:705-724 (unchanged by this work):
    ``Jste*1e-9``, ``/8`` for symmetricXYZ (also escapedPackets)."""
    J = (Jste * F32(1.0e-9)).astype(F32)
    E = escapedPackets.copy()
    if model.lgSymmetricXYZ:
        J = (J / F32(8.0)).astype(F32)
        E = (E / F32(8.0)).astype(F32)
    return J, E
