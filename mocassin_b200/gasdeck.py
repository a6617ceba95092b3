"""The shipped gas benchmarks (benchmarks/gas/HII40, benchmarks/gas/PN150) as transport workloads in
their first-iteration state: `input.in` + the abundance file + data/ph1.dat,
data/ph2.dat -> frequency mesh, cross-section stack, band list, species densities of the
`setMotherGrid` initial state (Te = TeStart, X(H0) = 1e-5), stellar CDF, geometry.

What a gas deck needs beyond this -- recPDF / totalLines of `emissionDriver`
(emission_mod.f90:905-1311: recombination, two-photon, free-free continua and ~2000 lines from a
dozen more data files) -- belongs to the host solver, which is out of scope.  The
loader therefore takes the re-emission tables as an input (`recPDF`, `totalLines`); without them it
fills in the smooth stand-in of mocassin_b200.workloads, and says so in `tables["recPDF_kind"]`.
"""
from __future__ import annotations

import os

import numpy as np

from . import gasdata as G
from .deck import read_input, stellar_cdf, _real
from .model import F32, I32, Grid, Model, auto_axis, number_active, set_star_position
from .workloads import recombination_cdf, wid_flx

PI = F32(3.141592654)
HPLANCK = F32(6.6262e-27)
SIGMA = F32(5.66956e-5)
CRYD = F32(3.2898423e15)
HCRYD = F32(2.1799153e-11)


def lstar_from_lphot(LPhot: float, T: float, nu: np.ndarray, widFlx: np.ndarray, lymanP: int) -> np.float32:
    """LStar [1e36 erg/s] of a blackbody from Q(H) = LPhot [1e36 phot/s] (continuum_mod.f90:303-313
    with the photon normalisation of setProbDen :447-470)."""
    from .deck import get_flux_blackbody

    erg = get_flux_blackbody(nu, T)
    phot = (erg / (nu * HCRYD).astype(F32)).astype(F32)
    norm = np.float64(0.0)
    acc = F32(0.0)
    for i in range(lymanP - 1, nu.shape[0]):
        acc = F32(np.float64(acc) + np.float64(phot[i]) * np.float64(widFlx[i]))
    normPhot = F32(F32(PI * acc) * HPLANCK)
    fourPi = F32(F32(4.0) * PI)
    RStar = F32(np.sqrt(F32(F32(LPhot) / F32(F32(fourPi * normPhot) * CRYD))))
    T2 = F32(F32(T) * F32(T))
    T4 = F32(T2 * T2)                                          # **4 = (T*T)*(T*T)
    return F32(F32(F32(F32(fourPi * RStar) * RStar) * SIGMA) * T4)


def load_gas_deck(run_dir: str, share_dir: str, input_file: str = "input.in", recPDF=None, totalLines=None):
    """(Model, tables, Deck) of a gas-only deck.  `tables` carries what K1 needs: xsec (XSecTables),
    bands (flat band list), den (species densities), ionDen, elemAbun, and the pointer dictionaries."""
    d = read_input(os.path.join(run_dir, input_file))
    if not d.lgGas or d.lgDust:
        raise NotImplementedError("gas-only decks (HII40, PN150)")
    if d.contShape != "blackbody":
        raise NotImplementedError(f"contShape {d.contShape}")
    o = d.other
    for k in ("densityFile", "densityLaw", "multiChemistry", "multiGrids", "planeIonization", "fillingFactor"):
        if k in o:
            raise NotImplementedError(f"keyword {k} in a gas deck")
    Hdensity = F32(_real(o["Hdensity"][0]))
    TeStart = F32(_real(o["TeStart"][0])) if "TeStart" in o else F32(10000.0)
    nstages = int(_real(o["nstages"][0])) if "nstages" in o else 7

    def resolve(p):
        for cand in (os.path.join(run_dir, p), os.path.join(run_dir, os.path.basename(p))):
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError(p)

    ab, on, xref, nUsed = G.read_abundances(resolve(d.abundanceFile))
    ph1, ph2 = G.read_ph_tables(os.path.join(share_dir, "data", "ph1.dat"), os.path.join(share_dir, "data", "ph2.dat"))
    ionEdge, nEdges = G.ion_edges(ph1, on, nstages, d.nuMax)
    nu = G.nu_mesh_gas(d.nbins, d.nuMin, d.nuMax, ionEdge, nEdges)
    nbins = d.nbins
    widFlx = wid_flx(nu)
    xt, ptr, xp = G.build_xsec_tables(nu, ph1, ph2, on, xref, nstages)

    if d.edges is None or min(d.edges) < 0:
        raise ValueError("readInput: Grid edges unspecified or non-valid grid edges")
    nx, ny, nz = d.nx, d.ny, d.nz
    xA, yA, zA = (auto_axis(n, e, d.lgSymmetricXYZ) for n, e in zip((nx, ny, nz), d.edges))
    r = F32(1.0e10) * np.sqrt(((xA / F32(1.0e10)) ** 2)[:, None, None] + ((yA / F32(1.0e10)) ** 2)[None, :, None]
                              + ((zA / F32(1.0e10)) ** 2)[None, None, :]).astype(F32)
    mask = ~(r < F32(d.R_in))
    if d.R_out > 0:
        mask &= ~(r > F32(d.R_out))
    active, nCells = number_active(mask)
    g = Grid(xAxis=xA, yAxis=yA, zAxis=zA, active=active, nCells=nCells)
    g.Hden = np.zeros(nCells + 1, dtype=F32)
    g.Hden[1:] = Hdensity
    g.Te = np.zeros(nCells + 1, dtype=F32)
    g.Te[1:] = TeStart
    g.Ne = g.Hden.copy()
    ionDen = G.initial_ion_state(nCells, on, xref, nUsed, nstages)
    elemAbun = np.asfortranarray(ab.reshape(1, 30))
    abIndex = np.ones(nCells + 1, dtype=I32)
    den = xt.species_densities(ionDen, elemAbun, abIndex, g.Hden)
    bands = xt.band_list(nbins)

    # the re-emission tables: the host solver's (emissionDriver); stand-in unless supplied
    kind = "supplied"
    if recPDF is None:
        kind = "stand-in (workloads.recombination_cdf at TeStart; emissionDriver is the host solver's)"
        row = recombination_cdf(nu, widFlx, float(TeStart)).astype(F32)
        recPDF = np.zeros((nCells + 1, nbins), dtype=F32, order="F")
        recPDF[1:, :] = row[None, :]
        totalLines = np.zeros(nCells + 1, dtype=F32)
        totalLines[1:] = F32(0.6)
    g.recPDF, g.totalLines = np.asfortranarray(recPDF, dtype=F32), np.asarray(totalLines, dtype=F32)

    cdf = stellar_cdf(d.TStellar, nu, widFlx)
    pos, sidx = set_star_position([g], [list(d.starPosition)])
    from .model import locate

    lymanP = int(locate(nu, F32(1.0)))                         # continuum_mod.f90:84
    LStar = F32(d.LStar) if d.LStar > 0 else lstar_from_lphot(d.LPhot, d.TStellar, nu, widFlx, lymanP)
    model = Model(grids=[g], nbins=nbins, nuArray=nu,
                  inSpectrumProbDen=np.stack([np.zeros(nbins, F32), cdf]).astype(F32),
                  deltaE=np.asarray([0.0, F32(LStar / F32(d.nPhotons))], dtype=F32),
                  starPosition=np.asarray(pos, dtype=F32), starIndeces=np.asarray(sidx, dtype=I32),
                  lgDust=False, lgGas=True, lgSymmetricXYZ=d.lgSymmetricXYZ, R_out=float(d.R_out),
                  ionEdge1=float(ionEdge[0]))
    tables = dict(xsec=xt, bands=bands, den=den, ionDen=ionDen, elemAbun=elemAbun, abIndex=abIndex, ptr=ptr, xp=xp,
                  widFlx=widFlx, ionEdge=ionEdge[:nEdges], nstages=nstages, LStar=float(LStar), recPDF_kind=kind,
                  ph1=ph1, ph2=ph2, lgElementOn=on, elementXref=xref)
    return model, tables, d


def host_opacity(model: Model, tables: dict) -> np.ndarray:
    """opacity(0:nCells, nbins) of the loaded state assembled on the host in addOpacity's order
    (ionization_mod.f90:349-484 without the free-free term of bin 1, which needs the host's
    BoltGaunt): the numpy twin of K1, for the oracle side of the parity tests."""
    g = model.grids[0]
    b, den, xs = tables["bands"], tables["den"], tables["xsec"].xSecArray
    op = np.zeros((g.nCells + 1, model.nbins), dtype=F32, order="F")
    for k in range(b["species"].shape[0]):
        lo, hi = int(b["low"][k]), min(int(b["high"][k]), model.nbins)
        if hi < lo:
            continue
        col = den[:, int(b["species"][k]) - 1]
        seg = xs[int(b["off"][k]) + lo - 1:int(b["off"][k]) + hi]
        op[:, lo - 1:hi] = (op[:, lo - 1:hi] + (col[:, None] * seg[None, :]).astype(F32)).astype(F32)
    op[0, :] = 0
    return op


# ---------------------------------------------------------------------------------------
# fixture format (tests/golden/deck_HII40.npz, deck_PN150.npz): the decks' files and the atomic
# data do not travel to the GPU box
# ---------------------------------------------------------------------------------------
_XT_SCALARS = ("nstages", "HlevXSecP1", "HlevNuP1", "HeISingXSecP1", "HeIlevNuP1", "HeIIXSecP1", "HeIIlevNuP1")


def gas_deck_to_arrays(model: Model, tables: dict, d) -> dict:
    import json

    g, xt = model.grids[0], tables["xsec"]
    scal = dict(lgSymmetricXYZ=bool(d.lgSymmetricXYZ), nPhotons=int(d.nPhotons), TStellar=float(d.TStellar),
                LPhot=float(d.LPhot), R_in=float(d.R_in), R_out=float(d.R_out), LStar=tables["LStar"],
                Hdensity=float(g.Hden[1]), TeStart=float(g.Te[1]), ionEdge1=float(model.ionEdge1),
                starPosition=list(d.starPosition), recPDF_kind=tables["recPDF_kind"],
                **{k: int(getattr(xt, k)) for k in _XT_SCALARS})
    return dict(nuArray=model.nuArray, xAxis=g.xAxis, yAxis=g.yAxis, zAxis=g.zAxis, active=g.active,
                cdf=model.inSpectrumProbDen[1], deltaE=model.deltaE, xSecArray=xt.xSecArray,
                lgElementOn=xt.lgElementOn, elementXref=xt.elementXref, elementP=xt.elementP, nShells=xt.nShells,
                elemAbun=tables["elemAbun"], recRow=g.recPDF[1], totalLines1=g.totalLines[1:2],
                deck_json=np.frombuffer(json.dumps(scal).encode(), dtype=np.uint8))


def gas_deck_from_arrays(a: dict):
    """Inverse of gas_deck_to_arrays: (Model, tables, scalars) without the deck's files."""
    import json

    s = json.loads(bytes(a["deck_json"]).decode())
    nu = np.asarray(a["nuArray"], dtype=F32)
    nbins = nu.shape[0]
    active = np.asfortranarray(a["active"], dtype=I32)
    g = Grid(xAxis=np.asarray(a["xAxis"], F32), yAxis=np.asarray(a["yAxis"], F32), zAxis=np.asarray(a["zAxis"], F32),
             active=active, nCells=int(active.max()))
    nCells = g.nCells
    from .opacity import XSecTables

    xt = XSecTables(xSecArray=np.asarray(a["xSecArray"], F32), lgElementOn=np.asarray(a["lgElementOn"], I32),
                    elementXref=np.asarray(a["elementXref"], I32), elementP=np.asfortranarray(a["elementP"], dtype=I32),
                    nShells=np.asfortranarray(a["nShells"], dtype=I32), **{k: int(s[k]) for k in _XT_SCALARS})
    g.Hden = np.zeros(nCells + 1, dtype=F32)
    g.Hden[1:] = F32(s["Hdensity"])
    g.Te = np.zeros(nCells + 1, dtype=F32)
    g.Te[1:] = F32(s["TeStart"])
    g.Ne = g.Hden.copy()
    on, xref = xt.lgElementOn, xt.elementXref
    ionDen = G.initial_ion_state(nCells, on, xref, int(on.sum()), xt.nstages)
    elemAbun = np.asfortranarray(a["elemAbun"], dtype=F32)
    abIndex = np.ones(nCells + 1, dtype=I32)
    g.recPDF = np.zeros((nCells + 1, nbins), dtype=F32, order="F")
    g.recPDF[1:, :] = np.asarray(a["recRow"], F32)[None, :]
    g.totalLines = np.zeros(nCells + 1, dtype=F32)
    g.totalLines[1:] = F32(a["totalLines1"][0])
    pos, sidx = set_star_position([g], [list(s["starPosition"])])
    model = Model(grids=[g], nbins=nbins, nuArray=nu,
                  inSpectrumProbDen=np.stack([np.zeros(nbins, F32), np.asarray(a["cdf"], F32)]).astype(F32),
                  deltaE=np.asarray(a["deltaE"], dtype=F32), starPosition=np.asarray(pos, dtype=F32),
                  starIndeces=np.asarray(sidx, dtype=I32), lgDust=False, lgGas=True,
                  lgSymmetricXYZ=bool(s["lgSymmetricXYZ"]), R_out=float(s["R_out"]), ionEdge1=float(s["ionEdge1"]))
    tables = dict(xsec=xt, bands=xt.band_list(nbins), den=xt.species_densities(ionDen, elemAbun, abIndex, g.Hden),
                  ionDen=ionDen, elemAbun=elemAbun, abIndex=abIndex, widFlx=wid_flx(nu), nstages=xt.nstages,
                  LStar=float(s["LStar"]), recPDF_kind=s["recPDF_kind"], lgElementOn=on, elementXref=xref)
    return model, tables, s
