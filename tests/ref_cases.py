"""Small seeded cases on which the CPU oracle is pinned against the reference's own code
(the f90py translation of photon_mod.f90, oracle/f90ref) and against the golden vectors
generated from it (tests/golden/make_golden.py -> tests/golden/ref_*.npz).

Sizes are chosen so that the translated reference (numpy.float32 scalar arithmetic in
Python, ~10^4 cell crossings per second) finishes each case in a few seconds."""
import numpy as np

from mocassin_b200 import workloads as W

SEED = 12345

# name: (builder, kwargs, packets, mode)   mode: "stellar" | ("diffext", gpLoc, cellLoc) | "reslines"
REF_CASES = {
    "hii_sym_gas": (W.hii_region, dict(n=9, nbins=80), 3000, "stellar"),
    "hii_sym_gas_debug": (W.hii_region, dict(n=7, nbins=60, debug=True, seed=9), 2000, "stellar"),
    "dust_shell_hg": (W.dust_shell, dict(n=8, nbins=40, tauV=3.0), 600, "stellar"),
    "dust_shell_iso": (W.dust_shell, dict(tauV=10.0, isotropic=True, n=7, nbins=40), 300, "stellar"),
    "multigrid_sym": (W.multigrid, dict(n=8, nsub=5, nbins=60), 1000, "stellar"),
    "multigrid_nonsym": (W.multigrid, dict(symmetric=False, n=9, nsub=5, nbins=60), 1000, "stellar"),
    "cube_uniform_gas": (W.synthetic_cube, dict(n=10, nbins=60, clumpy=False, dust=False, nPhotons=10**6), 2000, "stellar"),
    "cube_clumpy_gasdust": (W.synthetic_cube, dict(n=10, nbins=60, clumpy=True, dust=True, nPhotons=10**6), 800, "stellar"),
    "viewing_angles": (W.viewing_angles, dict(n=7, nbins=40), 2000, "stellar"),
    "viewing_angles_phifree": (W.viewing_angles, dict(n=7, nbins=40, phi_free=True), 1000, "stellar"),
    "plane_slab_gasdust": (W.plane_slab, dict(nx=5, ny=9, nz=5, nbins=60, Hden=30.0), 1000, "stellar"),
    "plane_slab_gas": (W.plane_slab, dict(nx=5, ny=9, nz=5, nbins=60, dust=False, Hden=30.0), 1000, "stellar"),
    "diffext_mother": (W.multigrid, dict(n=8, nsub=5, nbins=60), 500, ("diffext", 1, (3, 2, 5))),
    "diffext_subgrid": (W.multigrid, dict(n=8, nsub=5, nbins=60), 500, ("diffext", 2, (3, 2, 4))),
    "reslines_multigrid": (W.multigrid, dict(n=8, nsub=5, nbins=60), 0, "reslines"),
    "two_stars": (None, dict(), 500, ("stars", (1, 2))),
}
DIFFEXT_DELTAE = 1.0e-3


def two_stars():
    """lgMultistars: a second, off-centre source with its own spectrum and packet energy in the
    clumpy gas+dust cube; the two sources are transported one after the other (iteration_mod.f90:474-496)"""
    from mocassin_b200.model import star_indices

    F32 = np.float32
    m = W.synthetic_cube(n=9, nbins=40, clumpy=True, dust=True, nPhotons=10 ** 6)
    g = m.grids[0]
    pos = [float(g.xAxis[5]) * 0.9, float(g.yAxis[2]) * 1.1 + 1e15, float(g.zAxis[6])]
    idx = star_indices(g, pos)
    assert g.active[idx[0] - 1, idx[1] - 1, idx[2] - 1] > 0
    m.starPosition = np.vstack([m.starPosition, np.asarray(pos, F32)[None, :]]).astype(F32)
    m.starIndeces = np.vstack([m.starIndeces, np.asarray(idx + [1], np.int32)[None, :]]).astype(np.int32)
    m.deltaE = np.concatenate([m.deltaE, [F32(2e-6)]]).astype(F32)
    row = (m.inSpectrumProbDen[1:2] ** 2).astype(F32)
    row[0, -1] = 1.0
    m.inSpectrumProbDen = np.vstack([m.inSpectrumProbDen, row]).astype(F32)
    m.lgMultistars = True
    return m


def make(name):
    fn, kw, n, mode = REF_CASES[name]
    m = fn(**kw) if fn is not None else two_stars()
    if isinstance(mode, tuple):
        m.inSpectrumProbDen[0, :] = W.blackbody_cdf(20000.0, m.nuArray, np.gradient(m.nuArray).astype(np.float32))
        m.deltaE[0] = DIFFEXT_DELTAE
    if mode == "reslines":
        rs = np.random.default_rng(3)
        for g in m.grids:
            g.resLinePackets = np.zeros(g.nCells + 1, np.int32)
            g.resLinePackets[1:] = rs.integers(0, 3, g.nCells)
    return m, n, mode


def _collect(out, counters, fates, plane):
    res = {"Qphot": np.float32(counters["Qphot"]), "absInt": np.float32(counters["absInt"]),
           "scaInt": np.float32(counters["scaInt"]), "plane": np.asarray(plane, np.int64)}
    if fates is not None:
        res["fates"] = np.asarray(fates)[:, :2].astype(np.int32)      # segments, energyPacketRun calls
    for i, o in enumerate(out):
        for k in ("Jste", "escapedPackets", "Jdif", "linePackets"):
            if k in o:
                res[f"{k}_g{i + 1}"] = np.asarray(o[k], np.float32)
    return res


def _stars(run, stars):
    """one energyPacketDriver call per source into the same tallies; per-packet records concatenated,
    Qphot (reset by every call, photon_mod.f90:89) of the last source, absInt/scaInt/nSegments summed"""
    cs, fs = [], []
    for i in stars:
        c, f = run(i)
        cs.append(c); fs.append(np.asarray(f))
    c = dict(cs[-1])
    for k in ("absInt", "scaInt"):
        c[k] = np.float32(sum(np.float32(x[k]) for x in cs))
    c["nSegments"] = sum(int(x["nSegments"]) for x in cs)
    return c, np.concatenate(fs, axis=0)


def run_oracle(name):
    """the C oracle's faithful float32 tallies and per-packet records"""
    from oracle.oracle import Oracle

    m, n, mode = make(name)
    o = Oracle(m)
    if mode == "stellar":
        c, f = o.transport(1, 0, n, seed=SEED, want_fates=True)
    elif mode[0] == "stars":
        c, f = _stars(lambda i: o.transport(i, 0, n, seed=SEED, want_fates=True), mode[1])
    elif mode == "reslines":
        c, nrun = o.transport_reslines(1, seed=SEED)
        f = None
        c = dict(c, nRun=nrun)
    else:
        _, gp, cell = mode
        c, f = o.transport(0, 0, n, seed=SEED, gpLoc=gp, cellLoc=cell, want_fates=True)
    res = _collect(o.out, c, f, o.planeIonDistribution)
    res["nSegments"] = np.int64(c["nSegments"])
    return res


def run_reference(name, math="detmath", uninit_int=0):
    """the reference's own energyPacketDriver (translated), same Philox streams"""
    from oracle import oracle as orc
    from oracle.f90ref.harness import Reference

    m, n, mode = make(name)
    r = Reference(m, orc.load(), math=math, uninit_int=uninit_int)
    if mode == "stellar":
        c, f = r.transport(1, 0, n, seed=SEED)
    elif mode[0] == "stars":
        c, f = _stars(lambda i: r.transport(i, 0, n, seed=SEED), mode[1])
    elif mode == "reslines":
        c, f = r.transport_reslines(1, seed=SEED)
    else:
        _, gp, cell = mode
        g = m.grids[gp - 1]
        # the reference derives deltaE(0) = LdiffuseLoc(cell)/NphotonsDiffuseLoc (photon_mod.f90:64-66)
        r.grid[gp].ldiffuseloc[int(g.active[cell[0] - 1, cell[1] - 1, cell[2] - 1])] = np.float32(DIFFEXT_DELTAE)
        c, f = r.transport(0, 0, n, seed=SEED, gpLoc=gp, cellLoc=cell)
    res = _collect(r.out, c, f, r.G.planeiondistribution.a)
    res["draws"] = f[:, 2].astype(np.int32)
    res["nSegments"] = np.int64(c["nSegments"])
    return res


# ---------------------------------------------------------------------------------------------
# the callers either side of the transport: opacity block of iterateMC (K1), dust closure (K5, K6)
# ---------------------------------------------------------------------------------------------
AUX_CASES = ["opacity_multichem", "opacity_singlechem", "dust_closure", "dust_closure_debug"]
CONTBOLTZ1, GAUNTFF1, BREMS_P = 0.5, 1.1, 7      # what the harness's BoltGaunt stand-in sets for bin 1


def _opacity_inputs(multi):
    import types

    import opacity_case
    from mocassin_b200.model import number_active

    c = opacity_case.make(nCells=60, nbins=140, seed=5, multi_chem=multi)
    n = c["nCells"]
    mask = np.zeros((4, 4, 4), bool)
    mask.reshape(-1)[:n] = True
    active, nc = number_active(mask)
    assert nc == n and c["t"].band_list(c["nbins"])["low"].min() >= 2
    rng = np.random.default_rng(1)
    c["Ne"] = (rng.random(n + 1) * 1e12 + 1).astype(np.float32)
    c["Te"] = (rng.random(n + 1) * 1e4 + 5e3).astype(np.float32)
    c["active"] = active
    c["dm"] = types.SimpleNamespace(**c["dust_model"])
    return c


def _dust_inputs(debug):
    model, t = W.dust_closure(n=6, nbins=60, nPhotons=20000)
    g = model.grids[0]
    rng = np.random.default_rng(3)
    g.Tdust[1:, 1:, 1:] = rng.uniform(20.0, 1300.0, size=g.Tdust[1:, 1:, 1:].shape).astype(np.float32)
    g.Tdust[1, 1, 1:] = np.float32(300.0)
    g.Tdust[:, :, 5] = np.float32(2000.0)                 # one cell with every grain sublimed (0/0 row)
    g.Tdust[0, 0, 1:] = rng.uniform(50.0, 900.0, g.nCells).astype(np.float32)
    # radiation field: black-body baths of various temperatures (some close to the old mean
    # temperature -> converged), a few cells no packet crossed, a few far above the table
    nu = model.nuArray.astype(np.float64)
    J = np.zeros((g.nCells + 1, model.nbins), np.float32, order="F")
    for c in range(1, g.nCells + 1):
        Tb = float(g.Tdust[0, 0, c]) * (1.0 + 0.04 * rng.standard_normal()) if c % 3 else rng.uniform(30, 2500)
        x = np.minimum(157893.94 * nu / Tb, 80.0)
        J[c, :] = (4.0 * np.pi * 0.5250229 * nu ** 3 / np.expm1(x) * 3.28984e15 * np.asarray(t["widFlx"], np.float64)).astype(np.float32)
    J[rng.random(g.nCells + 1) < 0.12, :] = 0.0
    J[7, :] *= np.float32(1.0e6)
    Jd = None
    if debug:
        Jd = (J * np.float32(0.25)).astype(np.float32)
        Jd[9, :] = J[11, :]
        J[9, :] = 0.0                                     # hit only through Jdif
    return model, g, t, J, Jd


def run_oracle_aux(name):
    from oracle import oracle as O

    if name.startswith("opacity"):
        c = _opacity_inputs(name == "opacity_multichem")
        ff1 = dict(np.load(_gold(name)))["ff1"]          # FFOpacity(1) is an input of the oracle (BoltGaunt stays on the host)
        op, sca, ab = O.opacity(c["t"], c["nbins"], c["ionDen"], c["elemAbun"], c["abIndex"], c["Hden"], ff1=ff1,
                                dust=c["dust"], model=c["dm"])
        return dict(opacity=op, scaOpac=sca, absOpac=ab)
    model, g, t, J, Jd = _dust_inputs(name.endswith("debug"))
    model.lgDebug = Jd is not None
    pdf = O.dust_pdf(model, g, t)
    T, conv = O.dust_update(model, g, t, J, 0.05, Jdif=Jd)
    return dict(dustPDF=pdf, Tdust=T, lgConverged=conv)


def run_reference_aux(name):
    from oracle import oracle as O
    from oracle.f90ref.harness_aux import AuxReference

    a = AuxReference(O.load())
    if name.startswith("opacity"):
        c = _opacity_inputs(name == "opacity_multichem")
        args = (c["t"], c["nbins"], c["ionDen"], c["elemAbun"], c["abIndex"], c["Hden"], c["active"], CONTBOLTZ1, GAUNTFF1,
                c["Ne"], c["Te"], BREMS_P)
        op, sca, ab = a.opacity_block(*args, dust=c["dust"], dust_model=c["dm"])
        _, ff1 = AuxReference(O.load()).gas_opacity(*args)
        return dict(opacity=op, scaOpac=sca, absOpac=ab, ff1=ff1)
    model, g, t, J, Jd = _dust_inputs(name.endswith("debug"))
    pdf = a.dust_pdf(model, g, t)
    T, conv = a.dust_update(model, g, t, J, 0.05, Jdif=Jd)
    return dict(dustPDF=pdf, Tdust=np.array(T), lgConverged=conv)


def _gold(name):
    import os

    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_aux_{name}.npz")


# ---------------------------------------------------------------------------------------------
# photo-rate pre-integration (K8): updateCell's nPhotoSte/nPhotoDif and thermBalance's heatSte
# ---------------------------------------------------------------------------------------------
PHOTO_CASES = ["photo", "photo_debug"]


def _photo_inputs(debug):
    from mocassin_b200.model import number_active
    from mocassin_b200.opacity import XSecTables

    rng = np.random.default_rng(41)
    nbins, nstages, n = 70, 5, 20
    on = np.zeros(30, np.int32)
    for el in (1, 2, 6, 8, 20, 26):
        on[el - 1] = 1
    xref = np.zeros(30, np.int32)
    xref[on > 0] = np.arange(1, int(on.sum()) + 1)
    xs = [0.0, 0.0, 0.0]

    def push(k, holes=True):
        start = len(xs) + 1
        v = (1e-18 * rng.lognormal(0.0, 2.0, k) * (np.arange(1, k + 1) ** -3.0) * 50).astype(np.float32)
        if holes and k > 6:
            v[k // 2] = np.float32(1e-37)          # thermBalance leaves its loop here, updateCell treats it as 0
        xs.extend(v.tolist())
        return start

    tabs = dict(HlevNuP1=12, HeIlevNuP1=25, HeIIlevNuP1=40)
    tabs["HlevXSecP1"] = push(nbins - 12 + 1, holes=False)
    tabs["HeISingXSecP1"] = push(nbins - 25 + 1)
    tabs["HeIIXSecP1"] = push(nbins - 40 + 1, holes=False)
    elementP = np.zeros((30, 30, 7, 3), np.int32, order="F")
    for el in range(3, 31):
        if on[el - 1]:
            for ion in range(1, min(el, nstages) + 1):
                for sh in range(1, 8):
                    lo = int(rng.integers(2, nbins - 8))
                    hi = int(rng.integers(lo + 1, nbins + 1))
                    elementP[el - 1, ion - 1, sh - 1, :] = (lo, hi, push(hi - lo + 1, holes=bool(rng.integers(0, 2))))
    t = XSecTables(xSecArray=np.array(xs, np.float32), nstages=nstages, lgElementOn=on, elementXref=xref,
                   elementP=elementP, nShells=np.full((30, 30), 7, np.int32, order="F"), **tabs)
    mask = np.zeros((3, 3, 3), bool)
    mask.reshape(-1)[:n] = True
    active, nc = number_active(mask)
    nu = np.geomspace(0.05, 40.0, nbins).astype(np.float32)
    J = (rng.lognormal(-8.0, 2.0, (n + 1, nbins))).astype(np.float32, order="F")
    J[rng.random(J.shape) < 0.2] = 0.0
    J[0] = 0.0
    Jd = (J * np.float32(0.3)).astype(np.float32, order="F")[::-1].copy(order="F") if debug else None
    if Jd is not None:
        Jd[0] = 0.0
    ionDen = np.asfortranarray(rng.random((n + 1, int(on.sum()), nstages)).astype(np.float32))
    elemAbun = np.asfortranarray((rng.random((2, 30)) * 1e-3).astype(np.float32))
    abIndex = rng.integers(1, 3, n + 1).astype(np.int32)
    return dict(t=t, nbins=nbins, nu=nu, active=active, J=J, Jd=Jd, ionDen=ionDen, elemAbun=elemAbun, abIndex=abIndex)


def run_reference_photo(name):
    from oracle import oracle as O
    from oracle.f90ref.harness_aux import AuxReference

    c = _photo_inputs(name.endswith("debug"))
    return AuxReference(O.load()).photo(c["t"], c["nbins"], c["nu"], c["active"], c["J"], c["Jd"], c["ionDen"],
                                        c["elemAbun"], c["abIndex"])


def run_oracle_photo(name, outShell):
    """oracle_photo_integrals on the band list a host builds from the reference's pointer tables
    (one band per element and ion: outer shell from getOuterShell; update_mod.f90:175-204 for the
    rates, :1125-1157 for the heating, whose upper limit is one bin higher), then thermBalance's
    last step, heatSte = sum over ions of heat*ionDen*elemAbun (:1222-1231), in float32."""
    from oracle import oracle as O

    c = _photo_inputs(name.endswith("debug"))
    t, nb = c["t"], c["nbins"]
    ions, off, low, hiR, hiH = [], [], [], [], []
    for el in range(1, 31):
        if not t.lgElementOn[el - 1]:
            continue
        for ion in range(1, min(el, t.nstages - 1) + 1):
            if el == 1:
                o, l, h1, h2 = t.HlevXSecP1, t.HlevNuP1, nb, nb
            elif el == 2 and ion == 1:
                o, l, h1, h2 = t.HeISingXSecP1, t.HeIlevNuP1, nb, nb
            elif el == 2:
                o, l, h1, h2 = t.HeIIXSecP1, t.HeIIlevNuP1, nb, nb
            else:
                sh = int(outShell[el - 1, ion - 1])
                l, h, o = (int(v) for v in t.elementP[el - 1, ion - 1, sh - 1, :])
                h1, h2 = h - 1, h
            ions.append((el, ion)); off.append(o); low.append(l); hiR.append(h1); hiH.append(h2)
    nR = c["J"].shape[0]
    res = dict(nPhotoSte=np.zeros((nR, 30, t.nstages), np.float32), nPhotoDif=np.zeros((nR, 30, t.nstages), np.float32),
               heatSte=np.zeros(nR, np.float32), heatDif=np.zeros(nR, np.float32))
    for key, hkey, J in (("nPhotoSte", "heatSte", c["J"]), ("nPhotoDif", "heatDif", c["Jd"])):
        if J is None:
            res[key][1:] = np.float32(0.0)
            continue
        nP, _ = O.photo_integrals(nb, off, low, hiR, t.xSecArray, c["nu"], J)
        _, ht = O.photo_integrals(nb, off, low, hiH, t.xSecArray, c["nu"], J)
        res[key][:] = np.float32(1.0e-20)
        for b, (el, ion) in enumerate(ions):
            res[key][:, el - 1, ion - 1] = nP[:, b]
        ab = c["elemAbun"][np.maximum(c["abIndex"], 1) - 1]
        tot = np.zeros(nR, np.float32)
        for b, (el, ion) in enumerate(ions):
            h = (ht[:, b] * c["ionDen"][:, int(t.elementXref[el - 1]) - 1, ion - 1]).astype(np.float32)
            tot = (tot + (h * ab[:, el - 1]).astype(np.float32)).astype(np.float32)
        res[hkey][:] = tot
    res["nPhotoSte"][0] = 0
    res["nPhotoDif"][0] = 0
    res["heatSte"][0] = 0
    res["heatDif"][0] = 0
    return res


# ---------------------------------------------------------------------------------------------
# writeSED (K7 + the host scaling of mocassin_b200/output.py)
# ---------------------------------------------------------------------------------------------
SED_CASES = ["viewing_angles", "viewing_angles_phifree", "multigrid_sym", "hii_sym_gas"]


def sed_inputs(name):
    """model, widFlx, the reference's escapedPackets of the transport golden file, and `raw` = its
    sequential float32 sum over cells (the order writeSED adds in: grids, then cells)"""
    import os

    m, n, mode = make(name)
    gold = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_{name}.npz")))
    esc = [gold[f"escapedPackets_g{i + 1}"] for i in range(m.nGrids)]
    raw = np.zeros((m.nbins, m.nAngleBins + 1), np.float32)
    for e in esc:
        for i in range(e.shape[0]):
            raw = (raw + e[i, 1:, :]).astype(np.float32)
    return m, W.wid_flx(m.nuArray), esc, raw


def run_reference_sed(name):
    from oracle import oracle as O
    from oracle.f90ref.harness_aux import AuxReference

    m, wid, esc, raw = sed_inputs(name)
    # the host divides escapedPackets by 8 for symmetricXYZ before writeSED (iteration_mod.f90:719);
    # writeSED is host code: COS is the platform's (numpy's), not detmath
    scaled = [(e / np.float32(8.0)).astype(np.float32) if m.lgSymmetricXYZ else e for e in esc]
    sed, tot, rows = AuxReference(O.load(), math="libm").write_sed(m, wid, scaled)
    return dict(SED=sed, totalE=np.float32(tot), nu=np.array([r[0] for r in rows], np.float32),
                lambda_um=np.array([r[1] for r in rows], np.float32))


# ---------------------------------------------------------------------------------------------
# dust optics of the input side (mocassin_b200/deck.py): BHmie, getQs, linearMap, makeDustXsec's assembly
# ---------------------------------------------------------------------------------------------
def mie_inputs():
    """seeded (x, m) pairs for BHmie; optical constants on a 40-bin mesh, radii, weights and
    abundances for getQs and the assembly (2 species x 3 sizes)"""
    rng = np.random.default_rng(20261017)
    x = np.concatenate([10 ** rng.uniform(-2.5, 2.0, 150), [100.0, 0.5, 1.0e-3]]).astype(np.float32)
    m = (rng.uniform(0.4, 3.2, x.shape[0]) + 1j * 10 ** rng.uniform(-4.5, 0.6, x.shape[0])).astype(np.complex64)
    nu = (10 ** np.linspace(-4.0, 1.0, 40)).astype(np.float32)
    Ere = rng.uniform(0.6, 3.0, (2, 40)).astype(np.float32)
    Eim = (10 ** rng.uniform(-3.0, 0.4, (2, 40))).astype(np.float32)
    radius = np.array([0.005, 0.12, 1.5], np.float32)
    weight = np.array([0.7, 0.25, 0.05], np.float32)
    abun = np.array([0.6, 0.4], np.float32)
    # a table for linearMap: descending-wavelength optical constants onto the mesh, with points
    # outside the table on both sides
    xt = np.sort(10 ** rng.uniform(-3.0, 0.5, 25)).astype(np.float32)
    yt = rng.uniform(0.5, 3.0, 25).astype(np.float32)
    # an MRN-like size grid (10 sizes, weights ~ a^-3.5) for the weight normalisation
    sizes = (0.04 * (0.4 / 0.04) ** (np.arange(10) / 9.0)).astype(np.float32)
    return dict(x=x, m=m, nu=nu, Ere=Ere, Eim=Eim, radius=radius, weight=weight, abun=abun, xt=xt, yt=yt,
                sizes=sizes, size_weights=(sizes.astype(np.float64) ** -3.5).astype(np.float32))


def run_reference_mie():
    from oracle import oracle as O
    from oracle.f90ref.harness_aux import AuxReference

    I = mie_inputs()
    A = AuxReference(O.load(), math="libm")
    q = np.array([A.bhmie(x, m) for x, m in zip(I["x"], I["m"])], np.float32)
    Qa, Qs, G = (np.zeros((2, 3, 40), np.float32) for _ in range(3))
    for s in range(2):
        Qa[s], Qs[s], G[s] = A.get_qs(I["Ere"][s], I["Eim"][s], I["radius"], I["nu"])
    asm = A.dust_xsec_assembly(Qs, Qa, G, I["radius"], I["weight"], I["abun"], 40)
    wid = W.wid_flx(I["nu"])
    out = dict(bhmie=q, Qabs=Qa, Qsca=Qs, gCos=G, mapped=A.linear_map(I["yt"], I["xt"], I["nu"]),
               **{"asm_" + k: np.asarray(v) for k, v in asm.items()})
    # the stellar CDF (getFlux + setProbDen) at four temperatures, and dustEmissionInt for the first
    # 80 K of the assembled cross-sections (nTemps is 3000 in the reference: Python loops)
    for T in MIE_TSTAR:
        out[f"flux_{int(T)}"], out[f"cdf_{int(T)}"] = A.stellar_cdf(T, I["nu"], wid)
    out["grain_weights"] = A.grain_weights(I["sizes"], I["size_weights"])
    out["emint"] = A.dust_emission_int(asm["xSecArray"], asm["dustAbsXsecP"][1:, :], I["nu"], wid, MIE_NTEMPS)
    return out


MIE_TSTAR = (2500.0, 5800.0, 40000.0, 150000.0)
MIE_NTEMPS = 80


# ---------------------------------------------------------------------------------------------
# setStarPosition (mocassin_b200/model.py: set_star_position)
# ---------------------------------------------------------------------------------------------
def starpos_inputs():
    """{case: (grids, relative positions)}: multi-grid symmetric / non-symmetric and a plain cube;
    each list holds a star in the sub-grid followed by stars elsewhere (the order matters: the
    reference keeps the sub-grid's extents for the stars that follow)."""
    rng = np.random.default_rng(77)
    out = {}
    for name, m in (("multigrid_sym", W.multigrid()), ("multigrid_nonsym", W.multigrid(symmetric=False, n=15)),
                    ("cube", W.synthetic_cube(n=12, nbins=40, nPhotons=10))):
        lo = 0.0 if m.lgSymmetricXYZ else -0.9
        rel = [list(rng.uniform(lo, 0.9, 3)) for _ in range(6)] + [[0.0, 0.0, 0.0]] + [list(rng.uniform(lo, 0.9, 3)) for _ in range(20)]
        out[name] = (m.grids, np.array(rel, np.float32))
    return out


def run_reference_starpos():
    from oracle import oracle as O
    from oracle.f90ref.harness_aux import AuxReference

    A = AuxReference(O.load(), math="libm")
    res = {}
    for name, (grids, rel) in starpos_inputs().items():
        res[name + "_pos"], res[name + "_idx"] = A.set_star_position(grids, rel.tolist())
    for name, (g, sym, cells) in volume_inputs().items():
        res["vol_" + name] = A.get_volume(g, sym, cells)
    for name, (vt, vp, sym) in angle_inputs().items():
        for k, v in A.angle_tables(vt, vp, sym).items():
            res[f"ang_{name}_{k}"] = np.asarray(v)
    for n, R, sym in AXES_CASES:
        res[f"axes_{n}_{int(sym)}"] = np.stack(A.fill_axes(n, R, sym))
    for name, (grids, sym) in mask_inputs().items():
        act, geo = A.fill_mask(grids, sym)
        for i, a in enumerate(act):
            res[f"mask_{name}_g{i + 1}"] = a
        res[f"geo_{name}"] = np.array(geo, np.float32)
    for name, c in active_inputs().items():
        res["act_" + name], n = A.active_cells(c["x"], c["y"], c["z"], c["Hden"], c["Ndust"], c["lgGas"], c["lgDust"], c["R_in"], c["R_out"])
        res["nact_" + name] = np.int32(n)
    return res


AXES_CASES = ((13, 1.46e19, True), (15, 3.0e18, False), (16, 1.0e18, True))


def mask_inputs():
    """{case: ([mother, sub-grid] numbered as setMotherGrid leaves them, lgSymmetricXYZ)}"""
    from mocassin_b200.model import Grid, number_active

    out = {}
    for sym in (True, False):
        m = W.multigrid(symmetric=sym, n=16 if sym else 15)
        gm, gs = m.grids
        r = np.sqrt(gm.xAxis.astype(np.float64)[:, None, None] ** 2 + gm.yAxis.astype(np.float64)[None, :, None] ** 2
                    + gm.zAxis.astype(np.float64)[None, None, :] ** 2)
        act, nc = number_active((r >= 1e15) & (r <= 1e18))
        out["sym" if sym else "nonsym"] = ([Grid(xAxis=gm.xAxis, yAxis=gm.yAxis, zAxis=gm.zAxis, active=act, nCells=nc),
                                            Grid(xAxis=gs.xAxis, yAxis=gs.yAxis, zAxis=gs.zAxis, active=gs.active.copy(), nCells=gs.nCells,
                                                 motherP=1)], sym)
    return out


def active_inputs():
    """{case: axes, density fields, flags, radii}: the shipped deck p0tau10 (dust only, log axes from
    its Ndust file, through the fixture), an HII40-like gas shell on automatic axes, and a gas+dust
    cube with holes in both fields"""
    import os
    from mocassin_b200 import deck
    from mocassin_b200.model import auto_axis

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deck_p0tau10.npz")
    m, t, d = deck.deck_from_arrays(dict(np.load(gold)))
    g = m.grids[0]
    nd3 = np.where(g.active > 0, g.Ndust[np.maximum(g.active, 0)], 0.0).astype(np.float32)
    ax = auto_axis(13, 1.46e19, True)
    rng = np.random.default_rng(3)
    cx = auto_axis(9, 1.0e17, False)
    H = np.where(rng.random((9, 9, 9)) < 0.7, 50.0, 0.0).astype(np.float32)
    N = np.where(rng.random((9, 9, 9)) < 0.5, 1.0e-9, 0.0).astype(np.float32)
    return {"deck_p0tau10": dict(x=g.xAxis, y=g.yAxis, z=g.zAxis, Hden=np.zeros_like(nd3), Ndust=nd3, lgGas=False, lgDust=True,
                                 R_in=d.R_in, R_out=d.R_out, want=g.active),
            "gas_shell": dict(x=ax, y=ax, z=ax, Hden=np.full((13, 13, 13), 100.0, np.float32), Ndust=np.zeros((13, 13, 13), np.float32),
                              lgGas=True, lgDust=False, R_in=3.0e18, R_out=1.46e19),
            "gasdust_holes": dict(x=cx, y=cx, z=cx, Hden=H, Ndust=N, lgGas=True, lgDust=True, R_in=2.0e16, R_out=0.0)}


def angle_inputs():
    """{case: (viewPointTheta(0:n), viewPointPhi(0:n), lgSymmetricXYZ)}: two angles with phi, the same
    phi-free (as the 2-D disk decks: `inclination 2 0.218 -1. 1.35 -1.`), one angle, none"""
    f = np.float32
    return {"two": (f([0, 0.6981317, 1.9]), f([0, 3.4906585, 0.2]), False),
            "phifree": (f([0, 0.218, 1.35]), f([0, -1.0, -1.0]), True),
            "one": (f([0, 1.2]), f([0, 6.0]), True),
            "none": (f([0]), f([0]), True)}


def volume_inputs():
    """{case: (grid, symmetric, [(xP, yP, zP)])}: the log-spaced axes of the shipped p0tau100 deck
    (symmetric: half first cells) and a uniform non-symmetric cube, 300 active cells each incl. corners"""
    import os
    from mocassin_b200 import deck

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deck_p0tau100.npz")
    m, t, d = deck.deck_from_arrays(dict(np.load(gold)))
    out = {}
    for name, g, sym in (("p0tau100", m.grids[0], True), ("cube", W.synthetic_cube(n=9, nbins=40, nPhotons=10).grids[0], False)):
        idx = np.argwhere(np.asarray(g.active) > 0)
        sel = idx[np.random.default_rng(5).choice(len(idx), 300, replace=False)]
        cells = [(int(a) + 1, int(b) + 1, int(c) + 1) for a, b, c in sel] + [(g.nx, g.ny, g.nz), (1, g.ny, 1)]
        out[name] = (g, sym, cells)
    return out


# ---------------------------------------------------------------------------------------------
# writeTauNu / integratePathTauNu (mocassin_b200/output.py: tau_path, tau_nu)
# ---------------------------------------------------------------------------------------------
TAUNU_CASES = ["hii_sym_gas", "cube_clumpy_gasdust", "dust_shell_hg", "viewing_angles"]


def run_reference_taunu(name):
    from oracle import oracle as O
    from oracle.f90ref.harness_aux import AuxReference

    m, n, mode = make(name)
    taus, lam = AuxReference(O.load(), math="libm").write_tau_nu(m)
    return dict(tau_x=taus[0], tau_z=taus[1], tau_y=taus[2], lambda_um=lam)


# ---------------------------------------------------------------------------------------------
# writeContCube (K9 + the host scaling of mocassin_b200/output.py)
# ---------------------------------------------------------------------------------------------
def contcube_inputs(name):
    """model, the reference's escapedPackets per grid (transport golden file), `raws` = per grid the
    sequential float32 sum over freq = 1..nbins (the order writeContCube adds in), and an origin
    cell: an inactive mother-grid cell where there is one (the reference then reads row 0)"""
    m, wid, esc, _ = sed_inputs(name)
    raws = []
    for e in esc:
        c = np.zeros((e.shape[0], e.shape[2]), np.float32)
        for f in range(1, m.nbins + 1):
            c = (c + e[:, f, :]).astype(np.float32)
        raws.append(c)
    z = np.argwhere(np.asarray(m.grids[0].active) == 0)
    origin = tuple(int(v) + 1 for v in z[0]) if len(z) else (1, 1, 1)
    return m, esc, raws, origin


def run_reference_contcube(name):
    from oracle import oracle as O
    from oracle.f90ref.harness_aux import AuxReference

    m, esc, raws, origin = contcube_inputs(name)
    scaled = [(e / np.float32(8.0)).astype(np.float32) if m.lgSymmetricXYZ else e for e in esc]
    rows = AuxReference(O.load(), math="libm").write_cont_cube(m, scaled, 1.0, 10.0, origin=origin)
    idx = np.array([[int(v) for v in r[:4]] for r in rows], np.int32)
    val = np.array([np.asarray(r[4:], np.float32).ravel() for r in rows], np.float32)
    return dict(index=idx, contI=val, origin=np.array(origin, np.int32))


# ---------------------------------------------------------------------------------------------
# writeGrid: the checkpoint files grid0-3.out, dustGrid.out, photoSource.out
# ---------------------------------------------------------------------------------------------
GRID_FILES = {21: "grid0.out", 20: "grid1.out", 30: "grid2.out", 40: "grid3.out", 50: "dustGrid.out", 42: "photoSource.out"}


def writegrid_inputs(lg2D=False):
    from mocassin_b200 import checkpoint as ck

    F32 = np.float32
    m = W.multigrid(n=6, nsub=4, nbins=20)
    m.nAngleBins = 2
    m.viewPointTheta = np.array([0, 0.5, 1.9], F32)
    m.viewPointPhi = np.array([0, 0.7, 3.4], F32)
    rng = np.random.default_rng(2)
    on = np.zeros(30, np.int32)
    for e in (1, 2, 6, 8):
        on[e - 1] = 1
    xref = np.zeros(30, np.int32)
    xref[on > 0] = np.arange(1, 5)
    rp = ck.RunParams(abundanceFile=("abun/solar.dat",), dustSpeciesFile=("dust/sil.dat",), dustFile2="sizes.dat", nstages=5,
                      maxPhotons=10 ** 7, lgAutoPackets=True, convIncPercent=40.0, nPhotIncrease=2.0, lg2D=bool(lg2D))
    state = dict(lgConverged=[rng.integers(0, 2, g.nCells + 1) for g in m.grids],
                 lgBlack=[rng.integers(0, 2, g.nCells + 1) for g in m.grids],
                 Te=[(rng.random(g.nCells + 1) * 1e4).astype(F32) for g in m.grids],
                 Ne=[(rng.random(g.nCells + 1) * 1e3).astype(F32) for g in m.grids],
                 ionDen=[np.asfortranarray(rng.random((g.nCells + 1, 4, 5)).astype(F32)) for g in m.grids],
                 abFileIndex=[np.asfortranarray(rng.integers(1, 3, g.active.shape)) for g in m.grids],
                 lgElementOn=on, elementXref=xref, contShape=["blackbody"], TStellar=[80000.0], LStar=[1.0],
                 nPhotons=[1000000], spID=["mocassin"], tStep=[0.0], lgMultiChemistry=True, totalDustMass=1.5)
    for g in m.grids:
        g.Hden = (rng.random(g.nCells + 1) * 100).astype(F32)
        g.Ndust = (rng.random(g.nCells + 1) * 1e-9).astype(F32)
        g.dustAbunIndex = np.ones(g.nCells + 1, np.int32)
    return m, rp, state


def _norm(s):
    return " ".join(s.split())


def run_reference_writegrid(lg2D=False):
    """{file name: [normalised record lines]} from the reference's own writeGrid; values are
    rendered with checkpoint.py's number format (list-directed formatting is the compiler's).
    lg2D: the 2D flag set -- only plane j = 1 of the mother grid is written (grid_mod.f90:2709-2713)."""
    from mocassin_b200 import checkpoint as ck
    from oracle import oracle as O
    from oracle.f90ref.harness_aux import AuxReference

    m, rp, state = writegrid_inputs(lg2D)
    recs = AuxReference(O.load()).write_grid(m, rp, state)
    out = {}
    for unit, fn in GRID_FILES.items():
        lines = []
        for r in recs[unit]:
            items = [(v.rstrip() if isinstance(v, str) and v.strip() else v) for v in r]
            lines.append(_norm(ck._join(items)))
        out[fn] = lines
    return out


def run_writers(outdir, lg2D=False):
    from mocassin_b200 import checkpoint as ck

    m, rp, s = writegrid_inputs(lg2D)
    p = lambda f: f"{outdir}/{f}"
    ck.write_grid0(p("grid0.out"), m, lgConverged=s["lgConverged"], lgBlack=s["lgBlack"], lg2D=lg2D)
    ck.write_grid1(p("grid1.out"), m, s["Te"], s["Ne"], abFileIndex=s["abFileIndex"], lg2D=lg2D)
    ck.write_grid2(p("grid2.out"), m, s["ionDen"], s["lgElementOn"], s["elementXref"], rp.nstages, lg2D=lg2D)
    ck.write_grid3(p("grid3.out"), m, rp)
    ck.write_dust_grid(p("dustGrid.out"), m, lgMultiChemistry=True, totalDustMass=s["totalDustMass"], lg2D=lg2D)
    ck.write_photo_source(p("photoSource.out"), m, s["contShape"], s["TStellar"], s["LStar"], s["nPhotons"], s["spID"], s["tStep"])
    return {fn: [_norm(l) for l in open(p(fn)).read().splitlines()] for fn in GRID_FILES.values()}


# ---------------------------------------------------------------------------------------------
# per-element bound of the reference's float32 running sum against the fixed-point tally
# ---------------------------------------------------------------------------------------------
def j_error_bound(m, mode, n, unit_of):
    """For every (cell, nu) element of Jste (and Jdif) of every grid: how far the folded fixed-point
    tally may lie from the reference's sequential float32 sum.  With n_i segments added to element i,
    each rounded to the unit u of the tally, and J the value itself:
        |J_fixed - J_ref| <= n_i * u/2 * deltaE/dV(cell)  +  (n_i + 8) * 2^-24 * J
    (quantisation of the n_i path lengths; one float32 rounding per addition of the running sum, two
    per term, four in the fold).  n_i comes from the oracle, whose histories equal the reference's
    packet for packet (tests/test_reference_pin.py).  Returns bound(iG, name, got, want) -> array."""
    from oracle.oracle import Oracle

    o = Oracle(m, count_segments=True, fp32_tallies=False)
    if mode == "stellar":
        o.transport(1, 0, n, seed=SEED)
        dE = float(m.deltaE[1])
    elif mode[0] == "stars":
        for iStar in mode[1]:
            o.transport(iStar, 0, n, seed=SEED)
        dE = max(float(m.deltaE[i]) for i in mode[1])
    elif mode == "reslines":
        o.transport_reslines(1, seed=SEED)
        dE = float(m.deltaE[1])
    else:
        o.transport(0, 0, n, seed=SEED, gpLoc=mode[1], cellLoc=list(mode[2]))
        dE = float(m.deltaE[0])

    def bound(iG, name, got, want):
        N = o.out[iG - 1]["JsteN" if name == "Jste" else "JdifN"][1:].astype(np.float64)
        dV = m.grids[iG - 1].cell_volumes(m.lgSymmetricXYZ).astype(np.float64)[1:, None]
        u = float(unit_of(iG))
        return N * 0.5 * u * dE / dV * (1.0 + 1e-6) + (N + 8.0) * 2.0 ** -24 * np.maximum(got, want)

    return bound
