"""Randomised comparison of the C oracle with the reference's own code (the f90py translation,
oracle/f90ref) -- TEST INFRASTRUCTURE, runs where /root/reference exists.

    python tests/tools/fuzz_reference.py [--trials N] [--seed S] [--packets P]

Each trial builds one of the seeded workloads with random parameters, then perturbs what the
builders keep regular: non-uniform axes, per-(cell, nu) opacity scatter over several decades
(including exact zeros), random scattering albedo, random re-emission CDFs and line fractions,
random asymmetry parameters, sublimed grain species, R_out inside the grid, viewing angles, an
off-centre star; and picks the stellar packets, the extra diffuse source (iStar = 0) in a random
cell, or the resonance-line packet loop.  Both sides then transport the same packets (same Philox streams, detmath) and
every float32 tally element, Qphot/absInt/scaInt and the per-packet histories must be equal.
A trial in which both sides hit a `print; stop` condition counts as equal; one in which the
reference itself indexes an array out of bounds (undefined behaviour) is skipped.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mocassin_b200 import workloads as W  # noqa: E402
from mocassin_b200.model import star_indices  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle.f90ref import rt  # noqa: E402
from oracle.f90ref.harness import Reference  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

F32 = np.float32


def random_cdf_rows(rng, nrows, nb):
    p = rng.random((nrows, nb)) ** rng.integers(1, 6)
    p[:, : rng.integers(0, max(nb // 3, 1))] = 0.0           # leading zeros: bins never emitted
    p[rng.random((nrows, nb)) < 0.2] = 0.0                   # flat stretches in the CDF
    p[:, -2] += 1e-3
    c = np.cumsum(p, axis=1)
    c = (c / c[:, -1:]).astype(F32)
    c[c > F32(0.999998)] = F32(1.0)
    return c


def perturb_axis(rng, a, symmetric):
    n = a.shape[0]
    w = rng.uniform(0.3, 1.7, n - 1)
    span = float(a[-1] - a[0])
    b = np.concatenate([[0.0], np.cumsum(w)]) / w.sum() * span + float(a[0])
    return b.astype(F32)


def build(rng):
    kind = rng.choice(["hii", "dust", "multigrid", "cube", "view", "plane"], p=[0.2, 0.2, 0.2, 0.2, 0.12, 0.08])
    nb = int(rng.integers(14, 40))
    sym = bool(rng.integers(0, 2))
    desc = dict(kind=str(kind), nbins=nb)
    if kind == "hii":
        m = W.hii_region(n=int(rng.integers(4, 9)), nbins=nb, debug=bool(rng.integers(0, 2)), seed=int(rng.integers(1, 99)))
    elif kind == "dust":
        m = W.dust_shell(n=int(rng.integers(4, 9)), nbins=nb, tauV=float(rng.choice([0.3, 3.0, 30.0])),
                         isotropic=bool(rng.integers(0, 2)))
    elif kind == "multigrid":
        m = W.multigrid(n=int(rng.integers(6, 10)), nsub=int(rng.integers(3, 7)), nbins=nb, symmetric=sym)
    elif kind == "cube":
        m = W.synthetic_cube(n=int(rng.integers(4, 9)), nbins=nb, clumpy=bool(rng.integers(0, 2)), dust=bool(rng.integers(0, 2)),
                             nPhotons=10 ** 6)
    elif kind == "view":
        m = W.viewing_angles(n=int(rng.integers(5, 9)), nbins=nb, phi_free=bool(rng.integers(0, 2)))
    else:
        m = W.plane_slab(nx=int(rng.integers(3, 7)), ny=int(rng.integers(5, 12)), nz=int(rng.integers(3, 7)), nbins=nb,
                         dust=bool(rng.integers(0, 2)), Hden=float(rng.choice([3.0, 30.0, 300.0])))
    # --- perturbations -------------------------------------------------------------------
    single = m.nGrids == 1 and not m.lgPlaneIonization
    if single and rng.random() < 0.7:
        g = m.grids[0]
        g.xAxis, g.yAxis, g.zAxis = (perturb_axis(rng, a, m.lgSymmetricXYZ) for a in (g.xAxis, g.yAxis, g.zAxis))
        desc["axes"] = "non-uniform"
        if not m.lgSymmetricXYZ and rng.random() < 0.5:
            # off-centre star inside the grid
            pos = [float(rng.uniform(a[1], a[-2])) for a in (g.xAxis, g.yAxis, g.zAxis)]
            idx = star_indices(g, pos)
            if g.active[idx[0] - 1, idx[1] - 1, idx[2] - 1] > 0:
                m.starPosition[0, :] = np.asarray(pos, F32)
                m.starIndeces[0, :3] = idx
                desc["star"] = "off-centre"
    for g in m.grids:
        nR = g.nCells + 1
        scat = np.exp(rng.normal(0.0, rng.choice([0.5, 2.0, 4.0]), (nR, nb))).astype(F32)
        g.opacity = np.asfortranarray((g.opacity * scat).astype(F32))
        if rng.random() < 0.5:
            g.opacity[rng.random((nR, nb)) < 0.05] = 0.0
        g.opacity[0, :] = 0.0
        if g.scaOpac is not None:
            alb = rng.random((nR, nb)).astype(F32) * F32(rng.choice([0.3, 0.9, 1.0]))
            g.scaOpac = np.asfortranarray((g.opacity * alb).astype(F32))
            if g.absOpac is not None:
                g.absOpac = np.asfortranarray((g.opacity - g.scaOpac).astype(F32))
        if g.recPDF is not None and rng.random() < 0.7:
            g.recPDF = np.asfortranarray(random_cdf_rows(rng, nR, nb))
            g.totalLines = rng.random(nR).astype(F32) * F32(rng.choice([0.0, 0.3, 0.9]))
        if g.dustPDF is not None and rng.random() < 0.7:
            g.dustPDF = np.asfortranarray(random_cdf_rows(rng, nR, nb))
            g.dustPDF[:, -1] = 1.0
        if g.Tdust is not None and rng.random() < 0.5:
            # some grain species sublimed (Tdust(nS,0,cell) >= TdustSublime): scattering needs one that is not
            T = rng.uniform(100.0, 2000.0, g.Tdust.shape).astype(F32)
            T[1, :, :] = F32(300.0) if rng.random() < 0.8 else T[1, :, :]
            g.Tdust = np.asfortranarray(T)
            desc["Tdust"] = "random"
    m.inSpectrumProbDen[1, :] = random_cdf_rows(rng, 1, nb)[0] if rng.random() < 0.5 else m.inSpectrumProbDen[1, :]
    if m.gSca is not None and rng.random() < 0.7:
        gs = rng.random(nb).astype(F32) * F32(0.95)
        gs[rng.random(nb) < 0.2] = F32(5.0e-5)             # below the isotropic threshold of hg
        m.gSca = gs
    if single and rng.random() < 0.3:
        g = m.grids[0]
        m.R_out = float(rng.uniform(0.4, 0.9) * max(abs(float(g.xAxis[-1])), abs(float(g.xAxis[0]))))
        desc["R_out"] = m.R_out
    if rng.random() < 0.15 and not m.lgPlaneIonization:
        m.R_out = 0.0
    return m, desc


def pick_mode(rng, m):
    """stellar packets (mostly), the extra diffuse source in a random active cell, or the
    resonance-line packet loop"""
    u = rng.random()
    if u < 0.15 and m.lgGas:
        gp = int(rng.integers(1, m.nGrids + 1))
        g = m.grids[gp - 1]
        cells = np.argwhere(np.asarray(g.active) > 0)
        c = cells[int(rng.integers(0, len(cells)))]
        nb = m.nbins
        m.inSpectrumProbDen[0, :] = random_cdf_rows(rng, 1, nb)[0]
        m.deltaE[0] = F32(1.0e-3)
        return ("diffext", gp, tuple(int(v) + 1 for v in c))
    if u < 0.25 and m.lgDust and m.lgGas:
        for g in m.grids:
            r = rng.integers(0, 3, g.nCells + 1).astype(np.int32)
            r[0] = 0
            g.resLinePackets = r
        return ("reslines",)
    return ("stellar",)


def compare(m, n, seed, lib, mode=("stellar",)):
    o = Oracle(m)
    err_o = err_r = None
    fo = None
    try:
        if mode[0] == "stellar":
            co, fo = o.transport(1, 0, n, seed=seed, want_fates=True)
        elif mode[0] == "diffext":
            co, fo = o.transport(0, 0, n, seed=seed, gpLoc=mode[1], cellLoc=mode[2], want_fates=True)
        else:
            co, n = o.transport_reslines(1, seed=seed)
            n = max(n, 1)
    except RuntimeError as ex:
        err_o = str(ex)
    r = Reference(m, lib)
    try:
        if mode[0] == "stellar":
            cr, fr = r.transport(1, 0, n, seed=seed)
        elif mode[0] == "diffext":
            g = m.grids[mode[1] - 1]
            x, y, z = mode[2]
            r.grid[mode[1]].ldiffuseloc[int(g.active[x - 1, y - 1, z - 1])] = F32(m.deltaE[0])
            cr, fr = r.transport(0, 0, n, seed=seed, gpLoc=mode[1], cellLoc=mode[2])
        else:
            cr, fr = r.transport_reslines(1, seed=seed)
    except rt.FortranStop as ex:
        err_r = str(ex)
    except rt.FortranBoundsError as ex:
        # undefined behaviour in the reference itself (e.g. xAxis(zP) in the plane-parallel
        # emission, photon_mod.f90:612, when nz > nx): nothing to compare against
        return f"reference indexes out of bounds ({ex}): skipped", True
    if err_o or err_r:
        return ("both stop" if (err_o and err_r) else f"STOP MISMATCH oracle={err_o} reference={err_r}"), bool(err_o and err_r)
    bad = []
    if fo is None:
        if co["nSegments"] != cr["nSegments"]:
            bad.append(f"nSegments {co['nSegments']} vs {cr['nSegments']}")
    elif not (np.array_equal(fo[:, 0], fr[:, 0]) and np.array_equal(fo[:, 1], fr[:, 1])):
        k = np.flatnonzero((fo[:, 0] != fr[:, 0]) | (fo[:, 1] != fr[:, 1]))
        bad.append(f"{k.size} packet histories differ, first {k[:3]} oracle {fo[k[:3], :2].tolist()} reference {fr[k[:3], :2].tolist()}")
    for i in range(m.nGrids):
        for key in r.out[i]:
            a, b = o.out[i][key], r.out[i][key]
            if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                bad.append(f"grid {i + 1} {key}: {(a.view(np.uint32) != b.view(np.uint32)).sum()} elements differ")
    for key in ("Qphot", "absInt", "scaInt"):
        if F32(co[key]) != cr[key]:
            bad.append(f"{key}: {co[key]} vs {cr[key]}")
    if m.lgPlaneIonization and not np.array_equal(o.planeIonDistribution, r.G.planeiondistribution.a):
        bad.append("planeIonDistribution differs")
    stats = f"seg/pk {co['nSegments'] / n:.1f} abs {co['nAbs']} sca {co['nSca']} esc {co['nEscaped']} line {co['nLinePackets']} drop {co['nDropped']}"
    return ("; ".join(bad) if bad else "equal  " + stats), not bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=50)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--packets", type=int, default=300)
    a = ap.parse_args()
    lib = orc.load()
    nbad = 0
    t0 = time.time()
    for t in range(a.trials):
        rng = np.random.default_rng([a.seed, t])
        try:
            m, desc = build(rng)
        except Exception as ex:          # a builder rejecting random parameters is not a finding
            print(f"trial {t}: builder failed: {ex!r}")
            continue
        mode = pick_mode(rng, m)
        desc["mode"] = mode[0]
        msg, ok = compare(m, a.packets, 1000 + t, lib, mode)
        nbad += not ok
        print(f"trial {t} {'ok ' if ok else 'BAD'} {desc} :: {msg}", flush=True)
    print(f"{a.trials} trials, {nbad} mismatches, {time.time() - t0:.0f} s")
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
