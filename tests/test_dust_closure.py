"""CPU tests of the oracle restatement of the dust-only closure: getFlux
(continuum_mod.f90:359-416), setDustPDF (emission_mod.f90:1313-1387) and getDustT + the
dust-only updateCell branch (update_mod.f90:308-334, :1836-1945).  The reference ships no
fixtures for these routines, so the oracle is checked against float64 numpy restatements of
the same formulae and against the self-consistency the physics gives (a cell bathed in
B_nu(T0) gets T0 back)."""
import numpy as np
import pytest

from mocassin_b200 import workloads as W
from mocassin_b200.api import scale_estimators
from oracle import oracle as O

F32 = np.float32
HC = 157893.94
H = 6.6262e-27


def planck64(nu, T):
    return (0.5250229 / H) * nu ** 3 / np.expm1(HC * nu / T)


@pytest.mark.parametrize("nu,T", [(1.0, 1.0e4), (0.3, 2500.0), (1.0e-4, 30.0), (5.0, 1.0e5), (0.01, 300.0), (3.0, 6000.0)])
def test_get_flux_planck_branch(nu, T):
    assert HC * nu / T <= 86.0
    ref = planck64(np.float64(F32(nu)), np.float64(F32(T)))
    assert abs(O.get_flux(nu, T) - ref) <= 4e-6 * ref


@pytest.mark.parametrize("nu,T", [(10.0, 100.0), (1.0, 1000.0), (0.1, 150.0), (15.0, 3.0)])
def test_get_flux_wien_branch(nu, T):
    x = HC * np.float64(F32(nu)) / np.float64(F32(T))
    assert x > 86.0
    ref = (0.5250229 / H) * np.float64(F32(nu)) ** 3 * np.exp(-x)
    got = O.get_flux(nu, T)
    if ref < 1e-45:
        assert got == 0.0 or got < 2e-45
    else:
        assert abs(got - ref) <= 1e-5 * ref + 1.5e-45


def test_get_flux_rayleigh_jeans_branch():
    # exp(x) - 1 rounds to 0 in float32 when x < 2^-24
    nu, T = 1.0e-9, 1.0e8
    assert F32(np.exp(F32(HC * nu / T))) - F32(1.0) <= 0
    ref = 3.32154e-6 * nu * nu * T / H
    assert abs(O.get_flux(nu, T) - ref) <= 1e-5 * ref


@pytest.fixture(scope="module")
def case():
    return W.dust_closure(n=7, nbins=90, nPhotons=20000)


def _pdf64(model, g, t, cell):
    nb = model.nbins
    comp = int(g.dustAbunIndex[cell]) if model.lgMultiDustChemistry else 1
    dcp = int(model.dustComPoint[comp - 1])
    nu = model.nuArray.astype(np.float64)
    row = np.zeros(nb)
    for n in range(1, int(model.nSpeciesPart[comp - 1]) + 1):
        for ai in range(1, model.nSizes + 1):
            T = float(g.Tdust[n, ai, cell])
            if T > 0 and T < model.TdustSublime[dcp - 1 + n - 1]:
                o = int(t["dustAbsXsecP"][n + dcp - 2, ai - 1]) - 1
                x = HC * nu / T
                bb = (0.5250229 / H) * nu ** 3 * np.where(x > 86, np.exp(-x), 1.0 / np.expm1(np.minimum(x, 86.0)))
                row += t["xSecArray"][o:o + nb].astype(np.float64) * bb * t["widFlx"] * t["grainWeight"][ai - 1] * model.grainAbun[comp - 1, n - 1]
    c = np.cumsum(row)
    return c / c[-1]


def test_dust_pdf_rows(case):
    model, t = case
    g = model.grids[0]
    rng = np.random.default_rng(3)
    g.Tdust[1:, 1:, 1:] = rng.uniform(20.0, 1300.0, size=g.Tdust[1:, 1:, 1:].shape).astype(F32)
    g.Tdust[1, 1, 1:] = F32(300.0)            # at least one grain below every sublimation limit
    pdf = O.dust_pdf(model, g, t)
    assert pdf.shape == (g.nCells + 1, model.nbins)
    assert np.all(pdf[0] == 0)
    assert np.all(pdf[1:, -1] == 1.0)
    assert np.all(np.diff(np.minimum(pdf[1:], 1.0), axis=1) >= 0)
    for cell in (1, 2, g.nCells // 2, g.nCells):
        ref = _pdf64(model, g, t, cell)
        assert np.max(np.abs(pdf[cell] - ref)) < 3e-5


def test_dust_pdf_sublimed_grains_are_excluded(case):
    model, t = case
    g = model.grids[0]
    g.Tdust[:, :, 1:] = F32(100.0)
    g.Tdust[2, :, 1:] = F32(1300.0)           # species 2 of component 1 is above its 1200 K limit
    pdf = O.dust_pdf(model, g, t)
    cell = int(np.nonzero(g.dustAbunIndex[1:] == 1)[0][0]) + 1
    assert np.max(np.abs(pdf[cell] - _pdf64(model, g, t, cell))) < 3e-5
    g.Tdust[:, :, 1:] = F32(5000.0)           # everything sublimed: 0/0 rows, last entry forced to 1
    pdf = O.dust_pdf(model, g, t)
    assert np.all(np.isnan(pdf[1:, :-1])) and np.all(pdf[1:, -1] == 1.0)


def _bath(model, g, t, T0):
    """Host-scaled Jste of a cell bathed in B_nu(T0): Jste/pi = 4 B_nu dnu."""
    nu = model.nuArray.astype(np.float64)
    J = np.pi * 4.0 * planck64(nu, T0) * H * 3.28984e15 * t["widFlx"].astype(np.float64)
    out = np.zeros((g.nCells + 1, model.nbins), dtype=F32, order="F")
    out[1:, :] = J.astype(F32)[None, :]
    return out


@pytest.mark.parametrize("T0", [35.0, 250.5, 1100.0])
def test_dust_update_recovers_bath_temperature(case, T0):
    model, t = case
    g = model.grids[0]
    g.Tdust[:, :, 1:] = F32(100.0)
    T, conv = O.dust_update(model, g, t, _bath(model, g, t, T0), 0.05)
    c1 = g.dustAbunIndex[1:] == 1
    # component 1: species 1,2 use their own tables -> every grain sits at T0
    assert np.allclose(T[1:3, 1:, 1:][:, :, c1], T0, rtol=2e-3)
    assert np.allclose(T[0, 0, 1:][c1], T0, rtol=2e-3)
    # component 2 (global species 3): getDustT uses the local index, i.e. species 1's tables (sic)
    c2 = ~c1
    assert np.allclose(T[1, 1:, 1:][:, c2], T0, rtol=2e-3)
    assert np.all(T[2, :, 1:][:, c2] == 0)
    # weighted means: Tdust(nS,0) = sum_ai w T, Tdust(0,0) = sum_nS abun Tdust(nS,0)
    w = t["grainWeight"]
    assert np.allclose(T[1, 0, 1:], (T[1, 1:, 1:] * w[:, None]).sum(axis=0), rtol=1e-5)
    expect = abs(T[0, 0, 1:] - 100.0) / 100.0 <= 0.05
    assert np.array_equal(conv[1:].astype(bool), expect)
    assert conv[0] == 0


def test_dust_update_table_ends(case):
    model, t = case
    g = model.grids[0]
    g.Tdust[:, :, 1:] = F32(100.0)
    faint = np.full((g.nCells + 1, model.nbins), 1.0e-30, dtype=F32, order="F")
    T, conv = O.dust_update(model, g, t, faint, 0.05)
    assert np.all(T[1, 1:, 1:] == 1.0)                       # below the table: 1 K (lgTalk branch)
    # a cell no packet crossed is not updated at all (updateCell's lgHit test,
    # update_mod.f90:104-149): Tdust and lgConverged keep the values they came in with
    faint[3, :] = 0.0
    was = np.zeros(g.nCells + 1, np.int32)
    was[3] = 1
    T, conv = O.dust_update(model, g, t, faint, 0.05, lgConverged=was)
    assert np.all(T[:, :, 3] == 100.0) and conv[3] == 1
    assert np.all(T[1, 1:, 4] == 1.0) and conv[4] == 0
    hot = _bath(model, g, t, 2900.0) * F32(10.0)
    T, conv = O.dust_update(model, g, t, hot, 0.05)
    assert np.all(T[1, 1:, 1:] == 3000.0)                    # above the table: nTemps
    assert not conv[1:].any()


def test_dust_only_lucy_iteration_converges(case):
    """transport -> host scaling -> getDustT -> setDustPDF, twice round: temperatures fall
    with radius and the second pass moves them less than the first."""
    model, t = case
    g = model.grids[0]
    g.Tdust[:, :, 1:] = F32(100.0)
    n = 20000
    moves = []
    for it in range(3):
        g.dustPDF = O.dust_pdf(model, g, t)
        orc = O.Oracle(model)
        orc.transport(1, 0, n, seed=100 + it)
        J = orc.folded(1, float(model.deltaE[1]))["Jste"]
        Js, _ = scale_estimators(model, J, np.zeros((1, 1, 1), F32))
        Told = g.Tdust[0, 0, 1:].copy()
        T, conv = O.dust_update(model, g, t, Js, 0.05)
        g.Tdust = T
        moves.append(float(np.median(np.abs(T[0, 0, 1:] - Told) / Told)))
    assert moves[2] < moves[0]
    x = g.xAxis
    r = np.sqrt(sum(np.meshgrid(x.astype(np.float64) ** 2, x.astype(np.float64) ** 2, x.astype(np.float64) ** 2, indexing="ij")))
    rc = np.zeros(g.nCells + 1)
    m = g.active > 0
    rc[g.active[m]] = r[m]
    inner = T[0, 0, 1:][rc[1:] < 0.4 * model.R_out]
    outer = T[0, 0, 1:][rc[1:] > 0.8 * model.R_out]
    assert np.median(inner) > np.median(outer) > 5.0
