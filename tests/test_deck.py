"""The input side of a dust-only run (mocassin_b200/deck.py): the reference's shipped
benchmarks/dust decks read from their own files, Mie cross-sections, and the outer Lucy loop.

* BHmie restatement against an independent Mie series (scipy spherical Bessel functions);
* live (where /root/reference is mounted): every shipped dust deck loads, the 1-D benchmark
  decks have the optical depth at 1 um their names promise (tau = 1, 10, 100: Ivezic et al.
  1997 benchmark), and the committed fixtures tests/golden/deck_*.npz are what the loader builds;
* from the fixtures (everywhere): three Lucy iterations of p0tau1 on the CPU oracle conserve
  energy and give a dust temperature that falls monotonically outwards from ~800 K;
* the autoPackets / convergence rules of iterateMC's tail."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mocassin_b200 import deck  # noqa: E402

REF = os.environ.get("MOCASSIN_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
live = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "benchmarks", "dust")), reason="reference tree not mounted")


def _mie_reference(x, m, nmax=None):
    """Qext, Qsca, <cos> from the textbook series with scipy's spherical Bessel functions."""
    from scipy.special import spherical_jn, spherical_yn

    nmax = nmax or int(x + 4.0 * x ** (1.0 / 3.0) + 12)
    n = np.arange(1, nmax + 1)
    mx = m * x

    def psi(z):
        return z * spherical_jn(n, z), spherical_jn(n, z) + z * spherical_jn(n, z, derivative=True)

    px, dpx = psi(x)
    pmx, dpmx = psi(mx)
    xi = px + 1j * (x * spherical_yn(n, x))
    dxi = dpx + 1j * (spherical_yn(n, x) + x * spherical_yn(n, x, derivative=True))
    a = (m * pmx * dpx - px * dpmx) / (m * pmx * dxi - xi * dpmx)
    b = (pmx * dpx - m * px * dpmx) / (pmx * dxi - m * xi * dpmx)
    qext = 2.0 / x ** 2 * np.sum((2 * n + 1) * (a + b).real)
    qsca = 2.0 / x ** 2 * np.sum((2 * n + 1) * (abs(a) ** 2 + abs(b) ** 2))
    g = 4.0 / (x ** 2 * qsca) * (np.sum(n[:-1] * (n[:-1] + 2) / (n[:-1] + 1) * (a[:-1] * np.conj(a[1:]) + b[:-1] * np.conj(b[1:])).real)
                                  + np.sum((2 * n + 1) / (n * (n + 1)) * (a * np.conj(b)).real))
    return qext, qsca, g


@pytest.mark.parametrize("x", [0.01, 0.3, 1.0, 2.5, 7.0, 20.0, 60.0, 100.0])
@pytest.mark.parametrize("m", [1.7 + 0.03j, 1.33 + 1e-4j, 2.6 + 1.4j, 0.9 + 0.4j])
def test_bhmie_agrees_with_independent_mie_series(x, m):
    qe, qs, g = deck.bhmie(x, m)
    re, rs, rg = _mie_reference(float(np.float32(x)), complex(np.complex64(m)))
    assert abs(qe - re) <= 2e-5 * abs(re) + 1e-9, (qe, re)
    assert abs(qs - rs) <= 2e-5 * abs(rs) + 1e-9, (qs, rs)       # sums are kept in float32, as in the reference
    assert abs(g - rg) <= 5e-5 + 1e-4 * abs(rg), (g, rg)


def test_bhmie_small_particle_limits():
    # Rayleigh: Qsca = 8/3 x^4 |(m^2-1)/(m^2+2)|^2, Qabs = 4 x Im((m^2-1)/(m^2+2))
    x, m = 0.02, 1.6 + 0.05j
    qe, qs, g = deck.bhmie(x, m)
    k = (m * m - 1) / (m * m + 2)
    assert abs(qs - 8.0 / 3.0 * x ** 4 * abs(k) ** 2) < 2e-3 * qs
    assert abs((qe - qs) - 4 * x * k.imag) < 2e-3 * (qe - qs)
    assert abs(g) < 1e-3


def test_read_input_rejects_unknown_keywords(tmp_path):
    p = tmp_path / "input.in"
    p.write_text("nx 5\nnotAKeyword 3\n")
    with pytest.raises(ValueError):
        deck.read_input(str(p))
    p.write_text("autoPackets 0.20 2. 10000000\nsymmetricXYZ\nnebComposition noGas\nmaxIterateMC  10 95.\n"
                 "Ndust file 'input/x.ndust'\ndustFile 'a.dat' 'b.dat'\ninclination 2 0.218 -1. 1.35 -1.\nRout 2.18e17\n")
    d = deck.read_input(str(p))
    assert d.lgAutoPackets and d.maxPhotons == 10_000_000 and d.nPhotIncrease == 2.0 and d.convIncPercent == 0.2
    assert d.lgSymmetricXYZ and not d.lgGas and d.lgDust and d.NdustFile == "input/x.ndust"
    assert (d.maxIterateMC, d.minConvergence) == (10, 95.0) and d.dustFile == ("a.dat", "b.dat")
    assert d.nAngleBins == 2 and d.viewPointTheta == [0.0, 0.218, 1.35] and d.viewPointPhi == [0.0, -1.0, -1.0]


def _tau(model, lam_um):
    g = model.grids[0]
    lam = 2.9979250e14 / (model.nuArray.astype(np.float64) * 3.28984e15)
    i = int(np.argmin(abs(lam - lam_um)))
    x = g.xAxis.astype(np.float64)
    W = np.concatenate([[x[0]], (x[1:] + x[:-1]) / 2, [x[-1]]])
    return sum(float(g.opacity[c, i]) * (W[k + 1] - W[k]) for k, c in enumerate(g.active[:, 0, 0]) if c > 0)


@live
@pytest.mark.parametrize("name,tau", [("p0tau1", 1.0), ("p0tau10", 10.0), ("p0tau100", 100.0)])
def test_shipped_1d_decks_have_their_nominal_optical_depth_and_match_the_fixtures(name, tau):
    m, t, d = deck.load_dust_deck(os.path.join(REF, "benchmarks", "dust", "1D", name), REF)
    assert m.nbins == 215                        # nuDustRyd.dat up to nuMax = 15 Ryd (grid_mod.f90:180-211)
    assert abs(_tau(m, 1.0) - tau) < 0.015 * tau  # radial optical depth at 1 um: definition of the benchmark
    assert m.lgSymmetricXYZ and not m.lgGas and d.lgAutoPackets
    assert m.starIndeces.tolist() == [[1, 1, 1, 1]] and m.grids[0].active[0, 0, 0] == 0     # cavity around the star
    want = dict(np.load(os.path.join(GOLD, f"deck_{name}.npz")))
    got = deck.deck_to_arrays(m, t, d)
    assert sorted(want) == sorted(got)
    for k in want:
        assert np.array_equal(np.asarray(got[k]), want[k]), k


@live
def test_disk_fixture_is_what_the_loader_builds():
    m, t, d = deck.load_dust_deck(os.path.join(REF, "benchmarks", "dust", "2D", "tau1.000"), REF)
    want = dict(np.load(os.path.join(GOLD, "deck_tau1.000.npz")))
    got = deck.deck_to_arrays(m, t, d)
    for k in want:
        assert np.array_equal(np.asarray(got[k]), want[k]), k


@live
@pytest.mark.parametrize("name", ["tau0.100", "tau1.000", "tau10.00", "tau100.0"])
def test_shipped_2d_disk_decks_load(name):
    m, t, d = deck.load_dust_deck(os.path.join(REF, "benchmarks", "dust", "2D", name), REF)
    g = m.grids[0]
    assert (g.nx, g.ny, g.nz) == (40, 40, 40) and m.nbins == 215 and m.nAngleBins == 2
    at = m.angle_tables()
    assert at["totAngleBinsPhi"] == 1            # phi = -1: phi-free viewing angles (grid_mod.f90:416-429)
    tauV = float(name[3:])
    assert 0.8 * tauV < _tau(m, 0.55) < 1.05 * tauV      # midplane optical depth at 550 nm (Pascucci et al. 2004)


@live
def test_constant_density_deck_on_automatic_axes(tmp_path):
    """`Ndust constant` + `edges`: fillGrid's automatic axes, every cell inside [Rin, Rout] active,
    radial optical depth = Ndust * Cext * (Rout - Rin) to the half-cell the shell edges are resolved to."""
    import shutil

    src = os.path.join(REF, "benchmarks", "dust", "1D", "p0tau1")
    for f in ("p0tau1_grainsizes.dat", "p0tau1_grainspecies.dat"):
        shutil.copy(os.path.join(src, f), tmp_path / f)
    (tmp_path / "input.in").write_text(
        "symmetricXYZ\ncontShape blackbody\nnebComposition noGas\nmaxIterateMC 3 95.\nnPhotons 1000\nnx 21\nny 21\nnz 21\n"
        "LStar 1.\nTStellar 3000.\nRin 2.e16\nRout 1.e17\nedges 1.e17 1.e17 1.e17\nNdust constant 2.e-9\n"
        "dustFile 'input/p0tau1_grainspecies.dat' 'input/p0tau1_grainsizes.dat'\n")
    m, t, d = deck.load_dust_deck(str(tmp_path), REF)
    g = m.grids[0]
    assert (g.nx, g.ny, g.nz) == (21, 21, 21) and np.array_equal(g.xAxis, g.zAxis) and g.xAxis[-1] == np.float32(1.0e17)
    r = np.sqrt(sum(np.meshgrid(*(3 * [g.xAxis.astype(np.float64) ** 2]), indexing="ij")))
    diff = (g.active > 0) != ((r >= 2.0e16) & (r <= 1.0e17))         # float32 radius test vs float64: boundary cells only
    assert np.all((abs(r[diff] / 2.0e16 - 1) < 1e-6) | (abs(r[diff] / 1.0e17 - 1) < 1e-6)) and diff.sum() < 30
    assert np.all(g.Ndust[1:] == np.float32(2.0e-9))
    lam = 2.9979250e14 / (m.nuArray.astype(np.float64) * 3.28984e15)
    i = int(np.argmin(abs(lam - 1.0)))
    cext = float(g.opacity[1, i]) / 2.0e-9
    assert abs(_tau(m, 1.0) - 2.0e-9 * cext * 8.0e16) < 2.0e-9 * cext * 0.5e16
    (tmp_path / "input.in").write_text((tmp_path / "input.in").read_text().replace("edges 1.e17 1.e17 1.e17\n", ""))
    with pytest.raises(ValueError):
        deck.load_dust_deck(str(tmp_path), REF)


def test_three_lucy_iterations_of_p0tau1_on_the_oracle(oracle_lib):
    from deck_runner import oracle_step

    m, t, d = deck.deck_from_arrays(dict(np.load(os.path.join(GOLD, "deck_p0tau1.npz"))))
    d.maxIterateMC = 3
    step, st = oracle_step(m, t, d, seed=77, threads=2)
    hist = deck.iterate_dust(d, m, step)
    assert [h["nPhotons"] for h in hist] == [100000, 100000, 100000]
    assert hist[1]["converged_pct"] > hist[0]["converged_pct"]
    for c in st["counters"]:                     # dust only, no line packets: every packet escapes
        assert c["nEscaped"] == 100000 and c["nDropped"] == 0 and c["trapped"] == 0
    esc = st["escaped"]
    assert abs(float(esc.astype(np.float64).sum()) - 100000 * float(m.deltaE[1])) < 1e-3 * 38.26
    g = m.grids[0]
    cells = [c for c in g.active[:, 0, 0] if c > 0]
    T = g.Tdust[0, 0, cells]
    assert 650.0 < T[0] < 900.0                  # benchmark: 800 K at the inner edge of the shell
    assert np.all(np.diff(T[:9]) < 0) and np.all(T[9:] < 80.0)      # outer axis cells: few packets at 1e5, noisy


def test_disk_deck_iteration_on_the_oracle_conserves_energy_and_fills_the_viewing_angles(oracle_lib):
    """benchmarks/dust/2D/tau1.000 (40^3 disk, two phi-free inclinations) from its fixture: one
    iteration on the oracle; writeSED's total equals LStar, both viewing-angle columns are filled."""
    from deck_runner import oracle_step
    from mocassin_b200 import output

    m, t, d = deck.deck_from_arrays(dict(np.load(os.path.join(GOLD, "deck_tau1.000.npz"))))
    assert m.nAngleBins == 2 and m.grids[0].nx == 40
    d.maxIterateMC, d.nPhotons = 1, 60000
    m.deltaE[1] = np.float32(d.LStar) / np.float32(d.nPhotons)
    step, st = oracle_step(m, t, d, seed=9, threads=2)
    deck.iterate_dust(d, m, step)
    esc = st["escaped"]
    raw = esc[:, 1:, :].astype(np.float64).sum(axis=0).astype(np.float32)
    sed, totalE = output.sed_from_raw(m, t["widFlx"], raw)
    assert abs(totalE - d.LStar) < 2e-4 * d.LStar
    assert raw[:, 1].sum() > 0 and raw[:, 2].sum() > raw[:, 1].sum()      # 12.5 deg bin is smaller than the 77 deg one
    assert np.all(sed >= 0)
    T = m.grids[0].Tdust[0, 0, 1:]
    assert T.max() > 200.0 and T.min() >= 1.0


def test_iterate_dust_follows_the_autopackets_rule():
    """iteration_mod.f90:1106-1124: packets are multiplied by nPhotIncrease after an iteration
    (not the first) whose converged fraction grew by <= convIncPercent, while the (double
    counted) total stays below maxPhotons; the loop ends at minConvergence or maxIterateMC."""
    d = deck.Deck(lgAutoPackets=True, convIncPercent=0.2, nPhotIncrease=2.0, maxPhotons=1000, nPhotons=100,
                  maxIterateMC=10, minConvergence=95.0)

    class M:
        deltaE = np.array([0.0, 1.0], dtype=np.float32)

    seq = iter([10, 50, 55, 56, 80, 81, 96])
    calls = []

    def step(n, dE):
        calls.append((n, dE))
        return next(seq), 100

    hist = deck.iterate_dust(d, M, step)
    assert [c[0] for c in calls] == [100, 100, 100, 200, 400, 400, 800]
    assert [c[1] for c in calls] == [1.0, 1.0, 1.0, 0.5, 0.25, 0.25, 0.125]
    assert hist[-1]["converged_pct"] == 96.0 and len(hist) == 7
    # the cap: 2*nPhotons >= maxPhotons stops the growth
    seq = iter([10, 11, 12, 13, 14])
    calls.clear()
    d2 = deck.Deck(lgAutoPackets=True, convIncPercent=0.2, nPhotIncrease=2.0, maxPhotons=500, nPhotons=100,
                   maxIterateMC=5, minConvergence=95.0)
    deck.iterate_dust(d2, M, step)
    assert [c[0] for c in calls] == [100, 100, 200, 400, 400]


def test_path_length_quantum_leaves_head_room_on_graded_axes():
    """The fixed-point unit of the J tally is the smallest cell half-width / 2^24 -- unless the widest cell is
    more than 2^9 times wider, when it is the widest half-width / 2^33: the 2D disk deck (spacing ratio 10^5)
    would otherwise come within 3 bits of wrapping a 64-bit sum at its own 8x10^6 packets (measured on the
    oracle: max element 2^56.8 at 10^6 packets with the fine unit, 2^49.8 with the capped one)."""
    for name, want in (("p0tau1", 21), ("p0tau10", 22), ("p0tau100", 22), ("tau1.000", 17)):
        m, t, d = deck.deck_from_arrays(dict(np.load(os.path.join(GOLD, f"deck_{name}.npz"))))
        g = m.grids[0]
        e = m.len_unit_exponent(g)
        assert e == want, name
        fine = int(np.floor(np.log2(g.min_cell_width()))) - 24
        assert e >= fine and (e == fine or name == "tau1.000")
        # one diagonal crossing of the widest cell is below 2^37 units: > 6x10^7 of them fit a 63-bit sum
        assert 2.0 * np.sqrt(3.0) * 2.0 * g.max_cell_width() / 2.0 ** e < 2.0 ** 37
