"""Round trips of the reference's checkpoint text formats (writeGrid / resetGrid,
grid_mod.f90:2646-2870, :2967-3567) for the arrays this repository owns."""
import os

import numpy as np

from mocassin_b200 import checkpoint as ck
from mocassin_b200 import workloads as W

F32 = np.float32


def test_grid0_round_trip_multigrid(tmp_path):
    m = W.multigrid(n=9, nsub=5, nbins=40, nPhotons=10)
    rng = np.random.default_rng(1)
    conv = [rng.integers(0, 2, g.nCells + 1).astype(np.int32) for g in m.grids]
    for c in conv:
        c[0] = 0
    p = os.path.join(tmp_path, "grid0.out")
    ck.write_grid0(p, m, lgConverged=conv)
    grids, R_out, conv2, black2 = ck.read_grid0(p)
    assert len(grids) == m.nGrids and R_out == F32(m.R_out)
    for g, h, c, c2 in zip(m.grids, grids, conv, conv2):
        assert (h.nx, h.ny, h.nz, h.nCells, h.motherP) == (g.nx, g.ny, g.nz, g.nCells, g.motherP)
        for a, b in ((g.xAxis, h.xAxis), (g.yAxis, h.yAxis), (g.zAxis, h.zAxis)):
            assert np.array_equal(a.astype(F32), b)          # 9 significant digits: float32 exact
        assert np.array_equal(g.active, h.active)
        assert np.array_equal(c, c2)
    # layout: header lines, axes, then nx*ny*nz records per grid, z fastest
    lines = open(p).read().splitlines()
    g = m.grids[0]
    first = 2 + g.nx + g.ny + g.nz
    assert lines[0].split() == [str(m.nGrids)]
    assert [int(t) for t in lines[1].split()[:5]] == [g.nx, g.ny, g.nz, g.nCells, g.motherP]
    assert int(lines[first].split()[0]) == int(g.active[0, 0, 0]) and int(lines[first + 1].split()[0]) == int(g.active[0, 0, 1])


def test_dust_grid_round_trip(tmp_path):
    m, t = W.dust_closure(n=7, nbins=40)
    g = m.grids[0]
    rng = np.random.default_rng(2)
    g.Tdust[:, :, 1:] = rng.uniform(10, 900, size=g.Tdust[:, :, 1:].shape).astype(F32)
    want = (g.Ndust.copy(), g.dustAbunIndex.copy(), g.Tdust.copy())
    for multi in (True, False):
        p = os.path.join(tmp_path, f"dustGrid{int(multi)}.out")
        ck.write_dust_grid(p, m, lgMultiChemistry=multi, totalDustMass=1.5)
        txt = open(p).read().splitlines()
        assert "Total dust mass [Msol]" in txt[-1]
        per = 1 + (m.nSizes + 1)
        assert len(txt) == g.nx * g.ny * g.nz * per + 3
        assert len(txt[1].split()) == m.nSpeciesMax + 1          # one Tdust line per size: species 0..nSpeciesMax
        g.Ndust = None; g.Tdust = None
        if multi:
            g.dustAbunIndex = None
        ck.read_dust_grid(p, m, lgMultiChemistry=multi)
        assert np.array_equal(g.Ndust[1:], want[0][1:])
        assert np.array_equal(g.Tdust[:, :, 1:], want[2][:, :, 1:])
        assert np.array_equal(g.dustAbunIndex[1:], want[1][1:])


def test_photo_source_round_trip(tmp_path):
    m, t = W.dust_closure(n=7, nbins=40)
    p = os.path.join(tmp_path, "photoSource.out")
    ck.write_photo_source(p, m, ["blackbody"], [2500.0], [38.26], [100000])
    s = ck.read_photo_source(p)
    assert len(s) == 1 and s[0]["contShape"] == "blackbody" and s[0]["nPhotons"] == 100000
    assert s[0]["position"] == (0.0, 0.0, 0.0) and abs(s[0]["LStar"] - 38.26) < 1e-6


def test_checkpoint_feeds_a_second_engine_state(tmp_path):
    """A state written by one model object restores an identically configured second one."""
    a, t = W.dust_closure(n=7, nbins=40, T0=77.0)
    ck.write_checkpoint(str(tmp_path), a, lgMultiChemistry=True)
    b, _ = W.dust_closure(n=7, nbins=40, T0=5.0)
    grids, R_out, conv, _ = ck.read_grid0(os.path.join(tmp_path, "grid0.out"))
    assert np.array_equal(grids[0].active, b.grids[0].active)
    ck.read_dust_grid(os.path.join(tmp_path, "dustGrid.out"), b, lgMultiChemistry=True)
    assert np.array_equal(b.grids[0].Tdust[:, :, 1:], a.grids[0].Tdust[:, :, 1:])


def test_grid1_grid2_grid3_round_trip(tmp_path):
    """grid1.out (Te, Ne, Hden, abFileIndex), grid2.out (ionDen) and grid3.out (run parameters):
    what write_* writes, read_* reads back (the order resetGrid reads in, grid_mod.f90:3038-3111,
    :3423-3465)."""
    import ref_cases

    m, rp, s = ref_cases.writegrid_inputs()
    ref_cases.run_writers(str(tmp_path))
    Te, Ne, Hden, ab = ck.read_grid1(os.path.join(tmp_path, "grid1.out"), m.grids, multi_chemistry=True)
    for iG, g in enumerate(m.grids):
        assert np.array_equal(Te[iG][1:], s["Te"][iG][1:]) and np.array_equal(Ne[iG][1:], s["Ne"][iG][1:])
        assert np.array_equal(Hden[iG][1:], g.Hden[1:]) and np.array_equal(ab[iG], s["abFileIndex"][iG])
    ion = ck.read_grid2(os.path.join(tmp_path, "grid2.out"), m.grids, s["lgElementOn"], s["elementXref"], rp.nstages)
    for iG in range(m.nGrids):
        want = s["ionDen"][iG].copy()
        for e in (1, 2, 6, 8):                       # stages beyond min(elem+1, nstages) are not stored
            want[:, int(s["elementXref"][e - 1]) - 1, min(e + 1, rp.nstages):] = 0
        assert np.array_equal(ion[iG][1:], want[1:])
    d = ck.read_grid3(os.path.join(tmp_path, "grid3.out"))
    assert d["nGrids"] == 2 and d["nbins"] == m.nbins and d["lgSymmetricXYZ"] == m.lgSymmetricXYZ
    assert d["abundanceFile"] == ["abun/solar.dat"] and d["dustSpeciesFile"] == ["dust/sil.dat"] and d["dustFile2"] == "sizes.dat"
    assert d["lgAutoPackets"] and d["maxPhotons"] == 10 ** 7 and d["nstages"] == 5 and d["lgDust"] and d["lgGas"]
    assert d["nAngleBins"] == 2 and np.allclose(d["viewPointTheta"], [0, 0.5, 1.9]) and d["nSpeciesPart"] == [1]
    assert (d["nSpeciesMax"], d["nSizes"]) == (m.nSpeciesMax, m.nSizes) and d["lgNosource"] is False


def test_2d_checkpoints_hold_one_plane_of_the_mother_grid_and_read_back(tmp_path):
    """lg2D: grid0/1/2.out and dustGrid.out hold plane j = 1 of the mother grid only (grid_mod.f90:2709-2713);
    the readers complete the other planes of `active` as resetGrid does (:3476-3500) and hand the per-cell
    arrays back unchanged."""
    m, t = W.dust_closure(n=7, nbins=40)
    g = m.grids[0]
    scale = ck.fill_2d_planes(g)                     # make the grid a 2D one: planes j >= 2 alias plane 1
    ids = np.unique(g.active[g.active > 0])
    assert set(ids) <= set(np.unique(g.active[:, 0, :])) and scale[ids].sum() == (g.active > 0).sum()
    rng = np.random.default_rng(3)
    g.Tdust[:, :, 1:] = rng.uniform(10, 900, size=g.Tdust[:, :, 1:].shape).astype(F32)
    g.Hden = rng.uniform(1, 100, g.nCells + 1).astype(F32)
    Te = [rng.uniform(5e3, 2e4, g.nCells + 1).astype(F32)]
    Ne = [rng.uniform(1, 1e3, g.nCells + 1).astype(F32)]
    conv = [rng.integers(0, 2, g.nCells + 1).astype(np.int32)]
    p = lambda f: os.path.join(tmp_path, f)
    ck.write_grid0(p("grid0.out"), m, lgConverged=conv, lg2D=True)
    ck.write_grid1(p("grid1.out"), m, Te, Ne, lg2D=True)
    ck.write_dust_grid(p("dustGrid.out"), m, lg2D=True)
    per_cell = g.nx * g.nz
    assert len(open(p("grid0.out")).read().splitlines()) == 2 + g.nx + g.ny + g.nz + per_cell
    assert len(open(p("grid1.out")).read().splitlines()) == per_cell
    assert len(open(p("dustGrid.out")).read().splitlines()) == per_cell * (1 + m.nSizes + 1) + 3
    grids, _, conv2, _ = ck.read_grid0(p("grid0.out"), lg2D=True)
    assert np.array_equal(grids[0].active, g.active)
    seen = np.unique(g.active[:, 0, :][g.active[:, 0, :] > 0])
    assert np.array_equal(conv2[0][seen], conv[0][seen])
    Te2, Ne2, Hd2, _ = ck.read_grid1(p("grid1.out"), grids, lg2D=True)
    assert np.array_equal(Te2[0][seen], Te[0][seen]) and np.array_equal(Ne2[0][seen], Ne[0][seen]) and np.array_equal(Hd2[0][seen], g.Hden[seen])
    want = (g.Ndust.copy(), g.Tdust.copy())
    g.Ndust = None; g.Tdust = None
    ck.read_dust_grid(p("dustGrid.out"), m, lg2D=True)
    assert np.array_equal(g.Ndust[seen], want[0][seen]) and np.array_equal(g.Tdust[:, :, seen], want[1][:, :, seen])
