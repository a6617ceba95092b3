"""CPU tests that pin the oracle's building blocks (the reference ships no golden
vectors for this path -- SURVEY.md section 4 -- so these are the pins we create)."""
import ctypes as C

import numpy as np
import pytest

from mocassin_b200 import model as M
from oracle import oracle as orc

fp = orc.fp


def _fptr(a):
    return a.ctypes.data_as(fp)


def test_philox_known_answers(oracle_lib):
    # Random123 kat_vectors, philox4x32-10
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    out = (C.c_uint32 * 4)()
    for ctr, key, want in kats:
        oracle_lib.oracle_philox(*ctr, *key, out)
        assert tuple(out) == want


def test_uniforms_are_24bit_in_unit_interval(oracle_lib):
    u = np.zeros(4096, np.float32)
    oracle_lib.oracle_uniforms(12345, 7, 1, u.shape[0], _fptr(u))
    assert u.min() >= 0.0 and u.max() < 1.0
    assert np.all(u * 2.0 ** 24 == np.floor(u * 2.0 ** 24))
    assert abs(u.mean() - 0.5) < 0.02
    # counter based: stream of packet 7 does not depend on other packets
    v = np.zeros(16, np.float32)
    oracle_lib.oracle_uniforms(12345, 7, 1, 16, _fptr(v))
    assert np.array_equal(u[:16], v)
    w = np.zeros(16, np.float32)
    oracle_lib.oracle_uniforms(12345, 8, 1, 16, _fptr(w))
    assert not np.array_equal(v, w)


@pytest.mark.parametrize("which,fn,lo,hi", [
    (0, np.log, 2.0 ** -24, 1.0), (1, np.sin, -7.0, 7.0), (2, np.cos, -7.0, 7.0),
    (3, np.arccos, -1.0, 1.0), (4, np.arctan, -1.0e6, 1.0e6), (5, np.exp, -87.0, 88.0)])
def test_detmath_within_one_ulp_of_libm(oracle_lib, which, fn, lo, hi):
    rng = np.random.default_rng(which)
    x = rng.uniform(lo, hi, 200000).astype(np.float32)
    if which == 4:
        x = np.concatenate([x, rng.uniform(-3, 3, 100000).astype(np.float32), np.float32([0, 1, -1, 0.41421357, 1e-30])])
    if which == 0:
        x = np.concatenate([x, np.float32([1.0, 2.0 ** -24, 0.5, 0.70710677, 0.70710683])])
    if which == 3:
        x = np.concatenate([x, np.float32([1, -1, 0, 0.99999994, -0.99999994])])
    y = np.zeros_like(x)
    oracle_lib.oracle_detmath(which, _fptr(x), _fptr(y), x.shape[0])
    ref = fn(x.astype(np.float64))
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    err = np.abs(y.astype(np.float64) - ref) / np.maximum(ulp, 1e-45)
    assert err.max() <= 1.0, (err.max(), x[err.argmax()])


def test_locate_edge_semantics(oracle_lib):
    xa = np.float32([0.0, 1.0, 2.0, 4.0, 8.0])
    n = xa.shape[0]
    loc = lambda x: oracle_lib.oracle_locate(_fptr(xa), n, C.c_float(x))
    assert loc(-0.1) == 0            # below range
    assert loc(8.1) == n             # above range
    assert loc(0.0) == 1
    assert loc(0.5) == 1
    assert loc(1.0) == 2             # x == xa(2): first xa > x is xa(3)
    assert loc(3.9) == 3
    assert loc(7.9) == 4
    assert loc(8.0) == 1             # quirk: x == xa(n) -> minloc over empty mask -> max(-1,1)
    for x in np.random.default_rng(0).uniform(-1, 9, 500):
        assert loc(x) == M.locate(xa, x)


def test_getnu2_semantics(oracle_lib):
    nb = 50
    pdf = np.linspace(0.02, 1.0, nb).astype(np.float32)
    got = np.array([oracle_lib.oracle_getnu2(_fptr(pdf), 1, nb, 99, pid, 1) for pid in range(20000)])
    assert got.min() >= 2 and got.max() <= nb - 1      # bins 1 and nbins are never returned
    u = np.zeros(1, np.float32)
    for pid in range(200):
        oracle_lib.oracle_uniforms(99, pid, 1, 1, _fptr(u))
        k = max(int(np.searchsorted(pdf, u[0], side="right")), 1)     # leading entries with u >= cdf
        want = k + 1 if k < nb - 1 else k
        assert got[pid] == want
    # strided rows (recPDF(cell,:) has stride nCells+1)
    tab = np.zeros((7, nb), np.float32, order="F")
    tab[3, :] = pdf
    s = oracle_lib.oracle_getnu2(tab[3:, :].ctypes.data_as(fp), 7, nb, 99, 5, 1)
    assert s == got[5]


def test_random_unit_vector(oracle_lib):
    v = np.zeros((5000, 3), np.float32)
    for i in range(v.shape[0]):
        oracle_lib.oracle_random_unit_vector(1, i, 1, _fptr(v[i]))
    nrm = np.linalg.norm(v.astype(np.float64), axis=1)
    assert np.abs(nrm - 1).max() < 1e-6
    assert np.abs(v.mean(axis=0)).max() < 0.05
    # w = 2 r1 - 1 exactly
    u = np.zeros(2, np.float32)
    oracle_lib.oracle_uniforms(1, 17, 1, 2, _fptr(u))
    assert v[17, 2] == np.float32(2.0) * u[0] - np.float32(1.0)


@pytest.mark.parametrize("g", [0.0, 0.3, 0.85])
def test_hg_mean_cosine_is_g(oracle_lib, g):
    rng = np.random.default_rng(3)
    n = 40000
    mu = np.zeros(n)
    vout = np.zeros(3, np.float32)
    for i in range(n):
        vin = rng.standard_normal(3)
        vin = (vin / np.linalg.norm(vin)).astype(np.float32)
        ierr = oracle_lib.oracle_hg(C.c_float(g), _fptr(vin), 5, i, 1, _fptr(vout))
        assert ierr == 0
        mu[i] = float(np.dot(vin.astype(np.float64), vout.astype(np.float64)))
    assert abs(mu.mean() - g) < 4.0 / np.sqrt(n)
    # axis-aligned branch (denom <= 0.001)
    vin = np.float32([0, 0, -1])
    oracle_lib.oracle_hg(C.c_float(0.5), _fptr(vin), 5, 1, 1, _fptr(vout))
    assert abs(np.linalg.norm(vout) - 1) < 1e-6


def test_escape_bins(oracle_lib):
    from mocassin_b200 import workloads as W

    m = W.viewing_angles()
    o = orc.Oracle(m)
    T, P = C.c_int32(), C.c_int32()
    rng = np.random.default_rng(1)
    for _ in range(2000):
        d = rng.standard_normal(3)
        d = (d / np.linalg.norm(d)).astype(np.float32)
        assert oracle_lib.oracle_escape_bins(C.byref(o.P), _fptr(d), C.byref(T), C.byref(P)) == 0
        theta = np.arccos(np.float64(d[2]))
        assert T.value == min(int(np.float32(theta) / o.at["dTheta"]) + 1, 10) or abs(theta / np.pi * 10 % 1) < 1e-4
        assert 1 <= P.value <= 20
    # |dx| < 1e-35 -> idirP = 0 -> bin 1
    d = np.float32([0.0, 0.6, 0.8])
    oracle_lib.oracle_escape_bins(C.byref(o.P), _fptr(d), C.byref(T), C.byref(P))
    assert P.value == 1 and T.value == 3     # acos(0.8)=0.6435 rad / (pi/10) -> bin 3


@pytest.mark.parametrize("sym", [True, False])
def test_cell_volume_matches_getVolume_formula(oracle_lib, sym):
    from mocassin_b200 import workloads as W

    m = W.hii_region() if sym else W.viewing_angles()
    o = orc.Oracle(m)
    g = m.grids[0]
    dV = g.cell_volumes(sym)
    for (x, y, z) in [(1, 1, 1), (2, 3, 4), (g.nx, g.ny, g.nz), (1, g.ny, 2), (g.nx - 1, 1, g.nz)]:
        v = oracle_lib.oracle_cell_volume(C.byref(o.P), C.byref(o.G[0]), x, y, z)
        a = g.active[x - 1, y - 1, z - 1]
        if a > 0:
            assert np.float32(v) == dV[a]
    # interior cell: dx*dy*dz with dx = |x(i+1)-x(i-1)|/2 / 1e15
    ax = g.xAxis
    v = oracle_lib.oracle_cell_volume(C.byref(o.P), C.byref(o.G[0]), 3, 3, 3)
    d = np.float32(np.abs(ax[3] - ax[1]) / np.float32(2) / np.float32(1e15))
    assert np.float32(v) == np.float32(np.float32(d * d) * d)
