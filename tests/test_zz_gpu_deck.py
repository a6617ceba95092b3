"""The shipped dust benchmark deck p0tau1 (benchmarks/dust/1D, from tests/golden/deck_p0tau1.npz)
through the C ABI: three Lucy iterations on the device (K6 setDustPDF -> transport -> K5
getDustT, host rebuilds the dust opacities) must equal the same loop on the CPU oracle bit for
bit -- temperatures, convergence counts and packet counters."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name,packets", [("p0tau1", 100000), ("p0tau10", 60000)])
def test_deck_lucy_loop_on_device_matches_oracle(cuda_lib, oracle_lib, name, packets):
    from deck_runner import engine_step, oracle_step
    from mocassin_b200 import deck

    runs = {}
    for which in ("oracle", "cuda"):
        m, t, d = deck.deck_from_arrays(dict(np.load(os.path.join(GOLD, f"deck_{name}.npz"))))
        d.maxIterateMC, d.nPhotons = 3, packets
        m.deltaE[1] = np.float32(d.LStar) / np.float32(packets)
        step, st = (oracle_step(m, t, d, seed=500, threads=4) if which == "oracle" else engine_step(m, t, d, seed=500))
        hist = deck.iterate_dust(d, m, step)
        runs[which] = (hist, m.grids[0].Tdust.copy(), st["counters"])
        if which == "cuda":
            sed, cnt = st["eng"].fetch_sed()
            # dust only: every packet of the last iteration escapes (bar the odd one the iteration limits drop)
            assert int(cnt[:, 0].sum()) == st["counters"][-1]["nEscaped"] >= packets - 2
            st["eng"].close()
    (h0, T0, c0), (h1, T1, c1) = runs["oracle"], runs["cuda"]
    assert [(h["converged_pct"], h["nPhotons"]) for h in h0] == [(h["converged_pct"], h["nPhotons"]) for h in h1]
    for a, b in zip(c0, c1):
        for k in ("nAbs", "nSca", "nSegments", "nEscaped", "nDropped"):
            assert a[k] == b[k], k
    assert np.array_equal(T0[:, :, 1:].view(np.uint32), T1[:, :, 1:].view(np.uint32))


def test_disk_deck_iteration_on_device_matches_oracle(cuda_lib, oracle_lib):
    """benchmarks/dust/2D/tau1.000: 40^3 disk with two phi-free viewing angles.  One iteration:
    temperatures, escapedPackets (all three angle planes) and the device SED reduction equal the
    oracle's."""
    from deck_runner import engine_step, oracle_step
    from mocassin_b200 import deck

    packets = 80000
    out = {}
    for which in ("oracle", "cuda"):
        m, t, d = deck.deck_from_arrays(dict(np.load(os.path.join(GOLD, "deck_tau1.000.npz"))))
        d.maxIterateMC, d.nPhotons = 1, packets
        m.deltaE[1] = np.float32(d.LStar) / np.float32(packets)
        step, st = (oracle_step(m, t, d, seed=31, threads=4) if which == "oracle" else engine_step(m, t, d, seed=31))
        deck.iterate_dust(d, m, step)
        if which == "cuda":
            # the path-length quantum is capped by the widest cell on these axes (spacing ratio 10^5):
            # library and host model must pick the same one
            assert st["eng"].len_unit(1) == 2.0 ** m.len_unit_exponent(m.grids[0]) == 2.0 ** 17
            esc = st["eng"].fetch(1, want=("escapedPackets",))["escapedPackets"]
            _, cnt = st["eng"].fetch_sed()
            st["eng"].close()
        else:
            esc, cnt = st["escaped"], None
        out[which] = (m.grids[0].Tdust.copy(), esc, cnt, st["counters"][0])
    (T0, e0, _, c0), (T1, e1, cnt, c1) = out["oracle"], out["cuda"]
    for k in ("nAbs", "nSca", "nSegments", "nEscaped"):
        assert c0[k] == c1[k], k
    assert np.array_equal(T0[:, :, 1:].view(np.uint32), T1[:, :, 1:].view(np.uint32))
    assert np.array_equal(e0.view(np.uint32), e1.view(np.uint32))
    dE = np.float32(m.deltaE[1])
    want_cnt = np.rint(e0.astype(np.float64).sum(axis=0)[1:, :] / float(dE)).astype(np.int64)
    assert np.array_equal(cnt, want_cnt)


@pytest.mark.parametrize("name", ["viewing_angles", "multigrid_sym", "hii_sym_gas"])
def test_device_contcube_equals_frequency_sum_of_fetched_escaped_packets(cuda_lib, name):
    """K9 (contcube_kernel): per cell and viewing angle the float32 running sum over freq = 1..nbins
    of the folded escapedPackets -- the loop of writeContCube (output_mod.f90:2762-2772) -- equals
    the same loop done on the host over the array mcb200_fetch_estimators returns, bit for bit."""
    from cases import make
    from mocassin_b200.api import PacketEngine

    m, n = make(name)
    e = PacketEngine(m, seed=12345)
    e.upload_iteration_inputs()
    e.lucy_transport([n] * m.nStars)
    for iG in range(1, m.nGrids + 1):
        esc = e.fetch(iG, want=("escapedPackets",))["escapedPackets"]
        want = np.zeros((esc.shape[0], esc.shape[2]), np.float32)
        for f in range(1, m.nbins + 1):
            want = (want + esc[:, f, :]).astype(np.float32)
        got = e.fetch_contcube(iG)
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), iG
        assert np.count_nonzero(got) > 0
    e.close()


def test_deck_loop_with_device_side_dust_opacity_matches_oracle(cuda_lib, oracle_lib):
    """As the first test, but K1 rebuilds scaOpac/absOpac/opacity on the device from the
    device-resident dust state after every K5 (mcb200_assemble_opacity with Tdust = NULL): the whole
    iteration, sublimation test included, runs without Tdust or opacities crossing PCIe."""
    from deck_runner import engine_step, oracle_step
    from mocassin_b200 import deck

    packets = 60000
    runs = {}
    for which in ("oracle", "cuda"):
        m, t, d = deck.deck_from_arrays(dict(np.load(os.path.join(GOLD, "deck_p0tau1.npz"))))
        d.maxIterateMC, d.nPhotons = 3, packets
        m.deltaE[1] = np.float32(d.LStar) / np.float32(packets)
        step, st = (oracle_step(m, t, d, seed=900, threads=4) if which == "oracle"
                    else engine_step(m, t, d, seed=900, device_opacity=True))
        hist = deck.iterate_dust(d, m, step)
        g = m.grids[0]
        if which == "cuda":
            op, sca, ab = st["eng"].get_opacity(1, want_abs=True)
            st["eng"].close()
        else:
            op, sca, ab = g.opacity, g.scaOpac, g.absOpac
        runs[which] = (hist, g.Tdust.copy(), op, sca, ab)
    (h0, T0, op0, sca0, ab0), (h1, T1, op1, sca1, ab1) = runs["oracle"], runs["cuda"]
    assert [h["converged_pct"] for h in h0] == [h["converged_pct"] for h in h1]
    assert np.array_equal(T0[:, :, 1:].view(np.uint32), T1[:, :, 1:].view(np.uint32))
    for a, b in ((op0, op1), (sca0, sca1), (ab0, ab1)):
        assert np.array_equal(a[1:].view(np.uint32), b[1:].view(np.uint32))


def test_device_sublimation_mask_equals_host_mask(cuda_lib):
    """K1 with Tdust = NULL (mask built by dust_mask_kernel from the device dust state) against K1
    with the host's Tdust, on a 3-species, 2-component, 3-size model in which a third of the cells
    hold grains above their sublimation temperature."""
    from mocassin_b200 import workloads as W
    from mocassin_b200.api import PacketEngine

    model, t = W.dust_closure(n=9, nbins=60, nPhotons=1000, T0=100.0)
    g = model.grids[0]
    rng = np.random.default_rng(8)
    hot = rng.random(g.nCells + 1) < 0.33
    g.Tdust[1, 2, hot] = 1300.0          # species 1 of a component, size 2: above 1200 / 900 K, below 1400 K
    g.Tdust[2, 1, hot] = 2000.0
    e = PacketEngine(model)
    e.set_xsec(t["xSecArray"])
    e.set_dust_tables(t["widFlx"], t["grainWeight"], t["dustAbsXsecP"], t["dustEmIntegral"])
    e.set_opacity()
    e.set_dust_state()
    none = dict(species=[], off=[], low=[], high=[])
    den = np.zeros((g.nCells + 1, 0), np.float32)
    dust = dict(Ndust=g.Ndust, Tdust=g.Tdust, dustAbunIndex=g.dustAbunIndex, grainWeight=t["grainWeight"],
                dustScaXsecP=t["dustScaXsecP"], dustAbsXsecP=t["dustAbsXsecP"])
    e.assemble_opacity(1, none, den, None, dust)
    host = e.get_opacity(1, want_abs=True)
    e.assemble_opacity(1, none, den, None, dict(dust, Tdust=None))
    dev = e.get_opacity(1, want_abs=True)
    for a, b in zip(host, dev):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    cold = dict(dust, Tdust=np.full_like(g.Tdust, 50.0))
    e.assemble_opacity(1, none, den, None, cold)
    assert not np.array_equal(e.get_opacity(1)[0], host[0])      # the mask does matter in this model
    e.close()


@pytest.mark.parametrize("name", ["hii_sym_gas", "cube_clumpy_gasdust"])
def test_device_opacity_rows_feed_tau_nu(cuda_lib, name):
    """mcb200_get_opacity_rows: the gathered rows equal the rows of the full table, and tauNu.out's
    three optical-depth columns computed from them equal those computed from the host table."""
    from cases import make
    from mocassin_b200 import output
    from mocassin_b200.api import MocassinError, PacketEngine

    m, _ = make(name)
    g = m.grids[0]
    e = PacketEngine(m, seed=1)
    e.set_opacity()
    full = e.get_opacity(1)[0]
    cells = np.array([0, 1, g.nCells, 5, 5, g.nCells // 2], np.int32)
    got = e.get_opacity_rows(1, cells)
    assert np.array_equal(got.view(np.uint32), full[cells, :].view(np.uint32))
    dev = output.tau_nu(m, lambda c: e.get_opacity_rows(1, c))
    host = output.tau_nu(m, lambda c: g.opacity[c, :])
    for a, b in zip(dev, host):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and a.max() > 0
    with pytest.raises(MocassinError):
        e.get_opacity_rows(1, np.array([g.nCells + 1], np.int32))
    e.close()


def test_c_host_example_runs_on_the_device(cuda_lib, tmp_path):
    """examples/host_example.c: set_config -> set_grid -> set_spectra -> set_stars -> set_opacity ->
    set_pdfs -> transport -> fetch_estimators from plain C; its own checks (escaped energy = L, mean
    path per packet = 1.33 half edges as on the CPU oracle) decide the exit code."""
    import subprocess

    from mocassin_b200 import _lib

    exe = tmp_path / "host_example"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "host_example.c"),
                        _lib.LIB_PATH, "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-lm", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


def test_converged_deck_run_writes_the_files_the_oracle_run_wrote(cuda_lib, tmp_path):
    """benchmarks/dust/1D/p0tau10 to convergence on the device through scripts/run_dust_deck.py (10 Lucy
    iterations, autoPackets 1e5 -> 6.4e6 packets): output/SED.out, summary.out and tauNu.out are BYTE
    FOR BYTE the files the same script wrote with the CPU oracle standing in for the engine
    (profiles/r01_dust_deck_p0tau10_oracle_*.out) -- every packet history, temperature and
    convergence decision of all ten iterations included."""
    import subprocess

    out = tmp_path / "output"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_dust_deck.py"), "--golden",
                          os.path.join(GOLD, "deck_p0tau10.npz"), "--out", str(out)], capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    import json

    rows = [json.loads(x) for x in res.stdout.strip().splitlines()]
    want = [json.loads(x) for x in open(os.path.join(ROOT, "profiles", "r01_dust_deck_p0tau10_oracle.jsonl")) if x.startswith("{")]
    assert len(rows) == len(want) == 11
    for a, b in zip(rows[:-1], want[:-1]):
        for k in ("iteration", "converged_pct", "nPhotons", "segments", "nAbs", "nSca"):
            assert a[k] == b[k], (a["iteration"], k)
    assert rows[-1]["Tdust_along_x"] == want[-1]["Tdust_along_x"] and rows[-1]["escaped_packets"] == want[-1]["escaped_packets"]
    for fn in ("SED", "summary", "tauNu"):
        got = (out / f"{fn}.out").read_bytes()
        ref = open(os.path.join(ROOT, "profiles", f"r01_dust_deck_p0tau10_oracle_{fn}.out"), "rb").read()
        assert got == ref, fn
