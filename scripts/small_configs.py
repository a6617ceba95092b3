"""Throughput on the reference's own benchmark-sized configurations (BASELINE.json configs[0..3]) from their
fixtures (tests/golden/deck_*.npz, built from the shipped decks and the reference's atomic / optical data):
benchmarks/gas/HII40 and PN150 on their real ~100-150-band opacity tables (first-iteration state, and a partly
recombined state that exercises re-emission), examples/multigridgas and multigridgasdust (16^3 + 11^3 sub-grid),
the dust decks benchmarks/dust/1D/p0tau{1,10,100}, 2D/tau1.000 after two Lucy iterations have set a realistic
dustPDF; and, for continuity with round 1, the synthetic stand-ins of mocassin_b200/workloads.py.  One JSON line each."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
from mocassin_b200 import deck, workloads as W
from mocassin_b200.api import PacketEngine

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000


def report(name, e, n):
    for rep in range(3):
        e.zero_estimators()
        c = e.energyPacketDriver(1, n)
        e.reduce()
    print(json.dumps(dict(config=name, packets=n, packets_per_s=n / (c["total_ms"] * 1e-3), ms=c["total_ms"],
                          segments_per_packet=c["nSegments"] / n, segments_per_s=c["nSegments"] / (c["total_ms"] * 1e-3),
                          waves=c["nWaves"], launches=c["nLaunches"])), flush=True)


sys.path.insert(0, "tests")
from mocassin_b200 import gasdeck, multideck

for fx in ("HII40", "PN150"):
    for state in ("initial", "recombined"):
        m, t, _ = gasdeck.gas_deck_from_arrays(dict(np.load(os.path.join("tests", "golden", f"deck_{fx}.npz"))))
        ion = t["ionDen"]
        if state == "recombined":
            import test_zx_gpu_gas_deck as Tg
            ion = Tg._recombined(m, t)
        den = t["xsec"].species_densities(ion, t["elemAbun"], t["abIndex"], m.grids[0].Hden)
        e = PacketEngine(m, seed=12345)
        e.set_xsec(t["xsec"].xSecArray)
        t0 = time.perf_counter(); e.assemble_opacity(1, t["bands"], den, None); k1 = (time.perf_counter() - t0) * 1e3
        e.set_pdfs()
        report(f"deck {fx} 13^3 x{m.nbins} gas, {t['bands']['species'].shape[0]} bands, {state} state (K1 {k1:.2f} ms)", e, N)
        e.close()

for fx in ("multigridgas", "multigridgasdust"):
    m, t, _ = multideck.multideck_from_arrays(dict(np.load(os.path.join("tests", "golden", f"deck_{fx}.npz"))))
    e = PacketEngine(m, seed=12345)
    e.set_xsec(t["xsec"].xSecArray)
    for iG, ent in enumerate(t["grids"], start=1):
        e.assemble_opacity(iG, t["bands"], ent["den"], None, ent["dust"])
    e.set_pdfs()
    e.set_dust_state()
    report(f"deck {fx} 16^3 + 11^3 sub-grid x{m.nbins}", e, N)
    e.close()

cases = [("HII40-like 13^3 x600 gas", W.hii_region, dict(nPhotons=N)),
         ("dust shell 16^3 x215 tauV=10", W.dust_shell, dict(tauV=10.0, nPhotons=N)),
         ("multigrid 16^3+11^3 x600 gas+dust", W.multigrid, dict(nPhotons=N))]
for name, fn, kw in cases:
    m = fn(**kw)
    e = PacketEngine(m, seed=12345)
    e.upload_iteration_inputs()
    report(name, e, N)
    e.close()

for fx in ("p0tau1", "p0tau10", "p0tau100", "tau1.000"):
    path = os.path.join("tests", "golden", f"deck_{fx}.npz")
    if not os.path.exists(path):
        continue
    m, t, d = deck.deck_from_arrays(dict(np.load(path)))
    g = m.grids[0]
    e = PacketEngine(m, seed=12345)
    e.set_xsec(t["xSecArray"])
    e.set_dust_tables(t["widFlx"], t["grainWeight"], t["dustAbsXsecP"], t["dustEmIntegral"])
    e.set_opacity()
    e.set_dust_state()
    for it in range(2):                      # warm the dust temperatures: 50 K everywhere is not a workload
        e.setDustPDF(1)
        e.zero_estimators()
        e.energyPacketDriver(1, min(N, 1_000_000), deltaE=float(np.float32(d.LStar) / np.float32(min(N, 1_000_000))))
        e.reduce()
        e.getDustT(1, d.XHILimit)
    t0 = time.perf_counter(); e.setDustPDF(1); k6 = (time.perf_counter() - t0) * 1e3
    report(f"deck {fx} {g.nx}^3 x{m.nbins} dust ({g.nCells} cells; K6 {k6:.2f} ms)", e, N)
    e.close()
