"""Device-resident dust-only Lucy iteration: transport -> getDustT ->
setDustPDF with nothing but counters leaving the GPU.  Prints one JSON line per iteration and
a summary with the K5/K6 kernel costs.  Usage: python scripts/dust_lucy.py [n] [nbins] [packets] [iters]"""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mocassin_b200 import workloads as W          # noqa: E402
from mocassin_b200.api import PacketEngine        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nbins = int(sys.argv[2]) if len(sys.argv) > 2 else 300
packets = int(float(sys.argv[3])) if len(sys.argv) > 3 else 10_000_000
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 6

t0 = time.perf_counter()
model, t = W.dust_closure(n=n, nbins=nbins, nPhotons=packets, T0=100.0)
g = model.grids[0]
eng = PacketEngine(model)
eng.set_xsec(t["xSecArray"])
eng.set_dust_tables(t["widFlx"], t["grainWeight"], t["dustAbsXsecP"], t["dustEmIntegral"])
eng.set_opacity()
eng.set_dust_state()
setup = time.perf_counter() - t0
nconv = C.c_int64(0)
rows = []
for it in range(iters):
    a = time.perf_counter()
    eng._check(eng.lib.mcb200_dust_pdf(eng.h, 1, None))
    b = time.perf_counter()
    eng.zero_estimators()
    c = eng.energyPacketDriver(1, packets)
    eng.reduce()
    d = time.perf_counter()
    eng._check(eng.lib.mcb200_dust_update(eng.h, 1, 0.05, None, None, C.byref(nconv)))
    e = time.perf_counter()
    row = dict(iteration=it + 1, dust_pdf_ms=(b - a) * 1e3, transport_ms=(d - b) * 1e3, dust_update_ms=(e - d) * 1e3,
               converged_pct=100.0 * nconv.value / g.nCells, packets_per_s=packets / (d - b),
               segments=c["nSegments"], nAbs=c["nAbs"], nSca=c["nSca"])
    rows.append(row)
    print(json.dumps(row), flush=True)
T, conv, nc = eng.getDustT(1, 0.05)     # one more update just to fetch the temperatures
sed, cnt = eng.fetch_sed()
pairs = int(model.nSpeciesPart.max()) * model.nSizes
k5 = float(np.median([r["dust_update_ms"] for r in rows]))
k6 = float(np.median([r["dust_pdf_ms"] for r in rows]))
summary = dict(workload=f"dust_closure n={n} nbins={nbins} packets={packets}", nCells=g.nCells, setup_s=setup,
               K5_dust_update_ms=k5, K5_GBps=(g.nCells * (nbins * 4.0 * -(-pairs // 8) + 4.0 * T[:, :, 0].size)) / (k5 * 1e-3) / 1e9,
               K6_dust_pdf_ms=k6, K6_planck_evals_per_s=g.nCells * nbins * pairs / (k6 * 1e-3),
               Tdust_mean_min_max=[float(T[0, 0, 1:].mean()), float(T[0, 0, 1:].min()), float(T[0, 0, 1:].max())],
               escaped_fraction=float(cnt[:, 0].sum()) / packets)
print(json.dumps(summary))
