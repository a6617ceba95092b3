"""Debug helper: K6 (setDustPDF on the device) against the oracle on random dust temperatures; prints the first
differences.  Uses the oracle, so it lives under tests/ (run from the repo root on a GPU box)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from mocassin_b200 import workloads as W
from mocassin_b200.api import PacketEngine
from oracle import oracle as O
F32 = np.float32
model, t = W.dust_closure(n=9, nbins=120, multiChem=True)
g = model.grids[0]
rng = np.random.default_rng(5)
g.Tdust[1:, 1:, 1:] = rng.uniform(5.0, 1500.0, size=g.Tdust[1:, 1:, 1:].shape).astype(F32)
g.Tdust[1, 1, 1:4] = F32(0.0)
want = O.dust_pdf(model, g, t)
eng = PacketEngine(model)
eng.set_xsec(t["xSecArray"]); eng.set_dust_tables(t["widFlx"], t["grainWeight"], t["dustAbsXsecP"], t["dustEmIntegral"])
eng.set_opacity(); eng.set_dust_state()
got = eng.setDustPDF(1, fetch=True)
d = got.view(np.uint32) != want.view(np.uint32)
print("ndiff", d.sum(), "of", d.size)
i, j = np.nonzero(d)
for k in range(min(12, len(i))):
    print(i[k], j[k], got[i[k], j[k]], want[i[k], j[k]], hex(got.view(np.uint32)[i[k], j[k]]), hex(want.view(np.uint32)[i[k], j[k]]))
nanrows = np.isnan(want).any(axis=1)
print("nan rows", nanrows.sum(), "diff rows", np.unique(i).size, "diff in nan rows", np.isin(np.unique(i), np.nonzero(nanrows)[0]).sum())
print("first diff col per row", [int(j[i == r].min()) for r in np.unique(i)[:10]])
