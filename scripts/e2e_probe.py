import sys, time, argparse, numpy as np, torch
sys.path.insert(0,'.')
import bench
from mocassin_b200.api import PacketEngine
args = argparse.Namespace(grid=128, nbins=600, workload="clumpy")
m = bench.build_model(args, tables=False)
xsec, bands, den, dust = bench.compact_inputs(m)
table, idx, tl = bench.rec_tables(m, np.random.default_rng(2025))
g = m.grids[0]; nR, nb = g.nCells+1, m.nbins
def pinned(shape):
    t = torch.empty(int(np.prod(shape)), dtype=torch.float32, pin_memory=True)
    return t, t.numpy().reshape(shape, order="F")
t_rec, rec = pinned((nR, nb)); np.take(table.T, idx, axis=1, out=rec.T); rec[0,:]=0
g.recPDF, g.totalLines = rec, tl
t_J, Jh = pinned((nR, nb)); t_E, Eh = pinned((nR, nb+1, 1))
e = PacketEngine(m, seed=12345); e.set_xsec(xsec)
e.assemble_opacity(1, bands, den, None, dust); e.set_pdfs(); e.set_dust_state(); e.zero_estimators()
e.energyPacketDriver(1, 125000000)
for mode in (0, 1):
    e.set_option("async_pdfs", mode)
    for rep in range(2):
        T = {}
        def tick(name, f):
            t=time.perf_counter(); r=f(); T[name]=round(1e3*(time.perf_counter()-t),1); return r
        t0=time.perf_counter()
        tick('K1', lambda: e.assemble_opacity(1, bands, den, None, dust))
        tick('dust_state', e.set_dust_state)
        tick('set_pdfs', e.set_pdfs)
        tick('zero', e.zero_estimators)
        c = tick('transport', lambda: e.energyPacketDriver(1, 125000000))
        tick('fetch', lambda: e.fetch(1, out={"Jste": Jh, "escapedPackets": Eh}))
        print(mode, round(1e3*(time.perf_counter()-t0),1), T, c['kernel_ms'], c['total_ms'])
