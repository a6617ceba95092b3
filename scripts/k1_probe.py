import sys, time, argparse, numpy as np
sys.path.insert(0,'.')
import bench
from mocassin_b200.api import PacketEngine
args = argparse.Namespace(grid=128, nbins=600, workload="clumpy")
m = bench.build_model(args, tables=False)
xsec, bands, den, dust = bench.compact_inputs(m)
e = PacketEngine(m, seed=12345); e.set_xsec(xsec)
e.set_option("trace", 1)
for rep in range(3):
    t=time.perf_counter(); e.assemble_opacity(1, bands, den, None, dust); print("K1 call", round(1e3*(time.perf_counter()-t),1), "ms")
    t=time.perf_counter(); e.set_dust_state(); print("set_dust_state", round(1e3*(time.perf_counter()-t),1), "ms")
print({k:(v.shape, v.dtype, v.flags['F_CONTIGUOUS']) for k,v in dust.items() if hasattr(v,'shape')}, den.shape, den.dtype, den.flags['F_CONTIGUOUS'])
