"""Throughput on the reference's own benchmark-sized configurations (BASELINE.json configs[0]):
HII40-like 13^3 x 600 gas, dust 1-D shell 16^3 x 215, multigrid 16^3 + 11^3.  One JSON line each."""
import json, sys, time
sys.path.insert(0, ".")
from mocassin_b200 import workloads as W
from mocassin_b200.api import PacketEngine

cases = [("HII40-like 13^3 x600 gas", W.hii_region, dict(nPhotons=10_000_000)),
         ("dust shell 16^3 x215 tauV=10", W.dust_shell, dict(tauV=10.0, nPhotons=10_000_000)),
         ("multigrid 16^3+11^3 x600 gas+dust", W.multigrid, dict(nPhotons=10_000_000))]
for name, fn, kw in cases:
    m = fn(**kw)
    e = PacketEngine(m, seed=12345)
    e.upload_iteration_inputs()
    n = 10_000_000
    for rep in range(3):
        e.zero_estimators()
        t = time.perf_counter(); c = e.energyPacketDriver(1, n); dt = time.perf_counter() - t
    print(json.dumps(dict(config=name, packets=n, packets_per_s=n / (c["total_ms"] * 1e-3), ms=c["total_ms"],
                          segments_per_packet=c["nSegments"] / n, segments_per_s=c["nSegments"] / (c["total_ms"] * 1e-3),
                          waves=c["nWaves"], launches=c["nLaunches"])))
    e.close()
