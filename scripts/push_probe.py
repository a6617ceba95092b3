"""Device time of the merge's push kernels with local "peer" buffers (one GPU): structure, not link speed."""
import ctypes as C, sys
sys.path.insert(0, ".")
from mocassin_b200 import workloads as W
from mocassin_b200.api import PacketEngine
e = PacketEngine(W.hii_region())
for world, n in ((2, 1_180_000_000), (8, 1_180_000_000)):
    for mode in (0, 1):
        ms = C.c_double()
        e._check(e.lib.mcb200_test_push_kernels(e.h, mode, world, n, C.byref(ms)))
        cnt = n // world
        print(f"world {world} mode {'packed' if mode else '64-bit'}: {ms.value:.3f} ms for {(world-1)*cnt:.3e} elements pushed "
              f"({(world-1)*cnt*(4 if mode else 8)/ms.value/1e6:.0f} GB/s payload)")
