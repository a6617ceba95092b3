"""scripts/run_dust_deck.py with the CUDA engine replaced by the oracle-backed stand-in of
tests/deck_runner.py: runs a shipped dust deck end to end on the CPU (how the
profiles/r01_dust_deck_*_oracle.* files were made).  Test infrastructure, not a measurement.

    python tests/tools/run_deck_on_oracle.py <deck dir> <share dir> [--out DIR] [--threads N]
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import deck_runner  # noqa: E402
from mocassin_b200 import api  # noqa: E402

threads = os.cpu_count() or 1
if "--threads" in sys.argv:
    i = sys.argv.index("--threads")
    threads = int(sys.argv[i + 1])
    del sys.argv[i:i + 2]


def _transport(self, iStar, n, deltaE=None):
    from oracle import oracle as O

    self.orc = O.Oracle(self.model, fp32_tallies=False)
    c = self.orc.transport_mt(iStar, 0, n, seed=self.seed, threads=threads)
    self.dE = float(deltaE)
    c.update(total_ms=0.0, kernel_ms=0.0)
    return c


deck_runner.OracleEngine.energyPacketDriver = _transport
api.PacketEngine = deck_runner.OracleEngine
sys.argv[0] = "run_dust_deck.py"
runpy.run_path(os.path.join(ROOT, "scripts", "run_dust_deck.py"), run_name="__main__")
