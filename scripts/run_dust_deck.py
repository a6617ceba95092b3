"""Run one of the reference's dust-only decks (e.g. benchmarks/dust/1D/p0tau1) through the CUDA
library: the deck's own input.in, density, grain and n,k files -> Lucy iterations on the device
(setDustPDF -> energyPacketDriver -> getDustT) with iterateMC's convergence and autoPackets
rules -> output/SED.out, output/tauNu.out, output/summary.out, output/dustGrid.out, output/grid0.out,
output/photoSource.out in the reference's layouts.

    python scripts/run_dust_deck.py <deck dir> <share dir with dustData/> [--out DIR] [--max-iter N]
    python scripts/run_dust_deck.py --golden tests/golden/deck_p0tau1.npz      (no reference tree needed)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocassin_b200 import checkpoint, deck, output  # noqa: E402
from mocassin_b200.api import PacketEngine          # noqa: E402
from mocassin_b200.model import F32                 # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("deck_dir", nargs="?")
    ap.add_argument("share_dir", nargs="?")
    ap.add_argument("--golden")
    ap.add_argument("--out", default="output")
    ap.add_argument("--max-iter", type=int, default=0)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--device-opacity", action="store_true",
                    help="rebuild the dust opacities on the device (K1 from the device dust state) instead of on the host")
    a = ap.parse_args()
    if a.golden:
        model, tables, d = deck.deck_from_arrays(dict(np.load(a.golden)))
    else:
        model, tables, d = deck.load_dust_deck(a.deck_dir, a.share_dir)
    if a.max_iter:
        d.maxIterateMC = a.max_iter
    g = model.grids[0]
    os.makedirs(a.out, exist_ok=True)
    eng = PacketEngine(model, seed=a.seed)
    eng.set_xsec(tables["xSecArray"])
    eng.set_dust_tables(tables["widFlx"], tables["grainWeight"], tables["dustAbsXsecP"], tables["dustEmIntegral"])
    eng.set_opacity()
    eng.set_dust_state()
    summary = open(os.path.join(a.out, "summary.out"), "w")
    state = dict(it=0, conv=None, nPhotons=0)
    t0 = time.perf_counter()

    def step(nPhotons, deltaE):
        state["it"] += 1
        model.deltaE[1] = F32(deltaE)
        eng.set_option("seed", a.seed + state["it"])       # the reference reseeds every call (photon_mod.f90:68-87)
        eng.setDustPDF(1)
        eng.zero_estimators()
        c = eng.energyPacketDriver(1, nPhotons, deltaE=float(F32(deltaE)))
        eng.reduce()
        T, conv, nconv = eng.getDustT(1, d.XHILimit)
        if a.device_opacity:                               # iteration_mod.f90:166-227 on the device
            eng.assemble_opacity(1, dict(species=[], off=[], low=[], high=[]), np.zeros((g.nCells + 1, 0), F32), None,
                                 dict(Ndust=g.Ndust, Tdust=None, dustAbunIndex=g.dustAbunIndex,
                                      grainWeight=tables["grainWeight"], dustScaXsecP=tables["dustScaXsecP"],
                                      dustAbsXsecP=tables["dustAbsXsecP"]))
        else:                                              # ... or on the host (sublimed grains drop out)
            deck.dust_opacity(g, tables)
            eng.set_opacity()
            eng.set_dust_state()
        state.update(conv=conv, nPhotons=nPhotons, counters=c)
        return int(nconv), g.nCells

    def log(row):
        it, pct = row["iteration"], row["converged_pct"]
        summary.write(f" ! iterateMC: [Summary] Iteration  {it} ;  {int(pct)} % converged cells in grid  1\n")
        summary.write(f" ! iterateMC: [Summary] Iteration  {it} ; Total:   {pct:.5f} % converged cells over all grids\n")
        summary.write(f" ! iterateMC: [Summary] Iteration  {it} ;  {row['nPhotons']}  energy packets used\n\n")
        summary.flush()
        c = state["counters"]
        print(json.dumps(dict(row, seconds=round(time.perf_counter() - t0, 3), transport_ms=c["total_ms"],
                              segments=c["nSegments"], nAbs=c["nAbs"], nSca=c["nSca"])), flush=True)

    hist = deck.iterate_dust(d, model, step, log=log)
    summary.close()
    raw, cnt = eng.fetch_sed()
    totalE = output.write_sed(os.path.join(a.out, "SED.out"), model, tables["widFlx"], raw)
    output.write_tau_nu(os.path.join(a.out, "tauNu.out"), model, lambda cells: eng.get_opacity_rows(1, cells))
    checkpoint.write_checkpoint(a.out, model, lgConverged=[state["conv"]])
    checkpoint.write_photo_source(os.path.join(a.out, "photoSource.out"), model, ["blackbody"], [d.TStellar], [d.LStar],
                                  [state["nPhotons"]])
    cells = [int(c) for c in g.active[:, 0, 0] if c > 0]
    print(json.dumps(dict(iterations=len(hist), converged_pct=hist[-1]["converged_pct"], total_energy_out_e36=totalE,
                          LStar=d.LStar, escaped_packets=int(cnt[:, 0].sum()), packets=state["nPhotons"],
                          Tdust_along_x=[round(float(g.Tdust[0, 0, c]), 2) for c in cells], out=a.out)))
    eng.close()


if __name__ == "__main__":
    main()
