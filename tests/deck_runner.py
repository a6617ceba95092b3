"""One Lucy iteration of a dust-only deck (setDustPDF -> energyPacketDriver -> host scaling ->
getDustT -> dust opacities) on the CPU oracle and on the CUDA engine, as `step` callbacks for
mocassin_b200.deck.iterate_dust.  Test infrastructure: the product never imports this."""

import numpy as np

from mocassin_b200 import deck
from mocassin_b200.api import PacketEngine, scale_estimators
from mocassin_b200.model import F32


def oracle_step(model, tables, d, seed=12345, threads=1):
    from oracle import oracle as O

    g = model.grids[0]
    state = dict(it=0, counters=[])

    def step(nPhotons, deltaE):
        state["it"] += 1
        model.deltaE[1] = F32(deltaE)
        g.dustPDF = O.dust_pdf(model, g, tables)
        orc = O.Oracle(model, fp32_tallies=False)
        # a fresh packet stream per iteration (the reference reseeds from the clock, photon_mod.f90:68-87)
        c = orc.transport_mt(1, 0, nPhotons, seed=seed + state["it"], threads=threads)
        f = orc.folded(1, float(model.deltaE[1]))
        Js, _ = scale_estimators(model, f["Jste"], np.zeros((1, 1, 1), F32))
        # iterateMC zeroes grid%lgConverged before every iteration (iteration_mod.f90:85)
        T, conv = O.dust_update(model, g, tables, Js, d.XHILimit)
        g.Tdust = T
        deck.dust_opacity(g, tables)
        state["counters"].append(c)
        state["escaped"] = f["escapedPackets"]
        return int(conv[1:].sum()), g.nCells

    return step, state


def engine_step(model, tables, d, seed=12345, device_opacity=False):
    """The same iteration through the C ABI.  Tdust and the convergence flags stay on the device
    between K5 and K6.  device_opacity = False: the host rebuilds the dust opacities from the
    fetched temperatures and uploads them; True: K1 rebuilds them on the device from the
    device-resident dust state (mcb200_assemble_opacity with Tdust = NULL) and nothing is
    uploaded between iterations."""
    g = model.grids[0]
    eng = PacketEngine(model, seed=seed)
    eng.set_xsec(tables["xSecArray"])
    eng.set_dust_tables(tables["widFlx"], tables["grainWeight"], tables["dustAbsXsecP"], tables["dustEmIntegral"])
    eng.set_opacity()
    eng.set_dust_state()
    state = dict(it=0, counters=[], eng=eng)

    def step(nPhotons, deltaE):
        state["it"] += 1
        model.deltaE[1] = F32(deltaE)
        eng.set_option("seed", seed + state["it"])
        eng.setDustPDF(1)
        eng.zero_estimators()
        c = eng.energyPacketDriver(1, nPhotons, deltaE=float(F32(deltaE)))
        eng.reduce()
        T, conv, nconv = eng.getDustT(1, d.XHILimit)
        g.Tdust = T
        if device_opacity:
            eng.assemble_opacity(1, dict(species=[], off=[], low=[], high=[]), np.zeros((g.nCells + 1, 0), F32), None,
                                 dict(Ndust=g.Ndust, Tdust=None, dustAbunIndex=g.dustAbunIndex,
                                      grainWeight=tables["grainWeight"], dustScaXsecP=tables["dustScaXsecP"],
                                      dustAbsXsecP=tables["dustAbsXsecP"]))
        else:
            deck.dust_opacity(g, tables)
            eng.set_opacity()
            eng.set_dust_state()
        state["counters"].append(c)
        return int(nconv), g.nCells

    return step, state


class OracleEngine:
    """Stand-in for mocassin_b200.api.PacketEngine backed by the CPU oracle, with just the calls
    scripts/run_dust_deck.py makes -- lets the script's own logic (iteration loop, output files)
    be exercised without a GPU.  Test infrastructure only."""

    def __init__(self, model, seed=12345, **kw):
        self.model, self.seed, self.tables = model, seed, {}
        self.sed_cnt = None

    def set_xsec(self, x):
        self.tables["xSecArray"] = x

    def set_dust_tables(self, widFlx, grainWeight, dustAbsXsecP, dustEmIntegral):
        self.tables.update(widFlx=widFlx, grainWeight=grainWeight, dustAbsXsecP=dustAbsXsecP, dustEmIntegral=dustEmIntegral)

    def set_opacity(self, iG=0): pass
    def set_dust_state(self, iG=0): pass
    def close(self): pass

    def set_option(self, name, value):
        if name == "seed":
            self.seed = int(value)

    def setDustPDF(self, iG, fetch=False):
        from oracle import oracle as O

        g = self.model.grids[0]
        g.dustPDF = O.dust_pdf(self.model, g, self.tables)

    def zero_estimators(self):
        self.sed_cnt = None

    def energyPacketDriver(self, iStar, n, deltaE=None):
        from oracle import oracle as O

        self.orc = O.Oracle(self.model, fp32_tallies=False)
        c = self.orc.transport_mt(iStar, 0, n, seed=self.seed, threads=2)
        self.dE = float(deltaE)
        c.update(total_ms=0.0, kernel_ms=0.0)
        return c

    def reduce(self):
        self.f = self.orc.folded(1, self.dE)
        cnt, raw = self.orc.sed(self.dE)
        self.sed_cnt = (raw, cnt)

    def getDustT(self, iG, XHILimit):
        from oracle import oracle as O

        g = self.model.grids[0]
        Js, _ = scale_estimators(self.model, self.f["Jste"], np.zeros((1, 1, 1), F32))
        T, conv = O.dust_update(self.model, g, self.tables, Js, XHILimit)
        g.Tdust = T
        return T, conv, int(conv[1:].sum())

    def fetch_sed(self):
        return self.sed_cnt

    def assemble_opacity(self, iG, bands, den, ff1=None, dust=None):
        # dust-only, Tdust = None: what K1 does from the device-resident dust state
        assert dust is not None and dust.get("Tdust") is None and len(bands["species"]) == 0
        t = dict(self.tables, dustScaXsecP=dust["dustScaXsecP"], grainAbun1=self.model.grainAbun[0, :],
                 TdustSublime=self.model.TdustSublime)
        deck.dust_opacity(self.model.grids[iG - 1], t)

    def get_opacity_rows(self, iG, cells):
        return self.model.grids[iG - 1].opacity[np.asarray(cells), :]
